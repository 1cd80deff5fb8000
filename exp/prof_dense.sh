cd $GRAFT_REPO_ROOT
# ncu --set full of the dense engine's matrix kernel as shipped (n = 8, 10), each after its plain run exited 0
for spec in "8 262144" "10 65536"; do
  set -- $spec
  python exp/matrix/one.py $1 0 $2 > gpurun_out/plain_dm$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:gaast_dense_matrix -s 3 -c 1 -f -o gpurun_out/prof_r2f_dm$1 \
      python exp/matrix/one.py $1 0 $2 > gpurun_out/ncu_dm$1.log 2>&1
  tail -2 gpurun_out/plain_dm$1.log
done
