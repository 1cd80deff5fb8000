cd $GRAFT_REPO_ROOT
prof() { # name, args...
  name=$1; shift
  CMD="python bench.py --only --no-e2e --no-cpu --steps 2 --warmup 1 $*"
  GAAST_BENCH_NO_GRAPH=1 $CMD > gpurun_out/plain_$name.log 2>&1 && GAAST_BENCH_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:gaast_eval -s 3 -c 1 -f -o gpurun_out/prof_r2f_$name $CMD > gpurun_out/ncu_$name.log 2>&1
}
prof cfg1 --workload cfg1
prof cfg2 --workload cfg2
prof cfg3 --workload cfg3
prof cfg4 --workload cfg4
prof cfg5sum --workload cfg5
prof cfg5 --workload cfg5 --no-sum
D="python bench.py --steps 2 --warmup 1 --no-cpu"
$D > gpurun_out/plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv $D > gpurun_out/ncu_default.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
