#!/bin/bash
# tile / blocks-per-SM sweep of the dense-matrix kernel
cd /root/repo
for cfg in "8 0 262144" "10 0 65536" "9 0 131072" "7 0 1048576" "11 0 16384"; do
  for tb in "0 0" "16 2" "32 1" "8 3" "8 2" "4 2" "4 3" "2 2" "2 3" "16 1"; do
    set -- $tb
    echo "== $cfg  T=$1 blocks=$2"
    GAAST_DM_TILE=$1 GAAST_DM_BLOCKS=$2 timeout 120 python exp/matrix/one.py $cfg 2>&1 | tail -2 | cut -c1-230
  done
done
