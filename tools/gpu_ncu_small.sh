#!/bin/bash
# ncu --set full for the small kernels of the L2 step (K1 pack, K3 finish, K5 filter).  Usage: bash tools/gpu_ncu_small.sh <tag>
TAG=${1:-s}; O=gpurun_out; mkdir -p $O
SHORT="python bench.py --steps 10 --warmup 3 --no-ramp --no-cpu --no-ransac --no-hamming --no-cfg5"
$SHORT > $O/plain_$TAG.log 2>&1 || exit 1
for K in l2_pack_kernel l2_finish_kernel compact_lookback_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 -o $O/prof_${K}_$TAG -f $SHORT > $O/ncu_${K}_$TAG.log 2>&1
  echo "ncu $K exit $?"
done
