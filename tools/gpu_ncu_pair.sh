#!/bin/bash
# ncu --set full of the grouped RANSAC kernels of the pair pipeline (one group of 16 pairs).  Usage: bash tools/gpu_ncu_pair.sh <tag>
TAG=${1:-pg}; O=gpurun_out; mkdir -p $O
export PM_LANES=1 PM_PAIRS=16
timeout 120 python tools/pair_profile.py > $O/plain_pairg_$TAG.log 2>&1 || exit 1
for K in ransac_solve_kernel ransac_project_kernel ransac_score_kernel refit_solve_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o $O/prof_pairg_${K}_$TAG -f python tools/pair_profile.py > $O/ncu_pairg_${K}_$TAG.log 2>&1
  echo "ncu $K exit $?"
done
