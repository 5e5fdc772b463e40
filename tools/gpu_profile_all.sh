#!/bin/bash
# One gpurun call: parity tests, a full bench line, the ncu launch list of the bench command,
# and one `ncu --set full` capture of each dominant kernel (K2 L2 GEMM, K4 Hamming, K7 scoring).
# Usage (from the repo root, on the GPU box): bash tools/gpu_profile_all.sh <tag>
set -u
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $O/pytest_$TAG.log
tail -3 $O/pytest_$TAG.log
python bench.py --steps 1000 --warmup 10 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"
SHORT="python bench.py --steps 10 --warmup 3 --no-ramp --no-cpu --no-cfg5 --ransac-steps 2"
$SHORT > $O/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_$TAG.csv $SHORT > $O/ncu_list_$TAG.log 2>&1
echo "launch list exit $?"
python tools/step_timeline.py > $O/timeline_$TAG.txt 2>&1
for K in l2_tc_kernel ham_knn2_kernel ransac_score_kernel l2_pack_kernel l2_finish_kernel compact_lookback_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -o $O/prof_${K}_$TAG -f $SHORT > $O/ncu_${K}_$TAG.log 2>&1
  echo "ncu $K exit $?"
done
python tools/ham_only.py > $O/plain_ham_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:l2_tc_kernel -s 2 -c 1 -o $O/prof_l2_tc_kernel_fp8_$TAG -f python tools/ham_only.py > $O/ncu_l2_tc_kernel_fp8_$TAG.log 2>&1
echo "ncu l2_tc_kernel<FP8> exit $?"
PM_PAIRS=6 python tools/pair_profile.py > $O/plain_pair_$TAG.log 2>&1 &&
PM_PAIRS=6 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_pair_$TAG.csv python tools/pair_profile.py > $O/ncu_pair_$TAG.log 2>&1
echo "pair launch list exit $?"
PM_PAIRS=64 python tools/pair_profile.py | tee $O/pair_$TAG.log
PM_PIPE=1 python tools/step_timeline.py > $O/timeline_pipe_$TAG.txt 2>&1
