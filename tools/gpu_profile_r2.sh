#!/bin/bash
# round 2 profiling session (one gpurun call): a full bench line, the ncu launch list of a short bench command (it first
# exits 0 without ncu), one `ncu --set full` capture per dominant kernel, the split-mode and pair-pipeline launch lists.
set -u
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
PM_BENCH_WATCHDOG_S=500 timeout 600 python bench.py --steps 1000 --warmup 10 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"
SHORT="python bench.py --steps 10 --warmup 3 --no-ramp --no-cpu --no-cfg5 --ransac-steps 2"
timeout 300 $SHORT > $O/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_$TAG.csv $SHORT > $O/ncu_list_$TAG.log 2>&1
echo "launch list exit $?"
for K in l2_tc_kernel l2_finish_kernel compact_lookback_kernel l2_pack_kernel ransac_score_kernel ham_knn2_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 -o $O/prof_${K}_$TAG -f $SHORT > $O/ncu_${K}_$TAG.log 2>&1
  echo "ncu $K exit $?"
done
timeout 120 python tools/ham_only.py > $O/plain_ham_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:l2_tc_kernel -s 2 -c 1 -o $O/prof_l2_tc_kernel_fp8_$TAG -f python tools/ham_only.py > $O/ncu_l2_tc_kernel_fp8_$TAG.log 2>&1
echo "ncu l2_tc_kernel<FP8> exit $?"
PM_SURF=1 timeout 120 python tools/split_mode_bench.py > $O/plain_split_$TAG.log 2>&1 &&
PM_SURF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:l2_tc_kernel -s 4 -c 1 -o $O/prof_l2_tc_kernel_split_$TAG -f python tools/split_mode_bench.py > $O/ncu_l2_tc_kernel_split_$TAG.log 2>&1
echo "ncu l2_tc_kernel split exit $?"
PM_LANES=1 PM_PAIRS=16 timeout 120 python tools/pair_profile.py > $O/plain_pair_$TAG.log 2>&1 &&
PM_LANES=1 PM_PAIRS=16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_pair_$TAG.csv python tools/pair_profile.py > $O/ncu_pair_$TAG.log 2>&1
echo "pair launch list exit $?"
timeout 120 python tools/step_timeline.py > $O/timeline_$TAG.txt 2>&1
PM_PIPE=1 timeout 120 python tools/step_timeline.py > $O/timeline_pipe_$TAG.txt 2>&1
PM_SURF=1 timeout 120 python tools/step_timeline.py > $O/timeline_surf_$TAG.txt 2>&1
ls -la $O | grep $TAG | head -40
