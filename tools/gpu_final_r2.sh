#!/bin/bash
# end-of-round check on one GPU: the whole GPU test suite, smoke(), the default bench line, and the ncu launch list of the
# short bench command (all legs) + of one cfg3 cross-check step
TAG=${1:-r2b}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
PM_BENCH_WATCHDOG_S=500 timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "reference arm exit $?"; tail -c 600 $O/bench_ref_$TAG.json
SHORT="python bench.py --steps 10 --warmup 3 --no-ramp --no-cpu --no-cfg5 --ransac-steps 2 --lanes 1"
timeout 300 $SHORT > $O/plain_$TAG.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/launches_$TAG.csv $SHORT > $O/ncu_list_$TAG.log 2>&1
echo "launch list exit $?"
timeout 120 python tools/cross_only.py 3 > $O/plain_cross_$TAG.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_cross_$TAG.csv python tools/cross_only.py 2 > $O/ncu_cross_$TAG.log 2>&1
echo "cross launch list exit $?"
timeout 120 python tools/step_timeline.py > $O/timeline_$TAG.txt 2>&1
PM_SURF=1 timeout 120 python tools/step_timeline.py > $O/timeline_surf_$TAG.txt 2>&1
python tools/show_bench.py $O/bench_$TAG.json | head -30
