#!/bin/bash
# round 2, first GPU call: parity tests after the batch-1 changes, a bench line, the K2 clock64 trace and the step timeline
TAG=${1:-r2a}
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=400 > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -15 $O/pytest_$TAG.log
timeout 300 python bench.py --steps 200 --warmup 10 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"
python tools/show_bench.py $O/bench_$TAG.json 2>&1 | tail -30
PM_LIBPM_SO=$PWD/points_matching_b200/libpm_trace.so timeout 120 python tools/k2_trace2.py > $O/k2trace_$TAG.txt 2>&1; echo "trace exit $?"; cat $O/k2trace_$TAG.txt
timeout 120 python tools/step_timeline.py > $O/timeline_$TAG.txt 2>&1; tail -8 $O/timeline_$TAG.txt
PM_PIPE=1 timeout 120 python tools/step_timeline.py > $O/timeline_pipe_$TAG.txt 2>&1; tail -5 $O/timeline_pipe_$TAG.txt
