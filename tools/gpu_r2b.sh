#!/bin/bash
# round 2: parity tests, the new bench line, K2 trace + timelines
TAG=${1:-r2b}
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=400 > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -15 $O/pytest_$TAG.log
timeout 600 python bench.py --steps 200 --warmup 10 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"; tail -5 $O/bench_$TAG.err
python - <<PY
import json
d=json.load(open("$O/bench_$TAG.json"))
print(json.dumps(d["summary"], indent=1))
print("roofline frac", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "share", d["roofline"]["kernel_share_of_step"])
PY
PM_LIBPM_SO=$PWD/points_matching_b200/libpm_trace.so timeout 120 python tools/k2_trace2.py > $O/k2trace_$TAG.txt 2>&1; echo "trace exit $?"; cat $O/k2trace_$TAG.txt
PM_PIPE=1 timeout 120 python tools/step_timeline.py > $O/timeline_pipe_$TAG.txt 2>&1; tail -4 $O/timeline_pipe_$TAG.txt
