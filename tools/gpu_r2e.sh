#!/bin/bash
TAG=${1:-r2e}
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=300 > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -6 $O/pytest_$TAG.log
PM_BENCH_WATCHDOG_S=400 timeout 450 python bench.py --steps 200 --warmup 10 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"; tail -4 $O/bench_$TAG.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench_$TAG.json"))
    print(json.dumps(d["summary"], indent=1))
    print("roofline frac", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "share", d["roofline"]["kernel_share_of_step"])
except Exception as e:
    print("no bench json", e)
PY
for c in 0 3 40; do PM_TRACE_CTA=$c PM_LIBPM_SO=$PWD/points_matching_b200/libpm_trace.so timeout 120 python tools/k2_trace2.py > $O/k2trace_${TAG}_$c.txt 2>&1; echo "trace $c exit $?"; head -24 $O/k2trace_${TAG}_$c.txt; done
PM_PIPE=1 timeout 120 python tools/step_timeline.py > $O/timeline_pipe_$TAG.txt 2>&1; tail -4 $O/timeline_pipe_$TAG.txt
