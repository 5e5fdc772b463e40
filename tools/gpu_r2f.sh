#!/bin/bash
TAG=${1:-r2f}
O=gpurun_out; mkdir -p $O
PM_SURF=1 timeout 120 python tools/step_timeline.py > $O/timeline_surf_$TAG.txt 2>&1; tail -8 $O/timeline_surf_$TAG.txt
PM_BENCH_WATCHDOG_S=400 timeout 450 python bench.py --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$O/bench_$TAG.json"))
s=d["summary"]; print({k:(round(v,4) if isinstance(v,float) else v) for k,v in s.items()})
PY
timeout 600 python -m pytest tests -m gpu -x -q --timeout=300 > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -3 $O/pytest_$TAG.log
