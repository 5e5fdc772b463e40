#!/bin/bash
TAG=${1:-r2j}
O=gpurun_out; mkdir -p $O
PM_SURF=1 timeout 120 python tools/step_timeline.py > $O/timeline_surf_$TAG.txt 2>&1; head -3 $O/timeline_surf_$TAG.txt; tail -3 $O/timeline_surf_$TAG.txt
timeout 600 python -m pytest tests -m gpu -x -q --timeout=300 -k "l2 or split or surf or fused or pipelined or smoke" > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -3 $O/pytest_$TAG.log
