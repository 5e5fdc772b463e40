#!/bin/bash
TAG=${1:-r2h}
O=gpurun_out; mkdir -p $O
PM_SURF=1 timeout 120 python tools/step_timeline.py > $O/timeline_surf_$TAG.txt 2>&1; tail -4 $O/timeline_surf_$TAG.txt
PM_SURF=1 PM_TRACE_CTA=0 PM_LIBPM_SO=$PWD/points_matching_b200/libpm_trace.so timeout 120 python tools/k2_trace2.py > $O/k2trace_surf_$TAG.txt 2>&1; echo "trace exit $?"; head -24 $O/k2trace_surf_$TAG.txt
timeout 600 python -m pytest tests -m gpu -x -q --timeout=300 -k "l2 or split or surf or fused or pipelined" > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -3 $O/pytest_$TAG.log
