#!/bin/bash
TAG=${1:-r2i}
O=gpurun_out; mkdir -p $O
PM_SURF=1 timeout 120 python tools/step_timeline.py > $O/timeline_surf_$TAG.txt 2>&1; tail -7 $O/timeline_surf_$TAG.txt
timeout 120 python tools/step_timeline.py > $O/timeline_$TAG.txt 2>&1; tail -4 $O/timeline_$TAG.txt
