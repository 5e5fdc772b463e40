#!/bin/bash
TAG=${1:-r2l}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 -k "pair or batched or ransac or fundamental or smoke or lmeds" > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 $O/pytest_$TAG.log
for cfg in "1 1" "1 8" "2 8" "4 8" "8 8" "2 16" "4 16" "4 4" "8 2" "8 1"; do set -- $cfg; PM_LANES=$1 PM_PAIR_GROUP=$2 timeout 120 python tools/cfg5_sweep.py 2>&1 | tail -1; done
