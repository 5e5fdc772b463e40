#!/bin/bash
TAG=${1:-r2m}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 -k "pair or batched or ransac or fundamental or smoke or lmeds or sharded" > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 $O/pytest_$TAG.log
for cfg in "1 8" "2 8" "4 8" "8 8" "4 16" "8 4" "8 2" "8 1" "3 8" "6 8"; do set -- $cfg; PM_LANES=$1 PM_PAIR_GROUP=$2 timeout 120 python tools/cfg5_sweep.py 2>&1 | tail -1; done
PM_LANES=1 PM_PAIRS=8 timeout 120 python tools/pair_profile.py > $O/plain_pair_$TAG.log 2>&1 && PM_LANES=1 PM_PAIRS=8 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_pair_$TAG.csv python tools/pair_profile.py > $O/ncu_pair_$TAG.log 2>&1; echo "pair launch list exit $?"
python tools/show_bench.py $O/launches_pair_$TAG.csv 2>&1 | tail -25
