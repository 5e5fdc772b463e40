#!/bin/bash
for cfg in "4 16" "4 32" "8 16" "6 16" "2 32" "3 16" "4 24" "8 32"; do set -- $cfg; PM_LANES=$1 PM_PAIR_GROUP=$2 timeout 120 python tools/cfg5_sweep.py 2>&1 | tail -1; done
