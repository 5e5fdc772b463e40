#!/bin/bash
# Quick GPU loop: L2 parity tests, K2 debug compare, K2 scaling, short bench.  Usage: bash tools/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}; K=${2:-"l2 or cpp or smoke"}
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q -k "$K" > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -5 $O/pytest_$TAG.log
timeout 120 python tools/gpu_debug.py > $O/debug_$TAG.log 2>&1; echo "debug exit $?"; grep -c "equal 1.0000" $O/debug_$TAG.log
timeout 120 python tools/k2_scaling.py 2>&1 | tee $O/scaling_$TAG.log
timeout 300 python bench.py --steps 500 --warmup 10 --no-ransac --no-hamming --no-cpu > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"
python tools/show_bench.py $O/bench_$TAG.json
