#!/bin/bash
# Full GPU check: all parity tests, step timeline, bench (no CPU baseline).  Usage: bash tools/gpu_full.sh <tag>
TAG=${1:-f}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 $O/pytest_$TAG.log
timeout 120 python tools/step_timeline.py 2>&1 | tail -3
timeout 120 python tools/k2_scaling.py 2>&1 | tail -3
timeout 300 python bench.py --steps 1000 --warmup 10 --no-cpu > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"
python tools/show_bench.py $O/bench_$TAG.json
