#!/bin/bash
# N-GPU validation: the NCCL tests (two ranks) and the default bench line under torch.distributed.run.
# usage: bash tools/gpu_multi.sh <N> <tag>
N=${1:-2}; TAG=${2:-n$N}
O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -8
timeout 300 python -m pytest tests/test_sharded_nccl.py -x -q -m gpu 2>&1 | tail -4
PM_BENCH_WATCHDOG_S=400 timeout 460 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus $N --steps 2000 --warmup 20 > $O/bench_$TAG.json 2> $O/bench_$TAG.err
echo "bench N=$N exit $?"; grep "bench.py\|Error\|error" $O/bench_$TAG.err | tail -12
python tools/show_bench.py $O/bench_$TAG.json 2>/dev/null | head -40
