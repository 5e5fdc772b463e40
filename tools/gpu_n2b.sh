#!/bin/bash
TAG=${1:-n2b}
O=gpurun_out; mkdir -p $O
for i in 1 2 3; do timeout 120 ./examples/sharded_two_gpus 2 > $O/cpp_two_gpus_$TAG.log 2>&1; echo "cpp example run $i exit $?"; tail -2 $O/cpp_two_gpus_$TAG.log; done
PM_BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-ransac --no-hamming --no-cpu > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench n2 (cfg5 only) exit $?"; grep "bench.py\|File\|line" $O/bench_$TAG.err | tail -40
