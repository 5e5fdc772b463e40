#!/bin/bash
# two GPUs: the NCCL paths (torch.distributed protocol AND the C ABI's own communicator), the C++ host example, bench at N=2
TAG=${1:-n2}
O=gpurun_out; mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_sharded_nccl.py -m gpu -x -q --timeout=600 > $O/pytest_$TAG.log 2>&1; echo "pytest nccl exit $?"; tail -12 $O/pytest_$TAG.log
timeout 300 ./examples/sharded_two_gpus 2 > $O/cpp_two_gpus_$TAG.log 2>&1; echo "cpp example exit $?"; tail -3 $O/cpp_two_gpus_$TAG.log
PM_BENCH_WATCHDOG_S=500 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench n2 exit $?"; tail -5 $O/bench_$TAG.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench_$TAG.json"))
    print(json.dumps(d["summary"], indent=1))
except Exception as e:
    print("no bench json", e)
PY
