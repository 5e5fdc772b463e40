#!/bin/bash
# N = 8: host-buffer e2e with and without binding every rank to its GPU's local cores
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/topo_n8.txt 2>&1; lscpu | grep -i "numa\|socket\|^CPU(s)" >> $O/topo_n8.txt
for mode in aff noaff; do
  extra=""; [ $mode = noaff ] && extra="--no-affinity"
  PM_BENCH_WATCHDOG_S=200 timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus 8 --steps 500 --warmup 10 --no-ransac --no-hamming --no-cfg5 --no-cpu $extra > $O/bench_n8_$mode.json 2> $O/bench_n8_$mode.err
  echo "$mode exit $?"
  python - <<PY
import json
d=json.loads(open("$O/bench_n8_$mode.json").read().strip().splitlines()[-1])
e=d["e2e"]; print("$mode", d["value"], e["value"], e["ms_per_step"], e["lanes"], {k:v for k,v in e.items() if k.startswith("ms_per_step_")and k!="ms_per_step_blocks"}, e["u8_wire_format"]["ms_per_step"], d["notes"]["host_affinity"])
PY
done
cat $O/topo_n8.txt | head -30
