#!/usr/bin/env python
"""cfg5 pair pipeline (pm_match_estimate_batched_dev) on a few 8192 x 8192 pairs: wall clock per pair, and -- run under
`ncu --metrics gpu__time_duration.sum` -- the per-kernel launch list of one pair."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import points_matching_b200 as pm
from points_matching_b200 import synth
from points_matching_b200.pipeline import match_and_estimate_batch_native
N = int(os.environ.get("PM_PAIR_N", "8192")); NP = int(os.environ.get("PM_PAIRS", "16")); NH = int(os.environ.get("PM_PAIR_HYP", "4096"))
if os.environ.get("PM_INIT_DIST"):        # mimic bench.py under torchrun: NCCL process group in the same process
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    dist.barrier()
DEV = int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("PM_INIT_DIST") else 0
ctx = pm.Context(DEV)
ctx.set_batch_lanes(int(os.environ.get("PM_LANES", "4")))
pool = []
for k in range(4):
    d1, d2, k1, k2, _ = synth.image_pair(N, N, seed=100 + k)
    pool.append(tuple(torch.from_numpy(a).to(f"cuda:{DEV}") for a in (d1, d2, k1, k2)))
plist = [pool[p % 4] for p in range(NP)]
WORLD = int(os.environ.get("WORLD_SIZE", "1")) if os.environ.get("PM_INIT_DIST") else 1
ctx.batch_warmup(N, N, 128, False, NH)       # every lane exists and owns its workspaces before anything is timed
match_and_estimate_batch_native(ctx, plist[:4 * WORLD], n_hyp=NH)
torch.cuda.synchronize(); t0 = time.perf_counter()
out = match_and_estimate_batch_native(ctx, plist, n_hyp=NH)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
torch.cuda.synchronize(); t0 = time.perf_counter()
match_and_estimate_batch_native(ctx, plist, n_hyp=NH, sync=False)
t_enq = time.perf_counter() - t0
torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print("  enqueue returns after %.1f us per pair, everything done after %.1f us per pair" % (t_enq / (NP // WORLD) * 1e6, t_all / (NP // WORLD) * 1e6))
print("native batched: %.1f us per pair and rank (%d pairs in all, %d x %d, %d hypotheses); last: %d matches, %d inliers; %d launches"
      % (dt / (NP // WORLD) * 1e6, NP, N, N, NH, out[-1][1]["n_matches"], out[-1][1]["n_inliers"], ctx.launch_count()))
