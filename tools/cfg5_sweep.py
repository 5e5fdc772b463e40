#!/usr/bin/env python
"""cfg5 throughput of pm_match_estimate_batched_dev for the lanes / group size in the environment (PM_LANES, PM_PAIR_GROUP)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import points_matching_b200 as pm
from points_matching_b200 import synth
from points_matching_b200.pipeline import match_and_estimate_batch_native
dev = torch.device("cuda:0")
n, pool = 8192, []
for k in range(4):
    d1, d2, k1, k2, _ = synth.image_pair(n, n, seed=100 + k)
    pool.append(tuple(torch.from_numpy(a).to(dev) for a in (d1, d2, k1, k2)))
plist = [pool[p % 4] for p in range(1024)]
ctx = pm.Context(0)
lanes = int(os.environ.get("PM_LANES", "4"))
ctx.set_batch_lanes(lanes)
ctx.batch_warmup(n, n, 128, False, 4096)
match_and_estimate_batch_native(ctx, plist[:64], n_hyp=4096)
best = 1e9
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = match_and_estimate_batch_native(ctx, plist, n_hyp=4096)
    torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
print("lanes %d group %s: %.1f us per pair, %.0f pairs/s  (pair 0: %d matches, %d inliers)" %
      (lanes, os.environ.get("PM_PAIR_GROUP", "8"), best / 1024 * 1e6, 1024 / best, res[0][1]["n_matches"], res[0][1]["n_inliers"]))
