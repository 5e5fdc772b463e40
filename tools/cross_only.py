#!/usr/bin/env python
"""Runs only the cfg3 cross-check step (pm_match_cross_sharded_dev on one GPU) a few times and times it: the command ncu
wraps for the launch list of that step."""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import points_matching_b200 as pm
from points_matching_b200 import synth
ctx = pm.Context(0)
nq, nt = 12500, 100000
q, t = synth.orb_pair(nq, nt, seed=4321)
dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
knn = torch.zeros((nq, 2, 4), dtype=torch.int32, device="cuda")
col = torch.zeros(nt, dtype=torch.int64, device="cuda")
out = torch.zeros((nq, 4), dtype=torch.int32, device="cuda")
cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
def step():
    ctx.match_cross_sharded_dev(dq.data_ptr(), nq, dt.data_ptr(), nt, 32, pm.NORM_HAMMING, 0, knn.data_ptr(), col.data_ptr(),
                                out.data_ptr(), cnt.data_ptr())
for _ in range(3):
    step()
ctx.sync()
t0 = time.perf_counter()
for _ in range(n):
    step()
ctx.sync()
print("cross-check step %.1f us, %d mutual matches" % ((time.perf_counter() - t0) / n * 1e6, int(cnt[0].item())))
