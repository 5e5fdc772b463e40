#!/usr/bin/env python
"""Runs only the cfg3 Hamming shard (tensor-core path) a few times: the command ncu wraps to capture l2_tc_kernel<FP8>."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import points_matching_b200 as pm
from points_matching_b200 import synth
ctx = pm.Context(0)
nq, nt = 12500, 100000
q, t = synth.orb_pair(nq, nt, seed=4321)
dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
knn = torch.zeros((nq, 2, 4), dtype=torch.int32, device="cuda")
for _ in range(4):
    ctx.knn2_hamming_dev(dq.data_ptr(), nq, dt.data_ptr(), nt, 32, knn.data_ptr(), 0)
ctx.sync()
print("ok", knn[:2].cpu().tolist())
