#!/usr/bin/env python
"""cfg2-sized L2 step on SURF-like (unit-norm signed float) descriptors: K2 runs in split mode (3 bf16 products)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import points_matching_b200 as pm
from points_matching_b200 import synth
ctx = pm.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
N = 10000
for kind in ("surf", "sift"):
    q, t = (synth.surf_pair if kind == "surf" else synth.sift_pair)(N, N, seed=77)
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    knn = torch.zeros((N, 2, 4), dtype=torch.int32, device="cuda")
    good = torch.zeros((N, 4), dtype=torch.int32, device="cuda"); ng = torch.zeros(4, dtype=torch.int32, device="cuda")
    def step():
        ctx.knn2_l2_f32_dev(dq.data_ptr(), N, dt.data_ptr(), N, 128, knn.data_ptr(), 0)
        ctx.ratio_filter_dev(knn.data_ptr(), N, 0.75, good.data_ptr(), ng.data_ptr())
    for _ in range(50): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(500): step()
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 500
    ctx.profile_enable(True)
    for _ in range(100): step()
    k2, n = ctx.profile_read(0); ctx.profile_enable(False)
    print(f"{kind}: step {ms*1e3:.1f} us -> {N*N/(ms*1e-3):.3e} pairs/s; K2 {k2/n*1e3:.1f} us; stats {ctx.l2_stats()}")
