#!/usr/bin/env python
"""K2 timing vs tiles per CTA: nq = 148*256 rows (one 256-row work-item row per CTA), nt = 128*T."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import points_matching_b200 as pm

ctx = pm.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
nq = 148 * 256
rng = np.random.default_rng(0)
q = torch.from_numpy(rng.integers(0, 200, (nq, 128)).astype(np.float32)).cuda()
for T in (1, 2, 4, 8, 16, 32):
    nt = 128 * T
    t = torch.from_numpy(rng.integers(0, 200, (nt, 128)).astype(np.float32)).cuda()
    out = torch.zeros((nq, 2, 4), dtype=torch.int32, device="cuda")
    for _ in range(5):
        ctx.knn2_l2_f32_dev(q.data_ptr(), nq, t.data_ptr(), nt, 128, out.data_ptr())
    ctx.profile_enable(True)
    for _ in range(50):
        ctx.knn2_l2_f32_dev(q.data_ptr(), nq, t.data_ptr(), nt, 128, out.data_ptr())
    ms, n = ctx.profile_read(0)
    ctx.profile_enable(False)
    print(f"dbg={os.environ.get('PM_K2_DBG','0')} tiles/CTA={T:3d}  K2 = {1e3*ms/n:8.2f} us   per tile {1e3*ms/n/T:6.2f} us", ctx.l2_stats())
