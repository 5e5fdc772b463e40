#!/usr/bin/env python
"""Where the fixed cost of a K2 launch goes (PM_K2_TRACE build): CTA 0 clock stamps for 1 and 8 items per CTA."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import points_matching_b200 as pm
from points_matching_b200 import _lib
ctx = pm.Context(0)
rng = np.random.default_rng(0)
for T in (1, 8):
    nq, nt = 148 * 256, 128 * T
    q = torch.from_numpy(rng.integers(0, 200, (nq, 128)).astype(np.float32)).cuda()
    t = torch.from_numpy(rng.integers(0, 200, (nt, 128)).astype(np.float32)).cuda()
    out = torch.zeros((nq, 2, 4), dtype=torch.int32, device="cuda")
    tr = torch.zeros((64, 16), dtype=torch.int64, device="cuda")
    _lib.lib().pm_debug_set_k2_trace(C.c_void_p(tr.data_ptr()))
    for _ in range(3):
        ctx.knn2_l2_f32_dev(q.data_ptr(), nq, t.data_ptr(), nt, 128, out.data_ptr())
    ctx.sync(); torch.cuda.synchronize()
    r = tr.cpu().numpy()[63]
    names = ["entry", "setup_done(alloc,bars)", "pdl_wait_done", "syncthreads+extA", "epi: first tfull", "epi: loop done", "epi: flush done", "final sync"]
    print(f"items/CTA={T}: " + ", ".join(f"{n}=+{int(r[k] - r[0])}" for k, n in enumerate(names)))
