#!/usr/bin/env python
"""Per-tile clock64 stamps of CTA 0 in K2 (PM_K2_DBG must include 32)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import points_matching_b200 as pm
from points_matching_b200 import _lib
ctx = pm.Context(0)
nq, nt = 148 * 256, 128 * 32
rng = np.random.default_rng(0)
q = torch.from_numpy(rng.integers(0, 200, (nq, 128)).astype(np.float32)).cuda()
t = torch.from_numpy(rng.integers(0, 200, (nt, 128)).astype(np.float32)).cuda()
out = torch.zeros((nq, 2, 4), dtype=torch.int32, device="cuda")
tr = torch.zeros((64, 16), dtype=torch.int64, device="cuda")
_lib.lib().pm_debug_set_k2_trace(C.c_void_p(tr.data_ptr()))
for _ in range(3):
    ctx.knn2_l2_f32_dev(q.data_ptr(), nq, t.data_ptr(), nt, 128, out.data_ptr())
ctx.sync(); torch.cuda.synchronize()
a = tr.cpu().numpy()[:32]
print("dbg", os.environ.get("PM_K2_DBG"))
names = ["top", "waits", "order_sync", "issue", "commits"]
print("issuer A (even items) stamps, cycles; loop = top(lt+2) - top(lt) = two items:")
print(" ".join(f"{n:>10s}" for n in names[1:5]), "| loop(2 items) | epilogue (rel. top): wait_start tfull_seen arrive | busy")
for i in range(4, 24, 2):
    r = a[i]
    d = [r[k] - r[k - 1] for k in range(1, 5)]
    print(" ".join(f"{x:10d}" for x in d), "|", a[i + 2][0] - r[0], "|", r[12] - r[0], r[13] - r[0], r[14] - r[0], "|", r[14] - r[13])
