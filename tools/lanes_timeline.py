#!/usr/bin/env python
"""Timeline of ONE cfg2 step of lane 0 while four contexts take steps in turn (globaltimer marks, us):
first CTA in / first CTA past its waits / last CTA out of K1, K2, K3, K5.  Usage: python tools/lanes_timeline.py [lanes]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import points_matching_b200 as pm
from points_matching_b200 import _lib, synth

NQ, NT, DIM, RATIO = 10000, 10000, 128, 0.75
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
q0, t0 = synth.sift_pair(NQ, NT, seed=1234)
gq, gt = torch.from_numpy(q0).to(dev), torch.from_numpy(t0).to(dev)
fn = _lib.lib().pm_knn2_ratio_l2_f32_dev
lanes = []
for l in range(L):
    ctx = pm.Context(0)
    st = torch.cuda.Stream(device=dev)
    ctx.set_stream(st.cuda_stream)
    knn = torch.zeros((NQ, 2, 4), dtype=torch.int32, device=dev)
    good = torch.zeros((NQ, 4), dtype=torch.int32, device=dev)
    ngood = torch.zeros(4, dtype=torch.int32, device=dev)
    args = (ctx._h, C.c_void_p(gq.data_ptr()), C.c_int(NQ), C.c_void_p(gt.data_ptr()), C.c_int(NT), C.c_int(DIM),
            C.c_float(RATIO), C.c_int(0), C.c_void_p(knn.data_ptr()), C.c_void_p(good.data_ptr()), C.c_void_p(ngood.data_ptr()))
    lanes.append((ctx, st, args, knn, good, ngood))
for i in range(400):
    fn(*lanes[i % L][2])
torch.cuda.synchronize()
span = torch.zeros(15 + 17 + 2 * 1024, dtype=torch.int64, device=dev)
names = ["K1 pack", "K2 gemm", "K3 finish", None, "K5 filter"]
init = np.zeros(span.numel(), dtype=np.int64)
init[[0, 1, 3, 4, 6, 7, 9, 10, 12, 13]] = np.iinfo(np.int64).max
k2 = []
for rep in range(12):
    span.copy_(torch.from_numpy(init)); torch.cuda.synchronize()
    for i in range(40 * L):
        fn(*lanes[i % L][2])
    _lib.lib().pm_debug_set_span(C.c_void_p(span.data_ptr()))
    fn(*lanes[0][2])                                   # the instrumented step (lane 0)
    _lib.lib().pm_debug_set_span(C.c_void_p(0))
    for i in range(1, 4 * L):
        fn(*lanes[i % L][2])
    torch.cuda.synchronize()
    r = span.cpu().numpy()
    t0_ = r[0]
    if rep >= 7:
        print(" | ".join(f"{n}: in {(r[3*k]-t0_)/1e3:6.2f} dep {(r[3*k+1]-t0_)/1e3:6.2f} out {(r[3*k+2]-t0_)/1e3:6.2f}" for k, n in enumerate(names) if n), " (us)")
    k2.append(((r[4] - r[3]) / 1e3, (r[5] - r[4]) / 1e3))
print("K2 of lane 0 among %d lanes: first CTA in -> first past its waits, median %.2f us; past the waits -> last CTA out, median %.2f us (min %.2f max %.2f)"
      % (L, np.median([a for a, _ in k2]), np.median([b for _, b in k2]), min(b for _, b in k2), max(b for _, b in k2)))
