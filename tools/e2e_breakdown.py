#!/usr/bin/env python
"""Where the end-to-end (host buffer) cfg2 call spends its time: raw PCIe copies vs the libpm host call."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import points_matching_b200 as pm
from points_matching_b200 import synth
NQ = NT = 10000
ctx = pm.Context(0)
q, t = synth.sift_pair(NQ, NT, seed=1234)
hq, ht = torch.from_numpy(q).pin_memory(), torch.from_numpy(t).pin_memory()
dq, dt = torch.empty_like(hq, device="cuda"), torch.empty_like(ht, device="cuda")
hknn = torch.zeros((NQ, 2, 4), dtype=torch.int32).pin_memory()
hgood = torch.zeros((NQ, 4), dtype=torch.int32).pin_memory()
dk = torch.zeros((NQ, 2, 4), dtype=torch.int32, device="cuda")
def timeit(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
def h2d():
    dq.copy_(hq, non_blocking=True); dt.copy_(ht, non_blocking=True); torch.cuda.synchronize()
def d2h():
    hknn.copy_(dk, non_blocking=True); torch.cuda.synchronize()
def call():
    ctx.knn2_ratio_l2_ptr(hq.data_ptr(), NQ, ht.data_ptr(), NT, 128, 0.75, hknn.data_ptr(), hgood.data_ptr())
def dev():
    ctx.knn2_l2_f32_dev(dq.data_ptr(), NQ, dt.data_ptr(), NT, 128, dk.data_ptr(), 0); ctx.sync()
print(f"H2D 10.24 MB + sync: {timeit(h2d):7.1f} us  ({10.24e6 / timeit(h2d) / 1e3:.1f} GB/s)")
print(f"D2H 320 KB + sync:   {timeit(d2h):7.1f} us")
print(f"device-resident knn2 + sync: {timeit(dev):7.1f} us")
print(f"host call pm_knn2_ratio_l2_f32: {timeit(call):7.1f} us")
