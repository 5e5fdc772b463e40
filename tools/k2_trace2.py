#!/usr/bin/env python
"""clock64 trace of CTA 0 of K2 at cfg2 (needs a -DPM_K2_TRACE build of libpm: PM_LIBPM_SO=<path>).
Phases of the kernel (TRK slots) and per-item stamps of issuer warp A / epilogue warp 0."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import points_matching_b200 as pm
from points_matching_b200 import _lib, synth
ctx = pm.Context(0)
NQ = NT = 10000
q, t = (synth.surf_pair(NQ, NT, seed=77) if os.environ.get("PM_SURF", "0") != "0" else synth.sift_pair(NQ, NT, seed=1234))
dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
knn = torch.zeros((NQ, 2, 4), dtype=torch.int32, device="cuda")
tr = torch.zeros((64, 16), dtype=torch.int64, device="cuda")
_lib.lib().pm_debug_set_k2_trace(C.c_void_p(tr.data_ptr()))
CTA = int(os.environ.get("PM_TRACE_CTA", "0"))
_lib.lib().pm_debug_set_k2_trace_cta(CTA)
print("traced CTA", CTA)
for _ in range(5):
    ctx.knn2_l2_f32_dev(dq.data_ptr(), NQ, dt.data_ptr(), NT, 128, knn.data_ptr(), 0)
ctx.sync(); torch.cuda.synchronize()
a = tr.cpu().numpy()
k = a[63]
names = ["entry", "alloc_done", "pdl_wait_done", "tmem_ready(sync)", "first_tfull(epi)", "epi_loop_done", "flush_done", "end"]
print("kernel phases of CTA 0 (cycles since entry):")
for i, n in enumerate(names):
    print(f"  {n:20s} {int(k[i] - k[0]):8d}")
print("issuer A (even items): top->waits, ->order_sync, ->issue, ->commits | top(lt+2)-top(lt) | epilogue warp0: wait_start tfull_seen arrive (rel. top) | busy")
for i in range(0, 22, 2):
    r = a[i]
    if r[0] == 0:
        continue
    d = [int(r[j] - r[j - 1]) for j in range(1, 5)]
    nxt = int(a[i + 2][0] - r[0]) if a[i + 2][0] else -1
    print(f"  item {i:2d}: " + " ".join(f"{x:6d}" for x in d), "|", f"{nxt:6d}", "|", int(r[12] - r[0]), int(r[13] - r[0]), int(r[14] - r[0]), "|", int(r[14] - r[13]))
print("odd items (epilogue warp0 only): tfull_seen - wait_start, arrive - tfull_seen")
for i in range(1, 22, 2):
    r = a[i]
    print(f"  item {i:2d}: wait {int(r[13] - r[12]):6d} busy {int(r[14] - r[13]):6d}  start(rel. item0 top) {int(r[12] - a[0][0]):7d}")
