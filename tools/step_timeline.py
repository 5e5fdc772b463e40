#!/usr/bin/env python
"""Timeline of one cfg2 step (K1 pack -> K2 GEMM -> K3 finish (+ in-kernel fallback) -> K5 ratio filter) under the PDL
launch chain: per kernel, first CTA entry / first CTA past griddepcontrol.wait / last CTA exit (globaltimer, ns)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import points_matching_b200 as pm
from points_matching_b200 import _lib, synth
ctx = pm.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
NQ = NT = 10000
q, t = (synth.surf_pair(NQ, NT, seed=77) if os.environ.get("PM_SURF", "0") != "0" else synth.sift_pair(NQ, NT, seed=1234))
dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
knn = torch.zeros((NQ, 2, 4), dtype=torch.int32, device="cuda")
good = torch.zeros((NQ, 4), dtype=torch.int32, device="cuda")
ng = torch.zeros(4, dtype=torch.int32, device="cuda")
FUSED = os.environ.get("PM_FUSED", "1") != "0"
if os.environ.get("PM_PIPE", "0") != "0":
    ctx.set_pipelining(True)
def step():
    if FUSED:      # one call, same chain
        ctx.knn2_ratio_l2_f32_dev(dq.data_ptr(), NQ, dt.data_ptr(), NT, 128, 0.75, knn.data_ptr(), good.data_ptr(), ng.data_ptr(), 0)
        return
    ctx.knn2_l2_f32_dev(dq.data_ptr(), NQ, dt.data_ptr(), NT, 128, knn.data_ptr(), 0)
    ctx.ratio_filter_dev(knn.data_ptr(), NQ, 0.75, good.data_ptr(), ng.data_ptr())
for _ in range(200):
    step()
torch.cuda.synchronize()
span = torch.zeros(15 + 17 + 2 * 1024, dtype=torch.int64, device="cuda")
names = ["K1 pack", "K2 gemm", "K3 finish", None, "K5 filter"]     # slot 3 (stand-alone exact kernel) is no longer in the chain
rows = []
for rep in range(5):
    init = np.zeros(15 + 17 + 2 * 1024, dtype=np.int64)
    init[[0, 1, 3, 4, 6, 7, 9, 10, 12, 13]] = np.iinfo(np.int64).max
    span.copy_(torch.from_numpy(init)); torch.cuda.synchronize()
    _lib.lib().pm_debug_set_span(C.c_void_p(span.data_ptr()))
    for _ in range(3):
        step()                      # min/max over three back-to-back steps would smear: only the LAST step's marks are wanted
    torch.cuda.synchronize()
    _lib.lib().pm_debug_set_span(C.c_void_p(0))
    # three steps were recorded into the same slots: rerun with a single step for clean marks
    span.copy_(torch.from_numpy(init)); torch.cuda.synchronize()
    for _ in range(20):
        step()
    _lib.lib().pm_debug_set_span(C.c_void_p(span.data_ptr()))
    step()
    _lib.lib().pm_debug_set_span(C.c_void_p(0))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    rows.append(span.cpu().numpy().copy())
r = rows[-1]
hb = r[32 + 2 * 313:32 + 2 * 444].reshape(-1, 2)
stay = hb[:, 1] > 0
if stay.any():
    hs, he = (hb[stay, 0] - r[7]) / 1e3, (hb[stay, 1] - r[7]) / 1e3
    print("K3 helpers: %d of %d stayed; start (past their waits) min/med/max %.2f %.2f %.2f  end min/med/max %.2f %.2f %.2f us"
          % (stay.sum(), len(hb), hs.min(), np.median(hs), hs.max(), he.min(), np.median(he), he.max()))
else:
    print("K3 helpers: none stayed (or exact mode)")
blk = r[32:32 + 2 * 313].reshape(-1, 2)
st, en = (blk[:, 0] - r[7]) / 1e3, (blk[:, 1] - r[7]) / 1e3
print("K3 per-block (us after the first block passed the wait): start min/med/max %.2f %.2f %.2f  end min/med/max %.2f %.2f %.2f  dur med %.2f max %.2f"
      % (st.min(), np.median(st), st.max(), en.min(), np.median(en), en.max(), np.median(en - st), (en - st).max()))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(200):
    step()
ev0.record(stream)
for _ in range(2000):
    step()
ev1.record(stream); torch.cuda.synchronize()
print("fused" if FUSED else "two calls", "step %.2f us (2000 back-to-back steps, one input set)" % (ev0.elapsed_time(ev1) / 2000 * 1e3))
for r in rows:
    t0 = r[0]
    print(" | ".join(f"{n}: in {(r[3*k]-t0)/1e3:6.2f} dep {(r[3*k+1]-t0)/1e3:6.2f} out {(r[3*k+2]-t0)/1e3:6.2f}" for k, n in enumerate(names) if n), " (us)")
