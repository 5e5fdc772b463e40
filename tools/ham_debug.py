#!/usr/bin/env python
"""Bring-up diagnostics for the tensor-core Hamming path: dumps K2's (popc(b) - 2 a.b) tile values and compares
them with numpy, then compares the kNN result with the oracle."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import points_matching_b200 as pm
from points_matching_b200 import _lib, synth
from oracle import oracle as orc
L = _lib.lib()
ctx = pm.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for nq, nt in ((256, 256), (300, 700), (3000, 5000)):
    q, t = synth.orb_pair(nq, nt, seed=5)
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    mq, ntp = (nq + 255) // 256 * 256, (nt + 255) // 256 * 256
    dump = torch.full((mq, ntp), float("nan"), device="cuda")
    out = torch.zeros((nq, 2, 4), dtype=torch.int32, device="cuda")
    L.pm_debug_hamming_path(2)
    L.pm_debug_set_l2_dump(C.c_void_p(dump.data_ptr()))
    ctx.knn2_hamming_dev(dq.data_ptr(), nq, dt.data_ptr(), nt, 32, out.data_ptr(), 0)
    torch.cuda.synchronize(); ctx.sync()
    L.pm_debug_set_l2_dump(C.c_void_p(0)); L.pm_debug_hamming_path(0)
    qb = np.unpackbits(q, axis=1, bitorder="little").astype(np.float64)
    tb = np.unpackbits(t, axis=1, bitorder="little").astype(np.float64)
    ref = tb.sum(1)[None, :] - 2.0 * qb @ tb.T
    d = dump[:nq, :nt].double().cpu().numpy()
    err = np.abs(d - ref)
    main_got = d - tb.sum(1)[None, :]
    main_ref = -2.0 * qb @ tb.T
    print("   main part got", main_got[0, :8], "ref", main_ref[0, :8], " ratio", (main_got[0, :8] / main_ref[0, :8]))
    print(f"[{nq}x{nt}] nan={int(np.isnan(d).sum())} max_abs_err={np.nanmax(err):.3e}  sample got {d[0,:6]} ref {ref[0,:6]}")
    res = out.cpu().numpy().view(pm.DMATCH).reshape(nq, 2)
    r = orc.knn2_hamming(q, t)
    print("   idx equal rows:", (res["trainIdx"] == r["trainIdx"]).all(axis=1).mean(), " dist equal:", (res["distance"] == r["distance"]).mean(),
          " first rows got", res[:2].tolist(), "ref", r[:2].tolist())
