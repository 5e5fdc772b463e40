#!/usr/bin/env python
"""Bring-up diagnostics for the tcgen05 L2 kernel (run on the GPU box):
dumps the (||b||^2 - 2ab) tile values the tensor-core kernel produced and compares
them with a float64 computation, then checks kNN results against the exact FP32 kernel.
Usage: python tools/gpu_debug.py [nq nt]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from points_matching_b200 import Context, synth, _lib  # noqa: E402


def run(nq, nt, kind):
    L = _lib.lib()
    ctx = Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    if kind == "sift":
        q, t = synth.sift_pair(nq, nt, seed=5)
    else:
        q, t = synth.surf_pair(nq, nt, seed=5)
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    mq_pad, nt_pad = (nq + 255) // 256 * 256, (nt + 255) // 256 * 256
    dump = torch.full((mq_pad, nt_pad), float("nan"), device="cuda", dtype=torch.float32)
    out = torch.zeros((nq, 2, 4), device="cuda", dtype=torch.int32)
    L.pm_debug_set_l2_dump(C.c_void_p(dump.data_ptr()))
    ctx.knn2_l2_f32_dev(dq.data_ptr(), nq, dt.data_ptr(), nt, q.shape[1], out.data_ptr())
    torch.cuda.synchronize()
    L.pm_debug_set_l2_dump(C.c_void_p(0))
    st = ctx.l2_stats()
    d = dump[:nq, :nt].double().cpu().numpy()
    ref = (t.astype(np.float64) ** 2).sum(1)[None, :] - 2.0 * q.astype(np.float64) @ t.astype(np.float64).T
    err = np.abs(d - ref)
    print(f"[{kind} {nq}x{nt}] stats={st} nan={int(np.isnan(d).sum())} max_abs_err={np.nanmax(err):.3e} "
          f"ref_scale={np.abs(ref).max():.3e}")
    if np.isnan(d).any() or np.nanmax(err) > 1e-3 * np.abs(ref).max():
        bad = np.argwhere(~(err <= 1e-3 * np.abs(ref).max()))
        print("  first bad (row, col):", bad[:8].tolist())
        r, c = bad[0]
        print("  got", d[r, c:c + 8], "\n  ref", ref[r, c:c + 8])
        rows_bad = np.unique(bad[:, 0]); cols_bad = np.unique(bad[:, 1])
        print("  bad rows", rows_bad[:16], "... n=", len(rows_bad), " bad cols", cols_bad[:16], "... n=", len(cols_bad))
    res = out.cpu().numpy().view(_lib.DMATCH).reshape(nq, 2)
    # exact kernel cross-check
    L.pm_debug_force_exact(1)
    out2 = torch.zeros_like(out)
    ctx.knn2_l2_f32_dev(dq.data_ptr(), nq, dt.data_ptr(), nt, q.shape[1], out2.data_ptr())
    torch.cuda.synchronize()
    L.pm_debug_force_exact(0)
    res2 = out2.cpu().numpy().view(_lib.DMATCH).reshape(nq, 2)
    same_idx = (res["trainIdx"] == res2["trainIdx"]).all(axis=1)
    print(f"  tc vs exact kernel: idx rows equal {same_idx.mean():.4f}, "
          f"max |dist diff| {np.abs(res['distance'] - res2['distance']).max():.3e}")
    # numpy check of the exact kernel on a few rows
    rows = np.arange(0, nq, max(1, nq // 64))
    dd = ((q[rows, None, :].astype(np.float64) - t[None, :, :].astype(np.float64)) ** 2).sum(-1)
    order = np.argsort(dd, axis=1, kind="stable")[:, :2]
    print("  exact kernel vs numpy f64 (sampled rows): idx equal", (order == res2["trainIdx"][rows]).all(axis=1).mean())
    ctx.close()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    if len(sys.argv) >= 3:
        sizes = [(int(sys.argv[1]), int(sys.argv[2]))]
    else:
        sizes = [(128, 256), (300, 700), (1000, 3000)]
    for nq, nt in sizes:
        for kind in ("sift", "surf"):
            run(nq, nt, kind)
