#!/usr/bin/env python
"""Small end-to-end pass over every kernel family for compute-sanitizer --tool memcheck."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import points_matching_b200 as pm
from points_matching_b200 import synth, _lib
ctx = pm.Context(0)
q, t = synth.sift_pair(700, 900, seed=3)
knn = ctx.knn2(q, t, pm.NORM_L2); good = ctx.ratio_filter(knn, 0.75)
q2, t2 = synth.surf_pair(600, 1000, seed=4)
knn2 = ctx.knn2(q2, t2, pm.NORM_L2)
x = ctx.match_cross(q[:300], t[:400], pm.NORM_L2)
qb, tb = synth.orb_pair(2100, 2200, seed=5)
for path in (1, 2):
    _lib.lib().pm_debug_hamming_path(path)
    h = ctx.knn2(qb, tb, pm.NORM_HAMMING); hx = ctx.match_cross(qb, tb, pm.NORM_HAMMING)
_lib.lib().pm_debug_hamming_path(0)
p1, p2, gt = synth.correspondences(3000, seed=1)
r8 = ctx.find_fundamental(p1, p2, sample_size=8, n_hyp=1024, refit=True, seed=3)
r7 = ctx.find_fundamental(p1, p2, sample_size=7, n_hyp=512, metric=pm.METRIC_SYMEPI, refit=False, seed=4)
lm = ctx.find_fundamental_lmeds(p1[:500], p2[:500], n_hyp=100, seed=5)
lines = ctx.epilines(p1[:100], 1, r8[0]); res = ctx.residuals(p1, p2, r8[0])
# round 2: cross-check over marked rows (few queries, many train rows) and its full-pass twin, the dispatch table, the
# adaptive estimator, pair groups from host buffers (ragged sizes, two lanes), split-mode fallback with helper blocks
xs = ctx.match_cross(qb[:300], tb, pm.NORM_HAMMING); xl = ctx.match_cross(q[:200], t, pm.NORM_L2)
_lib.lib().pm_debug_cross_full(1); xf = ctx.match_cross(qb[:300], tb, pm.NORM_HAMMING); _lib.lib().pm_debug_cross_full(0)
assert len(xs) == len(xf)
for n, method in ((7, pm.FM_7POINT), (12, pm.FM_8POINT), (14, pm.FM_RANSAC), (200, pm.FM_RANSAC), (60, pm.FM_LMEDS)):
    ctx.find_fundamental_mat(p1[:n], p2[:n], method)
ad = ctx.find_fundamental_adaptive(p1, p2, max_iters=400, batch=128, seed=2)
import torch
pairs = [synth.image_pair(900 + 37 * i, 1000 - 41 * i, seed=20 + i) for i in range(5)]
ctx.set_batch_lanes(2)
d1 = [np.ascontiguousarray(pp[0].astype(np.uint8)) for pp in pairs]; d2 = [np.ascontiguousarray(pp[1].astype(np.uint8)) for pp in pairs]
hb = ctx.match_estimate_batched([a.ctypes.data for a in d1], [len(a) for a in d1], [a.ctypes.data for a in d2], [len(a) for a in d2],
                                128, True, [pp[2].ctypes.data for pp in pairs], [pp[3].ctypes.data for pp in pairs], 0.75, 256)
q3, t3 = synth.surf_pair(400, 3000, seed=9); t3[1500] = t3[7] * (1 + 1e-7); t3[2200] = t3[7] * (1 - 1e-7)
k3 = ctx.knn2(q3, t3, pm.NORM_L2)
print("ok", len(good), len(x), len(hx), r8[2], r7[2], lm[2], len(xs), len(xl), ctx.l2_stats())
