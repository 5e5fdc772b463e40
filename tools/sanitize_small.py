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
print("ok", len(good), len(x), len(hx), r8[2], r7[2], lm[2])
