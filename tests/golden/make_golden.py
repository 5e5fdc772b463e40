#!/usr/bin/env python
"""Generates the golden fixtures in this directory from OpenCV 4.13 (cv2).

The reference (/root/reference/Points Matching/main.cpp) has no tests or golden
vectors of its own and its arithmetic lives in un-vendored OpenCV 2.4.13, so the
fixtures are outputs of the same OpenCV calls made through cv2 in the build
container:
    main.cpp:43-46  BFMatcher(NORM_L2 / NORM_HAMMING).knnMatch / match(crossCheck)
    main.cpp:95-98  findFundamentalMat (FM_7POINT / FM_8POINT / FM_LMEDS / FM_RANSAC)
    main.cpp:127    computeCorrespondEpilines
Run here (needs cv2 and, for the image fixture, /root/reference); the .npz files
travel to the GPU box, this script does not need to.

    python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from points_matching_b200 import synth  # noqa: E402


def knn_arrays(m, k=2):
    idx = np.full((len(m), k), -1, np.int32)
    dist = np.full((len(m), k), np.float32(np.finfo(np.float32).max), np.float32)
    for i, row in enumerate(m):
        for j, d in enumerate(row):
            idx[i, j] = d.trainIdx
            dist[i, j] = d.distance
    return idx, dist


def match_arrays(m):
    return (np.array([d.queryIdx for d in m], np.int32), np.array([d.trainIdx for d in m], np.int32),
            np.array([d.distance for d in m], np.float32))


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name), **kw)
    print("wrote", name, {k: getattr(v, "shape", v) for k, v in kw.items()})


def gold_l2():
    out = {}
    # SIFT-like integers with duplicated train rows (forced exact ties)
    q, t = synth.sift_pair(300, 401, seed=11)
    t[17] = t[3]; t[200] = t[3]; t[399] = q[5]; t[7] = q[5]
    idx, dist = knn_arrays(cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, k=2))
    out.update(sift_q=q.astype(np.uint8), sift_t=t.astype(np.uint8), sift_idx=idx, sift_dist=dist)
    # SURF-like unit-norm floats
    q, t = synth.surf_pair(256, 333, seed=12)
    idx, dist = knn_arrays(cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, k=2))
    out.update(surf_q=q, surf_t=t, surf_idx=idx, surf_dist=dist)
    # all-equal rows: every distance ties -> indices 0,1
    q = np.ones((5, 128), np.float32); t = np.ones((9, 128), np.float32)
    idx, dist = knn_arrays(cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, k=2))
    out.update(eq_idx=idx, eq_dist=dist)
    # k > ntrain: rows shorter than k
    q, t = synth.sift_pair(4, 1, seed=13)
    m = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, k=2)
    out.update(short_len=np.array([len(r) for r in m], np.int32))
    # ratio test on the sift case
    m = cv2.BFMatcher(cv2.NORM_L2).knnMatch(out["sift_q"].astype(np.float32), out["sift_t"].astype(np.float32), k=2)
    good = [a for a, b in m if a.distance < 0.75 * b.distance]
    gq, gt, gd = match_arrays(good)
    out.update(ratio_q=gq, ratio_t=gt, ratio_d=gd)
    save("l2.npz", **out)


def gold_hamming():
    out = {}
    q, t = synth.orb_pair(300, 401, seed=21)
    idx, dist = knn_arrays(cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2))
    out.update(orb_q=q, orb_t=t, orb_idx=idx, orb_dist=dist)
    xq, xt, xd = match_arrays(cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t))
    out.update(orb_xq=xq, orb_xt=xt, orb_xd=xd)
    # low entropy, 4-byte rows: many exact ties
    rng = np.random.default_rng(22)
    q = rng.integers(0, 4, (300, 4), dtype=np.uint8)
    t = rng.integers(0, 4, (301, 4), dtype=np.uint8)
    idx, dist = knn_arrays(cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2))
    out.update(tie_q=q, tie_t=t, tie_idx=idx, tie_dist=dist)
    xq, xt, xd = match_arrays(cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t))
    out.update(tie_xq=xq, tie_xt=xt, tie_xd=xd)
    # odd byte width (OpenCV accepts any)
    q = rng.integers(0, 256, (50, 61), dtype=np.uint8)
    t = rng.integers(0, 256, (70, 61), dtype=np.uint8)
    idx, dist = knn_arrays(cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2))
    out.update(odd_q=q, odd_t=t, odd_idx=idx, odd_dist=dist)
    save("hamming.npz", **out)


def gold_fundamental():
    out = {}
    p1, p2, gt = synth.correspondences(600, seed=31)
    out.update(p1=p1, p2=p2, gt=gt)
    rng = np.random.default_rng(32)
    inl = np.nonzero(gt)[0]
    idx8 = np.stack([rng.choice(inl, 8, replace=False) for _ in range(64)]).astype(np.int32)
    idx7 = np.stack([rng.choice(inl, 7, replace=False) for _ in range(64)]).astype(np.int32)
    F8 = np.stack([cv2.findFundamentalMat(p1[i], p2[i], cv2.FM_8POINT)[0] for i in idx8])
    F7 = np.zeros((64, 3, 3, 3)); n7 = np.zeros(64, np.int32)
    for k, i in enumerate(idx7):
        F = cv2.findFundamentalMat(p1[i], p2[i], cv2.FM_7POINT)[0].reshape(-1, 3, 3)
        n7[k] = len(F); F7[k, : len(F)] = F
    out.update(idx8=idx8, idx7=idx7, F8=F8, F7=F7, n7=n7)
    # N-point 8-point on the true inliers
    out.update(F8_all=cv2.findFundamentalMat(p1[gt], p2[gt], cv2.FM_8POINT)[0])
    # full estimators (deterministic: cv::RNG((uint64)-1))
    for name, meth, thr in [("ransac1", cv2.FM_RANSAC, 1.0), ("ransac3", cv2.FM_RANSAC, 3.0),
                            ("lmeds", cv2.FM_LMEDS, 3.0), ("fm7", cv2.FM_7POINT, 3.0)]:
        F, m = cv2.findFundamentalMat(p1, p2, meth, thr, 0.99)
        out.update({name + "_F": F, name + "_mask": m.ravel().astype(np.uint8)})
    # dispatch table: N<7 -> None, N==7 -> 9x3 / 3x3 stacks
    out.update(n6_none=np.array(cv2.findFundamentalMat(p1[:6], p2[:6], cv2.FM_RANSAC)[0] is None))
    F, m = cv2.findFundamentalMat(p1[inl[:7]], p2[inl[:7]], cv2.FM_RANSAC)
    out.update(n7_F=F, n7_mask=m.ravel().astype(np.uint8))
    # residuals and epilines for the ransac1 model
    F = out["ransac1_F"]
    samp = np.array([cv2.sampsonDistance(np.append(a, 1.0), np.append(b, 1.0), F)
                     for a, b in zip(p1.astype(np.float64), p2.astype(np.float64))])
    out.update(sampson=samp)
    out.update(lines1=cv2.computeCorrespondEpilines(p1.reshape(-1, 1, 2), 1, F).reshape(-1, 3),
               lines2=cv2.computeCorrespondEpilines(p2.reshape(-1, 1, 2), 2, F).reshape(-1, 3))
    save("fundamental.npz", **out)


def gold_image_pair():
    """Config 1 stand-in: img01.JPG / img02.JPG are the commented-out inputs at main.cpp:12-13
    (img1.bmp / img2.bmp are absent from the mount); SIFT stands in for nonfree SURF."""
    d = "/root/reference/Points Matching"
    a = cv2.imread(os.path.join(d, "img01.JPG"), cv2.IMREAD_GRAYSCALE)
    b = cv2.imread(os.path.join(d, "img02.JPG"), cv2.IMREAD_GRAYSCALE)
    if a is None or b is None:
        print("image pair not available; skipping")
        return
    sift = cv2.SIFT_create()
    k1, d1 = sift.detectAndCompute(a, None)
    k2, d2 = sift.detectAndCompute(b, None)
    kp1 = np.array([k.pt for k in k1], np.float32)
    kp2 = np.array([k.pt for k in k2], np.float32)
    assert (d1 == np.round(d1)).all() and d1.max() <= 255
    idx, dist = knn_arrays(cv2.BFMatcher(cv2.NORM_L2).knnMatch(d1, d2, k=2))
    m = cv2.BFMatcher(cv2.NORM_L2).knnMatch(d1, d2, k=2)
    good = [x for x, y in m if x.distance < 0.75 * y.distance]
    gq, gt, gd = match_arrays(good)
    pts1, pts2 = kp1[gq], kp2[gt]
    Fr, mr = cv2.findFundamentalMat(pts1, pts2, cv2.FM_RANSAC, 1.0, 0.99)
    # the reference's literal flow: k=1 match, min/max-midpoint filter, FM_7POINT (-> LMedS)
    m1 = cv2.BFMatcher(cv2.NORM_L2).match(d1, d2)
    mn, mx = 1.0, 0.0
    for x in m1:
        mn = min(mn, x.distance); mx = max(mx, x.distance)
    lit = [x for x in m1 if x.distance < mn + (mx - mn) / 2]
    lq, lt, ld = match_arrays(lit)
    Fl, ml = cv2.findFundamentalMat(kp1[lq], kp2[lt], cv2.FM_7POINT)
    save("image_pair.npz", desc1=d1.astype(np.uint8), desc2=d2.astype(np.uint8), kp1=kp1, kp2=kp2,
         knn_idx=idx, knn_dist=dist, ratio_q=gq, ratio_t=gt, ratio_d=gd,
         ransac_F=Fr, ransac_mask=mr.ravel().astype(np.uint8),
         lit_q=lq, lit_t=lt, lit_d=ld, lit_min=np.float64(mn), lit_max=np.float64(mx),
         lit_F=Fl, lit_mask=ml.ravel().astype(np.uint8))


def gold_dispatch():
    """cv::findFundamentalMat's dispatch table at main.cpp:95-98: N == 7 -> every root of the 7-point
    solver stacked; FM_RANSAC below 15 points -> LMedS; parameter clamps.  (Generate with `--only dispatch`: the other
    fixtures stay as committed.)"""
    out = {}
    p1, p2, gt = synth.correspondences(600, seed=31)
    rng = np.random.default_rng(77)
    inl = np.nonzero(gt)[0]
    sets = [rng.choice(inl, 7, replace=False) for _ in range(10)] + [rng.choice(600, 7, replace=False) for _ in range(6)]
    idx7 = np.stack(sets).astype(np.int32)
    F7 = np.full((len(idx7), 9, 3), np.nan)
    k7 = np.zeros(len(idx7), np.int32)
    for j, i in enumerate(idx7):
        for meth in (cv2.FM_RANSAC, cv2.FM_LMEDS, cv2.FM_8POINT, cv2.FM_7POINT):     # any method: the 7-point roots
            F, m = cv2.findFundamentalMat(p1[i], p2[i], meth)
            assert F.shape[0] % 3 == 0 and (m == 1).all()
            if meth == cv2.FM_RANSAC:
                k7[j] = F.shape[0] // 3; F7[j, : F.shape[0]] = F
            else:
                assert np.array_equal(F, F7[j, : F.shape[0]])
    out.update(p1=p1, p2=p2, gt=gt, idx7=idx7, F7=F7, k7=k7)
    # below 15 points FM_RANSAC is LMedS; from 15 on it is RANSAC
    sub = rng.permutation(inl)[:40]
    out.update(sub=sub.astype(np.int32))
    for n in (8, 11, 14):
        Fr, mr = cv2.findFundamentalMat(p1[sub[:n]], p2[sub[:n]], cv2.FM_RANSAC, 1.0, 0.99)
        Fl, ml = cv2.findFundamentalMat(p1[sub[:n]], p2[sub[:n]], cv2.FM_LMEDS, 1.0, 0.99)
        F7p, m7 = cv2.findFundamentalMat(p1[sub[:n]], p2[sub[:n]], cv2.FM_7POINT, 1.0, 0.99)
        assert np.array_equal(Fr, Fl) and np.array_equal(mr, ml) and np.array_equal(Fr, F7p) and np.array_equal(mr, m7)
        out.update({f"n{n}_F": Fl, f"n{n}_mask": ml.ravel().astype(np.uint8)})
    mix = np.concatenate([sub[:30], rng.choice(np.nonzero(~gt)[0], 10, replace=False)])     # 25% outliers
    out.update(mix=mix.astype(np.int32))
    Fr, mr = cv2.findFundamentalMat(p1[mix], p2[mix], cv2.FM_RANSAC, 1.0, 0.99)
    Fl, ml = cv2.findFundamentalMat(p1[mix], p2[mix], cv2.FM_LMEDS, 1.0, 0.99)
    out.update(n40_ransac_F=Fr, n40_ransac_mask=mr.ravel().astype(np.uint8), n40_lmeds_F=Fl, n40_lmeds_mask=ml.ravel().astype(np.uint8))
    # clamps: param1 <= 0 -> 3, param2 outside (0, 1) -> 0.99
    F0, m0 = cv2.findFundamentalMat(p1[mix], p2[mix], cv2.FM_RANSAC, 0.0, 0.99)
    F3, m3 = cv2.findFundamentalMat(p1[mix], p2[mix], cv2.FM_RANSAC, 3.0, 0.99)
    Fc, mc = cv2.findFundamentalMat(p1[mix], p2[mix], cv2.FM_RANSAC, 3.0, 1.5)
    assert np.array_equal(F0, F3) and np.array_equal(m0, m3) and np.array_equal(Fc, F3) and np.array_equal(mc, m3)
    out.update(n40_thr3_F=F3, n40_thr3_mask=m3.ravel().astype(np.uint8))
    save("dispatch.npz", **out)


if __name__ == "__main__":
    cv2.setRNGSeed(0)
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "dispatch":
        gold_dispatch()
        sys.exit(0)
    gold_l2()
    gold_hamming()
    gold_fundamental()
    gold_image_pair()
    gold_dispatch()
