"""Pins the CPU oracle (oracle/pm_oracle.c) against the golden vectors generated from
OpenCV 4.13 by tests/golden/make_golden.py -- the reference itself ships no tests
(SURVEY.md section 4), and its arithmetic is OpenCV's (main.cpp:43-46, 95-98)."""
import numpy as np
import pytest

from points_matching_b200 import synth


def test_l2_sift_bit_exact(golden, orc):
    g = golden["l2"]
    o = orc.knn2_l2(g["sift_q"].astype(np.float32), g["sift_t"].astype(np.float32))
    assert (o["trainIdx"] == g["sift_idx"]).all()
    assert (o["distance"] == g["sift_dist"]).all()          # sqrtf((float)int) bit for bit
    assert (o["queryIdx"][:, 0] == np.arange(300)).all() and (o["imgIdx"] == 0).all()


def test_l2_surf_indices_and_tolerance(golden, orc):
    g = golden["l2"]
    o = orc.knn2_l2(g["surf_q"], g["surf_t"])
    assert (o["trainIdx"] == g["surf_idx"]).all()
    assert np.allclose(o["distance"], g["surf_dist"], rtol=1e-5, atol=0)


def test_l2_all_equal_ties_lowest_index(golden, orc):
    g = golden["l2"]
    o = orc.knn2_l2(np.ones((5, 128), np.float32), np.ones((9, 128), np.float32))
    assert (o["trainIdx"] == g["eq_idx"]).all() and (g["eq_idx"] == [[0, 1]] * 5).all()
    assert (o["distance"] == g["eq_dist"]).all()


def test_l2_short_rows(golden, orc):
    q, t = synth.sift_pair(4, 1, seed=13)
    o = orc.knn2_l2(q, t)
    assert (golden["l2"]["short_len"] == 1).all()
    assert (o["trainIdx"][:, 0] == 0).all() and (o["trainIdx"][:, 1] == -1).all()


def test_ratio_filter(golden, orc):
    g = golden["l2"]
    o = orc.knn2_l2(g["sift_q"].astype(np.float32), g["sift_t"].astype(np.float32))
    r = orc.ratio_filter(o, 0.75)
    assert (r["queryIdx"] == g["ratio_q"]).all() and (r["trainIdx"] == g["ratio_t"]).all()
    assert (r["distance"] == g["ratio_d"]).all()


@pytest.mark.parametrize("case", ["orb", "tie", "odd"])
def test_hamming_bit_exact(golden, orc, case):
    g = golden["hamming"]
    o = orc.knn2_hamming(g[case + "_q"], g[case + "_t"])
    assert (o["trainIdx"] == g[case + "_idx"]).all()
    assert (o["distance"] == g[case + "_dist"]).all()


@pytest.mark.parametrize("case", ["orb", "tie"])
def test_cross_check(golden, orc, case):
    g = golden["hamming"]
    q, t = g[case + "_q"], g[case + "_t"]
    o = orc.knn2_hamming(q, t)
    x = orc.cross_check(o, orc.col_best_hamming(q, t))
    assert (x["queryIdx"] == g[case + "_xq"]).all() and (x["trainIdx"] == g[case + "_xt"]).all()
    assert (x["distance"] == g[case + "_xd"]).all()


def test_cross_check_needs_only_the_marked_train_rows(orc):
    """The product's reverse pass runs over the train rows that are some query's best match only (pm_api.cu,
    pmk_cross_col_best).  The rule restated on the oracle: column minima computed over those rows alone, scattered into an
    otherwise empty column array, keep exactly the matches the full column pass keeps."""
    from points_matching_b200 import synth
    q, t = synth.orb_pair(300, 2500, seed=11)
    t[1200] = t[3]                                           # duplicate train rows: equal distances
    fwd = orc.knn2_hamming(q, t)
    full = orc.cross_check(fwd, orc.col_best_hamming(q, t))
    marked = np.unique(fwd["trainIdx"][:, 0])
    assert 0 < len(marked) <= 300
    col = np.full(t.shape[0], np.iinfo(np.uint64).max, dtype=np.uint64)
    col[marked] = orc.col_best_hamming(q, np.ascontiguousarray(t[marked]))
    part = orc.cross_check(fwd, col)
    assert len(part) == len(full) and (part == full).all()


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_8point_minimal_and_npoint(golden, orc):
    g = golden["fundamental"]
    p1, p2 = g["p1"], g["p2"]
    for idx, F in zip(g["idx8"], g["F8"]):
        assert _rel(orc.fm_8point(p1[idx], p2[idx]), F) < 1e-7
    gt = g["gt"]
    Fo = orc.fm_8point(p1[gt], p2[gt])
    assert _rel(Fo, g["F8_all"]) < 1e-10
    assert Fo[2, 2] == 1.0 and abs(np.linalg.svd(Fo, compute_uv=False)[2]) < 1e-12


def test_7point_solutions(golden, orc):
    g = golden["fundamental"]
    p1, p2 = g["p1"], g["p2"]
    for idx, F, n in zip(g["idx7"], g["F7"], g["n7"]):
        Fo = orc.fm_7point(p1[idx], p2[idx])
        assert len(Fo) == n
        for a in F[:n]:
            assert min(_rel(b, a) for b in Fo) < 1e-7


@pytest.mark.parametrize("name,method,thr", [("ransac1", 8, 1.0), ("ransac3", 8, 3.0),
                                            ("lmeds", 4, 3.0), ("fm7", 1, 3.0)])
def test_opencv_literal_estimator(golden, orc, name, method, thr):
    """cv::findFundamentalMat restated to the RNG stream: same F, same mask.  FM_7POINT with
    N>7 -- the reference's literal call at main.cpp:95-98 -- dispatches to LMedS (D4)."""
    g = golden["fundamental"]
    F, mask = orc.find_fundamental_cv(g["p1"], g["p2"], method, thr, 0.99, 1000)
    assert _rel(F, g[name + "_F"]) < 1e-8
    assert (mask == g[name + "_mask"]).all()
    if name == "fm7":
        assert (g["fm7_mask"] == g["lmeds_mask"]).all()


def test_dispatch_small_n(golden, orc):
    g = golden["fundamental"]
    assert bool(g["n6_none"])
    F, m = orc.find_fundamental_cv(g["p1"][:6], g["p2"][:6], 8)
    assert F is None
    inl = np.nonzero(g["gt"])[0][:7]
    F, m = orc.find_fundamental_cv(g["p1"][inl], g["p2"][inl], 8)
    assert F.shape == g["n7_F"].shape and (m == 1).all()


def test_dispatch_table_golden(golden, orc):
    """tests/golden/dispatch.npz (cv2 4.13): N == 7 -> all roots stacked whatever the method; FM_RANSAC with
    8 <= N < 15 -> LMedS; the parameter clamps.  The oracle follows cv::RNG's sample stream, so F and the masks are
    compared directly."""
    g = golden["dispatch"]
    p1, p2 = g["p1"], g["p2"]
    for idx, F, k in zip(g["idx7"], g["F7"], g["k7"]):
        for method in (8, 4, 2, 1):
            Fo, m = orc.find_fundamental_cv(p1[idx], p2[idx], method)
            assert Fo.shape == (3 * k, 3) and (m == 1).all()
            for a in F[: 3 * k].reshape(k, 3, 3):
                assert min(_rel(b, a) for b in Fo.reshape(k, 3, 3)) < 1e-7
    sub = g["sub"]
    for n in (8, 11, 14):
        # below 14 points the median (sorted error n/2) belongs to one of the seven sample points themselves, i.e. it is
        # rounding noise and so is the winner: only the dispatch (RANSAC == LMedS == 7POINT) is comparable there
        ref = orc.find_fundamental_cv(p1[sub[:n]], p2[sub[:n]], 4, 1.0, 0.99)
        for method in (8, 4, 1):
            Fo, m = orc.find_fundamental_cv(p1[sub[:n]], p2[sub[:n]], method, 1.0, 0.99)
            assert np.array_equal(Fo, ref[0]) and np.array_equal(m, ref[1]), (n, method)
            if n == 14:
                assert _rel(Fo, g[f"n{n}_F"]) < 1e-7 and (m == g[f"n{n}_mask"]).all(), (n, method)
    mix = g["mix"]
    Fo, m = orc.find_fundamental_cv(p1[mix], p2[mix], 8, 1.0, 0.99)
    assert _rel(Fo, g["n40_ransac_F"]) < 1e-7 and (m == g["n40_ransac_mask"]).all()
    Fo, m = orc.find_fundamental_cv(p1[mix], p2[mix], 4, 1.0, 0.99)
    assert _rel(Fo, g["n40_lmeds_F"]) < 1e-7 and (m == g["n40_lmeds_mask"]).all()
    for prm1, prm2 in ((0.0, 0.99), (3.0, 0.99), (3.0, 1.5), (-1.0, 0.0)):
        Fo, m = orc.find_fundamental_cv(p1[mix], p2[mix], 8, prm1, prm2)
        assert _rel(Fo, g["n40_thr3_F"]) < 1e-7 and (m == g["n40_thr3_mask"]).all(), (prm1, prm2)


def test_sampson_and_epilines(golden, orc):
    g = golden["fundamental"]
    F = g["ransac1_F"]
    s = orc.sampson_f64(F, g["p1"], g["p2"])
    assert np.allclose(s, g["sampson"], rtol=1e-12, atol=0)
    assert np.allclose(orc.epilines(g["p1"], 1, F), g["lines1"], rtol=0, atol=1e-6)
    assert np.allclose(orc.epilines(g["p2"], 2, F), g["lines2"], rtol=0, atol=1e-6)


def test_ransac_mask_is_symepi_threshold(golden, orc):
    """the cv2 RANSAC mask == (max symmetric epipolar distance <= thr^2)."""
    g = golden["fundamental"]
    e = orc.symepi_f64(g["ransac1_F"], g["p1"], g["p2"]).astype(np.float32)
    assert ((e <= np.float32(1.0)) == g["ransac1_mask"].astype(bool)).all()


def test_f32_scoring_matches_f64_away_from_threshold(golden, orc):
    g = golden["fundamental"]
    F = g["ransac1_F"]
    s = orc.sampson_f64(F, g["p1"], g["p2"])
    c, mask = orc.count_inliers_f32(F.astype(np.float32), g["p1"], g["p2"], 1.0, orc.METRIC_SAMPSON, True)
    clear = np.abs(s - 1.0) > 1e-3
    assert ((s <= 1.0)[clear] == mask.astype(bool)[clear]).all() and c == mask.sum()
    e = orc.symepi_f64(F, g["p1"], g["p2"])
    c, mask = orc.count_inliers_f32(F.astype(np.float32), g["p1"], g["p2"], 1.0, orc.METRIC_SYMEPI, True)
    clear = np.abs(e - 1.0) > 1e-3
    assert ((e <= 1.0)[clear] == mask.astype(bool)[clear]).all()


def test_ransac_given_samples_recovers_motion(orc):
    p1, p2, gt = synth.correspondences(2000, seed=3)
    idx = synth.sample_index_sets(2000, 4000, 8, seed=5)
    r = orc.ransac_f(p1, p2, idx, orc.METRIC_SAMPSON, 1.0, refit=True, want_models=True)
    assert r is not None
    assert r["n_inliers"] == r["mask"].sum() == r["counts"].max()
    assert r["best_model"] == int(np.argmax(r["counts"]))          # lowest id among ties
    # nearly all true inliers recovered, few outliers accepted
    assert r["mask"][gt].mean() > 0.8 and r["mask"][~gt].mean() < 0.05
    s = orc.sampson_f64(r["F"], p1[gt], p2[gt])
    assert np.mean(s) < 0.25          # refit F: mean Sampson on GT inliers ~ noise level
    # per-hypothesis counts are reproducible from the returned FP32 models
    for h in (0, 17, 3999):
        assert orc.count_inliers_f32(r["models"][h, 0], p1, p2, 1.0) == r["counts"][h]


def test_ransac_7pt_models(orc):
    p1, p2, gt = synth.correspondences(800, seed=4)
    idx = synth.sample_index_sets(800, 300, 7, seed=6)
    r = orc.ransac_f(p1, p2, idx, orc.METRIC_SYMEPI, 1.5, refit=False, want_models=True)
    assert r is not None and r["mask"][gt].mean() > 0.6
    h, k = divmod(r["best_model"], 3)
    assert orc.count_inliers_f32(r["models"][h, k], p1, p2, 1.5, orc.METRIC_SYMEPI) == r["n_inliers"]


def test_image_pair_config1(golden, orc):
    """Config 1 stand-in (img01/img02.JPG, SIFT for SURF): north_star flow and the reference's
    literal flow (main.cpp:46 k=1 match, :49-69 min/max filter, :95-98 FM_7POINT -> LMedS)."""
    g = golden["image_pair"]
    d1, d2 = g["desc1"].astype(np.float32), g["desc2"].astype(np.float32)
    o = orc.knn2_l2(d1, d2)
    assert (o["trainIdx"] == g["knn_idx"]).all() and (o["distance"] == g["knn_dist"]).all()
    r = orc.ratio_filter(o, 0.75)
    assert (r["queryIdx"] == g["ratio_q"]).all() and (r["trainIdx"] == g["ratio_t"]).all()
    pts1 = orc.gather_points(g["kp1"], r["queryIdx"])
    pts2 = orc.gather_points(g["kp2"], r["trainIdx"])
    assert (pts1 == g["kp1"][g["ratio_q"]]).all()
    F, mask = orc.find_fundamental_cv(pts1, pts2, 8, 1.0, 0.99, 1000)
    assert _rel(F, g["ransac_F"]) < 1e-6 and (mask == g["ransac_mask"]).all()
    lit, mn, mx = orc.minmax_filter(np.ascontiguousarray(o[:, 0]))
    assert mn == g["lit_min"] and mx == g["lit_max"]
    assert (lit["queryIdx"] == g["lit_q"]).all() and (lit["trainIdx"] == g["lit_t"]).all()
    F, mask = orc.find_fundamental_cv(g["kp1"][g["lit_q"]], g["kp2"][g["lit_t"]], 1)
    assert _rel(F, g["lit_F"]) < 1e-6 and (mask == g["lit_mask"]).all()


def test_lmeds_given_samples(orc):
    """orc_lmeds_f (the estimator behind main.cpp:95-98's CV_FM_7POINT with N > 7, on caller-supplied samples):
    medians equal a plain numpy restatement, the winner has the smallest median, the mask follows OpenCV's
    sigma rule, and the planted motion is recovered."""
    from points_matching_b200 import synth
    p1, p2, gt = synth.correspondences(400, seed=3, outlier_frac=0.3)
    idx = synth.sample_index_sets(400, 60, 7, seed=8)
    r = orc.lmeds_f(p1, p2, idx)
    assert r is not None
    med, models = r["medians"], r["models"]
    finite = np.isfinite(med)
    assert finite.sum() >= 60
    for m in np.nonzero(finite)[0][:25]:
        e = orc.symepi_f64(models[m].astype(np.float64), p1, p2).astype(np.float32)
        assert np.sort(e)[400 // 2] == med[m]
    best = int(np.argmin(np.where(finite, med, np.inf)))
    assert r["best_model"] == best
    sigma = max(2.5 * 1.4826 * (1 + 5.0 / (400 - 7)) * np.sqrt(float(med[best])), 0.001)
    e = orc.symepi_f64(models[best].astype(np.float64), p1, p2).astype(np.float32)
    assert np.array_equal(r["mask"], (e <= np.float32(sigma * sigma)).astype(np.uint8))
    assert r["mask"][gt].mean() > 0.9 and r["mask"][~gt].mean() < 0.1
    # given models reproduce the same answer
    r2 = orc.lmeds_f(p1, p2, idx, models=models)
    assert r2["best_model"] == best and np.array_equal(r2["mask"], r["mask"])
