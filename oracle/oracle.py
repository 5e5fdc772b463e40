"""ctypes binding of the CPU oracle (oracle/pm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpm_oracle.so")

DMATCH = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
METRIC_SAMPSON, METRIC_SYMEPI = 0, 1


def build(force=False):
    src = os.path.join(_HERE, "pm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_l2sq_f32_rerank.restype = C.c_float
        _lib.orc_sampson_f64.restype = C.c_double
        _lib.orc_symepi_f64.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def ncores():
    return os.cpu_count() or 1


def knn2_l2(q, t, nthreads=0):
    q, t = _f32(q), _f32(t)
    out = np.zeros((q.shape[0], 2), dtype=DMATCH)
    dim = q.shape[1] if q.ndim == 2 else t.shape[1]
    lib().orc_knn2_l2_f32(_p(q), q.shape[0], _p(t), t.shape[0], dim, _p(out), nthreads or ncores())
    return out


def l2sq_rerank(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_l2sq_f32_rerank(_p(a), _p(b), a.shape[0]))


def knn2_hamming(q, t, nthreads=0):
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    out = np.zeros((q.shape[0], 2), dtype=DMATCH)
    nb = q.shape[1] if q.ndim == 2 and q.shape[0] else t.shape[1]
    lib().orc_knn2_hamming(_p(q), q.shape[0], _p(t), t.shape[0], nb, _p(out), nthreads or ncores())
    return out


def col_best_hamming(q, t, nthreads=0):
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    out = np.zeros(t.shape[0], dtype=np.uint64)
    lib().orc_col_best_hamming(_p(q), q.shape[0], _p(t), t.shape[0], t.shape[1], _p(out), nthreads or ncores())
    return out


def ratio_filter(knn, ratio):
    knn = np.ascontiguousarray(knn)
    out = np.zeros(knn.shape[0], dtype=DMATCH)
    n = lib().orc_ratio_filter(_p(knn), knn.shape[0], C.c_float(ratio), _p(out))
    return out[:n]


def cross_check(knn, col_best):
    """knn: [nq] or [nq,2] DMatch; col_best: [nt] u64 packed column minima."""
    knn = np.ascontiguousarray(knn)
    stride = 1 if knn.ndim == 1 else knn.shape[1]
    col_best = np.ascontiguousarray(col_best, dtype=np.uint64)
    out = np.zeros(knn.shape[0], dtype=DMATCH)
    n = lib().orc_cross_check(_p(knn), knn.shape[0], stride, _p(col_best), col_best.shape[0], _p(out))
    return out[:n]


def minmax_filter(m):
    m = np.ascontiguousarray(m)
    out = np.zeros(m.shape[0], dtype=DMATCH)
    mn, mx = C.c_double(), C.c_double()
    n = lib().orc_minmax_filter(_p(m), m.shape[0], _p(out), C.byref(mn), C.byref(mx))
    return out[:n], mn.value, mx.value


def gather_points(kp_xy, idx):
    kp_xy = _f32(kp_xy)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    out = np.zeros((idx.shape[0], 2), dtype=np.float32)
    lib().orc_gather_points(_p(kp_xy), _p(idx), idx.shape[0], _p(out))
    return out


def fm_8point(p1, p2):
    p1, p2 = _f32(p1), _f32(p2)
    F = np.zeros(9)
    ok = lib().orc_fm_8point(_p(p1), _p(p2), p1.shape[0], _p(F))
    return F.reshape(3, 3) if ok else None


def fm_7point(p1, p2):
    p1, p2 = _f32(p1), _f32(p2)
    F = np.zeros(27)
    n = lib().orc_fm_7point(_p(p1), _p(p2), _p(F))
    return F[: 9 * n].reshape(n, 3, 3)


def sampson_f64(F, p1, p2):
    F = np.ascontiguousarray(F, dtype=np.float64).reshape(9)
    p1, p2 = np.asarray(p1, dtype=np.float64), np.asarray(p2, dtype=np.float64)
    f = lib().orc_sampson_f64
    return np.array([f(_p(F), C.c_double(a[0]), C.c_double(a[1]), C.c_double(b[0]), C.c_double(b[1]))
                     for a, b in zip(p1, p2)])


def symepi_f64(F, p1, p2):
    F = np.ascontiguousarray(F, dtype=np.float64).reshape(9)
    p1, p2 = np.asarray(p1, dtype=np.float64), np.asarray(p2, dtype=np.float64)
    f = lib().orc_symepi_f64
    return np.array([f(_p(F), C.c_double(a[0]), C.c_double(a[1]), C.c_double(b[0]), C.c_double(b[1]))
                     for a, b in zip(p1, p2)])


def count_inliers_f32(F32, p1, p2, thr, metric=METRIC_SAMPSON, want_mask=False):
    F32 = _f32(F32).reshape(9)
    p1, p2 = _f32(p1), _f32(p2)
    mask = np.zeros(p1.shape[0], dtype=np.uint8) if want_mask else None
    c = lib().orc_count_inliers_f32(_p(F32), _p(p1), _p(p2), p1.shape[0], C.c_float(thr), metric,
                                    _p(mask) if want_mask else None)
    return (c, mask) if want_mask else c


def ransac_f(p1, p2, sample_idx, metric=METRIC_SAMPSON, thr=1.0, refit=True, nthreads=0,
             want_models=False):
    """RANSAC over caller-supplied minimal samples.  Returns a dict or None."""
    p1, p2 = _f32(p1), _f32(p2)
    idx = np.ascontiguousarray(sample_idx, dtype=np.int32)
    nhyp, m = idx.shape
    per = 1 if m == 8 else 3
    F = np.zeros(9)
    mask = np.zeros(p1.shape[0], dtype=np.uint8)
    ninl = C.c_int(0)
    best = C.c_int64(-1)
    counts = np.zeros(nhyp, dtype=np.int32)
    Fs = np.zeros((nhyp, per, 9), dtype=np.float32) if want_models else None
    ok = lib().orc_ransac_f(_p(p1), _p(p2), p1.shape[0], _p(idx), nhyp, m, metric, C.c_float(thr),
                            int(bool(refit)), _p(F), _p(mask), C.byref(ninl), C.byref(best),
                            _p(counts), _p(Fs) if want_models else None, nthreads or ncores())
    if not ok:
        return None
    return dict(F=F.reshape(3, 3), mask=mask, n_inliers=ninl.value, best_model=best.value,
                counts=counts, models=Fs)


def lmeds_f(p1, p2, sample_idx, models=None):
    """LMedS over caller-supplied 7-point samples (or over given f32 models [nhyp*3, 9]).  Returns a dict or None."""
    p1, p2 = _f32(p1), _f32(p2)
    idx = np.ascontiguousarray(sample_idx, dtype=np.int32)
    nhyp = idx.shape[0]
    F = np.zeros(9)
    mask = np.zeros(p1.shape[0], dtype=np.uint8)
    ninl, best = C.c_int(0), C.c_int64(-1)
    med = np.zeros(nhyp * 3, dtype=np.float32)
    Fs = np.zeros((nhyp * 3, 9), dtype=np.float32) if models is None else np.ascontiguousarray(models, dtype=np.float32)
    ok = lib().orc_lmeds_f(_p(p1), _p(p2), p1.shape[0], _p(idx), nhyp, _p(F), _p(mask), C.byref(ninl), C.byref(best),
                           _p(med), _p(Fs), int(models is not None))
    if not ok:
        return None
    return dict(F=F.reshape(3, 3), mask=mask, n_inliers=ninl.value, best_model=best.value, medians=med, models=Fs)


def find_fundamental_cv(p1, p2, method, param1=3.0, param2=0.99, max_iters=1000):
    """OpenCV-literal dispatch (FM_7POINT=1, FM_8POINT=2, FM_LMEDS=4, FM_RANSAC=8)."""
    p1, p2 = _f32(p1), _f32(p2)
    F = np.zeros(27)
    mask = np.zeros(p1.shape[0], dtype=np.uint8)
    n = lib().orc_find_fundamental_cv(_p(p1), _p(p2), p1.shape[0], method, C.c_double(param1),
                                      C.c_double(param2), max_iters, _p(F), _p(mask))
    if n == 0:
        return None, None
    return F[: 9 * n].reshape(3 * n, 3), mask


def epilines(pts, which, F):
    pts = _f32(pts)
    F = np.ascontiguousarray(F, dtype=np.float64).reshape(9)
    out = np.zeros((pts.shape[0], 3), dtype=np.float32)
    lib().orc_epilines(_p(pts), pts.shape[0], which, _p(F), _p(out))
    return out
