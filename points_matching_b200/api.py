"""Host-side mirror of the reference's OpenCV call surface, over the C ABI (include/pm.h).

Names, argument meaning and error behaviour follow the calls made by
/root/reference/Points Matching/main.cpp:
    main.cpp:43-46   BruteForceMatcher<L2<float>> / matcher.match   -> BFMatcher
    main.cpp:49-69   good-match filter                               -> minmax_filter / ratio_test
    main.cpp:89-91   KeyPoint::convert(kps, pts, idx)                -> keypoints_convert
    main.cpp:95-98   findFundamentalMat                              -> findFundamentalMat
    main.cpp:127-132 computeCorrespondEpilines                       -> computeCorrespondEpilines
Matches come back as numpy structured arrays with cv::DMatch's layout (DMATCH) instead of
lists of DMatch objects.  OpenCV's exceptions map to PMError(PM_BAD_ARG); its "empty Mat"
results map to None.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import COMM_ID_BYTES, DMATCH, PAIR_RESULT, PM_EMPTY, PM_OK, FmOptions, RansacParams  # noqa: F401

NORM_L2 = 4          # cv::NORM_L2
NORM_HAMMING = 6     # cv::NORM_HAMMING
FM_7POINT, FM_8POINT, FM_LMEDS, FM_RANSAC = 1, 2, 4, 8
METRIC_SAMPSON, METRIC_SYMEPI = 0, 1


class PMError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libpm status {status}: {msg}")
        self.status = status


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One pm_ctx (device, stream, workspace).  Not thread-safe; one per host thread."""

    def __init__(self, device=0):
        self._L = _lib.lib()
        h = C.c_void_p()
        st = self._L.pm_create(C.byref(h), int(device))
        if st != PM_OK:
            raise PMError(st, "pm_create failed (no sm_100 CUDA device visible? libpm has no CPU fallback)")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.pm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, st, allow_empty=False):
        if st == PM_OK or (allow_empty and st == PM_EMPTY):
            return st
        raise PMError(st, (self._L.pm_last_error(self._h) or b"").decode())

    # ---- plumbing -------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        self._chk(self._L.pm_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def sync(self):
        self._chk(self._L.pm_sync(self._h))

    def set_pipelining(self, on):
        """Opt-in overlap of consecutive knn2_ratio_l2_*_dev calls (see pm_set_pipelining in include/pm.h)."""
        self._chk(self._L.pm_set_pipelining(self._h, 1 if on else 0))

    def launch_count(self):
        return int(self._L.pm_launch_count(self._h))

    def l2_stats(self):
        out = (C.c_int32 * 4)()
        self._chk(self._L.pm_l2_stats(self._h, out))
        return dict(exact_mode=bool(out[0]), fallback_rows=out[1], k_blocks=out[2], segments=out[3])

    def profile_enable(self, on=True):
        self._chk(self._L.pm_profile_enable(self._h, int(bool(on))))

    def profile_read(self, which):
        """(total_ms, launches) of kernel class `which` (0 L2 tensor-core, 1 Hamming, 2 RANSAC scoring)."""
        ms, n = C.c_double(0), C.c_int(0)
        self._chk(self._L.pm_profile_read(self._h, which, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def knn2_ratio_l2_ptr(self, q_ptr, nq, t_ptr, nt, dim, ratio, knn_ptr, good_ptr):
        """pm_knn2_ratio_l2_f32 on raw HOST pointers (pinned buffers in bench.py's e2e leg)."""
        n = C.c_int(0)
        self._chk(self._L.pm_knn2_ratio_l2_f32(self._h, C.c_void_p(q_ptr), nq, C.c_void_p(t_ptr), nt, dim,
                                               C.c_float(ratio), C.c_void_p(knn_ptr), C.c_void_p(good_ptr),
                                               C.byref(n)))
        return n.value

    def find_fundamental_ptr(self, p1_ptr, p2_ptr, n, sample_ptr, n_hyp, sample_size, metric, threshold, refit,
                             F_ptr, mask_ptr):
        """pm_find_fundamental on raw HOST pointers; returns n_inliers or None (PM_EMPTY)."""
        prm = RansacParams()
        prm.sample_size, prm.metric, prm.threshold = sample_size, metric, threshold
        prm.n_hyp, prm.refit, prm.sample_idx = n_hyp, int(bool(refit)), sample_ptr
        ninl = C.c_int(0)
        st = self._chk(self._L.pm_find_fundamental(self._h, C.c_void_p(p1_ptr), C.c_void_p(p2_ptr), n, C.byref(prm),
                                                   C.c_void_p(F_ptr), C.c_void_p(mask_ptr), C.byref(ninl)),
                       allow_empty=True)
        return None if st == PM_EMPTY else ninl.value

    def knn2_ratio(self, query, train, ratio=0.75):
        """(knn [nq,2], good [n]) == knnMatch(k=2) followed by the Lowe ratio test, one call."""
        q = np.ascontiguousarray(query, dtype=np.float32)
        t = np.ascontiguousarray(train, dtype=np.float32)
        nq, nt = q.shape[0], t.shape[0]
        knn = np.zeros((nq, 2), dtype=DMATCH)
        good = np.zeros(nq, dtype=DMATCH)
        if nq == 0:
            return knn, good
        n = self.knn2_ratio_l2_ptr(q.ctypes.data, nq, t.ctypes.data, nt, q.shape[1], ratio, knn.ctypes.data,
                                   good.ctypes.data)
        return knn, good[:n]

    # ---- host-buffer calls ------------------------------------------------------------
    def knn2(self, query, train, norm=NORM_L2):
        """[nq, 2] DMATCH array == BFMatcher(norm).knnMatch(query, train, k=2)."""
        q, t = np.ascontiguousarray(query), np.ascontiguousarray(train)
        if q.ndim != 2 or t.ndim != 2:
            raise PMError(_lib.PM_BAD_ARG, "descriptors must be 2-d arrays")
        nq, nt = q.shape[0], t.shape[0]
        if nq and nt and q.shape[1] != t.shape[1]:
            raise PMError(_lib.PM_BAD_ARG, "query/train descriptor width mismatch")
        if q.dtype != t.dtype:
            raise PMError(_lib.PM_BAD_ARG, "(-215) queryDescriptors.type() == trainDescType")
        width = q.shape[1] if nq else t.shape[1]
        out = np.zeros((nq, 2), dtype=DMATCH)
        if nq == 0:
            return out
        if norm == NORM_L2:
            if q.dtype == np.float32:
                fn = self._L.pm_knn2_l2_f32
            elif q.dtype == np.uint8:
                fn = self._L.pm_knn2_l2_u8
            else:
                raise PMError(_lib.PM_BAD_ARG, "NORM_L2 takes float32 (or uint8 SIFT) descriptors")
        elif norm == NORM_HAMMING:
            if q.dtype != np.uint8:
                raise PMError(_lib.PM_BAD_ARG, "NORM_HAMMING takes uint8 descriptors")
            fn = self._L.pm_knn2_hamming
        else:
            raise PMError(_lib.PM_BAD_ARG, "normType must be NORM_L2 or NORM_HAMMING")
        self._chk(fn(self._h, _p(q), nq, _p(t), nt, width, _p(out)))
        return out

    def ratio_filter(self, knn, ratio=0.75):
        knn = np.ascontiguousarray(knn, dtype=DMATCH)
        nq = knn.shape[0]
        out = np.zeros(nq, dtype=DMATCH)
        n = C.c_int(0)
        self._chk(self._L.pm_ratio_filter(self._h, _p(knn), nq, C.c_float(ratio), _p(out), C.byref(n)))
        return out[: n.value]

    def minmax_filter(self, matches):
        m = np.ascontiguousarray(matches, dtype=DMATCH)
        stride = 1 if m.ndim == 1 else m.shape[1]
        n = m.shape[0]
        out = np.zeros(n, dtype=DMATCH)
        cnt, mn, mx = C.c_int(0), C.c_double(0), C.c_double(0)
        self._chk(self._L.pm_minmax_filter(self._h, _p(m), n, stride, _p(out), C.byref(cnt), C.byref(mn), C.byref(mx)))
        return out[: cnt.value], mn.value, mx.value

    def match_cross(self, query, train, norm=NORM_L2):
        q, t = np.ascontiguousarray(query), np.ascontiguousarray(train)
        if q.dtype != t.dtype:
            raise PMError(_lib.PM_BAD_ARG, "(-215) queryDescriptors.type() == trainDescType")
        nq, nt = q.shape[0], t.shape[0]
        out = np.zeros(nq, dtype=DMATCH)
        n = C.c_int(0)
        if nq == 0 or nt == 0:
            return out[:0]
        if norm == NORM_L2:
            if q.dtype != np.float32:
                raise PMError(_lib.PM_BAD_ARG, "cross-check NORM_L2 takes float32 descriptors")
            fn = self._L.pm_match_cross_l2_f32
        else:
            if q.dtype != np.uint8:
                raise PMError(_lib.PM_BAD_ARG, "NORM_HAMMING takes uint8 descriptors")
            fn = self._L.pm_match_cross_hamming
        self._chk(fn(self._h, _p(q), nq, _p(t), nt, q.shape[1], _p(out), C.byref(n)))
        return out[: n.value]

    def gather_points(self, kp_xy, idx):
        kp = np.ascontiguousarray(kp_xy, dtype=np.float32).reshape(-1, 2)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        out = np.zeros((idx.shape[0], 2), dtype=np.float32)
        self._chk(self._L.pm_gather_points(self._h, _p(kp), kp.shape[0], _p(idx), idx.shape[0], _p(out)))
        return out

    def find_fundamental(self, p1, p2, sample_size=8, metric=METRIC_SAMPSON, threshold=1.0, n_hyp=4096,
                         refit=True, sample_idx=None, seed=0):
        """RANSAC over minimal samples.  Returns (F[3,3] f64, mask[n] u8, n_inliers) or None."""
        p1 = np.ascontiguousarray(p1, dtype=np.float32).reshape(-1, 2)
        p2 = np.ascontiguousarray(p2, dtype=np.float32).reshape(-1, 2)
        if p1.shape != p2.shape:
            raise PMError(_lib.PM_BAD_ARG, "points1/points2 size mismatch")
        n = p1.shape[0]
        prm = RansacParams()
        prm.sample_size, prm.metric, prm.threshold = sample_size, metric, threshold
        prm.refit, prm.seed = int(bool(refit)), seed
        keep = None
        if sample_idx is not None:
            keep = np.ascontiguousarray(sample_idx, dtype=np.int32)
            if keep.ndim != 2 or keep.shape[1] != sample_size:
                raise PMError(_lib.PM_BAD_ARG, "sample_idx must be [n_hyp, sample_size]")
            prm.n_hyp = keep.shape[0]
            prm.sample_idx = keep.ctypes.data
        else:
            prm.n_hyp = n_hyp
            prm.sample_idx = None
        F = np.zeros(9, dtype=np.float64)
        mask = np.zeros(n, dtype=np.uint8)
        ninl = C.c_int(0)
        st = self._chk(self._L.pm_find_fundamental(self._h, _p(p1), _p(p2), n, C.byref(prm), _p(F), _p(mask),
                                                   C.byref(ninl)), allow_empty=True)
        if st == PM_EMPTY:
            return None
        return F.reshape(3, 3), mask, ninl.value

    def find_fundamental_lmeds(self, p1, p2, n_hyp=512, sample_idx=None, seed=0):
        """LMedS over 7-point samples (the estimator behind the reference's CV_FM_7POINT call when N > 7).
        Returns (F[3,3] f64, mask[n] u8, n_inliers, median) or None."""
        p1 = np.ascontiguousarray(p1, dtype=np.float32).reshape(-1, 2)
        p2 = np.ascontiguousarray(p2, dtype=np.float32).reshape(-1, 2)
        n = p1.shape[0]
        keep = None
        if sample_idx is not None:
            keep = np.ascontiguousarray(sample_idx, dtype=np.int32)
            if keep.ndim != 2 or keep.shape[1] != 7:
                raise PMError(_lib.PM_BAD_ARG, "sample_idx must be [n_hyp, 7]")
            n_hyp = keep.shape[0]
        F = np.zeros(9, dtype=np.float64)
        mask = np.zeros(n, dtype=np.uint8)
        ninl, med = C.c_int(0), C.c_float(0)
        st = self._chk(self._L.pm_find_fundamental_lmeds(self._h, _p(p1), _p(p2), n, n_hyp, _p(keep), C.c_uint64(seed),
                                                         _p(F), _p(mask), C.byref(ninl), C.byref(med)), allow_empty=True)
        if st == PM_EMPTY:
            return None
        return F.reshape(3, 3), mask, ninl.value, med.value

    def lmeds_score_dev(self, dp1, dp2, n, dF32, n_models, dmedians):
        self._chk(self._L.pm_lmeds_score_dev(self._h, C.c_void_p(dp1), C.c_void_p(dp2), n, C.c_void_p(dF32), n_models,
                                             C.c_void_p(dmedians)))

    def fundamental_8point(self, p1, p2):
        p1 = np.ascontiguousarray(p1, dtype=np.float32).reshape(-1, 2)
        p2 = np.ascontiguousarray(p2, dtype=np.float32).reshape(-1, 2)
        F = np.zeros(9, dtype=np.float64)
        st = self._chk(self._L.pm_fundamental_8point(self._h, _p(p1), _p(p2), p1.shape[0], _p(F)), allow_empty=True)
        return None if st == PM_EMPTY else F.reshape(3, 3)

    def epilines(self, pts, which_image, F):
        pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 2)
        F = np.ascontiguousarray(F, dtype=np.float64).reshape(9)
        out = np.zeros((pts.shape[0], 3), dtype=np.float32)
        self._chk(self._L.pm_epilines(self._h, _p(pts), pts.shape[0], int(which_image), _p(F), _p(out)))
        return out

    def residuals(self, p1, p2, F, metric=METRIC_SAMPSON):
        p1 = np.ascontiguousarray(p1, dtype=np.float32).reshape(-1, 2)
        p2 = np.ascontiguousarray(p2, dtype=np.float32).reshape(-1, 2)
        F = np.ascontiguousarray(F, dtype=np.float64).reshape(9)
        out = np.zeros(p1.shape[0], dtype=np.float32)
        mean = C.c_double(0)
        self._chk(self._L.pm_residuals(self._h, _p(p1), _p(p2), p1.shape[0], _p(F), metric, _p(out), C.byref(mean)))
        return out, mean.value

    def find_fundamental_adaptive(self, p1, p2, sample_size=7, metric=METRIC_SYMEPI, threshold=3.0, confidence=0.99,
                                  max_iters=1000, batch=1024, refit=False, seed=0):
        """RANSAC with OpenCV's adaptive stop, batches generated / solved / scored on the device; only the 8-byte
        winner key returns per batch.  Returns (F[3,3], mask, n_inliers, hypotheses_run) or None."""
        p1 = np.ascontiguousarray(p1, dtype=np.float32).reshape(-1, 2)
        p2 = np.ascontiguousarray(p2, dtype=np.float32).reshape(-1, 2)
        if p1.shape != p2.shape:
            raise PMError(_lib.PM_BAD_ARG, "points1/points2 size mismatch")
        n = p1.shape[0]
        prm = RansacParams()
        prm.sample_size, prm.metric, prm.threshold, prm.n_hyp = sample_size, metric, threshold, batch
        prm.refit, prm.seed, prm.max_iters, prm.confidence, prm.sample_idx = int(bool(refit)), seed, max_iters, confidence, None
        F, mask = np.zeros(9, dtype=np.float64), np.zeros(n, dtype=np.uint8)
        ninl, run = C.c_int(0), C.c_int(0)
        st = self._chk(self._L.pm_find_fundamental_adaptive(self._h, _p(p1), _p(p2), n, C.byref(prm), _p(F), _p(mask),
                                                            C.byref(ninl), C.byref(run)), allow_empty=True)
        if st == PM_EMPTY:
            return None
        return F.reshape(3, 3), mask, ninl.value, run.value

    def find_fundamental_mat(self, p1, p2, method, param1=3.0, param2=0.99, max_iters=1000, options=None):
        """pm_find_fundamental_mat: cv::findFundamentalMat's dispatch table.  Returns (F[3k,3] f64, mask[n] u8) with
        k = 1..3 stacked models (k > 1 only for n == 7), or (None, None) for OpenCV's empty Mat."""
        p1 = np.ascontiguousarray(p1, dtype=np.float32).reshape(-1, 2)
        p2 = np.ascontiguousarray(p2, dtype=np.float32).reshape(-1, 2)
        if p1.shape != p2.shape:
            raise PMError(_lib.PM_BAD_ARG, "(-215) points1/points2 count mismatch")
        n = p1.shape[0]
        F, mask, k = np.zeros(27, dtype=np.float64), np.zeros(max(n, 1), dtype=np.uint8), C.c_int(0)
        st = self._chk(self._L.pm_find_fundamental_mat(self._h, _p(p1), _p(p2), n, int(method), C.c_double(param1),
                                                       C.c_double(param2), int(max_iters),
                                                       C.byref(options) if options is not None else None, _p(F), C.byref(k),
                                                       _p(mask)), allow_empty=True)
        if st == PM_EMPTY:
            return None, None
        return F[: 9 * k.value].reshape(3 * k.value, 3).copy(), mask[:n]

    def fundamental_7point(self, p1, p2):
        """run7Point on exactly seven correspondences: [k, 3, 3] f64 with k = 1..3 real roots, or None."""
        p1 = np.ascontiguousarray(p1, dtype=np.float32).reshape(-1, 2)
        p2 = np.ascontiguousarray(p2, dtype=np.float32).reshape(-1, 2)
        F, k = np.zeros(27, dtype=np.float64), C.c_int(0)
        st = self._chk(self._L.pm_fundamental_7point(self._h, _p(p1), _p(p2), p1.shape[0], _p(F), C.byref(k)), allow_empty=True)
        return None if st == PM_EMPTY else F[: 9 * k.value].reshape(k.value, 3, 3).copy()

    def measure_peak(self, which):
        """Measured issue peak: which 0 -> FP32 FFMA TFLOP/s, 1 -> 1e12 POPC.32 per second."""
        v = C.c_double(0)
        self._chk(self._L.pm_measure_peak(self._h, int(which), C.byref(v)))
        return v.value

    # ---- multi-GPU (NCCL inside the C ABI) ---------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * COMM_ID_BYTES)()
        st = _lib.lib().pm_comm_unique_id(buf)
        if st != PM_OK:
            raise PMError(st, "pm_comm_unique_id: libnccl.so.2 not loadable")
        return bytes(buf)

    def comm_init(self, n_ranks, rank, unique_id):
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        self._chk(self._L.pm_comm_init(self._h, n_ranks, rank, buf))

    def comm_init_from_torch(self, group=None):
        """Creates this ctx's own NCCL communicator over the ranks of a torch.distributed group (the unique id travels
        through the group); a no-op for a single process."""
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ids = [self.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.comm_init(world, rank, ids[0])

    def comm_info(self):
        n, r = C.c_int(1), C.c_int(0)
        self._chk(self._L.pm_comm_info(self._h, C.byref(n), C.byref(r)))
        return n.value, r.value

    def match_cross_sharded_dev(self, dq, nq, dt, nt, width, norm, q_index_base, dknn, dcol, dout, dn_out):
        self._chk(self._L.pm_match_cross_sharded_dev(self._h, C.c_void_p(dq), nq, C.c_void_p(dt), nt, width, int(norm),
                                                     q_index_base, C.c_void_p(dknn), C.c_void_p(dcol), C.c_void_p(dout),
                                                     C.c_void_p(dn_out)))

    def allgather_matches_dev(self, dlocal, dn_local, max_per_rank, dall, dcounts):
        self._chk(self._L.pm_allgather_matches_dev(self._h, C.c_void_p(dlocal), C.c_void_p(dn_local), max_per_rank,
                                                   C.c_void_p(dall), C.c_void_p(dcounts)))

    def find_fundamental_sharded_dev(self, dp1, dp2, n, dsamples_full, n_hyp_total, hyp_lo, n_hyp, sample_size, metric,
                                     threshold, refit, dF, dmask, dn_inl, dkey, seed=0):
        prm = RansacParams()
        prm.sample_size, prm.metric, prm.threshold = sample_size, metric, threshold
        prm.n_hyp, prm.refit, prm.sample_idx, prm.hyp_id_base, prm.seed = n_hyp, int(bool(refit)), dsamples_full or None, hyp_lo, seed
        self._chk(self._L.pm_find_fundamental_sharded_dev(self._h, C.c_void_p(dp1), C.c_void_p(dp2), n, C.byref(prm),
                                                          n_hyp_total, C.c_void_p(dF), C.c_void_p(dmask), C.c_void_p(dn_inl),
                                                          C.c_void_p(dkey)))

    # ---- device-resident calls (raw device pointers, e.g. torch_tensor.data_ptr()) ------
    def knn2_l2_f32_dev(self, dq, nq, dt, nt, dim, dout, q_index_base=0):
        self._chk(self._L.pm_knn2_l2_f32_dev(self._h, C.c_void_p(dq), nq, C.c_void_p(dt), nt, dim, q_index_base,
                                             C.c_void_p(dout)))

    def knn2_ratio_l2_f32_dev(self, dq, nq, dt, nt, dim, ratio, dknn, dgood, dn_good, q_index_base=0):
        """kNN-2 + ratio test enqueued as one chain (pack -> GEMM/top-2 -> re-rank -> filter)."""
        self._chk(self._L.pm_knn2_ratio_l2_f32_dev(self._h, C.c_void_p(dq), nq, C.c_void_p(dt), nt, dim, C.c_float(ratio),
                                                   q_index_base, C.c_void_p(dknn), C.c_void_p(dgood), C.c_void_p(dn_good)))

    def knn2_ratio_l2_u8_dev(self, dq, nq, dt, nt, dim, ratio, dknn, dgood, dn_good, q_index_base=0):
        self._chk(self._L.pm_knn2_ratio_l2_u8_dev(self._h, C.c_void_p(dq), nq, C.c_void_p(dt), nt, dim, C.c_float(ratio),
                                                  q_index_base, C.c_void_p(dknn), C.c_void_p(dgood), C.c_void_p(dn_good)))

    def knn2_l2_u8_dev(self, dq, nq, dt, nt, dim, dout, q_index_base=0):
        self._chk(self._L.pm_knn2_l2_u8_dev(self._h, C.c_void_p(dq), nq, C.c_void_p(dt), nt, dim, q_index_base,
                                            C.c_void_p(dout)))

    def knn2_hamming_dev(self, dq, nq, dt, nt, nbytes, dout, q_index_base=0):
        self._chk(self._L.pm_knn2_hamming_dev(self._h, C.c_void_p(dq), nq, C.c_void_p(dt), nt, nbytes, q_index_base,
                                              C.c_void_p(dout)))

    def ratio_filter_dev(self, dknn, nq, ratio, dout, dn_out):
        self._chk(self._L.pm_ratio_filter_dev(self._h, C.c_void_p(dknn), nq, C.c_float(ratio), C.c_void_p(dout),
                                              C.c_void_p(dn_out)))

    def minmax_filter_dev(self, dm, n, stride, dout, dn_out, dminmax=0):
        self._chk(self._L.pm_minmax_filter_dev(self._h, C.c_void_p(dm), n, stride, C.c_void_p(dout),
                                               C.c_void_p(dn_out), C.c_void_p(dminmax)))

    def col_best_hamming_dev(self, dq, nq, dt, nt, nbytes, dcol, q_index_base=0):
        self._chk(self._L.pm_col_best_hamming_dev(self._h, C.c_void_p(dq), nq, C.c_void_p(dt), nt, nbytes,
                                                  q_index_base, C.c_void_p(dcol)))

    def col_best_l2_f32_dev(self, dq, nq, dt, nt, dim, dcol, q_index_base=0):
        self._chk(self._L.pm_col_best_l2_f32_dev(self._h, C.c_void_p(dq), nq, C.c_void_p(dt), nt, dim,
                                                 q_index_base, C.c_void_p(dcol)))

    def cross_check_dev(self, dknn, nq, stride, dcol, nt, dout, dn_out):
        self._chk(self._L.pm_cross_check_dev(self._h, C.c_void_p(dknn), nq, stride, C.c_void_p(dcol), nt,
                                             C.c_void_p(dout), C.c_void_p(dn_out)))

    def gather_matches_dev(self, dm, dn, max_matches, dkp1, nkp1, dkp2, nkp2, dp1, dp2):
        self._chk(self._L.pm_gather_matches_dev(self._h, C.c_void_p(dm), C.c_void_p(dn), max_matches,
                                                C.c_void_p(dkp1), nkp1, C.c_void_p(dkp2), nkp2,
                                                C.c_void_p(dp1), C.c_void_p(dp2)))

    def find_fundamental_dev(self, dp1, dp2, n, dsamples, n_hyp, sample_size, metric, threshold, refit,
                             dF, dmask, dn_inl, dkey, hyp_id_base=0):
        prm = RansacParams()
        prm.sample_size, prm.metric, prm.threshold = sample_size, metric, threshold
        prm.n_hyp, prm.refit, prm.sample_idx, prm.hyp_id_base = n_hyp, int(bool(refit)), dsamples, hyp_id_base
        self._chk(self._L.pm_find_fundamental_dev(self._h, C.c_void_p(dp1), C.c_void_p(dp2), n, C.byref(prm),
                                                  C.c_void_p(dF), C.c_void_p(dmask), C.c_void_p(dn_inl),
                                                  C.c_void_p(dkey)))

    @staticmethod
    def _pair_params(n_hyp, sample_size, metric, threshold, refit, seed):
        prm = RansacParams()
        prm.sample_size, prm.metric, prm.threshold = sample_size, metric, threshold
        prm.n_hyp, prm.refit, prm.sample_idx, prm.seed, prm.hyp_id_base = n_hyp, int(bool(refit)), None, seed, 0
        return prm

    def match_estimate_pair_dev(self, ddesc1, n1, ddesc2, n2, dim, is_u8, dkp1, dkp2, ratio, dresult, n_hyp,
                                sample_size=8, metric=0, threshold=1.0, refit=True, seed=0):
        """main.cpp:43-98 for one image pair, asynchronous on the ctx stream (no host round trip);
        dresult: device pointer to one PAIR_RESULT record."""
        prm = self._pair_params(n_hyp, sample_size, metric, threshold, refit, seed)
        self._chk(self._L.pm_match_estimate_pair_dev(self._h, C.c_void_p(ddesc1), n1, C.c_void_p(ddesc2), n2, dim, int(is_u8),
                                                     C.c_void_p(dkp1), C.c_void_p(dkp2), C.c_float(ratio), C.byref(prm),
                                                     C.c_uint64(seed), C.c_void_p(dresult)))

    def match_estimate_batched_dev(self, ddesc1, n1, ddesc2, n2, dim, is_u8, dkp1, dkp2, ratio, dresults, n_hyp,
                                   sample_size=8, metric=0, threshold=1.0, refit=True, seed=0):
        """The batched form: sequences of device pointers / counts, one entry per pair; pair p uses seed + p and
        writes record p of dresults (device, [n_pairs] PAIR_RESULT).  Nothing is synchronised."""
        k = len(ddesc1)
        assert len(n1) == len(ddesc2) == len(n2) == len(dkp1) == len(dkp2) == k
        vp, i32 = C.c_void_p * max(k, 1), C.c_int32 * max(k, 1)
        prm = self._pair_params(n_hyp, sample_size, metric, threshold, refit, seed)
        self._chk(self._L.pm_match_estimate_batched_dev(self._h, k, vp(*ddesc1), i32(*n1), vp(*ddesc2), i32(*n2), dim, int(is_u8),
                                                        vp(*dkp1), vp(*dkp2), C.c_float(ratio), C.byref(prm), C.c_void_p(dresults)))

    def match_estimate_batched(self, desc1, n1, desc2, n2, dim, is_u8, kp1, kp2, ratio, n_hyp, sample_size=8, metric=0,
                               threshold=1.0, refit=True, seed=0):
        """pm_match_estimate_batched: sequences of HOST pointers (pinned buffers) / counts, one entry per pair; synchronous.
        Returns a PAIR_RESULT structured array [n_pairs]."""
        k = len(desc1)
        vp, i32 = C.c_void_p * max(k, 1), C.c_int32 * max(k, 1)
        prm = self._pair_params(n_hyp, sample_size, metric, threshold, refit, seed)
        out = np.zeros(max(k, 1), dtype=PAIR_RESULT)
        self._chk(self._L.pm_match_estimate_batched(self._h, k, vp(*desc1), i32(*n1), vp(*desc2), i32(*n2), dim, int(is_u8),
                                                    vp(*kp1), vp(*kp2), C.c_float(ratio), C.byref(prm), _p(out)))
        return out[:k]

    def set_batch_lanes(self, lanes):
        self._chk(self._L.pm_set_batch_lanes(self._h, lanes))

    def batch_warmup(self, n1, n2, dim, is_u8, n_hyp, sample_size=8, metric=0, threshold=1.0, refit=True):
        """Create the lanes of the batched pair call and their workspaces ahead of the first real batch."""
        prm = self._pair_params(n_hyp, sample_size, metric, threshold, refit, 0)
        self._chk(self._L.pm_batch_warmup(self._h, n1, n2, dim, int(is_u8), C.byref(prm)))

    def make_sample_sets_dev(self, n_points, n_hyp, m, seed, dout):
        self._chk(self._L.pm_make_sample_sets_dev(self._h, n_points, n_hyp, m, C.c_uint64(seed), C.c_void_p(dout)))

    def ransac_solve_dev(self, dp1, dp2, n, dsamples, n_hyp, sample_size, dF32):
        self._chk(self._L.pm_ransac_solve_dev(self._h, C.c_void_p(dp1), C.c_void_p(dp2), n, C.c_void_p(dsamples),
                                              n_hyp, sample_size, C.c_void_p(dF32)))

    def ransac_score_dev(self, dp1, dp2, n, dF32, n_models, threshold, metric, dcounts):
        self._chk(self._L.pm_ransac_score_dev(self._h, C.c_void_p(dp1), C.c_void_p(dp2), n, C.c_void_p(dF32),
                                              n_models, C.c_float(threshold), metric, C.c_void_p(dcounts)))

    def ransac_best_dev(self, dcounts, n_models, model_id_base, dkey):
        self._chk(self._L.pm_ransac_best_dev(self._h, C.c_void_p(dcounts), n_models, model_id_base, C.c_void_p(dkey)))

    def ransac_finish_dev(self, dp1, dp2, n, dFw, threshold, metric, refit, dF, dmask, dn_inl):
        self._chk(self._L.pm_ransac_finish_dev(self._h, C.c_void_p(dp1), C.c_void_p(dp2), n, C.c_void_p(dFw),
                                               C.c_float(threshold), metric, int(bool(refit)), C.c_void_p(dF),
                                               C.c_void_p(dmask), C.c_void_p(dn_inl)))


def make_sample_sets(n_points, n_hyp, m=8, seed=0):
    out = np.zeros((n_hyp, m), dtype=np.int32)
    st = _lib.lib().pm_make_sample_sets(n_points, n_hyp, m, C.c_uint64(seed), _p(out))
    if st != PM_OK:
        raise PMError(st, "pm_make_sample_sets: need n_points >= m, 0 < m <= 8")
    return out


_default = None


def default_context():
    global _default
    if _default is None:
        _default = Context(0)
    return _default


# ---------------------------------------------------------------------------------------
# OpenCV look-alikes
# ---------------------------------------------------------------------------------------
class BFMatcher:
    """cv::BFMatcher(normType, crossCheck) -- the 4.x spelling of BruteForceMatcher<L2<float>>
    (main.cpp:43).  knnMatch supports k in {1, 2} (the reference uses match(), i.e. k = 1)."""

    def __init__(self, normType=NORM_L2, crossCheck=False, ctx=None):
        if normType not in (NORM_L2, NORM_HAMMING):
            raise PMError(_lib.PM_BAD_ARG, "normType must be NORM_L2 or NORM_HAMMING")
        self.normType, self.crossCheck = normType, bool(crossCheck)
        self._ctx = ctx

    @property
    def ctx(self):
        return self._ctx or default_context()

    def knnMatch(self, queryDescriptors, trainDescriptors, k=2):
        """Returns a list of per-query DMATCH arrays (rows shorter than k when the train set is
        smaller than k, as in OpenCV)."""
        if k not in (1, 2):
            raise PMError(_lib.PM_BAD_ARG, "k must be 1 or 2")
        if self.crossCheck and k != 1:
            raise PMError(_lib.PM_BAD_ARG, "(-215) K == 1 && update == 0 (crossCheck needs k == 1)")
        if self.crossCheck:
            m = self.ctx.match_cross(queryDescriptors, trainDescriptors, self.normType)
            nq = np.asarray(queryDescriptors).shape[0]
            rows = [m[0:0]] * nq
            for i in range(m.shape[0]):
                rows[int(m["queryIdx"][i])] = m[i:i + 1]
            return rows
        knn = self.ctx.knn2(queryDescriptors, trainDescriptors, self.normType)
        return [row[: min(k, int((row["trainIdx"] >= 0).sum()))] for row in knn]

    def knnMatchArray(self, queryDescriptors, trainDescriptors):
        """[nq, 2] DMATCH array (absent neighbours: trainIdx = -1)."""
        return self.ctx.knn2(queryDescriptors, trainDescriptors, self.normType)

    def match(self, queryDescriptors, trainDescriptors):
        """matcher.match(d1, d2, matches) (main.cpp:46): best match per query, queryIdx order;
        with crossCheck only mutual nearest neighbours."""
        if self.crossCheck:
            return self.ctx.match_cross(queryDescriptors, trainDescriptors, self.normType)
        knn = self.ctx.knn2(queryDescriptors, trainDescriptors, self.normType)
        first = np.ascontiguousarray(knn[:, 0]) if knn.shape[0] else knn.reshape(0)
        return first[first["trainIdx"] >= 0]


def ratio_test(knn, ratio=0.75, ctx=None):
    """Lowe ratio test on a [nq, 2] kNN array (north_star's form of main.cpp:49-69)."""
    return (ctx or default_context()).ratio_filter(knn, ratio)


def minmax_filter(matches, ctx=None):
    """The reference's literal good-match rule (main.cpp:49-69).  Returns (good, min, max)."""
    return (ctx or default_context()).minmax_filter(matches)


def keypoints_convert(keypoints_xy, indices, ctx=None):
    """KeyPoint::convert(keypoints, points2f, keypointIndexes) (main.cpp:90-91)."""
    return (ctx or default_context()).gather_points(keypoints_xy, indices)


def findFundamentalMat(points1, points2, method=FM_RANSAC, ransacReprojThreshold=3.0, confidence=0.99,
                       maxIters=1000, *, sample_size=7, metric=METRIC_SYMEPI, refit=False, batch=1024, seed=0, ctx=None):
    """cv::findFundamentalMat look-alike (main.cpp:95-98) over pm_find_fundamental_mat.  Returns (F, mask) or
    (None, None); the dispatch is OpenCV's:

      N < 7 -> (None, None).  N == 7 (any method) -> the 7-point solver's 1..3 real roots stacked as a [3k, 3] array,
      mask of ones.  FM_8POINT -> N-point normalised 8-point, mask of ones.  FM_RANSAC with N >= 15 -> RANSAC with the
      adaptive stop niters = log(1 - confidence) / log(1 - w^m), run in batches of `batch` hypotheses on the GPU.
      Everything else (FM_7POINT with N > 7 -- the reference's literal call --, FM_LMEDS, FM_RANSAC with N < 15) -> LMedS
      over 7-point samples.  ransacReprojThreshold <= 0 -> 3, confidence outside (0, 1) -> 0.99.

    The keyword defaults are OpenCV's estimator (7-point samples, symmetric-epipolar error, no refit: the mask equals
    cv2's rule err <= thr^2 exactly).  BASELINE.json's north_star variant is sample_size=8, metric=METRIC_SAMPSON,
    refit=True.
    """
    ctx = ctx or default_context()
    opt = FmOptions()
    opt.sample_size, opt.metric, opt.refit, opt.batch, opt.seed = sample_size, metric, int(bool(refit)), batch, seed
    return ctx.find_fundamental_mat(points1, points2, method, ransacReprojThreshold, confidence, maxIters, opt)


def computeCorrespondEpilines(points, whichImage, F, ctx=None):
    """cv::computeCorrespondEpilines (main.cpp:128-132): [n, 3] lines with a^2 + b^2 = 1."""
    return (ctx or default_context()).epilines(points, whichImage, F)
