"""ctypes loader for libpm.so (the C ABI of include/pm.h).

There is no Python or CPU fallback: if the library is missing or no B200 is visible the
import / context creation fails loudly.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("PM_LIBPM_SO") or os.path.join(_HERE, "libpm.so")      # PM_LIBPM_SO: an instrumented build (tools/)

PM_OK, PM_EMPTY, PM_BAD_ARG, PM_CUDA_ERR, PM_NCCL_ERR, PM_NO_DEVICE = 0, 1, -1, -2, -3, -4
DMATCH = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
assert DMATCH.itemsize == 16
PAIR_RESULT = np.dtype([("F", "<f8", (9,)), ("key", "<u8"), ("n_matches", "<i4"), ("n_inliers", "<i4"),
                        ("has_model", "<i4"), ("reserved", "<i4")])
assert PAIR_RESULT.itemsize == 96


class RansacParams(C.Structure):       # pm_ransac_params, 56 bytes
    _fields_ = [("sample_size", C.c_int32), ("metric", C.c_int32), ("threshold", C.c_float),
                ("n_hyp", C.c_int32), ("refit", C.c_int32), ("sample_idx", C.c_void_p),
                ("seed", C.c_uint64), ("hyp_id_base", C.c_int32), ("max_iters", C.c_int32),
                ("confidence", C.c_double)]


class FmOptions(C.Structure):          # pm_fm_options, 24 bytes
    _fields_ = [("sample_size", C.c_int32), ("metric", C.c_int32), ("refit", C.c_int32), ("batch", C.c_int32),
                ("seed", C.c_uint64)]


assert C.sizeof(RansacParams) == 56 and C.sizeof(FmOptions) == 24
COMM_ID_BYTES = 128


# every symbol include/pm.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = [
    "pm_version", "pm_create", "pm_destroy", "pm_set_stream", "pm_sync", "pm_last_error", "pm_set_pipelining",
    "pm_launch_count", "pm_l2_stats", "pm_profile_enable", "pm_profile_read",
    "pm_knn2_l2_f32", "pm_knn2_l2_u8", "pm_knn2_hamming", "pm_knn2_ratio_l2_f32",
    "pm_knn2_l2_f32_dev", "pm_knn2_l2_u8_dev", "pm_knn2_hamming_dev",
    "pm_knn2_ratio_l2_f32_dev", "pm_knn2_ratio_l2_u8_dev",
    "pm_ratio_filter", "pm_ratio_filter_dev", "pm_minmax_filter", "pm_minmax_filter_dev",
    "pm_match_cross_l2_f32", "pm_match_cross_hamming",
    "pm_col_best_hamming_dev", "pm_col_best_l2_f32_dev", "pm_cross_check_dev",
    "pm_gather_points", "pm_gather_matches_dev",
    "pm_find_fundamental", "pm_find_fundamental_dev", "pm_make_sample_sets",
    "pm_ransac_solve_dev", "pm_ransac_score_dev", "pm_ransac_best_dev", "pm_ransac_finish_dev",
    "pm_match_estimate_pair_dev", "pm_match_estimate_batched_dev", "pm_set_batch_lanes", "pm_batch_warmup",
    "pm_fundamental_8point", "pm_epilines", "pm_residuals", "pm_find_fundamental_lmeds", "pm_lmeds_score_dev", "pm_make_sample_sets_dev",
    "pm_match_estimate_batched", "pm_find_fundamental_adaptive", "pm_find_fundamental_mat", "pm_fundamental_7point",
    "pm_comm_unique_id", "pm_comm_init", "pm_set_comm", "pm_comm_info", "pm_match_cross_sharded_dev",
    "pm_allgather_matches_dev", "pm_find_fundamental_sharded_dev", "pm_measure_peak",
    "pm_debug_set_span", "pm_debug_hamming_path", "pm_debug_force_exact", "pm_debug_fallback_no_helpers", "pm_debug_cross_full", "pm_debug_k2_repeat",
    "pm_debug_set_l2_dump", "pm_debug_set_k2_trace", "pm_debug_set_k2_trace_cta",
]


def build(force=False):
    """Compile libpm.so for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", src_dir, "clean", "-s"])
    subprocess.check_call(["make", "-C", src_dir, "-s", "-j8"])
    return SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(points_matching_b200 has no CPU fallback)")
        L = C.CDLL(SO_PATH)
        L.pm_last_error.restype = C.c_char_p
        L.pm_launch_count.restype = C.c_uint64
        for name in EXPORTS:
            getattr(L, name)
        _lib = L
    return _lib
