/*
 * pm_oracle.h -- CPU oracle for the matching + epipolar-geometry hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 *
 * What it restates.  The reference program
 *   /root/reference/Points Matching/main.cpp
 * contains no arithmetic of its own on this path; it calls OpenCV 2.4.13
 * (un-vendored, pinned by link.command.1.tlog:2 "OPENCV_*2413D.LIB"):
 *   main.cpp:43-46   BruteForceMatcher<L2<float>> / DescriptorMatcher::match
 *   main.cpp:49-69   min/max-midpoint "good match" filter
 *   main.cpp:89-91   KeyPoint::convert gather
 *   main.cpp:95-98   cv::findFundamentalMat(..., CV_FM_7POINT)
 *   main.cpp:127-132 cv::computeCorrespondEpilines
 * Each function below cites the call site it stands in for and restates the
 * published OpenCV algorithm behind it.  Parity pin: the reference ships no
 * tests or golden vectors and its OpenCV 2.4.13 Windows
 * binaries cannot run here, so PARITY WITH THE REFERENCE'S OWN BINARY IS
 * UNPINNED.  What pins the oracle instead is OpenCV 4.13 (cv2 -- the same
 * library, same algorithms, importable in the build container):
 * tests/golden/make_golden.py generated the vectors committed under
 * tests/golden/ and tests/test_oracle_golden.py checks every function here
 * against them.
 */
#ifndef PM_ORACLE_H
#define PM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Layout-identical to cv::DMatch (used at main.cpp:45,54,65,76-78). */
typedef struct {
    int32_t queryIdx;
    int32_t trainIdx;
    int32_t imgIdx;
    float   distance;
} orc_dmatch;

enum { ORC_METRIC_SAMPSON = 0, ORC_METRIC_SYMEPI = 1 };

/* ---- matching (main.cpp:43-46) ------------------------------------------ */
/* L2 kNN-2: distance = (float)sqrt(sum (a-b)^2) with the sum in f64; rows are
 * sorted ascending, ties -> lowest trainIdx.  Slots beyond nt get trainIdx=-1. */
void orc_knn2_l2_f32(const float *q, int nq, const float *t, int nt, int dim,
                     orc_dmatch *out /* [nq][2] */, int nthreads);
/* Same arithmetic as the product's FP32 re-rank (fixed summation order, see
 * DESIGN.md "re-rank order"): 32 lane-partials (fmaf chains) + xor-butterfly sum. */
float orc_l2sq_f32_rerank(const float *a, const float *b, int dim);
/* Hamming kNN-2 over `bytes`-wide rows; distance is an integer-valued float. */
void orc_knn2_hamming(const uint8_t *q, int nq, const uint8_t *t, int nt, int bytes,
                      orc_dmatch *out /* [nq][2] */, int nthreads);

/* ---- filters (main.cpp:49-69) ------------------------------------------- */
/* Lowe ratio test: keep knn[i][0] iff both slots valid and d0 < ratio*d1. */
int orc_ratio_filter(const orc_dmatch *knn, int nq, float ratio, orc_dmatch *out);
/* OpenCV crossCheck semantics: keep (i, j=fwd[i]) iff argmin_i' d(i', j) == i,
 * both argmins with lowest-index tie-break.  col_best[j] = packed
 * (dist_bits<<32 | queryIdx) column minima over all queries. */
int orc_cross_check(const orc_dmatch *knn, int nq, int knn_stride,
                    const uint64_t *col_best, int nt, orc_dmatch *out);
void orc_col_best_hamming(const uint8_t *q, int nq, const uint8_t *t, int nt, int bytes,
                          uint64_t *col_best /* [nt] */, int nthreads);
/* The reference's literal rule (main.cpp:49-69): minMatch starts at 1,
 * maxMatch at 0, keep distance < min + (max-min)/2 (double arithmetic). */
int orc_minmax_filter(const orc_dmatch *m, int n, orc_dmatch *out,
                      double *min_out, double *max_out);

/* ---- gather (main.cpp:89-91, KeyPoint::convert with an index list) ------- */
void orc_gather_points(const float *kp_xy /* [nkp][2] */, const int32_t *idx, int n,
                       float *out /* [n][2] */);

/* ---- fundamental matrix (main.cpp:95-98) -------------------------------- */
/* N-point normalised 8-point (OpenCV run8Point).  Returns 1, or 0 when degenerate. */
int orc_fm_8point(const float *p1, const float *p2, int n, double F[9]);
/* Same, over the subset idx[0..m). */
int orc_fm_8point_idx(const float *p1, const float *p2, const int32_t *idx, int m, double F[9]);
/* 7-point (OpenCV run7Point): returns the number of real solutions 0..3. */
int orc_fm_7point(const float *p1, const float *p2, double F[27]);
int orc_fm_7point_idx(const float *p1, const float *p2, const int32_t *idx, double F[27]);

/* Residuals.  *_f32: the product's exact FP32 op order (explicit fmaf), used for
 * bit-exact inlier counts; *_f64: plain double for quality metrics. */
float  orc_sampson_num_den_f32(const float F[9], float x1, float y1, float x2, float y2, float *den);
int    orc_is_inlier_f32(const float F[9], float x1, float y1, float x2, float y2,
                         float thr2, int metric);
double orc_sampson_f64(const double F[9], double x1, double y1, double x2, double y2);
double orc_symepi_f64(const double F[9], double x1, double y1, double x2, double y2);
/* Inlier count + optional mask for one model, FP32 product arithmetic. */
int orc_count_inliers_f32(const float F[9], const float *p1, const float *p2, int n,
                          float thr, int metric, uint8_t *mask /* or NULL */);

/* RANSAC over caller-supplied minimal samples (north_star contract).
 * m = 8 (normalised 8-point) or 7 (7-point, <=3 models per sample).
 * Winner: max inlier count, ties -> lowest model id (id = h for m=8, 3h+k for m=7).
 * Returns 1 and fills F (f64; the refit if refit!=0, else the winning model),
 * mask (winner's inliers, before refit) and n_inliers; 0 if no model was found.
 * counts (optional) receives the per-hypothesis best count.  Fs32 (optional)
 * receives the FP32 model(s) per hypothesis: [nhyp][9] (m=8) or [nhyp][3][9]. */
int orc_ransac_f(const float *p1, const float *p2, int n,
                 const int32_t *sample_idx, int nhyp, int m,
                 int metric, float thr, int refit,
                 double F[9], uint8_t *mask, int *n_inliers, int64_t *best_model,
                 int32_t *counts, float *Fs32, int nthreads);

/* LMedS over caller-supplied 7-point samples (the estimator behind main.cpp:95-98's CV_FM_7POINT when
 * N > 7): models rounded to f32, error = (float) max sym-epi distance in f64, median = sorted err[n/2],
 * winner = smallest median (lowest model id 3*hyp+k on ties), inliers by the 2.5*1.4826*(1+5/(n-7))*sqrt(med)
 * rule, no refit.  medians: [nhyp*3] or NULL; Fs32: [nhyp*3][9] models out (or in when models_given). */
int orc_lmeds_f(const float *p1, const float *p2, int n, const int32_t *sample_idx, int nhyp,
                double F[9], uint8_t *mask, int *n_inliers, int64_t *best_model, float *medians,
                float *Fs32, int models_given);

/* OpenCV-literal estimator restatement (cv::findFundamentalMat dispatch as probed in
 * SURVEY.md section 8 a6): cv::RNG(-1) sample stream, 7-point models, sym-epi metric in
 * double cast to float, adaptive iteration count, no refit.  method: 8 = FM_RANSAC,
 * 4 = FM_LMEDS, 1 = FM_7POINT, 2 = FM_8POINT.  Returns number of 3x3 models written
 * to F (0 = empty Mat, 1, or up to 3 for N==7). */
int orc_find_fundamental_cv(const float *p1, const float *p2, int n, int method,
                            double param1, double param2, int max_iters,
                            double F[27], uint8_t *mask);

/* ---- epilines (main.cpp:127-132) ---------------------------------------- */
/* l = F x (which==1) or F^T x (which==2), scaled so a^2+b^2 = 1. */
void orc_epilines(const float *pts, int n, int which, const double F[9], float *lines /* [n][3] */);

#ifdef __cplusplus
}
#endif
#endif
