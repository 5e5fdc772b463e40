/*
 * pm.h -- C ABI of libpm: B200 (sm_100a) descriptor matching + epipolar geometry.
 *
 * Drop-in boundary for the hot path of /root/reference/Points Matching/main.cpp.
 * The reference has no plugin/FFI layer: the boundary is the OpenCV C++ API subset
 * that main.cpp calls (and that "x64/Debug/Points Matching.exe" imports).  Each entry
 * point cites the call it replaces.  Plain C: no exceptions or STL cross the
 * boundary, outputs are caller-owned buffers sized by the caller, every function
 * returns a pm_status.  One pm_ctx per host thread (it owns the device, a stream,
 * a workspace arena and the TMA descriptors); a ctx is not thread-safe.
 *
 * There is NO CPU fallback: without a CUDA device pm_create fails with
 * PM_NO_DEVICE and nothing else can be called.
 *
 * Two families:
 *   host-buffer calls   (pm_knn2_*, pm_match_*, pm_find_fundamental ...):
 *       synchronous, take HOST pointers, copy in/out -- what main.cpp would call.
 *   device-resident calls (pm_*_dev): take DEVICE pointers, enqueue on the ctx
 *       stream and return without synchronising -- for pipelines that keep data
 *       in HBM (bench `value`, batched config 5, multi-GPU shards).
 */
#ifndef PM_H
#define PM_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_VERSION 200

typedef struct pm_ctx pm_ctx;

/* Layout-identical to cv::DMatch (main.cpp:45 vector<DMatch>; fields used at
 * main.cpp:54,65,76-78,110-113).  16 bytes.  distance is L2 (not squared) for
 * float descriptors and an integer-valued float for Hamming; imgIdx is 0.
 * An absent neighbour (k > number of train rows) has trainIdx = -1, distance = FLT_MAX. */
typedef struct pm_dmatch {
    int32_t queryIdx;
    int32_t trainIdx;
    int32_t imgIdx;
    float   distance;
} pm_dmatch;

typedef enum pm_status {
    PM_OK        = 0,
    PM_EMPTY     = 1,   /* OpenCV returns an empty Mat (e.g. N < 7): outputs untouched */
    PM_BAD_ARG   = -1,  /* OpenCV would raise cv::Exception -215 (type/dim mismatch)   */
    PM_CUDA_ERR  = -2,
    PM_NCCL_ERR  = -3,  /* libnccl.so.2 not loadable, no communicator set, or an NCCL call failed */
    PM_NO_DEVICE = -4
} pm_status;

/* ---- context ---------------------------------------------------------------- */
int  pm_version(void);
int  pm_create(pm_ctx **ctx, int device);
int  pm_destroy(pm_ctx *ctx);
/* Run on an existing cudaStream_t (e.g. torch's current stream); NULL = ctx-owned stream.  Changing the
 * stream first waits for the work this ctx enqueued on the old one (the workspaces are shared). */
int  pm_set_stream(pm_ctx *ctx, void *cuda_stream);
int  pm_sync(pm_ctx *ctx);
/* Opt-in overlap of consecutive pm_knn2_ratio_l2_*_dev calls of equal shapes (default off).  With it on,
 * the pack kernel of call i+1 starts while the finish / filter kernels of call i still run (it skips the
 * stream-order wait; the library orders the two chains itself and every intermediate buffer exists twice).
 * Contract: the input descriptors of such a call must be complete and visible when the call is enqueued
 * (host-synchronised, or ordered by an event the ctx stream waits on) -- NOT produced by a kernel enqueued
 * on the ctx stream right before it.  Outputs are ordered as usual: call i+1 never writes before call i
 * has finished.  Results are identical with and without it. */
int  pm_set_pipelining(pm_ctx *ctx, int on);
const char *pm_last_error(pm_ctx *ctx);
/* Number of libpm kernels launched by this ctx since creation (bench "gpu_launches"). */
uint64_t pm_launch_count(pm_ctx *ctx);
/* Counters of the last L2 call: [0] exact-integer mode (1/0), [1] rows sent to the
 * exact fallback, [2] MMA k-blocks per tile, [3] segments per row tile. */
int  pm_l2_stats(pm_ctx *ctx, int32_t out[4]);
/* Device-side timing of the dominant kernels with CUDA events on the ctx stream (bench
 * roofline): enable, run any number of calls, then read.  which: 0 = L2 tensor-core kernel
 * (K2), 1 = Hamming kNN kernel (K4), 2 = RANSAC scoring kernel (K7).  pm_profile_read
 * synchronises the stream, returns the summed milliseconds and the launch count since the
 * last read, and resets them. */
int  pm_profile_enable(pm_ctx *ctx, int on);
int  pm_profile_read(pm_ctx *ctx, int which, double *total_ms, int *n_launches);

/* ---- descriptor matching ----------------------------------------------------
 * Replaces BruteForceMatcher<L2<float>> matcher; matcher.match(d1, d2, matches)
 * (main.cpp:43-46; 4.x spelling BFMatcher(NORM_L2).knnMatch(q, t, k=2)).
 * out is [nq][2], sorted ascending per row, ties -> lowest trainIdx; out[i][0] is
 * exactly what match() (k=1) returns.  q_index_base is added to queryIdx (row shards). */
int pm_knn2_l2_f32(pm_ctx *ctx, const float *q, int nq, const float *t, int nt, int dim,
                   pm_dmatch *out);
/* SIFT stored as bytes (0..255): same result as the f32 call on the widened data. */
int pm_knn2_l2_u8(pm_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t, int nt, int dim,
                  pm_dmatch *out);
/* BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2) for `bytes`-wide binary rows (ORB: 32). */
int pm_knn2_hamming(pm_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t, int nt, int bytes,
                    pm_dmatch *out);

/* knnMatch(k=2) + Lowe ratio test in one call (the north_star flow for main.cpp:43-69):
 * knn_out (optional, [nq][2]) and the compacted good matches come back together, so the
 * kNN result never makes a host round trip between the two steps.  good_out must hold nq
 * entries; entries past *n_good are unspecified. */
int pm_knn2_ratio_l2_f32(pm_ctx *ctx, const float *q, int nq, const float *t, int nt, int dim,
                         float ratio, pm_dmatch *knn_out, pm_dmatch *good_out, int *n_good);

int pm_knn2_l2_f32_dev(pm_ctx *ctx, const float *dq, int nq, const float *dt, int nt, int dim,
                       int q_index_base, pm_dmatch *dout);
int pm_knn2_l2_u8_dev(pm_ctx *ctx, const uint8_t *dq, int nq, const uint8_t *dt, int nt, int dim,
                      int q_index_base, pm_dmatch *dout);
int pm_knn2_hamming_dev(pm_ctx *ctx, const uint8_t *dq, int nq, const uint8_t *dt, int nt,
                        int bytes, int q_index_base, pm_dmatch *dout);
/* Device-resident knnMatch(k=2) + Lowe ratio test (main.cpp:43-69 as the north_star restates it) in
 * ONE call: the whole chain pack -> GEMM/top-2 -> re-rank -> filter is enqueued together.  Same results
 * as pm_knn2_l2_f32_dev followed by pm_ratio_filter_dev.  dknn [nq][2], dgood [nq], dn_good [1], all
 * device memory. */
int pm_knn2_ratio_l2_f32_dev(pm_ctx *ctx, const float *dq, int nq, const float *dt, int nt, int dim,
                             float ratio, int q_index_base, pm_dmatch *dknn, pm_dmatch *dgood,
                             int32_t *dn_good);
int pm_knn2_ratio_l2_u8_dev(pm_ctx *ctx, const uint8_t *dq, int nq, const uint8_t *dt, int nt, int dim,
                            float ratio, int q_index_base, pm_dmatch *dknn, pm_dmatch *dgood,
                            int32_t *dn_good);

/* ---- good-match filters (main.cpp:49-69) -------------------------------------
 * Lowe ratio test over a [nq][2] kNN result: keep knn[i][0] iff both neighbours
 * exist and d0 < ratio*d1.  Output is compacted in queryIdx order. */
int pm_ratio_filter(pm_ctx *ctx, const pm_dmatch *knn, int nq, float ratio,
                    pm_dmatch *out, int *n_out);
int pm_ratio_filter_dev(pm_ctx *ctx, const pm_dmatch *dknn, int nq, float ratio,
                        pm_dmatch *dout, int32_t *dn_out);
/* The reference's literal rule (main.cpp:49-69): min starts at 1, max at 0, keep
 * distance < min + (max-min)/2.  m is a k=1 match list (stride_elems 1) or the
 * first column of a kNN-2 result (stride_elems 2). */
int pm_minmax_filter(pm_ctx *ctx, const pm_dmatch *m, int n, int stride_elems,
                     pm_dmatch *out, int *n_out, double *min_out, double *max_out);
int pm_minmax_filter_dev(pm_ctx *ctx, const pm_dmatch *dm, int n, int stride_elems,
                         pm_dmatch *dout, int32_t *dn_out, double *dminmax /* [2] */);
/* BFMatcher(norm, crossCheck=true).match(q, t): (i, j) survives iff j is i's nearest
 * train row and i is j's nearest query row, both with lowest-index ties. */
int pm_match_cross_l2_f32(pm_ctx *ctx, const float *q, int nq, const float *t, int nt, int dim,
                          pm_dmatch *out, int *n_out);
int pm_match_cross_hamming(pm_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t, int nt,
                           int bytes, pm_dmatch *out, int *n_out);
/* Building blocks of the sharded cross-check: column minima over a query
 * shard as packed u64 (float_bits(dist) << 32 | queryIdx) -- all-reduce(min) them
 * across ranks -- then the local filter bwd[fwd[i]] == i. */
int pm_col_best_hamming_dev(pm_ctx *ctx, const uint8_t *dq, int nq, const uint8_t *dt, int nt,
                            int bytes, int q_index_base, uint64_t *dcol_best /* [nt] */);
int pm_col_best_l2_f32_dev(pm_ctx *ctx, const float *dq, int nq, const float *dt, int nt, int dim,
                           int q_index_base, uint64_t *dcol_best /* [nt] */);
int pm_cross_check_dev(pm_ctx *ctx, const pm_dmatch *dknn, int nq, int knn_stride_elems,
                       const uint64_t *dcol_best, int nt, pm_dmatch *dout, int32_t *dn_out);

/* ---- KeyPoint::convert(keypoints, points, indices) (main.cpp:89-91) ---------- */
int pm_gather_points(pm_ctx *ctx, const float *kp_xy /* [nkp][2] */, int nkp,
                     const int32_t *idx, int n, float *out /* [n][2] */);
/* Device hand-off: matches -> the two correspondence lists, no host trip. */
int pm_gather_matches_dev(pm_ctx *ctx, const pm_dmatch *dmatches, const int32_t *dn_matches,
                          int max_matches, const float *dkp1, int nkp1, const float *dkp2, int nkp2,
                          float *dp1, float *dp2);

/* ---- cv::findFundamentalMat (main.cpp:95-98) --------------------------------- */
enum { PM_METRIC_SAMPSON = 0,      /* r^2 / (a^2+b^2+a'^2+b'^2)      (north_star)          */
       PM_METRIC_SYMEPI  = 1 };    /* max(d(x2,Fx1)^2, d(x1,F^Tx2)^2) (OpenCV computeError) */

typedef struct pm_ransac_params {
    int32_t sample_size;   /* 8: normalised 8-point; 7: 7-point (up to 3 models per sample) */
    int32_t metric;        /* PM_METRIC_*                                                     */
    float   threshold;     /* pixels; inlier <=> err <= threshold^2                           */
    int32_t n_hyp;         /* hypotheses (minimal samples) to evaluate; pm_find_fundamental_adaptive: per batch */
    int32_t refit;         /* !=0: N-point normalised 8-point on the winner's inliers         */
    const int32_t *sample_idx; /* [n_hyp][sample_size] HOST (host call) / DEVICE (_dev call)  */
                           /* index sets, distinct within a row; NULL: generated from `seed`  */
    uint64_t seed;
    int32_t hyp_id_base;   /* global id of this shard's hypothesis 0 (multi-GPU)              */
    int32_t max_iters;     /* pm_find_fundamental_adaptive: cap on the hypotheses (cv maxIters); <= 0: 1000 */
    double  confidence;    /* pm_find_fundamental_adaptive: cv `confidence`; outside (0,1): 0.99 */
} pm_ransac_params;        /* 56 bytes */

/* RANSAC over minimal samples.  Winner = max inlier count, ties -> lowest model id
 * (id = hyp for 8-point, 3*hyp+k for 7-point).  F is row-major 3x3 f64 with F[8]=1;
 * mask[n] in {0,1} is the winner's inlier set (before the refit); PM_EMPTY when n <
 * sample_size or no hypothesis produced a model (F, mask untouched). */
int pm_find_fundamental(pm_ctx *ctx, const float *p1 /* [n][2] */, const float *p2, int n,
                        const pm_ransac_params *prm,
                        double F[9], uint8_t *mask, int *n_inliers);

/* The same estimator with OpenCV's adaptive termination (findFundamentalMat's `confidence` / `maxIters`), driven from
 * the device: the correspondences are uploaded ONCE, the minimal samples of every batch (prm->n_hyp per batch, <= 0:
 * 1024) are generated on the device from prm->seed (prm->sample_idx is ignored), and only the 8-byte running winner key
 * returns to the host per batch -- it feeds  niters = log(1 - confidence) / log(1 - w^sample_size)  (cv's
 * RANSACUpdateNumIters), w = best inlier ratio so far.  The winner over all batches is "most inliers, lowest global
 * model id on ties"; F, mask and n_inliers are read back once at the end.  *n_hyp_run (optional) = hypotheses evaluated. */
int pm_find_fundamental_adaptive(pm_ctx *ctx, const float *p1, const float *p2, int n, const pm_ransac_params *prm,
                                 double F[9], uint8_t *mask, int *n_inliers, int *n_hyp_run);

/* cv::findFundamentalMat(points1, points2, method, param1, param2, mask) itself -- the call at main.cpp:95-98 with OpenCV's
 * dispatch table (probed on cv2 4.13):
 *     n < 7                                  -> PM_EMPTY (empty Mat)
 *     n == 7 (any method)                    -> 7-point on the seven points: 1..3 stacked 3x3 models (cv returns 9x3), mask ones
 *     PM_FM_8POINT                           -> N-point normalised 8-point, mask ones
 *     PM_FM_RANSAC and n >= 15               -> RANSAC (adaptive, as pm_find_fundamental_adaptive)
 *     everything else (PM_FM_7POINT with n > 7 = the reference's literal call, PM_FM_LMEDS, PM_FM_RANSAC with n < 15)
 *                                            -> LMedS over 7-point samples, niters from `param2` at outlier ratio 0.45
 *     param1 <= 0 -> 3;  param2 outside (DBL_EPSILON, 1 - DBL_EPSILON) -> 0.99;  max_iters <= 0 -> 1000
 * F receives *n_models (1..3) row-major 3x3 f64 matrices with F[8] = 1; mask (optional) n bytes in {0, 1}.
 * opt == NULL is OpenCV's estimator: 7-point samples, symmetric-epipolar error, no refit.  The north_star's variant
 * (8-point samples, Sampson error, 8-point refit on the inliers) is opt = {8, PM_METRIC_SAMPSON, 1, ...}.
 * The sample stream is libpm's own (splitmix64 from opt->seed), not cv::RNG: F is an equally valid RANSAC / LMedS answer,
 * not bit-equal to OpenCV's. */
enum { PM_FM_7POINT = 1, PM_FM_8POINT = 2, PM_FM_LMEDS = 4, PM_FM_RANSAC = 8 };   /* == cv::FM_* / CV_FM_* */
typedef struct pm_fm_options {
    int32_t sample_size;   /* RANSAC minimal sample: 7 (OpenCV) or 8; 0 -> 7                                 */
    int32_t metric;        /* PM_METRIC_SYMEPI (OpenCV) or PM_METRIC_SAMPSON                                   */
    int32_t refit;         /* RANSAC: !=0 refits the winner on its inliers (OpenCV does not)                   */
    int32_t batch;         /* RANSAC hypotheses per batch between two looks at the adaptive stop; <= 0: 1024   */
    uint64_t seed;
} pm_fm_options;           /* 24 bytes */
int pm_find_fundamental_mat(pm_ctx *ctx, const float *p1, const float *p2, int n, int method, double param1, double param2,
                            int max_iters, const pm_fm_options *opt, double F[27], int *n_models, uint8_t *mask);
/* The n == 7 case on its own: run7Point on seven correspondences, all real roots (1..3), full double precision. */
int pm_fundamental_7point(pm_ctx *ctx, const float *p1, const float *p2, int n, double F[27], int *n_models);

/* Deterministic minimal-sample index sets ([n_hyp][m], distinct within a row); the
 * same (n_points, n_hyp, m, seed) gives the same sets on every rank. */
int pm_make_sample_sets(int n_points, int n_hyp, int m, uint64_t seed, int32_t *out);
/* The same sets generated on the device (one thread per hypothesis), asynchronous on the ctx stream. */
int pm_make_sample_sets_dev(pm_ctx *ctx, int n_points, int n_hyp, int m, uint64_t seed, int32_t *dout);

/* Staged device API (what pm_find_fundamental runs; exposed for shards and tests).
 *  solve : dF32 [n_hyp][models][12] f32 (9 used; NaN = no model), models = 1 or 3
 *  score : dcounts [n_hyp*models] inlier count per model, FP32 arithmetic as in DESIGN.md
 *  best  : *dkey = max over models of (count << 32 | (0xFFFFFFFF - model_id)), 0 if none
 *  finish: winner re-scored -> dmask, dn_inliers; optional refit -> dF (f64[9]) */
int pm_ransac_solve_dev(pm_ctx *ctx, const float *dp1, const float *dp2, int n,
                        const int32_t *dsample_idx, int n_hyp, int sample_size, float *dF32);
int pm_ransac_score_dev(pm_ctx *ctx, const float *dp1, const float *dp2, int n,
                        const float *dF32, int n_models, float threshold, int metric,
                        int32_t *dcounts);
int pm_ransac_best_dev(pm_ctx *ctx, const int32_t *dcounts, int n_models, int model_id_base,
                       uint64_t *dkey);
int pm_ransac_finish_dev(pm_ctx *ctx, const float *dp1, const float *dp2, int n,
                         const float *dF32_winner /* [12] */, float threshold, int metric, int refit,
                         double *dF /* [9] */, uint8_t *dmask, int32_t *dn_inliers);

/* All four stages, device resident and asynchronous on the ctx stream (prm->sample_idx is a
 * DEVICE pointer and must not be NULL).  dkey receives the winner key (0 = no model: dF,
 * dmask are then undefined and *dn_inliers is 0). */
int pm_find_fundamental_dev(pm_ctx *ctx, const float *dp1, const float *dp2, int n,
                            const pm_ransac_params *prm, double *dF /* [9] */, uint8_t *dmask,
                            int32_t *dn_inliers, uint64_t *dkey);

/* ---- whole pairs, device-resident and asynchronous (BASELINE config 5; main.cpp:43-98 end to end) ----
 * knnMatch(k=2) -> ratio test -> KeyPoint::convert on both sides -> findFundamentalMat(RANSAC) for one image
 * pair, enqueued on the ctx stream WITHOUT any host round trip: the number of good matches stays on the
 * device and every RANSAC kernel reads it there.  ddesc1/ddesc2: [n1|n2][dim] f32 (is_u8 = 0) or u8 (1)
 * descriptors, dkp1/dkp2: [n1|n2][2] f32 keypoint coordinates, all device memory.  prm->sample_idx is
 * ignored: the minimal-sample index sets are generated on the device from `seed` (the sets
 * pm_make_sample_sets(n_matches, n_hyp, m, seed) returns).  *dresult (device) is complete once the stream
 * has passed the call; has_model = 0 when fewer than sample_size matches survive or no sample gave a model. */
typedef struct pm_pair_result {
    double   F[9];          /* row-major 3x3, F[8] = 1 (zeros when has_model == 0)       */
    uint64_t key;           /* winner key: inliers << 32 | (0xFFFFFFFF - model id); 0 = none */
    int32_t  n_matches;     /* good matches after the ratio test                          */
    int32_t  n_inliers;     /* inliers of the winning model                               */
    int32_t  has_model;
    int32_t  reserved;
} pm_pair_result;           /* 96 bytes */
int pm_match_estimate_pair_dev(pm_ctx *ctx, const void *ddesc1, int n1, const void *ddesc2, int n2, int dim,
                               int is_u8, const float *dkp1, const float *dkp2, float ratio,
                               const pm_ransac_params *prm, uint64_t seed, pm_pair_result *dresult);
/* The batched form: n_pairs independent pairs given as HOST arrays of device pointers and counts; pair p
 * uses seed prm->seed + p and writes dresults[p].  Everything is enqueued before the call returns; nothing
 * is synchronised.  Ranks of a multi-GPU job call it on their own slice of the pairs (no collective). */
int pm_match_estimate_batched_dev(pm_ctx *ctx, int n_pairs, const void *const *ddesc1, const int32_t *n1,
                                  const void *const *ddesc2, const int32_t *n2, int dim, int is_u8,
                                  const float *const *dkp1, const float *const *dkp2, float ratio,
                                  const pm_ransac_params *prm, pm_pair_result *dresults);

/* The same batch from HOST memory, synchronous (config 5 end to end): desc1 / desc2 / kp1 / kp2 are arrays of HOST pointers
 * (pinned memory for full PCIe speed), results [n_pairs] is host memory.  Every lane uploads its pairs on its own stream
 * right before their kernels, so uploads run under the other lanes' kernels; 96 bytes per pair come back, once, at the end. */
int pm_match_estimate_batched(pm_ctx *ctx, int n_pairs, const void *const *desc1, const int32_t *n1,
                              const void *const *desc2, const int32_t *n2, int dim, int is_u8,
                              const float *const *kp1, const float *const *kp2, float ratio,
                              const pm_ransac_params *prm, pm_pair_result *results);

/* The batched calls cut the pairs into groups of 16 consecutive pairs (PM_PAIR_GROUP in the environment overrides): the
 * pairs of a group run their matching chains one after the other and then share ONE launch of every RANSAC kernel (the
 * pair is a grid dimension), so those kernels fill the GPU instead of running at a fraction of a wave per pair.  Group g
 * runs on lane g mod `lanes` (1..8 internal streams with their own workspaces, one host thread enqueues each; default 4), so
 * the latency-bound kernels of one group overlap the matching of another.  Results do not depend on either number.
 * Measured per 8192 x 8192 pair with 4096 hypotheses: 1 lane, groups of 1 = 203 us; 8 lanes, groups of 1 (the round-1 form) =
 * 60 us; 4 lanes, groups of 16 = 43 us.  A lane and its workspaces are created the first time it is used (cudaMalloc
 * synchronises the device), so the first batch after a change of lanes or shapes is slow: warm up once.
 * PM_BATCH_THREADS=0 (environment) makes one host thread enqueue all lanes. */
int pm_set_batch_lanes(pm_ctx *ctx, int lanes);
/* Optional: create the lanes and their workspaces for pairs of up to n1 x n2 descriptors now (a throw-away batch
 * of empty pairs; synchronises), so that the first real batch does not pay for it. */
int pm_batch_warmup(pm_ctx *ctx, int n1, int n2, int dim, int is_u8, const pm_ransac_params *prm);

/* LMedS over 7-point minimal samples -- what cv::findFundamentalMat(p1, p2, CV_FM_7POINT) actually runs when
 * N > 7, i.e. the reference's literal call at main.cpp:95-98.  Per model the error is OpenCV's
 * max(d(x2,Fx1)^2, d(x1,F^T x2)^2) in FP64 cast to float; the model with the smallest median wins (lowest
 * model id = 3*hyp+k on ties); inliers are err <= (2.5*1.4826*(1+5/(n-7))*sqrt(median))^2; no refit.
 * sample_idx: [n_hyp][7] HOST index sets or NULL (generated from seed).  PM_EMPTY when n < 8 or no model. */
int pm_find_fundamental_lmeds(pm_ctx *ctx, const float *p1, const float *p2, int n, int n_hyp,
                              const int32_t *sample_idx, uint64_t seed, double F[9], uint8_t *mask,
                              int *n_inliers, float *median_out);
/* The scoring stage alone: dmedians[m] = median error of model m (+inf for a NaN model). */
int pm_lmeds_score_dev(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const float *dF32, int n_models,
                       float *dmedians);

/* N-point normalised 8-point (findFundamentalMat(..., FM_8POINT)); mask all ones. */
int pm_fundamental_8point(pm_ctx *ctx, const float *p1, const float *p2, int n, double F[9]);

/* ---- multi-GPU: one process (or host thread) per GPU, NCCL over NVLink -------------------------------
 * The reference is single-process; BASELINE.json's north_star shards its two calls: query rows per rank for the matcher
 * (cross-check needs ONE exchange: min-reduce of the packed column minima), hypothesis batches per rank for RANSAC (ONE
 * exchange: max-reduce of the 8-byte winner key).  libpm loads libnccl.so.2 at run time (dlopen: a process that never calls
 * these entries does not need NCCL); every failure on that side returns PM_NCCL_ERR with pm_last_error() set.
 *   pm_comm_unique_id  rank 0 makes the 128-byte ncclUniqueId and ships it to the other ranks (MPI, torch.distributed, a file)
 *   pm_comm_init       ncclCommInitRank on the ctx's device; the ctx owns the communicator (destroyed by pm_destroy)
 *   pm_set_comm        or: borrow the caller's ncclComm_t (void* = ncclComm_t); NULL detaches
 * The collectives run on the ctx stream between the kernels they connect. */
#define PM_COMM_ID_BYTES 128
int pm_comm_unique_id(void *id /* [PM_COMM_ID_BYTES] */);
int pm_comm_init(pm_ctx *ctx, int n_ranks, int rank, const void *id /* [PM_COMM_ID_BYTES] */);
int pm_set_comm(pm_ctx *ctx, void *nccl_comm, int n_ranks, int rank);
int pm_comm_info(pm_ctx *ctx, int *n_ranks, int *rank);   /* 1, 0 without a communicator */
/* BFMatcher(norm, crossCheck=true).match over a query set sharded by rows: this rank holds rows [q_index_base,
 * q_index_base + nq) and the whole (replicated) train set.  kNN-2 of the shard -> dknn ([nq][2], optional); the train rows that
 * are the best match of some query are marked (ncclAllReduce(ncclMax, ncclUint8) of the [nt] mark bytes: the union over the
 * ranks) and only those rows take part in the reverse pass against the shard (all rows when more than 3/4 are marked) -> packed
 * column minima -> dcol_best ([nt], scratch the caller provides; rows nobody points at hold "none") -> ncclAllReduce(ncclMin,
 * ncclUint64) in place -> local filter: the shard's mutual matches in queryIdx order -> dout ([nq]), *dn_out.  Concatenating
 * the ranks' lists in rank order gives exactly the single-GPU result.  With several ranks the call reads ONE 4-byte count back
 * from the device (the number of marked rows sizes the reverse pass), i.e. it waits for the forward pass and the rest stays
 * enqueued; on one rank the reverse pass is sized by the bound min(nq, nt) and nothing is synchronised.
 * norm: 4 = L2 (f32 rows of `width` floats), 6 = Hamming (rows of `width` bytes). */
int pm_match_cross_sharded_dev(pm_ctx *ctx, const void *dq, int nq, const void *dt, int nt, int width, int norm,
                               int q_index_base, pm_dmatch *dknn, uint64_t *dcol_best, pm_dmatch *dout, int32_t *dn_out);
/* Fixed-width all-gather of per-rank match lists (the "gather of match results" of the north_star): dall is
 * [n_ranks][max_per_rank], dcounts [n_ranks]; rank r's list is dall[r][0 .. dcounts[r]). */
int pm_allgather_matches_dev(pm_ctx *ctx, const pm_dmatch *dlocal, const int32_t *dn_local, int max_per_rank,
                             pm_dmatch *dall, int32_t *dcounts);
/* RANSAC-F with the hypotheses sharded by batch: this rank solves and scores hypotheses [prm->hyp_id_base, + prm->n_hyp) of a
 * job of n_hyp_total; prm->sample_idx is the FULL [n_hyp_total][sample_size] DEVICE array, identical on every rank, or NULL
 * (every rank generates the sets of its shard, and later of the winner, from prm->seed: same sets as
 * pm_make_sample_sets(n, n_hyp_total, m, seed)).  One ncclAllReduce(ncclMax, ncclUint64) of the winner key; every rank then
 * re-solves the winning index set itself (no broadcast) and finishes (mask, optional refit): dF, dmask, dn_inliers and dkey
 * are identical on every rank and identical to a single-GPU run over all n_hyp_total hypotheses. */
int pm_find_fundamental_sharded_dev(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const pm_ransac_params *prm,
                                    int n_hyp_total, double *dF /* [9] */, uint8_t *dmask, int32_t *dn_inliers, uint64_t *dkey);

/* ---- measured pipe peaks (bench.py roofline denominators; BASELINE.md: "measure them on the box") -----------------
 * Independent-chain microkernels, one resident wave, timed with CUDA events on the ctx stream.
 * which 0: FP32 FFMA issue -> *value = TFLOP/s (2 FLOP per FFMA);  1: POPC.32 -> *value = 1e12 POPC per second. */
int pm_measure_peak(pm_ctx *ctx, int which, double *value);

/* ---- NVTX: every entry point above opens an NVTX range named after itself (domain "libpm") when a profiler is attached;
 * no cost otherwise. ---- */

/* ---- debug hooks (process-global, NOT thread-safe, not part of the drop-in surface; used by tools/ and two bench legs) ----
 *   pm_debug_set_span(p)            every kernel of the L2 chain stamps %globaltimer marks into p (tools/step_timeline.py)
 *   pm_debug_hamming_path(k)        0 auto, 1 force the POPC kernel, 2 force the tensor-core kernel
 *   pm_debug_force_exact(on)        L2: exact FP32 kernel for every row (cross-check of the two paths)
 *   pm_debug_fallback_no_helpers(on) L2 split mode: no helper blocks, the last row block of K3 runs the flagged-row scan alone
 *   pm_debug_cross_full(on)         cross-check: reverse pass over the whole train set instead of the marked rows only
 *   pm_debug_k2_repeat(n)           the L2 GEMM kernel is launched n times back to back inside one profiling event pair (same result)
 *   pm_debug_set_l2_dump(p), pm_debug_set_k2_trace(p)   K2 tile dump / clock64 trace (PM_K2_TRACE builds) */
void pm_debug_set_span(unsigned long long *p);
void pm_debug_hamming_path(int path);
void pm_debug_force_exact(int on);
void pm_debug_fallback_no_helpers(int on);
void pm_debug_cross_full(int on);
void pm_debug_k2_repeat(int n);
void pm_debug_set_l2_dump(float *ddump);
void pm_debug_set_k2_trace(long long *p);
void pm_debug_set_k2_trace_cta(int cta);

/* ---- diagnostics (main.cpp:103-123, 127-132) --------------------------------- */
/* cv::computeCorrespondEpilines: l = F x (which_image 1) or F^T x (2), a^2+b^2 = 1. */
int pm_epilines(pm_ctx *ctx, const float *pts, int n, int which_image, const double F[9],
                float *lines /* [n][3] */);
/* Per-correspondence residuals in the correct convention x2^T F x1 (main.cpp:110-117
 * evaluates x1^T F x2): out[i] = Sampson or sym-epi distance (f32). */
int pm_residuals(pm_ctx *ctx, const float *p1, const float *p2, int n, const double F[9],
                 int metric, float *out, double *mean_out);

#ifdef __cplusplus
}
#endif
#endif /* PM_H */
