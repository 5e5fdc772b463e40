// ransac.cu -- K6/K7/K8: RANSAC fundamental-matrix estimation for sm_100a, plus the
// epipolar diagnostics.
//
// Replaces cv::findFundamentalMat (/root/reference/Points Matching/main.cpp:95-98) --
// CvFMEstimator run8Point / run7Point, computeReprojError + findInliers, model
// selection -- cv::computeCorrespondEpilines (main.cpp:127-132) and the residual loop
// at main.cpp:103-123 (in the correct convention x2^T F x1).
//
//   K6 solve : 8 lanes per hypothesis (4 hypotheses per warp).  Lane j holds the j-th
//              sample's 9-vector; Householder QR of the 9x8 (9x7) matrix across lanes via
//              width-8 shuffles; null vector(s) = Q e9 (Q e8, Q e9); rank-2 projection by
//              a 3x3 one-sided Jacobi; all in registers, FP64 (the solve is <2% of the
//              work, and FP64 keeps F within ~1e-12 of OpenCV's double-precision path).
//   K7 score : all-pairs (model, point) -> FP32-issue bound, not HBM bound.  Each thread
//              owns 4 models (36 coefficient registers), the CTA streams the packed
//              correspondences through a double-buffered shared-memory tile and reads
//              them with broadcast LDS.128; counts stay private -> no reduction in the
//              hot loop.  Arithmetic is the fixed FP32 order of DESIGN.md "scoring op
//              order" (explicit fmaf), bit-exact against the oracle.
//   K8 finish: winner's mask + count, N-point normalised 8-point refit on the inliers
//              (deterministic two-level FP64 reductions, 9x9 Jacobi on one warp).
#include <cfloat>
#include "pm_internal.h"

namespace {

// =====================================================================================
// small FP64 helpers (register resident)
// =====================================================================================
__device__ __forceinline__ double grp_sum8(double v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 4, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 2, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 1, 8);
    return v;
}

// One-sided Jacobi on a 3x3 (columns rotated until orthogonal), then drop the column of
// smallest norm: F <- rank-2 projection (OpenCV run8Point "make F0 singular").
__device__ void rank2_project3(double (&F)[9])
{
    double A[9], V[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { A[i] = F[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool changed = false;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            double a = 0, b = 0, g = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) { a += A[k * 3 + p] * A[k * 3 + p]; b += A[k * 3 + q] * A[k * 3 + q]; g += A[k * 3 + p] * A[k * 3 + q]; }
            // converged when the columns are orthogonal to working precision.  (The first version asked for 1e-17, below
            // the rounding noise of g itself: almost every matrix then ran all 30 sweeps -- 62 us per launch of the projection.)
            if (g == 0.0 || g * g <= 5.3e-32 * (a * b)) continue;
            changed = true;
            const double zeta = (b - a) / (2 * g);
            const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
            const double c = 1 / sqrt(1 + tt * tt), s = c * tt;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double x = A[k * 3 + p], y = A[k * 3 + q];
                A[k * 3 + p] = c * x - s * y; A[k * 3 + q] = s * x + c * y;
                x = V[k * 3 + p]; y = V[k * 3 + q];
                V[k * 3 + p] = c * x - s * y; V[k * 3 + q] = s * x + c * y;
            }
        }
        if (!changed) break;
    }
    double nrm[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) nrm[j] = A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j];
    const int drop = (nrm[0] <= nrm[1] && nrm[0] <= nrm[2]) ? 0 : (nrm[1] <= nrm[2] ? 1 : 2);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double v = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (k != drop) v += A[i * 3 + k] * V[j * 3 + k];
            F[i * 3 + j] = v;
        }
}

// cv::solveCubic restatement (c0 x^3 + c1 x^2 + c2 x + c3 = 0); returns the root count.
__device__ int solve_cubic(const double (&c)[4], double (&r)[3])
{
    const double PI = 3.1415926535897932384626433832795;
    double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
    double x0 = 0, x1 = 0, x2 = 0;
    int n = 0;
    if (a0 == 0) {
        if (a1 == 0) {
            if (a2 == 0) n = a3 == 0 ? -1 : 0;
            else { x0 = -a3 / a2; n = 1; }
        } else {
            double d = a2 * a2 - 4 * a1 * a3;
            if (d >= 0) {
                d = sqrt(d);
                const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
                if (fabs(q1) > fabs(q2)) { x0 = q1 / a1; x1 = a3 / q1; }
                else { x0 = q2 / a1; x1 = a3 / q2; }
                n = d > 0 ? 2 : 1;
            }
        }
    } else {
        a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
        const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
        const double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
        const double Qcubed = Q * Q * Q;
        double d = Qcubed - R * R;
        if (d > 0) {
            const double theta = acos(R / sqrt(Qcubed));
            const double sqrtQ = sqrt(Q);
            const double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
            x0 = t0 * cos(t1) - t2;
            x1 = t0 * cos(t1 + (2. * PI / 3)) - t2;
            x2 = t0 * cos(t1 + (4. * PI / 3)) - t2;
            n = 3;
        } else if (d == 0) {
            if (R >= 0) { x0 = -2 * pow(R, 1. / 3) - a1 / 3; x1 = pow(R, 1. / 3) - a1 / 3; }
            else { x0 = 2 * pow(-R, 1. / 3) - a1 / 3; x1 = -pow(-R, 1. / 3) - a1 / 3; }
            x2 = 0;
            n = x0 == x1 ? 1 : 2;
            x1 = x0 == x1 ? 0 : x1;
        } else {
            d = sqrt(-d);
            double e = pow(d + fabs(R), 1. / 3);
            if (R > 0) e = -e;
            x0 = (e + Q / e) - a1 * (1. / 3);
            n = 1;
        }
    }
    r[0] = x0; r[1] = x1; r[2] = x2;
    return n;
}

// dst64 (optional): the same model in full double precision, 9 per model (NaN = no model) -- what the N == 7 case of
// cv::findFundamentalMat returns to the caller (the RANSAC / LMedS scorers read the f32 copy)
__device__ __forceinline__ void store_model(float *dst, const double (&F)[9], bool valid, double *dst64 = nullptr)
{
#pragma unroll
    for (int i = 0; i < 9; ++i) dst[i] = valid ? (float)F[i] : __int_as_float(0x7fc00000);
    dst[9] = dst[10] = dst[11] = 0.f;
    if (dst64) {
#pragma unroll
        for (int i = 0; i < 9; ++i) dst64[i] = valid ? F[i] : __longlong_as_double(0x7ff8000000000000ll);
    }
}

// Device-side point count (asynchronous pair pipeline, pm_match_estimate_*_dev): when n_dev is given, the
// number of correspondences is *n_dev (written by the ratio filter earlier in the stream), bounded by the
// host's n, which only sizes the launch.
__device__ __forceinline__ int eff_n(int n, const int32_t *n_dev)
{
    return n_dev ? min(n, max(*n_dev, 0)) : n;
}

// Grouped launches of the pair pipeline (BASELINE config 5): several image pairs share ONE launch of every RANSAC kernel --
// the pair is a grid dimension and every per-pair array advances by its stride.  s[k] is the element stride of the k-th
// pointer argument of the kernel in declaration order (zero strides + a unit grid dimension = the single-problem call).
struct PairStride { long long s[8]; };
static const PairStride PS0 = {{0, 0, 0, 0, 0, 0, 0, 0}};

// =====================================================================================
// K6: minimal solvers, 8 lanes per hypothesis
// =====================================================================================
constexpr int SOLVE_THREADS = 256;    // 32 hypotheses per block
template <int M>
__global__ void __launch_bounds__(SOLVE_THREADS)
ransac_solve_kernel(const float2 *__restrict__ p1, const float2 *__restrict__ p2, int n,
                    const int32_t *__restrict__ samples, int n_hyp, float *__restrict__ Fout, const int32_t *n_dev,
                    double *__restrict__ Fout64, double *__restrict__ hand, PairStride ps)
{
    { const long long pb = blockIdx.y; p1 += pb * ps.s[0]; p2 += pb * ps.s[1]; samples += pb * ps.s[2]; Fout += pb * ps.s[3]; if (n_dev) n_dev += pb * ps.s[4];
      if (hand) hand += pb * (long long)n_hyp * 16; }
    n = eff_n(n, n_dev);
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int h = gtid >> 3, sub = threadIdx.x & 7;
    const bool live = h < n_hyp;
    const bool has_pt = live && sub < M && n >= M;       // fewer points than a minimal sample: no model
    double x1 = 0, y1 = 0, x2 = 0, y2 = 0;
    if (has_pt) {
        int k = samples[(size_t)h * M + sub];
        k = min(max(k, 0), n - 1);
        const float2 a = p1[k], b = p2[k];
        x1 = a.x; y1 = a.y; x2 = b.x; y2 = b.y;
    }
    bool valid = n >= M;
    double c1x = 0, c1y = 0, c2x = 0, c2y = 0, s1 = 1, s2 = 1;
    if (M == 8) {
        // Hartley normalisation: centroid to origin, mean distance sqrt(2)
        c1x = grp_sum8(x1) * 0.125; c1y = grp_sum8(y1) * 0.125;
        c2x = grp_sum8(x2) * 0.125; c2y = grp_sum8(y2) * 0.125;
        double d1 = sqrt((x1 - c1x) * (x1 - c1x) + (y1 - c1y) * (y1 - c1y));
        double d2 = sqrt((x2 - c2x) * (x2 - c2x) + (y2 - c2y) * (y2 - c2y));
        d1 = grp_sum8(d1) * 0.125; d2 = grp_sum8(d2) * 0.125;
        if (d1 < FLT_EPSILON || d2 < FLT_EPSILON) valid = false;
        s1 = sqrt(2.) / d1; s2 = sqrt(2.) / d2;
        x1 = (x1 - c1x) * s1; y1 = (y1 - c1y) * s1;
        x2 = (x2 - c2x) * s2; y2 = (y2 - c2y) * s2;
    }
    // column `sub` of the 9 x M matrix = this sample's epipolar constraint row
    double c[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0};
    if (sub >= M) {
#pragma unroll
        for (int i = 0; i < 9; ++i) c[i] = 0;
    }
    // Householder QR across lanes; lane k keeps its reflector w (entries k..8) and beta
    double w[9], beta = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) w[i] = 0;
#pragma unroll
    for (int k = 0; k < M; ++k) {
        double v[9];
        double nrm2 = 0;
#pragma unroll
        for (int i = k; i < 9; ++i) nrm2 += c[i] * c[i];
        const double alpha = c[k] >= 0 ? -sqrt(nrm2) : sqrt(nrm2);
#pragma unroll
        for (int i = k; i < 9; ++i) v[i] = c[i];
        v[k] -= alpha;
        double vtv = 0;
#pragma unroll
        for (int i = k; i < 9; ++i) vtv += v[i] * v[i];
        double b = vtv > 0 ? 2.0 / vtv : 0.0;
        // broadcast lane k's reflector to its group
#pragma unroll
        for (int i = k; i < 9; ++i) v[i] = __shfl_sync(0xffffffffu, v[i], k, 8);
        b = __shfl_sync(0xffffffffu, b, k, 8);
        if (sub == k) {
#pragma unroll
            for (int i = k; i < 9; ++i) w[i] = v[i];
            beta = b;
        }
        double s = 0;
#pragma unroll
        for (int i = k; i < 9; ++i) s += v[i] * c[i];
        s *= b;
#pragma unroll
        for (int i = k; i < 9; ++i) c[i] -= s * v[i];
    }
    // null-space basis: z = Q e_j = H_0 H_1 ... H_{M-1} e_j, j = 8 (and 7 for the 7-point)
    double z[9], z2[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { z[i] = i == 8 ? 1.0 : 0.0; z2[i] = i == 7 ? 1.0 : 0.0; }
#pragma unroll
    for (int k = M - 1; k >= 0; --k) {
        double v[9];
#pragma unroll
        for (int i = k; i < 9; ++i) v[i] = __shfl_sync(0xffffffffu, w[i], k, 8);
        const double b = __shfl_sync(0xffffffffu, beta, k, 8);
        double s = 0, t = 0;
#pragma unroll
        for (int i = k; i < 9; ++i) { s += v[i] * z[i]; t += v[i] * z2[i]; }
        s *= b; t *= b;
#pragma unroll
        for (int i = k; i < 9; ++i) { z[i] -= s * v[i]; if (M == 7) z2[i] -= t * v[i]; }
    }
    if (M == 8) {
        // The rank-2 projection (3x3 Jacobi with FP64 divides and square roots) is the long, serial tail of the solve and
        // needs ONE lane per hypothesis: it runs as a kernel of its own (ransac_project_kernel, one thread per hypothesis)
        // so that the 8-lane blocks here retire as soon as the null vector is known instead of idling behind one warp.
        if (live && sub == 0) {
            double *hnd = hand + (size_t)h * 16;
#pragma unroll
            for (int i = 0; i < 9; ++i) hnd[i] = z[i];
            hnd[9] = s1; hnd[10] = s2; hnd[11] = c1x; hnd[12] = c1y; hnd[13] = c2x; hnd[14] = c2y; hnd[15] = valid ? 1.0 : 0.0;
        }
    } else {
        if (!live || sub != 0) return;
        // 7-point: det(lambda*f1 + (1-lambda)*f2) = 0  (OpenCV run7Point)
        double f1[9], f2[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { f1[i] = z2[i]; f2[i] = z[i]; f1[i] -= f2[i]; }
        double cc[4], t0, t1, t2;
        t0 = f2[4] * f2[8] - f2[5] * f2[7];
        t1 = f2[3] * f2[8] - f2[5] * f2[6];
        t2 = f2[3] * f2[7] - f2[4] * f2[6];
        cc[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
        cc[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 -
                f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
                f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
                f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
                f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) -
                f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
                f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
        t0 = f1[4] * f1[8] - f1[5] * f1[7];
        t1 = f1[3] * f1[8] - f1[5] * f1[6];
        t2 = f1[3] * f1[7] - f1[4] * f1[6];
        cc[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 -
                f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
                f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
                f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
                f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) -
                f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
                f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
        cc[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
        double r[3];
        int nr = solve_cubic(cc, r);
        if (nr < 1 || nr > 3) nr = 0;
        for (int k = 0; k < 3; ++k) {
            double F[9];
            bool ok = k < nr;
            if (ok) {
                double lambda = r[k], mu = 1;
                const double sc = f1[8] * r[k] + f2[8];
                if (fabs(sc) > DBL_EPSILON) { mu = 1. / sc; lambda *= mu; F[8] = 1; }
                else F[8] = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) F[i] = f1[i] * lambda + f2[i] * mu;
#pragma unroll
                for (int i = 0; i < 9; ++i) ok = ok && isfinite(F[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 9; ++i) F[i] = 0;
            }
            store_model(Fout + ((size_t)h * 3 + k) * 12, F, ok, Fout64 ? Fout64 + ((size_t)h * 3 + k) * 9 : nullptr);
        }
    }
}

// Second half of the 8-point solve: rank-2 projection of the null vector's 3x3, de-normalisation, F[8] = 1.  One thread per
// hypothesis (hand[h] = {z[9], s1, s2, c1x, c1y, c2x, c2y, valid}); blockIdx.y = pair of a grouped launch.
__global__ void __launch_bounds__(128)
ransac_project_kernel(const double *__restrict__ hand, int n_hyp, float *__restrict__ Fout, double *__restrict__ Fout64, long long f_stride)
{
    const int hh = blockIdx.x * blockDim.x + threadIdx.x;
    if (hh >= n_hyp) return;
    hand += (size_t)blockIdx.y * n_hyp * 16 + (size_t)hh * 16;
    Fout += (long long)blockIdx.y * f_stride;
    double F0[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) F0[i] = hand[i];
    const double s1 = hand[9], s2 = hand[10], c1x = hand[11], c1y = hand[12], c2x = hand[13], c2y = hand[14];
    bool valid = hand[15] != 0.0;
    rank2_project3(F0);
    // F = T2^T F0 T1, T = [s 0 -s cx; 0 s -s cy; 0 0 1]
    double Mx[9];
    Mx[0] = s2 * F0[0]; Mx[1] = s2 * F0[1]; Mx[2] = s2 * F0[2];
    Mx[3] = s2 * F0[3]; Mx[4] = s2 * F0[4]; Mx[5] = s2 * F0[5];
    Mx[6] = -s2 * c2x * F0[0] - s2 * c2y * F0[3] + F0[6];
    Mx[7] = -s2 * c2x * F0[1] - s2 * c2y * F0[4] + F0[7];
    Mx[8] = -s2 * c2x * F0[2] - s2 * c2y * F0[5] + F0[8];
    double F[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        F[i * 3 + 0] = Mx[i * 3 + 0] * s1;
        F[i * 3 + 1] = Mx[i * 3 + 1] * s1;
        F[i * 3 + 2] = -Mx[i * 3 + 0] * s1 * c1x - Mx[i * 3 + 1] * s1 * c1y + Mx[i * 3 + 2];
    }
    if (fabs(F[8]) > FLT_EPSILON) {
        const double inv = 1.0 / F[8];
#pragma unroll
        for (int i = 0; i < 9; ++i) F[i] *= inv;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) valid = valid && isfinite(F[i]);
    store_model(Fout + (size_t)hh * 12, F, valid, Fout64 ? Fout64 + (size_t)hh * 9 : nullptr);
}

// =====================================================================================
// K7: scoring
// =====================================================================================
constexpr int SC_THREADS = 128;
#ifndef SC_MPT_V
#define SC_MPT_V 4
#endif
#ifndef SC_UNROLL_V
#define SC_UNROLL_V 16
#endif
#define PM_STR2(x) #x
#define PM_STR(x) PM_STR2(x)
#define PM_UNROLL(n) _Pragma(PM_STR(unroll n))
constexpr int SC_MPT = SC_MPT_V;     // models per thread
constexpr int SC_TILE = 512;       // correspondences per shared-memory tile (8 KB)

__global__ void pack_points_kernel(const float2 *__restrict__ p1, const float2 *__restrict__ p2, int n,
                                   float4 *__restrict__ out, const int32_t *n_dev)
{
    n = eff_n(n, n_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 a = p1[i], b = p2[i];
    out[i] = make_float4(a.x, a.y, b.x, b.y);
}

// The FP32 inlier test, operation for operation (DESIGN.md "scoring op order"): the margin
//   Sampson: m = fma(thr2, den, -r*r)            inlier <=> m >= 0
//   sym-epi: m1 = fma(thr2, a*a+b*b, -r*r), m2 likewise with the transposed line; inlier <=> both >= 0
// is returned as float bits whose SIGN bit says "outlier" (m is never -0: x - x rounds to +0).
template <int METRIC>
__device__ __forceinline__ unsigned outlier_bits(const float (&F)[9], const float4 p, const float thr2)
{
    const float a = fmaf(F[0], p.x, fmaf(F[1], p.y, F[2]));
    const float b = fmaf(F[3], p.x, fmaf(F[4], p.y, F[5]));
    const float c = fmaf(F[6], p.x, fmaf(F[7], p.y, F[8]));
    const float r = fmaf(p.z, a, fmaf(p.w, b, c));
    const float at = fmaf(F[0], p.z, fmaf(F[3], p.w, F[6]));
    const float bt = fmaf(F[1], p.z, fmaf(F[4], p.w, F[7]));
    const float nr2 = -__fmul_rn(r, r);
    if (METRIC == PM_METRIC_SAMPSON) {
        const float den = fmaf(a, a, fmaf(b, b, fmaf(at, at, __fmul_rn(bt, bt))));
        return __float_as_uint(fmaf(thr2, den, nr2));
    } else {
        const float n2 = fmaf(a, a, __fmul_rn(b, b));
        const float n1 = fmaf(at, at, __fmul_rn(bt, bt));
        return __float_as_uint(fmaf(thr2, n2, nr2)) | __float_as_uint(fmaf(thr2, n1, nr2));
    }
}
template <int METRIC>
__device__ __forceinline__ bool is_inlier(const float (&F)[9], const float4 p, const float thr2)
{
    return (outlier_bits<METRIC>(F, p, thr2) >> 31) == 0u;
}
// out += bits >> 31 in ONE instruction on the FMA pipe (IMAD.HI: hi32(bits * 2) + out) instead of FSETP + IADD
__device__ __forceinline__ unsigned add_sign(unsigned bits, unsigned out)
{
    unsigned r;
    asm("mad.hi.u32 %0, %1, 2, %2;" : "=r"(r) : "r"(bits), "r"(out));
    return r;
}
// A model with a NaN / infinite coefficient (degenerate sample) is replaced by F = e9: a = b = a' = b' = 0, r = 1,
// margin = -1 for every point -> count 0, exactly what "no model" must score.
__device__ __forceinline__ void sanitize_model(float (&F)[9])
{
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 9; ++i) ok = ok && (fabsf(F[i]) < 3.0e38f);
    if (!ok) {
#pragma unroll
        for (int i = 0; i < 8; ++i) F[i] = 0.f;
        F[8] = 1.f;
    }
}

__device__ __forceinline__ void sc_cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}

template <int METRIC, int MPT>
__global__ void __launch_bounds__(SC_THREADS)
ransac_score_kernel(const float4 *__restrict__ pts, int n, int chunk_pts, const float *__restrict__ Fm,
                    int n_models, float thr2, int32_t *__restrict__ counts, int use_atomic, const int32_t *n_dev, PairStride ps)
{
    __shared__ __align__(16) float4 tile[2][SC_TILE];
    { const long long pb = blockIdx.z; pts += pb * ps.s[0]; Fm += pb * ps.s[1]; counts += pb * ps.s[2]; if (n_dev) n_dev += pb * ps.s[3]; }
    n = eff_n(n, n_dev);
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * (SC_THREADS * MPT);
    float F[MPT][9];
#pragma unroll
    for (int r = 0; r < MPT; ++r) {
        const int m = m0 + r * SC_THREADS + tid;
        const float4 *src = reinterpret_cast<const float4 *>(Fm + (size_t)min(m, n_models - 1) * 12);
        const float4 u = __ldg(src), v = __ldg(src + 1), w = __ldg(src + 2);
        F[r][0] = u.x; F[r][1] = u.y; F[r][2] = u.z; F[r][3] = u.w;
        F[r][4] = v.x; F[r][5] = v.y; F[r][6] = v.z; F[r][7] = v.w; F[r][8] = w.x;
    }
#pragma unroll
    for (int r = 0; r < MPT; ++r) sanitize_model(F[r]);
    unsigned outl[MPT];                   // OUTLIERS seen so far (one IMAD.HI per evaluation)
#pragma unroll
    for (int r = 0; r < MPT; ++r) outl[r] = 0u;

    const int p0 = blockIdx.y * chunk_pts, p1 = min(n, p0 + chunk_pts);
    const int ntiles = (p1 - p0 + SC_TILE - 1) / SC_TILE;
    auto load_tile = [&](int ti, int buf) {
        const int b = p0 + ti * SC_TILE;
        const int cntp = min(SC_TILE, p1 - b);
        for (int v = tid; v < cntp; v += SC_THREADS) sc_cp_async16(&tile[buf][v], pts + b + v);
        asm volatile("cp.async.commit_group;");
    };
    if (ntiles > 0) load_tile(0, 0);
    for (int ti = 0; ti < ntiles; ++ti) {
        const int buf = ti & 1;
        if (ti + 1 < ntiles) { load_tile(ti + 1, buf ^ 1); asm volatile("cp.async.wait_group 1;"); }
        else asm volatile("cp.async.wait_group 0;");
        __syncthreads();
        const int cntp = min(SC_TILE, p1 - (p0 + ti * SC_TILE));
        PM_UNROLL(SC_UNROLL_V)
        for (int j = 0; j < cntp; ++j) {
            const float4 p = tile[buf][j];
#pragma unroll
            for (int r = 0; r < MPT; ++r) outl[r] = add_sign(outlier_bits<METRIC>(F[r], p, thr2), outl[r]);
        }
        __syncthreads();
    }
    const int seen = max(p1 - p0, 0);
#pragma unroll
    for (int r = 0; r < MPT; ++r) {
        const int m = m0 + r * SC_THREADS + tid;
        if (m < n_models) {
            const int c = seen - (int)outl[r];
            if (use_atomic) atomicAdd(&counts[m], c);
            else counts[m] = c;
        }
    }
}

// key = count << 32 | (0xFFFFFFFF - model_id): max key = max count, lowest id on ties
__global__ void __launch_bounds__(256)
ransac_best_kernel(const int32_t *__restrict__ counts, int n_models, int id_base, unsigned long long *key)
{
    unsigned long long best = 0;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n_models; m += gridDim.x * blockDim.x) {
        const int c = counts[m];
        if (c > 0) {
            const unsigned long long k = ((unsigned long long)(unsigned)c << 32) | (0xFFFFFFFFu - (unsigned)(id_base + m));
            best = k > best ? k : best;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o);
        best = y > best ? y : best;
    }
    if ((threadIdx.x & 31) == 0 && best) atomicMax(key, best);
}

// Small model sets (the pair pipeline: a few thousand hypotheses): ONE block finds the winner key, copies the
// winning model and zeroes the inlier counter -- one launch instead of memset + best + pick + memset.  Same key,
// same winner as ransac_best_kernel + ransac_pick_kernel (max over the same keys).
__global__ void __launch_bounds__(1024)
ransac_best_pick_kernel(const int32_t *__restrict__ counts, int n_models, int id_base, const float *__restrict__ Fm,
                        unsigned long long *key_out, float *Fw, int32_t *n_inl, PairStride ps)
{
    __shared__ unsigned long long sh[32];
    { const long long pb = blockIdx.x; counts += pb * ps.s[0]; Fm += pb * ps.s[1]; key_out += pb * ps.s[2]; Fw += pb * ps.s[3]; if (n_inl) n_inl += pb * ps.s[4]; }
    unsigned long long best = 0;
    for (int m = threadIdx.x; m < n_models; m += blockDim.x) {
        const int c = counts[m];
        if (c > 0) {
            const unsigned long long k = ((unsigned long long)(unsigned)c << 32) | (0xFFFFFFFFu - (unsigned)(id_base + m));
            best = k > best ? k : best;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o);
        best = y > best ? y : best;
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        best = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o);
            best = y > best ? y : best;
        }
        if (threadIdx.x == 0) { *key_out = best; if (n_inl) *n_inl = 0; }
        if (threadIdx.x < 12) {
            const long long m = (long long)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFu)) - id_base;
            Fw[threadIdx.x] = (best != 0 && m >= 0 && m < n_models) ? Fm[(size_t)m * 12 + threadIdx.x] : __int_as_float(0x7fc00000);
        }
    }
}

__global__ void ransac_pick_kernel(const unsigned long long *key, const float *__restrict__ Fm, int id_base,
                                   int n_models, float *Fw)
{
    const int i = threadIdx.x;
    if (i >= 12) return;
    const unsigned long long k = *key;
    const long long m = (long long)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFu)) - id_base;
    Fw[i] = (k != 0 && m >= 0 && m < n_models) ? Fm[(size_t)m * 12 + i] : __int_as_float(0x7fc00000);
}

// Batched RANSAC with an adaptive stop (pm_find_fundamental_adaptive): after each batch the running winner
// (*best_key, Fw_best) absorbs the batch winner.  Model ids are global, so the plain maximum of the keys keeps
// "most inliers, lowest model id on ties" across batches.
__global__ void ransac_update_best_kernel(const unsigned long long *key, const float *__restrict__ Fm, int id_base,
                                          int n_models, unsigned long long *best_key, float *Fw_best)
{
    const unsigned long long k = *key, b = *best_key;
    const long long m = (long long)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFu)) - id_base;
    const bool take = k > b && m >= 0 && m < n_models;
    __syncwarp();
    if (take) {
        if (threadIdx.x < 12) Fw_best[threadIdx.x] = Fm[(size_t)m * 12 + threadIdx.x];
        if (threadIdx.x == 0) *best_key = k;
    }
}

// =====================================================================================
// K8: mask + refit
// =====================================================================================
template <int METRIC>
__global__ void __launch_bounds__(256)
ransac_mask_kernel(const float4 *__restrict__ pts, int n, const float *__restrict__ Fw, float thr2,
                   uint8_t *__restrict__ mask, int32_t *n_inl, const int32_t *n_dev, PairStride ps)
{
    { const long long pb = blockIdx.y; pts += pb * ps.s[0]; Fw += pb * ps.s[1]; mask += pb * ps.s[2]; n_inl += pb * ps.s[3]; if (n_dev) n_dev += pb * ps.s[4]; }
    n = eff_n(n, n_dev);
    float F[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = Fw[i];
    sanitize_model(F);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool in = false;
    if (i < n) { in = is_inlier<METRIC>(F, pts[i], thr2); mask[i] = in ? 1 : 0; }
    const int c = __syncthreads_count(in);
    if (threadIdx.x == 0 && c) atomicAdd(n_inl, c);
}

constexpr int RF_BLOCKS = 64;      // fixed grid -> deterministic two-level reductions
constexpr int RF_THREADS = 256;

template <int NV>
__device__ void block_reduce_store(double (&v)[NV], double *dst)
{
    __shared__ double sh[RF_THREADS / 32][NV];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh[w][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
        for (int ww = 0; ww < RF_THREADS / 32; ++ww) s += sh[ww][threadIdx.x];
        dst[threadIdx.x] = s;
    }
    __syncthreads();
}

// pass 1: sums of x1,y1,x2,y2 and the count over the selected points
__global__ void __launch_bounds__(RF_THREADS)
refit_sum_kernel(const float4 *__restrict__ pts, int n, const uint8_t *__restrict__ mask, double *partial, const int32_t *n_dev,
                 PairStride ps)
{
    { const long long pb = blockIdx.y; pts += pb * ps.s[0]; if (mask) mask += pb * ps.s[1]; partial += pb * ps.s[2]; if (n_dev) n_dev += pb * ps.s[3]; }
    n = eff_n(n, n_dev);
    double v[5] = {0, 0, 0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (!mask || mask[i]) { const float4 p = pts[i]; v[0] += p.x; v[1] += p.y; v[2] += p.z; v[3] += p.w; v[4] += 1.0; }
    block_reduce_store<5>(v, partial + blockIdx.x * 5);
}
// stats: [0..3] centroid, [4] count, [5] s1, [6] s2
// The finalize steps between the passes (centroids from pass 1's partial sums, scales from pass 2's) are a few dozen
// additions: every block of the NEXT pass repeats them itself, entry k summed over the blocks in ascending order,
// and block 0 publishes them -- no single-warp kernels between the passes.
__device__ __forceinline__ void refit_mean_inblock(const double *__restrict__ partial, double *sh, double *stats_out)
{
    const int k = threadIdx.x;
    if (k < 32) {
        double s = 0;
        if (k < 5)
            for (int b = 0; b < RF_BLOCKS; ++b) s += partial[b * 5 + k];
        const double cnt = __shfl_sync(0xffffffffu, s, 4);
        const double inv = cnt > 0 ? 1.0 / cnt : 0.0;
        if (k < 4) sh[k] = s * inv;
        if (k == 4) sh[4] = s;
        if (blockIdx.x == 0 && k < 5) stats_out[k] = k < 4 ? s * inv : s;
    }
    __syncthreads();
}
__device__ __forceinline__ void refit_scale_inblock(const double *__restrict__ partial, const double cnt, double *sh,
                                                    double *stats_out)
{
    const int e = threadIdx.x;
    if (e < 2) {
        double a = 0;
        for (int k = 0; k < RF_BLOCKS; ++k) a += partial[k * 2 + e];
        a = cnt > 0 ? a / cnt : 0;
        const double sc = a >= FLT_EPSILON ? sqrt(2.) / a : 0.0;
        sh[5 + e] = sc;
        if (blockIdx.x == 0) stats_out[5 + e] = sc;
    }
    __syncthreads();
}

// pass 2: mean distances to the centroids.  fused != 0: the centroids come from pass 1's partial sums (partial_in)
__global__ void __launch_bounds__(RF_THREADS)
refit_scale_kernel(const float4 *__restrict__ pts, int n, const uint8_t *__restrict__ mask,
                   double *__restrict__ stats, double *partial, const int32_t *n_dev, const double *partial_in, PairStride ps)
{
    __shared__ double sst[8];
    { const long long pb = blockIdx.y; pts += pb * ps.s[0]; if (mask) mask += pb * ps.s[1]; stats += pb * ps.s[2]; partial += pb * ps.s[3];
      if (n_dev) n_dev += pb * ps.s[4]; if (partial_in) partial_in += pb * ps.s[5]; }
    n = eff_n(n, n_dev);
    if (partial_in) refit_mean_inblock(partial_in, sst, stats);
    else { if (threadIdx.x < 5) sst[threadIdx.x] = stats[threadIdx.x]; __syncthreads(); }
    const double c1x = sst[0], c1y = sst[1], c2x = sst[2], c2y = sst[3];
    double v[2] = {0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (!mask || mask[i]) {
            const float4 p = pts[i];
            v[0] += sqrt((p.x - c1x) * (p.x - c1x) + (p.y - c1y) * (p.y - c1y));
            v[1] += sqrt((p.z - c2x) * (p.z - c2x) + (p.w - c2y) * (p.w - c2y));
        }
    block_reduce_store<2>(v, partial + blockIdx.x * 2);
}
// pass 3: upper triangle of the 9x9 normal matrix sum r r^T (45 entries)
__global__ void __launch_bounds__(RF_THREADS)
refit_ata_kernel(const float4 *__restrict__ pts, int n, const uint8_t *__restrict__ mask,
                 double *__restrict__ stats, double *partial, const int32_t *n_dev, const double *partial_in, PairStride ps)
{
    __shared__ double sst[8];
    { const long long pb = blockIdx.y; pts += pb * ps.s[0]; if (mask) mask += pb * ps.s[1]; stats += pb * ps.s[2]; partial += pb * ps.s[3];
      if (n_dev) n_dev += pb * ps.s[4]; if (partial_in) partial_in += pb * ps.s[5]; }
    n = eff_n(n, n_dev);
    if (threadIdx.x < 5) sst[threadIdx.x] = stats[threadIdx.x];      // centroids and count (pass 2's block 0 wrote them)
    __syncthreads();
    if (partial_in) refit_scale_inblock(partial_in, sst[4], sst, stats);
    else { if (threadIdx.x < 2) sst[5 + threadIdx.x] = stats[5 + threadIdx.x]; __syncthreads(); }
    const double c1x = sst[0], c1y = sst[1], c2x = sst[2], c2y = sst[3], s1 = sst[5], s2 = sst[6];
    double acc[45];
#pragma unroll
    for (int k = 0; k < 45; ++k) acc[k] = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (!mask || mask[i]) {
            const float4 p = pts[i];
            const double x1 = (p.x - c1x) * s1, y1 = (p.y - c1y) * s1, x2 = (p.z - c2x) * s2, y2 = (p.w - c2y) * s2;
            const double r[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0};
            int k = 0;
#pragma unroll
            for (int a = 0; a < 9; ++a)
#pragma unroll
                for (int b = a; b < 9; ++b) acc[k++] += r[a] * r[b];
        }
    block_reduce_store<45>(acc, partial + blockIdx.x * 45);
}

// The pair pipeline's result record, written by the last kernel of the chain (null res: none).
struct PairOut {
    const unsigned long long *key; const int32_t *n_good; const int32_t *n_inl; int n_max; int m; pm_pair_result *res;
};
__device__ __forceinline__ void write_pair_result(const PairOut &po, const double *F)
{
    const unsigned long long k = *po.key;
    const int n = min(max(*po.n_good, 0), po.n_max);
    const bool ok = k != 0ull && n >= po.m;
    for (int i = 0; i < 9; ++i) po.res->F[i] = ok ? F[i] : 0.0;
    po.res->key = ok ? k : 0ull; po.res->n_matches = n; po.res->n_inliers = ok ? *po.n_inl : 0;
    po.res->has_model = ok; po.res->reserved = 0;
}

// no refit: the record carries the winning minimal model itself (one warp per pair of a grouped launch)
__global__ void pair_group_result_kernel(const float *__restrict__ Fw, long long fw_stride, PairOut po, PairStride ps)
{
    const long long pb = blockIdx.x;
    Fw += pb * fw_stride;
    po.key += pb * ps.s[4]; po.n_good += pb * ps.s[5]; po.n_inl += pb * ps.s[6]; po.res += pb * ps.s[7];
    if (threadIdx.x != 0) return;
    double F[9];
    for (int i = 0; i < 9; ++i) F[i] = (double)Fw[i];
    write_pair_result(po, F);
}

// pass 4 (one warp): 9x9 cyclic Jacobi (lanes 0..8 each own index k of the rotation
// updates), smallest eigenvector, rank-2 projection, de-normalisation.  Falls back to the
// winning minimal model when there are < 8 points or the system is degenerate.
__global__ void __launch_bounds__(32)
refit_solve_kernel(const double *__restrict__ partial, const double *__restrict__ stats,
                   const float *__restrict__ Ffallback, double *__restrict__ Fout, int32_t *ok_out, PairOut po, PairStride ps)
{
    __shared__ double A[81], V[81];
    {   // one warp (block) per pair in a grouped launch
        const long long pb = blockIdx.x;
        partial += pb * ps.s[0]; stats += pb * ps.s[1]; if (Ffallback) Ffallback += pb * ps.s[2]; Fout += pb * ps.s[3];
        if (po.res) { po.key += pb * ps.s[4]; po.n_good += pb * ps.s[5]; po.n_inl += pb * ps.s[6]; po.res += pb * ps.s[7]; }
    }
    const int lane = threadIdx.x;
    for (int e = lane; e < 81; e += 32) { A[e] = 0; V[e] = (e % 10 == 0) ? 1.0 : 0.0; }
    __syncwarp();
    for (int k = lane; k < 45; k += 32) {       // entry k of the upper triangle, blocks summed in ascending order
        int a = 0, rem = k;
        while (rem >= 9 - a) { rem -= 9 - a; ++a; }
        const int b = a + rem;
        double s = 0;
        for (int blk = 0; blk < RF_BLOCKS; ++blk) s += partial[blk * 45 + k];
        A[a * 9 + b] = s; A[b * 9 + a] = s;
    }
    __syncwarp();
    const double cnt = stats[4], s1 = stats[5], s2 = stats[6];
    bool ok = cnt >= 8 && s1 > 0 && s2 > 0;
    for (int sweep = 0; sweep < 60 && ok; ++sweep) {
        double off = 0, diag = 0;
        for (int i = 0; i < 9; ++i) {
            diag += A[i * 9 + i] * A[i * 9 + i];
            for (int j = i + 1; j < 9; ++j) off += A[i * 9 + j] * A[i * 9 + j];
        }
        if (off <= 1e-34 * diag || off == 0) break;
        // parallel-order Jacobi: round r rotates the four disjoint pairs {i, j} with i + j = r (mod 9) at once
        // (every pair appears in exactly one of the nine rounds); lane g < 4 computes the rotation of pair g
        for (int r = 0; r < 9; ++r) {
            int p = 0, q = 0;
            double c = 1.0, s = 0.0;
            if (lane < 4) {
                int seen = 0;
                for (int i = 0; i < 9; ++i) {
                    const int j = (r - i + 9) % 9;
                    if (i < j) { if (seen == lane) { p = i; q = j; } ++seen; }
                }
                const double apq = A[p * 9 + q];
                if (apq != 0) {
                    const double theta = (A[q * 9 + q] - A[p * 9 + p]) / (2 * apq);
                    const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                    c = 1 / sqrt(tt * tt + 1); s = tt * c;
                }
            }
            int pg[4], qg[4]; double cg[4], sg[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                pg[g] = __shfl_sync(0xffffffffu, p, g); qg[g] = __shfl_sync(0xffffffffu, q, g);
                cg[g] = __shfl_sync(0xffffffffu, c, g); sg[g] = __shfl_sync(0xffffffffu, s, g);
            }
            __syncwarp();
            if (lane < 9) {
                const int k = lane;
#pragma unroll
                for (int g = 0; g < 4; ++g) {      // columns p, q of row k (disjoint across the four pairs)
                    const double akp = A[k * 9 + pg[g]], akq = A[k * 9 + qg[g]];
                    A[k * 9 + pg[g]] = cg[g] * akp - sg[g] * akq; A[k * 9 + qg[g]] = sg[g] * akp + cg[g] * akq;
                }
            }
            __syncwarp();
            if (lane < 9) {
                const int k = lane;
#pragma unroll
                for (int g = 0; g < 4; ++g) {      // rows p, q (and the eigenvector accumulation)
                    const double apk = A[pg[g] * 9 + k], aqk = A[qg[g] * 9 + k];
                    A[pg[g] * 9 + k] = cg[g] * apk - sg[g] * aqk; A[qg[g] * 9 + k] = sg[g] * apk + cg[g] * aqk;
                    const double vpk = V[pg[g] * 9 + k], vqk = V[qg[g] * 9 + k];
                    V[pg[g] * 9 + k] = cg[g] * vpk - sg[g] * vqk; V[qg[g] * 9 + k] = sg[g] * vpk + cg[g] * vqk;
                }
            }
            __syncwarp();
        }
    }
    if (lane != 0) return;
    double F[9];
    if (ok) {
        int mn = 0, nsmall = 0;
        for (int i = 0; i < 9; ++i) {
            if (A[i * 9 + i] < A[mn * 9 + mn]) mn = i;
            if (fabs(A[i * 9 + i]) < DBL_EPSILON) ++nsmall;
        }
        if (nsmall > 1) ok = false;        // rank < 8 (OpenCV: any of the 8 largest < DBL_EPSILON)
        double F0[9];
        for (int i = 0; i < 9; ++i) F0[i] = V[mn * 9 + i];
        rank2_project3(F0);
        const double c1x = stats[0], c1y = stats[1], c2x = stats[2], c2y = stats[3];
        double Mx[9];
        Mx[0] = s2 * F0[0]; Mx[1] = s2 * F0[1]; Mx[2] = s2 * F0[2];
        Mx[3] = s2 * F0[3]; Mx[4] = s2 * F0[4]; Mx[5] = s2 * F0[5];
        Mx[6] = -s2 * c2x * F0[0] - s2 * c2y * F0[3] + F0[6];
        Mx[7] = -s2 * c2x * F0[1] - s2 * c2y * F0[4] + F0[7];
        Mx[8] = -s2 * c2x * F0[2] - s2 * c2y * F0[5] + F0[8];
        for (int i = 0; i < 3; ++i) {
            F[i * 3 + 0] = Mx[i * 3 + 0] * s1;
            F[i * 3 + 1] = Mx[i * 3 + 1] * s1;
            F[i * 3 + 2] = -Mx[i * 3 + 0] * s1 * c1x - Mx[i * 3 + 1] * s1 * c1y + Mx[i * 3 + 2];
        }
        if (fabs(F[8]) > FLT_EPSILON) { const double inv = 1.0 / F[8]; for (int i = 0; i < 9; ++i) F[i] *= inv; }
        for (int i = 0; i < 9; ++i) ok = ok && isfinite(F[i]);
    }
    if (!ok && Ffallback) for (int i = 0; i < 9; ++i) F[i] = (double)Ffallback[i];
    if (ok || Ffallback) for (int i = 0; i < 9; ++i) Fout[i] = F[i];
    if (ok_out) *ok_out = ok ? 1 : 0;
    if (po.res) write_pair_result(po, F);      // (F is only read when a model exists, and then ok || Ffallback holds)
}

__global__ void copy_f32_to_f64_kernel(const float *src, double *dst, int n)
{
    if ((int)threadIdx.x < n) dst[threadIdx.x] = (double)src[threadIdx.x];
}

// =====================================================================================
// diagnostics
// =====================================================================================
__global__ void epilines_kernel(const float2 *__restrict__ pts, int n, int which, const double *__restrict__ Fd,
                                float *__restrict__ lines)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double f[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) f[r * 3 + c] = which == 2 ? Fd[c * 3 + r] : Fd[r * 3 + c];
    const double x = pts[i].x, y = pts[i].y;
    const double a = f[0] * x + f[1] * y + f[2], b = f[3] * x + f[4] * y + f[5], c = f[6] * x + f[7] * y + f[8];
    double nu = a * a + b * b;
    nu = nu ? 1. / sqrt(nu) : 1.;
    lines[3 * i] = (float)(a * nu); lines[3 * i + 1] = (float)(b * nu); lines[3 * i + 2] = (float)(c * nu);
}

__global__ void __launch_bounds__(256)
residuals_kernel(const float2 *__restrict__ p1, const float2 *__restrict__ p2, int n, const double *__restrict__ F,
                 int metric, float *__restrict__ out, double *partial)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0;
    if (i < n) {
        const double x1 = p1[i].x, y1 = p1[i].y, x2 = p2[i].x, y2 = p2[i].y;
        const double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5], c = F[6] * x1 + F[7] * y1 + F[8];
        const double r = x2 * a + y2 * b + c;
        const double at = F[0] * x2 + F[3] * y2 + F[6], bt = F[1] * x2 + F[4] * y2 + F[7];
        if (metric == PM_METRIC_SAMPSON) e = r * r / (a * a + b * b + at * at + bt * bt);
        else { const double e2 = r * r / (a * a + b * b), e1 = r * r / (at * at + bt * bt); e = e1 > e2 ? e1 : e2; }
        out[i] = (float)e;
    }
    // mean: fixed-shape reductions only (warp tree, 8 warps in order, then residuals_sum_kernel over the blocks in
    // order), so the last bits do not depend on the order in which blocks finish
    __shared__ double sh[8];
    double x = e;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x == 0 && partial) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += sh[w];
        partial[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(256) residuals_sum_kernel(const double *__restrict__ partial, int nblocks, double *sum)
{
    __shared__ double sh[8];
    double x = 0;
    for (int b = threadIdx.x; b < nblocks; b += 256) x += partial[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += sh[w];
        *sum = s;
    }
}

// =====================================================================================
// LMedS (what cv::findFundamentalMat(..., CV_FM_7POINT) runs when N > 7: the reference's literal call,
// main.cpp:95-98).  Per model: err_i = (float) max(d(x2,Fx1)^2, d(x1,F^T x2)^2) in FP64
// (OpenCV's computeError), median = sorted err[n / 2]; winner = smallest median (lowest model id on
// ties); inliers = err <= (2.5 * 1.4826 * (1 + 5/(n-7)) * sqrt(median))^2.  No fused multiply-adds in the
// error: the oracle (orc_symepi_f64, -ffp-contract=off) is matched bit for bit.
// =====================================================================================
__device__ __forceinline__ float symepi_err_f64(const double (&F)[9], const float4 p)
{
    const double x1 = p.x, y1 = p.y, x2 = p.z, y2 = p.w;
    double a = __dadd_rn(__dadd_rn(__dmul_rn(F[0], x1), __dmul_rn(F[1], y1)), F[2]);
    double b = __dadd_rn(__dadd_rn(__dmul_rn(F[3], x1), __dmul_rn(F[4], y1)), F[5]);
    double c = __dadd_rn(__dadd_rn(__dmul_rn(F[6], x1), __dmul_rn(F[7], y1)), F[8]);
    const double s2 = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(x2, a), __dmul_rn(y2, b)), c);
    a = __dadd_rn(__dadd_rn(__dmul_rn(F[0], x2), __dmul_rn(F[3], y2)), F[6]);
    b = __dadd_rn(__dadd_rn(__dmul_rn(F[1], x2), __dmul_rn(F[4], y2)), F[7]);
    c = __dadd_rn(__dadd_rn(__dmul_rn(F[2], x2), __dmul_rn(F[5], y2)), F[8]);
    const double s1 = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
    const double d1 = __dadd_rn(__dadd_rn(__dmul_rn(x1, a), __dmul_rn(y1, b)), c);
    const double e1 = __dmul_rn(__dmul_rn(d1, d1), s1), e2 = __dmul_rn(__dmul_rn(d2, d2), s2);
    return (float)(e1 > e2 ? e1 : e2);
}

constexpr int LM_THREADS = 256;
// One CTA per model.  errs: scratch [gridDim.x][n].  medians[model] = err bits of rank n/2 (NaN -> +inf
// ordering: a model with a NaN coefficient or error gets median +inf and can never win).
__global__ void __launch_bounds__(LM_THREADS)
lmeds_median_kernel(const float4 *__restrict__ pts, int n, const float *__restrict__ Fm, int model0, int n_models,
                    float *__restrict__ errs, float *__restrict__ medians)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_rank;
    __shared__ int s_bad;
    const int m = model0 + blockIdx.x;
    if (m >= n_models) return;
    double F[9];
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 9; ++i) { const float v = Fm[(size_t)m * 12 + i]; ok = ok && isfinite(v); F[i] = (double)v; }
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    if (!ok) { if (threadIdx.x == 0) medians[m] = __int_as_float(0x7f800000); return; }
    unsigned *e = reinterpret_cast<unsigned *>(errs + (size_t)blockIdx.x * n);
    for (int i = threadIdx.x; i < n; i += LM_THREADS) {
        const float err = symepi_err_f64(F, pts[i]);
        if (!(err >= 0.f)) s_bad = 1;                 // NaN
        e[i] = __float_as_uint(err);                  // err >= 0: the bits order like the values
    }
    __syncthreads();
    if (s_bad) { if (threadIdx.x == 0) medians[m] = __int_as_float(0x7f800000); return; }
    // radix select of rank n/2 (0-based): four 8-bit passes, most significant byte first
    if (threadIdx.x == 0) { s_prefix = 0u; s_rank = (unsigned)(n / 2); }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        hist[threadIdx.x] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix, himask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (int i = threadIdx.x; i < n; i += LM_THREADS) {
            const unsigned v = e[i];
            if ((v & himask) == prefix) atomicAdd(&hist[(v >> shift) & 0xFFu], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned r = s_rank, b = 0;
            for (; b < 255u; ++b) { if (r < hist[b]) break; r -= hist[b]; }
            s_rank = r; s_prefix = prefix | (b << shift);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) medians[m] = __uint_as_float(s_prefix);
}

// key = median bits << 32 | model id: the minimum is the smallest median, lowest id on ties
__global__ void __launch_bounds__(256)
lmeds_best_kernel(const float *__restrict__ medians, int n_models, int id_base, unsigned long long *key)
{
    unsigned long long best = ~0ull;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n_models; m += gridDim.x * blockDim.x) {
        const unsigned b = __float_as_uint(medians[m]);
        if (b < 0x7f800000u) {
            const unsigned long long k = ((unsigned long long)b << 32) | (unsigned)(id_base + m);
            best = k < best ? k : best;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o); best = y < best ? y : best; }
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(key, best);
}

// winner -> Fw (12 floats), F (f64), sigma^2 (f32 at Fw[9]); key == ~0: no model
__global__ void lmeds_pick_kernel(const unsigned long long *key, const float *__restrict__ Fm, int id_base, int n_models,
                                  int n, float *Fw, double *Fout)
{
    if (threadIdx.x != 0) return;
    const unsigned long long k = *key;
    const long long m = (long long)(unsigned)(k & 0xFFFFFFFFull) - id_base;
    if (k == ~0ull || m < 0 || m >= n_models) { for (int i = 0; i < 12; ++i) Fw[i] = __int_as_float(0x7fc00000); return; }
    for (int i = 0; i < 9; ++i) { Fw[i] = Fm[(size_t)m * 12 + i]; Fout[i] = (double)Fw[i]; }
    const double med = (double)__uint_as_float((unsigned)(k >> 32));
    double sigma = 2.5 * 1.4826 * (1 + 5. / (n - 7)) * sqrt(med);
    if (sigma < 0.001) sigma = 0.001;
    Fw[9] = (float)(sigma * sigma);
}

__global__ void __launch_bounds__(256)
lmeds_mask_kernel(const float4 *__restrict__ pts, int n, const float *__restrict__ Fw, uint8_t *__restrict__ mask,
                  int32_t *n_inl)
{
    double F[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = (double)Fw[i];
    const float t = Fw[9];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool in = false;
    if (i < n) { in = symepi_err_f64(F, pts[i]) <= t; mask[i] = in ? 1 : 0; }
    const int c = __syncthreads_count(in);
    if (threadIdx.x == 0 && c) atomicAdd(n_inl, c);
}

int get_pts4(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const float4 **out, const int32_t *dn = nullptr)
{
    PM_WS(ctx, pts, float4 *, WS_MISC, (size_t)(n > 0 ? n : 1) * sizeof(float4));
    if (n > 0) {
        pack_points_kernel<<<pm_cdiv(n, 256), 256, 0, ctx->stream>>>((const float2 *)dp1, (const float2 *)dp2, n, pts, dn);
        PM_CHECK_LAUNCH(ctx);
    }
    *out = pts;
    return PM_OK;
}

int run_refit(pm_ctx *ctx, const float4 *pts, int n, const uint8_t *dmask, const float *dFfallback, double *dF,
              int32_t *dok, const int32_t *dn = nullptr, PairOut po = PairOut{nullptr, nullptr, nullptr, 0, 0, nullptr})
{
    // one partial-sum region per pass (a pass reads the previous pass's partials while it writes its own)
    PM_WS(ctx, ws, double *, WS_REFIT, (size_t)(RF_BLOCKS * (5 + 2 + 45) + 16) * sizeof(double));
    double *part1 = ws, *part2 = ws + RF_BLOCKS * 5, *part3 = part2 + RF_BLOCKS * 2, *stats = part3 + RF_BLOCKS * 45;
    refit_sum_kernel<<<RF_BLOCKS, RF_THREADS, 0, ctx->stream>>>(pts, n, dmask, part1, dn, PS0);
    PM_CHECK_LAUNCH(ctx);
    refit_scale_kernel<<<RF_BLOCKS, RF_THREADS, 0, ctx->stream>>>(pts, n, dmask, stats, part2, dn, part1, PS0);
    PM_CHECK_LAUNCH(ctx);
    refit_ata_kernel<<<RF_BLOCKS, RF_THREADS, 0, ctx->stream>>>(pts, n, dmask, stats, part3, dn, part2, PS0);
    PM_CHECK_LAUNCH(ctx);
    refit_solve_kernel<<<1, 32, 0, ctx->stream>>>(part3, stats, dFfallback, dF, dok, po, PS0);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

}  // namespace

int pmk_ransac_solve(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const int32_t *dsamples, int n_hyp,
                     int m, float *dF32, const int32_t *dn, double *dF64)
{
    if (n_hyp <= 0) return PM_OK;
    const int blocks = pm_cdiv(n_hyp * 8, SOLVE_THREADS);
    if (m == 8) {
        PM_WS(ctx, hand, double *, WS_HAND, (size_t)n_hyp * 16 * sizeof(double));
        ransac_solve_kernel<8><<<blocks, SOLVE_THREADS, 0, ctx->stream>>>((const float2 *)dp1, (const float2 *)dp2, n, dsamples, n_hyp, dF32, dn, dF64, hand, PS0);
        PM_CHECK_LAUNCH(ctx);
        ransac_project_kernel<<<pm_cdiv(n_hyp, 128), 128, 0, ctx->stream>>>(hand, n_hyp, dF32, dF64, 0);
    } else {
        ransac_solve_kernel<7><<<blocks, SOLVE_THREADS, 0, ctx->stream>>>((const float2 *)dp1, (const float2 *)dp2, n, dsamples, n_hyp, dF32, dn, dF64, nullptr, PS0);
    }
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

int pmk_ransac_score(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const float *dF32, int n_models,
                     float thr, int metric, int32_t *dcounts, const int32_t *dn, const float *dpts4)
{
    if (n_models <= 0) return PM_OK;
    const float4 *pts = reinterpret_cast<const float4 *>(dpts4);       // already packed {x1, y1, x2, y2}, or null
    if (!pts) {
        int st = get_pts4(ctx, dp1, dp2, n, &pts, dn);
        if (st != PM_OK) return st;
    }
    // 4 models per thread when there are enough models to fill the GPU; 1 model per thread for small batches
    // (e.g. 4096 hypotheses x 4096 matches per image pair in config 5), where 4 would leave most SMs idle
    int mpt = SC_MPT;
    if (pm_cdiv(n_models, SC_THREADS * SC_MPT) * pm_cdiv(n > 0 ? n : 1, SC_TILE) < 2 * ctx->num_sms) mpt = 1;
    const int mblocks = pm_cdiv(n_models, SC_THREADS * mpt);
    // split the correspondences across blockIdx.y when the model blocks alone would leave the SMs unevenly
    // loaded: aim at >= 8 CTAs per SM so the tail imbalance stays below ~10% (256 model blocks on 148 SMs ran
    // at 1.73 CTAs per SM = 86% balance on the 8-GPU shard of cfg4)
    const int min_chunk = mpt == 1 ? SC_TILE : 4 * SC_TILE;
    int chunks = 1;
    if (mblocks < 8 * ctx->num_sms && n > min_chunk)
        chunks = min(pm_cdiv(8 * ctx->num_sms, mblocks), pm_cdiv(n, min_chunk));
    int chunk_pts = pm_round_up(pm_cdiv(n > 0 ? n : 1, chunks), SC_TILE);
    chunks = pm_cdiv(n > 0 ? n : 1, chunk_pts);
    const int use_atomic = chunks > 1;
    if (use_atomic) PM_CUDA(ctx, cudaMemsetAsync(dcounts, 0, (size_t)n_models * 4, ctx->stream));
    const float thr2 = thr * thr;
    dim3 grid(mblocks, chunks);
    {
        pm_prof_scope prof(ctx, 2);
        if (metric == PM_METRIC_SAMPSON) {
            if (mpt == 1) ransac_score_kernel<PM_METRIC_SAMPSON, 1><<<grid, SC_THREADS, 0, ctx->stream>>>(pts, n, chunk_pts, dF32, n_models, thr2, dcounts, use_atomic, dn, PS0);
            else ransac_score_kernel<PM_METRIC_SAMPSON, SC_MPT><<<grid, SC_THREADS, 0, ctx->stream>>>(pts, n, chunk_pts, dF32, n_models, thr2, dcounts, use_atomic, dn, PS0);
        } else {
            if (mpt == 1) ransac_score_kernel<PM_METRIC_SYMEPI, 1><<<grid, SC_THREADS, 0, ctx->stream>>>(pts, n, chunk_pts, dF32, n_models, thr2, dcounts, use_atomic, dn, PS0);
            else ransac_score_kernel<PM_METRIC_SYMEPI, SC_MPT><<<grid, SC_THREADS, 0, ctx->stream>>>(pts, n, chunk_pts, dF32, n_models, thr2, dcounts, use_atomic, dn, PS0);
        }
    }
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

int pmk_ransac_best(pm_ctx *ctx, const int32_t *dcounts, int n_models, int id_base, uint64_t *dkey)
{
    PM_CUDA(ctx, cudaMemsetAsync(dkey, 0, 8, ctx->stream));
    if (n_models <= 0) return PM_OK;
    const int blocks = min(pm_cdiv(n_models, 256), 4 * ctx->num_sms);
    ransac_best_kernel<<<blocks, 256, 0, ctx->stream>>>(dcounts, n_models, id_base, (unsigned long long *)dkey);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

// winner key + winning model + zeroed inlier counter in one single-block launch (small model sets)
int pmk_ransac_best_pick(pm_ctx *ctx, const int32_t *dcounts, int n_models, int id_base, const float *dF32, uint64_t *dkey,
                         float *dFw, int32_t *dn_inl)
{
    ransac_best_pick_kernel<<<1, 1024, 0, ctx->stream>>>(dcounts, n_models, id_base, dF32, (unsigned long long *)dkey, dFw, dn_inl, PS0);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

int pmk_ransac_pick(pm_ctx *ctx, const uint64_t *dkey, const float *dF32, int id_base, int n_models, float *dFw)
{
    ransac_pick_kernel<<<1, 32, 0, ctx->stream>>>((const unsigned long long *)dkey, dF32, id_base, n_models, dFw);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

int pmk_ransac_update_best(pm_ctx *ctx, const uint64_t *dkey, const float *dF32, int id_base, int n_models, uint64_t *dbest_key,
                           float *dFw_best)
{
    ransac_update_best_kernel<<<1, 32, 0, ctx->stream>>>((const unsigned long long *)dkey, dF32, id_base, n_models,
                                                         (unsigned long long *)dbest_key, dFw_best);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

int pmk_ransac_finish(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const float *dFw, float thr, int metric,
                      int refit, double *dF, uint8_t *dmask, int32_t *dn_inl, const int32_t *dn, const float *dpts4,
                      int ninl_is_zero, const uint64_t *dkey, int sample_size, pm_pair_result *dres)
{
    const float4 *pts = reinterpret_cast<const float4 *>(dpts4);
    if (!pts) {
        int st = get_pts4(ctx, dp1, dp2, n, &pts, dn);
        if (st != PM_OK) return st;
    }
    if (!ninl_is_zero) PM_CUDA(ctx, cudaMemsetAsync(dn_inl, 0, 4, ctx->stream));
    const float thr2 = thr * thr;
    if (n > 0) {
        if (metric == PM_METRIC_SAMPSON)
            ransac_mask_kernel<PM_METRIC_SAMPSON><<<pm_cdiv(n, 256), 256, 0, ctx->stream>>>(pts, n, dFw, thr2, dmask, dn_inl, dn, PS0);
        else
            ransac_mask_kernel<PM_METRIC_SYMEPI><<<pm_cdiv(n, 256), 256, 0, ctx->stream>>>(pts, n, dFw, thr2, dmask, dn_inl, dn, PS0);
        PM_CHECK_LAUNCH(ctx);
    }
    // dres: the pair pipeline's record; with a refit its last kernel writes it, else a kernel of its own
    if (refit) return run_refit(ctx, pts, n, dmask, dFw, dF, nullptr, dn,
                                PairOut{(const unsigned long long *)dkey, dn, dn_inl, n, sample_size, dres});
    copy_f32_to_f64_kernel<<<1, 32, 0, ctx->stream>>>(dFw, dF, 9);
    PM_CHECK_LAUNCH(ctx);
    if (dres) return pmk_pair_result(ctx, dkey, dn, dn_inl, dF, n, sample_size, dres);
    return PM_OK;
}

int pmk_fundamental_npoint(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const uint8_t *dmask, double *dF,
                           int32_t *dok)
{
    const float4 *pts;
    int st = get_pts4(ctx, dp1, dp2, n, &pts);
    if (st != PM_OK) return st;
    return run_refit(ctx, pts, n, dmask, nullptr, dF, dok);
}

// medians of the symmetric-epipolar error, one per model (NaN models: +inf)
int pmk_lmeds_score(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const float *dF32, int n_models, float *dmedians)
{
    if (n_models <= 0 || n <= 0) return PM_OK;
    const float4 *pts;
    int st = get_pts4(ctx, dp1, dp2, n, &pts);
    if (st != PM_OK) return st;
    // error scratch: at most ~64 MB, models in batches
    int batch = (int)((16ll << 20) / n);
    if (batch < 1) batch = 1;
    if (batch > n_models) batch = n_models;
    PM_WS(ctx, errs, float *, WS_LINES, (size_t)batch * n * sizeof(float));
    for (int m0 = 0; m0 < n_models; m0 += batch) {
        lmeds_median_kernel<<<min(batch, n_models - m0), LM_THREADS, 0, ctx->stream>>>(pts, n, dF32, m0, n_models, errs, dmedians);
        PM_CHECK_LAUNCH(ctx);
    }
    return PM_OK;
}

// winner of the medians -> F (f64[9], no refit: OpenCV returns the minimal model), mask, inlier count;
// *dkey = median bits << 32 | model id, ~0 when no model
int pmk_lmeds_finish(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const float *dF32, const float *dmedians,
                     int n_models, float *dFw, double *dF, uint8_t *dmask, int32_t *dn_inl, uint64_t *dkey)
{
    const float4 *pts;
    int st = get_pts4(ctx, dp1, dp2, n, &pts);
    if (st != PM_OK) return st;
    PM_CUDA(ctx, cudaMemsetAsync(dkey, 0xFF, 8, ctx->stream));
    PM_CUDA(ctx, cudaMemsetAsync(dn_inl, 0, 4, ctx->stream));
    if (n_models > 0) {
        lmeds_best_kernel<<<min(pm_cdiv(n_models, 256), 4 * ctx->num_sms), 256, 0, ctx->stream>>>(dmedians, n_models, 0, (unsigned long long *)dkey);
        PM_CHECK_LAUNCH(ctx);
    }
    lmeds_pick_kernel<<<1, 32, 0, ctx->stream>>>((const unsigned long long *)dkey, dF32, 0, n_models, n, dFw, dF);
    PM_CHECK_LAUNCH(ctx);
    if (n > 0) {
        lmeds_mask_kernel<<<pm_cdiv(n, 256), 256, 0, ctx->stream>>>(pts, n, dFw, dmask, dn_inl);
        PM_CHECK_LAUNCH(ctx);
    }
    return PM_OK;
}

// Device twin of pm_make_sample_sets (pm_api.cu): same splitmix64 stream, same rejection of repeats, so the
// sets are identical to the host generator's -- one thread per hypothesis.
// h_base: row h of `out` is hypothesis h_base + h of the stream (batches / shards of one job draw disjoint ranges).
__global__ void sample_sets_kernel(int n_points, int n_hyp, int m, unsigned long long seed, int32_t *__restrict__ out,
                                   const int32_t *n_dev, int h_base, PairStride ps)
{
    { const long long pb = blockIdx.y; out += pb * ps.s[0]; if (n_dev) n_dev += pb * ps.s[1]; seed += (unsigned long long)(pb * ps.s[2]); }
    n_points = eff_n(n_points, n_dev);
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_hyp) return;
    if (n_points < m) {                  // no minimal sample exists (the solver writes "no model" for this case)
        for (int i = 0; i < m; ++i) out[(size_t)h * m + i] = 0;
        return;
    }
    unsigned long long s = seed ^ (0xD1B54A32D192ED03ull * ((unsigned long long)h_base + (unsigned long long)h + 1ull));
    int32_t row[8];
    for (int i = 0; i < m; ++i) {
        for (;;) {
            unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            const int32_t v = (int32_t)(z % (unsigned long long)n_points);
            bool dup = false;
            for (int k = 0; k < i; ++k) dup = dup || row[k] == v;
            if (!dup) { row[i] = v; break; }
        }
    }
    for (int i = 0; i < m; ++i) out[(size_t)h * m + i] = row[i];
}

int pmk_sample_sets(pm_ctx *ctx, int n_points, int n_hyp, int m, uint64_t seed, int32_t *dout, const int32_t *dn, int h_base)
{
    if (n_hyp <= 0) return PM_OK;
    sample_sets_kernel<<<pm_cdiv(n_hyp, 128), 128, 0, ctx->stream>>>(n_points, n_hyp, m, (unsigned long long)seed, dout, dn, h_base, PS0);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

// Sharded RANSAC (pm_find_fundamental_sharded_dev): after the max-reduce every rank holds the global winner key and
// re-solves the winning minimal sample itself.  Step 1: the winner's index set -> out[0..m) -- row id / per of the full
// sample array, or regenerated from the seed (same splitmix64 stream as sample_sets_kernel).  key == 0: zeros.
__global__ void winner_samples_kernel(const unsigned long long *key, const int32_t *__restrict__ samples_full, int n_points,
                                      int m, int per, unsigned long long seed, int32_t *__restrict__ out)
{
    if (threadIdx.x != 0) return;
    const unsigned long long k = *key;
    if (k == 0ull || n_points < m) { for (int i = 0; i < m; ++i) out[i] = 0; return; }
    const unsigned id = 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull);
    const unsigned h = id / (unsigned)per;
    if (samples_full) { for (int i = 0; i < m; ++i) out[i] = samples_full[(size_t)h * m + i]; return; }
    unsigned long long s = seed ^ (0xD1B54A32D192ED03ull * ((unsigned long long)h + 1ull));
    for (int i = 0; i < m; ++i) {
        for (;;) {
            unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            const int32_t v = (int32_t)(z % (unsigned long long)n_points);
            bool dup = false;
            for (int kk = 0; kk < i; ++kk) dup = dup || out[kk] == v;
            if (!dup) { out[i] = v; break; }
        }
    }
}
// Step 2 (after the solver ran on that one sample): model id % per of its <= 3 models -> Fw
__global__ void winner_pick_kernel(const unsigned long long *key, const float *__restrict__ Fm, int per, float *Fw)
{
    const int i = threadIdx.x;
    if (i >= 12) return;
    const unsigned long long k = *key;
    const unsigned id = 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull);
    Fw[i] = k != 0ull ? Fm[(size_t)(id % (unsigned)per) * 12 + i] : __int_as_float(0x7fc00000);
}

int pmk_ransac_winner_resolve(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const uint64_t *dkey,
                              const int32_t *dsamples_full, uint64_t seed, int m, int32_t *dwin_idx, float *dF3, float *dFw)
{
    const int per = m == 8 ? 1 : 3;
    winner_samples_kernel<<<1, 32, 0, ctx->stream>>>((const unsigned long long *)dkey, dsamples_full, n, m, per,
                                                     (unsigned long long)seed, dwin_idx);
    PM_CHECK_LAUNCH(ctx);
    int st = pmk_ransac_solve(ctx, dp1, dp2, n, dwin_idx, 1, m, dF3);
    if (st != PM_OK) return st;
    winner_pick_kernel<<<1, 32, 0, ctx->stream>>>((const unsigned long long *)dkey, dF3, per, dFw);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

// The RANSAC half of the pair pipeline for a GROUP of image pairs in ONE launch per kernel (the pair is a grid
// dimension): sample sets, minimal solves, scoring, winner, mask, the three refit passes and the refit solve -- nine
// launches for the whole group instead of nine per pair, and grids large enough to fill the GPU (one pair's 4096
// hypotheses are 0.43 waves of the solver and 0.27 of the scorer).  Slot b of every array belongs to pair b of the group;
// the arithmetic per pair is exactly that of the single-pair calls.
int pmk_pair_group_ransac(pm_ctx *ctx, const pm_pair_group &G)
{
    const int P = G.n_pairs, m = G.m, per = m == 8 ? 1 : 3, nh = G.n_hyp, nm = nh * per, nmax = G.nmax;
    if (P <= 0) return PM_OK;
    const long long sF = (long long)(nm + 1) * 12, sKeyI = 16 /* ints per 64-byte key record */;
    const int32_t *n_good = reinterpret_cast<const int32_t *>(G.key + 2);     // [b]: + b * 16 ints
    int32_t *n_inl = reinterpret_cast<int32_t *>(G.key + 2) + 1;
    float *Fw = G.F32 + (size_t)nm * 12;                                       // slot b: + b * sF
    // sample sets: pair b draws from seed0 + b, bounded by its own match count
    sample_sets_kernel<<<dim3(pm_cdiv(nh, 128), P), 128, 0, ctx->stream>>>(nmax, nh, m, (unsigned long long)G.seed0, G.samples, n_good, 0,
                                                                           PairStride{{(long long)nh * m, sKeyI, 1, 0, 0, 0, 0, 0}});
    PM_CHECK_LAUNCH(ctx);
    {
        const dim3 grid(pm_cdiv(nh * 8, SOLVE_THREADS), P);
        const PairStride ps = {{nmax, nmax, (long long)nh * m, sF, sKeyI, 0, 0, 0}};
        if (m == 8) {
            PM_WS(ctx, hand, double *, WS_HAND, (size_t)P * nh * 16 * sizeof(double));
            ransac_solve_kernel<8><<<grid, SOLVE_THREADS, 0, ctx->stream>>>(G.p1, G.p2, nmax, G.samples, nh, G.F32, n_good, nullptr, hand, ps);
            PM_CHECK_LAUNCH(ctx);
            ransac_project_kernel<<<dim3(pm_cdiv(nh, 128), P), 128, 0, ctx->stream>>>(hand, nh, G.F32, nullptr, sF);
        } else {
            ransac_solve_kernel<7><<<grid, SOLVE_THREADS, 0, ctx->stream>>>(G.p1, G.p2, nmax, G.samples, nh, G.F32, n_good, nullptr, nullptr, ps);
        }
        PM_CHECK_LAUNCH(ctx);
    }
    {
        // scoring: 4 models per thread when that still leaves >= 8 CTAs per SM with 512-point chunks, else 1 model per
        // thread (a group of 8 pairs x 4096 hypotheses x 4096 matches: 2048 CTAs instead of 128)
        int mpt = SC_MPT;
        if ((long long)pm_cdiv(nm, SC_THREADS * SC_MPT) * pm_cdiv(nmax, SC_TILE) * P < 8 * ctx->num_sms) mpt = 1;
        const int mblocks = pm_cdiv(nm, SC_THREADS * mpt);
        int chunks = 1;
        if ((long long)mblocks * P < 8 * ctx->num_sms && nmax > SC_TILE)
            chunks = min(pm_cdiv(8 * ctx->num_sms, mblocks * P), pm_cdiv(nmax, SC_TILE));
        int chunk_pts = pm_round_up(pm_cdiv(nmax, chunks), SC_TILE);
        chunks = pm_cdiv(nmax, chunk_pts);
        const int use_atomic = chunks > 1;
        if (use_atomic) PM_CUDA(ctx, cudaMemsetAsync(G.counts, 0, (size_t)P * nm * 4, ctx->stream));
        const float thr2 = G.threshold * G.threshold;
        const dim3 grid(mblocks, chunks, P);
        const PairStride ps = {{nmax, sF, nm, sKeyI, 0, 0, 0, 0}};
        pm_prof_scope prof(ctx, 2);
        if (G.metric == PM_METRIC_SAMPSON) {
            if (mpt == 1) ransac_score_kernel<PM_METRIC_SAMPSON, 1><<<grid, SC_THREADS, 0, ctx->stream>>>(G.pts4, nmax, chunk_pts, G.F32, nm, thr2, G.counts, use_atomic, n_good, ps);
            else ransac_score_kernel<PM_METRIC_SAMPSON, SC_MPT><<<grid, SC_THREADS, 0, ctx->stream>>>(G.pts4, nmax, chunk_pts, G.F32, nm, thr2, G.counts, use_atomic, n_good, ps);
        } else {
            if (mpt == 1) ransac_score_kernel<PM_METRIC_SYMEPI, 1><<<grid, SC_THREADS, 0, ctx->stream>>>(G.pts4, nmax, chunk_pts, G.F32, nm, thr2, G.counts, use_atomic, n_good, ps);
            else ransac_score_kernel<PM_METRIC_SYMEPI, SC_MPT><<<grid, SC_THREADS, 0, ctx->stream>>>(G.pts4, nmax, chunk_pts, G.F32, nm, thr2, G.counts, use_atomic, n_good, ps);
        }
    }
    PM_CHECK_LAUNCH(ctx);
    ransac_best_pick_kernel<<<P, 1024, 0, ctx->stream>>>(G.counts, nm, 0, G.F32, (unsigned long long *)G.key, Fw, n_inl,
                                                         PairStride{{nm, sF, 8, sF, sKeyI, 0, 0, 0}});
    PM_CHECK_LAUNCH(ctx);
    const float thr2 = G.threshold * G.threshold;
    {
        const dim3 grid(pm_cdiv(nmax, 256), P);
        const PairStride ps = {{nmax, sF, nmax, sKeyI, sKeyI, 0, 0, 0}};
        if (G.metric == PM_METRIC_SAMPSON) ransac_mask_kernel<PM_METRIC_SAMPSON><<<grid, 256, 0, ctx->stream>>>(G.pts4, nmax, Fw, thr2, G.mask, n_inl, n_good, ps);
        else ransac_mask_kernel<PM_METRIC_SYMEPI><<<grid, 256, 0, ctx->stream>>>(G.pts4, nmax, Fw, thr2, G.mask, n_inl, n_good, ps);
        PM_CHECK_LAUNCH(ctx);
    }
    const long long sR = RF_BLOCKS * (5 + 2 + 45) + 16;
    double *part1 = G.refit, *part2 = part1 + RF_BLOCKS * 5, *part3 = part2 + RF_BLOCKS * 2, *stats = part3 + RF_BLOCKS * 45;
    const PairOut po = {(const unsigned long long *)G.key, n_good, n_inl, nmax, m, G.res};
    if (G.refit_on) {
        refit_sum_kernel<<<dim3(RF_BLOCKS, P), RF_THREADS, 0, ctx->stream>>>(G.pts4, nmax, G.mask, part1, n_good, PairStride{{nmax, nmax, sR, sKeyI, 0, 0, 0, 0}});
        PM_CHECK_LAUNCH(ctx);
        refit_scale_kernel<<<dim3(RF_BLOCKS, P), RF_THREADS, 0, ctx->stream>>>(G.pts4, nmax, G.mask, stats, part2, n_good, part1,
                                                                              PairStride{{nmax, nmax, sR, sR, sKeyI, sR, 0, 0}});
        PM_CHECK_LAUNCH(ctx);
        refit_ata_kernel<<<dim3(RF_BLOCKS, P), RF_THREADS, 0, ctx->stream>>>(G.pts4, nmax, G.mask, stats, part3, n_good, part2,
                                                                            PairStride{{nmax, nmax, sR, sR, sKeyI, sR, 0, 0}});
        PM_CHECK_LAUNCH(ctx);
        refit_solve_kernel<<<P, 32, 0, ctx->stream>>>(part3, stats, Fw, G.Fout, nullptr, po, PairStride{{sR, sR, sF, 16, 8, sKeyI, sKeyI, 1}});
        PM_CHECK_LAUNCH(ctx);
    } else {
        pair_group_result_kernel<<<P, 32, 0, ctx->stream>>>(Fw, sF, po, PairStride{{0, 0, 0, 0, 8, sKeyI, sKeyI, 1}});
        PM_CHECK_LAUNCH(ctx);
    }
    return PM_OK;
}

// One record per image pair of the asynchronous pair pipeline (pm_match_estimate_*_dev).
__global__ void pair_result_kernel(const unsigned long long *key, const int32_t *n_good, const int32_t *n_inl,
                                   const double *F, int n_max, int m, pm_pair_result *res)
{
    const int i = threadIdx.x;
    const unsigned long long k = *key;
    const int n = min(max(*n_good, 0), n_max);
    const bool ok = k != 0ull && n >= m;
    if (i < 9) res->F[i] = ok ? F[i] : 0.0;
    if (i == 9) { res->key = ok ? k : 0ull; res->n_matches = n; res->n_inliers = ok ? *n_inl : 0; res->has_model = ok; res->reserved = 0; }
}

int pmk_pair_result(pm_ctx *ctx, const uint64_t *dkey, const int32_t *dn_good, const int32_t *dn_inl, const double *dF,
                    int n_max, int m, pm_pair_result *dres)
{
    pair_result_kernel<<<1, 32, 0, ctx->stream>>>((const unsigned long long *)dkey, dn_good, dn_inl, dF, n_max, m, dres);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

int pmk_epilines(pm_ctx *ctx, const float *dpts, int n, int which, const double *dF, float *dlines)
{
    if (n <= 0) return PM_OK;
    epilines_kernel<<<pm_cdiv(n, 256), 256, 0, ctx->stream>>>((const float2 *)dpts, n, which, dF, dlines);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

int pmk_residuals(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const double *dF, int metric, float *dout,
                  double *dsum)
{
    if (dsum) PM_CUDA(ctx, cudaMemsetAsync(dsum, 0, 8, ctx->stream));
    if (n <= 0) return PM_OK;
    const int nb = pm_cdiv(n, 256);
    double *partial = nullptr;
    if (dsum) { PM_WS(ctx, p, double *, WS_REFIT, (size_t)max(nb, RF_BLOCKS * (5 + 2 + 45) + 16) * sizeof(double)); partial = p; }
    residuals_kernel<<<nb, 256, 0, ctx->stream>>>((const float2 *)dp1, (const float2 *)dp2, n, dF, metric, dout, partial);
    PM_CHECK_LAUNCH(ctx);
    if (dsum) {
        residuals_sum_kernel<<<1, 256, 0, ctx->stream>>>(partial, nb, dsum);
        PM_CHECK_LAUNCH(ctx);
    }
    return PM_OK;
}
