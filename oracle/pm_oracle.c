/*
 * pm_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see pm_oracle.h).
 *
 * Plain C99 restatement of the OpenCV behaviour reached from
 * /root/reference/Points Matching/main.cpp:43-46, 49-69, 89-91, 95-98, 127-132.
 * Build: see oracle/Makefile (must use -ffp-contract=off: the *_f32 functions
 * define the product's bit-exact FP32 operation order with explicit fmaf()).
 */
#include "pm_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.1415926535897932384626433832795

/* ------------------------------------------------------------------------- */
/* matching: main.cpp:43-46 (BruteForceMatcher<L2<float>>, match/knnMatch)    */
/* ------------------------------------------------------------------------- */

static void put2(orc_dmatch *o, int qi, double d0, int i0, double d1, int i1, int sqrt_it)
{
    o[0].queryIdx = qi; o[0].trainIdx = i0; o[0].imgIdx = 0;
    o[0].distance = i0 < 0 ? FLT_MAX : (float)(sqrt_it ? sqrt(d0) : d0);
    o[1].queryIdx = qi; o[1].trainIdx = i1; o[1].imgIdx = 0;
    o[1].distance = i1 < 0 ? FLT_MAX : (float)(sqrt_it ? sqrt(d1) : d1);
}

void orc_knn2_l2_f32(const float *q, int nq, const float *t, int nt, int dim,
                     orc_dmatch *out, int nthreads)
{
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int i = 0; i < nq; ++i) {
        const float *a = q + (size_t)i * dim;
        double b0 = INFINITY, b1 = INFINITY;
        int i0 = -1, i1 = -1;
        for (int j = 0; j < nt; ++j) {
            const float *b = t + (size_t)j * dim;
            double s = 0.0;
            for (int k = 0; k < dim; ++k) {
                double d = (double)a[k] - (double)b[k];
                s += d * d;
            }
            /* strict < keeps the lowest trainIdx on ties (OpenCV batchDistance) */
            if (s < b0) { b1 = b0; i1 = i0; b0 = s; i0 = j; }
            else if (s < b1) { b1 = s; i1 = j; }
        }
        put2(out + (size_t)i * 2, i, b0, i0, b1, i1, 1);
    }
}

float orc_l2sq_f32_rerank(const float *a, const float *b, int dim)
{
    /* The product's "re-rank order" (DESIGN.md): 32 lanes, lane l owns k = 128c + 4l + e
     * (c ascending, e = 0..3) as one fmaf chain, then an xor-butterfly sum 16,8,4,2,1. */
    float p[32], q[32];
    for (int l = 0; l < 32; ++l) p[l] = 0.f;
    for (int c = 0; c < dim; c += 128)
        for (int l = 0; l < 32; ++l)
            for (int e = 0; e < 4; ++e) {
                int k = c + 4 * l + e;
                if (k < dim) { float d = a[k] - b[k]; p[l] = fmaf(d, d, p[l]); }
            }
    for (int off = 16; off > 0; off >>= 1) {
        for (int l = 0; l < 32; ++l) q[l] = p[l] + p[l ^ off];
        memcpy(p, q, sizeof(p));
    }
    return p[0];
}

static inline int popc8(const uint8_t *a, const uint8_t *b, int bytes)
{
    int s = 0, k = 0;
    for (; k + 8 <= bytes; k += 8) {
        uint64_t x, y;
        memcpy(&x, a + k, 8); memcpy(&y, b + k, 8);
        s += __builtin_popcountll(x ^ y);
    }
    for (; k < bytes; ++k) s += __builtin_popcount((unsigned)(a[k] ^ b[k]));
    return s;
}

void orc_knn2_hamming(const uint8_t *q, int nq, const uint8_t *t, int nt, int bytes,
                      orc_dmatch *out, int nthreads)
{
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int i = 0; i < nq; ++i) {
        const uint8_t *a = q + (size_t)i * bytes;
        int b0 = INT32_MAX, b1 = INT32_MAX, i0 = -1, i1 = -1;
        for (int j = 0; j < nt; ++j) {
            int s = popc8(a, t + (size_t)j * bytes, bytes);
            if (s < b0) { b1 = b0; i1 = i0; b0 = s; i0 = j; }
            else if (s < b1) { b1 = s; i1 = j; }
        }
        put2(out + (size_t)i * 2, i, (double)b0, i0, (double)b1, i1, 0);
    }
}

void orc_col_best_hamming(const uint8_t *q, int nq, const uint8_t *t, int nt, int bytes,
                          uint64_t *col_best, int nthreads)
{
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int j = 0; j < nt; ++j) {
        const uint8_t *b = t + (size_t)j * bytes;
        uint64_t best = UINT64_MAX;
        for (int i = 0; i < nq; ++i) {
            float d = (float)popc8(q + (size_t)i * bytes, b, bytes);
            uint32_t bits; memcpy(&bits, &d, 4);
            uint64_t key = ((uint64_t)bits << 32) | (uint32_t)i;
            if (key < best) best = key;
        }
        col_best[j] = best;
    }
}

/* ------------------------------------------------------------------------- */
/* filters: main.cpp:49-69                                                    */
/* ------------------------------------------------------------------------- */

int orc_ratio_filter(const orc_dmatch *knn, int nq, float ratio, orc_dmatch *out)
{
    int n = 0;
    for (int i = 0; i < nq; ++i) {
        const orc_dmatch *m = knn + (size_t)i * 2;
        if (m[0].trainIdx < 0 || m[1].trainIdx < 0) continue;
        if (m[0].distance < ratio * m[1].distance) out[n++] = m[0];
    }
    return n;
}

int orc_cross_check(const orc_dmatch *knn, int nq, int knn_stride,
                    const uint64_t *col_best, int nt, orc_dmatch *out)
{
    int n = 0;
    for (int i = 0; i < nq; ++i) {
        const orc_dmatch *m = knn + (size_t)i * knn_stride;
        int j = m->trainIdx;
        if (j < 0 || j >= nt) continue;
        if ((uint32_t)(col_best[j] & 0xFFFFFFFFu) == (uint32_t)m->queryIdx) out[n++] = *m;   /* queryIdx == i unless the rows are a shard */
    }
    return n;
}

int orc_minmax_filter(const orc_dmatch *m, int n, orc_dmatch *out,
                      double *min_out, double *max_out)
{
    double mn = 1, mx = 0;                 /* main.cpp:49-50 */
    for (int i = 0; i < n; ++i) {          /* main.cpp:51-56 */
        mn = mn > m[i].distance ? m[i].distance : mn;
        mx = mx < m[i].distance ? m[i].distance : mx;
    }
    int k = 0;
    for (int i = 0; i < n; ++i)            /* main.cpp:63-69 */
        if (m[i].distance < mn + (mx - mn) / 2) out[k++] = m[i];
    if (min_out) *min_out = mn;
    if (max_out) *max_out = mx;
    return k;
}

void orc_gather_points(const float *kp_xy, const int32_t *idx, int n, float *out)
{
    for (int i = 0; i < n; ++i) {
        out[2 * i] = kp_xy[2 * (size_t)idx[i]];
        out[2 * i + 1] = kp_xy[2 * (size_t)idx[i] + 1];
    }
}

/* ------------------------------------------------------------------------- */
/* small dense linear algebra (f64)                                           */
/* ------------------------------------------------------------------------- */

/* Cyclic Jacobi for a symmetric n x n matrix (n <= 9).  On return w[] holds the
 * eigenvalues in DESCENDING order and V[k*n..] the k-th eigenvector (cv::eigen
 * convention used by run8Point). */
static void jacobi_eig_sym(const double *A_in, int n, double *w, double *V)
{
    double A[81];
    memcpy(A, A_in, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * n + j] = (i == j);
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, diag = 0;
        for (int i = 0; i < n; ++i) {
            diag += A[i * n + i] * A[i * n + i];
            for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
        }
        if (off <= 1e-34 * diag || off == 0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double apq = A[p * n + q];
                if (apq == 0) continue;
                double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
                double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                double c = 1 / sqrt(tt * tt + 1), s = tt * c;
                for (int k = 0; k < n; ++k) {
                    double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {   /* rows of V are eigenvectors */
                    double vpk = V[p * n + k], vqk = V[q * n + k];
                    V[p * n + k] = c * vpk - s * vqk;
                    V[q * n + k] = s * vpk + c * vqk;
                }
            }
    }
    for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
    for (int i = 0; i < n - 1; ++i) {           /* selection sort, descending */
        int m = i;
        for (int j = i + 1; j < n; ++j) if (w[j] > w[m]) m = j;
        if (m != i) {
            double tw = w[i]; w[i] = w[m]; w[m] = tw;
            for (int k = 0; k < n; ++k) {
                double tv = V[i * n + k]; V[i * n + k] = V[m * n + k]; V[m * n + k] = tv;
            }
        }
    }
}

/* One-sided (Hestenes) Jacobi SVD of A (m x n, row-major), any m,n <= 9:
 * A*V = U*S.  On return the columns of Acols (m x n) are U_k*s_k, V is n x n
 * (columns = right singular vectors), s[] the column norms, sorted descending. */
static void jacobi_svd_cols(double *A, int m, int n, double *V, double *s)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * n + j] = (i == j);
    for (int sweep = 0; sweep < 80; ++sweep) {
        int changed = 0;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double a = 0, b = 0, g = 0;
                for (int k = 0; k < m; ++k) {
                    double x = A[k * n + p], y = A[k * n + q];
                    a += x * x; b += y * y; g += x * y;
                }
                if (g == 0 || fabs(g) <= 1e-17 * sqrt(a * b)) continue;
                changed = 1;
                double zeta = (b - a) / (2 * g);
                double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
                double c = 1 / sqrt(1 + tt * tt), sn = c * tt;
                for (int k = 0; k < m; ++k) {
                    double x = A[k * n + p], y = A[k * n + q];
                    A[k * n + p] = c * x - sn * y;
                    A[k * n + q] = sn * x + c * y;
                }
                for (int k = 0; k < n; ++k) {
                    double x = V[k * n + p], y = V[k * n + q];
                    V[k * n + p] = c * x - sn * y;
                    V[k * n + q] = sn * x + c * y;
                }
            }
        if (!changed) break;
    }
    for (int j = 0; j < n; ++j) {
        double a = 0;
        for (int k = 0; k < m; ++k) a += A[k * n + j] * A[k * n + j];
        s[j] = sqrt(a);
    }
    for (int i = 0; i < n - 1; ++i) {
        int mx = i;
        for (int j = i + 1; j < n; ++j) if (s[j] > s[mx]) mx = j;
        if (mx != i) {
            double ts = s[i]; s[i] = s[mx]; s[mx] = ts;
            for (int k = 0; k < m; ++k) {
                double tv = A[k * n + i]; A[k * n + i] = A[k * n + mx]; A[k * n + mx] = tv;
            }
            for (int k = 0; k < n; ++k) {
                double tv = V[k * n + i]; V[k * n + i] = V[k * n + mx]; V[k * n + mx] = tv;
            }
        }
    }
}

/* Zero the smallest singular value of a 3x3 (run8Point "make F0 singular"). */
static void rank2_project(double F[9])
{
    double A[9], V[9], s[3];
    memcpy(A, F, sizeof(A));
    jacobi_svd_cols(A, 3, 3, V, s);
    /* F = sum_k (A[:,k]) V[:,k]^T over k = 0,1 */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            F[i * 3 + j] = A[i * 3 + 0] * V[j * 3 + 0] + A[i * 3 + 1] * V[j * 3 + 1];
}

/* ------------------------------------------------------------------------- */
/* solvers: inside cv::findFundamentalMat, main.cpp:95-98                     */
/* ------------------------------------------------------------------------- */

static int fm_8point_core(const float *p1, const float *p2, const int32_t *idx, int n, double F[9])
{
    /* OpenCV run8Point: isotropic normalisation, 9x9 normal matrix, smallest
     * eigenvector, rank-2 projection, de-normalise, F[8] -> 1. */
    if (n < 8) return 0;
    double c1x = 0, c1y = 0, c2x = 0, c2y = 0;
    for (int i = 0; i < n; ++i) {
        int k = idx ? idx[i] : i;
        c1x += p1[2 * k]; c1y += p1[2 * k + 1];
        c2x += p2[2 * k]; c2y += p2[2 * k + 1];
    }
    double t = 1.0 / n;
    c1x *= t; c1y *= t; c2x *= t; c2y *= t;
    double s1 = 0, s2 = 0;
    for (int i = 0; i < n; ++i) {
        int k = idx ? idx[i] : i;
        double dx = p1[2 * k] - c1x, dy = p1[2 * k + 1] - c1y;
        s1 += sqrt(dx * dx + dy * dy);
        dx = p2[2 * k] - c2x; dy = p2[2 * k + 1] - c2y;
        s2 += sqrt(dx * dx + dy * dy);
    }
    s1 *= t; s2 *= t;
    if (s1 < FLT_EPSILON || s2 < FLT_EPSILON) return 0;
    s1 = sqrt(2.) / s1; s2 = sqrt(2.) / s2;

    double A[81];
    memset(A, 0, sizeof(A));
    for (int i = 0; i < n; ++i) {
        int k = idx ? idx[i] : i;
        double x1 = (p1[2 * k] - c1x) * s1, y1 = (p1[2 * k + 1] - c1y) * s1;
        double x2 = (p2[2 * k] - c2x) * s2, y2 = (p2[2 * k + 1] - c2y) * s2;
        double r[9] = { x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1 };
        for (int a = 0; a < 9; ++a)
            for (int b = 0; b < 9; ++b) A[a * 9 + b] += r[a] * r[b];
    }
    double W[9], V[81];
    jacobi_eig_sym(A, 9, W, V);
    for (int i = 0; i < 8; ++i)
        if (fabs(W[i]) < DBL_EPSILON) return 0;
    double F0[9];
    memcpy(F0, V + 8 * 9, sizeof(F0));
    rank2_project(F0);
    /* F = T2^T F0 T1 with T = [s 0 -s*cx; 0 s -s*cy; 0 0 1] */
    double T1[9] = { s1, 0, -s1 * c1x, 0, s1, -s1 * c1y, 0, 0, 1 };
    double T2[9] = { s2, 0, -s2 * c2x, 0, s2, -s2 * c2y, 0, 0, 1 };
    double M[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double a = 0;
            for (int k = 0; k < 3; ++k) a += T2[k * 3 + i] * F0[k * 3 + j];
            M[i * 3 + j] = a;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double a = 0;
            for (int k = 0; k < 3; ++k) a += M[i * 3 + k] * T1[k * 3 + j];
            F[i * 3 + j] = a;
        }
    if (fabs(F[8]) > FLT_EPSILON) {
        double inv = 1.0 / F[8];
        for (int i = 0; i < 9; ++i) F[i] *= inv;
    }
    return 1;
}

int orc_fm_8point(const float *p1, const float *p2, int n, double F[9])
{ return fm_8point_core(p1, p2, NULL, n, F); }

int orc_fm_8point_idx(const float *p1, const float *p2, const int32_t *idx, int m, double F[9])
{ return fm_8point_core(p1, p2, idx, m, F); }

/* cv::solveCubic restatement (coefficients a0 x^3 + a1 x^2 + a2 x + a3). */
static int solve_cubic(const double c[4], double r[3])
{
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
                double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
                if (fabs(q1) > fabs(q2)) { x0 = q1 / a1; x1 = a3 / q1; }
                else { x0 = q2 / a1; x1 = a3 / q2; }
                n = d > 0 ? 2 : 1;
            }
        }
    } else {
        a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
        double Q = (a1 * a1 - 3 * a2) * (1. / 9);
        double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
        double Qcubed = Q * Q * Q;
        double d = Qcubed - R * R;
        if (d > 0) {
            double theta = acos(R / sqrt(Qcubed));
            double sqrtQ = sqrt(Q);
            double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
            x0 = t0 * cos(t1) - t2;
            x1 = t0 * cos(t1 + (2. * ORC_PI / 3)) - t2;
            x2 = t0 * cos(t1 + (4. * ORC_PI / 3)) - t2;
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

static int fm_7point_core(const float *p1, const float *p2, const int32_t *idx, double F[27])
{
    /* OpenCV run7Point: rows on UN-normalised coordinates, 2-d null space f1,f2,
     * real roots of det(l*f1 + (1-l)*f2) = 0, each scaled so F[8] = 1. */
    double A[63], V[81], s[9];
    for (int i = 0; i < 7; ++i) {
        int k = idx ? idx[i] : i;
        double x1 = p1[2 * k], y1 = p1[2 * k + 1], x2 = p2[2 * k], y2 = p2[2 * k + 1];
        double *r = A + i * 9;
        r[0] = x2 * x1; r[1] = x2 * y1; r[2] = x2;
        r[3] = y2 * x1; r[4] = y2 * y1; r[5] = y2;
        r[6] = x1; r[7] = y1; r[8] = 1;
    }
    jacobi_svd_cols(A, 7, 9, V, s);
    double f1[9], f2[9];
    for (int k = 0; k < 9; ++k) { f1[k] = V[k * 9 + 7]; f2[k] = V[k * 9 + 8]; }
    /* f1 <- f1 - f2 so that lambda*f1 + f2 = lambda*F1 + (1-lambda)*F2 */
    for (int k = 0; k < 9; ++k) f1[k] -= f2[k];
    double c[4], t0, t1, t2;
    t0 = f2[4] * f2[8] - f2[5] * f2[7];
    t1 = f2[3] * f2[8] - f2[5] * f2[6];
    t2 = f2[3] * f2[7] - f2[4] * f2[6];
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
    c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 -
           f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
           f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
           f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
           f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) -
           f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
           f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
    t0 = f1[4] * f1[8] - f1[5] * f1[7];
    t1 = f1[3] * f1[8] - f1[5] * f1[6];
    t2 = f1[3] * f1[7] - f1[4] * f1[6];
    c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 -
           f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
           f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
           f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
           f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) -
           f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
           f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
    double r[3];
    int n = solve_cubic(c, r);
    if (n < 1 || n > 3) return 0;
    for (int k = 0; k < n; ++k) {
        double lambda = r[k], mu = 1;
        double sc = f1[8] * r[k] + f2[8];
        double *Fk = F + 9 * k;
        if (fabs(sc) > DBL_EPSILON) { mu = 1. / sc; lambda *= mu; Fk[8] = 1; }
        else Fk[8] = 0;
        for (int i = 0; i < 8; ++i) Fk[i] = f1[i] * lambda + f2[i] * mu;
    }
    return n;
}

int orc_fm_7point(const float *p1, const float *p2, double F[27])
{ return fm_7point_core(p1, p2, NULL, F); }

int orc_fm_7point_idx(const float *p1, const float *p2, const int32_t *idx, double F[27])
{ return fm_7point_core(p1, p2, idx, F); }

/* ------------------------------------------------------------------------- */
/* residuals                                                                  */
/* ------------------------------------------------------------------------- */

/* The product's FP32 scoring arithmetic, operation for operation
 * (DESIGN.md "scoring op order").  Returns r (= x2^T F x1) and *den. */
float orc_sampson_num_den_f32(const float F[9], float x1, float y1, float x2, float y2, float *den)
{
    float a  = fmaf(F[0], x1, fmaf(F[1], y1, F[2]));
    float b  = fmaf(F[3], x1, fmaf(F[4], y1, F[5]));
    float c  = fmaf(F[6], x1, fmaf(F[7], y1, F[8]));
    float r  = fmaf(x2, a, fmaf(y2, b, c));
    float at = fmaf(F[0], x2, fmaf(F[3], y2, F[6]));
    float bt = fmaf(F[1], x2, fmaf(F[4], y2, F[7]));
    *den = fmaf(a, a, fmaf(b, b, fmaf(at, at, bt * bt)));
    return r;
}

int orc_is_inlier_f32(const float F[9], float x1, float y1, float x2, float y2,
                      float thr2, int metric)
{
    float a  = fmaf(F[0], x1, fmaf(F[1], y1, F[2]));
    float b  = fmaf(F[3], x1, fmaf(F[4], y1, F[5]));
    float c  = fmaf(F[6], x1, fmaf(F[7], y1, F[8]));
    float r  = fmaf(x2, a, fmaf(y2, b, c));
    float at = fmaf(F[0], x2, fmaf(F[3], y2, F[6]));
    float bt = fmaf(F[1], x2, fmaf(F[4], y2, F[7]));
    float nr2 = -(r * r);
    /* margin form (one fused multiply-add, sign = outlier): the product's "scoring op order" */
    if (metric == ORC_METRIC_SAMPSON) {
        float den = fmaf(a, a, fmaf(b, b, fmaf(at, at, bt * bt)));
        return fmaf(thr2, den, nr2) >= 0.0f;
    } else {
        /* max(d1^2 s1, d2^2 s2) <= thr^2, division-free; x1^T F^T x2 == x2^T F x1 */
        float n2 = fmaf(a, a, b * b);
        float n1 = fmaf(at, at, bt * bt);
        return (fmaf(thr2, n2, nr2) >= 0.0f) && (fmaf(thr2, n1, nr2) >= 0.0f);
    }
}

double orc_sampson_f64(const double F[9], double x1, double y1, double x2, double y2)
{
    double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5];
    double c = F[6] * x1 + F[7] * y1 + F[8];
    double r = x2 * a + y2 * b + c;
    double at = F[0] * x2 + F[3] * y2 + F[6], bt = F[1] * x2 + F[4] * y2 + F[7];
    return r * r / (a * a + b * b + at * at + bt * bt);
}

double orc_symepi_f64(const double F[9], double x1, double y1, double x2, double y2)
{
    /* OpenCV FMEstimatorCallback::computeError */
    double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5];
    double c = F[6] * x1 + F[7] * y1 + F[8];
    double s2 = 1. / (a * a + b * b);
    double d2 = x2 * a + y2 * b + c;
    a = F[0] * x2 + F[3] * y2 + F[6]; b = F[1] * x2 + F[4] * y2 + F[7];
    c = F[2] * x2 + F[5] * y2 + F[8];
    double s1 = 1. / (a * a + b * b);
    double d1 = x1 * a + y1 * b + c;
    double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
    return e1 > e2 ? e1 : e2;
}

int orc_count_inliers_f32(const float F[9], const float *p1, const float *p2, int n,
                          float thr, int metric, uint8_t *mask)
{
    float thr2 = thr * thr;
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        int in = orc_is_inlier_f32(F, p1[2 * i], p1[2 * i + 1], p2[2 * i], p2[2 * i + 1], thr2, metric);
        if (mask) mask[i] = (uint8_t)in;
        cnt += in;
    }
    return cnt;
}

/* ------------------------------------------------------------------------- */
/* RANSAC over caller-supplied samples (north_star contract)                  */
/* ------------------------------------------------------------------------- */

int orc_ransac_f(const float *p1, const float *p2, int n,
                 const int32_t *sample_idx, int nhyp, int m,
                 int metric, float thr, int refit,
                 double F[9], uint8_t *mask, int *n_inliers, int64_t *best_model,
                 int32_t *counts, float *Fs32, int nthreads)
{
    (void)nthreads;
    if (n < m || (m != 7 && m != 8) || nhyp <= 0) return 0;
    int per = m == 8 ? 1 : 3;
    int32_t *best_cnt = (int32_t *)malloc(sizeof(int32_t) * (size_t)nhyp);
    int8_t *best_k = (int8_t *)malloc((size_t)nhyp);
    float *Fs = Fs32 ? Fs32 : (float *)malloc(sizeof(float) * 9 * (size_t)per * nhyp);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int h = 0; h < nhyp; ++h) {
        double Fd[27];
        int ns = m == 8 ? orc_fm_8point_idx(p1, p2, sample_idx + (size_t)h * 8, 8, Fd)
                        : orc_fm_7point_idx(p1, p2, sample_idx + (size_t)h * 7, Fd);
        int bc = 0, bk = -1;
        for (int k = 0; k < per; ++k) {
            float *Ff = Fs + ((size_t)h * per + k) * 9;
            if (k >= ns) { for (int i = 0; i < 9; ++i) Ff[i] = NAN; continue; }
            for (int i = 0; i < 9; ++i) Ff[i] = (float)Fd[9 * k + i];
            int c = orc_count_inliers_f32(Ff, p1, p2, n, thr, metric, NULL);
            if (c > bc) { bc = c; bk = k; }
        }
        best_cnt[h] = bc; best_k[h] = (int8_t)bk;
    }
    int bh = -1, bc = 0;
    for (int h = 0; h < nhyp; ++h)
        if (best_cnt[h] > bc) { bc = best_cnt[h]; bh = h; }
    if (counts) memcpy(counts, best_cnt, sizeof(int32_t) * (size_t)nhyp);
    int ok = 0;
    if (bh >= 0) {
        const float *Fw = Fs + ((size_t)bh * per + best_k[bh]) * 9;
        uint8_t *mk = mask ? mask : (uint8_t *)malloc((size_t)n);
        int c = orc_count_inliers_f32(Fw, p1, p2, n, thr, metric, mk);
        if (n_inliers) *n_inliers = c;
        if (best_model) *best_model = (int64_t)bh * per + best_k[bh];
        for (int i = 0; i < 9; ++i) F[i] = (double)Fw[i];
        ok = 1;
        if (refit && c >= 8) {
            int32_t *ii = (int32_t *)malloc(sizeof(int32_t) * (size_t)c);
            int k = 0;
            for (int i = 0; i < n; ++i) if (mk[i]) ii[k++] = i;
            double Fr[9];
            if (orc_fm_8point_idx(p1, p2, ii, c, Fr)) memcpy(F, Fr, sizeof(Fr));
            free(ii);
        }
        if (!mask) free(mk);
    }
    free(best_cnt); free(best_k);
    if (!Fs32) free(Fs);
    return ok;
}

/* ------------------------------------------------------------------------- */
/* OpenCV-literal estimator: cv::findFundamentalMat as called at main.cpp:95  */
/* ------------------------------------------------------------------------- */

typedef struct { uint64_t state; } cv_rng;
static unsigned cv_rng_next(cv_rng *r)
{
    r->state = (uint64_t)(unsigned)r->state * 4164903690U + (unsigned)(r->state >> 32);
    return (unsigned)r->state;
}
static int cv_rng_uniform(cv_rng *r, int a, int b)
{ return a == b ? a : (int)(cv_rng_next(r) % (unsigned)(b - a) + a); }

static int have_collinear(const float *p, const int32_t *idx, int count)
{
    int i = count - 1;
    for (int j = 0; j < i; ++j) {
        double dx1 = p[2 * idx[j]] - p[2 * idx[i]], dy1 = p[2 * idx[j] + 1] - p[2 * idx[i] + 1];
        for (int k = 0; k < j; ++k) {
            double dx2 = p[2 * idx[k]] - p[2 * idx[i]], dy2 = p[2 * idx[k] + 1] - p[2 * idx[i] + 1];
            if (fabs(dx2 * dy1 - dy2 * dx1) <=
                FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2)))
                return 1;
        }
    }
    return 0;
}

static int cv_get_subset(const float *p1, const float *p2, int count, int mp,
                         cv_rng *rng, int max_attempts, int32_t *idx)
{
    int iters = 0;
    for (; iters < max_attempts; ++iters) {
        int i;
        for (i = 0; i < mp; ++i) {
            int v, dup;
            do {
                v = cv_rng_uniform(rng, 0, count);
                dup = 0;
                for (int k = 0; k < i; ++k) if (idx[k] == v) { dup = 1; break; }
            } while (dup);
            idx[i] = v;
        }
        if (have_collinear(p1, idx, mp) || have_collinear(p2, idx, mp)) continue;
        break;
    }
    return iters < max_attempts;
}

static int cv_update_num_iters(double p, double ep, int model_points, int max_iters)
{
    p = p < 0 ? 0 : (p > 1 ? 1 : p);
    ep = ep < 0 ? 0 : (ep > 1 ? 1 : ep);
    double num = 1. - p > DBL_MIN ? 1. - p : DBL_MIN;
    double denom = 1. - pow(1. - ep, model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num); denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)lrint(num / denom);
}

static void cv_compute_error(const double F[9], const float *p1, const float *p2, int n, float *err)
{
    for (int i = 0; i < n; ++i)
        err[i] = (float)orc_symepi_f64(F, p1[2 * i], p1[2 * i + 1], p2[2 * i], p2[2 * i + 1]);
}

static int cv_find_inliers(const float *err, int n, double thresh, uint8_t *mask)
{
    float t = (float)(thresh * thresh);
    int nz = 0;
    for (int i = 0; i < n; ++i) { int f = err[i] <= t; mask[i] = (uint8_t)f; nz += f; }
    return nz;
}

static int cmp_float(const void *a, const void *b)
{
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

int orc_find_fundamental_cv(const float *p1, const float *p2, int n, int method,
                            double param1, double param2, int max_iters,
                            double F[27], uint8_t *mask)
{
    const int mp = 7;
    if (n < 7) return 0;
    if (n == 7 || method == 2 /* FM_8POINT */) {
        int r = n == 7 ? orc_fm_7point(p1, p2, F) : orc_fm_8point(p1, p2, n, F);
        if (mask) memset(mask, 1, (size_t)n);
        return r;
    }
    if (param1 <= 0) param1 = 3;
    if (param2 < DBL_EPSILON || param2 > 1 - DBL_EPSILON) param2 = 0.99;
    float *err = (float *)malloc(sizeof(float) * (size_t)n);
    uint8_t *cur = (uint8_t *)malloc((size_t)n), *best = (uint8_t *)calloc((size_t)n, 1);
    cv_rng rng = { (uint64_t)-1 };
    int32_t idx[8];
    double model[27], best_model[9];
    int result = 0;
    if ((method & ~3) == 8 /* FM_RANSAC */ && n >= 15) {
        int niters = max_iters > 0 ? max_iters : 1000, max_good = 0;
        for (int iter = 0; iter < niters; ++iter) {
            if (!cv_get_subset(p1, p2, n, mp, &rng, 10000, idx)) { if (iter == 0) goto done; break; }
            int nm = orc_fm_7point_idx(p1, p2, idx, model);
            for (int k = 0; k < nm; ++k) {
                cv_compute_error(model + 9 * k, p1, p2, n, err);
                int good = cv_find_inliers(err, n, param1, cur);
                if (good > (max_good > mp - 1 ? max_good : mp - 1)) {
                    uint8_t *tmp = cur; cur = best; best = tmp;
                    memcpy(best_model, model + 9 * k, sizeof(best_model));
                    max_good = good;
                    niters = cv_update_num_iters(param2, (double)(n - good) / n, mp, niters);
                }
            }
        }
        if (max_good > 0) {
            memcpy(F, best_model, sizeof(best_model));
            if (mask) memcpy(mask, best, (size_t)n);
            result = 1;
        }
    } else {
        /* LMedS: what FM_7POINT with N>7 (the reference's call) dispatches to */
        int niters = cv_update_num_iters(param2, 0.45, mp, max_iters > 0 ? max_iters : 1000);
        if (niters < 3) niters = 3;
        double min_median = DBL_MAX;
        for (int iter = 0; iter < niters; ++iter) {
            if (!cv_get_subset(p1, p2, n, mp, &rng, 1000, idx)) { if (iter == 0) goto done; break; }
            int nm = orc_fm_7point_idx(p1, p2, idx, model);
            for (int k = 0; k < nm; ++k) {
                cv_compute_error(model + 9 * k, p1, p2, n, err);
                qsort(err, (size_t)n, sizeof(float), cmp_float);
                double median = err[n / 2];   /* 4.x: nth_element at count/2 */
                if (median < min_median) {
                    min_median = median;
                    memcpy(best_model, model + 9 * k, sizeof(best_model));
                }
            }
        }
        if (min_median < DBL_MAX) {
            double sigma = 2.5 * 1.4826 * (1 + 5. / (n - mp)) * sqrt(min_median);
            if (sigma < 0.001) sigma = 0.001;
            cv_compute_error(best_model, p1, p2, n, err);
            int good = cv_find_inliers(err, n, sigma, best);
            if (good > 0) {
                memcpy(F, best_model, sizeof(best_model));
                if (mask) memcpy(mask, best, (size_t)n);
                result = 1;
            }
        }
    }
done:
    free(err); free(cur); free(best);
    return result;
}

/* LMedS over caller-supplied 7-point samples: the estimator cv::findFundamentalMat(..., CV_FM_7POINT) runs
 * when N > 7 (main.cpp:95-98), with the product's conventions -- models rounded to f32 first, winner =
 * smallest median (lowest model id 3*hyp+k on ties).  Fs32 in/out: when models_given != 0 the nhyp*3 models
 * ([9] f32 each, NaN = absent) are taken from Fs32 instead of being solved here. */
int orc_lmeds_f(const float *p1, const float *p2, int n, const int32_t *sample_idx, int nhyp,
                double F[9], uint8_t *mask, int *n_inliers, int64_t *best_model, float *medians,
                float *Fs32, int models_given)
{
    if (n < 8 || nhyp <= 0) return 0;
    float *err = (float *)malloc(sizeof(float) * (size_t)n);
    float *Fs = Fs32 ? Fs32 : (float *)malloc(sizeof(float) * 27 * (size_t)nhyp);
    double best_med = DBL_MAX; int64_t bm = -1;
    for (int h = 0; h < nhyp; ++h) {
        double Fd[27];
        int ns = 3;
        if (!models_given) ns = orc_fm_7point_idx(p1, p2, sample_idx + (size_t)h * 7, Fd);
        for (int k = 0; k < 3; ++k) {
            float *Ff = Fs + ((size_t)h * 3 + k) * 9;
            if (!models_given) for (int i = 0; i < 9; ++i) Ff[i] = k < ns ? (float)Fd[9 * k + i] : NAN;
            float med = INFINITY;
            int ok = 1;
            for (int i = 0; i < 9; ++i) ok = ok && isfinite(Ff[i]);
            if (ok) {
                double Fw[9];
                for (int i = 0; i < 9; ++i) Fw[i] = (double)Ff[i];
                cv_compute_error(Fw, p1, p2, n, err);
                for (int i = 0; i < n; ++i) if (!(err[i] >= 0.f)) ok = 0;
                if (ok) { qsort(err, (size_t)n, sizeof(float), cmp_float); med = err[n / 2]; }
            }
            if (medians) medians[(size_t)h * 3 + k] = med;
            if (med < INFINITY && (double)med < best_med) { best_med = med; bm = (int64_t)h * 3 + k; }
        }
    }
    int result = 0;
    if (bm >= 0) {
        double Fw[9];
        for (int i = 0; i < 9; ++i) Fw[i] = (double)Fs[(size_t)bm * 9 + i];
        double sigma = 2.5 * 1.4826 * (1 + 5. / (n - 7)) * sqrt(best_med);
        if (sigma < 0.001) sigma = 0.001;
        cv_compute_error(Fw, p1, p2, n, err);
        uint8_t *mk = mask ? mask : (uint8_t *)malloc((size_t)n);
        float t = (float)(sigma * sigma);
        int nz = 0;
        for (int i = 0; i < n; ++i) { int f = err[i] <= t; mk[i] = (uint8_t)f; nz += f; }
        if (!mask) free(mk);
        memcpy(F, Fw, sizeof(Fw));
        if (n_inliers) *n_inliers = nz;
        if (best_model) *best_model = bm;
        result = 1;
    }
    free(err);
    if (!Fs32) free(Fs);
    return result;
}

/* ------------------------------------------------------------------------- */
/* epilines: main.cpp:127-132                                                 */
/* ------------------------------------------------------------------------- */

void orc_epilines(const float *pts, int n, int which, const double F[9], float *lines)
{
    double f[9];
    if (which == 2) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) f[i * 3 + j] = F[j * 3 + i]; }
    else memcpy(f, F, sizeof(f));
    for (int i = 0; i < n; ++i) {
        double x = pts[2 * i], y = pts[2 * i + 1];
        double a = f[0] * x + f[1] * y + f[2];
        double b = f[3] * x + f[4] * y + f[5];
        double c = f[6] * x + f[7] * y + f[8];
        double nu = a * a + b * b;
        nu = nu ? 1. / sqrt(nu) : 1.;
        lines[3 * i] = (float)(a * nu); lines[3 * i + 1] = (float)(b * nu); lines[3 * i + 2] = (float)(c * nu);
    }
}
