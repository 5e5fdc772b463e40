// hamming.cu -- K4: binary-descriptor Hamming kNN-2 (ORB/BRIEF) for sm_100a.
//
// Replaces BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2) -- the binary-descriptor form of
// the matcher the reference instantiates at /root/reference/Points Matching/main.cpp:43-46.
//
// Design (DESIGN.md "K4"): all-pairs, so operands are served from registers/shared
// memory, not HBM.  Each thread keeps QPT query rows in registers; the CTA streams a
// chunk of train rows through a double-buffered shared-memory tile (cp.async) and every
// thread reads the same train row with broadcast LDS.128.  Per pair: W xor + W popc,
// then a packed (dist<<16 | local train index) key goes through a 3-op min/max top-2
// network -- keys are unique, so the lowest train index wins ties exactly like OpenCV.
// Train rows are split into chunks across blockIdx.y; a finalize kernel merges the
// per-chunk (best, second) pairs.
#include <cuda_bf16.h>
#include "pm_internal.h"
#include "l2_common.h"

namespace {

constexpr int HAM_THREADS = 128;
constexpr int HAM_QPT = 4;       // query rows per thread
constexpr int HAM_TT = 128;      // train rows per shared-memory tile

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// Pads `bytes`-wide rows to W 32-bit words (zero fill: xor of the pads is 0).
__global__ void ham_pack_kernel(const uint8_t *__restrict__ src, int n, int bytes, int W,
                                uint32_t *__restrict__ dst)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)n * W;
    if (i >= total) return;
    int row = (int)(i / W), w = (int)(i % W);
    uint32_t v = 0;
    const uint8_t *p = src + (size_t)row * bytes + (size_t)w * 4;
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (w * 4 + b < bytes) v |= (uint32_t)p[b] << (8 * b);
    dst[i] = v;
}

// Hamming distance of two W-word rows.  POPC runs on the 16-lane XU pipe, which bounds the kernel (ncu r1a:
// XU 91.5% busy), while the 64-lane ALU pipe idles; carry-save adders (xor3 / majority = one LOP3 each)
// fold three words into a "ones" and a "twos" word first, so 8 words cost 5 POPC + 6 LOP3 instead of 8 POPC.
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }
// `one` is the constant 1 passed as a kernel argument: x * one + y stays an IMAD (FMA pipe) instead of an
// IADD on the ALU pipe, which is the pipe that bounds the kernel once the POPC count is down (ncu r1c: ALU 89%).
template <int W>
__device__ __forceinline__ uint32_t ham_dist(const uint32_t (&q)[W], const uint32_t (&t)[W], uint32_t one)
{
    uint32_t x[W];
#pragma unroll
    for (int w = 0; w < W; ++w) x[w] = q[w] ^ t[w];
    uint32_t ones = 0, twos = 0;
#pragma unroll
    for (int g = 0; g + 8 <= W; g += 8) {
        const uint32_t s1 = xor3(x[g], x[g + 1], x[g + 2]), c1 = maj3(x[g], x[g + 1], x[g + 2]);
        const uint32_t s2 = xor3(x[g + 3], x[g + 4], x[g + 5]), c2 = maj3(x[g + 3], x[g + 4], x[g + 5]);
        const uint32_t s3 = xor3(s1, s2, x[g + 6]), c3 = maj3(s1, s2, x[g + 6]);
        ones = __popc(s3) * one + (__popc(x[g + 7]) * one + ones);
        twos = __popc(c1) * one + (__popc(c2) * one + (__popc(c3) * one + twos));
    }
    if (W % 8 == 4) {
        constexpr int g = W - 4;
        ones += __popc(xor3(x[g], x[g + 1], x[g + 2])) + __popc(x[g + 3]);
        twos += __popc(maj3(x[g], x[g + 1], x[g + 2]));
    }
    return twos * (one + one) + ones;
}

// part[(chunk * nq + qi) * 2 + {0,1}] = (dist << 32 | global train index), ~0 if absent.
template <int W>
__global__ void __launch_bounds__(HAM_THREADS)
ham_knn2_kernel(const uint32_t *__restrict__ q, int nq, const uint32_t *__restrict__ t, int nt,
                int chunk_rows, unsigned long long *__restrict__ part, uint32_t one)
{
    __shared__ __align__(16) uint32_t tile[2][HAM_TT * W];
    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * (HAM_THREADS * HAM_QPT);
    const int c0 = blockIdx.y * chunk_rows;
    const int c1 = min(nt, c0 + chunk_rows);
    if (c0 >= c1) {
        // empty chunk: still publish "absent" so finalize can read unconditionally
#pragma unroll
        for (int r = 0; r < HAM_QPT; ++r) {
            int qi = q0 + r * HAM_THREADS + tid;
            if (qi < nq) {
                size_t o = ((size_t)blockIdx.y * nq + qi) * 2;
                part[o] = ~0ull; part[o + 1] = ~0ull;
            }
        }
        return;
    }

    uint32_t qr[HAM_QPT][W];
#pragma unroll
    for (int r = 0; r < HAM_QPT; ++r) {
        int qi = min(q0 + r * HAM_THREADS + tid, nq - 1);
        const uint4 *src = reinterpret_cast<const uint4 *>(q + (size_t)qi * W);
#pragma unroll
        for (int w = 0; w < W / 4; ++w) {
            uint4 v = __ldg(src + w);
            qr[r][4 * w] = v.x; qr[r][4 * w + 1] = v.y; qr[r][4 * w + 2] = v.z; qr[r][4 * w + 3] = v.w;
        }
    }
    uint32_t m1[HAM_QPT], m2[HAM_QPT];
#pragma unroll
    for (int r = 0; r < HAM_QPT; ++r) { m1[r] = 0xFFFFFFFFu; m2[r] = 0xFFFFFFFFu; }

    const int ntiles = (c1 - c0 + HAM_TT - 1) / HAM_TT;
    constexpr int V16 = HAM_TT * W / 4;                  // 16-byte vectors per tile
    auto load_tile = [&](int ti, int buf) {
        const int r0 = c0 + ti * HAM_TT;
        const int rows = min(HAM_TT, c1 - r0);
        const uint4 *src = reinterpret_cast<const uint4 *>(t + (size_t)r0 * W);
        for (int v = tid; v < rows * W / 4; v += HAM_THREADS)
            cp_async16(&tile[buf][v * 4], src + v);
        (void)V16;
        cp_async_commit();
    };
    load_tile(0, 0);
    for (int ti = 0; ti < ntiles; ++ti) {
        const int buf = ti & 1;
        if (ti + 1 < ntiles) { load_tile(ti + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const int rows = min(HAM_TT, c1 - (c0 + ti * HAM_TT));
        const uint32_t jbase = (uint32_t)(ti * HAM_TT);
#pragma unroll 2
        for (int j = 0; j < rows; ++j) {
            uint32_t tw[W];
            const uint4 *row = reinterpret_cast<const uint4 *>(&tile[buf][j * W]);
#pragma unroll
            for (int w = 0; w < W / 4; ++w) {
                uint4 v = row[w];
                tw[4 * w] = v.x; tw[4 * w + 1] = v.y; tw[4 * w + 2] = v.z; tw[4 * w + 3] = v.w;
            }
#pragma unroll
            for (int r = 0; r < HAM_QPT; ++r) {
                const uint32_t d = ham_dist<W>(qr[r], tw, one);
                uint32_t key = d * (one << 16) + (jbase + (uint32_t)j);
                m2[r] = min(m2[r], max(m1[r], key));
                m1[r] = min(m1[r], key);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < HAM_QPT; ++r) {
        int qi = q0 + r * HAM_THREADS + tid;
        if (qi < nq) {
            size_t o = ((size_t)blockIdx.y * nq + qi) * 2;
            part[o] = m1[r] == 0xFFFFFFFFu ? ~0ull
                      : ((unsigned long long)(m1[r] >> 16) << 32) | (unsigned)(c0 + (m1[r] & 0xFFFFu));
            part[o + 1] = m2[r] == 0xFFFFFFFFu ? ~0ull
                          : ((unsigned long long)(m2[r] >> 16) << 32) | (unsigned)(c0 + (m2[r] & 0xFFFFu));
        }
    }
}

// Any width (W words, runtime): one thread per query, rows re-read through L1.
__global__ void ham_knn2_generic_kernel(const uint32_t *__restrict__ q, int nq,
                                        const uint32_t *__restrict__ t, int nt, int W,
                                        int chunk_rows, unsigned long long *__restrict__ part)
{
    int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const int c0 = blockIdx.y * chunk_rows, c1 = min(nt, c0 + chunk_rows);
    unsigned long long m1 = ~0ull, m2 = ~0ull;
    const uint32_t *a = q + (size_t)qi * W;
    for (int j = c0; j < c1; ++j) {
        const uint32_t *b = t + (size_t)j * W;
        uint32_t d = 0;
        for (int w = 0; w < W; ++w) d += __popc(__ldg(a + w) ^ __ldg(b + w));
        unsigned long long key = ((unsigned long long)d << 32) | (unsigned)j;
        m2 = min(m2, max(m1, key));
        m1 = min(m1, key);
    }
    size_t o = ((size_t)blockIdx.y * nq + qi) * 2;
    part[o] = m1; part[o + 1] = m2;
}

// Merges the per-chunk pairs.  mode 0: DMatch[nq][2]; mode 1: packed column minimum
// (float_bits(dist) << 32 | index + base) for the cross-check exchange.
__global__ void ham_finalize_kernel(const unsigned long long *__restrict__ part, int nq, int nchunks,
                                    int q_index_base, int mode, pm_dmatch *__restrict__ out,
                                    unsigned long long *__restrict__ col_best)
{
    int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    unsigned long long m1 = ~0ull, m2 = ~0ull;
    for (int c = 0; c < nchunks; ++c) {
        size_t o = ((size_t)c * nq + qi) * 2;
        unsigned long long a = part[o], b = part[o + 1];
        m2 = min(m2, max(m1, a)); m1 = min(m1, a);
        m2 = min(m2, max(m1, b)); m1 = min(m1, b);
    }
    if (mode == 0) {
        pm_dmatch r0, r1;
        r0.queryIdx = r1.queryIdx = qi + q_index_base;
        r0.imgIdx = r1.imgIdx = 0;
        r0.trainIdx = m1 == ~0ull ? -1 : (int)(m1 & 0xFFFFFFFFu);
        r0.distance = m1 == ~0ull ? 3.402823466e+38f : (float)(unsigned)(m1 >> 32);
        r1.trainIdx = m2 == ~0ull ? -1 : (int)(m2 & 0xFFFFFFFFu);
        r1.distance = m2 == ~0ull ? 3.402823466e+38f : (float)(unsigned)(m2 >> 32);
        out[(size_t)qi * 2] = r0;
        out[(size_t)qi * 2 + 1] = r1;
    } else {
        // here "query" rows are the train set and the neighbours are query indices
        unsigned long long k = ~0ull;
        if (m1 != ~0ull) {
            float d = (float)(unsigned)(m1 >> 32);
            k = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)((int)(m1 & 0xFFFFFFFFu) + q_index_base);
        }
        col_best[qi] = k;
    }
}

// =====================================================================================
// Tensor-core Hamming.  For 0/1 vectors ||a - b||^2 IS the Hamming distance, so K2
// (l2_tc.cu) runs unchanged on E4M3 operands: bits are expanded to bytes (query 1.0 = 0x38, train
// -2.0 = 0xC0), ||b||^2 = popc(b) rides in the norm step, everything is an exact small integer.
// The POPC kernel above is pinned at its pipe limits (ncu r1c: ALU 89%, XU 68%); this path is bound by
// K2's MMA rate instead.  Rows of up to 256 bits.
// =====================================================================================
constexpr int HT_W = 8;          // words per row (256 bits)

// eight lanes per row (four rows per warp at a time): lane s expands word s of the row into 32 operand bytes
__global__ void __launch_bounds__(256)
ham_expand_kernel(const uint32_t *__restrict__ q, int nq, int mq_pad, const uint32_t *__restrict__ t, int nt, int nt_pad,
                  uint8_t *__restrict__ qx, uint8_t *__restrict__ tx, float *__restrict__ qnorm, uint8_t *__restrict__ text,
                  L2Flags *flags, L2Cand *__restrict__ part, int part_per_row)
{
    pm_pdl_prologue();
    const int lane = threadIdx.x & 31, sub = lane & 7;
    const int total = mq_pad + nt_pad, stride = gridDim.x * (blockDim.x >> 3);
    for (int grow = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); grow < total; grow += stride) {   // mq_pad, nt_pad: multiples of 256
        const bool is_train = grow >= mq_pad;
        const int row = is_train ? grow - mq_pad : grow;
        const int n = is_train ? nt : nq;
        const uint32_t w = row < n ? __ldg((is_train ? t : q) + (size_t)row * HT_W + sub) : 0u;
        int pc = __popc(w);
        pc += __shfl_xor_sync(0xffffffffu, pc, 4);
        pc += __shfl_xor_sync(0xffffffffu, pc, 2);
        pc += __shfl_xor_sync(0xffffffffu, pc, 1);
        const uint32_t one = is_train ? 0xC0u : 0x38u;                 // E4M3 -2.0 / 1.0
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)                                     // four bits -> four bytes (bit b of the row -> byte b)
            o[k] = ((((w >> (4 * k)) & 0xFu) * 0x00204081u) & 0x01010101u) * one;
        uint4 *dst = reinterpret_cast<uint4 *>((is_train ? tx : qx) + (size_t)row * L2_PACK_COLS + 32 * sub);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        if (is_train) {
            if (sub < 2) {
                const uint4 s3 = bf16_split3(row < n ? (float)pc : __uint_as_float(L2_PAD_NORM_BITS));
                uint8_t *e = text + (size_t)(row >> 7) * L2_EXT_BYTES + ext_row_offset(row & 127) + sub * 128;
                *reinterpret_cast<uint4 *>(e) = sub == 0 ? make_uint4(s3.x, s3.y | 0x3F800000u, 0x3F803F80u, 0u)
                                                         : make_uint4(0u, 0u, 0u, 0u);
            }
        } else {
            if (sub == 0) qnorm[row] = (float)pc;
            for (int k = sub; k < part_per_row; k += 8) part[(size_t)row * part_per_row + k] = L2Cand{L2_INF, -1};
        }
    }
    // integer data with tiny norms: K2 / the finish kernel run in exact mode
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        flags->nonexact = 0;
        flags->max_tnorm_bits = __float_as_uint(256.f);
        flags->max_qnorm_bits = __float_as_uint(256.f);
    }
}

// 8 lanes per query row: merge K2's quad candidates, recompute the 8 columns of the two best quads with
// popc on the raw rows (lane s owns word s), top-2 by (distance, index).  mode 0: DMatch rows; mode 1:
// packed column minimum float_bits(dist) << 32 | (index + base) for the cross-check exchange.
__global__ void __launch_bounds__(256)
ham_tc_finish_kernel(const L2Cand *__restrict__ part, int ncand, const uint32_t *__restrict__ q, const uint32_t *__restrict__ t,
                     int nq, int nt, L2Flags *flags_next, int q_index_base, int mode, pm_dmatch *__restrict__ out,
                     unsigned long long *__restrict__ col_best)
{
    pm_pdl_prologue();
    const int lane = threadIdx.x & 31, sub = lane & 7;
    if (blockIdx.x == 0 && threadIdx.x == 0) *flags_next = L2Flags{0, 0u, 0u, 0, 0u, 0u, 0, {0}};
    const int ngroups = gridDim.x * (blockDim.x >> 3);
    const int nq_round = (nq + 3) & ~3;
    for (int i = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); i < nq_round; i += ngroups) {
        const bool live_row = i < nq;
        const int ir = live_row ? i : nq - 1;
        const L2Cand *prow = part + (size_t)ir * ncand;
        L2Cand ca = L2Cand{L2_INF, -1}, cb = L2Cand{L2_INF, -1};
        if (sub < ncand) ca = prow[sub];
        if (sub + 8 < ncand) cb = prow[sub + 8];
        const uint32_t aw = q[(size_t)ir * HT_W + sub];
        const unsigned long long ka = cand_key(ca, nt), kb = cand_key(cb, nt);
        unsigned long long h0 = min_u64(ka, kb), h1 = max_u64(ka, kb), h2 = ~0ull;
        for (int c0 = sub + 16; c0 < ncand; c0 += 8) {
            const unsigned long long key = cand_key(prow[c0], nt);
            h2 = min_u64(h2, max_u64(h1, key));
            h1 = min_u64(max_u64(h0, key), h1);
            h0 = min_u64(h0, key);
        }
        unsigned long long k[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            k[r] = group_min_u64(0u, h0);
            const bool pop = h0 == k[r] && h0 != ~0ull;
            h0 = pop ? h1 : h0; h1 = pop ? h2 : h1; h2 = pop ? ~0ull : h2;
        }
        const int cand = sub >> 2, mem = sub & 3;
        const bool live0 = k[0] != ~0ull, live1 = k[1] != ~0ull;
        const int jq0 = live0 ? (int)(k[0] & 0xFFFFFFFFu) : 0, jq1 = live1 ? (int)(k[1] & 0xFFFFFFFFu) : 0;
        const int my_col = (cand ? jq1 : jq0) + mem;
        unsigned dots[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int col = (c >> 2 ? jq1 : jq0) + (c & 3);
            dots[c] = (unsigned)__popc(aw ^ __ldg(t + (size_t)min(col, nt - 1) * HT_W + sub));
        }
        const unsigned my_d = group_transpose_sum(dots, sub);
        const bool mine = (cand ? live1 : live0) && my_col < nt;
        unsigned long long key = mine ? (((unsigned long long)__float_as_uint((float)my_d) << 32) | (unsigned)my_col) : ~0ull;
        const unsigned long long w0 = group_min_u64(0u, key);
        key = key == w0 ? ~0ull : key;
        const unsigned long long w1 = group_min_u64(0u, key);
        if (!live_row) continue;
        if (mode == 0) {
            if (sub < 2) {
                const unsigned long long w = sub == 0 ? w0 : w1;
                reinterpret_cast<uint4 *>(out + (size_t)i * 2)[sub] =
                    w == ~0ull ? make_uint4((unsigned)(i + q_index_base), 0xFFFFFFFFu, 0u, __float_as_uint(3.402823466e+38f))
                               : make_uint4((unsigned)(i + q_index_base), (unsigned)(w & 0xFFFFFFFFull), 0u, (unsigned)(w >> 32));
            }
        } else if (sub == 0) {
            // here the "query" rows are the train set and the neighbours are query indices of this shard
            col_best[i] = w0 == ~0ull ? ~0ull : ((w0 & 0xFFFFFFFF00000000ull) | (unsigned)((int)(w0 & 0xFFFFFFFFull) + q_index_base));
        }
    }
}

static int g_ham_path = 0;       // 0 auto, 1 POPC kernel, 2 tensor-core kernel (debug / bench switch)

int ham_tc_run(pm_ctx *ctx, const uint32_t *pq, int nq, const uint32_t *pt, int nt, int q_index_base, int mode,
               pm_dmatch *dout, uint64_t *dcol)
{
    const int mq_pad = pm_round_up(nq, 256), nt_pad = pm_round_up(nt, 256);
    const int MT = mq_pad / 256, NT = nt_pad / 128;
    const int smax = l2_tc_smax(ctx, MT, NT);
    L2Flags *flags, *flags_next, *tflags_unused;
    { int fst = l2_flags_acquire(ctx, &flags, &flags_next, &tflags_unused); if (fst != PM_OK) return fst; }
    PM_WS(ctx, qx, uint8_t *, WS_Q_PACK, (size_t)mq_pad * L2_PACK_COLS);
    PM_WS(ctx, tx, uint8_t *, WS_T_PACK, (size_t)nt_pad * L2_PACK_COLS);
    PM_WS(ctx, qnorm, float *, WS_Q_NORM, (size_t)mq_pad * 4);
    PM_WS(ctx, text, uint8_t *, WS_T_NORM, (size_t)(nt_pad / 128) * L2_EXT_BYTES);
    PM_WS(ctx, part, L2Cand *, WS_L2_PART, (size_t)mq_pad * smax * 3 * sizeof(L2Cand));
    const int xblocks = min(pm_cdiv(mq_pad + nt_pad, 32), 8 * ctx->num_sms);
    PM_CUDA(ctx, pm_launch_pdl(ham_expand_kernel, dim3(xblocks), dim3(256), 0, ctx->stream, pq, nq, mq_pad, pt, nt, nt_pad, qx, tx,
                               qnorm, text, flags, part, smax * 3));
    PM_CHECK_LAUNCH(ctx);
    int st = l2_tc_launch(ctx, qx, mq_pad, tx, nt_pad, text, flags, part, smax, pm_l2_dump_ptr(), 1);
    if (st != PM_OK) return st;
    const int fblocks = min(pm_cdiv(nq, 32), 4 * ctx->num_sms);
    PM_CUDA(ctx, pm_launch_pdl(ham_tc_finish_kernel, dim3(fblocks), dim3(256), 0, ctx->stream, (const L2Cand *)part, smax * 3, pq, pt,
                               nq, nt, flags_next, q_index_base, mode, dout, (unsigned long long *)dcol));
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

int padded_words(int bytes)
{
    int w = (bytes + 3) / 4;
    if (w <= 4) return 4;
    if (w <= 8) return 8;
    if (w <= 16) return 16;
    return (w + 3) / 4 * 4;
}

// Returns device pointer to W-word rows (in place when already in that layout).
int ham_prepare(pm_ctx *ctx, const uint8_t *d, int n, int bytes, int W, int slot, const uint32_t **out)
{
    if (bytes == W * 4 && ((uintptr_t)d & 15) == 0) { *out = reinterpret_cast<const uint32_t *>(d); return PM_OK; }
    PM_WS(ctx, p, uint32_t *, slot, (size_t)(n > 0 ? n : 1) * W * 4);
    long long total = (long long)n * W;
    if (total > 0) {
        ham_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d, n, bytes, W, p);
        PM_CHECK_LAUNCH(ctx);
    }
    *out = p;
    return PM_OK;
}

int ham_run(pm_ctx *ctx, const uint8_t *dq, int nq, const uint8_t *dt, int nt, int bytes,
            int q_index_base, int mode, pm_dmatch *dout, uint64_t *dcol)
{
    if (nq <= 0) return PM_OK;
    // tensor-core path: rows of up to 256 bits and enough pairs to fill the persistent GEMM
    const bool tc_ok = bytes <= 4 * HT_W && nt > 0;
    const bool use_tc = tc_ok && (g_ham_path == 2 || (g_ham_path == 0 && (long long)nq * nt >= (1ll << 22)));
    const int W = use_tc ? HT_W : padded_words(bytes);
    const uint32_t *pq, *pt;
    int st;
    if ((st = ham_prepare(ctx, dq, nq, bytes, W, WS_HAM_Q, &pq)) != PM_OK) return st;
    if ((st = ham_prepare(ctx, dt, nt, bytes, W, WS_HAM_T, &pt)) != PM_OK) return st;

    if (use_tc) return ham_tc_run(ctx, pq, nq, pt, nt, q_index_base, mode, dout, dcol);
    const bool fast = (W == 4 || W == 8 || W == 16);
    const int qblocks = fast ? pm_cdiv(nq, HAM_THREADS * HAM_QPT) : pm_cdiv(nq, 128);
    // One resident wave: as many CTAs as the GPU holds at once (occupancy x SMs), so every SM carries the
    // same number of equal work units (4 CTAs per SM on 600 CTAs left SMs with 5 vs 4: 81% balance, ncu
    // r1a).  Chunks are a multiple of the tile and <= 65536 rows (16-bit local index).
    static int occ = 0;
    if (!occ) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ham_knn2_kernel<8>, HAM_THREADS, 0) != cudaSuccess || occ < 1) occ = 4;
    }
    int want = max(1, (occ * ctx->num_sms) / qblocks);
    int chunks = nt > 0 ? max(1, min(want, pm_cdiv(nt, HAM_TT))) : 1;
    int chunk_rows = nt > 0 ? pm_round_up(pm_cdiv(nt, chunks), HAM_TT) : HAM_TT;
    if (chunk_rows > 65536) chunk_rows = 65536;
    chunks = nt > 0 ? pm_cdiv(nt, chunk_rows) : 1;

    PM_WS(ctx, part, unsigned long long *, WS_HAM_PART, (size_t)chunks * nq * 2 * sizeof(unsigned long long));
    dim3 grid(qblocks, chunks);
    {
    pm_prof_scope prof(ctx, 1);
    if (W == 4)       ham_knn2_kernel<4><<<grid, HAM_THREADS, 0, ctx->stream>>>(pq, nq, pt, nt, chunk_rows, part, 1u);
    else if (W == 8)  ham_knn2_kernel<8><<<grid, HAM_THREADS, 0, ctx->stream>>>(pq, nq, pt, nt, chunk_rows, part, 1u);
    else if (W == 16) ham_knn2_kernel<16><<<grid, HAM_THREADS, 0, ctx->stream>>>(pq, nq, pt, nt, chunk_rows, part, 1u);
    else ham_knn2_generic_kernel<<<grid, 128, 0, ctx->stream>>>(pq, nq, pt, nt, W, chunk_rows, part);
    }
    PM_CHECK_LAUNCH(ctx);
    ham_finalize_kernel<<<pm_cdiv(nq, 256), 256, 0, ctx->stream>>>(part, nq, chunks, q_index_base, mode,
                                                                  dout, (unsigned long long *)dcol);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

}  // namespace

// Debug / bench switch: 0 auto (tensor cores for large problems), 1 force the POPC kernel, 2 force tensor cores.
extern "C" void pm_debug_hamming_path(int path) { g_ham_path = path; }

int pmk_hamming_knn2(pm_ctx *ctx, const uint8_t *dq, int nq, const uint8_t *dt, int nt, int bytes,
                     int q_index_base, pm_dmatch *dout)
{
    return ham_run(ctx, dq, nq, dt, nt, bytes, q_index_base, 0, dout, nullptr);
}

// Column minima over the query shard == nearest "query" for every train row: the same
// kernel with the roles swapped (rows = train set, columns = this rank's queries).
int pmk_hamming_col_best(pm_ctx *ctx, const uint8_t *dq, int nq, const uint8_t *dt, int nt, int bytes,
                         int q_index_base, uint64_t *dcol_best)
{
    return ham_run(ctx, dt, nt, dq, nq, bytes, q_index_base, 1, nullptr, dcol_best);
}
