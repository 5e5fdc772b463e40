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
#include "pm_internal.h"

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
    const int W = padded_words(bytes);
    const uint32_t *pq, *pt;
    int st;
    if ((st = ham_prepare(ctx, dq, nq, bytes, W, WS_HAM_Q, &pq)) != PM_OK) return st;
    if ((st = ham_prepare(ctx, dt, nt, bytes, W, WS_HAM_T, &pt)) != PM_OK) return st;

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
