// l2_fallback.cuh -- exact FP32 kNN-2 scan for the query rows the split-mode finish kernel (K3, l2.cu) could not
// certify.  Barrier-free in the sense that matters: no block ever waits for a block that may not have started.  The work
// is a list of (flagged row, train segment) items handed out by an atomic counter to whatever blocks take part; every item
// leaves its top-2 in a scratch slot and the block that completes a row's LAST segment merges the row (a per-row
// countdown).  K3 runs it in its own tail (l2_finish_kernel): extra "helper" blocks that first check that every row block
// has STARTED (else they leave), then wait for the row blocks to finish; the last row block to finish takes part as well,
// so the scan completes even if no helper stayed.
// Distances follow the "re-rank order" of l2.cu / DESIGN.md bit for bit (the same group_l2sq).
#pragma once
#include "pm_internal.h"
#include "l2_common.h"

struct L2FallbackArgs {
    const void *q, *t;            // raw descriptors (f32 or u8 rows of `dim` elements)
    int is_u8, nq, nt, dim, vec, q_index_base;
    L2Flags *flags;               // n_flagged (the row blocks wrote it), next_item (the work queue)
    const int *flagged;           // [n_flagged] query rows
    unsigned long long *fb_part;  // [<= 2 * items] per-item (best, second) keys
    unsigned *fb_cnt;             // [n_flagged when a row is split] segments done; returns to zero
    pm_dmatch *out;               // [nq][2]
    int workers;                  // blocks expected to take part (sizes the split of a row into segments)
};

// the scan runs on at most this many blocks (grid stride); bounds the scratch: < 2 * L2FB_MAX_GRID items of two
// 8-byte keys, then L2FB_MAX_GRID row countdowns
constexpr int L2FB_MAX_GRID = 1024;
constexpr size_t L2FB_SCRATCH_BYTES = (size_t)4 * L2FB_MAX_GRID * 8 + (size_t)L2FB_MAX_GRID * 4;

#ifdef __CUDACC__
constexpr int L2FB_MAX_WARPS = 32;

__device__ __forceinline__ void l2fb_load_row8(const float *p, int sub, int dim, bool vec, float (&v)[4][4])
{
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int l = sub + 8 * e;
        if (vec) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(p) + l);
            v[e][0] = x.x; v[e][1] = x.y; v[e][2] = x.z; v[e][3] = x.w;
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) v[e][c] = 4 * l + c < dim ? __ldg(p + 4 * l + c) : 0.f;
        }
    }
}
__device__ __forceinline__ void l2fb_load_row8(const uint8_t *p, int sub, int dim, bool vec, float (&v)[4][4])
{
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int l = sub + 8 * e;
        if (vec) {
            const uchar4 x = __ldg(reinterpret_cast<const uchar4 *>(p) + l);
            v[e][0] = x.x; v[e][1] = x.y; v[e][2] = x.z; v[e][3] = x.w;
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) v[e][c] = 4 * l + c < dim ? (float)__ldg(p + 4 * l + c) : 0.f;
        }
    }
}
// sum (a-b)^2 in the re-rank order: virtual-lane partials (fma chain over e = 0..3), then the xor butterfly
// 16 | 8 (inside the lane) and 4 | 2 | 1 (across the 8 lanes of the group)
__device__ __forceinline__ float l2fb_group_l2sq(const float (&a)[4][4], const float (&b)[4][4], int sub, int dim)
{
    float p[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        p[e] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (4 * (sub + 8 * e) + c < dim) { const float d = a[e][c] - b[e][c]; p[e] = fmaf(d, d, p[e]); }
    }
    float r = (p[0] + p[2]) + (p[1] + p[3]);          // l ^ 16, then l ^ 8
    r += __shfl_xor_sync(0xffffffffu, r, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

__device__ __forceinline__ uint4 l2fb_record(int qidx, unsigned long long w)
{
    return w == ~0ull ? make_uint4((unsigned)qidx, 0xFFFFFFFFu, 0u, __float_as_uint(3.402823466e+38f))
                      : make_uint4((unsigned)qidx, (unsigned)(w & 0xFFFFFFFFull), 0u,
                                   __float_as_uint(sqrtf(__uint_as_float((unsigned)(w >> 32)))));
}

// Takes items from the queue until it is empty (whole block; blockDim.x a multiple of 32, <= 1024).  n_flagged must be
// final: call it only once every row block is known to have finished.
template <typename T>
__device__ void l2_fallback_items(const L2FallbackArgs &A)
{
    __shared__ float x_qs[L2_KDIM];
    __shared__ unsigned long long x_k[L2FB_MAX_WARPS][2];
    __shared__ int x_item;
    const int nf = *reinterpret_cast<volatile int *>(&A.flags->n_flagged);
    if (nf <= 0) return;
    const int stride = A.workers > 0 ? A.workers : 1;
    const T *q = reinterpret_cast<const T *>(A.q), *t = reinterpret_cast<const T *>(A.t);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, sub = lane & 7, g = lane >> 3;
    const int chunk = nwarps * 32;                                  // train rows per pass of the block
    const int nchunk = (A.nt + chunk - 1) / chunk;
    // few flagged rows: split every row into `split` segments so the whole grid has work; many: one item per row
    int split = 1;
    if (nf < stride) { split = (stride + nf - 1) / nf; if (split > nchunk) split = nchunk; }
    const int seg_chunks = (nchunk + split - 1) / split;
    split = (nchunk + seg_chunks - 1) / seg_chunks;
    const int items = nf * split;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) x_item = atomicAdd(&A.flags->next_item, 1);
        __syncthreads();
        const int item = x_item;
        if (item >= items) break;
        const int r = item / split, sg = item - r * split;
        const int i = *reinterpret_cast<volatile const int *>(&A.flagged[r]);
        if (threadIdx.x < L2_KDIM) x_qs[threadIdx.x] = (int)threadIdx.x < A.dim ? (float)q[(size_t)i * A.dim + threadIdx.x] : 0.f;
        __syncthreads();
        float a[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int c = 0; c < 4; ++c) a[e][c] = x_qs[4 * (sub + 8 * e) + c];
        unsigned long long k0 = ~0ull, k1 = ~0ull;
        const int ch_end = min(nchunk, (sg + 1) * seg_chunks);
        for (int ch = sg * seg_chunks; ch < ch_end; ++ch) {
            const int j0 = ch * chunk + warp * 32;
#pragma unroll 2
            for (int it = 0; it < 8; ++it) {                    // 4 train rows per warp and iteration
                const int j = j0 + it * 4 + g;
                float b[4][4];
                l2fb_load_row8(t + (size_t)min(j, A.nt - 1) * A.dim, sub, A.dim, A.vec != 0, b);
                const float d = l2fb_group_l2sq(a, b, sub, A.dim);
                const unsigned long long key = j < A.nt ? (((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j) : ~0ull;
                k1 = min_u64(k1, max_u64(k0, key));
                k0 = min_u64(k0, key);
            }
        }
        // merge the 4 groups of the warp (lanes of a group hold identical keys), then the warps
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            const unsigned long long y0 = __shfl_xor_sync(0xffffffffu, k0, o), y1 = __shfl_xor_sync(0xffffffffu, k1, o);
            const unsigned long long lo = min_u64(k0, y0), hi = max_u64(k0, y0);
            k1 = min_u64(min_u64(k1, y1), hi); k0 = lo;
        }
        if (lane == 0) { x_k[warp][0] = k0; x_k[warp][1] = k1; }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long m0 = ~0ull, m1 = ~0ull;
            for (int w = 0; w < nwarps; ++w)
                for (int e = 0; e < 2; ++e) { const unsigned long long key = x_k[w][e]; m1 = min_u64(m1, max_u64(m0, key)); m0 = min_u64(m0, key); }
            int last = 1;
            if (split > 1) {
                A.fb_part[(size_t)item * 2] = m0;
                A.fb_part[(size_t)item * 2 + 1] = m1;
                __threadfence();
                last = atomicAdd(&A.fb_cnt[r], 1u) == (unsigned)(split - 1);
                if (last) {
                    __threadfence();
                    A.fb_cnt[r] = 0u;                               // ready for the next call
                    m0 = m1 = ~0ull;
                    for (int c = 0; c < 2 * split; ++c) {
                        const unsigned long long key = *reinterpret_cast<volatile unsigned long long *>(&A.fb_part[(size_t)r * split * 2 + c]);
                        m1 = min_u64(m1, max_u64(m0, key)); m0 = min_u64(m0, key);
                    }
                }
            }
            if (last) {
                uint4 *o = reinterpret_cast<uint4 *>(A.out + (size_t)i * 2);
                o[0] = l2fb_record(i + A.q_index_base, m0);
                o[1] = l2fb_record(i + A.q_index_base, m1);
            }
        }
    }
}
#endif
