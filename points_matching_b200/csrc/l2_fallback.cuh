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
    unsigned long long *fb_part;  // [flagged rows][train chunks][2]: per-chunk (best, second) keys
    unsigned *fb_cnt;             // [flagged rows] chunks done; returns to zero
    pm_dmatch *out;               // [nq][2]
    int workers;                  // blocks expected to take part (sizes the split of a row into segments)
};

// at most this many helper blocks
constexpr int L2FB_MAX_GRID = 1024;
// Scratch: one (best, second) slot per (flagged row, segment) and one countdown per flagged row.  flagged rows <= 16 *
// batches, segments <= workers / batches + 1  =>  slots <= 16 * (workers + batches), batches <= nq / 16 + 1.
static inline size_t l2_fb_slots(int nq) { return (size_t)16 * ((size_t)L2FB_MAX_GRID + (size_t)(nq > 0 ? nq : 1) / 16 + 2); }
static inline size_t l2_fb_scratch_bytes(int nq) { return l2_fb_slots(nq) * 16 + (size_t)(nq > 0 ? nq : 1) * 4 + 256; }

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

__device__ __forceinline__ float4 l2fb_load_quad(const float *row, int c4) { return __ldg(reinterpret_cast<const float4 *>(row) + c4); }
__device__ __forceinline__ float4 l2fb_load_quad(const uint8_t *row, int c4)
{
    const uchar4 x = __ldg(reinterpret_cast<const uchar4 *>(row) + c4);
    return make_float4((float)x.x, (float)x.y, (float)x.z, (float)x.w);
}
__device__ __forceinline__ uint4 l2fb_record(int qidx, unsigned long long w)
{
    return w == ~0ull ? make_uint4((unsigned)qidx, 0xFFFFFFFFu, 0u, __float_as_uint(3.402823466e+38f))
                      : make_uint4((unsigned)qidx, (unsigned)(w & 0xFFFFFFFFull), 0u,
                                   __float_as_uint(sqrtf(__uint_as_float((unsigned)(w >> 32)))));
}

// Takes items from the queue until it is empty (whole block of 256 threads).  n_flagged must be final: call it only once
// every row block is known to have finished.
//
// Item = (batch of up to L2FB_ROWS flagged query rows, chunk of L2FB_CHUNK train rows): the train chunk is staged in shared
// memory ONCE and every query row of the batch is evaluated against it, so the train set crosses L2 -> SM once per batch
// of rows instead of once per row (the first version re-read the whole 5 MB train set for each of ~18 flagged rows: 92 MB
// from L2, ~22 us at cfg2 size).  8 lanes per train row, two train rows per group, the group's 16-float slices of them in
// registers; the query rows come from shared memory (the four groups of a warp read the same addresses: broadcasts).
// Per (row, chunk) the block's top-2 goes to a scratch slot; the block that completes a row's last chunk merges the row.
constexpr int L2FB_ROWS = 16;        // query rows per batch
constexpr int L2FB_CHUNK = 64;       // train rows per item

template <typename T>
__device__ __noinline__ void l2_fallback_items(const L2FallbackArgs &A)
{
    __shared__ __align__(16) float x_t[L2FB_CHUNK][L2_KDIM];        // 32 KB
    __shared__ __align__(16) float x_q[L2FB_ROWS][L2_KDIM];         // 8 KB
    __shared__ unsigned long long x_k[L2FB_ROWS][8][2];
    __shared__ int x_item, x_last[L2FB_ROWS], x_idx[L2FB_ROWS];
    const int nf = *reinterpret_cast<volatile int *>(&A.flags->n_flagged);
    if (nf <= 0) return;
    const T *q = reinterpret_cast<const T *>(A.q), *t = reinterpret_cast<const T *>(A.t);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7, grp = threadIdx.x >> 3;   // 32 groups
    const int nchunk = (A.nt + L2FB_CHUNK - 1) / L2FB_CHUNK;
    const int nbatch = (nf + L2FB_ROWS - 1) / L2FB_ROWS;
    const int rpb = (nf + nbatch - 1) / nbatch;           // rows per batch, balanced (18 rows: 9 + 9, not 16 + 2)
    // few batches: split the train chunks of a batch into `split` segments so that every worker has an item
    const int workers = A.workers > 0 ? A.workers : 1;
    int split = min(nchunk, (workers + nbatch - 1) / nbatch);
    const int seg_chunks = (nchunk + split - 1) / split;
    split = (nchunk + seg_chunks - 1) / seg_chunks;
    const int items = nbatch * split;                     // item = batch * split + segment
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) x_item = atomicAdd(&A.flags->next_item, 1);
        __syncthreads();
        const int item = x_item;
        if (item >= items) break;
        const int bt = item / split, sg = item - bt * split;
        const int r0 = bt * rpb, nr = min(rpb, nf - r0);
        // stage the query rows of the batch (zero past dim).  The row numbers first, once per item: they were written by
        // other blocks of this launch, so they are read past L1 (ld.cg) -- as plain loads the compiler may keep in flight
        // together (one volatile load per ELEMENT, each followed by its dependent row load, made an item ~10 us)
        if ((int)threadIdx.x < L2FB_ROWS) x_idx[threadIdx.x] = (int)threadIdx.x < nr ? __ldcg(&A.flagged[r0 + (int)threadIdx.x]) : 0;
        __syncthreads();
        if (A.vec) {
            float4 v[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int idx = (int)threadIdx.x + 256 * i, r = idx >> 5, c4 = idx & 31;      // 16 rows x 32 quads
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < nr) v[i] = l2fb_load_quad(q + (size_t)x_idx[r] * L2_KDIM, c4);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int idx = (int)threadIdx.x + 256 * i, r = idx >> 5, c4 = idx & 31;
                *reinterpret_cast<float4 *>(&x_q[r][4 * c4]) = v[i];
            }
        } else {
#pragma unroll 4
            for (int e = threadIdx.x; e < L2FB_ROWS * L2_KDIM; e += blockDim.x) {
                const int r = e >> 7, k = e & (L2_KDIM - 1);
                float v = 0.f;
                if (r < nr && k < A.dim) v = (float)q[(size_t)x_idx[r] * A.dim + k];
                x_q[r][k] = v;
            }
        }
        unsigned long long m0 = ~0ull, m1 = ~0ull;          // thread r < nr: running top-2 of row r0 + r over this segment
        const int ch_end = min(nchunk, (sg + 1) * seg_chunks);
        for (int ch = sg * seg_chunks; ch < ch_end; ++ch) {
        const int j0 = ch * L2FB_CHUNK;
        __syncthreads();                                    // the previous chunk's x_t / x_k have been consumed
        if (A.vec) {
            // dim == 128, aligned rows: eight independent 16-byte (f32) / 4-byte (u8) loads per thread in flight, then the stores
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int idx = (int)threadIdx.x + 256 * i, r = idx >> 5, c4 = idx & 31;      // 64 rows x 32 quads
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j0 + r < A.nt) v[i] = l2fb_load_quad(t + (size_t)(j0 + r) * L2_KDIM, c4);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int idx = (int)threadIdx.x + 256 * i, r = idx >> 5, c4 = idx & 31;
                *reinterpret_cast<float4 *>(&x_t[r][4 * c4]) = v[i];
            }
        } else {
            for (int e = threadIdx.x; e < L2FB_CHUNK * L2_KDIM; e += blockDim.x) {
                const int r = e >> 7, k = e & (L2_KDIM - 1);
                x_t[r][k] = (j0 + r < A.nt && k < A.dim) ? (float)t[(size_t)(j0 + r) * A.dim + k] : 0.f;
            }
        }
        __syncthreads();
        // this group's two train rows, their slices in registers
        float b0[4][4], b1[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float4 u = *reinterpret_cast<const float4 *>(&x_t[grp][4 * (sub + 8 * e)]);
            const float4 v = *reinterpret_cast<const float4 *>(&x_t[grp + 32][4 * (sub + 8 * e)]);
            b0[e][0] = u.x; b0[e][1] = u.y; b0[e][2] = u.z; b0[e][3] = u.w;
            b1[e][0] = v.x; b1[e][1] = v.y; b1[e][2] = v.z; b1[e][3] = v.w;
        }
        const int ja = j0 + grp, jb = j0 + grp + 32;
#pragma unroll 1
        for (int r = 0; r < nr; ++r) {
            float a[4][4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 u = *reinterpret_cast<const float4 *>(&x_q[r][4 * (sub + 8 * e)]);
                a[e][0] = u.x; a[e][1] = u.y; a[e][2] = u.z; a[e][3] = u.w;
            }
            const float da = l2fb_group_l2sq(a, b0, sub, A.dim), db = l2fb_group_l2sq(a, b1, sub, A.dim);
            const unsigned long long ka = ja < A.nt ? (((unsigned long long)__float_as_uint(da) << 32) | (unsigned)ja) : ~0ull;
            const unsigned long long kb = jb < A.nt ? (((unsigned long long)__float_as_uint(db) << 32) | (unsigned)jb) : ~0ull;
            unsigned long long k0 = min_u64(ka, kb), k1 = max_u64(ka, kb);
            // merge the 4 groups of the warp (lanes of a group hold identical keys)
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
                const unsigned long long y0 = __shfl_xor_sync(0xffffffffu, k0, o), y1 = __shfl_xor_sync(0xffffffffu, k1, o);
                const unsigned long long lo = min_u64(k0, y0), hi = max_u64(k0, y0);
                k1 = min_u64(min_u64(k1, y1), hi); k0 = lo;
            }
            if (lane == 0) { x_k[r][warp][0] = k0; x_k[r][warp][1] = k1; }
        }
        __syncthreads();
        if ((int)threadIdx.x < nr) {        // thread r folds the 8 warps' pairs of row r into its running top-2
            const int r = threadIdx.x;
            for (int w = 0; w < 8; ++w)
                for (int e = 0; e < 2; ++e) { const unsigned long long key = x_k[r][w][e]; m1 = min_u64(m1, max_u64(m0, key)); m0 = min_u64(m0, key); }
        }
        }   // chunks of the segment
        // the segment's top-2 of row r -> its scratch slot; whoever completes a row's LAST segment merges the row: thread r
        // runs the countdown, then the warps of the block merge the completed rows (lanes load the slots in parallel --
        // one thread reading ~100 slots one volatile load after the other took 75 us)
        __syncthreads();
        if ((int)threadIdx.x < L2FB_ROWS) x_last[threadIdx.x] = 0;
        __syncthreads();
        if ((int)threadIdx.x < nr) {
            const int fr = r0 + (int)threadIdx.x;
            unsigned long long *slot = A.fb_part + ((size_t)fr * split + sg) * 2;
            slot[0] = m0; slot[1] = m1;
            __threadfence();
            if (atomicAdd(&A.fb_cnt[fr], 1u) == (unsigned)(split - 1)) {
                __threadfence();
                A.fb_cnt[fr] = 0u;                                   // ready for the next call
                x_last[threadIdx.x] = 1;
            }
        }
        __syncthreads();
        for (int r = warp; r < nr; r += 8) {
            if (!x_last[r]) continue;
            const int fr = r0 + r;
            unsigned long long k0 = ~0ull, k1 = ~0ull;
            const unsigned long long *slots = A.fb_part + (size_t)fr * split * 2;       // other blocks wrote them: ld.cg
            for (int c = lane; c < 2 * split; c += 128) {                                // four loads in flight per lane
                unsigned long long key[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) key[u] = c + 32 * u < 2 * split ? __ldcg(slots + c + 32 * u) : ~0ull;
#pragma unroll
                for (int u = 0; u < 4; ++u) { k1 = min_u64(k1, max_u64(k0, key[u])); k0 = min_u64(k0, key[u]); }
            }
#pragma unroll
            for (int o = 1; o <= 16; o <<= 1) {
                const unsigned long long y0 = __shfl_xor_sync(0xffffffffu, k0, o), y1 = __shfl_xor_sync(0xffffffffu, k1, o);
                const unsigned long long lo = min_u64(k0, y0), hi = max_u64(k0, y0);
                k1 = min_u64(min_u64(k1, y1), hi); k0 = lo;
            }
            if (lane < 2) {
                const int i = x_idx[r];
                reinterpret_cast<uint4 *>(A.out + (size_t)i * 2)[lane] = l2fb_record(i + A.q_index_base, lane == 0 ? k0 : k1);
            }
        }
    }
}
#endif
