// l2.cu -- float L2 kNN-2: K1 (pack + squared-norm pre-pass), K3 (merge, FP32 re-rank,
// certification), the exact FP32 kernel used for uncertified rows / wide descriptors,
// and the orchestration around the tensor-core kernel K2 (l2_tc.cu).
//
// Replaces BruteForceMatcher<L2<float>>::match / knnMatch (k = 2), the matcher named at
// /root/reference/Points Matching/main.cpp:43 and called at main.cpp:46; results follow
// OpenCV's batchDistance semantics: distance = sqrt(sum (a-b)^2) in f32, rows sorted
// ascending, ties -> lowest trainIdx.
//
// "re-rank order" (DESIGN.md): the FP32 squared distance is defined as
//   lane l (0..31): p_l = fma-chain over k = 128c + 4l + e (c ascending, e = 0..3) of (a_k-b_k)^2
//   then the xor-butterfly p += shfl_xor(p, 16|8|4|2|1)
// which the oracle restates bit for bit (orc_l2sq_f32_rerank).
#include <atomic>
#include <cuda_bf16.h>
#include "pm_internal.h"
#include "l2_common.h"
#include "l2_fallback.cuh"

namespace {

constexpr float L2_EPS_REL = 6.2e-5f;   // bound on |approx - exact| / (||a|| ||b||max), split mode
#define L2_INF __int_as_float(0x7f800000)

__device__ __forceinline__ float warp_sum_butterfly(float p)
{
    p += __shfl_xor_sync(0xffffffffu, p, 16);
    p += __shfl_xor_sync(0xffffffffu, p, 8);
    p += __shfl_xor_sync(0xffffffffu, p, 4);
    p += __shfl_xor_sync(0xffffffffu, p, 2);
    p += __shfl_xor_sync(0xffffffffu, p, 1);
    return p;
}

template <typename T> __device__ __forceinline__ float load_elem(const T *p, size_t i) { return (float)p[i]; }

// ---------------------------------------------------------------------------------
// K1: 8 lanes per row, query and train rows in ONE launch (rows [0, mq_pad) are query
// rows, the rest train rows).  Writes [hi|lo] bf16 (train rows pre-scaled by -2),
// ||row||^2 (train pad rows: 1.7e38 inside K2's norm image), the integer-valued flag, the max norms, and (query
// side) the +inf candidate init.  Pure streaming: 16 B loads, 8 B stores per lane.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void load_row4(const float *p, int lane, int dim, bool vec, float (&x)[4])
{
    if (vec) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p) + lane);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) x[e] = 4 * lane + e < dim ? __ldg(p + 4 * lane + e) : 0.f;
    }
}
__device__ __forceinline__ void load_row4(const uint8_t *p, int lane, int dim, bool vec, float (&x)[4])
{
    if (vec) {
        const uchar4 v = __ldg(reinterpret_cast<const uchar4 *>(p) + lane);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) x[e] = 4 * lane + e < dim ? (float)__ldg(p + 4 * lane + e) : 0.f;
    }
}

// chunk e of a 128-wide row for lane `sub` of an 8-lane group: elements 4 * (sub + 8e) .. + 3
__device__ __forceinline__ void load_chunk4(const float *p, int l, int dim, bool vec, float (&x)[4])
{
    if (vec) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p) + l);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = 4 * l + c < dim ? __ldg(p + 4 * l + c) : 0.f;
    }
}
__device__ __forceinline__ void load_chunk4(const uint8_t *p, int l, int dim, bool vec, float (&x)[4])
{
    if (vec) {
        const uchar4 v = __ldg(reinterpret_cast<const uchar4 *>(p) + l);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = 4 * l + c < dim ? (float)__ldg(p + 4 * l + c) : 0.f;
    }
}

// K1: 8 lanes per row (4 rows per warp: the per-row bookkeeping -- reduction, norm image, flags -- is
// amortised over four rows; one warp per row was instruction bound at ~230 warp instructions per row).
template <typename T>
__global__ void __launch_bounds__(256)
l2_pack_kernel(const T *__restrict__ q, int nq, int mq_pad, const T *__restrict__ t, int nt, int nt_pad, int dim,
               int vec, __nv_bfloat16 *__restrict__ qpack, __nv_bfloat16 *__restrict__ tpack,
               float *__restrict__ qnorm, uint8_t *__restrict__ text, uint8_t *__restrict__ q8, uint8_t *__restrict__ t8,
               float *__restrict__ tnorm, L2Flags *flags,
               L2Cand *__restrict__ part, int part_per_row, const L2Flags *tflags_in, unsigned long long *span,
               const unsigned long long *chain_done, unsigned long long wait_seq)
{
    __shared__ unsigned s_max[2][8];
    __shared__ int s_nonint;
    pm_span_mark(span, 0, false);
    if (chain_done) {
        // pipelined chain (pm_set_pipelining): this launch does NOT wait for its stream predecessors -- it
        // overlaps the finish / filter kernels of the previous chain.  It reads only the caller's inputs and
        // writes only this chain's buffer set and flags block, which the previous chain does not touch; it
        // waits for "K2 of the previous chain is past its waits" = everything before that chain is complete.
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        pm_chain_wait(chain_done, wait_seq);
        __syncthreads();
    } else {
        pm_pdl_prologue();
    }
    pm_span_mark(span, 1, false);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7;
    if (threadIdx.x == 0) s_nonint = 0;
    unsigned mx_q = 0u, mx_t = 0u;          // running max of the norm bits (norms are >= 0: bits order like floats)
    bool integral = true;
    // mq_pad and nt_pad are multiples of 4, so the four rows of a warp are all query rows or all train rows
    const int total = mq_pad + nt_pad, stride = gridDim.x * (blockDim.x >> 3);
    for (int grow = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); grow < total; grow += stride) {
        const bool is_train = grow >= mq_pad;
        const int row = is_train ? grow - mq_pad : grow;
        const int n = is_train ? nt : nq;
        const bool live = row < n;
        const T *src = (is_train ? t : q) + (size_t)(live ? row : 0) * dim;
        float x[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            load_chunk4(src, sub + 8 * e, dim, vec != 0, x[e]);
            if (!live) { x[e][0] = x[e][1] = x[e][2] = x[e][3] = 0.f; }
        }
        const float scale = is_train ? -2.f : 1.f;
        __nv_bfloat16 *dst = (is_train ? tpack : qpack) + (size_t)row * L2_PACK_COLS;
        unsigned *dst8 = reinterpret_cast<unsigned *>((is_train ? t8 : q8) + (size_t)row * L2_KDIM);
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            // byte copy for K3's exact-mode re-check; the row is "integral" (0..255 integers) iff the saturating
            // conversion round-trips
            unsigned ub[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                s = fmaf(x[e][c], x[e][c], s);
                ub[c] = __float2uint_rn(fminf(fmaxf(x[e][c], 0.f), 255.f));
                integral = integral && (float)ub[c] == x[e][c];
            }
            // [hi | lo] bf16 split, two elements per conversion instruction
            const float v0 = x[e][0] * scale, v1 = x[e][1] * scale, v2 = x[e][2] * scale, v3 = x[e][3] * scale;
            const __nv_bfloat162 h01 = __floats2bfloat162_rn(v0, v1), h23 = __floats2bfloat162_rn(v2, v3);
            const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
            const __nv_bfloat162 l01 = __floats2bfloat162_rn(v0 - f01.x, v1 - f01.y), l23 = __floats2bfloat162_rn(v2 - f23.x, v3 - f23.y);
            const int l = sub + 8 * e;
            *reinterpret_cast<uint2 *>(dst + 4 * l) =
                make_uint2(*reinterpret_cast<const unsigned *>(&h01), *reinterpret_cast<const unsigned *>(&h23));
            *reinterpret_cast<uint2 *>(dst + L2_KDIM + 4 * l) =
                make_uint2(*reinterpret_cast<const unsigned *>(&l01), *reinterpret_cast<const unsigned *>(&l23));
            dst8[l] = ub[0] | (ub[1] << 8) | (ub[2] << 16) | (ub[3] << 24);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (is_train) {
            if (sub == 2) tnorm[row] = live ? s : 0.f;
            // K2's norm operand: [n_h n_m n_l 1 1 1 0 0 | 0 x 8] bf16 in the smem image of the row's column tile
            if (sub < 2) {
                const uint4 s3 = bf16_split3(live ? s : __uint_as_float(L2_PAD_NORM_BITS));
                uint8_t *e = text + (size_t)(row >> 7) * L2_EXT_BYTES + ext_row_offset(row & 127) + sub * 128;
                *reinterpret_cast<uint4 *>(e) = sub == 0 ? make_uint4(s3.x, s3.y | 0x3F800000u, 0x3F803F80u, 0u)
                                                         : make_uint4(0u, 0u, 0u, 0u);
            }
            if (live) mx_t = max(mx_t, __float_as_uint(s));
        } else {
            if (sub == 0) qnorm[row] = live ? s : 0.f;
            if (live) mx_q = max(mx_q, __float_as_uint(s));
            for (int k = sub; k < part_per_row; k += 8) part[(size_t)row * part_per_row + k] = L2Cand{L2_INF, -1};
        }
    }
    // one atomic per block and side (same-address traffic serialises in L2)
    if (!__all_sync(0xffffffffu, integral) && lane == 0) s_nonint = 1;
    if (lane == 0) { s_max[0][warp] = mx_q; s_max[1][warp] = mx_t; }
    __syncthreads();
    if (threadIdx.x < 2) {
        unsigned mx = 0u;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mx = max(mx, s_max[threadIdx.x][w]);
        if (mx) atomicMax(threadIdx.x == 0 ? &flags->max_qnorm_bits : &flags->max_tnorm_bits, mx);
    }
    if (threadIdx.x == 2 && s_nonint) flags->nonexact = 1;
    // query-only launch of the chunked host path: fold in the train side's flags (packed earlier)
    if (tflags_in && blockIdx.x == 0 && threadIdx.x == 3) {
        if (tflags_in->nonexact) flags->nonexact = 1;
        atomicMax(&flags->max_tnorm_bits, tflags_in->max_tnorm_bits);
    }
    pm_span_mark(span, 2, true);
}

// exact FP32 squared distance between the query chunk held in registers (dim <= 128) and a train row
template <typename T>
__device__ __forceinline__ float warp_l2sq_regs(const float (&a)[4], const T *__restrict__ b, int dim, int lane)
{
    float p = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (4 * lane + e < dim) { const float d = a[e] - load_elem(b, 4 * lane + e); p = fmaf(d, d, p); }
    return warp_sum_butterfly(p);
}

// ---------------------------------------------------------------------------------
// K3: one warp per query row.  K2 hands over, per segment, the best adjacent-column
// PAIR minima (value, index).  K3 merges them, re-computes the members of the winning
// pairs in FP32 (the "re-rank"), and
//   exact mode: the answer is exact -- the overall second best is either the second
//               pair minimum or the partner (index ^ 1) of the best;
//   split mode: both members of the three best pairs are re-ranked; every column outside
//               them is >= the third pair minimum, which certifies the top-2 (else the
//               row goes to the exact kernel).
// ---------------------------------------------------------------------------------
// Exact FP32 kNN-2 of ONE query row by a whole CTA (256 threads): warps stride over the train rows,
// lexicographic (d^2, index) top-2.  qs[dim_pad], md[16], mi[16] are shared-memory scratch.
template <typename T>
__device__ void l2_exact_row(const T *__restrict__ q, const T *__restrict__ t, int nt, int dim, int i,
                             int q_index_base, pm_dmatch *__restrict__ out, float *qs, float *md, int *mi)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int dim_pad = (dim + 127) / 128 * 128;
    __syncthreads();
    for (int k = threadIdx.x; k < dim_pad; k += blockDim.x) qs[k] = k < dim ? load_elem(q, (size_t)i * dim + k) : 0.f;
    __syncthreads();
    float b0 = L2_INF, b1 = L2_INF; int i0 = -1, i1 = -1;
    for (int j = warp; j < nt; j += nwarps) {
        const T *b = t + (size_t)j * dim;
        float p = 0.f;
        for (int c = 0; c < dim_pad; c += 128) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int kk = c + 4 * lane + e;
                if (kk < dim) { const float d = qs[kk] - load_elem(b, kk); p = fmaf(d, d, p); }
            }
        }
        p = warp_sum_butterfly(p);
        if (p < b0) { b1 = b0; i1 = i0; b0 = p; i0 = j; }      // j ascending within a warp
        else if (p < b1) { b1 = p; i1 = j; }
    }
    if (lane == 0) { md[2 * warp] = b0; md[2 * warp + 1] = b1; mi[2 * warp] = i0; mi[2 * warp + 1] = i1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float d0 = L2_INF, d1 = L2_INF; int j0 = -1, j1 = -1;
        for (int w = 0; w < 2 * nwarps; ++w) {
            const float d = md[w]; const int j = mi[w];
            if (j < 0) continue;
            if (d < d0 || (d == d0 && (unsigned)j < (unsigned)j0)) { d1 = d0; j1 = j0; d0 = d; j0 = j; }
            else if (d < d1 || (d == d1 && (unsigned)j < (unsigned)j1)) { d1 = d; j1 = j; }
        }
        pm_dmatch r0, r1;
        r0.queryIdx = r1.queryIdx = i + q_index_base;
        r0.imgIdx = r1.imgIdx = 0;
        r0.trainIdx = j0; r0.distance = j0 < 0 ? 3.402823466e+38f : sqrtf(d0);
        r1.trainIdx = j1; r1.distance = j1 < 0 ? 3.402823466e+38f : sqrtf(d1);
        out[(size_t)i * 2] = r0;
        out[(size_t)i * 2 + 1] = r1;
    }
}

// ---- K3 works in groups of 8 lanes per query row (4 rows per warp) ----
// A 128-wide row held by 8 lanes: lane s keeps the float4 chunks s, s+8, s+16, s+24, i.e. the elements
// that "virtual lanes" l = s + 8e of the 32-lane re-rank order own.
__device__ __forceinline__ void load_row8(const float *p, int sub, int dim, bool vec, float (&v)[4][4])
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
__device__ __forceinline__ void load_row8(const uint8_t *p, int sub, int dim, bool vec, float (&v)[4][4])
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
// sum (a-b)^2 in the re-rank order: virtual-lane partials (fma chain over e = 0..3), then the xor
// butterfly 16 | 8 (inside the lane) and 4 | 2 | 1 (across the group's lanes)
__device__ __forceinline__ float group_l2sq(unsigned gmask, const float (&a)[4][4], const float (&b)[4][4], int sub, int dim)
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

template <typename T>
__global__ void __launch_bounds__(256, 3)
l2_finish_kernel(const L2Cand *__restrict__ part, int ncand, const float *__restrict__ qnorm,
                 const uint8_t *__restrict__ q8, const uint8_t *__restrict__ t8, const float *__restrict__ tnorm,
                 const T *__restrict__ q, const T *__restrict__ t, int nq, int nt, int dim, int vec,
                 L2Flags *flags, L2Flags *flags_next, int *__restrict__ flagged,
                 int q_index_base, pm_dmatch *__restrict__ out, unsigned long long *span, int row_blocks, L2FallbackArgs fb)
{
    // Blocks [0, row_blocks) classify the rows; blocks past them are HELPERS of the split-mode fallback scan (see the tail
    // of this kernel).  A row block announces itself before the wait below, so that a helper -- which looks after its
    // own wait -- sees every row block that was resident by then.
    __shared__ int s_role;
    const bool is_helper = (int)blockIdx.x >= row_blocks;
    if (!is_helper && threadIdx.x == 0) atomicAdd(&flags->rows_started, 1u);
    pm_span_mark(span, 6, false);
    pm_pdl_prologue();
    pm_span_mark(span, 7, false);
    if (is_helper) {
        if (l2_exact_mode(*flags)) return;                 // exact-integer data: no row is ever flagged
        if (threadIdx.x == 0) {
            // stay only if EVERY row block has started (then waiting for them cannot starve them); else the last row
            // block to finish, which takes the same work queue, completes the scan without this helper.  The row blocks'
            // announcements may still be in flight when a helper dispatched right behind them looks, so it looks for a
            // bounded time (~50 us at most, then it leaves: a bounded wait cannot deadlock anything)
            unsigned v = 0;
            for (int spin = 0; spin < 512; ++spin) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&flags->rows_started) : "memory");
                if (v >= (unsigned)row_blocks) break;
                __nanosleep(100);
            }
            int stay = v >= (unsigned)row_blocks;
            if (stay) {
                // every row block is running: they finish in a few microseconds.  (Bounded all the same -- ~0.3 s -- so
                // that no fault elsewhere can turn this wait into a hang; a helper that gives up leaves the scan to the
                // last row block.)
                int spin = 0;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&flags->rows_done) : "memory");
                    if (v < (unsigned)row_blocks) __nanosleep(64);
                } while (v < (unsigned)row_blocks && ++spin < (1 << 22));
                stay = v >= (unsigned)row_blocks;
            }
            s_role = stay;
        }
        __syncthreads();
        if (span && threadIdx.x == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); span[32 + 2 * blockIdx.x] = tt; }
        if (s_role) l2_fallback_items<T>(fb);
        if (span && threadIdx.x == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); span[33 + 2 * blockIdx.x] = s_role ? tt : 0ull; }
        return;
    }
    if (span && threadIdx.x == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); span[32 + 2 * blockIdx.x] = tt; }
    const int lane = threadIdx.x & 31, sub = lane & 7;
    const unsigned gmask = 0xFFu << (lane & 24);
    if (blockIdx.x == 0 && threadIdx.x == 0) *flags_next = L2Flags{0, 0u, 0u, 0, 0u, 0u, 0, {0}};   // the next call's block
    const int ngroups = row_blocks * (blockDim.x >> 3);
    const int nq_round = (nq + 3) & ~3;              // whole warps stay in the loop together (group shuffles)
    for (int i = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); i < nq_round; i += ngroups) {
        const bool live_row = i < nq;
        const int ir = live_row ? i : nq - 1;
        // every load that does not depend on the candidates goes out first (two candidates per lane up front)
        const L2Cand *prow = part + (size_t)ir * ncand;
        L2Cand ca = L2Cand{L2_INF, -1}, cb = L2Cand{L2_INF, -1};
        if (sub < ncand) ca = prow[sub];
        if (sub + 8 < ncand) cb = prow[sub + 8];
        const float na = qnorm[ir];
        const L2Flags fl = *flags;
        const bool split = !l2_exact_mode(fl);
        const uint4 aq = __ldg(reinterpret_cast<const uint4 *>(q8 + (size_t)ir * L2_KDIM) + sub);   // byte row, 16 per lane
        // lane-local sorted triple (branch-free for the first two), then three rounds of group arg-min over
        // the lane heads (keys are unique: distinct train indices)
        const unsigned long long ka = cand_key(ca, nt), kb = cand_key(cb, nt);
        unsigned long long h0 = min_u64(ka, kb), h1 = max_u64(ka, kb), h2 = ~0ull;
        for (int c0 = sub + 16; c0 < ncand; c0 += 8) {       // more than 16 candidates: many CTAs share the row tile
            const unsigned long long key = cand_key(prow[c0], nt);
            h2 = min_u64(h2, max_u64(h1, key));
            h1 = min_u64(max_u64(h0, key), h1);
            h0 = min_u64(h0, key);
        }
        unsigned long long k[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            k[r] = group_min_u64(gmask, h0);
            const bool pop = h0 == k[r] && h0 != ~0ull;
            h0 = pop ? h1 : h0; h1 = pop ? h2 : h1; h2 = pop ? ~0ull : h2;
        }
        float b0 = L2_INF, b1 = L2_INF; int j0 = -1, j1 = -1;
        bool certified = true;
        if (!split) {
            // exact mode: K2 hands over the two best column QUADS (index = first column of the quad, value =
            // exact minimum over its four columns).  All eight members are recomputed exactly in integers from
            // the byte copies K1 keeps: d^2 = ||a||^2 + ||b||^2 - 2 a.b with a.b by dp4a (everything < 2^24, so
            // the float results equal the re-rank-order distances bit for bit).
            const int cand = sub >> 2, mem = sub & 3;                 // lane `sub` owns candidate (quad cand, member mem)
            const bool live0 = k[0] != ~0ull, live1 = k[1] != ~0ull;
            const int jq0 = live0 ? (int)(k[0] & 0xFFFFFFFFu) : 0, jq1 = live1 ? (int)(k[1] & 0xFFFFFFFFu) : 0;
            const int my_col = (cand ? jq1 : jq0) + mem;
            const float my_nb = tnorm[min(my_col, nt - 1)];
            uint4 bq[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int col = (c >> 2 ? jq1 : jq0) + (c & 3);
                bq[c] = __ldg(reinterpret_cast<const uint4 *>(t8 + (size_t)min(col, nt - 1) * L2_KDIM) + sub);
            }
            unsigned dots[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                unsigned dot = __dp4a(aq.x, bq[c].x, 0u);
                dot = __dp4a(aq.y, bq[c].y, dot);
                dot = __dp4a(aq.z, bq[c].z, dot);
                dots[c] = __dp4a(aq.w, bq[c].w, dot);
            }
            const unsigned my_dot = group_transpose_sum(dots, sub);
            const bool mine = (cand ? live1 : live0) && my_col < nt;
            const float my_d = (float)((int)na + (int)my_nb - 2 * (int)my_dot);        // >= 0, exact
            // two rounds of group arg-min over (d^2, column)
            unsigned long long key = mine ? (((unsigned long long)__float_as_uint(my_d) << 32) | (unsigned)my_col) : ~0ull;
            const unsigned long long w0 = group_min_u64(gmask, key);
            key = key == w0 ? ~0ull : key;
            const unsigned long long w1 = group_min_u64(gmask, key);
            if (w0 != ~0ull) { j0 = (int)(w0 & 0xFFFFFFFFu); b0 = __uint_as_float((unsigned)(w0 >> 32)); }
            if (w1 != ~0ull) { j1 = (int)(w1 & 0xFFFFFFFFu); b1 = __uint_as_float((unsigned)(w1 >> 32)); }
        } else {
            // split mode: both members of the three best pairs are re-ranked in FP32 (direct differences)
            float a[4][4];
            load_row8(q + (size_t)ir * dim, sub, dim, vec != 0, a);
            float d2[6]; int idx[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) { d2[r] = L2_INF; idx[r] = -1; }
            const float bound = k[2] == ~0ull ? L2_INF : ord2f((unsigned)(k[2] >> 32)) + na;   // approx d^2 of the 3rd pair minimum
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const bool live = k[r] != ~0ull;
                const int j = live ? (int)(k[r] & 0xFFFFFFFFu) : 0, pj = j ^ 1;
                float b[4][4];
                load_row8(t + (size_t)j * dim, sub, dim, vec != 0, b);
                const float dj = group_l2sq(gmask, a, b, sub, dim);
                load_row8(t + (size_t)(pj < nt ? pj : j) * dim, sub, dim, vec != 0, b);
                const float dpj = group_l2sq(gmask, a, b, sub, dim);
                if (live) { idx[2 * r] = j; d2[2 * r] = dj; if (pj < nt) { idx[2 * r + 1] = pj; d2[2 * r + 1] = dpj; } }
            }
            // the two smallest by (d^2, index)
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                if (idx[r] < 0) continue;
                const float d = d2[r]; const int j = idx[r];
                if (d < b0 || (d == b0 && (unsigned)j < (unsigned)j0)) { b1 = b0; j1 = j0; b0 = d; j0 = j; }
                else if (d < b1 || (d == b1 && (unsigned)j < (unsigned)j1)) { b1 = d; j1 = j; }
            }
            if (k[2] != ~0ull) {
                const float eps = L2_EPS_REL * sqrtf(na * __uint_as_float(fl.max_tnorm_bits));
                certified = b1 < bound - eps;
            }
        }
        if (live_row && sub < 2) {
            // lanes 0 and 1 of the group write the two 16-byte DMatch records of the row
            const int jj = sub == 0 ? j0 : j1;
            const float dd = sub == 0 ? b0 : b1;
            reinterpret_cast<uint4 *>(out + (size_t)i * 2)[sub] =
                make_uint4((unsigned)(i + q_index_base), (unsigned)jj, 0u,
                           __float_as_uint(jj < 0 ? 3.402823466e+38f : sqrtf(fmaxf(dd, 0.f))));
            if (sub == 0 && !certified) flagged[atomicAdd(&flags->n_flagged, 1)] = i;
        }
    }
    // Split mode only: rows that could not be certified (flagged[0 .. n_flagged)) get an exact FP32 scan of the whole
    // train set (l2_fallback.cuh).  The list is complete once every row block has passed the counter below; the helper
    // blocks wait for exactly that, and the last row block to arrive takes part as well, so the scan completes whatever
    // the helpers did.  Nobody waits for a block that may not have started.  (An earlier version ran the scan behind a
    // hand-rolled grid barrier, which assumed that the whole grid is resident at once -- not guaranteed beside other
    // streams, e.g. the lanes of the batched pair call.)
    if (!l2_exact_mode(*flags)) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            int role = atomicAdd(&flags->rows_done, 1u) == (unsigned)(row_blocks - 1);     // the last one: the list is complete
            if (!role) {
                // not the last: help all the same, IF every row block is known to have started (then waiting for the rest
                // of them cannot starve anybody) -- with the whole grid resident, which is the normal case, the scan runs
                // on every SM instead of on the few that host the helper blocks
                unsigned v;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&flags->rows_started) : "memory");
                if (v >= (unsigned)row_blocks) {
                    int spin = 0;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&flags->rows_done) : "memory");
                        if (v < (unsigned)row_blocks) __nanosleep(64);
                    } while (v < (unsigned)row_blocks && ++spin < (1 << 22));
                    role = v >= (unsigned)row_blocks;
                }
            } else {
                __threadfence();
            }
            s_role = role;
        }
        __syncthreads();
        if (s_role) l2_fallback_items<T>(fb);
    }
    if (span && threadIdx.x == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); span[33 + 2 * blockIdx.x] = tt; }
    pm_span_mark(span, 8, true);
}

// ---------------------------------------------------------------------------------
// Exact FP32 kNN-2 for a list of rows (or all rows): one CTA per query row, warps
// stride over the train rows, lexicographic (d^2, index) top-2.
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
l2_exact_kernel(const T *__restrict__ q, const T *__restrict__ t, int nq, int nt, int dim,
                int q_index_base, pm_dmatch *__restrict__ out, unsigned long long *span)
{
    extern __shared__ float qs[];                 // [dim_pad] query row, then merge scratch
    pm_span_mark(span, 9, false);
    pm_pdl_prologue();
    pm_span_mark(span, 10, false);
    const int dim_pad = (dim + 127) / 128 * 128;
    float *md = qs + dim_pad;                     // [nwarps][2]
    int *mi = reinterpret_cast<int *>(md + 2 * (blockDim.x >> 5));
    for (int r = blockIdx.x; r < nq; r += gridDim.x) l2_exact_row(q, t, nt, dim, r, q_index_base, out, qs, md, mi);
    pm_span_mark(span, 11, true);
}

__global__ void l2_knn_to_colbest_kernel(const pm_dmatch *__restrict__ knn, int n, int base,
                                         unsigned long long *__restrict__ col)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const pm_dmatch m = knn[(size_t)j * 2];
    col[j] = m.trainIdx < 0 ? ~0ull
             : ((unsigned long long)__float_as_uint(m.distance) << 32) | (unsigned)(m.trainIdx + base);
}

template <typename T>
int run_exact(pm_ctx *ctx, const T *dq, const T *dt, int nq, int nt, int dim, int base, pm_dmatch *dout)
{
    const int dim_pad = (dim + 127) / 128 * 128;
    const size_t smem = (size_t)dim_pad * 4 + 8 * 4 * 4;
    const int grid = min(nq, 8 * ctx->num_sms);
    PM_CUDA(ctx, pm_launch_pdl(l2_exact_kernel<T>, dim3(grid), dim3(256), smem, ctx->stream, dq, dt, nq, nt, dim,
                               base, dout, g_pm_span));
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

}  // namespace

// Debug hook (tools/gpu_debug.py): dump of (||b||^2 - 2ab) from the tensor-core kernel.
static float *g_l2_dump = nullptr;
extern "C" void pm_debug_set_l2_dump(float *ddump) { g_l2_dump = ddump; }
float *pm_l2_dump_ptr() { return g_l2_dump; }
// Force the exact FP32 kernel for every row (parity cross-check of the two paths).
static int g_l2_force_exact = 0;
extern "C" void pm_debug_force_exact(int on) { g_l2_force_exact = on; }
// A/B switch: no helper blocks -- the last row block of K3 runs the whole fallback scan alone (the path taken when the
// helpers find that not every row block is resident)
static int g_l2_fb_no_helpers = 0;
extern "C" void pm_debug_fallback_no_helpers(int on) { g_l2_fb_no_helpers = on; }

int l2_flags_acquire(pm_ctx *ctx, L2Flags **cur, L2Flags **zero_next, L2Flags **tflags, bool advance)
{
    const bool fresh = ctx->slot_ptr[WS_L2_FLAGS] == nullptr;
    PM_WS(ctx, f, L2Flags *, WS_L2_FLAGS, 4 * sizeof(L2Flags) + 64);
    if (fresh) { PM_CUDA(ctx, cudaMemsetAsync(f, 0, 4 * sizeof(L2Flags) + 64, ctx->stream)); ctx->l2_rot = 0; }
    *cur = f + ctx->l2_rot;
    *zero_next = f + (ctx->l2_rot + 2) % 3;
    *tflags = f + 3;
    if (advance) ctx->l2_rot = (ctx->l2_rot + 1) % 3;      // only calls whose finish kernel zeroes *zero_next may advance
    return PM_OK;
}

// One chain K1 -> K2 -> K3 (-> K5 when dgood is given).  phase 0: pack query + train, match (one call).  The
// chunked host path (pm_api.cu) overlaps the H2D copies with compute: phase 1 packs the train set only
// (flags -> the train-side block), phase 2 runs one query chunk against the train set packed by phase 1.
//
// Chain pipelining (opt-in, pm_set_pipelining; one-call kNN-2 + ratio chains only): every buffer K1 writes
// exists twice and chains alternate between the sets, so K1 of chain s+1 may start while K3 / K5 of chain s
// still run -- it skips griddepcontrol.wait when the previous kernel on the stream is the tail of chain s
// with the same shapes.  Ordering is then carried by two device words: K2 of every signalling chain stores
// its number to chain_mark once past its waits (= K1 of that chain and everything enqueued before the chain
// have completed), the tail K5 stores it to chain_done.  K1(s+1) spins for chain_mark >= s, K2(s+1) spins
// for chain_done >= s after its own griddepcontrol.wait, and K3 / K5 follow K2 by PDL.
// Measured on the B200: grid completion is transitive along a PDL chain -- K2(s+1)'s griddepcontrol.wait is not
// released when K1(s+1) exits but when K5(s), K1's own stream predecessor, has completed (a dependent that
// never executes griddepcontrol.wait gets it at its exit).  So K2(s+1) cannot be started under K3(s) / K5(s)
// on one stream (tried: K3 carrying the chain wait and a 128-thread K5 that fits beside a K2 CTA only made
// K5 slower, 36.1 vs 33.6 us per step); what the mode buys is K1's 4 us.  (K3 applying the ratio test itself was tried as well: with 313 32-row tiles the
// ordered prefix inside K3 cost 8-10 us against 3.8 us for K5 behind one PDL boundary.)
static int l2_chain(pm_ctx *ctx, const void *dq, int nq, const void *dt, int nt, int dim, int is_u8,
                    int q_index_base, pm_dmatch *dout, int phase, float ratio, pm_dmatch *dgood, int32_t *dn_good,
                    const pm_gather_out *gather = nullptr)
{
    if (nq <= 0 && phase != 1) {
        if (dgood) PM_CUDA(ctx, cudaMemsetAsync(dn_good, 0, sizeof(int32_t), ctx->stream));
        return PM_OK;
    }
    // K2 work items are 256 query rows x 128 train columns
    const int mq_pad = phase == 1 ? 0 : pm_round_up(nq, 256), nt_pad = pm_round_up(nt > 0 ? nt : 1, 256);
    const int MT = mq_pad / 256, NT = nt_pad / 128;
    const int smax = phase == 1 ? 1 : l2_tc_smax(ctx, MT, NT);
    const bool use_tc = dim <= L2_KDIM && nt > 0 && !g_l2_force_exact;
    ctx->l2_stats[2] = 0; ctx->l2_stats[3] = 0;
    if (!use_tc) {
        if (phase == 1) return PM_OK;
        int st = is_u8 ? run_exact(ctx, (const uint8_t *)dq, (const uint8_t *)dt, nq, nt, dim, q_index_base, dout)
                       : run_exact(ctx, (const float *)dq, (const float *)dt, nq, nt, dim, q_index_base, dout);
        if (st != PM_OK || !dgood) return st;
        return pmk_ratio_filter(ctx, dout, nq, ratio, dgood, dn_good, gather);
    }
    // ---- chain pipelining state ----
    const bool signalling = ctx->pipelining && phase == 0 && dgood != nullptr;     // this chain's tail stores its number
    const bool same_shape = ctx->chain_shape[0] == nq && ctx->chain_shape[1] == nt && ctx->chain_shape[2] == dim &&
                            ctx->chain_shape[3] == is_u8;
    const bool run_ahead = signalling && ctx->tail_is_chain && same_shape;          // K1 skips griddepcontrol.wait
    const unsigned long long seq = signalling ? ctx->chain_seq + 1 : 0;
    const int set = signalling ? (int)(seq & 1) : 0;                                // buffer set of this chain
    const int nset = ctx->pipelining ? 2 : 1;
    L2Flags *flags, *flags_next, *tflags;
    { int fst = l2_flags_acquire(ctx, &flags, &flags_next, &tflags, phase != 1); if (fst != PM_OK) return fst; }
    unsigned long long *chain_done = reinterpret_cast<unsigned long long *>(tflags + 1);   // [0] done, [1] mark, [2] block counter
    unsigned long long *chain_mark = chain_done + 1;
    unsigned *chain_ctr = reinterpret_cast<unsigned *>(chain_done + 2);
    // every workspace K1 writes, sized for `nset` copies; `set` selects one
    auto ws2 = [&](int slot, size_t bytes) -> uint8_t * {
        const size_t each = (bytes + 255) & ~(size_t)255;
        uint8_t *base = (uint8_t *)pm_ws(ctx, slot, each * nset);
        return base ? base + each * set : nullptr;
    };
#define L2_WS2(var, type, slot, bytes) type var = (type)ws2(slot, bytes); if (!var) return PM_CUDA_ERR
    L2_WS2(tpack, __nv_bfloat16 *, WS_T_PACK, (size_t)nt_pad * L2_PACK_COLS * 2);
    L2_WS2(text, uint8_t *, WS_T_NORM, (size_t)(nt_pad / 128) * L2_EXT_BYTES);
    L2_WS2(t8, uint8_t *, WS_T_U8, (size_t)nt_pad * L2_KDIM);
    L2_WS2(tnormf, float *, WS_T_NORMF, (size_t)nt_pad * 4);
    const int vec_u8 = dim == L2_KDIM && (((uintptr_t)dq | (uintptr_t)dt) & 3) == 0;
    const int vec_f32 = dim == L2_KDIM && (((uintptr_t)dq | (uintptr_t)dt) & 15) == 0;
    const int vec = is_u8 ? vec_u8 : vec_f32;
    const unsigned long long *no_chain = nullptr;
    if (phase == 1) {
        PM_CUDA(ctx, cudaMemsetAsync(tflags, 0, sizeof(L2Flags), ctx->stream));
        const int blocks = min(pm_cdiv(nt_pad, 32), 8 * ctx->num_sms);
        if (is_u8)
            PM_CUDA(ctx, pm_launch_pdl(l2_pack_kernel<uint8_t>, dim3(blocks), dim3(256), 0, ctx->stream, (const uint8_t *)nullptr, 0, 0,
                                       (const uint8_t *)dt, nt, nt_pad, dim, vec, (__nv_bfloat16 *)nullptr, tpack, (float *)nullptr, text,
                                       (uint8_t *)nullptr, t8, tnormf, tflags, (L2Cand *)nullptr, 0, (const L2Flags *)nullptr, g_pm_span,
                                       no_chain, 0ull));
        else
            PM_CUDA(ctx, pm_launch_pdl(l2_pack_kernel<float>, dim3(blocks), dim3(256), 0, ctx->stream, (const float *)nullptr, 0, 0,
                                       (const float *)dt, nt, nt_pad, dim, vec, (__nv_bfloat16 *)nullptr, tpack, (float *)nullptr, text,
                                       (uint8_t *)nullptr, t8, tnormf, tflags, (L2Cand *)nullptr, 0, (const L2Flags *)nullptr, g_pm_span,
                                       no_chain, 0ull));
        PM_CHECK_LAUNCH(ctx);
        return PM_OK;
    }
    L2_WS2(qpack, __nv_bfloat16 *, WS_Q_PACK, (size_t)mq_pad * L2_PACK_COLS * 2);
    L2_WS2(qnorm, float *, WS_Q_NORM, (size_t)mq_pad * 4);
    L2_WS2(q8, uint8_t *, WS_Q_U8, (size_t)mq_pad * L2_KDIM);
    L2_WS2(part, L2Cand *, WS_L2_PART, (size_t)mq_pad * smax * 3 * sizeof(L2Cand));
    L2_WS2(flagged, int *, WS_L2_FLAGGED, (size_t)nq * 4);
    // fallback scratch (l2_fallback.cuh): per-(row, segment) key pairs, then the per-row countdowns (zero between calls)
    const size_t fb_bytes = l2_fb_scratch_bytes(nq);
    const bool fb_fresh = ctx->slot_bytes[WS_L2_FBPART] < (size_t)nset * ((fb_bytes + 255) & ~(size_t)255);
    L2_WS2(fbpart, unsigned long long *, WS_L2_FBPART, fb_bytes);
    if (fb_fresh) PM_CUDA(ctx, cudaMemsetAsync(ctx->slot_ptr[WS_L2_FBPART], 0, ctx->slot_bytes[WS_L2_FBPART], ctx->stream));
#undef L2_WS2
    L2FallbackArgs fb;
    fb.q = dq; fb.t = dt; fb.is_u8 = is_u8; fb.nq = nq; fb.nt = nt; fb.dim = dim; fb.vec = vec; fb.q_index_base = q_index_base;
    fb.flags = flags; fb.flagged = flagged; fb.fb_part = fbpart;
    fb.fb_cnt = reinterpret_cast<unsigned *>(fbpart + 2 * l2_fb_slots(nq)); fb.out = dout;
    // K3: 8 lanes per row, 32 rows per block, at most one resident wave of row blocks (a second wave would double its
    // latency), plus the helper blocks of the split-mode fallback scan (they leave at once in exact-integer mode)
    const int fin_blocks = min(pm_cdiv(nq, 32), 3 * ctx->num_sms);
    const int fin_helpers = g_l2_fb_no_helpers ? 0 : min(max(3 * ctx->num_sms - fin_blocks, ctx->num_sms), L2FB_MAX_GRID - 1);
    fb.workers = fin_helpers + fin_blocks;
    const int pack_nt = phase == 2 ? 0 : nt, pack_nt_pad = phase == 2 ? 0 : nt_pad;
    const L2Flags *tflags_in = phase == 2 ? tflags : nullptr;
    const int pack_blocks = min(pm_cdiv(mq_pad + pack_nt_pad, 32), 8 * ctx->num_sms);
    const unsigned long long *k1_done = run_ahead ? chain_mark : nullptr;
    const unsigned long long k1_wait = run_ahead ? seq - 1 : 0;
    if (is_u8)
        PM_CUDA(ctx, pm_launch_pdl(l2_pack_kernel<uint8_t>, dim3(pack_blocks), dim3(256), 0, ctx->stream, (const uint8_t *)dq, nq, mq_pad,
                                   (const uint8_t *)dt, pack_nt, pack_nt_pad, dim, vec, qpack, tpack, qnorm, text, q8, t8, tnormf, flags, part, smax * 3,
                                   tflags_in, g_pm_span, k1_done, k1_wait));
    else
        PM_CUDA(ctx, pm_launch_pdl(l2_pack_kernel<float>, dim3(pack_blocks), dim3(256), 0, ctx->stream, (const float *)dq, nq, mq_pad,
                                   (const float *)dt, pack_nt, pack_nt_pad, dim, vec, qpack, tpack, qnorm, text, q8, t8, tnormf, flags, part, smax * 3,
                                   tflags_in, g_pm_span, k1_done, k1_wait));
    PM_CHECK_LAUNCH(ctx);
    int st = l2_tc_launch(ctx, qpack, mq_pad, tpack, nt_pad, text, flags, part, smax, g_l2_dump, 0, set,
                          run_ahead ? chain_done : nullptr, run_ahead ? seq - 1 : 0, signalling ? chain_mark : nullptr, seq);
    if (st != PM_OK) return st;
    {
        // K3 carries 42 KB of static shared memory (the fallback scan's staging): ask for the full carveout so that three
        // blocks still fit on an SM (the default carveout admitted one, which tripled the row phase)
        static std::atomic<unsigned long long> k3_attr_devices{0};
        const unsigned long long dev_bit = 1ull << (ctx->device & 63);
        if (!(k3_attr_devices.load(std::memory_order_acquire) & dev_bit)) {
            PM_CUDA(ctx, cudaFuncSetAttribute(l2_finish_kernel<uint8_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            PM_CUDA(ctx, cudaFuncSetAttribute(l2_finish_kernel<float>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            k3_attr_devices.fetch_or(dev_bit, std::memory_order_release);
        }
    }
    if (is_u8)
        PM_CUDA(ctx, pm_launch_pdl(l2_finish_kernel<uint8_t>, dim3(fin_blocks + fin_helpers), dim3(256), 0, ctx->stream, (const L2Cand *)part, smax * 3,
                                   (const float *)qnorm, (const uint8_t *)q8, (const uint8_t *)t8, (const float *)tnormf,
                                   (const uint8_t *)dq, (const uint8_t *)dt, nq, nt, dim, vec_u8, flags, flags_next, flagged,
                                   q_index_base, dout, g_pm_span, fin_blocks, fb));
    else
        PM_CUDA(ctx, pm_launch_pdl(l2_finish_kernel<float>, dim3(fin_blocks + fin_helpers), dim3(256), 0, ctx->stream, (const L2Cand *)part, smax * 3,
                                   (const float *)qnorm, (const uint8_t *)q8, (const uint8_t *)t8, (const float *)tnormf,
                                   (const float *)dq, (const float *)dt, nq, nt, dim, vec_f32, flags, flags_next, flagged,
                                   q_index_base, dout, g_pm_span, fin_blocks, fb));
    PM_CHECK_LAUNCH(ctx);
    ctx->l2_stats[3] = smax;
    if (!dgood) return PM_OK;
    if (!signalling) return pmk_ratio_filter(ctx, dout, nq, ratio, dgood, dn_good, gather);
    st = pmk_ratio_filter_tail(ctx, dout, nq, ratio, dgood, dn_good, chain_done, chain_ctr, seq, gather);
    if (st != PM_OK) return st;
    ctx->chain_seq = seq;
    ctx->tail_is_chain = true;              // cleared by the next launch of any other kind (PM_CHECK_LAUNCH)
    ctx->chain_shape[0] = nq; ctx->chain_shape[1] = nt; ctx->chain_shape[2] = dim; ctx->chain_shape[3] = is_u8;
    return PM_OK;
}

int pmk_l2_knn2_fused(pm_ctx *ctx, const void *dq, int nq, const void *dt, int nt, int dim, int is_u8,
                      int q_index_base, pm_dmatch *dout, int phase, float ratio, pm_dmatch *dgood, int32_t *dn_good,
                      const pm_gather_out *gather)
{
    return l2_chain(ctx, dq, nq, dt, nt, dim, is_u8, q_index_base, dout, phase, ratio, dgood, dn_good, gather);
}

int pmk_l2_knn2_phase(pm_ctx *ctx, const void *dq, int nq, const void *dt, int nt, int dim, int is_u8,
                      int q_index_base, pm_dmatch *dout, int phase)
{
    return l2_chain(ctx, dq, nq, dt, nt, dim, is_u8, q_index_base, dout, phase, 0.f, nullptr, nullptr);
}

int pmk_l2_knn2(pm_ctx *ctx, const void *dq, int nq, const void *dt, int nt, int dim, int is_u8,
                int q_index_base, pm_dmatch *dout)
{
    return pmk_l2_knn2_phase(ctx, dq, nq, dt, nt, dim, is_u8, q_index_base, dout, 0);
}

int pm_l2_stats(pm_ctx *ctx, int32_t out[4])
{
    if (!ctx) return PM_BAD_ARG;
    L2Flags h = {};
    if (ctx->slot_ptr[WS_L2_FLAGS]) {     // the block the last call used (parity was flipped after it)
        PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PM_CUDA(ctx, cudaMemcpy(&h, (const L2Flags *)ctx->slot_ptr[WS_L2_FLAGS] + (ctx->l2_rot + 2) % 3, sizeof(h),
                                cudaMemcpyDeviceToHost));
    }
    const bool exact = l2_exact_mode(h);
    out[0] = ctx->l2_stats[3] ? exact : 0;
    out[1] = h.n_flagged;
    out[2] = ctx->l2_stats[3] ? (exact ? 2 : 6) : 0;
    out[3] = ctx->l2_stats[3];
    return PM_OK;
}

// Nearest query for every train row = kNN of the train set against the query shard.
int pmk_l2_col_best(pm_ctx *ctx, const float *dq, int nq, const float *dt, int nt, int dim, int q_index_base,
                    uint64_t *dcol_best)
{
    if (nt <= 0) return PM_OK;
    PM_WS(ctx, knn, pm_dmatch *, WS_KNN2, (size_t)nt * 2 * sizeof(pm_dmatch));
    if (nq <= 0) {
        PM_CUDA(ctx, cudaMemsetAsync(dcol_best, 0xFF, (size_t)nt * 8, ctx->stream));
        return PM_OK;
    }
    int st = pmk_l2_knn2(ctx, dt, nt, dq, nq, dim, 0, 0, knn);
    if (st != PM_OK) return st;
    // trainIdx here is a (local) query index: shift by the shard base
    l2_knn_to_colbest_kernel<<<pm_cdiv(nt, 256), 256, 0, ctx->stream>>>(knn, nt, q_index_base,
                                                                        (unsigned long long *)dcol_best);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

int pm_hang_init_l2(pm_hang_rec *dev_view)
{
    return cudaMemcpyToSymbol(g_pm_hang_rec, &dev_view, sizeof(dev_view)) == cudaSuccess ? PM_OK : PM_CUDA_ERR;
}
