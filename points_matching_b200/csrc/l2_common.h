// l2_common.h -- types shared by the L2 kernels (K1 pack, K2 tensor-core GEMM, K3 finish).
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#include <cuda_bf16.h>
#endif

// Packed operand row: [hi bf16 x128 | lo bf16 x128] = 512 B (K padded to 128).
#define L2_PACK_COLS 256
#define L2_KDIM 128
// exact-integer mode: t + L2_EXACT_BIAS lies in [2^23, 2^24) -> unit spacing, so the low 24
// bits of the float are (0x400000 + t) and order like the integers themselves
#define L2_EXACT_BIAS 12582912.0f          /* 1.5 * 2^23 */
/* ||b||^2 of pad columns: float bits 0x7EFFFFFF (1.7e38).  Absorbs any bias, is larger than every real
 * value in split mode, and its exact-mode key (bits * 256) is 0xFFFFFF00 = "absent" */
#define L2_PAD_NORM_BITS 0x7EFFFFFFu
#define L2_EXT_BYTES 4096                  /* norm operand image per 128-column tile */
#define L2_EXACT_NORM_LIMIT_BITS 0x4A800000u /* 2^22 as float bits: max ||.||^2 for exact mode */

struct L2Cand {        // one candidate: approximate (||b||^2 - 2ab) and train index
    float d;
    int idx;
};

struct L2Flags {
    int nonexact;              // !=0: some value is not an integer in [0,255]
    unsigned max_tnorm_bits;   // max ||b||^2 over real train rows (float bits)
    unsigned max_qnorm_bits;   // max ||a||^2 over real query rows (float bits)
    int n_flagged;             // rows K3 could not certify -> exact fallback
    unsigned rows_done;        // K3 (split mode): row blocks that have finished classifying their rows
    unsigned rows_started;     // K3 (split mode): row blocks that have started
    int next_item;             // work queue of the fallback scan (l2_fallback.cuh)
    int pad[1];
};

// exact-integer mode: integer-valued data and norms small enough for the biased key
static __host__ __device__ __forceinline__ bool l2_exact_mode(const L2Flags &f)
{
    return !f.nonexact && f.max_tnorm_bits < L2_EXACT_NORM_LIMIT_BITS && f.max_qnorm_bits < L2_EXACT_NORM_LIMIT_BITS;
}
#ifdef __CUDACC__
// split mode: shift added to every column norm so that (||b||^2 - 2ab + shift) > 0 and its
// float bits order like unsigned integers
static __device__ __forceinline__ float l2_split_shift(unsigned max_qnorm_bits)
{
    return __uint_as_float(max_qnorm_bits) * 1.001f + 1e-30f;
}
// "ext" (norm step) operand layout, K-major, no swizzle: byte offset of row r's first 16 B;
// the second K half (elements 8..15, all zero) sits 128 B further
static __device__ __forceinline__ uint32_t ext_row_offset(int r) { return (uint32_t)((r >> 3) * 256 + (r & 7) * 16); }
// exact three-term bf16 split of an fp32 value (24-bit mantissa = 3 x 8 bits): .x = h | m << 16, .y = l
static __device__ __forceinline__ uint4 bf16_split3(float v)
{
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);
    const __nv_bfloat16 l = __float2bfloat16_rn(r2);
    uint4 o;
    o.x = (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(m) << 16);
    o.y = (uint32_t)__bfloat16_as_ushort(l);
    o.z = 0u; o.w = 0u;
    return o;
}
#endif

#ifdef __CUDACC__
#ifndef L2_INF
#define L2_INF __int_as_float(0x7f800000)
#endif
// ---- helpers shared by the finish kernels (l2.cu, hamming.cu) ----
// order-preserving float -> uint map (handles negatives; t = ||b||^2 - 2ab can be < 0)
static __device__ __forceinline__ unsigned f2ord(float f)
{
    const unsigned b = __float_as_uint(f);
    return b ^ ((unsigned)((int)b >> 31) | 0x80000000u);
}
static __device__ __forceinline__ float ord2f(unsigned u)
{
    const unsigned b = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u;
    return __uint_as_float(b);
}
static __device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long k)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(0xffffffffu, k, o);
        k = y < k ? y : k;
    }
    return k;
}

// group (8 aligned lanes) minimum of 64-bit keys.  Full-mask xor butterflies stay inside the group; a
// redux.sync with a sub-warp mask is serialised per group by the compiler (4 passes + a convergence loop).
static __device__ __forceinline__ unsigned long long group_min_u64(unsigned, unsigned long long k)
{
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(0xffffffffu, k, o);
        k = y < k ? y : k;
    }
    return k;
}
// Each of the 8 lanes of a group holds partial sums v[0..7]; returns sum over the group's lanes of v[sub]
// (lane `sub` ends up with candidate `sub`): a transposing reduction, 4 + 2 + 1 shuffles instead of 8 x 3.
static __device__ __forceinline__ unsigned group_transpose_sum(unsigned (&v)[8], int sub)
{
    unsigned w[4], u[2];
    const bool b4 = (sub & 4) != 0, b2 = (sub & 2) != 0, b1 = (sub & 1) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const unsigned keep = b4 ? v[i + 4] : v[i], send = b4 ? v[i] : v[i + 4];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const unsigned keep = b2 ? w[i + 2] : w[i], send = b2 ? w[i] : w[i + 2];
        u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const unsigned keep = b1 ? u[1] : u[0], send = b1 ? u[0] : u[1];
    return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}
static __device__ __forceinline__ unsigned long long cand_key(const L2Cand c, int nt)
{
    const bool ok = c.idx >= 0 && c.idx < nt;                 // absent, or a pad column / pad quad
    return ok ? (((unsigned long long)f2ord(c.d) << 32) | (unsigned)c.idx) : ~0ull;
}
static __device__ __forceinline__ unsigned long long min_u64(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static __device__ __forceinline__ unsigned long long max_u64(unsigned long long a, unsigned long long b) { return a < b ? b : a; }

#endif

struct pm_ctx;
float *pm_l2_dump_ptr();      // debug: K2 dumps its (||b||^2 - 2ab) tile values here when set
int l2_tc_grid(pm_ctx *ctx, int MT, int NT);
int l2_tc_smax(pm_ctx *ctx, int MT, int NT);
int l2_tc_launch(pm_ctx *ctx, const void *qpack, int mq_pad, const void *tpack, int nt_pad,
                 const void *text, const L2Flags *flags, L2Cand *part, int smax, float *dump, int fp8,
                 int tmap_set = 0, const unsigned long long *chain_done = nullptr, unsigned long long wait_seq = 0,
                 unsigned long long *chain_mark = nullptr, unsigned long long mark_seq = 0);
// This call's flags block (zero on entry), the block its finish kernel must zero (it serves the call after
// the next: three blocks rotate so that a pipelined chain never touches a block the previous chain still
// uses) and the train-side block of the chunked host path.
int l2_flags_acquire(pm_ctx *ctx, L2Flags **cur, L2Flags **zero_next, L2Flags **tflags, bool advance = true);
