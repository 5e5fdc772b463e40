// l2_tc.cu -- K2: L2 distance GEMM on tcgen05 tensor cores with the top-k selection fused
// into the epilogue (the distance matrix never reaches HBM).  sm_100a only.
//
// Replaces the all-pairs distance + per-row k-smallest inside
// BruteForceMatcher<L2<float>>::match / knnMatch, the matcher named at
// /root/reference/Points Matching/main.cpp:43 (OpenCV batchDistance).
//
//   d^2(i,j) = ||a_i||^2 + ||b_j||^2 - 2 a_i.b_j
// The -2 is folded into the packed train operand, ||a_i||^2 is row-constant and added
// after selection, and ||b_j||^2 (plus the key bias) rides in the GEMM itself: one extra
// K=16 MMA step multiplies a constant [1 1 1 0...] row block by the column norm split
// exactly into three bf16 terms, so the epilogue does no arithmetic on the distance at all.
// Operands are bf16 "hi|lo" rows written by K1 (l2.cu):
//   exact-integer mode (SIFT, values 0..255): lo == 0, 2 k-blocks, the GEMM is exact
//   split mode (general floats):  a.b ~ ah.bh + ah.bl + al.bh, 6 k-blocks
//
// Work item = 256 query rows x 128 train columns: TWO 128-row A tiles stay resident in shared
// memory and every B k-block that TMA brings in feeds both, which halves the L2 -> SM operand
// traffic (the binding resource for K = 128: ncu measured 4.4 TB/s of TMA reads with one A tile).
//
// Structure (one persistent CTA per SM, 576 threads):
//   warp 0      TMA producer   : 2 A row-tiles resident (2 x <= 4 x 16 KB), B k-blocks through
//                                a 4-stage 16 KB ring (cp.async.bulk.tensor, SWIZZLE_128B)
//   warp 1      MMA issuer     : tcgen05.mma cta_group::1 kind::f16, M=128 N=128 K=16, one
//                                accumulator per A tile, double-buffered in TMEM (2 x 2 x 128
//                                columns); also allocates / frees TMEM
//   warps 2-17  epilogue       : 4 warps per scheduler, each owns 32 rows x 64 columns of a
//                                work item: tcgen05.ld 32x32b.x32, pack (value | column) into one
//                                u32 key (IMAD / PRMT), branch-free min/max
//                                tournament (VIMNMX/VIMNMX3) over adjacent column PAIRS: top-2
//                                pair minima in exact mode, top-3 in split mode, kept in
//                                registers across the sweep (K3 re-checks the partners exactly).
// Each CTA walks a contiguous range of the (row-tile, column-tile) space; per row tile it
// writes one "segment" of candidates which K3 merges, re-ranks in FP32 and certifies.
#include <cuda.h>
#include <cstdlib>
#include <cuda_bf16.h>
#include "pm_internal.h"
#include "l2_common.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int MH = 2;                        // A row-tiles (m-halves) per work item: 256 rows
constexpr int A_BLK_BYTES = BM * BK * 2;     // 16 KB
constexpr int B_BLK_BYTES = BN * BK * 2;     // 16 KB
constexpr int NSTAGE = 4;
constexpr int A_MAXBLK = 4;
constexpr int EPI_WARP0 = 2;        // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-17 epilogue
constexpr int EPI_WARPS = 16;       // 4 per scheduler: enough TLP to keep the issue slots busy
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int TC_THREADS = (EPI_WARP0 + EPI_WARPS) * 32;   // 576
constexpr int NSLICE = EPI_WARPS / 4 / MH;                 // column slices per accumulator: 2
constexpr int SLICE = BN / NSLICE;                         // 64 columns per warp
constexpr int SMEM_A = 0;
constexpr int SMEM_B = MH * A_MAXBLK * A_BLK_BYTES;            // 131072
constexpr int SMEM_BAR = SMEM_B + NSTAGE * B_BLK_BYTES;        // 196608
constexpr int SMEM_SCRATCH = SMEM_BAR + 256;                   // MH x (NSLICE-1) x 128 rows x 3 candidates
// "ext" operands of the norm MMA step, K-major, no swizzle: core matrix = 8 rows x 16 B,
// the two 8-element K halves 128 B apart (LBO), 8-row groups 256 B apart (SBO)
constexpr int EXT_A_BYTES = BM * 32;                           // 4 KB, constant [1 1 1 0...] rows
constexpr int EXT_B_BYTES = BN * 32;                           // 8 KB per tile, double buffered
constexpr int SMEM_EXTA = SMEM_SCRATCH + MH * (NSLICE - 1) * BM * 3 * 8;
constexpr int SMEM_EXTB = SMEM_EXTA + EXT_A_BYTES;
constexpr int SMEM_TOTAL = SMEM_EXTB + 2 * EXT_B_BYTES + 1024; // + alignment slack

// instruction descriptor: D=F32, A=B=BF16, K-major both, N=128, M=128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// K-major, SWIZZLE_128B smem operand: 8-row atoms of 1024 B (SBO), LBO = 1 (16 B, unused
// inside one 128 B swizzle span), descriptor version 1 (Blackwell), layout type 2.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// K-major, no swizzle (ext operands): LBO = 128 B between the K halves, SBO = 256 B between 8-row groups
__device__ __forceinline__ uint64_t make_sdesc_ext(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t ext_row_offset(int r) { return (uint32_t)((r >> 3) * 256 + (r & 7) * 16); }
// exact three-term bf16 split of an fp32 value (24-bit mantissa = 3 x 8 bits)
__device__ __forceinline__ uint4 bf16_split3(float v)
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// epilogue: branch-free selection on packed keys.
//   exact mode: acc + (||b||^2 + 1.5*2^23) is an integer-valued float in [2^23, 2^24), so
//               bits * 256 + (c+1) = ((0x400000 + t) << 8) | (c+1) orders by (t, column).
//   split mode: acc + (||b||^2 + shift) > 0, key = (bits & ~0xFF) | (c+1) (PRMT): the value is
//               truncated by <= 2^-15 relative, which only lowers the certification bound.
// The low byte is the column within the warp's 64-column slice, plus one (1..64): non-zero
// for keys of the current tile and zeroed on carried keys, so an equal distance from an
// earlier tile always wins (lowest train index) and "which entries are new" needs no compare.
// The 32 keys of a chunk are 16 adjacent (even, odd) column pairs; only the smaller key of a
// pair competes: the overall runner-up is either another pair's minimum or the partner
// (column ^ 1) of a winner, which K3 re-checks exactly.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t umin3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u32(a, b, c); }

struct Sel2 {   // exact mode: running (best, second) pair minima
    uint32_t m1, m2; int i1, i2;
    __device__ __forceinline__ void reset() { m1 = m2 = 0xFFFFFFFFu; i1 = i2 = -1; }
    // 56 min/max ops per 32 elements; chb = column base of the chunk inside the slice
    __device__ __forceinline__ void chunk(uint32_t (&k)[32], uint32_t chb)
    {
        uint32_t lo[8], hi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t p0 = min(k[4 * i], k[4 * i + 1]), p1 = min(k[4 * i + 2], k[4 * i + 3]);
            lo[i] = min(p0, p1); hi[i] = max(p0, p1);
        }
#pragma unroll
        for (int w = 4; w >= 1; w >>= 1)
#pragma unroll
            for (int i = 0; i < w; ++i) {
                const uint32_t L = min(lo[i], lo[i + w]);
                const uint32_t S = umin3(max(lo[i], lo[i + w]), hi[i], hi[i + w]);
                lo[i] = L; hi[i] = S;
            }
        // keys carry the column within the chunk (1..32); make it the column within the slice
        const uint32_t L = lo[0] + chb, S = hi[0] + chb;
        m2 = umin3(m2, S, max(m1, L));
        m1 = min(m1, L);
    }
    // after a tile: resolve the indices of entries that came from it, zero their column byte
    __device__ __forceinline__ void end_tile(int col0)
    {
        const uint32_t n1 = m1 & 0xFFu, n2 = m2 & 0xFFu;
        const int g1 = col0 + (int)n1 - 1, g2 = col0 + (int)n2 - 1;
        i2 = n2 ? g2 : (n1 ? i1 : i2);      // a carried second is the old best when a new best arrived
        i1 = n1 ? g1 : i1;
        m1 &= ~0xFFu; m2 &= ~0xFFu;
    }
    __device__ __forceinline__ float value(uint32_t m) const { return (float)((int)(m >> 8) - 0x400000); }
};

struct Sel3 {   // split mode: running (best, second, third) pair minima
    uint32_t m1, m2, m3; int i1, i2, i3;
    __device__ __forceinline__ void reset() { m1 = m2 = m3 = 0xFFFFFFFFu; i1 = i2 = i3 = -1; }
    __device__ __forceinline__ void chunk(uint32_t (&k)[32], uint32_t chb)
    {
        uint32_t a1 = 0xFFFFFFFFu, a2 = 0xFFFFFFFFu, a3 = 0xFFFFFFFFu;     // chunk-local sorted triple
#pragma unroll
        for (int i = 0; i < 8; ++i) {       // two pair minima (sorted) at a time into the triple
            const uint32_t p0 = min(k[4 * i], k[4 * i + 1]), p1 = min(k[4 * i + 2], k[4 * i + 3]);
            const uint32_t lo = min(p0, p1), hi = max(p0, p1);
            const uint32_t c3 = umin3(max(a1, hi), max(a2, lo), a3);
            const uint32_t c2 = umin3(hi, max(a1, lo), a2);
            a1 = min(a1, lo); a2 = c2; a3 = c3;
        }
        a1 += chb; a2 += chb; a3 += chb;    // 16 distinct pair minima per chunk: all three are real
        // merge two sorted triples
        const uint32_t c1 = min(m1, a1);
        const uint32_t c2 = umin3(a2, max(m1, a1), m2);
        const uint32_t c3 = min(umin3(a3, max(m1, a2), max(m2, a1)), m3);
        m1 = c1; m2 = c2; m3 = c3;
    }
    __device__ __forceinline__ void end_tile(int col0)
    {
        const uint32_t n1 = m1 & 0xFFu, n2 = m2 & 0xFFu, n3 = m3 & 0xFFu;
        const int g1 = col0 + (int)n1 - 1, g2 = col0 + (int)n2 - 1, g3 = col0 + (int)n3 - 1;
        const int o1 = i1, o2 = i2, o3 = i3;
        // carried entries keep their order: the p-th carried slot takes the p-th old index
        const int c1 = n1 ? 0 : 1;                   // carried entries consumed before slot 2
        const int c2 = c1 + (n2 ? 0 : 1);            // ... before slot 3
        i1 = n1 ? g1 : o1;
        i2 = n2 ? g2 : (c1 ? o2 : o1);
        i3 = n3 ? g3 : (c2 == 0 ? o1 : (c2 == 1 ? o2 : o3));
        m1 &= ~0xFFu; m2 &= ~0xFFu; m3 &= ~0xFFu;
    }
};

// lexicographic (value, index) insert used when merging the column slices at a flush
struct Cand3 {
    float d[3]; int i[3];
    __device__ __forceinline__ void reset() { d[0] = d[1] = d[2] = __int_as_float(0x7f800000); i[0] = i[1] = i[2] = -1; }
    __device__ __forceinline__ static bool less(float da, int ia, float db, int ib)
    { return da < db || (da == db && (unsigned)ia < (unsigned)ib); }
    __device__ __forceinline__ void insert(float t, int idx)
    {
        if (idx < 0) return;
        if (less(t, idx, d[2], i[2])) {
            if (less(t, idx, d[1], i[1])) {
                d[2] = d[1]; i[2] = i[1];
                if (less(t, idx, d[0], i[0])) { d[1] = d[0]; i[1] = i[0]; d[0] = t; i[0] = idx; }
                else { d[1] = t; i[1] = idx; }
            } else { d[2] = t; i[2] = idx; }
        }
    }
};

struct TcParams {
    const float *tnorm;          // [nt_pad] ||b||^2 (pad columns: +inf)
    const L2Flags *flags;
    L2Cand *part;                // [mq_pad][smax][3]; MT counts 256-row super tiles
    float *dump;                 // debug: [mq_pad][nt_pad] of (||b||^2 - 2ab), or null
    int MT, NT, smax, nt_pad;
    long long *trace;            // PM_K2_DBG & 32: clock64 stamps of CTA 0, [tile][4]
    uint32_t dbg;                // timing experiments only (PM_K2_DBG): 1 skip selection math, 2 skip main MMAs, 4 skip TMA of B
    uint32_t mul256;             // == 256, passed at run time so the key pack stays an IMAD (FMA pipe), not an ALU LEA
};

__global__ void __launch_bounds__(TC_THREADS, 1)
l2_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_t, TcParams P)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t sA = sbase + SMEM_A, sB = sbase + SMEM_B, sBar = sbase + SMEM_BAR;
    const uint32_t bar_full = sBar, bar_empty = sBar + 8 * NSTAGE;
    const uint32_t bar_afull = sBar + 16 * NSTAGE, bar_aempty = bar_afull + 8;
    const uint32_t bar_tfull = bar_aempty + 8, bar_tempty = bar_tfull + 16;
    const uint32_t bar_ext = bar_tempty + 16;                    // [2]: norm operand of a tile staged
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(sgen + SMEM_BAR + 16 * NSTAGE + 64);
    const uint32_t sExtA = sbase + SMEM_EXTA, sExtB = sbase + SMEM_EXTB;
    L2Cand *scratch = reinterpret_cast<L2Cand *>(sgen + SMEM_SCRATCH);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pm_pdl_prologue();      // K1's outputs (flags, packed operands, norms) are complete past this point
    const L2Flags fl = *P.flags;
    const bool exact = l2_exact_mode(fl);
    const int nkb = exact ? 2 : 6;
    const int nablk = exact ? 2 : 4;

    const long long T = (long long)P.MT * P.NT;
    const int G = gridDim.x;
    const int t_begin = (int)((T * blockIdx.x) / G), t_end = (int)((T * (blockIdx.x + 1)) / G);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_t)) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
            mbar_init(bar_afull, 1);
            mbar_init(bar_aempty, 1);
            for (int a = 0; a < 2; ++a) {
                mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, EPI_WARPS);
                mbar_init(bar_ext + 8 * a, BN / 32);             // one arrival per staging warp
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32((const void *)tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= EPI_WARP0 && (int)threadIdx.x - EPI_WARP0 * 32 < BM) {
        // constant A operand of the norm step: row r = [1 1 1 0 0 0 0 0 | 0 x 8] in bf16
        const int r = (int)threadIdx.x - EPI_WARP0 * 32;
        uint8_t *dst = sgen + SMEM_EXTA + ext_row_offset(r);
        *reinterpret_cast<uint4 *>(dst) = make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u);
        *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, apar = 0; int cur_m = -1;
            for (int tile = t_begin; tile < t_end; ++tile) {
                const int m = tile / P.NT, n = tile % P.NT;
                if (m != cur_m) {
                    if (cur_m >= 0) { mbar_wait(bar_aempty, apar); apar ^= 1; }
                    mbar_expect_tx(bar_afull, (uint32_t)(MH * nablk * A_BLK_BYTES));
                    for (int h = 0; h < MH; ++h)
                        for (int b = 0; b < nablk; ++b)
                            tma_load_2d(sA + (h * A_MAXBLK + b) * A_BLK_BYTES, &tmap_q, bar_afull, b * BK, (m * MH + h) * BM);
                    cur_m = m;
                }
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    if (P.dbg & 4u) { mbar_arrive(bar_full + 8 * stage); if (++stage == NSTAGE) { stage = 0; phase ^= 1; } continue; }
                    mbar_expect_tx(bar_full + 8 * stage, B_BLK_BYTES);
                    // k-blocks: hi0 hi1 | lo0 lo1 | hi0 hi1   (packed row = [hi 0..127 | lo 128..255])
                    const int kc = ((kb == 2 || kb == 3) ? 128 : 0) + (kb & 1) * BK;
                    tma_load_2d(sB + stage * B_BLK_BYTES, &tmap_t, bar_full + 8 * stage, kc, n * BN);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, apar = 0, acc = 0, acc_phase = 0; int cur_m = -1;
#define TR(slot) do { if (P.trace && blockIdx.x == 0 && tile - t_begin < 64) P.trace[(tile - t_begin) * 16 + (slot)] = clock64(); } while (0)
            for (int tile = t_begin; tile < t_end; ++tile) {
                TR(0);
                const int m = tile / P.NT;
                if (m != cur_m) { mbar_wait(bar_afull, apar); apar ^= 1; cur_m = m; }
                TR(1);
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                TR(2);
                tc_fence_after();
                TR(3);
                const uint32_t d_tmem = tmem_base + acc * (MH * BN);     // accumulator (acc, h) at + h * BN
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    if (kb == 0) TR(4); else if (kb == 1) TR(7);
                    // A block: hi for kb 0..3 (x B hi, x B lo), lo for kb 4,5 (x B hi)
                    const int ablk = kb < 4 ? (kb & 1) : 2 + (kb & 1);
                    const uint64_t bdesc = make_sdesc(sB + stage * B_BLK_BYTES);
                    if (!(P.dbg & 2u))
#pragma unroll
                    for (int h = 0; h < MH; ++h) {      // the same B k-block feeds both A tiles
                        const uint64_t adesc = make_sdesc(sA + (h * A_MAXBLK + ablk) * A_BLK_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)   // +32 B per K=16 step inside the swizzle span
                            umma_bf16(d_tmem + h * BN, adesc + 2 * k, bdesc + 2 * k, IDESC, (uint32_t)((kb | k) != 0));
                    }
                    if (kb == 0) TR(5); else if (kb == 1) TR(8);
                    umma_commit(bar_empty + 8 * stage);
                    if (kb == 0) TR(6); else if (kb == 1) TR(9);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                // norm step: acc += [1 1 1 0..] x split3(||b||^2 + bias)
                if (!(P.dbg & 8u)) {
                mbar_wait(bar_ext + 8 * acc, acc_phase);
                TR(10);
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < MH; ++h)
                    umma_bf16(d_tmem + h * BN, make_sdesc_ext(sExtA), make_sdesc_ext(sExtB + acc * EXT_B_BYTES), IDESC, 1u);
                }
                TR(11);
                umma_commit(bar_tfull + 8 * acc);
                TR(12);
                if (tile + 1 < t_end && (tile + 1) / P.NT != m) umma_commit(bar_aempty);
                acc ^= 1; if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ===================== epilogue =====================
        const int e = warp - EPI_WARP0;
        const int quarter = warp & 3;          // TMEM lane quarter this warp may access
        const int mh = (e >> 2) & (MH - 1);    // which A row-tile (accumulator) of the work item
        const int slice = e >> 3;              // which 64 of the item's 128 columns
        const int row = mh * BM + quarter * 32 + lane;   // row within the 256-row work item
        const float shift = l2_split_shift(fl.max_qnorm_bits);
        const float nb_off = exact ? L2_EXACT_BIAS : shift;            // bias folded into the norm operand
        const float nb_pad = exact ? L2_EXACT_PAD : 3.0e38f;           // pad columns: never selected
        const uint32_t mul = P.mul256;
        const int et = e * 32 + lane;                                  // threads 0..255 stage one column each
        uint32_t acc = 0, acc_phase = 0; int cur_m = -1;
        // stage_ext(buf, nb): this thread's column of the norm operand, then one arrival per warp
        auto stage_ext = [&](uint32_t buf, float nb) {
            const float v = nb == __int_as_float(0x7f800000) ? nb_pad : nb + nb_off;
            const uint4 s3 = bf16_split3(v);
            uint8_t *dst = sgen + SMEM_EXTB + buf * EXT_B_BYTES + ext_row_offset(et);
            *reinterpret_cast<uint4 *>(dst) = s3;
            *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(0u, 0u, 0u, 0u);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // visible to the tensor core's proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ext + 8 * buf);
        };
        float nb_pref = 0.f;
        if (et < BN && t_begin < t_end) {
            stage_ext(0u, __ldg(P.tnorm + (t_begin % P.NT) * BN + et));
            if (t_begin + 1 < t_end) nb_pref = __ldg(P.tnorm + ((t_begin + 1) % P.NT) * BN + et);
        }
        Sel2 s2; s2.reset();
        Sel3 s3; s3.reset();

        auto flush = [&](int m) {
            Cand3 c; c.reset();
            if (exact) {
                if ((s2.m1 >> 8) != 0xFFFFFFu) c.insert(s2.value(s2.m1), s2.i1);
                if ((s2.m2 >> 8) != 0xFFFFFFu) c.insert(s2.value(s2.m2), s2.i2);
            } else {
                if ((s3.m1 >> 8) != 0xFFFFFFu) c.insert(__uint_as_float(s3.m1) - shift, s3.i1);
                if ((s3.m2 >> 8) != 0xFFFFFFu) c.insert(__uint_as_float(s3.m2) - shift, s3.i2);
                if ((s3.m3 >> 8) != 0xFFFFFFu) c.insert(__uint_as_float(s3.m3) - shift, s3.i3);
            }
            if (slice > 0) {
#pragma unroll
                for (int k = 0; k < 3; ++k) scratch[((slice - 1) * MH * BM + row) * 3 + k] = L2Cand{c.d[k], c.i[k]};
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            if (slice == 0) {
                for (int sl = 0; sl < NSLICE - 1; ++sl)
#pragma unroll
                    for (int k = 0; k < 3; ++k) { const L2Cand o = scratch[(sl * MH * BM + row) * 3 + k]; c.insert(o.d, o.idx); }
                // segment slot = index of this CTA among the CTAs that touch row tile m
                const long long first_tile = (long long)m * P.NT;
                int c0 = (int)((first_tile * G) / T);
                while ((T * (c0 + 1)) / G <= first_tile) ++c0;
                while ((T * c0) / G > first_tile) --c0;
                const int slot = (int)blockIdx.x - c0;
                L2Cand *dst = P.part + ((size_t)(m * MH * BM + row) * P.smax + slot) * 3;
#pragma unroll
                for (int k = 0; k < 3; ++k) dst[k] = L2Cand{c.d[k], c.i[k]};
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            s2.reset(); s3.reset();
        };

        for (int tile = t_begin; tile < t_end; ++tile) {
            const int m = tile / P.NT, n = tile % P.NT;
            if (m != cur_m) { if (cur_m >= 0) flush(cur_m); cur_m = m; }
            // stage the NEXT tile's norm operand (its buffer was last read by the MMA of tile-1,
            // which this warp has already seen complete) and prefetch the one after
            if (et < BN && tile + 1 < t_end && !(P.dbg & 8u)) {
                stage_ext(acc ^ 1u, nb_pref);
                if (tile + 2 < t_end) nb_pref = __ldg(P.tnorm + ((tile + 2) % P.NT) * BN + et);
            }
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            if (P.trace && blockIdx.x == 0 && e == 0 && lane == 0 && tile - t_begin < 64) P.trace[(tile - t_begin) * 16 + 13] = clock64();
            const int col0 = n * BN + slice * SLICE;
            const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * (MH * BN) + mh * BN + slice * SLICE;
#pragma unroll 1
            for (int ch = 0; ch < SLICE / 32; ++ch) {
                uint32_t r[32];
                if (P.dbg & 16u) continue;
                tmem_ld32(taddr0 + ch * 32, r);
                tmem_ld_wait();
                if (P.dump) {
                    float *drow = P.dump + (size_t)(m * MH * BM + row) * P.nt_pad + col0 + ch * 32;
#pragma unroll
                    for (int c = 0; c < 32; ++c) drow[c] = __uint_as_float(r[c]) - nb_off;
                }
                if (P.dbg & 1u) { s2.m1 ^= r[0] ^ r[31]; continue; }
                if (exact) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) r[c] = r[c] * mul + (uint32_t)(c + 1);
                    s2.chunk(r, (uint32_t)(ch * 32));
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) r[c] = __byte_perm(r[c], (uint32_t)(c + 1), 0x3214);
                    s3.chunk(r, (uint32_t)(ch * 32));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
            if (P.trace && blockIdx.x == 0 && e == 0 && lane == 0 && tile - t_begin < 64) P.trace[(tile - t_begin) * 16 + 14] = clock64();
            if (exact) s2.end_tile(col0); else s3.end_tile(col0);
            acc ^= 1; if (acc == 0) acc_phase ^= 1;
        }
        if (cur_m >= 0) flush(cur_m);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap(pm_ctx *ctx, CUtensorMap *tm, const void *base, int rows_pad, int box_rows)
{
    if (!ctx->tmap_encode) {
        cudaDriverEntryPointQueryResult qres;
        void *fn = nullptr;
        PM_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess)
            return pm_fail(ctx, PM_CUDA_ERR, "cuTensorMapEncodeTiled entry point not found");
        ctx->tmap_encode = fn;
    }
    cuuint64_t dims[2] = {(cuuint64_t)L2_PACK_COLS, (cuuint64_t)rows_pad};
    cuuint64_t strides[1] = {(cuuint64_t)L2_PACK_COLS * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((tmap_encode_fn)ctx->tmap_encode)(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base),
                                                    dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return pm_fail(ctx, PM_CUDA_ERR, "cuTensorMapEncodeTiled failed: %d", (int)r);
    return PM_OK;
}

}  // namespace

static long long *g_k2_trace = nullptr;
extern "C" void pm_debug_set_k2_trace(long long *p) { g_k2_trace = p; }

int l2_tc_grid(pm_ctx *ctx, int MT, int NT)
{
    long long T = (long long)MT * NT;
    return (int)(T < ctx->num_sms ? T : ctx->num_sms);
}

// Max number of CTAs (segments) that can touch one row tile.
int l2_tc_smax(pm_ctx *ctx, int MT, int NT)
{
    const long long T = (long long)MT * NT;
    const int G = l2_tc_grid(ctx, MT, NT);
    int smax = 1;
    for (int m = 0; m < MT; ++m) {
        const long long first = (long long)m * NT, last = first + NT - 1;
        int c0 = (int)((first * G) / T);
        while ((T * (c0 + 1)) / G <= first) ++c0;
        while ((T * c0) / G > first) --c0;
        int c1 = (int)((last * G) / T);
        while ((T * (c1 + 1)) / G <= last) ++c1;
        while ((T * c1) / G > last) --c1;
        if (c1 - c0 + 1 > smax) smax = c1 - c0 + 1;
    }
    return smax;
}

int l2_tc_launch(pm_ctx *ctx, const void *qpack, int mq_pad, const void *tpack, int nt_pad,
                 const float *tnorm, const L2Flags *flags, L2Cand *part, int smax, float *dump)
{
    static bool attr_set = false;
    if (!attr_set) {
        PM_CUDA(ctx, cudaFuncSetAttribute(l2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
        attr_set = true;
    }
    static_assert(sizeof(CUtensorMap) == 128, "tmap_store size");
    CUtensorMap *tmaps = reinterpret_cast<CUtensorMap *>(ctx->tmap_store);
    const void *bases[2] = {qpack, tpack};
    const int rows[2] = {mq_pad, nt_pad}, boxes[2] = {BM, BN};
    for (int k = 0; k < 2; ++k)
        if (ctx->tmap_base[k] != bases[k] || ctx->tmap_rows[k] != rows[k]) {   // re-encode only when the operand moved
            int st = make_tmap(ctx, &tmaps[k], bases[k], rows[k], boxes[k]);
            if (st != PM_OK) return st;
            ctx->tmap_base[k] = bases[k]; ctx->tmap_rows[k] = rows[k];
        }
    const CUtensorMap &tq = tmaps[0], &tt = tmaps[1];
    TcParams P;
    P.tnorm = tnorm; P.flags = flags; P.part = part; P.dump = dump;
    P.MT = mq_pad / (MH * BM); P.NT = nt_pad / BN; P.smax = smax; P.nt_pad = nt_pad; P.mul256 = 256u;
    { const char *dv = getenv("PM_K2_DBG"); P.dbg = dv ? (uint32_t)atoi(dv) : 0u; }
    P.trace = (P.dbg & 32u) ? g_k2_trace : nullptr;
    const int G = l2_tc_grid(ctx, P.MT, P.NT);
    {
        pm_prof_scope prof(ctx, 0);
        cudaError_t le = pm_launch_pdl(l2_tc_kernel, dim3(G), dim3(TC_THREADS), (size_t)SMEM_TOTAL, ctx->stream, tq, tt, P);
        if (le != cudaSuccess) return pm_fail(ctx, PM_CUDA_ERR, "l2_tc_kernel launch: %s", cudaGetErrorString(le));
    }
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}
