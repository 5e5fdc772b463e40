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
// (Measured alternative: letting the epilogue warps write the column constants into the drained accumulator with
// tcgen05.st and accumulating from there saves those 2 of 18 tensor-pipe dispatches per item but doubles the
// epilogue's TMEM traffic -- 1090 instead of 1230 cycles of MMA issue per item, 1300-1700 instead of 880 cycles of
// epilogue: the kernel went from 21.4 to 27.2 us.)
// Operands are bf16 "hi|lo" rows written by K1 (l2.cu):
//   exact-integer mode (SIFT, values 0..255): lo == 0, 2 k-blocks, the GEMM is exact
//   split mode (general floats):  a.b ~ ah.bh + ah.bl + al.bh, 6 k-blocks
//
// Work item = 256 query rows x 128 train columns: TWO 128-row A tiles stay resident in shared
// memory and every B k-block that TMA brings in feeds both, which halves the L2 -> SM operand
// traffic (the binding resource for K = 128: ncu measured 4.4 TB/s of TMA reads with one A tile).
//
// Structure (one persistent CTA per SM, 608 threads):
//   warp 0      TMA producer   : 2 A row-tiles resident (k-block major, one mbarrier per k-block pair so the first
//                                MMAs start when half of A has landed), B stages (two 16 KB k-blocks + the 4 KB norm
//                                image) through a ring; the first loads go out right after griddepcontrol.wait,
//                                before anything else of the prologue
//   warps 1, 18 MMA issuers    : tcgen05.mma cta_group::1 kind::f16, M=128 N=128 K=16, one accumulator per A tile,
//                                double-buffered in TMEM (2 x 2 x 128 columns); warp 1 allocates / frees TMEM
//   warps 2-17  epilogue       : 4 warps per scheduler, each owns 32 rows x 64 columns of a work item:
//                                tcgen05.ld 32x32b.x32, quad minima on the raw bits, branch-free min/max
//                                tournament (VIMNMX/VIMNMX3): top-2 quad minima in exact mode, top-3 pair minima in
//                                split mode, kept in registers across the sweep (K3 re-checks the members exactly).
// Each CTA walks a contiguous range of the (row-tile, column-tile) space; per row tile it
// writes one "segment" of candidates which K3 merges, re-ranks in FP32 and certifies.
#include <cuda.h>
#include <cstdlib>
#include <cuda_bf16.h>
#include <atomic>
#include <vector>
#include "pm_internal.h"
#include "l2_common.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int MH = 2;                        // A row-tiles (m-halves) per work item: 256 rows
constexpr int BLK_BYTES = BM * BK * 2;       // 16 KB: one 128-row x 64-column bf16 k-block of A or B
constexpr int EXT_BYTES = BN * 32;           // 4 KB: K=16 "ext" operand (norm step), K-major, no swizzle
constexpr int STAGE_BYTES = 2 * BLK_BYTES + EXT_BYTES;   // a B stage: two k-blocks (K = 128) + the norm operand of the tile
// A blocks (k-block major: block (kb, h) at (kb * MH + h) * 16 KB) grow up from 0: 64 KB exact, 128 KB split.
// B stages grow DOWN from the top: stage s at AB_BYTES - (s + 1) * 36 KB, so stages 0 and 1 sit at the same
// addresses in both modes (the first loads are issued before the mode is known); exact: 4 stages, split: 2.
constexpr int AB_BYTES = 212992;
constexpr int EPI_WARP0 = 2;        // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-17 epilogue
constexpr int EPI_WARPS = 16;       // 4 per scheduler: enough TLP to keep the issue slots busy
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int ISS2_WARP = EPI_WARP0 + EPI_WARPS;           // warp 18: second MMA issuer (odd work items)
constexpr int TC_THREADS = (ISS2_WARP + 1) * 32;           // 608
constexpr int NSLICE = EPI_WARPS / 4 / MH;                 // column slices per accumulator: 2
constexpr int SLICE = BN / NSLICE;                         // 64 columns per warp
constexpr int SMEM_A = 0;
constexpr int SMEM_BAR = AB_BYTES;
constexpr int SMEM_SCRATCH = SMEM_BAR + 256;                   // MH x (NSLICE-1) x 128 rows x 3 candidates
// "ext" operands of the norm MMA step, K-major, no swizzle: core matrix = 8 rows x 16 B,
// the two 8-element K halves 128 B apart (LBO), 8-row groups 256 B apart (SBO).
//   A side (built here):  row = [1 1 1 bias_h bias_m bias_l 0 0 | 0 x 8]
//   B side (K1, l2.cu):   row = [n_h n_m n_l 1 1 1 0 0 | 0 x 8],  n = ||b||^2 split exactly into three bf16
// so the step adds ||b||^2 + bias to every accumulator element.
constexpr int SMEM_EXTA = SMEM_SCRATCH + MH * (NSLICE - 1) * BM * 3 * 8;
constexpr int SMEM_TOTAL = SMEM_EXTA + BM * 32 + 1024;         // + alignment slack

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
// tag: call site << 16 | work item (hang report only)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t tag = 0)
{
    uint32_t ok, spins = 0;
    unsigned long long t0 = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && (++spins & 1023u) == 0) {          // try_wait itself may block for a while: judge by the clock
            const unsigned long long now = pm_now_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > PM_WAIT_LIMIT_NS) pm_hang_trap(0x30u | (threadIdx.x >> 5 << 8), (bar & 0x3FFu) | (parity << 12), tag, blockIdx.x);
        }
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// one elected lane of a converged warp (the async-proxy instructions take warp-uniform operands:
// issuing them from converged code lets the compiler keep descriptors in uniform registers)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// kind::f8f6f4, A = B = E4M3 (format code 0), D = F32: the Hamming path feeds 0/1 (and 0/-2) bytes, K = 32 per instruction
constexpr uint32_t IDESC_E4M3 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
__device__ __forceinline__ void umma_e4m3(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
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

struct Sel2 {   // exact mode: running (best, second) QUAD minima
    uint32_t m1, m2; int i1, i2;
    __device__ __forceinline__ void reset() { m1 = m2 = 0xFFFFFFFFu; i1 = i2 = -1; }
    // In: the raw accumulator bits of 32 columns (integer-valued floats of one binade: they order like
    // unsigned integers).  The minimum of each aligned group of four columns is taken on the raw bits
    // (16 ops) and only then packed with its QUAD number: key = bits * 256 + (quad + 1), 8 IMADs instead
    // of 32.  K3 recomputes the four members of the two winning quads exactly, so neither the column inside
    // the quad nor ties inside it need to be tracked here.  38 ALU ops + 8 IMAD per 32 elements.
    // qb = quad base of the chunk inside the slice (0 or 8).
    __device__ __forceinline__ void chunk(const uint32_t (&k)[32], uint32_t qb, uint32_t mul)
    {
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t q0 = min(umin3(k[8 * j], k[8 * j + 1], k[8 * j + 2]), k[8 * j + 3]);
            uint32_t q1 = min(umin3(k[8 * j + 4], k[8 * j + 5], k[8 * j + 6]), k[8 * j + 7]);
            q0 = q0 * mul + (uint32_t)(2 * j + 1);
            q1 = q1 * mul + (uint32_t)(2 * j + 2);
            lo[j] = min(q0, q1); hi[j] = max(q0, q1);
        }
#pragma unroll
        for (int w = 2; w >= 1; w >>= 1)
#pragma unroll
            for (int i = 0; i < w; ++i) {
                const uint32_t L = min(lo[i], lo[i + w]);
                const uint32_t S = umin3(max(lo[i], lo[i + w]), hi[i], hi[i + w]);
                lo[i] = L; hi[i] = S;
            }
        // keys carry the quad within the chunk (1..8); make it the quad within the slice (1..16)
        const uint32_t L = lo[0] + qb, S = hi[0] + qb;
        m2 = umin3(m2, S, max(m1, L));
        m1 = min(m1, L);
    }
    // after a tile: resolve the quad base columns of entries that came from it, zero their low byte
    __device__ __forceinline__ void end_tile(int col0)
    {
        const uint32_t n1 = m1 & 0xFFu, n2 = m2 & 0xFFu;
        const int g1 = col0 + 4 * ((int)n1 - 1), g2 = col0 + 4 * ((int)n2 - 1);
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
    const uint8_t *text;         // [NT][4 KB] norm operand images of the train column tiles (K1)
    const L2Flags *flags;
    L2Cand *part;                // [mq_pad][smax][3]; MT counts 256-row super tiles
    float *dump;                 // debug: [mq_pad][nt_pad] of (||b||^2 - 2ab), or null
    int MT, NT, smax, nt_pad;
    const int *sched;            // [G][4]: first item, item count, segment slot of the CTA's first row tile (l2_tc_schedule)
    int trace_cta;               // PM_K2_TRACE builds: the CTA that stamps
    const unsigned long long *chain_done;   // chain pipelining: spin until *chain_done >= wait_seq (null: no spin)
    unsigned long long wait_seq;
    unsigned long long *chain_mark;         // signalling chains: store mark_seq once past the waits (null: none)
    unsigned long long mark_seq;
    unsigned long long *span;    // debug timeline (pm_internal.h), or null
    long long *trace;            // PM_K2_TRACE builds: clock64 stamps of CTA 0, [tile][16]
    uint32_t mul256;             // == 256, passed at run time so the key pack stays an IMAD (FMA pipe), not an ALU LEA
};

#ifdef PM_K2_TRACE
#define TR(slot) do { if (P.trace && (int)blockIdx.x == P.trace_cta && lane == 0 && lt < 64) P.trace[lt * 16 + (slot)] = clock64(); } while (0)
#define TRE(slot) do { if (P.trace && (int)blockIdx.x == P.trace_cta && e == 0 && lane == 0 && lt < 64) P.trace[lt * 16 + (slot)] = clock64(); } while (0)
#else
#define TR(slot) do { } while (0)
#define TRE(slot) do { } while (0)
#endif
#ifdef PM_K2_TRACE
#define TRK(slot) do { if (P.trace && (int)blockIdx.x == P.trace_cta && threadIdx.x == (slot >= 4 ? EPI_WARP0 * 32 : 32)) P.trace[63 * 16 + (slot)] = clock64(); } while (0)
#else
#define TRK(slot) do { } while (0)
#endif

// FP8 = false: bf16 hi|lo operands (L2, K = 128 per stage).  FP8 = true: E4M3 0/1 operands, 256 per row
// (binary descriptors expanded by hamming.cu: ||a-b||^2 of 0/1 vectors IS the Hamming distance), always
// exact mode; the two 128-byte k-blocks of a stage then hold K = 256.
template <bool FP8>
__global__ void __launch_bounds__(TC_THREADS, 1)
l2_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_t, TcParams P)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t sA = sbase + SMEM_A, sBar = sbase + SMEM_BAR;
    const uint32_t bar_full = sBar, bar_empty = sBar + 32;       // [4] each: B stage loaded / consumed
    const uint32_t bar_ak0 = sBar + 64, bar_ak1 = sBar + 72;     // A hi k-block 0 / 1 (both row tiles) loaded
    const uint32_t bar_alo = sBar + 80, bar_aempty = sBar + 88;  // A lo k-blocks loaded (split mode) / A consumed
    const uint32_t bar_tfull = sBar + 96, bar_tempty = sBar + 112;  // [2]: accumulator ready / drained
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(sgen + SMEM_BAR + 128);
    const uint32_t sExtA = sbase + SMEM_EXTA;
    L2Cand *scratch = reinterpret_cast<L2Cand *>(sgen + SMEM_SCRATCH);
    auto stage_addr = [&](int st) { return sA + (uint32_t)(AB_BYTES - (st + 1) * STAGE_BYTES); };
    constexpr int KBW = FP8 ? 2 * BK : BK;     // tensor-map columns per k-block (elements)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TRK(0);

    // ---- setup that does not depend on the previous kernel: overlaps its tail under PDL ----
    const int cta = (int)blockIdx.x;
    const int4 sch = __ldg(reinterpret_cast<const int4 *>(P.sched) + cta);     // written by the host before this chain was enqueued
    const int t_begin = sch.x, ntiles = sch.y;
    const int NT = P.NT;
    const int m_first = t_begin / NT, n_first = t_begin - m_first * NT;   // the only divisions: per-role counters follow
    if (warp == 0 && lane == 0) {
        // the producer initialises the barriers itself: it is their first user (the loads below), everybody else
        // meets them after the block-wide sync further down
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_t)) : "memory");
        for (int s = 0; s < 4; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_ak0, 1); mbar_init(bar_ak1, 1); mbar_init(bar_alo, 1);
        mbar_init(bar_aempty, 2);                       // one arrival from each MMA issuer
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32((const void *)tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    TRK(1);
    pm_span_mark(P.span, 3, false);
    pm_pdl_prologue();
    // pipelined chains: K1 of this chain ran ahead of the previous chain's tail; nothing below may start
    // before that chain has completed (the __syncthreads further down holds the other threads)
    pm_chain_wait(P.chain_done, P.wait_seq);
    // "K2 of chain s is past its waits" = K1(s) and everything enqueued before chain s have completed: the
    // word K1 of chain s+1 spins on when it runs ahead
    if (P.chain_mark && blockIdx.x == 0 && threadIdx.x == 0)
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(P.chain_mark), "l"(P.mark_seq) : "memory");
    pm_span_mark(P.span, 4, false);      // K1's outputs (flags, packed operands, norm images) are complete past this point
    // First loads, issued by the thread that waited above, before the flags are even read: the hi k-blocks of the first
    // row tile's A and the first B stage (hi k-blocks + norm image) have the same shared-memory addresses in both modes.
    if (warp == 0 && lane == 0 && ntiles > 0) {
        mbar_expect_tx(bar_full, (uint32_t)STAGE_BYTES);
        bulk_load_1d(stage_addr(0) + 2 * BLK_BYTES, P.text + (size_t)n_first * EXT_BYTES, EXT_BYTES, bar_full);
        mbar_expect_tx(bar_ak0, (uint32_t)(MH * BLK_BYTES));
        for (int h = 0; h < MH; ++h) tma_load_2d(sA + (0 * MH + h) * BLK_BYTES, &tmap_q, bar_ak0, 0, (m_first * MH + h) * BM);
        tma_load_2d(stage_addr(0), &tmap_t, bar_full, 0, n_first * BN);
        mbar_expect_tx(bar_ak1, (uint32_t)(MH * BLK_BYTES));
        for (int h = 0; h < MH; ++h) tma_load_2d(sA + (1 * MH + h) * BLK_BYTES, &tmap_q, bar_ak1, KBW, (m_first * MH + h) * BM);
        tma_load_2d(stage_addr(0) + BLK_BYTES, &tmap_t, bar_full, KBW, n_first * BN);
    }
    TRK(2);
    const L2Flags fl = *P.flags;
    const bool exact = l2_exact_mode(fl);
    const int nsp = exact ? 1 : 2;           // B stages (k-block pairs) per work item
    const int nstage = exact ? 4 : 2;        // B ring depth
    const float shift = l2_split_shift(fl.max_qnorm_bits);
    const float nb_off = exact ? L2_EXACT_BIAS : shift;            // bias folded into the norm step

    if (warp >= EPI_WARP0 && (int)threadIdx.x - EPI_WARP0 * 32 < BM) {
        // constant A operand of the norm step: row r = [1 1 1 bias_h bias_m bias_l 0 0 | 0 x 8] in bf16
        const int r = (int)threadIdx.x - EPI_WARP0 * 32;
        const uint4 b3 = bf16_split3(nb_off);          // .x = h | m << 16, .y = l
        uint8_t *dst = sgen + SMEM_EXTA + ext_row_offset(r);
        *reinterpret_cast<uint4 *>(dst) = make_uint4(0x3F803F80u, 0x00003F80u | (b3.x << 16), (b3.x >> 16) | (b3.y << 16), 0u);
        *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    TRK(3);

    if (warp == 0) {
        // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
        int m = m_first, n = n_first, cur_m = -1;
        int stage = 0; uint32_t sphase = 0, apar = 0;
        for (int lt = 0; lt < ntiles; ++lt) {
            if (m != cur_m) {
                if (cur_m >= 0) { mbar_wait(bar_aempty, apar, (1u << 16) | (uint32_t)lt); apar ^= 1; }    // every MMA of the old row tile is done
                if (elect_one()) {
                    if (lt > 0) {                                               // (the first row tile's hi blocks are on their way)
                        mbar_expect_tx(bar_ak0, (uint32_t)(MH * BLK_BYTES));
                        for (int h = 0; h < MH; ++h) tma_load_2d(sA + (0 * MH + h) * BLK_BYTES, &tmap_q, bar_ak0, 0, (m * MH + h) * BM);
                        mbar_expect_tx(bar_ak1, (uint32_t)(MH * BLK_BYTES));
                        for (int h = 0; h < MH; ++h) tma_load_2d(sA + (1 * MH + h) * BLK_BYTES, &tmap_q, bar_ak1, KBW, (m * MH + h) * BM);
                    }
                    if (!exact) {
                        mbar_expect_tx(bar_alo, (uint32_t)(2 * MH * BLK_BYTES));
                        for (int kb = 2; kb < 4; ++kb)
                            for (int h = 0; h < MH; ++h)
                                tma_load_2d(sA + (kb * MH + h) * BLK_BYTES, &tmap_q, bar_alo, kb * KBW, (m * MH + h) * BM);
                    }
                }
                __syncwarp();
                cur_m = m;
            }
            for (int sp = 0; sp < nsp; ++sp) {
                // (not for the very first stage: it was filled before this loop, and an issuer that has already consumed
                // it has moved `empty` into its next phase -- waiting for the "previous" phase here would never return)
                if (lt > 0 || sp > 0) mbar_wait(bar_empty + 8 * stage, sphase ^ 1, (2u << 16) | (uint32_t)lt);
                if (elect_one()) {
                    const uint32_t fb = bar_full + 8 * stage, dst = stage_addr(stage);
                    if (lt > 0 || sp > 0) {                                     // (the very first stage is on its way)
                        mbar_expect_tx(fb, sp == 0 ? STAGE_BYTES : 2 * BLK_BYTES);
                        // stages: hi0 hi1 (+ norm image) | lo0 lo1   (packed row = [hi 0..127 | lo 128..255])
                        const int kc = sp == 1 ? 2 * KBW : 0;
                        tma_load_2d(dst, &tmap_t, fb, kc, n * BN);
                        tma_load_2d(dst + BLK_BYTES, &tmap_t, fb, kc + KBW, n * BN);
                        if (sp == 0) bulk_load_1d(dst + 2 * BLK_BYTES, P.text + (size_t)n * EXT_BYTES, EXT_BYTES, fb);
                    }
                }
                __syncwarp();
                if (++stage == nstage) { stage = 0; sphase ^= 1; }
            }
            if (++n == NT) { n = 0; ++m; }
        }
    } else if (warp == 1 || warp == ISS2_WARP) {
        // ===================== MMA issuers (whole warp converged, one elected lane issues) =====================
        // Two issuer warps alternate work items (warp 1: even, warp 18: odd).  An mbarrier wait that follows
        // tcgen05.commit in the same warp stalls until the committed MMAs drain; with two issuers those
        // stalls (and the barrier round trips) of one warp hide behind the other warp's MMAs.  A pair of
        // named barriers hands the issue order over, so MMAs still reach the tensor pipe in item order.
        const int x = warp == 1 ? 0 : 1;
        int m = m_first, n = n_first + x;
        if (n >= NT) { n -= NT; ++m; }
        const uint64_t extA = make_sdesc_ext(sExtA);
        // one 64-wide k-block: A block `ablk` (0,1 hi / 2,3 lo) of both row tiles x B block `bk` of the stage at sb;
        // `first`: the norm step goes in front (it overwrites the accumulators)
        auto issue_kblock = [&](int ablk, int bk, bool first, uint32_t d_tmem, uint32_t sb) {
            if (first) {
                // acc = [1 1 1 bias..] x [split3(||b||^2) 1 1 1 ..]
                const uint64_t extB = make_sdesc_ext(sb + 2 * BLK_BYTES);
#pragma unroll
                for (int h = 0; h < MH; ++h) umma_bf16(d_tmem + h * BN, extA, extB, IDESC, 0u);
            }
            const uint64_t bdesc = make_sdesc(sb + bk * BLK_BYTES);
#pragma unroll
            for (int h = 0; h < MH; ++h) {      // the same B k-block feeds both A tiles
                const uint64_t adesc = make_sdesc(sA + (ablk * MH + h) * BLK_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) { // +32 B per MMA (K=16 bf16 / K=32 e4m3) inside the swizzle span
                    if (FP8) umma_e4m3(d_tmem + h * BN, adesc + 2 * k, bdesc + 2 * k, IDESC_E4M3, 1u);
                    else umma_bf16(d_tmem + h * BN, adesc + 2 * k, bdesc + 2 * k, IDESC, 1u);
                }
            }
        };
        int a_seen = 0;                      // A row tiles this warp has waited for (never skips a phase of the A barriers)
        for (int lt = x; lt < ntiles; lt += 2) {
            TR(0);
            const uint32_t acc = (uint32_t)x;
            const int a_need = m - m_first;
            while (a_seen < a_need) {        // row tiles this warp had no item in
                const uint32_t ap = (uint32_t)(a_seen & 1);
                mbar_wait(bar_ak0, ap, (3u << 16) | (uint32_t)lt); mbar_wait(bar_ak1, ap, (3u << 16) | (uint32_t)lt);
                if (!exact) mbar_wait(bar_alo, ap, (3u << 16) | (uint32_t)lt);
                ++a_seen;
            }
            const bool a_fresh = a_seen == a_need;       // first item of this warp in the row tile: A may still be landing
            const uint32_t ap = (uint32_t)(a_need & 1);
            mbar_wait(bar_tempty + 8 * acc, (uint32_t)(((lt >> 1) & 1) ^ 1), (4u << 16) | (uint32_t)lt);     // accumulator drained by the epilogue of item lt - 2
            const uint32_t d_tmem = tmem_base + acc * (MH * BN);     // accumulator (acc, h) at + h * BN
            const bool more = lt + 1 < ntiles;
            // split mode: the B ring (2 stages) holds exactly one item, so this warp may only LOOK at the ring once the
            // other issuer is done with item lt - 1 -- an mbarrier parity wait one whole ring ahead would alias with the
            // phase before it and return at once
            if (!exact && lt > 0) asm volatile("bar.sync %0, 64;" ::"r"(2 + x) : "memory");
            for (int sp = 0; sp < nsp; ++sp) {
                // exact: one stage per item, ring of 4.  split: two stages per item -- [B hi + norm image], [B lo] -- ring
                // of 2: stage 0 serves A hi x B hi AND A lo x B hi, stage 1 serves A hi x B lo (an earlier version loaded
                // B hi a second time as a third stage)
                const int sidx = exact ? lt : 2 * lt + sp;           // running B stage index
                const int st = exact ? (sidx & 3) : (sidx & 1);
                mbar_wait(bar_full + 8 * st, (uint32_t)((exact ? (sidx >> 2) : (sidx >> 1)) & 1), (5u << 16) | (uint32_t)lt);
                if (a_fresh && sp == 0) mbar_wait(bar_ak0, ap, (6u << 16) | (uint32_t)lt);
                tc_fence_after();
                if (sp == 0) {
                    TR(1);
                    if (exact && lt > 0) asm volatile("bar.sync %0, 64;" ::"r"(2 + x) : "memory");   // item lt-1 has been issued
                    TR(2);
                }
                const uint32_t sb = stage_addr(st);
                if (sp == 0) {
                    if (a_fresh) {
                        // first item of this warp in the row tile: start on k-block 0 while the rest of A is still landing
                        if (elect_one()) issue_kblock(0, 0, true, d_tmem, sb);
                        __syncwarp();
                        mbar_wait(bar_ak1, ap, (8u << 16) | (uint32_t)lt);
                        tc_fence_after();
                        if (elect_one()) issue_kblock(1, 1, false, d_tmem, sb);
                        __syncwarp();
                        if (!exact) { mbar_wait(bar_alo, ap, (7u << 16) | (uint32_t)lt); tc_fence_after(); }
                    } else if (elect_one()) {
                        issue_kblock(0, 0, true, d_tmem, sb);
                        issue_kblock(1, 1, false, d_tmem, sb);
                    }
                    if (!exact && elect_one()) {       // A lo x B hi
                        issue_kblock(2, 0, false, d_tmem, sb);
                        issue_kblock(3, 1, false, d_tmem, sb);
                    }
                } else if (elect_one()) {              // A hi x B lo
                    issue_kblock(0, 0, false, d_tmem, sb);
                    issue_kblock(1, 1, false, d_tmem, sb);
                }
                __syncwarp();
                if (sp == nsp - 1) {
                    TR(3);
                    if (more) asm volatile("bar.arrive %0, 64;" ::"r"(3 - x) : "memory");
                }
                if (elect_one()) {
                    umma_commit(bar_empty + 8 * st);
                    if (sp == nsp - 1) {
                        umma_commit(bar_tfull + 8 * acc);
                        // the A tiles may be overwritten once the last TWO items of the row (one per issuer) are done
                        if (n == NT - 1 && more) { umma_commit(bar_aempty); if (lt == 0) mbar_arrive(bar_aempty); }
                        if (n == NT - 2 && lt + 2 < ntiles) umma_commit(bar_aempty);
                    }
                }
                __syncwarp();
            }
            if (a_fresh) ++a_seen;
            TR(4);
            n += 2;
            if (n >= NT) { n -= NT; ++m; }
        }
    } else {
        // ===================== epilogue =====================
        const int e = warp - EPI_WARP0;
        const int quarter = warp & 3;          // TMEM lane quarter this warp may access
        const int mh = (e >> 2) & (MH - 1);    // which A row-tile (accumulator) of the work item
        const int slice = e >> 3;              // which 64 of the item's 128 columns
        const int row = mh * BM + quarter * 32 + lane;   // row within the 256-row work item
        const uint32_t mul = P.mul256;
        int cur_m = -1;
        int m = m_first, n = n_first;
        const int slot_first = sch.z;          // this CTA's index among the CTAs that touch row tile m_first
        Sel2 s2; s2.reset();
        Sel3 s3; s3.reset();
        const uint32_t taddr_warp = tmem_base + ((uint32_t)(quarter * 32) << 16) + mh * BN + slice * SLICE;

        auto flush = [&](int mm) {
            // segment slot = index of this CTA among the CTAs that touch row tile mm: only the
            // CTA's first row tile can have started in an earlier CTA
            const int slot = mm == m_first ? slot_first : 0;
            L2Cand *dst = P.part + ((size_t)(mm * MH * BM + row) * P.smax + slot) * 3;
            if (exact) {
                // (value, quad base column) as one u64 per entry: the two slices of a row merge with four min / max
                unsigned long long k1 = (s2.m1 >> 8) != 0xFFFFFFu ? (((unsigned long long)(s2.m1 >> 8) << 32) | (unsigned)s2.i1) : ~0ull;
                unsigned long long k2 = (s2.m2 >> 8) != 0xFFFFFFu ? (((unsigned long long)(s2.m2 >> 8) << 32) | (unsigned)s2.i2) : ~0ull;
                unsigned long long *sc = reinterpret_cast<unsigned long long *>(scratch);
                if (slice > 0) { sc[((slice - 1) * MH * BM + row) * 2] = k1; sc[((slice - 1) * MH * BM + row) * 2 + 1] = k2; }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                if (slice == 0) {
                    for (int sl = 0; sl < NSLICE - 1; ++sl) {
                        const unsigned long long o1 = sc[(sl * MH * BM + row) * 2], o2 = sc[(sl * MH * BM + row) * 2 + 1];
                        const unsigned long long lo = min_u64(k1, o1), hi = max_u64(k1, o1);
                        k2 = min_u64(min_u64(k2, o2), hi); k1 = lo;
                    }
                    dst[0] = k1 == ~0ull ? L2Cand{__int_as_float(0x7f800000), -1} : L2Cand{(float)((int)(k1 >> 32) - 0x400000), (int)(unsigned)k1};
                    dst[1] = k2 == ~0ull ? L2Cand{__int_as_float(0x7f800000), -1} : L2Cand{(float)((int)(k2 >> 32) - 0x400000), (int)(unsigned)k2};
                    dst[2] = L2Cand{__int_as_float(0x7f800000), -1};
                }
            } else {
                Cand3 c; c.reset();
                if ((s3.m1 >> 8) < 0x7EFFFFu) c.insert(__uint_as_float(s3.m1) - shift, s3.i1);
                if ((s3.m2 >> 8) < 0x7EFFFFu) c.insert(__uint_as_float(s3.m2) - shift, s3.i2);
                if ((s3.m3 >> 8) < 0x7EFFFFu) c.insert(__uint_as_float(s3.m3) - shift, s3.i3);
                if (slice > 0) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) scratch[((slice - 1) * MH * BM + row) * 3 + k] = L2Cand{c.d[k], c.i[k]};
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                if (slice == 0) {
                    for (int sl = 0; sl < NSLICE - 1; ++sl)
#pragma unroll
                        for (int k = 0; k < 3; ++k) { const L2Cand o = scratch[(sl * MH * BM + row) * 3 + k]; c.insert(o.d, o.idx); }
#pragma unroll
                    for (int k = 0; k < 3; ++k) dst[k] = L2Cand{c.d[k], c.i[k]};
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            s2.reset(); s3.reset();
        };

        for (int lt = 0; lt < ntiles; ++lt) {
            const uint32_t acc = (uint32_t)(lt & 1), acc_phase = (uint32_t)((lt >> 1) & 1);
            if (m != cur_m) { if (cur_m >= 0) flush(cur_m); cur_m = m; }
            TRE(12);
            mbar_wait(bar_tfull + 8 * acc, acc_phase, (9u << 16) | (uint32_t)lt);
            tc_fence_after();
            TRE(13);
            if (lt == 0) TRK(4);
            const int col0 = n * BN + slice * SLICE;
            const uint32_t taddr0 = taddr_warp + acc * (MH * BN);
#pragma unroll 1
            for (int ch = 0; ch < SLICE / 32; ++ch) {
                uint32_t r[32];
                tmem_ld32(taddr0 + ch * 32, r);
                tmem_ld_wait();
                if (P.dump) {
                    float *drow = P.dump + (size_t)(m * MH * BM + row) * P.nt_pad + col0 + ch * 32;
#pragma unroll
                    for (int c = 0; c < 32; ++c) drow[c] = __uint_as_float(r[c]) - nb_off;
                }
                if (exact) {
                    s2.chunk(r, (uint32_t)(ch * 8), mul);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) r[c] = __byte_perm(r[c], (uint32_t)(c + 1), 0x3214);
                    s3.chunk(r, (uint32_t)(ch * 32));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
            TRE(14);
            if (exact) s2.end_tile(col0); else s3.end_tile(col0);
            if (++n == NT) { n = 0; ++m; }
        }
        TRK(5);
        if (cur_m >= 0) flush(cur_m);
        TRK(6);
    }

    tc_fence_before();
    __syncthreads();
    pm_span_mark(P.span, 5, true);
    TRK(7);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap(pm_ctx *ctx, CUtensorMap *tm, const void *base, int rows_pad, int box_rows, int fp8)
{
    if (!ctx->tmap_encode) {
        cudaDriverEntryPointQueryResult qres;
        void *fn = nullptr;
        PM_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess)
            return pm_fail(ctx, PM_CUDA_ERR, "cuTensorMapEncodeTiled entry point not found");
        ctx->tmap_encode = fn;
    }
    // bf16: rows of 256 elements (hi|lo), boxes of 64; e4m3: rows of 256 bytes, boxes of 128 -- 128-byte spans either way
    cuuint64_t dims[2] = {(cuuint64_t)L2_PACK_COLS, (cuuint64_t)rows_pad};
    cuuint64_t strides[1] = {(cuuint64_t)L2_PACK_COLS * (fp8 ? 1 : 2)};
    cuuint32_t box[2] = {(cuuint32_t)(fp8 ? 2 * BK : BK), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((tmap_encode_fn)ctx->tmap_encode)(tm, fp8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base),
                                                    dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return pm_fail(ctx, PM_CUDA_ERR, "cuTensorMapEncodeTiled failed: %d", (int)r);
    return PM_OK;
}

}  // namespace

static long long *g_k2_trace = nullptr;
static int g_k2_trace_cta = 0;
extern "C" void pm_debug_set_k2_trace(long long *p) { g_k2_trace = p; }
extern "C" void pm_debug_set_k2_trace_cta(int cta) { g_k2_trace_cta = cta; }

// Work partition of K2: CTA c owns `count` consecutive items of the (row tile, column tile) space starting at `first`.
// A CTA whose range crosses into the next row tile pays for it -- it reloads its two A tiles (a bubble of the tensor pipe)
// and flushes one more candidate segment, ~2 work items' worth of time at cfg2 -- and those CTAs used to finish last.  The
// partition therefore balances COST = items + L2_SCHED_BOUNDARY_COST per row-tile crossing instead of item counts: the
// smallest makespan for which a left-to-right greedy fill covers all items is found by bisection.  The table ([G][4] ints:
// first, count, segment slot of the first row tile, 0) lives in device memory and is rebuilt only when the shape changes.
constexpr double L2_SCHED_BOUNDARY_COST = 2.0;

struct L2Sched {
    int MT = -1, NT = -1, G = 0, smax = 1;
    bool uploaded = false;
    unsigned long long last_use = 0;
    std::vector<int> table;      // [G][4]
};
constexpr int L2_SCHED_CACHE = 8;          // shapes kept (device copies side by side in one workspace slot)
struct L2SchedCache {
    L2Sched e[L2_SCHED_CACHE];
    unsigned long long clock = 0;
};

static void l2_sched_build(L2Sched &S, int MT, int NT, int G)
{
    const long long T = (long long)MT * NT;
    S.MT = MT; S.NT = NT; S.G = G;
    S.table.assign((size_t)G * 4, 0);
    auto fill = [&](double L, std::vector<int> *out) -> bool {
        long long s = 0;
        for (int c = 0; c < G; ++c) {
            // largest k >= 1 with k + B * crossings(s, k) <= L
            long long k = 0;
            if (s < T) {
                long long lo = 1, hi = T - s;
                auto cost = [&](long long kk) { return (double)kk + L2_SCHED_BOUNDARY_COST * (double)((s + kk - 1) / NT - s / NT); };
                if (cost(1) > L) k = 1;          // every CTA that has work left takes at least one item
                else {
                    while (lo < hi) { const long long mid = (lo + hi + 1) / 2; if (cost(mid) <= L) lo = mid; else hi = mid - 1; }
                    k = lo;
                }
            }
            if (out) { (*out)[(size_t)c * 4] = (int)s; (*out)[(size_t)c * 4 + 1] = (int)k; }
            s += k;
        }
        return s >= T;
    };
    double lo = (double)T / G, hi = (double)T / G + L2_SCHED_BOUNDARY_COST * ((double)MT / G + 2.0) + 2.0;
    while (!fill(hi, nullptr)) hi *= 1.5;
    for (int it = 0; it < 40; ++it) { const double mid = 0.5 * (lo + hi); if (fill(mid, nullptr)) hi = mid; else lo = mid; }
    fill(hi, &S.table);
    // segment slots: CTA c's slot in row tile m = number of CTAs before it that touch m; only a CTA's FIRST row tile can
    // have started in an earlier CTA
    S.smax = 1;
    std::vector<int> touch((size_t)MT, 0);
    for (int c = 0; c < G; ++c) {
        const int first = S.table[(size_t)c * 4], cnt = S.table[(size_t)c * 4 + 1];
        if (cnt <= 0) continue;
        const int m0 = first / NT, m1 = (first + cnt - 1) / NT;
        S.table[(size_t)c * 4 + 2] = touch[(size_t)m0];
        for (int m = m0; m <= m1; ++m) { ++touch[(size_t)m]; if (touch[(size_t)m] > S.smax) S.smax = touch[(size_t)m]; }
    }
}

// the cached partition for this shape (least recently used entry replaced); *slot = its index
static L2Sched &l2_sched_get(pm_ctx *ctx, int MT, int NT, int *slot = nullptr)
{
    if (!ctx->l2_sched) ctx->l2_sched = new L2SchedCache();
    L2SchedCache &C = *static_cast<L2SchedCache *>(ctx->l2_sched);
    const int G = l2_tc_grid(ctx, MT, NT);
    int hit = -1, lru = 0;
    for (int k = 0; k < L2_SCHED_CACHE; ++k) {
        if (C.e[k].MT == MT && C.e[k].NT == NT && C.e[k].G == G) hit = k;
        if (C.e[k].last_use < C.e[lru].last_use) lru = k;
    }
    if (hit < 0) { hit = lru; l2_sched_build(C.e[hit], MT, NT, G); C.e[hit].uploaded = false; }
    C.e[hit].last_use = ++C.clock;
    if (slot) *slot = hit;
    return C.e[hit];
}
void l2_sched_free(pm_ctx *ctx) { delete static_cast<L2SchedCache *>(ctx->l2_sched); ctx->l2_sched = nullptr; }

int l2_tc_grid(pm_ctx *ctx, int MT, int NT)
{
    long long T = (long long)MT * NT;
    return (int)(T < ctx->num_sms ? T : ctx->num_sms);
}

// Max number of CTAs (segments) that can touch one row tile under the partition above.
int l2_tc_smax(pm_ctx *ctx, int MT, int NT) { return l2_sched_get(ctx, MT, NT).smax; }

static int g_k2_repeat = 1;
extern "C" void pm_debug_k2_repeat(int n) { g_k2_repeat = n > 1 ? (n < 64 ? n : 64) : 1; }

int l2_tc_launch(pm_ctx *ctx, const void *qpack, int mq_pad, const void *tpack, int nt_pad,
                 const void *text, const L2Flags *flags, L2Cand *part, int smax, float *dump, int fp8,
                 int tmap_set, const unsigned long long *chain_done, unsigned long long wait_seq,
                 unsigned long long *chain_mark, unsigned long long mark_seq)
{
    // per DEVICE, not per process: a host with one thread (and one ctx) per GPU needs it on every device it drives
    static std::atomic<unsigned long long> attr_devices{0};     // lanes of the batched pair call launch from several host threads
    const unsigned long long dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_devices.load(std::memory_order_acquire) & dev_bit)) {
        PM_CUDA(ctx, cudaSetDevice(ctx->device));
        PM_CUDA(ctx, cudaFuncSetAttribute(l2_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
        PM_CUDA(ctx, cudaFuncSetAttribute(l2_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
        attr_devices.fetch_or(dev_bit, std::memory_order_release);
    }
    static_assert(sizeof(CUtensorMap) == 128, "tmap_store size");
    CUtensorMap *tmaps = reinterpret_cast<CUtensorMap *>(ctx->tmap_store) + 2 * tmap_set;   // one cached pair per buffer set
    const void *bases[2] = {qpack, tpack};
    const int rows[2] = {mq_pad, nt_pad}, boxes[2] = {BM, BN};
    for (int k = 0; k < 2; ++k) {
        const int c = 2 * tmap_set + k;
        if (ctx->tmap_base[c] != bases[k] || ctx->tmap_rows[c] != rows[k] || ctx->tmap_fp8[c] != fp8) {   // re-encode only when the operand moved
            int st = make_tmap(ctx, &tmaps[k], bases[k], rows[k], boxes[k], fp8);
            if (st != PM_OK) return st;
            ctx->tmap_base[c] = bases[k]; ctx->tmap_rows[c] = rows[k]; ctx->tmap_fp8[c] = fp8;
        }
    }
    const CUtensorMap &tq = tmaps[0], &tt = tmaps[1];
    TcParams P;
    P.text = (const uint8_t *)text; P.flags = flags; P.part = part; P.dump = dump;
    P.MT = mq_pad / (MH * BM); P.NT = nt_pad / BN; P.smax = smax; P.nt_pad = nt_pad; P.mul256 = 256u;
    P.trace = g_k2_trace; P.span = g_pm_span;
    P.chain_done = chain_done; P.wait_seq = wait_seq; P.chain_mark = chain_mark; P.mark_seq = mark_seq;
    if ((long long)P.MT * P.NT >= (1ll << 30)) return pm_fail(ctx, PM_BAD_ARG, "L2 matching: more than 2^30 work items");
    int sslot = 0;
    L2Sched &S = l2_sched_get(ctx, P.MT, P.NT, &sslot);
    const int G = S.G;
    const bool ws_fresh = ctx->slot_ptr[WS_L2_SCHED] == nullptr;
    PM_WS(ctx, dsched_all, int *, WS_L2_SCHED, (size_t)L2_SCHED_CACHE * ctx->num_sms * 16);
    if (ws_fresh) for (int k = 0; k < L2_SCHED_CACHE; ++k) static_cast<L2SchedCache *>(ctx->l2_sched)->e[k].uploaded = false;
    int *dsched = dsched_all + (size_t)sslot * ctx->num_sms * 4;
    if (!S.uploaded) {
        // A new shape: the copy is enqueued on the stream, i.e. after every earlier kernel that may still read the entry
        // it replaces and before this launch (a copy is not subject to programmatic early launch).  Pageable source: the
        // runtime stages it before the call returns, so the host vector may change afterwards.
        PM_CUDA(ctx, cudaMemcpyAsync(dsched, S.table.data(), S.table.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        S.uploaded = true;
    }
    P.sched = dsched; P.trace_cta = g_k2_trace_cta;
    {
        pm_prof_scope prof(ctx, fp8 ? 1 : 0);       // profile class 1 = the Hamming matching kernel
        // pm_debug_k2_repeat(n): the launch is repeated n times inside ONE event pair (the kernel only reads its operands and
        // rewrites the same candidates, so the result does not change) -- its steady-state duration, with the launch
        // latency and prologue of launch k + 1 under launch k as in any back-to-back use
        for (int rep = 0; rep < (g_k2_repeat > 1 ? g_k2_repeat : 1); ++rep) {
            cudaError_t le = fp8 ? pm_launch_pdl(l2_tc_kernel<true>, dim3(G), dim3(TC_THREADS), (size_t)SMEM_TOTAL, ctx->stream, tq, tt, P)
                                 : pm_launch_pdl(l2_tc_kernel<false>, dim3(G), dim3(TC_THREADS), (size_t)SMEM_TOTAL, ctx->stream, tq, tt, P);
            if (le != cudaSuccess) return pm_fail(ctx, PM_CUDA_ERR, "l2_tc_kernel launch: %s", cudaGetErrorString(le));
            ++ctx->launches;
        }
        --ctx->launches;                            // PM_CHECK_LAUNCH below counts one
    }
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

int pm_hang_init_l2_tc(pm_hang_rec *dev_view)
{
    return cudaMemcpyToSymbol(g_pm_hang_rec, &dev_view, sizeof(dev_view)) == cudaSuccess ? PM_OK : PM_CUDA_ERR;
}
