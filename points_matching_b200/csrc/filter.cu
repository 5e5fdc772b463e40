// filter.cu -- K5: good-match filters with order-preserving stream compaction, and the
// match -> correspondence gather.
//
// Replaces the host loops at /root/reference/Points Matching/main.cpp:49-69 (min/max
// midpoint rule), the ratio test / cross-check the north_star names for the same step,
// and main.cpp:71-79 + 89-91 (index lists + KeyPoint::convert).
//
// All of it is streaming, HBM-bound work over <= a few MB: ONE kernel does predicate, scan
// (decoupled look-back across tiles) and scatter; output stays in queryIdx order, exactly
// like the reference's push_back loop.
#include "pm_internal.h"

namespace {

constexpr int FB = 1024;   // rows per compaction block

struct RatioPred {
    const pm_dmatch *knn; float ratio;
    __device__ bool operator()(int i, pm_dmatch &m) const {
        const pm_dmatch a = knn[(size_t)i * 2], b = knn[(size_t)i * 2 + 1];
        m = a;
        return a.trainIdx >= 0 && b.trainIdx >= 0 && a.distance < ratio * b.distance;
    }
};
struct CrossPred {
    const pm_dmatch *knn; int stride; const unsigned long long *col_best; int nt;
    __device__ bool operator()(int i, pm_dmatch &m) const {
        m = knn[(size_t)i * stride];
        const int j = m.trainIdx;
        if (j < 0 || j >= nt) return false;
        return (unsigned)(col_best[j] & 0xFFFFFFFFull) == (unsigned)m.queryIdx;
    }
};
struct MinMaxPred {
    const pm_dmatch *m_in; int stride; const unsigned *mm_bits;
    __device__ bool operator()(int i, pm_dmatch &m) const {
        m = m_in[(size_t)i * stride];
        const double mn = (double)__uint_as_float(mm_bits[0]), mx = (double)__uint_as_float(mm_bits[1]);
        return (double)m.distance < mn + (mx - mn) / 2;      // main.cpp:65
    }
};

__device__ __forceinline__ int block_exclusive_scan(int v, int *total)
{
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_sums[w] = x;
    __syncthreads();
    if (w == 0) {
        int s = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
        warp_sums[lane] = s;
    }
    __syncthreads();
    const int base = w > 0 ? warp_sums[w - 1] : 0;
    *total = warp_sums[(blockDim.x >> 5) - 1];
    __syncthreads();
    return base + x - v;
}

// Single-pass order-preserving compaction (decoupled look-back).  Tiles ALWAYS take tickets from an atomic counter, so
// a tile only ever waits on tiles that have already started -- no assumption about how many CTAs are resident at once
// or in which order the hardware dispatches them (other streams, the lanes of the batched pair call and other
// processes may share the GPU).  The counter is never reset: it counts the tiles of ALL calls of this ctx, and the host
// passes the number handed out before this call (ticket_base; 64 bits), so tile = ticket - ticket_base.  A block takes
// its ticket BEFORE it lets the next kernel of the stream launch (griddepcontrol.launch_dependents), so the tickets of
// consecutive calls cannot interleave, and before its own griddepcontrol.wait, so the atomic's round trip hides under the
// predecessor's tail.  status[t] = epoch << 34 | state << 32 | value with state 1 = tile aggregate, 2 = inclusive prefix;
// the epoch (one per call) makes stale words from earlier calls invisible, so nothing is cleared between calls.
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <class Pred>
__global__ void __launch_bounds__(FB) compact_lookback_kernel(Pred pred, int n, pm_dmatch *out, int32_t *n_out,
                                                              unsigned long long *status, unsigned long long *counter,
                                                              unsigned long long ticket_base, unsigned epoch,
                                                              unsigned long long *span, unsigned long long *chain_done,
                                                              unsigned *chain_ctr, unsigned long long chain_seq, pm_gather_out g)
{
    __shared__ int s_tile, s_prefix;
    pm_span_mark(span, 12, false);
    if (threadIdx.x == 0) s_tile = (int)(atomicAdd(counter, 1ull) - ticket_base);
    pm_pdl_prologue();
    pm_span_mark(span, 13, false);
    __syncthreads();
    const int tile = s_tile;
    const int ntiles = (n + FB - 1) / FB;
    if ((unsigned)tile >= (unsigned)ntiles) {      // cannot happen while host and device ticket counts agree; never spin on it
        pm_chain_signal(chain_done, chain_ctr, chain_seq);
        return;
    }
    const int i = tile * FB + threadIdx.x;
    pm_dmatch m;
    const int f = i < n ? (int)pred(i, m) : 0;
    int total;
    const int ex = block_exclusive_scan(f, &total);
    if (threadIdx.x < 32) {
        // warp-wide decoupled look-back: 32 predecessors per round trip
        const int lane = threadIdx.x;
        const unsigned long long tag = (unsigned long long)epoch << 34;
        int prefix = 0;
        if (tile == 0) {
            if (lane == 0) {
                st_volatile_u64(&status[0], tag | (2ull << 32) | (unsigned)total);
            }
        } else {
            if (lane == 0) st_volatile_u64(&status[tile], tag | (1ull << 32) | (unsigned)total);
            unsigned spins = 0;
            for (int hi = tile - 1; hi >= 0;) {
                if (++spins > (1u << 22)) pm_hang_trap(0x20u, (unsigned)tile, (unsigned)hi, epoch);
                const int p = hi - lane;
                unsigned long long v = tag | (2ull << 32);                 // before tile 0: inclusive prefix 0
                if (p >= 0) v = ld_volatile_u64(&status[p]);
                const bool ready = (v >> 34) == (unsigned long long)epoch;
                if (!__all_sync(0xffffffffu, ready)) continue;              // some predecessor not published yet
                const unsigned incl = __ballot_sync(0xffffffffu, ((v >> 32) & 3ull) == 2ull);
                const int stop = incl ? __ffs(incl) - 1 : 31;               // nearest predecessor with an inclusive prefix
                prefix += __reduce_add_sync(0xffffffffu, lane <= stop ? (int)(unsigned)(v & 0xFFFFFFFFull) : 0);
                if (incl) break;
                hi -= 32;
            }
            if (lane == 0) st_volatile_u64(&status[tile], tag | (2ull << 32) | (unsigned)(prefix + total));
        }
        if (lane == 0) {
            s_prefix = prefix;
            if (tile == ntiles - 1) *n_out = prefix + total;
        }
    }
    __syncthreads();
    if (f) {
        const int pos = s_prefix + ex;
        out[pos] = m;
        if (g.kp1) {         // fused KeyPoint::convert on both sides (same bounds rule as gather_matches_kernel)
            const float2 a = (m.queryIdx >= 0 && m.queryIdx < g.nkp1) ? reinterpret_cast<const float2 *>(g.kp1)[m.queryIdx] : make_float2(0.f, 0.f);
            const float2 b = (m.trainIdx >= 0 && m.trainIdx < g.nkp2) ? reinterpret_cast<const float2 *>(g.kp2)[m.trainIdx] : make_float2(0.f, 0.f);
            reinterpret_cast<float2 *>(g.p1)[pos] = a;
            reinterpret_cast<float2 *>(g.p2)[pos] = b;
            if (g.pts4) reinterpret_cast<float4 *>(g.pts4)[pos] = make_float4(a.x, a.y, b.x, b.y);
        }
    }
    pm_chain_signal(chain_done, chain_ctr, chain_seq);
    pm_span_mark(span, 14, true);
}

template <class Pred>
int run_compact(pm_ctx *ctx, Pred pred, int n, pm_dmatch *dout, int32_t *dn_out, unsigned long long *chain_done = nullptr,
                unsigned *chain_ctr = nullptr, unsigned long long chain_seq = 0, const pm_gather_out *gather = nullptr)
{
    const pm_gather_out g = gather ? *gather : pm_gather_out{nullptr, 0, nullptr, 0, nullptr, nullptr, nullptr};
    if (n <= 0) {
        PM_CUDA(ctx, cudaMemsetAsync(dn_out, 0, sizeof(int32_t), ctx->stream));
        return PM_OK;
    }
    const int nb = pm_cdiv(n, FB);
    const size_t need = (size_t)(nb + 2) * 8;
    const bool fresh = ctx->slot_bytes[WS_COUNT] < need;
    PM_WS(ctx, st, unsigned long long *, WS_COUNT, need);
    if (fresh) {        // newly (re)allocated: all epochs 0, ticket counter 0
        PM_CUDA(ctx, cudaMemsetAsync(st, 0, ctx->slot_bytes[WS_COUNT], ctx->stream));
        ctx->compact_epoch = 0;
        ctx->compact_tickets = 0;
    }
    const unsigned epoch = ++ctx->compact_epoch;
    if (epoch >= (1u << 29)) {   // keep the 30-bit tag from wrapping into a stale match
        PM_CUDA(ctx, cudaMemsetAsync(st, 0, ctx->slot_bytes[WS_COUNT], ctx->stream));
        ctx->compact_epoch = 0;
        ctx->compact_tickets = 0;
        return run_compact(ctx, pred, n, dout, dn_out, chain_done, chain_ctr, chain_seq, gather);
    }
    const unsigned long long base = ctx->compact_tickets;          // st[0]: the ticket counter of this ctx, never reset
    const bool after_memset = base == 0;                            // the counter was just cleared by a memset on the stream
    ctx->compact_tickets += (unsigned long long)nb;
    if (after_memset) {
        // the tickets are taken BEFORE griddepcontrol.wait: the first launch after the memset must not start early (a
        // programmatic launch could take a ticket before the memset's stores are visible) -> plain stream order
        compact_lookback_kernel<Pred><<<nb, FB, 0, ctx->stream>>>(pred, n, dout, dn_out, st + 1, st, base, epoch, g_pm_span,
                                                                   chain_done, chain_ctr, chain_seq, g);
    } else {
        PM_CUDA(ctx, pm_launch_pdl(compact_lookback_kernel<Pred>, dim3(nb), dim3(FB), 0, ctx->stream, pred, n, dout, dn_out, st + 1,
                                   st, base, epoch, g_pm_span, chain_done, chain_ctr, chain_seq, g));
    }
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

// main.cpp:49-56: minMatch = 1, maxMatch = 0, then a running min/max of `distance`.
// Distances are non-negative floats, so their bit patterns order like unsigned ints.
__global__ void minmax_init_kernel(unsigned *mm) { mm[0] = __float_as_uint(1.0f); mm[1] = __float_as_uint(0.0f); }
__global__ void minmax_reduce_kernel(const pm_dmatch *m, int n, int stride, unsigned *mm)
{
    unsigned lo = __float_as_uint(1.0f), hi = 0u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned b = __float_as_uint(m[(size_t)i * stride].distance);
        lo = min(lo, b); hi = max(hi, b);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[0], lo); atomicMax(&mm[1], hi); }
}
__global__ void minmax_export_kernel(const unsigned *mm, double *out)
{
    out[0] = (double)__uint_as_float(mm[0]);
    out[1] = (double)__uint_as_float(mm[1]);
}

__global__ void gather_points_kernel(const float2 *__restrict__ kp, int nkp, const int32_t *__restrict__ idx,
                                     int n, float2 *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = idx[i];
    out[i] = (k >= 0 && k < nkp) ? kp[k] : make_float2(0.f, 0.f);
}

// pts4 (optional): the same correspondences once more as {x1, y1, x2, y2}, the layout the scoring / mask / refit
// kernels read -- saves the pack launches of the pair pipeline
__global__ void gather_matches_kernel(const pm_dmatch *__restrict__ m, const int32_t *__restrict__ n_ptr,
                                      int max_matches, const float2 *__restrict__ kp1, int nkp1,
                                      const float2 *__restrict__ kp2, int nkp2,
                                      float2 *__restrict__ p1, float2 *__restrict__ p2, float4 *__restrict__ pts4)
{
    const int n = min(*n_ptr, max_matches);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const pm_dmatch d = m[i];
    const float2 a = (d.queryIdx >= 0 && d.queryIdx < nkp1) ? kp1[d.queryIdx] : make_float2(0.f, 0.f);
    const float2 b = (d.trainIdx >= 0 && d.trainIdx < nkp2) ? kp2[d.trainIdx] : make_float2(0.f, 0.f);
    p1[i] = a; p2[i] = b;
    if (pts4) pts4[i] = make_float4(a.x, a.y, b.x, b.y);
}

// ---- cross-check: only the train rows that are somebody's best match need a column minimum ----
__global__ void cross_mark_kernel(const pm_dmatch *__restrict__ knn, int nq, int stride, int nt, uint8_t *__restrict__ mark)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const int j = knn[(size_t)i * stride].trainIdx;
    if (j >= 0 && j < nt) mark[j] = 1;
}

// marked rows -> list (unordered: one warp-aggregated atomic per 32 rows)
__global__ void cross_list_kernel(const uint8_t *__restrict__ mark, int nt, int *__restrict__ list, int *__restrict__ count)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool m = j < nt && mark[j] != 0;
    const unsigned b = __ballot_sync(0xffffffffu, m);
    if (!b) return;
    const int lane = threadIdx.x & 31, leader = __ffs(b) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(b));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (m) list[base + __popc(b & ((1u << lane) - 1u))] = j;
}

template <typename V>
__global__ void cross_gather_rows_kernel(const V *__restrict__ src, int row_elems, const int *__restrict__ list,
                                         const int *__restrict__ count, long long total, V *__restrict__ dst)
{
    const int live = *count;                       // rows past the listed ones (the launch is sized by a bound) become zero rows
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / row_elems), w = (int)(idx - (long long)r * row_elems);
        V v = V();
        if (r < live) v = src[(size_t)list[r] * row_elems + w];
        dst[idx] = v;
    }
}

__global__ void cross_scatter_kernel(const int *__restrict__ list, const int *__restrict__ count, int n,
                                     const unsigned long long *__restrict__ small_, unsigned long long *__restrict__ col)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && k < *count) col[list[k]] = small_[k];
}

}  // namespace

int pmk_ratio_filter(pm_ctx *ctx, const pm_dmatch *dknn, int nq, float ratio, pm_dmatch *dout, int32_t *dn_out,
                     const pm_gather_out *gather)
{
    return run_compact(ctx, RatioPred{dknn, ratio}, nq, dout, dn_out, nullptr, nullptr, 0, gather);
}

int pmk_ratio_filter_tail(pm_ctx *ctx, const pm_dmatch *dknn, int nq, float ratio, pm_dmatch *dout, int32_t *dn_out,
                          unsigned long long *chain_done, unsigned *chain_ctr, unsigned long long seq, const pm_gather_out *gather)
{
    return run_compact(ctx, RatioPred{dknn, ratio}, nq, dout, dn_out, chain_done, chain_ctr, seq, gather);
}

int pmk_cross_mark(pm_ctx *ctx, const pm_dmatch *dknn, int nq, int stride, int nt, uint8_t *dmark)
{
    PM_CUDA(ctx, cudaMemsetAsync(dmark, 0, (size_t)nt, ctx->stream));
    if (nq > 0) {
        cross_mark_kernel<<<pm_cdiv(nq, 256), 256, 0, ctx->stream>>>(dknn, nq, stride, nt, dmark);
        PM_CHECK_LAUNCH(ctx);
    }
    return PM_OK;
}

int pmk_cross_list(pm_ctx *ctx, const uint8_t *dmark, int nt, int32_t *dlist, int32_t *dcount)
{
    PM_CUDA(ctx, cudaMemsetAsync(dcount, 0, 4, ctx->stream));
    cross_list_kernel<<<pm_cdiv(nt, 256), 256, 0, ctx->stream>>>(dmark, nt, dlist, dcount);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

int pmk_cross_gather_rows(pm_ctx *ctx, const void *dsrc, size_t row_bytes, const int32_t *dlist, const int32_t *dcount, int n,
                          void *ddst)
{
    if (n <= 0) return PM_OK;
    const uintptr_t al = (uintptr_t)dsrc | (uintptr_t)ddst | (uintptr_t)row_bytes;
    const int blocks = 8 * ctx->num_sms;
    if ((al & 15) == 0)
        cross_gather_rows_kernel<uint4><<<blocks, 256, 0, ctx->stream>>>((const uint4 *)dsrc, (int)(row_bytes / 16), dlist, dcount,
                                                                           (long long)n * (long long)(row_bytes / 16), (uint4 *)ddst);
    else if ((al & 3) == 0)
        cross_gather_rows_kernel<uint32_t><<<blocks, 256, 0, ctx->stream>>>((const uint32_t *)dsrc, (int)(row_bytes / 4), dlist, dcount,
                                                                              (long long)n * (long long)(row_bytes / 4), (uint32_t *)ddst);
    else
        cross_gather_rows_kernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>((const uint8_t *)dsrc, (int)row_bytes, dlist, dcount,
                                                                             (long long)n * (long long)row_bytes, (uint8_t *)ddst);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

int pmk_cross_scatter(pm_ctx *ctx, const int32_t *dlist, const int32_t *dcount, int n, const uint64_t *dsmall, uint64_t *dcol_best,
                      int nt)
{
    PM_CUDA(ctx, cudaMemsetAsync(dcol_best, 0xFF, (size_t)nt * 8, ctx->stream));      // never-marked rows: no candidate
    if (n > 0) {
        cross_scatter_kernel<<<pm_cdiv(n, 256), 256, 0, ctx->stream>>>(dlist, dcount, n, (const unsigned long long *)dsmall,
                                                                         (unsigned long long *)dcol_best);
        PM_CHECK_LAUNCH(ctx);
    }
    return PM_OK;
}

int pmk_cross_check(pm_ctx *ctx, const pm_dmatch *dknn, int nq, int stride, const uint64_t *dcol_best, int nt,
                    pm_dmatch *dout, int32_t *dn_out)
{
    return run_compact(ctx, CrossPred{dknn, stride, (const unsigned long long *)dcol_best, nt}, nq, dout, dn_out);
}

int pmk_minmax_filter(pm_ctx *ctx, const pm_dmatch *dm, int n, int stride, pm_dmatch *dout, int32_t *dn_out,
                      double *dminmax)
{
    PM_WS(ctx, mm, unsigned *, WS_MISC, 64);
    minmax_init_kernel<<<1, 1, 0, ctx->stream>>>(mm);
    PM_CHECK_LAUNCH(ctx);
    if (n > 0) {
        const int blocks = min(pm_cdiv(n, 256), 4 * ctx->num_sms);
        minmax_reduce_kernel<<<blocks, 256, 0, ctx->stream>>>(dm, n, stride, mm);
        PM_CHECK_LAUNCH(ctx);
    }
    if (dminmax) {
        minmax_export_kernel<<<1, 1, 0, ctx->stream>>>(mm, dminmax);
        PM_CHECK_LAUNCH(ctx);
    }
    return run_compact(ctx, MinMaxPred{dm, stride, mm}, n, dout, dn_out);
}

int pmk_gather_points(pm_ctx *ctx, const float *dkp, int nkp, const int32_t *didx, int n, float *dout)
{
    if (n <= 0) return PM_OK;
    gather_points_kernel<<<pm_cdiv(n, 256), 256, 0, ctx->stream>>>((const float2 *)dkp, nkp, didx, n, (float2 *)dout);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

int pmk_gather_matches(pm_ctx *ctx, const pm_dmatch *dm, const int32_t *dn, int max_matches, const float *dkp1,
                       int nkp1, const float *dkp2, int nkp2, float *dp1, float *dp2, float *dpts4)
{
    if (max_matches <= 0) return PM_OK;
    gather_matches_kernel<<<pm_cdiv(max_matches, 256), 256, 0, ctx->stream>>>(
        dm, dn, max_matches, (const float2 *)dkp1, nkp1, (const float2 *)dkp2, nkp2, (float2 *)dp1, (float2 *)dp2, (float4 *)dpts4);
    PM_CHECK_LAUNCH(ctx);
    return PM_OK;
}

int pm_hang_init_filter(pm_hang_rec *dev_view)
{
    return cudaMemcpyToSymbol(g_pm_hang_rec, &dev_view, sizeof(dev_view)) == cudaSuccess ? PM_OK : PM_CUDA_ERR;
}
