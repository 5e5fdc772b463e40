// pm_internal.h -- shared internals of libpm (context, workspace arena, launch helpers).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/pm.h"

#define PM_NSLOTS 48
#define PM_PROF_RING 4096
#define PM_MAX_LANES 8   /* measured, us per cfg5 pair (every lane warmed up first): 1 / 2 / 4 / 6 / 8 lanes = 207 / 130 / 86 / 71 / 67 */

// Workspace slots (one growable device buffer each).
enum pm_slot {
    WS_Q_RAW = 0, WS_T_RAW, WS_OUT, WS_OUT2, WS_COUNT,
    WS_Q_PACK, WS_T_PACK, WS_Q_NORM, WS_T_NORM, WS_L2_PART, WS_L2_FLAGS, WS_L2_FLAGGED,
    WS_HAM_Q, WS_HAM_T, WS_HAM_PART, WS_COLBEST,
    WS_P1, WS_P2, WS_SAMPLES, WS_F32, WS_COUNTS, WS_KEY, WS_MASK, WS_FOUT, WS_REFIT, WS_MISC,
    WS_KNN, WS_KNN2, WS_IDX, WS_KP, WS_LINES,
    WS_Q_U8, WS_T_U8, WS_T_NORMF, WS_L2_FBPART, WS_SHARD, WS_KP2, WS_PAIRRES, WS_L2_SCHED, WS_HAND,
    WS_X_MARK, WS_X_LIST, WS_X_ROWS, WS_X_COL
};
static_assert(WS_X_COL < PM_NSLOTS, "workspace slots");
void l2_sched_free(pm_ctx *ctx);        // l2_tc.cu
void pm_comm_release(pm_ctx *ctx);      // pm_nccl.cu: destroys an owned communicator (pm_destroy)

struct pm_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    void *slot_ptr[PM_NSLOTS] = {};
    size_t slot_bytes[PM_NSLOTS] = {};
    uint64_t launches = 0;
    std::string err;
    int32_t l2_stats[4] = {0, 0, 0, 0};
    unsigned compact_epoch = 0;    // filter.cu: tag of the current compaction call
    unsigned long long compact_tickets = 0;   // filter.cu: tiles handed out by all compaction calls so far (ticket base of the next)
    int l2_rot = 0;                // l2.cu: which of the three rotating L2Flags blocks the next call uses
    // opt-in chain pipelining (pm_set_pipelining, l2.cu): consecutive one-call kNN-2 + ratio chains overlap
    int pipelining = 0;
    bool tail_is_chain = false;    // the last kernel enqueued through this ctx is the tail of signalling chain `chain_seq`
    unsigned long long chain_seq = 0;   // signalling chains enqueued so far (the tail of chain s stores s to chain_done)
    int chain_shape[4] = {0, 0, 0, 0};  // nq, nt, dim, is_u8 of the last signalling chain
    void *l2_sched = nullptr;      // l2_tc.cu: K2's work partitions (small cache keyed by shape)
    void *tmap_encode = nullptr;   // cuTensorMapEncodeTiled entry point
    // cached TMA descriptors (l2_tc.cu): [2 * set + 0] query operand, [2 * set + 1] train operand
    alignas(64) unsigned char tmap_store[4][128] = {};
    const void *tmap_base[4] = {nullptr, nullptr, nullptr, nullptr};
    int tmap_rows[4] = {0, 0, 0, 0};
    int tmap_fp8[4] = {-1, -1, -1, -1};
    int32_t *h_pinned = nullptr;   // 4 KB pinned scratch for small D2H reads
    // batched pair pipeline (pm_match_estimate_batched_dev): child contexts (own stream + workspaces) so that
    // the latency-bound kernels of one pair overlap the other pairs' work
    pm_ctx *lane[PM_MAX_LANES] = {};
    cudaEvent_t ev_lane[PM_MAX_LANES] = {};
    cudaEvent_t ev_fork = nullptr;
    int batch_lanes = 4;
    // chunked host path (pm_api.cu): uploads run on their own stream, one event per query chunk
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_fence = nullptr, ev_train = nullptr, ev_chunk[8] = {};
    // multi-GPU (pm_nccl.cu): ncclComm_t, owned (pm_comm_init) or borrowed (pm_set_comm)
    void *nccl_comm = nullptr;
    bool comm_owned = false;
    int n_ranks = 1, rank = 0;
    int ham_path = 0;              // per-ctx override of the Hamming kernel choice is the process-global debug hook for now
    // optional device-side kernel timing (pm_profile_*): ring of event pairs per kernel class
    bool profile = false;
    cudaEvent_t prof_ev[3][PM_PROF_RING][2] = {};
    int prof_n[3] = {0, 0, 0};
    bool prof_alloc = false;
};

// Brackets one launch with events when profiling is on (which: 0 K2, 1 K4, 2 K7).
struct pm_prof_scope {
    pm_ctx *c; int which; int slot;
    pm_prof_scope(pm_ctx *ctx, int w) : c(ctx), which(w), slot(-1) {
        if (c->profile && c->prof_n[w] < PM_PROF_RING) {
            slot = c->prof_n[w]++;
            cudaEventRecord(c->prof_ev[w][slot][0], c->stream);
        }
    }
    ~pm_prof_scope() { if (slot >= 0) cudaEventRecord(c->prof_ev[which][slot][1], c->stream); }
};

int pm_fail(pm_ctx *ctx, int status, const char *fmt, ...);
void *pm_ws(pm_ctx *ctx, int slot, size_t bytes);   // nullptr on failure (ctx->err set)

#define PM_CUDA(ctx, call)                                                              \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess)                                                         \
            return pm_fail(ctx, PM_CUDA_ERR, "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                           cudaGetErrorString(e__));                                    \
    } while (0)

#define PM_CHECK_LAUNCH(ctx)                                                            \
    do {                                                                                \
        (ctx)->launches++;                                                              \
        (ctx)->tail_is_chain = false;                                                   \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess)                                                         \
            return pm_fail(ctx, PM_CUDA_ERR, "%s:%d launch: %s", __FILE__, __LINE__,    \
                           cudaGetErrorString(e__));                                    \
    } while (0)

#define PM_WS(ctx, var, type, slot, bytes)                                              \
    type var = (type)pm_ws(ctx, slot, bytes);                                           \
    if (!var) return PM_CUDA_ERR

// Debug timeline (tools/step_timeline.py): when set, every kernel of the L2 chain records
// span[3k+0] = min CTA entry, span[3k+1] = min time past griddepcontrol.wait, span[3k+2] = max CTA exit
// (%globaltimer, ns).  nullptr in normal operation.
extern unsigned long long *g_pm_span;

// Hang detector: every device-side wait in libpm (mbarrier waits of K2, the look-back of K5, the chain waits) counts its
// polls; a wait that exceeds PM_SPIN_LIMIT polls (seconds: far beyond any legitimate wait) writes what it was waiting for
// into a host-mapped record and traps, so a protocol fault surfaces as PM_CUDA_ERR with a message instead of a silent
// hang.  g_pm_hang_rec (one per translation unit, set by pm_hang_init_<tu> at pm_create) points at mapped host memory.
struct pm_hang_rec { unsigned code, a, b, c; };
#define PM_SPIN_LIMIT (1u << 27)
#define PM_WAIT_LIMIT_NS 3000000000ull     /* 3 s on %globaltimer */
extern pm_hang_rec *pm_hang_host;       // the host view of the record (pm_api.cu); nullptr until the first pm_create
int pm_hang_init_filter(pm_hang_rec *dev_view);
int pm_hang_init_l2(pm_hang_rec *dev_view);
int pm_hang_init_l2_tc(pm_hang_rec *dev_view);
#ifdef __CUDACC__
static __device__ pm_hang_rec *g_pm_hang_rec = nullptr;
static __device__ __forceinline__ unsigned long long pm_now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
static __device__ __noinline__ void pm_hang_trap(unsigned code, unsigned a, unsigned b, unsigned c)
{
    pm_hang_rec *r = g_pm_hang_rec;
    if (r && atomicCAS_system(&r->code, 0u, code) == 0u) {      // the first waiter to give up describes itself
        r->a = a; r->b = b; r->c = c;
        __threadfence_system();
        // give the other waiters of this kernel a moment to reach their own limit before the context dies
        const unsigned long long t0 = pm_now_ns();
        while (pm_now_ns() - t0 < 2000000ull) { }
    }
    __trap();
}
#endif

// Programmatic dependent launch (PDL): every kernel of a chain is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and begins with pm_pdl_prologue():
// it lets the NEXT kernel of the stream start launching right away (its CTAs become
// resident and run their own prologue), then waits until the PREVIOUS kernel has fully
// completed and flushed.  Data order is unchanged; only launch latency overlaps.
#ifdef __CUDACC__
__device__ __forceinline__ void pm_span_mark(unsigned long long *span, int slot, bool is_max)
{
    if (span && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (is_max) atomicMax(span + slot, t); else atomicMin(span + slot, t);
    }
}
__device__ __forceinline__ void pm_pdl_prologue()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// Chain pipelining (pm_set_pipelining): the tail kernel of signalling chain s stores s to *chain_done once its
// last block is done; a kernel of a later chain that may run ahead of its stream predecessors spins here
// (one thread; the caller follows with __syncthreads) until chain `seq` has completed.  Every CTA of the
// awaited chain was started before the spinning kernel could launch (PDL launches cascade in stream
// order), so the spin cannot starve it.
__device__ __forceinline__ void pm_chain_wait(const unsigned long long *chain_done, unsigned long long seq)
{
    if (chain_done && threadIdx.x == 0) {
        unsigned long long v;
        unsigned spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(chain_done) : "memory");
            if (++spins > (1u << 24)) pm_hang_trap(0x10u, (unsigned)seq, (unsigned)v, blockIdx.x);
        } while (v < seq);
    }
}
// tail of a chain: the last block to arrive publishes the sequence number (ctr returns to 0)
__device__ __forceinline__ void pm_chain_signal(unsigned long long *chain_done, unsigned *ctr, unsigned long long seq)
{
    if (!chain_done) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ctr, 1u) == gridDim.x - 1) {
            *ctr = 0u;
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(chain_done), "l"(seq) : "memory");
        }
    }
}
template <typename... KArgs, typename... Args>
static inline cudaError_t pm_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                        cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

static inline int pm_cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int pm_round_up(int a, int b) { return pm_cdiv(a, b) * b; }

// Optional gather fused into the ratio filter's scatter (pair pipeline): survivor i also leaves as the two
// keypoint coordinates (KeyPoint::convert, main.cpp:89-91) and as the packed {x1, y1, x2, y2}.
struct pm_gather_out {
    const float *kp1; int nkp1; const float *kp2; int nkp2; float *p1; float *p2; float *pts4;
};

// ---- kernels-by-file entry points (host launchers) --------------------------------
// hamming.cu
int pmk_hamming_knn2(pm_ctx *ctx, const uint8_t *dq, int nq, const uint8_t *dt, int nt, int bytes,
                     int q_index_base, pm_dmatch *dout);
int pmk_hamming_col_best(pm_ctx *ctx, const uint8_t *dq, int nq, const uint8_t *dt, int nt, int bytes,
                         int q_index_base, uint64_t *dcol_best);
// filter.cu
int pmk_ratio_filter(pm_ctx *ctx, const pm_dmatch *dknn, int nq, float ratio, pm_dmatch *dout,
                     int32_t *dn_out, const pm_gather_out *gather = nullptr);
// same, as the tail of signalling chain `seq` (stores seq to chain_done when the last block is done)
int pmk_ratio_filter_tail(pm_ctx *ctx, const pm_dmatch *dknn, int nq, float ratio, pm_dmatch *dout,
                          int32_t *dn_out, unsigned long long *chain_done, unsigned *chain_ctr, unsigned long long seq,
                          const pm_gather_out *gather = nullptr);
// Cross-check, column side restricted to the train rows that ARE somebody's best match (filter.cu): the rows of the
// forward result mark their best train row; the marked rows are listed (any order), gathered into a compact row set and,
// after the reverse pass over that set, their packed minima are scattered back into the [nt] column array.
int pmk_cross_mark(pm_ctx *ctx, const pm_dmatch *dknn, int nq, int stride, int nt, uint8_t *dmark);
int pmk_cross_list(pm_ctx *ctx, const uint8_t *dmark, int nt, int32_t *dlist, int32_t *dcount);
int pmk_cross_gather_rows(pm_ctx *ctx, const void *dsrc, size_t row_bytes, const int32_t *dlist, const int32_t *dcount, int n,
                          void *ddst);        // n: launch size (a bound); rows past *dcount are written as zeros
int pmk_cross_scatter(pm_ctx *ctx, const int32_t *dlist, const int32_t *dcount, int n, const uint64_t *dsmall, uint64_t *dcol_best,
                      int nt);
// The column side of a cross-check for this rank's query shard (pm_api.cu).  reduce_marks (may be null: one rank) makes the
// mark bytes the union over the ranks, in place, on the ctx stream.  One rank: nothing is synchronised (the reverse pass is
// sized by the bound min(nq, nt)); several ranks: the stream is synchronised once to read the 4-byte count of marked rows.
typedef int (*pm_mark_reduce_fn)(pm_ctx *ctx, uint8_t *dmark, size_t n);
int pmk_cross_col_best(pm_ctx *ctx, int hamming, const void *dq, int nq, const void *dt, int nt, int width, int q_index_base,
                       const pm_dmatch *dknn, uint64_t *dcol_best, pm_mark_reduce_fn reduce_marks);
int pmk_cross_check(pm_ctx *ctx, const pm_dmatch *dknn, int nq, int stride, const uint64_t *dcol_best,
                    int nt, pm_dmatch *dout, int32_t *dn_out);
int pmk_minmax_filter(pm_ctx *ctx, const pm_dmatch *dm, int n, int stride, pm_dmatch *dout,
                      int32_t *dn_out, double *dminmax);
int pmk_gather_points(pm_ctx *ctx, const float *dkp, int nkp, const int32_t *didx, int n, float *dout);
int pmk_gather_matches(pm_ctx *ctx, const pm_dmatch *dm, const int32_t *dn, int max_matches,
                       const float *dkp1, int nkp1, const float *dkp2, int nkp2, float *dp1, float *dp2,
                       float *dpts4 = nullptr);
// l2.cu / l2_tc.cu
int pmk_l2_knn2(pm_ctx *ctx, const void *dq, int nq, const void *dt, int nt, int dim, int is_u8,
                int q_index_base, pm_dmatch *dout);
int pmk_l2_knn2_phase(pm_ctx *ctx, const void *dq, int nq, const void *dt, int nt, int dim, int is_u8,
                      int q_index_base, pm_dmatch *dout, int phase);
int pmk_l2_knn2_fused(pm_ctx *ctx, const void *dq, int nq, const void *dt, int nt, int dim, int is_u8,
                      int q_index_base, pm_dmatch *dout, int phase, float ratio, pm_dmatch *dgood, int32_t *dn_good,
                      const pm_gather_out *gather = nullptr);
int pmk_l2_col_best(pm_ctx *ctx, const float *dq, int nq, const float *dt, int nt, int dim,
                    int q_index_base, uint64_t *dcol_best);
// ransac.cu
// dn (optional, device): the point count is read from *dn by the kernels; n then only bounds it
int pmk_ransac_solve(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const int32_t *dsamples,
                     int n_hyp, int m, float *dF32, const int32_t *dn = nullptr, double *dF64 = nullptr);
int pmk_ransac_score(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const float *dF32,
                     int n_models, float thr, int metric, int32_t *dcounts, const int32_t *dn = nullptr,
                     const float *dpts4 = nullptr);
int pmk_ransac_best_pick(pm_ctx *ctx, const int32_t *dcounts, int n_models, int id_base, const float *dF32, uint64_t *dkey,
                         float *dFw, int32_t *dn_inl);
int pmk_ransac_best(pm_ctx *ctx, const int32_t *dcounts, int n_models, int id_base, uint64_t *dkey);
int pmk_ransac_finish(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const float *dFw,
                      float thr, int metric, int refit, double *dF, uint8_t *dmask, int32_t *dn_inl,
                      const int32_t *dn = nullptr, const float *dpts4 = nullptr, int ninl_is_zero = 0,
                      const uint64_t *dkey = nullptr, int sample_size = 0, pm_pair_result *dres = nullptr);
int pmk_sample_sets(pm_ctx *ctx, int n_points, int n_hyp, int m, uint64_t seed, int32_t *dout,
                    const int32_t *dn = nullptr, int h_base = 0);
// a group of image pairs whose RANSAC kernels run as one launch each (ransac.cu); slot b = pair b of the group:
//   p1 / p2 [b][nmax] float2, pts4 [b][nmax] float4, samples [b][n_hyp * m], F32 [b][(n_models + 1) * 12] (winner last),
//   counts [b][n_models], key [b][8] u64 (key, -, {n_good, n_inl} as ints at +2), mask [b][nmax], refit [b][RF workspace],
//   Fout [b][16] f64, res [b]
struct pm_pair_group {
    int n_pairs, nmax, n_hyp, m, metric, refit_on;
    float threshold;
    uint64_t seed0;
    const float2 *p1, *p2; const float4 *pts4;
    int32_t *samples; float *F32; int32_t *counts; uint64_t *key; uint8_t *mask; double *refit; double *Fout;
    pm_pair_result *res;
};
#define PM_REFIT_WS_DOUBLES (64 * (5 + 2 + 45) + 16)
int pmk_pair_group_ransac(pm_ctx *ctx, const pm_pair_group &G);
int pmk_pair_result(pm_ctx *ctx, const uint64_t *dkey, const int32_t *dn_good, const int32_t *dn_inl, const double *dF,
                    int n_max, int m, pm_pair_result *dres);
int pmk_ransac_pick(pm_ctx *ctx, const uint64_t *dkey, const float *dF32, int id_base, int n_models, float *dFw);
int pmk_ransac_update_best(pm_ctx *ctx, const uint64_t *dkey, const float *dF32, int id_base, int n_models, uint64_t *dbest_key,
                           float *dFw_best);
int pmk_ransac_winner_resolve(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const uint64_t *dkey,
                              const int32_t *dsamples_full, uint64_t seed, int m, int32_t *dwin_idx, float *dF3, float *dFw);
int pmk_lmeds_score(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const float *dF32, int n_models, float *dmedians);
int pmk_lmeds_finish(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const float *dF32, const float *dmedians,
                     int n_models, float *dFw, double *dF, uint8_t *dmask, int32_t *dn_inl, uint64_t *dkey);
int pmk_fundamental_npoint(pm_ctx *ctx, const float *dp1, const float *dp2, int n,
                           const uint8_t *dmask_or_null, double *dF, int32_t *dok);
int pmk_epilines(pm_ctx *ctx, const float *dpts, int n, int which, const double *dF, float *dlines);
int pmk_residuals(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const double *dF, int metric,
                  float *dout, double *dsum);
