// pm_nccl.cu -- the multi-GPU entry points of include/pm.h: NCCL from inside the C ABI, so that a C++ host can run the
// sharded forms of the reference's two calls (BASELINE.json north_star: "query descriptors shard by rows and
// RANSAC hypotheses shard by batch; an NCCL max-allreduce picks the winning model"):
//   BFMatcher(crossCheck=true).match (main.cpp:43-46)  -> pm_match_cross_sharded_dev   one ncclAllReduce(min, u64), Nt x 8 B
//   cv::findFundamentalMat(RANSAC)  (main.cpp:95-98)  -> pm_find_fundamental_sharded_dev one ncclAllReduce(max, u64), 8 B
// The collectives are enqueued on the ctx stream between the kernels they connect.  libnccl.so.2 is loaded with dlopen on
// first use (in a process that imported torch this resolves to the copy torch already loaded); nothing here is needed,
// or loaded, by single-GPU callers.
#include <dlfcn.h>
#include <mutex>
#include "pm_internal.h"

#if __has_include(<nccl.h>)
#include <nccl.h>
#else   // the handful of NCCL declarations used here (stable ABI since NCCL 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5 } ncclDataType_t;
#endif
static_assert(sizeof(ncclUniqueId) == PM_COMM_ID_BYTES, "ncclUniqueId size");

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

NcclApi &nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.why = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return; }
        auto sym = [&](const char *name) -> void * {
            void *p = dlsym(api.handle, name);
            if (!p && api.why.empty()) api.why = std::string("libnccl: missing symbol ") + name;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return api;
}

int nccl_ready(pm_ctx *ctx, NcclApi **out)
{
    NcclApi &a = nccl_api();
    if (!a.why.empty()) return pm_fail(ctx, PM_NCCL_ERR, "%s", a.why.c_str());
    *out = &a;
    return PM_OK;
}

#define PM_NCCL(ctx, api, call)                                                                           \
    do {                                                                                                  \
        ncclResult_t r__ = (call);                                                                        \
        if (r__ != ncclSuccess)                                                                           \
            return pm_fail(ctx, PM_NCCL_ERR, "%s:%d %s: %s", __FILE__, __LINE__, #call, (api)->GetErrorString(r__)); \
    } while (0)

// keys of rows this rank does not own must lose the max: nothing to do (0).  Column minima of an empty shard are ~0
// (all ones) and lose the unsigned min: nothing to do either -- ncclUint64 makes both reductions exact as they are.

__global__ void fill_u64_kernel(unsigned long long *p, int n, unsigned long long v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// fixed-width block for the all-gather: rows past the local count are zeroed so the gathered buffer is deterministic
__global__ void pad_matches_kernel(const pm_dmatch *__restrict__ src, const int32_t *__restrict__ n_src, int width,
                                   pm_dmatch *__restrict__ dst, int32_t *__restrict__ cnt_dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(max(*n_src, 0), width);
    if (i == 0) *cnt_dst = n;
    if (i < width) dst[i] = i < n ? src[i] : pm_dmatch{0, 0, 0, 0.f};
}

}  // namespace

void pm_comm_release(pm_ctx *ctx)
{
    if (ctx->nccl_comm && ctx->comm_owned) {
        NcclApi &a = nccl_api();
        if (a.CommDestroy) a.CommDestroy((ncclComm_t)ctx->nccl_comm);
    }
    ctx->nccl_comm = nullptr; ctx->comm_owned = false; ctx->n_ranks = 1; ctx->rank = 0;
}

extern "C" {

int pm_comm_unique_id(void *id)
{
    if (!id) return PM_BAD_ARG;
    NcclApi &a = nccl_api();
    if (!a.why.empty()) return PM_NCCL_ERR;
    return a.GetUniqueId(reinterpret_cast<ncclUniqueId *>(id)) == ncclSuccess ? PM_OK : PM_NCCL_ERR;
}

int pm_comm_init(pm_ctx *ctx, int n_ranks, int rank, const void *id)
{
    if (!ctx) return PM_BAD_ARG;
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return pm_fail(ctx, PM_BAD_ARG, "pm_comm_init: bad rank / id");
    NcclApi *a;
    int st = nccl_ready(ctx, &a);
    if (st != PM_OK) return st;
    pm_comm_release(ctx);
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    PM_NCCL(ctx, a, a->CommInitRank(&comm, n_ranks, uid, rank));
    ctx->nccl_comm = comm; ctx->comm_owned = true; ctx->n_ranks = n_ranks; ctx->rank = rank;
    return PM_OK;
}

int pm_set_comm(pm_ctx *ctx, void *nccl_comm, int n_ranks, int rank)
{
    if (!ctx) return PM_BAD_ARG;
    pm_comm_release(ctx);
    if (!nccl_comm) return PM_OK;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return pm_fail(ctx, PM_BAD_ARG, "pm_set_comm: bad rank");
    NcclApi *a;
    int st = nccl_ready(ctx, &a);
    if (st != PM_OK) return st;
    ctx->nccl_comm = nccl_comm; ctx->comm_owned = false; ctx->n_ranks = n_ranks; ctx->rank = rank;
    return PM_OK;
}

int pm_comm_info(pm_ctx *ctx, int *n_ranks, int *rank)
{
    if (!ctx) return PM_BAD_ARG;
    if (n_ranks) *n_ranks = ctx->n_ranks;
    if (rank) *rank = ctx->rank;
    return PM_OK;
}

// union of the "is somebody's best match" marks over the ranks (pmk_cross_col_best)
static int reduce_marks_nccl(pm_ctx *ctx, uint8_t *dmark, size_t n)
{
    NcclApi *a = nullptr;
    int st = nccl_ready(ctx, &a);
    if (st != PM_OK) return st;
    PM_NCCL(ctx, a, a->AllReduce(dmark, dmark, n, ncclUint8, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return PM_OK;
}

int pm_match_cross_sharded_dev(pm_ctx *ctx, const void *dq, int nq, const void *dt, int nt, int width, int norm,
                               int q_index_base, pm_dmatch *dknn, uint64_t *dcol_best, pm_dmatch *dout, int32_t *dn_out)
{
    if (!ctx) return PM_BAD_ARG;
    if (nq < 0 || nt < 0 || width <= 0 || !dn_out || (norm != 4 && norm != 6))
        return pm_fail(ctx, PM_BAD_ARG, "pm_match_cross_sharded_dev: bad argument (norm must be 4 = L2 or 6 = Hamming)");
    if ((nq > 0 && (!dq || !dout)) || (nt > 0 && (!dt || !dcol_best)))
        return pm_fail(ctx, PM_BAD_ARG, "pm_match_cross_sharded_dev: null pointer");
    NcclApi *a = nullptr;
    if (ctx->n_ranks > 1) {
        if (!ctx->nccl_comm) return pm_fail(ctx, PM_NCCL_ERR, "no communicator: call pm_comm_init or pm_set_comm first");
        int st = nccl_ready(ctx, &a);
        if (st != PM_OK) return st;
    }
    if (nt == 0) { PM_CUDA(ctx, cudaMemsetAsync(dn_out, 0, 4, ctx->stream)); return PM_OK; }
    if (!dknn && nq > 0) {
        PM_WS(ctx, k, pm_dmatch *, WS_KNN, (size_t)nq * 2 * sizeof(pm_dmatch));
        dknn = k;
    }
    int st;
    if (nq > 0) {
        st = norm == 6 ? pmk_hamming_knn2(ctx, (const uint8_t *)dq, nq, (const uint8_t *)dt, nt, width, q_index_base, dknn)
                       : pmk_l2_knn2(ctx, dq, nq, dt, nt, width, 0, q_index_base, dknn);
        if (st != PM_OK) return st;
    }
    // reverse pass over the train rows some rank's query points at (every rank takes part in the mark exchange, also one
    // with an empty shard: its column minima are all "none")
    st = pmk_cross_col_best(ctx, norm == 6, dq, nq, dt, nt, width, q_index_base, dknn, dcol_best,
                            ctx->n_ranks > 1 ? reduce_marks_nccl : nullptr);
    if (st != PM_OK) return st;
    if (ctx->n_ranks > 1)
        PM_NCCL(ctx, a, a->AllReduce(dcol_best, dcol_best, (size_t)nt, ncclUint64, ncclMin, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return pmk_cross_check(ctx, dknn, nq, 2, dcol_best, nt, dout, dn_out);
}

int pm_allgather_matches_dev(pm_ctx *ctx, const pm_dmatch *dlocal, const int32_t *dn_local, int max_per_rank, pm_dmatch *dall,
                             int32_t *dcounts)
{
    if (!ctx) return PM_BAD_ARG;
    if (max_per_rank <= 0 || !dlocal || !dn_local || !dall || !dcounts)
        return pm_fail(ctx, PM_BAD_ARG, "pm_allgather_matches_dev: bad argument");
    // every rank writes its own padded block in place, then the blocks travel (in-place all-gather)
    pm_dmatch *mine = dall + (size_t)ctx->rank * max_per_rank;
    pad_matches_kernel<<<pm_cdiv(max_per_rank, 256), 256, 0, ctx->stream>>>(dlocal, dn_local, max_per_rank, mine, dcounts + ctx->rank);
    PM_CHECK_LAUNCH(ctx);
    if (ctx->n_ranks == 1) return PM_OK;
    if (!ctx->nccl_comm) return pm_fail(ctx, PM_NCCL_ERR, "no communicator: call pm_comm_init or pm_set_comm first");
    NcclApi *a;
    int st = nccl_ready(ctx, &a);
    if (st != PM_OK) return st;
    PM_NCCL(ctx, a, a->AllGather(mine, dall, (size_t)max_per_rank * sizeof(pm_dmatch), ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    PM_NCCL(ctx, a, a->AllGather(dcounts + ctx->rank, dcounts, 1, ncclInt32, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return PM_OK;
}

int pm_find_fundamental_sharded_dev(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const pm_ransac_params *prm,
                                    int n_hyp_total, double *dF, uint8_t *dmask, int32_t *dn_inliers, uint64_t *dkey)
{
    if (!ctx) return PM_BAD_ARG;
    if (!prm || !dp1 || !dp2 || !dF || !dmask || !dn_inliers || !dkey) return pm_fail(ctx, PM_BAD_ARG, "pm_find_fundamental_sharded_dev: null pointer");
    const int m = prm->sample_size, nh = prm->n_hyp, base = prm->hyp_id_base;
    if ((m != 7 && m != 8) || (prm->metric != PM_METRIC_SAMPSON && prm->metric != PM_METRIC_SYMEPI) || n < m || nh < 0 || base < 0 ||
        n_hyp_total < base + nh)
        return pm_fail(ctx, PM_BAD_ARG, "pm_find_fundamental_sharded_dev: need sample_size 7|8, n >= sample_size, 0 <= base, base + n_hyp <= n_hyp_total");
    NcclApi *a = nullptr;
    int st;
    if (ctx->n_ranks > 1) {
        if (!ctx->nccl_comm) return pm_fail(ctx, PM_NCCL_ERR, "no communicator: call pm_comm_init or pm_set_comm first");
        if ((st = nccl_ready(ctx, &a)) != PM_OK) return st;
    }
    const int per = m == 8 ? 1 : 3;
    const int nhw = nh > 0 ? nh : 1;
    PM_WS(ctx, dF32, float *, WS_F32, ((size_t)nhw * per + 1) * 12 * 4);
    PM_WS(ctx, dcounts, int32_t *, WS_COUNTS, (size_t)nhw * per * 4);
    PM_WS(ctx, dsh, int32_t *, WS_SHARD, 64 + 4 * 12 * 4);
    int32_t *dwin_idx = dsh;                                          // [8]
    float *dF3 = reinterpret_cast<float *>(dsh + 16);                 // [3][12]
    float *dFw = dF32 + (size_t)nhw * per * 12;
    const int32_t *ds = prm->sample_idx ? prm->sample_idx + (size_t)base * m : nullptr;
    if (!ds && nh > 0) {                                              // this shard's slice of the seed's stream
        PM_WS(ctx, gen, int32_t *, WS_SAMPLES, (size_t)nh * m * 4);
        if ((st = pmk_sample_sets(ctx, n, nh, m, prm->seed, gen, nullptr, base)) != PM_OK) return st;
        ds = gen;
    }
    if (nh > 0) {
        if ((st = pmk_ransac_solve(ctx, dp1, dp2, n, ds, nh, m, dF32)) != PM_OK) return st;
        if ((st = pmk_ransac_score(ctx, dp1, dp2, n, dF32, nh * per, prm->threshold, prm->metric, dcounts)) != PM_OK) return st;
    }
    if ((st = pmk_ransac_best(ctx, dcounts, nh * per, base * per, dkey)) != PM_OK) return st;          // 0 for an empty shard
    if (ctx->n_ranks > 1)      // the one exchange: 8 bytes, max count wins, lowest global model id on ties
        PM_NCCL(ctx, a, a->AllReduce(dkey, dkey, 1, ncclUint64, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    // every rank re-solves the winning index set (it knows the whole index array / the seed): no broadcast of F or the mask
    if ((st = pmk_ransac_winner_resolve(ctx, dp1, dp2, n, dkey, prm->sample_idx, prm->seed, m, dwin_idx, dF3, dFw)) != PM_OK) return st;
    return pmk_ransac_finish(ctx, dp1, dp2, n, dFw, prm->threshold, prm->metric, prm->refit, dF, dmask, dn_inliers);
}

}  // extern "C"
