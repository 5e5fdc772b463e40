// pm_api.cu -- the C ABI of libpm (include/pm.h): context, workspace arena, and the
// host-buffer entry points that stand in for the OpenCV calls of
// /root/reference/Points Matching/main.cpp:43-46, 49-69, 89-91, 95-98, 103-132.
// Every path runs CUDA kernels; there is no CPU fallback.
#include <cstdarg>
#include <cstdlib>
#include <vector>
#include <mutex>
#include <thread>
#include <vector>
#include <cfloat>
#include <cmath>
#include <nvtx3/nvToolsExt.h>
#include "pm_internal.h"

// NVTX range per entry point (domain "libpm"): header-only nvtx3, a no-op unless a profiler injects itself.
static nvtxDomainHandle_t pm_nvtx_domain()
{
    static nvtxDomainHandle_t d = nvtxDomainCreateA("libpm");
    return d;
}
struct pm_nvtx_scope {
    explicit pm_nvtx_scope(const char *name)
    {
        nvtxEventAttributes_t a = {};
        a.version = NVTX_VERSION; a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
        a.messageType = NVTX_MESSAGE_TYPE_ASCII; a.message.ascii = name;
        nvtxDomainRangePushEx(pm_nvtx_domain(), &a);
    }
    ~pm_nvtx_scope() { nvtxDomainRangePop(pm_nvtx_domain()); }
};
#define PM_NVTX() pm_nvtx_scope nvtx_scope__(__func__)

pm_hang_rec *pm_hang_host = nullptr;
static int pm_hang_setup()
{
    // one host-mapped record per process (every device can write it); per-device symbol initialisation
    static std::mutex mu;
    static unsigned long long done_devices = 0;
    std::lock_guard<std::mutex> lk(mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return PM_CUDA_ERR;
    if (!pm_hang_host) {
        void *h = nullptr;
        if (cudaHostAlloc(&h, sizeof(pm_hang_rec), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return PM_CUDA_ERR;
        memset(h, 0, sizeof(pm_hang_rec));
        pm_hang_host = static_cast<pm_hang_rec *>(h);
    }
    if (done_devices & (1ull << (dev & 63))) return PM_OK;
    pm_hang_rec *d = nullptr;
    if (cudaHostGetDevicePointer((void **)&d, pm_hang_host, 0) != cudaSuccess) return PM_CUDA_ERR;
    if (pm_hang_init_filter(d) != PM_OK || pm_hang_init_l2(d) != PM_OK || pm_hang_init_l2_tc(d) != PM_OK) return PM_CUDA_ERR;
    done_devices |= 1ull << (dev & 63);
    return PM_OK;
}

unsigned long long *g_pm_span = nullptr;
extern "C" void pm_debug_set_span(unsigned long long *p) { g_pm_span = p; }

int pm_fail(pm_ctx *ctx, int status, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) {
        ctx->err = buf;
        if (status == PM_CUDA_ERR && pm_hang_host && pm_hang_host->code) {
            char more[160];
            snprintf(more, sizeof(more), " [device-side wait gave up: code 0x%x, a=%u b=%u c=%u]", pm_hang_host->code, pm_hang_host->a,
                     pm_hang_host->b, pm_hang_host->c);
            ctx->err += more;
        }
    }
    return status;
}

void *pm_ws(pm_ctx *ctx, int slot, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (ctx->slot_bytes[slot] >= bytes) return ctx->slot_ptr[slot];
    if (ctx->slot_ptr[slot]) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->slot_ptr[slot]);
        ctx->slot_ptr[slot] = nullptr; ctx->slot_bytes[slot] = 0;
    }
    size_t cap = bytes + bytes / 4;            // growth slack
    cap = (cap + 255) & ~(size_t)255;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, cap);
    if (e != cudaSuccess) {
        pm_fail(ctx, PM_CUDA_ERR, "cudaMalloc(%zu) for workspace slot %d: %s", cap, slot, cudaGetErrorString(e));
        return nullptr;
    }
    ctx->slot_ptr[slot] = p; ctx->slot_bytes[slot] = cap;
    return p;
}

extern "C" {

int pm_version(void) { return PM_VERSION; }

int pm_create(pm_ctx **out, int device)
{
    if (!out) return PM_BAD_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return PM_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PM_NO_DEVICE;
    if (prop.major != 10) return PM_NO_DEVICE;          // sm_100a kernels only
    if (cudaSetDevice(device) != cudaSuccess) return PM_CUDA_ERR;
    if (pm_hang_setup() != PM_OK) return PM_CUDA_ERR;
    pm_ctx *ctx = new pm_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PM_CUDA_ERR; }
    ctx->stream = ctx->own_stream;
    if (cudaMallocHost((void **)&ctx->h_pinned, 4096) != cudaSuccess) { cudaStreamDestroy(ctx->own_stream); delete ctx; return PM_CUDA_ERR; }
    *out = ctx;
    return PM_OK;
}

int pm_destroy(pm_ctx *ctx)
{
    if (!ctx) return PM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    pm_comm_release(ctx);
    l2_sched_free(ctx);
    for (int s = 0; s < PM_NSLOTS; ++s) if (ctx->slot_ptr[s]) cudaFree(ctx->slot_ptr[s]);
    if (ctx->prof_alloc)
        for (int w = 0; w < 3; ++w)
            for (int i = 0; i < PM_PROF_RING; ++i)
                for (int k = 0; k < 2; ++k) cudaEventDestroy(ctx->prof_ev[w][i][k]);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->copy_stream) {
        cudaStreamDestroy(ctx->copy_stream);
        cudaEventDestroy(ctx->ev_fence); cudaEventDestroy(ctx->ev_train);
        for (int i = 0; i < 8; ++i) cudaEventDestroy(ctx->ev_chunk[i]);
    }
    for (int k = 0; k < PM_MAX_LANES; ++k) {
        if (ctx->lane[k]) pm_destroy(ctx->lane[k]);
        if (ctx->ev_lane[k]) cudaEventDestroy(ctx->ev_lane[k]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return PM_OK;
}

int pm_set_batch_lanes(pm_ctx *ctx, int lanes)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    if (lanes < 1 || lanes > PM_MAX_LANES) return pm_fail(ctx, PM_BAD_ARG, "pm_set_batch_lanes: 1..%d lanes", PM_MAX_LANES);
    ctx->batch_lanes = lanes;
    return PM_OK;
}

int pm_set_stream(pm_ctx *ctx, void *s)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    cudaStream_t ns = s ? (cudaStream_t)s : ctx->own_stream;
    if (ns != ctx->stream) {
        // the workspaces are shared by everything this ctx enqueues: work still running on the old stream must not
        // meet work enqueued on the new one
        PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->stream = ns;
    }
    ctx->tail_is_chain = false;
    return PM_OK;
}

int pm_set_pipelining(pm_ctx *ctx, int on)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    if ((on != 0) != (ctx->pipelining != 0)) {
        // the workspace layout changes (one buffer set <-> two): drain the stream first
        PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->pipelining = on != 0;
        ctx->tail_is_chain = false;
        for (int k = 0; k < 4; ++k) ctx->tmap_base[k] = nullptr;
    }
    return PM_OK;
}

int pm_sync(pm_ctx *ctx)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PM_OK;
}

int pm_profile_enable(pm_ctx *ctx, int on)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    if (on && !ctx->prof_alloc) {
        for (int w = 0; w < 3; ++w)
            for (int i = 0; i < PM_PROF_RING; ++i)
                for (int k = 0; k < 2; ++k) PM_CUDA(ctx, cudaEventCreate(&ctx->prof_ev[w][i][k]));
        ctx->prof_alloc = true;
    }
    ctx->profile = on != 0;
    for (int w = 0; w < 3; ++w) ctx->prof_n[w] = 0;
    return PM_OK;
}

int pm_profile_read(pm_ctx *ctx, int which, double *total_ms, int *n_launches)
{
    if (!ctx || which < 0 || which > 2) return PM_BAD_ARG;
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double tot = 0;
    const int n = ctx->prof_n[which];
    for (int i = 0; i < n; ++i) {
        float ms = 0;
        PM_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->prof_ev[which][i][0], ctx->prof_ev[which][i][1]));
        tot += ms;
    }
    ctx->prof_n[which] = 0;
    if (total_ms) *total_ms = tot;
    if (n_launches) *n_launches = n;
    return PM_OK;
}

const char *pm_last_error(pm_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t pm_launch_count(pm_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------
// device-resident entry points
// ---------------------------------------------------------------------------------
#define PM_REQUIRE(ctx, cond, msg) do { if (!(cond)) return pm_fail(ctx, PM_BAD_ARG, "%s: %s", __func__, msg); } while (0)

int pm_knn2_l2_f32_dev(pm_ctx *ctx, const float *dq, int nq, const float *dt, int nt, int dim, int base, pm_dmatch *dout)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, nq >= 0 && nt >= 0 && dim > 0, "negative size or dim <= 0");
    PM_REQUIRE(ctx, nq == 0 || (dq && dout), "null pointer");
    return pmk_l2_knn2(ctx, dq, nq, dt, nt, dim, 0, base, dout);
}
int pm_knn2_l2_u8_dev(pm_ctx *ctx, const uint8_t *dq, int nq, const uint8_t *dt, int nt, int dim, int base, pm_dmatch *dout)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, nq >= 0 && nt >= 0 && dim > 0, "negative size or dim <= 0");
    PM_REQUIRE(ctx, nq == 0 || (dq && dout), "null pointer");
    return pmk_l2_knn2(ctx, dq, nq, dt, nt, dim, 1, base, dout);
}
int pm_knn2_ratio_l2_f32_dev(pm_ctx *ctx, const float *dq, int nq, const float *dt, int nt, int dim, float ratio, int base,
                             pm_dmatch *dknn, pm_dmatch *dgood, int32_t *dn_good)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, nq >= 0 && nt >= 0 && dim > 0 && dn_good, "negative size, dim <= 0 or null count");
    PM_REQUIRE(ctx, nq == 0 || (dq && dknn && dgood), "null pointer");
    return pmk_l2_knn2_fused(ctx, dq, nq, dt, nt, dim, 0, base, dknn, 0, ratio, dgood, dn_good);
}
int pm_knn2_ratio_l2_u8_dev(pm_ctx *ctx, const uint8_t *dq, int nq, const uint8_t *dt, int nt, int dim, float ratio, int base,
                            pm_dmatch *dknn, pm_dmatch *dgood, int32_t *dn_good)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, nq >= 0 && nt >= 0 && dim > 0 && dn_good, "negative size, dim <= 0 or null count");
    PM_REQUIRE(ctx, nq == 0 || (dq && dknn && dgood), "null pointer");
    return pmk_l2_knn2_fused(ctx, dq, nq, dt, nt, dim, 1, base, dknn, 0, ratio, dgood, dn_good);
}
int pm_knn2_hamming_dev(pm_ctx *ctx, const uint8_t *dq, int nq, const uint8_t *dt, int nt, int bytes, int base, pm_dmatch *dout)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, nq >= 0 && nt >= 0 && bytes > 0, "negative size or bytes <= 0");
    PM_REQUIRE(ctx, nq == 0 || (dq && dout), "null pointer");
    return pmk_hamming_knn2(ctx, dq, nq, dt, nt, bytes, base, dout);
}
int pm_ratio_filter_dev(pm_ctx *ctx, const pm_dmatch *dknn, int nq, float ratio, pm_dmatch *dout, int32_t *dn)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, nq >= 0 && dn, "bad argument");
    return pmk_ratio_filter(ctx, dknn, nq, ratio, dout, dn);
}
int pm_minmax_filter_dev(pm_ctx *ctx, const pm_dmatch *dm, int n, int stride, pm_dmatch *dout, int32_t *dn, double *dminmax)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, n >= 0 && (stride == 1 || stride == 2) && dn, "bad argument");
    return pmk_minmax_filter(ctx, dm, n, stride, dout, dn, dminmax);
}
int pm_col_best_hamming_dev(pm_ctx *ctx, const uint8_t *dq, int nq, const uint8_t *dt, int nt, int bytes, int base, uint64_t *dcol)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, nq >= 0 && nt >= 0 && bytes > 0, "bad argument");
    if (nq == 0 && nt > 0) { PM_CUDA(ctx, cudaMemsetAsync(dcol, 0xFF, (size_t)nt * 8, ctx->stream)); return PM_OK; }
    return pmk_hamming_col_best(ctx, dq, nq, dt, nt, bytes, base, dcol);
}
int pm_col_best_l2_f32_dev(pm_ctx *ctx, const float *dq, int nq, const float *dt, int nt, int dim, int base, uint64_t *dcol)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, nq >= 0 && nt >= 0 && dim > 0, "bad argument");
    return pmk_l2_col_best(ctx, dq, nq, dt, nt, dim, base, dcol);
}
int pm_cross_check_dev(pm_ctx *ctx, const pm_dmatch *dknn, int nq, int stride, const uint64_t *dcol, int nt, pm_dmatch *dout, int32_t *dn)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, nq >= 0 && (stride == 1 || stride == 2) && dn, "bad argument");
    return pmk_cross_check(ctx, dknn, nq, stride, dcol, nt, dout, dn);
}
int pm_gather_matches_dev(pm_ctx *ctx, const pm_dmatch *dm, const int32_t *dn, int max_matches, const float *dkp1, int nkp1,
                          const float *dkp2, int nkp2, float *dp1, float *dp2)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    return pmk_gather_matches(ctx, dm, dn, max_matches, dkp1, nkp1, dkp2, nkp2, dp1, dp2);
}
int pm_ransac_solve_dev(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const int32_t *ds, int n_hyp, int m, float *dF32)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, (m == 7 || m == 8) && n >= m && n_hyp >= 0, "sample_size must be 7 or 8 and n >= sample_size");
    return pmk_ransac_solve(ctx, dp1, dp2, n, ds, n_hyp, m, dF32);
}
int pm_ransac_score_dev(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const float *dF32, int n_models, float thr,
                        int metric, int32_t *dcounts)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, n >= 0 && n_models >= 0 && (metric == PM_METRIC_SAMPSON || metric == PM_METRIC_SYMEPI), "bad argument");
    return pmk_ransac_score(ctx, dp1, dp2, n, dF32, n_models, thr, metric, dcounts);
}
int pm_ransac_best_dev(pm_ctx *ctx, const int32_t *dcounts, int n_models, int id_base, uint64_t *dkey)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    return pmk_ransac_best(ctx, dcounts, n_models, id_base, dkey);
}
int pm_ransac_finish_dev(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const float *dFw, float thr, int metric,
                         int refit, double *dF, uint8_t *dmask, int32_t *dn)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, n >= 0 && (metric == PM_METRIC_SAMPSON || metric == PM_METRIC_SYMEPI), "bad argument");
    return pmk_ransac_finish(ctx, dp1, dp2, n, dFw, thr, metric, refit, dF, dmask, dn);
}

// ---------------------------------------------------------------------------------
// host-buffer entry points (synchronous; H2D + kernels + D2H)
// ---------------------------------------------------------------------------------
#define H2D(ctx, dst, src, bytes) PM_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (ctx)->stream))
#define H2D_ON(ctx, strm, dst, src, bytes) PM_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, strm))

// the ctx's upload stream and its events (chunked host kNN, host-buffer pair groups), created on first use
static int copy_stream_ready(pm_ctx *ctx)
{
    if (ctx->copy_stream) return PM_OK;
    PM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    PM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fence, cudaEventDisableTiming));
    PM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_train, cudaEventDisableTiming));
    for (int i = 0; i < 8; ++i) PM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming));
    return PM_OK;
}
#define D2H(ctx, dst, src, bytes) PM_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (ctx)->stream))

// Upload + L2 kNN-2 with the H2D copies overlapped with compute: the train set goes first and is packed
// while the query chunks are still crossing PCIe; every query chunk is matched as soon as it lands.
// The kNN result of row i does not depend on other query rows, so chunking changes nothing in the output.
// (Measured alternatives, both slower at cfg2 on the same box: chunk sizes falling linearly so that the last
// chunk is small, 263 -> 278 us per call; sending each chunk's kNN rows back on a second copy stream while the
// later chunks upload, 263 -> 271 us.)
static int l2_upload_and_match(pm_ctx *ctx, const void *q, int nq, const void *t, int nt, int dim, size_t elem, int is_u8,
                               uint8_t *dq, uint8_t *dt, pm_dmatch *dknn)
{
    const size_t row = (size_t)dim * elem;
    int nchunks = nq / 2560;                         // >= 10 row tiles per chunk keeps K2's grid busy
    if (nchunks > 8) nchunks = 8;
    if (nchunks < 2 || nt == 0 || dim > 128) {       // small problem: plain upload
        H2D(ctx, dq, q, (size_t)nq * row);
        if (nt) H2D(ctx, dt, t, (size_t)nt * row);
        return pmk_l2_knn2(ctx, dq, nq, dt, nt, dim, is_u8, 0, dknn);
    }
    { int cst = copy_stream_ready(ctx); if (cst != PM_OK) return cst; }
    const int chunk = pm_round_up(pm_cdiv(nq, nchunks), 256);
    // the copy stream may not overwrite the raw buffers before earlier work on the compute stream is done
    PM_CUDA(ctx, cudaEventRecord(ctx->ev_fence, ctx->stream));
    PM_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fence, 0));
    PM_CUDA(ctx, cudaMemcpyAsync(dt, t, (size_t)nt * row, cudaMemcpyHostToDevice, ctx->copy_stream));
    PM_CUDA(ctx, cudaEventRecord(ctx->ev_train, ctx->copy_stream));
    int nc = 0;
    for (int lo = 0; lo < nq; lo += chunk, ++nc) {
        const int rows = nq - lo < chunk ? nq - lo : chunk;
        PM_CUDA(ctx, cudaMemcpyAsync(dq + (size_t)lo * row, (const uint8_t *)q + (size_t)lo * row, (size_t)rows * row,
                                     cudaMemcpyHostToDevice, ctx->copy_stream));
        PM_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[nc], ctx->copy_stream));
    }
    PM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_train, 0));
    int st = pmk_l2_knn2_phase(ctx, nullptr, 0, dt, nt, dim, is_u8, 0, nullptr, 1);
    if (st != PM_OK) return st;
    nc = 0;
    for (int lo = 0; lo < nq; lo += chunk, ++nc) {
        const int rows = nq - lo < chunk ? nq - lo : chunk;
        PM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[nc], 0));
        st = pmk_l2_knn2_phase(ctx, dq + (size_t)lo * row, rows, dt, nt, dim, is_u8, lo, dknn + (size_t)lo * 2, 2);
        if (st != PM_OK) return st;
    }
    return PM_OK;
}

static int knn2_host(pm_ctx *ctx, const void *q, int nq, const void *t, int nt, int width, size_t elem, int kind,
                     pm_dmatch *out)
{
    // kind 0: l2 f32, 1: l2 u8, 2: hamming
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, nq >= 0 && nt >= 0 && width > 0, "negative size or zero width");
    if (nq == 0) return PM_OK;                       // empty query -> empty result
    PM_REQUIRE(ctx, q && out && (nt == 0 || t), "null pointer");
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t qb = (size_t)nq * width * elem, tb = (size_t)nt * width * elem;
    PM_WS(ctx, dq, uint8_t *, WS_Q_RAW, qb);
    PM_WS(ctx, dt, uint8_t *, WS_T_RAW, tb);
    PM_WS(ctx, dout, pm_dmatch *, WS_OUT, (size_t)nq * 2 * sizeof(pm_dmatch));
    int st;
    if (kind == 0 || kind == 1) st = l2_upload_and_match(ctx, q, nq, t, nt, width, elem, kind, dq, dt, dout);
    else {
        H2D(ctx, dq, q, qb);
        if (tb) H2D(ctx, dt, t, tb);
        st = pmk_hamming_knn2(ctx, dq, nq, dt, nt, width, 0, dout);
    }
    if (st != PM_OK) return st;
    D2H(ctx, out, dout, (size_t)nq * 2 * sizeof(pm_dmatch));
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PM_OK;
}

int pm_knn2_l2_f32(pm_ctx *ctx, const float *q, int nq, const float *t, int nt, int dim, pm_dmatch *out)
{ return knn2_host(ctx, q, nq, t, nt, dim, 4, 0, out); }
int pm_knn2_l2_u8(pm_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t, int nt, int dim, pm_dmatch *out)
{ return knn2_host(ctx, q, nq, t, nt, dim, 1, 1, out); }
int pm_knn2_hamming(pm_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t, int nt, int bytes, pm_dmatch *out)
{ return knn2_host(ctx, q, nq, t, nt, bytes, 1, 2, out); }

int pm_knn2_ratio_l2_f32(pm_ctx *ctx, const float *q, int nq, const float *t, int nt, int dim, float ratio,
                         pm_dmatch *knn_out, pm_dmatch *good_out, int *n_good)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, nq >= 0 && nt >= 0 && dim > 0 && n_good, "bad argument");
    *n_good = 0;
    if (nq == 0) return PM_OK;
    PM_REQUIRE(ctx, q && good_out && (nt == 0 || t), "null pointer");
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t qb = (size_t)nq * dim * 4, tb = (size_t)nt * dim * 4;
    PM_WS(ctx, dq, float *, WS_Q_RAW, qb);
    PM_WS(ctx, dt, float *, WS_T_RAW, tb);
    PM_WS(ctx, dknn, pm_dmatch *, WS_OUT, (size_t)nq * 2 * sizeof(pm_dmatch));
    PM_WS(ctx, dgood, pm_dmatch *, WS_OUT2, (size_t)nq * sizeof(pm_dmatch));
    PM_WS(ctx, dn, int32_t *, WS_KEY, 64);
    int st;
    if ((st = l2_upload_and_match(ctx, q, nq, t, nt, dim, 4, 0, (uint8_t *)dq, (uint8_t *)dt, dknn)) != PM_OK) return st;
    if ((st = pmk_ratio_filter(ctx, dknn, nq, ratio, dgood, dn)) != PM_OK) return st;
    // one synchronisation: the count, the kNN rows and the (at most nq) survivors travel together --
    // copying the unused tail of good_out (<= 16 B x nq) is cheaper than a second host round trip
    D2H(ctx, ctx->h_pinned, dn, 4);
    if (knn_out) D2H(ctx, knn_out, dknn, (size_t)nq * 2 * sizeof(pm_dmatch));
    D2H(ctx, good_out, dgood, (size_t)nq * sizeof(pm_dmatch));
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_good = ctx->h_pinned[0];
    return PM_OK;
}

static int read_count(pm_ctx *ctx, const int32_t *dn, int *n_out)
{
    D2H(ctx, ctx->h_pinned, dn, 4);
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = ctx->h_pinned[0];
    return PM_OK;
}

int pm_ratio_filter(pm_ctx *ctx, const pm_dmatch *knn, int nq, float ratio, pm_dmatch *out, int *n_out)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, nq >= 0 && n_out, "bad argument");
    *n_out = 0;
    if (nq == 0) return PM_OK;
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    PM_WS(ctx, dknn, pm_dmatch *, WS_KNN, (size_t)nq * 2 * sizeof(pm_dmatch));
    PM_WS(ctx, dout, pm_dmatch *, WS_OUT2, (size_t)nq * sizeof(pm_dmatch));
    PM_WS(ctx, dn, int32_t *, WS_KEY, 64);
    H2D(ctx, dknn, knn, (size_t)nq * 2 * sizeof(pm_dmatch));
    int st = pmk_ratio_filter(ctx, dknn, nq, ratio, dout, dn);
    if (st != PM_OK) return st;
    if ((st = read_count(ctx, dn, n_out)) != PM_OK) return st;
    if (*n_out) { D2H(ctx, out, dout, (size_t)*n_out * sizeof(pm_dmatch)); PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); }
    return PM_OK;
}

int pm_minmax_filter(pm_ctx *ctx, const pm_dmatch *m, int n, int stride, pm_dmatch *out, int *n_out, double *mn, double *mx)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, n >= 0 && (stride == 1 || stride == 2) && n_out, "bad argument");
    *n_out = 0;
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    PM_WS(ctx, dm, pm_dmatch *, WS_KNN, (size_t)(n ? n : 1) * stride * sizeof(pm_dmatch));
    PM_WS(ctx, dout, pm_dmatch *, WS_OUT2, (size_t)(n ? n : 1) * sizeof(pm_dmatch));
    PM_WS(ctx, dn, int32_t *, WS_KEY, 64);
    double *dmm = reinterpret_cast<double *>(dn + 4);
    if (n) H2D(ctx, dm, m, (size_t)n * stride * sizeof(pm_dmatch));
    int st = pmk_minmax_filter(ctx, dm, n, stride, dout, dn, dmm);
    if (st != PM_OK) return st;
    D2H(ctx, ctx->h_pinned, dn, 32);
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = ctx->h_pinned[0];
    const double *hmm = reinterpret_cast<const double *>(ctx->h_pinned + 4);
    if (mn) *mn = hmm[0];
    if (mx) *mx = hmm[1];
    if (*n_out) { D2H(ctx, out, dout, (size_t)*n_out * sizeof(pm_dmatch)); PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); }
    return PM_OK;
}

static int match_cross_host(pm_ctx *ctx, const void *q, int nq, const void *t, int nt, int width, size_t elem, int kind,
                            pm_dmatch *out, int *n_out)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, nq >= 0 && nt >= 0 && width > 0 && n_out, "bad argument");
    *n_out = 0;
    if (nq == 0 || nt == 0) return PM_OK;
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t qb = (size_t)nq * width * elem, tb = (size_t)nt * width * elem;
    PM_WS(ctx, dq, uint8_t *, WS_Q_RAW, qb);
    PM_WS(ctx, dt, uint8_t *, WS_T_RAW, tb);
    PM_WS(ctx, dknn, pm_dmatch *, WS_OUT, (size_t)nq * 2 * sizeof(pm_dmatch));
    PM_WS(ctx, dout, pm_dmatch *, WS_OUT2, (size_t)nq * sizeof(pm_dmatch));
    PM_WS(ctx, dcol, uint64_t *, WS_COLBEST, (size_t)nt * 8);
    PM_WS(ctx, dn, int32_t *, WS_KEY, 64);
    H2D(ctx, dq, q, qb);
    H2D(ctx, dt, t, tb);
    int st;
    if (kind == 0) {
        if ((st = pmk_l2_knn2(ctx, dq, nq, dt, nt, width, 0, 0, dknn)) != PM_OK) return st;
    } else {
        if ((st = pmk_hamming_knn2(ctx, dq, nq, dt, nt, width, 0, dknn)) != PM_OK) return st;
    }
    if ((st = pmk_cross_col_best(ctx, kind != 0, dq, nq, dt, nt, width, 0, dknn, dcol, nullptr)) != PM_OK) return st;
    if ((st = pmk_cross_check(ctx, dknn, nq, 2, dcol, nt, dout, dn)) != PM_OK) return st;
    if ((st = read_count(ctx, dn, n_out)) != PM_OK) return st;
    if (*n_out) { D2H(ctx, out, dout, (size_t)*n_out * sizeof(pm_dmatch)); PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); }
    return PM_OK;
}

int pm_match_cross_l2_f32(pm_ctx *ctx, const float *q, int nq, const float *t, int nt, int dim, pm_dmatch *out, int *n_out)
{ return match_cross_host(ctx, q, nq, t, nt, dim, 4, 0, out, n_out); }
int pm_match_cross_hamming(pm_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t, int nt, int bytes, pm_dmatch *out, int *n_out)
{ return match_cross_host(ctx, q, nq, t, nt, bytes, 1, 1, out, n_out); }

int pm_gather_points(pm_ctx *ctx, const float *kp, int nkp, const int32_t *idx, int n, float *out)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, nkp >= 0 && n >= 0, "bad argument");
    if (n == 0) return PM_OK;
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    PM_WS(ctx, dkp, float *, WS_KP, (size_t)(nkp ? nkp : 1) * 8);
    PM_WS(ctx, didx, int32_t *, WS_IDX, (size_t)n * 4);
    PM_WS(ctx, dout, float *, WS_P1, (size_t)n * 8);
    if (nkp) H2D(ctx, dkp, kp, (size_t)nkp * 8);
    H2D(ctx, didx, idx, (size_t)n * 4);
    int st = pmk_gather_points(ctx, dkp, nkp, didx, n, dout);
    if (st != PM_OK) return st;
    D2H(ctx, out, dout, (size_t)n * 8);
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PM_OK;
}

// splitmix64: deterministic, rank-invariant sample sets
static inline uint64_t splitmix64(uint64_t &s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int pm_make_sample_sets(int n_points, int n_hyp, int m, uint64_t seed, int32_t *out)
{
    if (n_points < m || m <= 0 || m > 8 || n_hyp < 0 || !out) return PM_BAD_ARG;
    for (int h = 0; h < n_hyp; ++h) {
        uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(h + 1));
        int32_t *row = out + (size_t)h * m;
        for (int i = 0; i < m; ++i) {
            for (;;) {
                const int32_t v = (int32_t)(splitmix64(s) % (uint64_t)n_points);
                bool dup = false;
                for (int k = 0; k < i; ++k) dup = dup || row[k] == v;
                if (!dup) { row[i] = v; break; }
            }
        }
    }
    return PM_OK;
}

int pm_make_sample_sets_dev(pm_ctx *ctx, int n_points, int n_hyp, int m, uint64_t seed, int32_t *dout)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, n_points >= m && m > 0 && m <= 8 && n_hyp >= 0 && dout, "need n_points >= m, 0 < m <= 8");
    return pmk_sample_sets(ctx, n_points, n_hyp, m, seed, dout);
}

int pm_find_fundamental(pm_ctx *ctx, const float *p1, const float *p2, int n, const pm_ransac_params *prm,
                        double F[9], uint8_t *mask, int *n_inliers)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, prm && p1 && p2 && F, "null pointer");
    const int m = prm->sample_size;
    PM_REQUIRE(ctx, m == 7 || m == 8, "sample_size must be 7 or 8");
    PM_REQUIRE(ctx, prm->metric == PM_METRIC_SAMPSON || prm->metric == PM_METRIC_SYMEPI, "unknown metric");
    PM_REQUIRE(ctx, prm->n_hyp >= 0 && n >= 0, "negative size");
    if (n_inliers) *n_inliers = 0;
    if (n < m || prm->n_hyp == 0) return PM_EMPTY;          // cv::findFundamentalMat: empty Mat
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int per = m == 8 ? 1 : 3, nh = prm->n_hyp;
    PM_WS(ctx, dp1, float *, WS_P1, (size_t)n * 8);
    PM_WS(ctx, dp2, float *, WS_P2, (size_t)n * 8);
    PM_WS(ctx, ds, int32_t *, WS_SAMPLES, (size_t)nh * m * 4);
    PM_WS(ctx, dF32, float *, WS_F32, ((size_t)nh * per + 1) * 12 * 4);
    PM_WS(ctx, dcounts, int32_t *, WS_COUNTS, (size_t)nh * per * 4);
    PM_WS(ctx, dkey, uint64_t *, WS_KEY, 64);
    PM_WS(ctx, dmask, uint8_t *, WS_MASK, (size_t)n);
    PM_WS(ctx, dFout, double *, WS_FOUT, 16 * 8);
    float *dFw = dF32 + (size_t)nh * per * 12;
    int32_t *dninl = reinterpret_cast<int32_t *>(dkey + 2);
    H2D(ctx, dp1, p1, (size_t)n * 8);
    H2D(ctx, dp2, p2, (size_t)n * 8);
    int st;
    if (prm->sample_idx) H2D(ctx, ds, prm->sample_idx, (size_t)nh * m * 4);
    else if ((st = pmk_sample_sets(ctx, n, nh, m, prm->seed, ds)) != PM_OK) return st;   // same sets as pm_make_sample_sets
    if ((st = pmk_ransac_solve(ctx, dp1, dp2, n, ds, nh, m, dF32)) != PM_OK) return st;
    if ((st = pmk_ransac_score(ctx, dp1, dp2, n, dF32, nh * per, prm->threshold, prm->metric, dcounts)) != PM_OK) return st;
    if ((st = pmk_ransac_best(ctx, dcounts, nh * per, 0, dkey)) != PM_OK) return st;
    if ((st = pmk_ransac_pick(ctx, dkey, dF32, 0, nh * per, dFw)) != PM_OK) return st;
    if ((st = pmk_ransac_finish(ctx, dp1, dp2, n, dFw, prm->threshold, prm->metric, prm->refit, dFout, dmask, dninl)) != PM_OK) return st;
    D2H(ctx, ctx->h_pinned, dkey, 32);
    D2H(ctx, ctx->h_pinned + 16, dFout, 72);
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint64_t key = *reinterpret_cast<const uint64_t *>(ctx->h_pinned);
    if (key == 0) return PM_EMPTY;
    memcpy(F, ctx->h_pinned + 16, 72);
    if (n_inliers) *n_inliers = ctx->h_pinned[4];
    if (mask) { D2H(ctx, mask, dmask, (size_t)n); PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); }
    return PM_OK;
}

// cv::RANSACUpdateNumIters (OpenCV ptsetreg.cpp), restated: iterations needed so that with probability p at least one
// sample of `model_points` is outlier-free when the outlier ratio is ep
static int cv_update_num_iters(double p, double ep, int model_points, int max_iters)
{
    p = p < 0 ? 0 : (p > 1 ? 1 : p);
    ep = ep < 0 ? 0 : (ep > 1 ? 1 : ep);
    double num = 1 - p > DBL_MIN ? 1 - p : DBL_MIN;
    double denom = 1 - std::pow(1 - ep, model_points);
    if (denom < DBL_MIN) return 0;
    num = std::log(num); denom = std::log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)std::lround(num / denom);
}

int pm_find_fundamental_adaptive(pm_ctx *ctx, const float *p1, const float *p2, int n, const pm_ransac_params *prm,
                                 double F[9], uint8_t *mask, int *n_inliers, int *n_hyp_run)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, prm && p1 && p2 && F, "null pointer");
    const int m = prm->sample_size;
    PM_REQUIRE(ctx, m == 7 || m == 8, "sample_size must be 7 or 8");
    PM_REQUIRE(ctx, prm->metric == PM_METRIC_SAMPSON || prm->metric == PM_METRIC_SYMEPI, "unknown metric");
    PM_REQUIRE(ctx, n >= 0, "negative size");
    if (n_inliers) *n_inliers = 0;
    if (n_hyp_run) *n_hyp_run = 0;
    if (n < m) return PM_EMPTY;
    const int max_iters = prm->max_iters > 0 ? prm->max_iters : 1000;
    const double conf = (prm->confidence > DBL_EPSILON && prm->confidence < 1 - DBL_EPSILON) ? prm->confidence : 0.99;
    const int batch = prm->n_hyp > 0 ? (prm->n_hyp < max_iters ? prm->n_hyp : max_iters) : (max_iters < 1024 ? max_iters : 1024);
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int per = m == 8 ? 1 : 3;
    PM_WS(ctx, dp1, float *, WS_P1, (size_t)n * 8);
    PM_WS(ctx, dp2, float *, WS_P2, (size_t)n * 8);
    PM_WS(ctx, ds, int32_t *, WS_SAMPLES, (size_t)batch * m * 4);
    PM_WS(ctx, dF32, float *, WS_F32, ((size_t)batch * per + 2) * 12 * 4);
    PM_WS(ctx, dcounts, int32_t *, WS_COUNTS, (size_t)batch * per * 4);
    PM_WS(ctx, dkey, uint64_t *, WS_KEY, 64);
    PM_WS(ctx, dmask, uint8_t *, WS_MASK, (size_t)n);
    PM_WS(ctx, dFout, double *, WS_FOUT, 16 * 8);
    float *dFw = dF32 + (size_t)batch * per * 12;           // running winner (12 floats)
    uint64_t *dbest = dkey + 1;                             // running winner key
    int32_t *dninl = reinterpret_cast<int32_t *>(dkey + 2);
    H2D(ctx, dp1, p1, (size_t)n * 8);                       // the correspondences cross PCIe once
    H2D(ctx, dp2, p2, (size_t)n * 8);
    PM_CUDA(ctx, cudaMemsetAsync(dbest, 0, 8, ctx->stream));
    int st, done = 0, need = max_iters;
    uint64_t best = 0;
    while (done < (need < max_iters ? need : max_iters)) {
        const int nh = batch < max_iters - done ? batch : max_iters - done;
        // batch b draws the sets of hypotheses [done, done + nh) of the stream pm_make_sample_sets(n, max_iters, m, seed) defines
        if ((st = pmk_sample_sets(ctx, n, nh, m, prm->seed, ds, nullptr, done)) != PM_OK) return st;
        if ((st = pmk_ransac_solve(ctx, dp1, dp2, n, ds, nh, m, dF32)) != PM_OK) return st;
        if ((st = pmk_ransac_score(ctx, dp1, dp2, n, dF32, nh * per, prm->threshold, prm->metric, dcounts)) != PM_OK) return st;
        if ((st = pmk_ransac_best(ctx, dcounts, nh * per, done * per, dkey)) != PM_OK) return st;
        if ((st = pmk_ransac_update_best(ctx, dkey, dF32, done * per, nh * per, dbest, dFw)) != PM_OK) return st;
        D2H(ctx, ctx->h_pinned, dbest, 8);                  // the only per-batch traffic: 8 bytes
        PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        done += nh;
        const uint64_t k = *reinterpret_cast<const uint64_t *>(ctx->h_pinned);
        if (k > best) {
            best = k;
            const int good = (int)(k >> 32);
            need = cv_update_num_iters(conf, (double)(n - good) / n, m, max_iters);
        }
    }
    if (n_hyp_run) *n_hyp_run = done;
    if (best == 0) return PM_EMPTY;
    if ((st = pmk_ransac_finish(ctx, dp1, dp2, n, dFw, prm->threshold, prm->metric, prm->refit, dFout, dmask, dninl)) != PM_OK) return st;
    D2H(ctx, ctx->h_pinned, dninl, 4);
    D2H(ctx, ctx->h_pinned + 16, dFout, 72);
    if (mask) D2H(ctx, mask, dmask, (size_t)n);
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(F, ctx->h_pinned + 16, 72);
    if (n_inliers) *n_inliers = ctx->h_pinned[0];
    return PM_OK;
}

int pm_fundamental_7point(pm_ctx *ctx, const float *p1, const float *p2, int n, double F[27], int *n_models)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, p1 && p2 && F && n_models, "null pointer");
    *n_models = 0;
    PM_REQUIRE(ctx, n == 7, "the 7-point solver takes exactly seven correspondences");
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    PM_WS(ctx, dp1, float *, WS_P1, 7 * 8);
    PM_WS(ctx, dp2, float *, WS_P2, 7 * 8);
    PM_WS(ctx, ds, int32_t *, WS_SAMPLES, 8 * 4);
    PM_WS(ctx, dF32, float *, WS_F32, 4 * 12 * 4);
    PM_WS(ctx, dF64, double *, WS_FOUT, 32 * 8);
    for (int i = 0; i < 7; ++i) ctx->h_pinned[i] = i;
    H2D(ctx, dp1, p1, 7 * 8);
    H2D(ctx, dp2, p2, 7 * 8);
    H2D(ctx, ds, ctx->h_pinned, 7 * 4);
    int st = pmk_ransac_solve(ctx, dp1, dp2, 7, ds, 1, 7, dF32, nullptr, dF64);
    if (st != PM_OK) return st;
    double *h = reinterpret_cast<double *>(ctx->h_pinned + 16);
    D2H(ctx, h, dF64, 27 * 8);
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int k = 0;                                              // the solver stores NaN for absent / non-finite roots
    for (int r = 0; r < 3; ++r)
        if (h[9 * r] == h[9 * r]) { memcpy(F + 9 * k, h + 9 * r, 72); ++k; }
    *n_models = k;
    return k ? PM_OK : PM_EMPTY;
}

int pm_find_fundamental_mat(pm_ctx *ctx, const float *p1, const float *p2, int n, int method, double param1, double param2,
                            int max_iters, const pm_fm_options *opt, double F[27], int *n_models, uint8_t *mask)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, p1 && p2 && F && n_models && n >= 0, "bad argument");
    PM_REQUIRE(ctx, method == PM_FM_7POINT || method == PM_FM_8POINT || method == PM_FM_LMEDS || method == PM_FM_RANSAC,
               "method must be FM_7POINT, FM_8POINT, FM_LMEDS or FM_RANSAC");
    *n_models = 0;
    if (n < 7) return PM_EMPTY;                             // cv: empty Mat
    int st;
    if (n == 7) {                                           // any method: run7Point on the seven points, every real root
        if ((st = pm_fundamental_7point(ctx, p1, p2, 7, F, n_models)) != PM_OK) return st;
        if (mask) memset(mask, 1, 7);
        return PM_OK;
    }
    if (method == PM_FM_8POINT) {
        if ((st = pm_fundamental_8point(ctx, p1, p2, n, F)) != PM_OK) return st;
        *n_models = 1;
        if (mask) memset(mask, 1, (size_t)n);
        return PM_OK;
    }
    if (param1 <= 0) param1 = 3.;
    if (!(param2 > DBL_EPSILON && param2 < 1 - DBL_EPSILON)) param2 = 0.99;
    if (max_iters <= 0) max_iters = 1000;
    const uint64_t seed = opt ? opt->seed : 0;
    int ninl = 0;
    if (method == PM_FM_RANSAC && n >= 15) {
        pm_ransac_params prm;
        memset(&prm, 0, sizeof(prm));
        prm.sample_size = opt && opt->sample_size ? opt->sample_size : 7;
        prm.metric = opt ? opt->metric : PM_METRIC_SYMEPI;
        prm.threshold = (float)param1; prm.refit = opt ? opt->refit : 0;
        prm.n_hyp = opt && opt->batch > 0 ? opt->batch : 1024;
        prm.seed = seed; prm.max_iters = max_iters; prm.confidence = param2;
        st = pm_find_fundamental_adaptive(ctx, p1, p2, n, &prm, F, mask, &ninl, nullptr);
    } else {
        // LMedS: niters = RANSACUpdateNumIters(confidence, outlier ratio 0.45, 7 points, max_iters), at least 3
        int niters = cv_update_num_iters(param2, 0.45, 7, max_iters);
        if (niters < 3) niters = 3;
        st = pm_find_fundamental_lmeds(ctx, p1, p2, n, niters, nullptr, seed, F, mask, &ninl, nullptr);
    }
    if (st == PM_OK) *n_models = 1;
    return st;
}

int pm_find_fundamental_dev(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const pm_ransac_params *prm,
                            double *dF, uint8_t *dmask, int32_t *dn_inliers, uint64_t *dkey)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, prm && dp1 && dp2 && dF && dmask && dn_inliers && dkey && prm->sample_idx, "null pointer");
    const int m = prm->sample_size;
    PM_REQUIRE(ctx, m == 7 || m == 8, "sample_size must be 7 or 8");
    PM_REQUIRE(ctx, prm->metric == PM_METRIC_SAMPSON || prm->metric == PM_METRIC_SYMEPI, "unknown metric");
    PM_REQUIRE(ctx, prm->n_hyp > 0 && n >= m, "need n >= sample_size and n_hyp > 0");
    const int per = m == 8 ? 1 : 3, nh = prm->n_hyp;
    PM_WS(ctx, dF32, float *, WS_F32, ((size_t)nh * per + 1) * 12 * 4);
    PM_WS(ctx, dcounts, int32_t *, WS_COUNTS, (size_t)nh * per * 4);
    float *dFw = dF32 + (size_t)nh * per * 12;
    int st;
    if ((st = pmk_ransac_solve(ctx, dp1, dp2, n, prm->sample_idx, nh, m, dF32)) != PM_OK) return st;
    if ((st = pmk_ransac_score(ctx, dp1, dp2, n, dF32, nh * per, prm->threshold, prm->metric, dcounts)) != PM_OK) return st;
    if ((st = pmk_ransac_best(ctx, dcounts, nh * per, prm->hyp_id_base * per, dkey)) != PM_OK) return st;
    if ((st = pmk_ransac_pick(ctx, dkey, dF32, prm->hyp_id_base * per, nh * per, dFw)) != PM_OK) return st;
    return pmk_ransac_finish(ctx, dp1, dp2, n, dFw, prm->threshold, prm->metric, prm->refit, dF, dmask, dn_inliers);
}

// main.cpp:43-98 for a GROUP of image pairs, entirely in stream order (no host round trip): per pair the matching chain
// K1, K2, K3, K5 (which also gathers the keypoint coordinates of the survivors into the pair's slot), then ONE launch per
// RANSAC kernel for the whole group (pmk_pair_group_ransac: sample sets, K6, K7, best+pick, mask, three refit passes and
// the refit solve, which writes the result records).  The match counts stay on the device and bound every RANSAC kernel
// through its n_dev argument.  A group of one pair is the single-pair entry.
static int pair_group_enqueue(pm_ctx *ctx, int cnt, const void *const *dd1, const int32_t *n1, const void *const *dd2,
                              const int32_t *n2, int dim, int is_u8, const float *const *dkp1, const float *const *dkp2, float ratio,
                              const pm_ransac_params *prm, uint64_t seed0, pm_pair_result *dres, bool host_inputs)
{
    const int m = prm->sample_size, per = m == 8 ? 1 : 3, nh = prm->n_hyp;
    int nmax = 1, n2max = 1;
    for (int b = 0; b < cnt; ++b) { if (n1[b] > nmax) nmax = n1[b]; if (n2[b] > n2max) n2max = n2[b]; }
    const size_t P = (size_t)cnt;
    PM_WS(ctx, dknn, pm_dmatch *, WS_KNN, (size_t)nmax * 2 * sizeof(pm_dmatch));
    PM_WS(ctx, dgood, pm_dmatch *, WS_OUT2, (size_t)nmax * sizeof(pm_dmatch));
    PM_WS(ctx, dkey, uint64_t *, WS_KEY, P * 64);
    PM_WS(ctx, dp1, float *, WS_P1, P * nmax * 8);
    PM_WS(ctx, dp2, float *, WS_P2, P * nmax * 8);
    PM_WS(ctx, dpts, float *, WS_MISC, P * nmax * 16);          // the matches as {x1, y1, x2, y2}
    PM_WS(ctx, ds, int32_t *, WS_SAMPLES, P * nh * m * 4);
    PM_WS(ctx, dF32, float *, WS_F32, P * ((size_t)nh * per + 1) * 12 * 4);
    PM_WS(ctx, dcounts, int32_t *, WS_COUNTS, P * nh * per * 4);
    PM_WS(ctx, dmask, uint8_t *, WS_MASK, P * nmax);
    PM_WS(ctx, dF, double *, WS_FOUT, P * 16 * 8);
    PM_WS(ctx, drefit, double *, WS_REFIT, P * PM_REFIT_WS_DOUBLES * sizeof(double));
    uint8_t *s1 = nullptr, *s2 = nullptr; float *sk1 = nullptr, *sk2 = nullptr;
    const size_t elem = is_u8 ? 1 : 4;
    // staging for HOST descriptors / keypoints: TWO sets, filled by the ctx's upload stream -- the upload of pair b + 1
    // runs under the matching chain of pair b (ev_chunk[set]: "set uploaded", ev_chunk[2 + set]: "the chain that read the
    // set is done"; an event that was never recorded counts as complete)
    const size_t st1 = ((size_t)nmax * dim * elem + 255) & ~(size_t)255, st2 = ((size_t)n2max * dim * elem + 255) & ~(size_t)255;
    const size_t stk1 = ((size_t)nmax * 8 + 255) & ~(size_t)255, stk2 = ((size_t)n2max * 8 + 255) & ~(size_t)255;
    if (host_inputs) {
        PM_WS(ctx, a1, uint8_t *, WS_Q_RAW, 2 * st1);
        PM_WS(ctx, a2, uint8_t *, WS_T_RAW, 2 * st2);
        PM_WS(ctx, a3, float *, WS_KP, 2 * stk1);
        PM_WS(ctx, a4, float *, WS_KP2, 2 * stk2);
        s1 = a1; s2 = a2; sk1 = a3; sk2 = a4;
        int cst = copy_stream_ready(ctx);
        if (cst != PM_OK) return cst;
        // the first uploads may not overtake earlier work of this stream that still reads the staging sets or the slots
        PM_CUDA(ctx, cudaEventRecord(ctx->ev_fence, ctx->stream));
        PM_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fence, 0));
    }
    int st;
    for (int b = 0; b < cnt; ++b) {
        const void *d1 = dd1[b], *d2 = dd2[b];
        const float *k1 = dkp1[b], *k2 = dkp2[b];
        if (host_inputs) {
            const int set = b & 1;
            uint8_t *u1 = s1 + set * st1, *u2 = s2 + set * st2;
            float *uk1 = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(sk1) + set * stk1);
            float *uk2 = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(sk2) + set * stk2);
            cudaStream_t cs = ctx->copy_stream;
            if (b >= 2) PM_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_chunk[2 + set], 0));      // chain b - 2 has read the set
            if (n1[b]) { H2D_ON(ctx, cs, u1, d1, (size_t)n1[b] * dim * elem); H2D_ON(ctx, cs, uk1, k1, (size_t)n1[b] * 8); }
            if (n2[b]) { H2D_ON(ctx, cs, u2, d2, (size_t)n2[b] * dim * elem); H2D_ON(ctx, cs, uk2, k2, (size_t)n2[b] * 8); }
            PM_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[set], cs));
            PM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[set], 0));
            d1 = u1; d2 = u2; k1 = uk1; k2 = uk2;
        }
        int32_t *dn_good = reinterpret_cast<int32_t *>(dkey + (size_t)b * 8 + 2);
        // K5 gathers while it scatters: the good matches leave as DMatch records AND as the pair's point lists
        const pm_gather_out g = {k1, n1[b], k2, n2[b], dp1 + (size_t)b * nmax * 2, dp2 + (size_t)b * nmax * 2, dpts + (size_t)b * nmax * 4};
        if ((st = pmk_l2_knn2_fused(ctx, d1, n1[b], d2, n2[b], dim, is_u8, 0, dknn, 0, ratio, dgood, dn_good, &g)) != PM_OK) return st;
        if (host_inputs) PM_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[2 + (b & 1)], ctx->stream));
    }
    pm_pair_group G;
    G.n_pairs = cnt; G.nmax = nmax; G.n_hyp = nh; G.m = m; G.metric = prm->metric; G.refit_on = prm->refit != 0;
    G.threshold = prm->threshold; G.seed0 = seed0;
    G.p1 = reinterpret_cast<const float2 *>(dp1); G.p2 = reinterpret_cast<const float2 *>(dp2);
    G.pts4 = reinterpret_cast<const float4 *>(dpts);
    G.samples = ds; G.F32 = dF32; G.counts = dcounts; G.key = dkey; G.mask = dmask; G.refit = drefit; G.Fout = dF; G.res = dres;
    return pmk_pair_group_ransac(ctx, G);
}

// pairs per group of the batched calls (PM_PAIR_GROUP overrides; results do not depend on it)
static int pair_group_size()
{
    static const int grp_env = getenv("PM_PAIR_GROUP") ? atoi(getenv("PM_PAIR_GROUP")) : 0;
    return grp_env > 0 ? (grp_env < 64 ? grp_env : 64) : 16;
}

static int pair_check(pm_ctx *ctx, int dim, const pm_ransac_params *prm)
{
    PM_REQUIRE(ctx, prm != nullptr && dim > 0, "null parameters or dim <= 0");
    PM_REQUIRE(ctx, prm->sample_size == 7 || prm->sample_size == 8, "sample_size must be 7 or 8");
    PM_REQUIRE(ctx, prm->metric == PM_METRIC_SAMPSON || prm->metric == PM_METRIC_SYMEPI, "unknown metric");
    PM_REQUIRE(ctx, prm->n_hyp > 0, "n_hyp must be positive");
    return PM_OK;
}

int pm_match_estimate_pair_dev(pm_ctx *ctx, const void *dd1, int n1, const void *dd2, int n2, int dim, int is_u8,
                               const float *dkp1, const float *dkp2, float ratio, const pm_ransac_params *prm, uint64_t seed,
                               pm_pair_result *dres)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    int st = pair_check(ctx, dim, prm);
    if (st != PM_OK) return st;
    PM_REQUIRE(ctx, n1 >= 0 && n2 >= 0 && dres, "negative size or null result");
    PM_REQUIRE(ctx, (n1 == 0 || (dd1 && dkp1)) && (n2 == 0 || (dd2 && dkp2)), "null pointer");
    const int32_t c1 = n1, c2 = n2;
    return pair_group_enqueue(ctx, 1, &dd1, &c1, &dd2, &c2, dim, is_u8, &dkp1, &dkp2, ratio, prm, seed, dres, false);
}

static int batched_impl(pm_ctx *ctx, int n_pairs, const void *const *dd1, const int32_t *n1, const void *const *dd2,
                        const int32_t *n2, int dim, int is_u8, const float *const *dkp1, const float *const *dkp2,
                        float ratio, const pm_ransac_params *prm, pm_pair_result *dres, bool host_inputs);

int pm_match_estimate_batched_dev(pm_ctx *ctx, int n_pairs, const void *const *dd1, const int32_t *n1, const void *const *dd2,
                                  const int32_t *n2, int dim, int is_u8, const float *const *dkp1, const float *const *dkp2,
                                  float ratio, const pm_ransac_params *prm, pm_pair_result *dres)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    return batched_impl(ctx, n_pairs, dd1, n1, dd2, n2, dim, is_u8, dkp1, dkp2, ratio, prm, dres, false);
}

// The host-buffer form (BASELINE config 5 end to end): descriptors and keypoints of every pair in HOST memory (pinned for
// full PCIe speed), results back in host memory; synchronous.  Every lane uploads its pairs on its own stream right before
// their kernels, so the uploads of one lane run under the kernels of the others; the only device-to-host traffic is the
// 96-byte record per pair, read once at the end.
int pm_match_estimate_batched(pm_ctx *ctx, int n_pairs, const void *const *desc1, const int32_t *n1, const void *const *desc2,
                              const int32_t *n2, int dim, int is_u8, const float *const *kp1, const float *const *kp2,
                              float ratio, const pm_ransac_params *prm, pm_pair_result *results)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, n_pairs >= 0 && (n_pairs == 0 || results), "bad argument");
    if (n_pairs == 0) return PM_OK;
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    PM_WS(ctx, dres, pm_pair_result *, WS_PAIRRES, (size_t)n_pairs * sizeof(pm_pair_result));
    int st = batched_impl(ctx, n_pairs, desc1, n1, desc2, n2, dim, is_u8, kp1, kp2, ratio, prm, dres, true);
    if (st != PM_OK) return st;
    D2H(ctx, results, dres, (size_t)n_pairs * sizeof(pm_pair_result));
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PM_OK;
}

static int batched_impl(pm_ctx *ctx, int n_pairs, const void *const *dd1, const int32_t *n1, const void *const *dd2,
                        const int32_t *n2, int dim, int is_u8, const float *const *dkp1, const float *const *dkp2,
                        float ratio, const pm_ransac_params *prm, pm_pair_result *dres, bool host_inputs)
{
    int st = pair_check(ctx, dim, prm);
    if (st != PM_OK) return st;
    PM_REQUIRE(ctx, n_pairs >= 0, "negative pair count");
    if (n_pairs == 0) return PM_OK;
    PM_REQUIRE(ctx, dd1 && n1 && dd2 && n2 && dkp1 && dkp2 && dres, "null pointer");
    for (int p = 0; p < n_pairs; ++p) {
        PM_REQUIRE(ctx, n1[p] >= 0 && n2[p] >= 0, "negative size");
        PM_REQUIRE(ctx, (n1[p] == 0 || (dd1[p] && dkp1[p])) && (n2[p] == 0 || (dd2[p] && dkp2[p])), "null pointer");
    }
    // consecutive pairs form groups of `grp` (their RANSAC kernels share launches); group g runs on lane g % L
    const int grp = pair_group_size();
    const int n_groups = pm_cdiv(n_pairs, grp);
    const int L = n_groups < ctx->batch_lanes ? n_groups : ctx->batch_lanes;
    auto enqueue_group = [&](pm_ctx *c, int g) {
        const int p0 = g * grp, cnt = n_pairs - p0 < grp ? n_pairs - p0 : grp;
        return pair_group_enqueue(c, cnt, dd1 + p0, n1 + p0, dd2 + p0, n2 + p0, dim, is_u8, dkp1 + p0, dkp2 + p0, ratio, prm,
                                  prm->seed + (uint64_t)p0, dres + p0, host_inputs);
    };
    if (L <= 1) {
        for (int g = 0; g < n_groups; ++g)
            if ((st = enqueue_group(ctx, g)) != PM_OK) return st;
        return PM_OK;
    }
    // Groups are independent: group g runs on lane g % L (a child context with its own stream and workspaces), so
    // the latency-bound kernels of one group (minimal solves, the refit eigen-solves) overlap the other lanes' matching.  The lanes fork from and join the ctx stream through events: to the caller the
    // call still behaves like one stream-ordered operation.
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->ev_fork) PM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    for (int k = 0; k < L; ++k) {
        if (!ctx->lane[k]) {
            st = pm_create(&ctx->lane[k], ctx->device);
            if (st != PM_OK) return pm_fail(ctx, st, "pm_match_estimate_batched_dev: cannot create lane %d", k);
            PM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_lane[k], cudaEventDisableTiming));
        }
    }
    PM_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    for (int k = 0; k < L; ++k) PM_CUDA(ctx, cudaStreamWaitEvent(ctx->lane[k]->stream, ctx->ev_fork, 0));
    // one host thread per lane: a pair is 13 launches, and a single thread enqueuing eight lanes measured 94 instead
    // of 67 us per pair.  (Lanes and their workspaces are created on first use -- cudaMalloc synchronises the device --
    // so a batch that meets a cold lane is slow; an earlier measurement without a warm-up made 6 and 8 lanes look 5x
    // slower than 4.)
    uint64_t before[PM_MAX_LANES];
    int lane_st[PM_MAX_LANES];
    for (int k = 0; k < L; ++k) { before[k] = ctx->lane[k]->launches; lane_st[k] = PM_OK; }
    auto run_lane = [&](int k) {
        cudaSetDevice(ctx->device);
        for (int g = k; g < n_groups && lane_st[k] == PM_OK; g += L) lane_st[k] = enqueue_group(ctx->lane[k], g);
    };
    static const bool lane_threads = !(getenv("PM_BATCH_THREADS") && atoi(getenv("PM_BATCH_THREADS")) == 0);
    if (n_groups >= 2 * L && lane_threads) {
        std::vector<std::thread> workers;
        for (int k = 1; k < L; ++k) workers.emplace_back(run_lane, k);
        run_lane(0);
        for (auto &w : workers) w.join();
    } else {
        for (int k = 0; k < L; ++k) run_lane(k);
    }
    // join first, report afterwards: whatever the lanes did enqueue (it writes dres and reads the caller's buffers) must
    // be ordered before anything the caller enqueues on the ctx stream next, also when a lane failed half-way
    int first_bad = -1;
    for (int k = 0; k < L; ++k) {
        ctx->launches += ctx->lane[k]->launches - before[k];
        if (lane_st[k] != PM_OK && first_bad < 0) first_bad = k;
        cudaError_t e1 = cudaEventRecord(ctx->ev_lane[k], ctx->lane[k]->stream);
        cudaError_t e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(ctx->stream, ctx->ev_lane[k], 0) : e1;
        if (e2 != cudaSuccess && first_bad < 0) {
            first_bad = k; lane_st[k] = PM_CUDA_ERR;
            ctx->lane[k]->err = std::string("joining the lane: ") + cudaGetErrorString(e2);
        }
    }
    ctx->tail_is_chain = false;
    if (first_bad >= 0) return pm_fail(ctx, lane_st[first_bad], "lane %d: %s", first_bad, ctx->lane[first_bad]->err.c_str());
    return PM_OK;
}

// Creates the lanes and lets every one of them allocate its workspaces for pairs of up to n1 x n2 descriptors: a
// throw-away batch of all-zero descriptors (no match survives the ratio test, so the RANSAC kernels run empty).
int pm_batch_warmup(pm_ctx *ctx, int n1, int n2, int dim, int is_u8, const pm_ransac_params *prm)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    int st = pair_check(ctx, dim, prm);
    if (st != PM_OK) return st;
    PM_REQUIRE(ctx, n1 > 0 && n2 > 0, "sizes must be positive");
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t elem = is_u8 ? 1 : 4, b1 = (size_t)n1 * dim * elem, b2 = (size_t)n2 * dim * elem;
    const int np = ctx->batch_lanes * pair_group_size();       // one full group for every lane
    uint8_t *buf = nullptr;
    const size_t off_d2 = (b1 + 255) & ~(size_t)255, off_k1 = off_d2 + ((b2 + 255) & ~(size_t)255),
                 off_k2 = off_k1 + (((size_t)n1 * 8 + 255) & ~(size_t)255), off_res = off_k2 + (((size_t)n2 * 8 + 255) & ~(size_t)255),
                 total = off_res + (size_t)np * sizeof(pm_pair_result);
    PM_CUDA(ctx, cudaMalloc((void **)&buf, total));
    cudaError_t e = cudaMemsetAsync(buf, 0, total, ctx->stream);
    std::vector<const void *> d1(np, buf), d2(np, buf + off_d2);
    std::vector<const float *> k1(np, (const float *)(buf + off_k1)), k2(np, (const float *)(buf + off_k2));
    std::vector<int32_t> c1(np, n1), c2(np, n2);
    if (e == cudaSuccess)
        st = pm_match_estimate_batched_dev(ctx, np, d1.data(), c1.data(), d2.data(), c2.data(), dim, is_u8, k1.data(), k2.data(), 0.75f, prm,
                                           (pm_pair_result *)(buf + off_res));
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(buf);
    if (e != cudaSuccess || e2 != cudaSuccess)
        return pm_fail(ctx, PM_CUDA_ERR, "pm_batch_warmup: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return st;
}

int pm_lmeds_score_dev(pm_ctx *ctx, const float *dp1, const float *dp2, int n, const float *dF32, int n_models, float *dmedians)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, n >= 0 && n_models >= 0, "bad argument");
    return pmk_lmeds_score(ctx, dp1, dp2, n, dF32, n_models, dmedians);
}

int pm_find_fundamental_lmeds(pm_ctx *ctx, const float *p1, const float *p2, int n, int n_hyp, const int32_t *sample_idx,
                              uint64_t seed, double F[9], uint8_t *mask, int *n_inliers, float *median_out)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, p1 && p2 && F && n >= 0 && n_hyp >= 0, "bad argument");
    if (n_inliers) *n_inliers = 0;
    if (n < 8 || n_hyp == 0) return PM_EMPTY;
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int nm = n_hyp * 3;
    PM_WS(ctx, dp1, float *, WS_P1, (size_t)n * 8);
    PM_WS(ctx, dp2, float *, WS_P2, (size_t)n * 8);
    PM_WS(ctx, ds, int32_t *, WS_SAMPLES, (size_t)n_hyp * 7 * 4);
    PM_WS(ctx, dF32, float *, WS_F32, ((size_t)nm + 1) * 12 * 4);
    PM_WS(ctx, dmed, float *, WS_COUNTS, (size_t)nm * 4);
    PM_WS(ctx, dkey, uint64_t *, WS_KEY, 64);
    PM_WS(ctx, dmask, uint8_t *, WS_MASK, (size_t)n);
    PM_WS(ctx, dFout, double *, WS_FOUT, 16 * 8);
    float *dFw = dF32 + (size_t)nm * 12;
    int32_t *dninl = reinterpret_cast<int32_t *>(dkey + 2);
    H2D(ctx, dp1, p1, (size_t)n * 8);
    H2D(ctx, dp2, p2, (size_t)n * 8);
    std::vector<int32_t> gen;
    const int32_t *hs = sample_idx;
    if (!hs) {
        gen.resize((size_t)n_hyp * 7);
        pm_make_sample_sets(n, n_hyp, 7, seed, gen.data());
        hs = gen.data();
    }
    H2D(ctx, ds, hs, (size_t)n_hyp * 7 * 4);
    if (!sample_idx) PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int st;
    if ((st = pmk_ransac_solve(ctx, dp1, dp2, n, ds, n_hyp, 7, dF32)) != PM_OK) return st;
    if ((st = pmk_lmeds_score(ctx, dp1, dp2, n, dF32, nm, dmed)) != PM_OK) return st;
    if ((st = pmk_lmeds_finish(ctx, dp1, dp2, n, dF32, dmed, nm, dFw, dFout, dmask, dninl, dkey)) != PM_OK) return st;
    D2H(ctx, ctx->h_pinned, dkey, 32);
    D2H(ctx, ctx->h_pinned + 16, dFout, 72);
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint64_t key = *reinterpret_cast<const uint64_t *>(ctx->h_pinned);
    if (key == ~0ull) return PM_EMPTY;
    memcpy(F, ctx->h_pinned + 16, 72);
    if (n_inliers) *n_inliers = ctx->h_pinned[4];
    if (median_out) { const uint32_t b = (uint32_t)(key >> 32); memcpy(median_out, &b, 4); }
    if (mask) { D2H(ctx, mask, dmask, (size_t)n); PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); }
    return PM_OK;
}

int pm_fundamental_8point(pm_ctx *ctx, const float *p1, const float *p2, int n, double F[9])
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, p1 && p2 && F && n >= 0, "bad argument");
    if (n < 8) return PM_EMPTY;
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    PM_WS(ctx, dp1, float *, WS_P1, (size_t)n * 8);
    PM_WS(ctx, dp2, float *, WS_P2, (size_t)n * 8);
    PM_WS(ctx, dFout, double *, WS_FOUT, 16 * 8);
    int32_t *dok = reinterpret_cast<int32_t *>(dFout + 12);
    H2D(ctx, dp1, p1, (size_t)n * 8);
    H2D(ctx, dp2, p2, (size_t)n * 8);
    int st = pmk_fundamental_npoint(ctx, dp1, dp2, n, nullptr, dFout, dok);
    if (st != PM_OK) return st;
    D2H(ctx, ctx->h_pinned, dFout, 13 * 8);
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_pinned[24] == 0) return PM_EMPTY;
    memcpy(F, ctx->h_pinned, 72);
    return PM_OK;
}

int pm_epilines(pm_ctx *ctx, const float *pts, int n, int which, const double F[9], float *lines)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, n >= 0 && (which == 1 || which == 2) && F, "bad argument");
    if (n == 0) return PM_OK;
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    PM_WS(ctx, dp, float *, WS_P1, (size_t)n * 8);
    PM_WS(ctx, dF, double *, WS_FOUT, 16 * 8);
    PM_WS(ctx, dl, float *, WS_LINES, (size_t)n * 12);
    H2D(ctx, dp, pts, (size_t)n * 8);
    memcpy(ctx->h_pinned, F, 72);
    H2D(ctx, dF, ctx->h_pinned, 72);
    int st = pmk_epilines(ctx, dp, n, which, dF, dl);
    if (st != PM_OK) return st;
    D2H(ctx, lines, dl, (size_t)n * 12);
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PM_OK;
}

int pm_residuals(pm_ctx *ctx, const float *p1, const float *p2, int n, const double F[9], int metric, float *out, double *mean_out)
{
    if (!ctx) return PM_BAD_ARG;
    PM_NVTX();
    PM_REQUIRE(ctx, n >= 0 && F && (metric == PM_METRIC_SAMPSON || metric == PM_METRIC_SYMEPI), "bad argument");
    if (mean_out) *mean_out = 0;
    if (n == 0) return PM_OK;
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    PM_WS(ctx, dp1, float *, WS_P1, (size_t)n * 8);
    PM_WS(ctx, dp2, float *, WS_P2, (size_t)n * 8);
    PM_WS(ctx, dF, double *, WS_FOUT, 16 * 8);
    PM_WS(ctx, dl, float *, WS_LINES, (size_t)n * 12);
    H2D(ctx, dp1, p1, (size_t)n * 8);
    H2D(ctx, dp2, p2, (size_t)n * 8);
    memcpy(ctx->h_pinned, F, 72);
    H2D(ctx, dF, ctx->h_pinned, 72);
    int st = pmk_residuals(ctx, dp1, dp2, n, dF, metric, dl, dF + 10);
    if (st != PM_OK) return st;
    if (out) D2H(ctx, out, dl, (size_t)n * 4);
    D2H(ctx, ctx->h_pinned + 32, dF + 10, 8);
    PM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (mean_out) *mean_out = *reinterpret_cast<const double *>(ctx->h_pinned + 32) / n;
    return PM_OK;
}

}  // extern "C"

// Column side of a cross-check (BFMatcher crossCheck = true: (i, j) stays iff i is also the nearest query of train row j).
// Only train rows that are the best match of some query can ever be asked about, so the reverse pass runs over THOSE
// rows only: at most (number of queries over all ranks) of the nt rows -- one eighth of them at BASELINE config 3.
// Falls back to the reverse pass over the whole train set when more than 3/4 of the rows are marked.
static int g_cross_full = 0;       // debug / test switch: always the full reverse pass
extern "C" void pm_debug_cross_full(int on) { g_cross_full = on; }

int pmk_cross_col_best(pm_ctx *ctx, int hamming, const void *dq, int nq, const void *dt, int nt, int width, int q_index_base,
                       const pm_dmatch *dknn, uint64_t *dcol_best, pm_mark_reduce_fn reduce_marks)
{
    if (nt <= 0) return PM_OK;
    const size_t row_bytes = (size_t)width * (hamming ? 1 : 4);
    int st, n_marked = nt;
    PM_WS(ctx, mark, uint8_t *, WS_X_MARK, (size_t)nt + 64);
    PM_WS(ctx, list, int32_t *, WS_X_LIST, (size_t)nt * 4 + 64);
    int32_t *count = list + nt;
    const bool bound_says_full = !reduce_marks && (long long)(nq < nt ? nq : nt) * 4 > (long long)nt * 3;
    if (!g_cross_full && !bound_says_full) {
        if ((st = pmk_cross_mark(ctx, dknn, nq, 2, nt, mark)) != PM_OK) return st;
        if (reduce_marks && (st = reduce_marks(ctx, mark, (size_t)nt)) != PM_OK) return st;
        if ((st = pmk_cross_list(ctx, mark, nt, list, count)) != PM_OK) return st;
        if (reduce_marks) {
            // several ranks: the number of marked rows (queries of ALL ranks) is only known on the device
            if ((st = read_count(ctx, count, &n_marked)) != PM_OK) return st;
        } else {
            n_marked = nq < nt ? nq : nt;            // one rank: at most one marked row per query -- no read-back, the
        }                                            // kernels below take the exact count from the device
    }
    if (nq <= 0) {                                   // an empty shard loses every minimum
        PM_CUDA(ctx, cudaMemsetAsync(dcol_best, 0xFF, (size_t)nt * 8, ctx->stream));
        return PM_OK;
    }
    if (g_cross_full || (long long)n_marked * 4 > (long long)nt * 3)
        return hamming ? pmk_hamming_col_best(ctx, (const uint8_t *)dq, nq, (const uint8_t *)dt, nt, width, q_index_base, dcol_best)
                       : pmk_l2_col_best(ctx, (const float *)dq, nq, (const float *)dt, nt, width, q_index_base, dcol_best);
    if (n_marked > 0) {
        PM_WS(ctx, rows, uint8_t *, WS_X_ROWS, (size_t)n_marked * row_bytes);
        PM_WS(ctx, small_, uint64_t *, WS_X_COL, (size_t)n_marked * 8);
        if ((st = pmk_cross_gather_rows(ctx, dt, row_bytes, list, count, n_marked, rows)) != PM_OK) return st;
        st = hamming ? pmk_hamming_col_best(ctx, (const uint8_t *)dq, nq, rows, n_marked, width, q_index_base, small_)
                     : pmk_l2_col_best(ctx, (const float *)dq, nq, (const float *)rows, n_marked, width, q_index_base, small_);
        if (st != PM_OK) return st;
        return pmk_cross_scatter(ctx, list, count, n_marked, small_, dcol_best, nt);
    }
    return pmk_cross_scatter(ctx, list, count, 0, nullptr, dcol_best, nt);
}

