// pm_ubench.cu -- pm_measure_peak: measured issue peaks of the two CUDA-core pipes that bound the non-tensor kernels
// (BASELINE.md / SURVEY 8d: "measure them on the box", MEASURED_PEAKS.json has HBM and bf16 only):
//   which 0  FP32 FFMA   bounds K7, RANSAC scoring (ransac.cu)          -> TFLOP/s, 2 FLOP per FFMA
//   which 1  POPC.32     bounds K4a, the POPC Hamming kernel (hamming.cu) -> 1e12 POPC per second
// Independent dependency chains per thread (ILP 8) x 2048 resident threads per SM, so neither the 4-cycle FFMA latency
// nor the POPC pipe's latency limits the issue rate; one resident wave; a warm-up launch, then the best of five launches
// timed with CUDA events on the ctx stream.
#include "pm_internal.h"

namespace {

constexpr int UB_ILP = 8;
constexpr int UB_THREADS = 1024;

__global__ void __launch_bounds__(UB_THREADS) ub_ffma_kernel(float *out, int iters, float a, float b)
{
    float x[UB_ILP];
#pragma unroll
    for (int k = 0; k < UB_ILP; ++k) x[k] = (float)(threadIdx.x + k) * 1e-3f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < UB_ILP; ++k) x[k] = fmaf(x[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < UB_ILP; ++k) s += x[k];
    if (s == 123456.789f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // never true: keeps the chains alive
}

__global__ void __launch_bounds__(UB_THREADS) ub_popc_kernel(unsigned *out, int iters, unsigned m)
{
    unsigned x[UB_ILP], key[UB_ILP];
#pragma unroll
    for (int k = 0; k < UB_ILP; ++k) { x[k] = threadIdx.x * 2654435761u + k; key[k] = m * (k + 1) + threadIdx.x; }
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < UB_ILP; ++k) x[k] = __popc(x[k] ^ key[k]);     // one LOP3 (ALU pipe) + one POPC per step
    }
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < UB_ILP; ++k) s += x[k];
    if (s == 0xFFFFFFFFu) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int pm_measure_peak(pm_ctx *ctx, int which, double *value)
{
    if (!ctx) return PM_BAD_ARG;
    if (!value || which < 0 || which > 1) return pm_fail(ctx, PM_BAD_ARG, "pm_measure_peak: which must be 0 (FFMA) or 1 (POPC)");
    PM_CUDA(ctx, cudaSetDevice(ctx->device));
    PM_WS(ctx, dout, float *, WS_MISC, (size_t)2 * ctx->num_sms * UB_THREADS * 4);
    const int grid = 2 * ctx->num_sms;                   // 2048 threads per SM: one full resident wave
    const int iters = which == 0 ? 4096 : 1024;          // ~1 ms per launch either way
    cudaEvent_t e0, e1;
    PM_CUDA(ctx, cudaEventCreate(&e0));
    PM_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0, ctx->stream);
        if (which == 0) ub_ffma_kernel<<<grid, UB_THREADS, 0, ctx->stream>>>(dout, iters, 0.999f, 0.001f);
        else ub_popc_kernel<<<grid, UB_THREADS, 0, ctx->stream>>>((unsigned *)dout, iters, 0x9E3779B9u);
        ctx->launches++;
        cudaEventRecord(e1, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e != cudaSuccess) {
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            return pm_fail(ctx, PM_CUDA_ERR, "pm_measure_peak: %s", cudaGetErrorString(e));
        }
        if (rep > 0 && ms < best) best = ms;             // rep 0 is the warm-up
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double ops = (double)grid * UB_THREADS * (double)iters * 8.0 * UB_ILP;
    *value = (which == 0 ? 2.0 * ops : ops) / (best * 1e-3) / 1e12;
    return PM_OK;
}
