// ubench.cu -- issue-rate microbenchmarks for the roofline denominators that
// MEASURED_PEAKS.json does not carry (B200, sm_100a): warp-instructions per clock per SM for
// the pipes the hot kernels lean on (integer/float min-max, IMAD/LEA, FFMA, POPC, LOP3, PRMT).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; run: ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

#define DEF_KERNEL(NAME, DECL, ...)                                                         \
    __global__ void __launch_bounds__(256) k_##NAME(uint32_t *out, long long *cycles, uint32_t seed) \
    {                                                                                       \
        DECL                                                                                \
        __syncthreads();                                                                    \
        const long long t0 = clock64();                                                     \
        _Pragma("unroll 1") for (int it = 0; it < ITERS; ++it) {                            \
            _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { __VA_ARGS__ }                   \
        }                                                                                   \
        const long long t1 = clock64();                                                     \
        uint32_t acc = 0;                                                                   \
        _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) acc ^= x[c] ^ y[c];              \
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc;                                   \
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;                                 \
    }

#define DECL_U32 uint32_t x[CHAINS], y[CHAINS]; \
    _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { x[c] = seed * (threadIdx.x + 1 + c); y[c] = seed ^ (c * 2654435761u + threadIdx.x); }

// two ops per body statement so the pair (min,max) keeps changing
DEF_KERNEL(vimnmx, DECL_U32, { uint32_t a, b; asm volatile("min.u32 %0, %2, %3;\n\tmax.u32 %1, %2, %3;" : "=r"(a), "=r"(b) : "r"(x[c]), "r"(y[c])); x[c] = a + 1; y[c] = b; })
DEF_KERNEL(iadd_only, DECL_U32, { x[c] = x[c] + 1; asm volatile("" : "+r"(x[c])); })
DEF_KERNEL(vimnmx3, DECL_U32, { uint32_t a; asm volatile("min.u32 %0, %1, %2;" : "=r"(a) : "r"(x[c]), "r"(y[c])); uint32_t b = __vimin3_u32(a, x[c] ^ 5u, y[c] + 3u); asm volatile("" : "+r"(b)); x[c] = b; })
DEF_KERNEL(fmnmx, DECL_U32, { uint32_t a, b; asm volatile("min.f32 %0, %2, %3;\n\tmax.f32 %1, %2, %3;" : "=r"(a), "=r"(b) : "r"(x[c]), "r"(y[c])); x[c] = a; y[c] = b; })
DEF_KERNEL(fmnmx3, DECL_U32, { uint32_t a; asm volatile("min.f32 %0, %1, %2, %3;" : "=r"(a) : "r"(x[c]), "r"(y[c]), "r"(seed)); x[c] = a; })
DEF_KERNEL(imad, DECL_U32, { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(seed), "r"(y[c])); })
DEF_KERNEL(lea, DECL_U32, { x[c] = (x[c] << 8) + y[c]; asm volatile("" : "+r"(x[c])); })
DEF_KERNEL(lop3, DECL_U32, { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(seed)); })
DEF_KERNEL(prmt, DECL_U32, { asm volatile("prmt.b32 %0, %0, %1, 0x3214;" : "+r"(x[c]) : "r"(y[c])); })
DEF_KERNEL(popc, DECL_U32, { uint32_t p; asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(x[c])); x[c] = p ^ y[c]; })
DEF_KERNEL(ffma, DECL_U32, { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y[c]), "r"(seed)); })
DEF_KERNEL(fadd, DECL_U32, { asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(x[c]) : "r"(y[c])); })
DEF_KERNEL(fsetp_sel, DECL_U32, { asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.b32 %0, %0, %1, p;}" : "+r"(x[c]) : "r"(y[c])); })
DEF_KERNEL(isetp_sel, DECL_U32, { asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.b32 %0, %0, %1, p;}" : "+r"(x[c]) : "r"(y[c])); })
DEF_KERNEL(vmin_u16x2, DECL_U32, { x[c] = __vminu2(x[c], y[c]) + 1; asm volatile("" : "+r"(x[c])); })
DEF_KERNEL(hmnmx2, DECL_U32, { asm volatile("min.f16x2 %0, %0, %1;" : "+r"(x[c]) : "r"(y[c])); x[c] += 1; })


// ---- TMEM read throughput: NW warps of one CTA per SM loop tcgen05.ld over a 512-column allocation
template <int X>
__device__ __forceinline__ void ldtm(uint32_t taddr, uint32_t &sink);
template <> __device__ __forceinline__ void ldtm<32>(uint32_t taddr, uint32_t &sink)
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
          "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= r[i];
}
template <> __device__ __forceinline__ void ldtm<8>(uint32_t taddr, uint32_t &sink)
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) sink ^= r[i];
}
template <int X>
__global__ void k_tmem(uint32_t *out, long long *cycles, int iters)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t sink = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 512 / X; c += 1) ldtm<X>(base + ((c * X + (warp >> 2) * X) & 511), sink);
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sink;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}
template <int X>
static void run_tmem(int nwarps, int nsm)
{
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, nsm * nwarps * 32 * 4); cudaMalloc(&cyc, nsm * 8);
    const int iters = 64;
    k_tmem<X><<<nsm, nwarps * 32>>>(out, cyc, iters);
    k_tmem<X><<<nsm, nwarps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long *h = new long long[nsm];
    cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nsm; ++i) avg += h[i]; avg /= nsm;
    const double bytes = (double)nwarps * iters * 512 * 32 * 4;     // per SM
    printf("tmem ld 32x32b.x%-3d warps=%2d : %7.1f B/clk/SM  (%.1f cyc per ld per warp)  err=%s\n", X, nwarps, bytes / avg,
           avg / (iters * (512 / X)), cudaGetErrorString(cudaGetLastError()));
    delete[] h; cudaFree(out); cudaFree(cyc);
}

template <typename K>
static void run(const char *name, K kernel, double ops_per_body, int nsm)
{
    const int blocks = nsm * 4, threads = 256;           // 32 warps per SM, all resident
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
    kernel<<<blocks, threads>>>(out, cyc, 12345u);
    kernel<<<blocks, threads>>>(out, cyc, 12345u);
    cudaDeviceSynchronize();
    long long *h = new long long[blocks];
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    // per SM: 4 blocks x 8 warps; warp-instructions per clock per SM
    const double winst = 4.0 * 8 * ITERS * CHAINS * ops_per_body;
    printf("%-12s %8.2f warp-inst/clk/SM  (%6.1f lanes/clk/SM)  cycles=%.0f  err=%s\n", name, winst / avg, 32 * winst / avg, avg,
           cudaGetErrorString(cudaGetLastError()));
    delete[] h; cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s  SMs=%d\n", p.name, p.multiProcessorCount);
    const int n = p.multiProcessorCount;
    run("iadd", k_iadd_only, 1, n);
    run("vimnmx(+iadd)", k_vimnmx, 3, n);
    run("vimnmx3(+..)", k_vimnmx3, 5, n);
    run("fmnmx", k_fmnmx, 2, n);
    run("fmnmx3", k_fmnmx3, 1, n);
    run("imad", k_imad, 1, n);
    run("lea/shl-add", k_lea, 1, n);
    run("lop3", k_lop3, 1, n);
    run("prmt", k_prmt, 1, n);
    run("popc(+lop)", k_popc, 2, n);
    run("ffma", k_ffma, 1, n);
    run("fadd", k_fadd, 1, n);
    run("fsetp+sel", k_fsetp_sel, 2, n);
    run("isetp+sel", k_isetp_sel, 2, n);
    run("vminu2(+add)", k_vmin_u16x2, 2, n);
    run("hmnmx2(+add)", k_hmnmx2, 2, n);
    for (int nw : {4, 8, 16}) { run_tmem<32>(nw, n); run_tmem<8>(nw, n); }
    return 0;
}
