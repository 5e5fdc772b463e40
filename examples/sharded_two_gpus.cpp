// sharded_two_gpus.cpp -- a C++ host driving the multi-GPU entry points of include/pm.h: one host thread per GPU,
// one pm_ctx per thread, NCCL inside the C ABI (pm_comm_init).  The two sharded calls stand in for the two OpenCV
// calls of /root/reference/Points Matching/main.cpp on an N-GPU box:
//   BFMatcher(NORM_HAMMING, crossCheck=true).match (main.cpp:43-46) -> pm_match_cross_sharded_dev + pm_allgather_matches_dev
//   cv::findFundamentalMat(RANSAC)                  (main.cpp:95-98) -> pm_find_fundamental_sharded_dev
// and are checked against the same calls on ONE GPU: sharding must not change a single bit.
//
//   sharded_two_gpus [n_gpus]      exit 0: identical results, 3: fewer GPUs than asked, 1: mismatch / error
#include <cuda_runtime_api.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "pm.h"

#define CK(call) do { int st__ = (call); if (st__ != PM_OK) { std::fprintf(stderr, "%s -> %d: %s\n", #call, st__, ctx ? pm_last_error(ctx) : ""); ok = false; std::_Exit(1); } } while (0)
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e__)); ok = false; return; } } while (0)

static uint64_t rng_state = 0x1234567ull;
static uint32_t rnd() { rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng_state >> 33); }

struct RankOut {
    std::vector<pm_dmatch> all; std::vector<int32_t> counts;     // gathered mutual matches
    double F[9]; std::vector<uint8_t> mask; int32_t n_inl; uint64_t key;
};

int main(int argc, char **argv)
{
    const int R = argc > 1 ? std::atoi(argv[1]) : 2;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < R) { std::printf("needs %d GPUs, %d visible: skipped\n", R, ndev); return 3; }
    const int nq = 6001, nt = 5000, bytes = 32, npts = 4000, nhyp = 3001;
    std::vector<uint8_t> q((size_t)nq * bytes), t((size_t)nt * bytes);
    for (auto &b : t) b = (uint8_t)rnd();
    for (int i = 0; i < nq; ++i) {                       // half the queries are noisy copies of a train row
        const bool planted = i % 2 == 0;
        const int j = (int)(rnd() % nt);
        for (int k = 0; k < bytes; ++k) q[(size_t)i * bytes + k] = planted ? t[(size_t)j * bytes + k] : (uint8_t)rnd();
        if (planted) for (int f = 0; f < 20; ++f) { const uint32_t bit = rnd() % (bytes * 8); q[(size_t)i * bytes + bit / 8] ^= (uint8_t)(1u << (bit % 8)); }
    }
    // correspondences of a pure sideways translation (F = [t]x) plus 40% outliers
    std::vector<float> p1((size_t)npts * 2), p2((size_t)npts * 2);
    for (int i = 0; i < npts; ++i) {
        const float x = (float)(rnd() % 19200) * 0.1f, y = (float)(rnd() % 10800) * 0.1f, d = 5.f + (float)(rnd() % 400) * 0.1f;
        p1[2 * i] = x; p1[2 * i + 1] = y;
        if (i % 5 < 3) { p2[2 * i] = x + d; p2[2 * i + 1] = y + ((float)(rnd() % 100) - 50.f) * 0.004f; }
        else { p2[2 * i] = (float)(rnd() % 19200) * 0.1f; p2[2 * i + 1] = (float)(rnd() % 10800) * 0.1f; }
    }
    pm_ransac_params prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.sample_size = 8; prm.metric = PM_METRIC_SAMPSON; prm.threshold = 1.f; prm.refit = 1; prm.seed = 77; prm.sample_idx = nullptr;

    unsigned char id[PM_COMM_ID_BYTES];
    if (pm_comm_unique_id(id) != PM_OK) { std::fprintf(stderr, "libnccl.so.2 not loadable\n"); return 1; }
    std::vector<RankOut> out((size_t)R);
    std::vector<char> oks((size_t)R, 1);
    auto rank_main = [&](int r) {
        bool ok = true;
        RankOut &o = out[(size_t)r];
        [&] {
            pm_ctx *ctx = nullptr;
            CK(pm_create(&ctx, r));
            CK(pm_comm_init(ctx, R, r, id));
            const int lo = (int)((long long)nq * r / R), hi = (int)((long long)nq * (r + 1) / R), mine = hi - lo, width = (nq + R - 1) / R;
            uint8_t *dq, *dt, *dmask; pm_dmatch *dlocal, *dall; uint64_t *dcol, *dkey; int32_t *dn, *dcounts, *dninl; float *dp1, *dp2; double *dF;
            CU(cudaMalloc((void **)&dq, (size_t)mine * bytes)); CU(cudaMalloc((void **)&dt, (size_t)nt * bytes));
            CU(cudaMalloc((void **)&dlocal, (size_t)width * sizeof(pm_dmatch))); CU(cudaMalloc((void **)&dall, (size_t)R * width * sizeof(pm_dmatch)));
            CU(cudaMalloc((void **)&dcol, (size_t)nt * 8)); CU(cudaMalloc((void **)&dn, 4)); CU(cudaMalloc((void **)&dcounts, (size_t)R * 4));
            CU(cudaMalloc((void **)&dp1, (size_t)npts * 8)); CU(cudaMalloc((void **)&dp2, (size_t)npts * 8)); CU(cudaMalloc((void **)&dF, 72));
            CU(cudaMalloc((void **)&dmask, (size_t)npts)); CU(cudaMalloc((void **)&dninl, 4)); CU(cudaMalloc((void **)&dkey, 8));
            CU(cudaMemcpy(dq, q.data() + (size_t)lo * bytes, (size_t)mine * bytes, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(dt, t.data(), (size_t)nt * bytes, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(dp1, p1.data(), (size_t)npts * 8, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(dp2, p2.data(), (size_t)npts * 8, cudaMemcpyHostToDevice));
            // cross-check: one min-reduce of the packed column minima, then the gather of the survivors
            CK(pm_match_cross_sharded_dev(ctx, dq, mine, dt, nt, bytes, 6, lo, nullptr, dcol, dlocal, dn));
            CK(pm_allgather_matches_dev(ctx, dlocal, dn, width, dall, dcounts));
            // RANSAC-F: this rank's slice of the hypotheses, one 8-byte max-reduce, local re-solve of the winner
            pm_ransac_params mine_prm = prm;
            mine_prm.hyp_id_base = (int)((long long)nhyp * r / R);
            mine_prm.n_hyp = (int)((long long)nhyp * (r + 1) / R) - mine_prm.hyp_id_base;
            CK(pm_find_fundamental_sharded_dev(ctx, dp1, dp2, npts, &mine_prm, nhyp, dF, dmask, dninl, dkey));
            CK(pm_sync(ctx));
            o.all.resize((size_t)R * width); o.counts.resize((size_t)R); o.mask.resize((size_t)npts);
            CU(cudaMemcpy(o.all.data(), dall, o.all.size() * sizeof(pm_dmatch), cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(o.counts.data(), dcounts, (size_t)R * 4, cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(o.F, dF, 72, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(o.mask.data(), dmask, (size_t)npts, cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(&o.n_inl, dninl, 4, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(&o.key, dkey, 8, cudaMemcpyDeviceToHost));
            for (void *p : {(void *)dq, (void *)dt, (void *)dlocal, (void *)dall, (void *)dcol, (void *)dn, (void *)dcounts, (void *)dp1,
                            (void *)dp2, (void *)dF, (void *)dmask, (void *)dninl, (void *)dkey}) cudaFree(p);
            pm_destroy(ctx);
        }();
        oks[(size_t)r] = ok;
    };
    std::vector<std::thread> th;
    for (int r = 0; r < R; ++r) th.emplace_back(rank_main, r);
    for (auto &x : th) x.join();
    for (int r = 0; r < R; ++r) if (!oks[(size_t)r]) return 1;

    // the same two calls on one GPU
    pm_ctx *ctx = nullptr;
    if (pm_create(&ctx, 0) != PM_OK) return 1;
    std::vector<pm_dmatch> ref((size_t)nq);
    int nref = 0;
    if (pm_match_cross_hamming(ctx, q.data(), nq, t.data(), nt, bytes, ref.data(), &nref) != PM_OK) return 1;
    prm.n_hyp = nhyp;
    double Fref[9]; std::vector<uint8_t> mref((size_t)npts); int ninl_ref = 0;
    if (pm_find_fundamental(ctx, p1.data(), p2.data(), npts, &prm, Fref, mref.data(), &ninl_ref) != PM_OK) return 1;
    pm_destroy(ctx);

    bool same = true;
    const int width = (nq + R - 1) / R;
    for (int r = 0; r < R && same; ++r) {                // every rank holds the same gathered list == the single-GPU list
        int k = 0;
        for (int s = 0; s < R && same; ++s)
            for (int i = 0; i < out[(size_t)r].counts[(size_t)s] && same; ++i, ++k)
                same = k < nref && std::memcmp(&out[(size_t)r].all[(size_t)s * width + i], &ref[(size_t)k], sizeof(pm_dmatch)) == 0;
        same = same && k == nref;
        same = same && std::memcmp(out[(size_t)r].F, Fref, 72) == 0 && out[(size_t)r].n_inl == ninl_ref &&
               std::memcmp(out[(size_t)r].mask.data(), mref.data(), (size_t)npts) == 0;
    }
    std::printf("ranks %d: mutual matches %d, RANSAC inliers %d of %d, winner key %llx -> %s\n", R, nref, ninl_ref, npts,
                (unsigned long long)out[0].key, same ? "identical to one GPU" : "MISMATCH");
    return same ? 0 : 1;
}
