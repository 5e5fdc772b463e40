// l2_common.h -- types shared by the L2 kernels (K1 pack, K2 tensor-core GEMM, K3 finish).
#pragma once
#include <cstdint>

// Packed operand row: [hi bf16 x128 | lo bf16 x128] = 512 B (K padded to 128).
#define L2_PACK_COLS 256
#define L2_KDIM 128

struct L2Cand {        // one candidate: approximate (||b||^2 - 2ab) and train index
    float d;
    int idx;
};

struct L2Flags {
    int nonexact;              // !=0: some value is not an integer in [0,255] -> split mode
    unsigned max_tnorm_bits;   // max ||b||^2 over real train rows (float bits)
    int n_flagged;             // rows K3 could not certify -> exact fallback
    int pad;
};

struct pm_ctx;
int l2_tc_grid(pm_ctx *ctx, int MT, int NT);
int l2_tc_smax(pm_ctx *ctx, int MT, int NT);
int l2_tc_launch(pm_ctx *ctx, const void *qpack, int mq_pad, const void *tpack, int nt_pad, const float *tnorm,
                 const L2Flags *flags, L2Cand *part, int smax, float *dump);
