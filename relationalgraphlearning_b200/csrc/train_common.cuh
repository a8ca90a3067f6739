// Argument structs shared by the linear-layer backward kernels (train_kernels.cu: mma.sync; linear_bwd_tc.cu: tcgen05).
#pragma once
#include "common.cuh"

namespace rgl {

// row r of a logical [R, width] matrix: ptr + (r / rpg) * gstride + (r % rpg) * ld   (grouped rows: e.g. the robot
// row of every state inside a [B, n, 32] tensor is rpg = 1, gstride = n*32)
struct Rows {
    float* ptr;
    int ld;
    int rpg;
    long long gstride;
    __device__ __forceinline__ float* row(int r) const {
        if (rpg == 1) return ptr + (long long)r * gstride;              // plain matrix / one row per group: no division
        return ptr + (long long)(r / rpg) * gstride + (long long)(r % rpg) * ld;
    }
};

struct LinBwdArgs {
    Rows G, mask, Xin, Gin;
    int N, K, R;
    const float* W;       // optional (data gradient)
    int w_layout;         // 0: W is [N][K] (nn.Linear.weight), 1: W is [K][N] (w_a / Ws used as x @ W)
    int accumulate;       // Gin += instead of =
    float* dW;            // optional, same layout as W
    float* db;            // optional [N]
    int ntiles;
    int vecG, vecX;       // rows of G(+mask) / Xin are 16-byte aligned and N / K are multiples of 4
    int vecW;             // W is [N][K] with K a multiple of 4 and a 16-byte aligned base: staged with cp.async
    // fused first layer of a two-layer MLP (rgl_mlp2_bwd): x0 rows [R, K0], dW0 [K][K0] (nn.Linear layout), db0 [K]
    Rows X0;
    int K0;
    float* dW0;
    float* db0;
};

// tcgen05 form of the backward for the 32-wide layers (linear_bwd_tc.cu); cudaErrorNotSupported: shape / layout not covered
cudaError_t run_linear_bwd_tc(LinBwdArgs& a, int num_sms, size_t max_smem, cudaStream_t st);

}  // namespace rgl
