// Shared declarations of the FaceNeRF MLP kernels.
#pragma once
#include "common.cuh"

namespace inerf {

// Kernel-side view of one FaceNeRF call.
struct MlpArgs {
    const float* w[INERF_N_PARAMS];   // nn.Linear layout (out,in), device pointers
    const float* cond;                // folded biases (CondLayout)
    const void* packed;               // mode-specific packed weights (bf16 path)
    // fused-PE input: points p = o + d*z
    const float* rays; int ray_stride;
    const float* z; int s;
    // embedded input (FaceNeRF.forward signature): x (P, 90)
    const float* x;
    long long P;                      // number of points
    float* out;                       // (P,4)
    int cond_dim;                     // dim_aud + dim_expr + dim_latent
    int dim_expr;
    float* trace;                     // optional (bf16 path): post-activation values of the first 256 points, [11][256][256]
};

int mlp_fp32_launch(const MlpArgs& a, bool embedded, cudaStream_t st);
int mlp_bf16_launch(const MlpArgs& a, bool embedded, cudaStream_t st);
int mlp_bf16_hang_info(int32_t* out8);
int mlp_bf16_packed_bytes(const InerfNetDims* d, size_t* bytes);
int mlp_bf16_pack(const InerfNetDims* d, const float* const* params_host, void* packed, cudaStream_t st);

}  // namespace inerf
