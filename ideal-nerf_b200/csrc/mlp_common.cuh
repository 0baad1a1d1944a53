// Shared declarations of the FaceNeRF MLP kernels.
#pragma once
#include "common.cuh"

namespace inerf {

// Training-time stores (point-major rows).  acts: h0..h7 (8x256) | v0..v2 (3x128) | gamma(p) (64) | gamma(v) (32)
constexpr int SAVE_W = 2528, SAVE_PE = 2432, SAVE_DIR = 2496;
// deltas: gradient w.r.t. every pre-activation, same column map as the first 2432 columns of acts
constexpr int DELTA_W = 2432;

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
    float* save;                      // optional (fp32 path, training): activation store [ceil(P/64)*64][SAVE_W]
    float* trace;                     // optional (bf16 path): post-activation values of the first 256 points, [11][256][256]
};

int mlp_fp32_launch(const MlpArgs& a, bool embedded, cudaStream_t st);
int mlp_bf16_launch(const MlpArgs& a, bool embedded, cudaStream_t st);
int mlp_fp32_bwd_launch(const InerfNetDims* dims, const float* const* params_host, float* const* grads_host,
                        const float* aud, const float* expr, const float* latent, const float* acts, float* deltas,
                        const float* d_raw, long long P, float* d_cond, void* dw_args_dev, cudaStream_t st);
size_t mlp_fp32_bwd_args_bytes();
int mlp_bf16_hang_info(int32_t* out8);
int mlp_bf16_packed_bytes(const InerfNetDims* d, size_t* bytes);
int mlp_bf16_pack(const InerfNetDims* d, const float* const* params_host, void* packed, cudaStream_t st);

}  // namespace inerf
