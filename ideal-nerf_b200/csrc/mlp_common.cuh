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
    // optional (bf16 path, training): per 128-point tile T = point/128, the activations as the forward kernel holds them in shared
    // memory -- TRAIN_IMGS 16 KB images [128 points][64 features] bf16, K-major, 128-byte swizzle -- and the ReLU masks
    uint8_t* save_img;                // [n_tiles][TRAIN_IMGS][16384]
    uint32_t* save_mask;              // [n_tiles][TRAIN_MASK_WORDS][128]: bit (31-j) of word w = pre-activation 32w+j of the row is >= 0
    float tc_comp;                    // fp16x2 kernel, calibration runs only: uniform main-accumulator scale - 1 (< 0: use tc_ulps)
    float tc_ulps[4];                 // fp16x2 kernel: main-accumulator compensation in ulps of 1.0 for {L0, L1-7, V0, V1-2}
};

// image index inside a tile: h0..h7 at 4l+kb, v0..v2 at 32+2v+kb, gamma(p) (63 cols) at 38, gamma(v) (27 cols) at 39; the backward
// kernels store the deltas with the same map and d_raw (r,g,b,sigma in columns 0..3) at 38
constexpr int TRAIN_IMGS = 40, TRAIN_IMG_PE = 38, TRAIN_IMG_DIR = 39, TRAIN_IMG_DOUT = 38, TRAIN_MASK_WORDS = 76;
__host__ __device__ constexpr int train_img_of(int l) { return l < 8 ? 4 * l : 32 + 2 * (l - 8); }
__host__ __device__ constexpr int train_mask_of(int l) { return l < 8 ? 8 * l : 64 + 4 * (l - 8); }

int mlp_fp32_launch(const MlpArgs& a, bool embedded, cudaStream_t st);
int mlp_bf16_launch(const MlpArgs& a, bool embedded, cudaStream_t st);     // a.save_img != NULL: the activation-saving build
int mlp_fp32_bwd_launch(const InerfNetDims* dims, const float* const* params_host, float* const* grads_host,
                        const float* aud, const float* expr, const float* latent, const float* acts, float* deltas,
                        const float* d_raw, long long P, float* d_cond, void* dw_args_dev, cudaStream_t st);
size_t mlp_fp32_bwd_args_bytes();
int mlp_f16x2_packed_bytes(const InerfNetDims* d, size_t* bytes);                                     // fp32-gate tensor-core mode (mlp_f16x2.cu)
int mlp_f16x2_pack(const InerfNetDims* d, const float* const* params_host, void* packed, cudaStream_t st);
int mlp_f16x2_launch(const MlpArgs& a, cudaStream_t st);
int mlp_bf16_hang_info(int32_t* out8);
void mlp_bf16_stage_offsets(uint32_t (*off)[2][5]);      // [11 layers][half][K-block index in issue order] -> byte offset in the packed blob
// bf16 training path (mlp_bf16_bwd.cu, mlp_bf16_dw.cu, mlp_fp32_bwd.cu)
int mlp_bf16_bwd_packed_bytes(const InerfNetDims* d, size_t* bytes);
int mlp_bf16_bwd_pack(const InerfNetDims* d, const float* const* params_host, void* packed, cudaStream_t st);
int mlp_bf16_bwd_chain_launch(const InerfNetDims* dims, const float* const* params_host, const void* packed_t, const uint32_t* mask,
                              const float* d_raw, uint8_t* delta_img, long long P, cudaStream_t st);
size_t mlp_bf16_dw_scratch_bytes();
int mlp_bf16_dw_launch(const InerfNetDims* dims, float* const* grads_host, const uint8_t* delta_img, const uint8_t* acts_img,
                       long long n_tiles, void* scratch, cudaStream_t st);
int mlp_bwd_cond_launch(const InerfNetDims* dims, const float* const* params_host, float* const* grads_host, const float* aud,
                        const float* expr, const float* latent, float* d_cond, cudaStream_t st);
int mlp_bf16_packed_bytes(const InerfNetDims* d, size_t* bytes);
int mlp_bf16_pack(const InerfNetDims* d, const float* const* params_host, void* packed, cudaStream_t st);

}  // namespace inerf
