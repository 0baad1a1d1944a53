// Device-side pieces of ray generation and of the stratified depths, shared by rays.cu and the set-up kernel of
// inerf_render_rays_fused (render_fused.cu).  Every operation that feeds the bit-exact gates uses explicit round-to-nearest intrinsics.
// Reference: NeRFs/HeadNeRF/helper.py:228-243, NeRFs/HeadNeRF/train/audio_exp_nerf.py:306-328,409-427.
#pragma once
#include "common.cuh"

namespace inerf {

__device__ __forceinline__ void store_ray(float* __restrict__ r, float ox, float oy, float oz, float dx, float dy,
                                          float dz, float near_, float far_) {
    // viewdirs = d / ||d||_2   (torch.norm = sqrt(sum of squares), then a true division)
    float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    r[0] = ox; r[1] = oy; r[2] = oz;
    r[3] = dx; r[4] = dy; r[5] = dz;
    r[6] = near_; r[7] = far_;
    r[8] = __fdiv_rn(dx, nrm); r[9] = __fdiv_rn(dy, nrm); r[10] = __fdiv_rn(dz, nrm);
}

// direction of pixel (row, col) in world space: camera-frame ((i-cx)/f, -(j-cy)/f, -1) through c2w[:3,:3]
__device__ __forceinline__ void pixel_dir(float row, float col, float focal, float cx, float cy, const float* __restrict__ c2w, int rs, float (&d)[3]) {
    float c0 = __fdiv_rn(__fsub_rn(col, cx), focal);
    float c1 = -__fdiv_rn(__fsub_rn(row, cy), focal);
    float c2 = -1.0f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        // torch.sum(dirs[..., None, :] * c2w[:3,:3], -1): products rounded, added left to right
        float a = __fmul_rn(c0, c2w[r * rs + 0]);
        float b = __fmul_rn(c1, c2w[r * rs + 1]);
        float c = __fmul_rn(c2, c2w[r * rs + 2]);
        d[r] = __fadd_rn(__fadd_rn(a, b), c);
    }
}

__device__ __forceinline__ float coarse_z(float near_, float far_, float t, int lindisp) {
    if (!lindisp)   // near * (1. - t) + far * t
        return __fadd_rn(__fmul_rn(near_, __fsub_rn(1.0f, t)), __fmul_rn(far_, t));
    // 1. / (1. / near * (1. - t) + 1. / far * t)
    float a = __fmul_rn(__fdiv_rn(1.0f, near_), __fsub_rn(1.0f, t));
    float b = __fmul_rn(__fdiv_rn(1.0f, far_), t);
    return __fdiv_rn(1.0f, __fadd_rn(a, b));
}

}  // namespace inerf
