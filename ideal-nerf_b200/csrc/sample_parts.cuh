// The importance sampler of the reference's configuration (N_samples = 64, N_importance = 128) with the draws made in the kernel, as a
// per-warp device function: shared by importance_rng_64_128_kernel (sample_pdf.cu, inerf_importance_sample_rng) and the fused
// compositor + sampler of inerf_render_rays_fused (render_fused.cu), which feeds it the coarse weights straight from registers.
// Reference: NeRFs/HeadNeRF/helper.py:269-313, NeRFs/HeadNeRF/train/audio_exp_nerf.py:342-347,364.
#pragma once
#include <math_constants.h>

#include "common.cuh"
#include "philox.cuh"

namespace inerf {

constexpr int IMP64_WARP_FLOATS = 64 + 64 + 64 + 128 + 192;      // shared-memory floats per warp (16-byte aligned)

// z2 / w2: the lane's two adjacent coarse depths / weights (samples 2 lane, 2 lane + 1).  Returns z_std of the ray (0 when z_std is NULL).
__device__ __forceinline__ float importance_rng_64_128_ray(const float2 z2, const float2 w2, int ray, int lane, float* __restrict__ wsm,
                                                          const unsigned long long* __restrict__ rng_state, uint32_t stream_id,
                                                          float* __restrict__ z_samples, float* __restrict__ z_merged, float* __restrict__ z_std) {
    constexpr int S1 = 64, NB = 63, NI = 128, TOT = 192;
    float* zc = wsm;                   // [64] coarse depths
    float* bsm = zc + 64;             // [63] bin mid-points (+1 pad)
    float* csm = bsm + 64;            // [63] cdf, [63] = +inf
    float* zsm = csm + 64;            // [128] samples
    float* osm = zsm + 128;           // [192] merged row

    reinterpret_cast<float2*>(zc)[lane] = z2;
    const float z_next = __shfl_down_sync(0xffffffffu, z2.x, 1), w_next = __shfl_down_sync(0xffffffffu, w2.x, 1);
    // pdf bins j = 2 lane, 2 lane + 1 of the 62 (weights[..., 1:-1]): w[j + 1]
    const float wa = lane < 31 ? w2.y + 1e-5f : 0.f, wb = lane < 31 ? w_next + 1e-5f : 0.f;
    float incl = wa + wb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const float inv_S = 1.0f / __shfl_sync(0xffffffffu, incl, 31);
    const float excl = incl - (wa + wb);
    bsm[2 * lane] = 0.5f * (z2.y + z2.x);
    if (lane < 31) {
        bsm[2 * lane + 1] = 0.5f * (z_next + z2.y);
        csm[2 * lane + 1] = (excl + wa) * inv_S;
        csm[2 * lane + 2] = (excl + wa + wb) * inv_S;
    } else {
        csm[63] = CUDART_INF_F;
    }
    if (lane == 0) csm[0] = 0.0f;

    // sorted uniforms: U_(k) = (E_1 + .. + E_k) / (E_1 + .. + E_129), draws k = 4 lane .. 4 lane + 3
    const Philox4 q = philox_at(rng_state, (uint64_t)ray * 33u + (uint64_t)lane, stream_id);
    float e[4] = {-__logf(u01_open0(q.x)), -__logf(u01_open0(q.y)), -__logf(u01_open0(q.z)), -__logf(u01_open0(q.w))};
    e[1] += e[0]; e[2] += e[1]; e[3] += e[2];
    float eincl = e[3];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, eincl, o);
        if (lane >= o) eincl += t;
    }
    const uint32_t spare = (q.x & 0xFFu) | ((q.y & 0xFFu) << 8) | ((q.z & 0xFFu) << 16) | (q.w << 24);      // bits the uniforms above do not use
    const float e_tail = -__logf(u01_open0(__shfl_sync(0xffffffffu, spare, 0)));
    const float inv_T = 1.0f / (__shfl_sync(0xffffffffu, eincl, 31) + e_tail);
    const float e_off = eincl - e[3];
    __syncwarp();

    // invert the CDF: c = #{j : cdf_j <= u} by bisection over the 64-entry padded table
    float u[4], zs[4];
    int c[4] = {0, 0, 0, 0};
#pragma unroll
    for (int t = 0; t < 4; ++t) u[t] = fminf((e[t] + e_off) * inv_T, 1.0f);
#pragma unroll
    for (int step = 32; step >= 1; step >>= 1)
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (csm[c[t] + step - 1] <= u[t]) c[t] += step;
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int below = max(c[t] - 1, 0), above = min(c[t], NB - 1);
        const float cb = csm[below], ca = csm[above];
        float den = ca - cb;
        if (den < 1e-5f) den = 1.0f;
        const float tt = (u[t] - cb) / den;
        const float bb = bsm[below], ba = bsm[above];
        zs[t] = bb + tt * (ba - bb);
        sum += zs[t];
    }
    reinterpret_cast<float4*>(zsm)[lane] = make_float4(zs[0], zs[1], zs[2], zs[3]);
    float std_v = 0.f;
    if (z_std) {
        const float mean = warp_sum(sum) * (1.0f / NI);
        float sq = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) { const float d = zs[t] - mean; sq += d * d; }
        sq = warp_sum(sq);
        std_v = sqrtf(sq * (1.0f / NI));
        if (lane == 0) z_std[ray] = std_v;
    }
    if (z_samples) reinterpret_cast<float4*>(z_samples + (size_t)ray * NI)[lane] = make_float4(zs[0], zs[1], zs[2], zs[3]);
    __syncwarp();

    // rank merge of the two ascending lists (ties: coarse depths first)
    {
        const float v[2] = {z2.x, z2.y};
        int r[2] = {0, 0};
#pragma unroll
        for (int step = 64; step >= 1; step >>= 1)
#pragma unroll
            for (int t = 0; t < 2; ++t)
                if (zsm[r[t] + step - 1] < v[t]) r[t] += step;
        const float last = zsm[NI - 1];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            r[t] += (last < v[t]) ? 1 : 0;                     // the bisection counts among the first 127 samples
            osm[2 * lane + t + r[t]] = v[t];
        }
    }
    {
        int r[4] = {0, 0, 0, 0};
#pragma unroll
        for (int step = 32; step >= 1; step >>= 1)
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (zc[r[t] + step - 1] <= zs[t]) r[t] += step;
        const float last = zc[S1 - 1];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            r[t] += (last <= zs[t]) ? 1 : 0;
            osm[4 * lane + t + r[t]] = zs[t];
        }
    }
    __syncwarp();
    float4* out = reinterpret_cast<float4*>(z_merged + (size_t)ray * TOT);
    out[lane] = reinterpret_cast<const float4*>(osm)[lane];
    if (lane < 16) out[32 + lane] = reinterpret_cast<const float4*>(osm)[32 + lane];
    return std_v;
}

}  // namespace inerf
