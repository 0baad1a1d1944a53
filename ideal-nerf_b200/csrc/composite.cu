// raw2outputs: alpha compositing with the background colour injected as the last sample.
//
// Reference: NeRFs/HeadNeRF/train/baseline.py:325-375 (head) and NeRFs/TorsoNeRF/test_torso.py:352-402
// (adds rgb_map_fg).  Backward = the reference's autograd of those lines, restated analytically.
//
// Mapping: one warp per ray; lane l owns samples l, l+32, ... so every global access is a fully
// coalesced 128 B (z, weights) or 512 B (raw, float4 per lane) request.  The exclusive cumprod is a
// warp shuffle scan carried across 32-sample chunks; the 5 (8 with rgb_fg) per-ray sums are reduced
// together with a 9-shuffle transpose-reduce instead of 8 separate 5-shuffle reductions.
// Algorithmic traffic: 24*S + 48 B per ray forward, 36*S + 60 B per ray backward.
#include "common.cuh"
#include "composite_parts.cuh"

using namespace inerf;

namespace {

template <int C>
__global__ void __launch_bounds__(256) composite_fwd_kernel(
    const float4* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d, int d_stride,
    const float* __restrict__ bc_rgb, const float* __restrict__ noise, int n, int s, int white_bkgd,
    float* __restrict__ rgb, float* __restrict__ disp, float* __restrict__ acc, float* __restrict__ depth,
    float* __restrict__ weights, float* __restrict__ rgb_fg) {
    const int lane = threadIdx.x & 31;
    const int ray = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ray >= n) return;
    const size_t base = (size_t)ray * s;

    // issue every load of the ray before any math
    float4 rv[C];
    float zv[C], nv[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        int i = c * 32 + lane;
        bool ok = i < s;
        rv[c] = ok ? ldg_stream4(raw + base + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        zv[c] = ok ? ldg_stream(z + base + i) : 0.f;
        nv[c] = (ok && noise) ? ldg_stream(noise + base + i) : 0.f;
    }
    const float* dptr = rays_d + (size_t)ray * d_stride;
    float dx = dptr[0], dy = dptr[1], dz = dptr[2];
    float norm = sqrtf(dx * dx + dy * dy + dz * dz);
    float bcr = bc_rgb[ray * 3], bcg = bc_rgb[ray * 3 + 1], bcb = bc_rgb[ray * 3 + 2];

    float carry = 1.0f;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < C; ++c) {
        int i = c * 32 + lane;
        bool valid = i < s, last = i == s - 1;
        float znext = __shfl_down_sync(0xffffffffu, zv[c], 1);
        if (c + 1 < C) {
            float z0 = __shfl_sync(0xffffffffu, zv[c + 1 < C ? c + 1 : c], 0);
            if (lane == 31) znext = z0;
        }
        Sample sm = make_sample(rv[c].w, nv[c], zv[c], znext, norm, last, valid);
        float incl = warp_scan_mul(sm.q, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        float T = carry * excl;
        carry *= __shfl_sync(0xffffffffu, incl, 31);
        float w = sm.alpha * T;
        if (valid) weights[base + i] = w;
        float cr = last ? bcr : sigmoidf_(rv[c].x);
        float cg = last ? bcg : sigmoidf_(rv[c].y);
        float cb = last ? bcb : sigmoidf_(rv[c].z);
        v[0] += w * cr; v[1] += w * cg; v[2] += w * cb;
        v[3] += w * zv[c];
        v[4] += w;
        if (!last) { v[5] += w * cr; v[6] += w * cg; v[7] += w * cb; }
    }
    float tot = reduce8(v, lane);                                  // lane 4q holds quantity q
    float acc_t = __shfl_sync(0xffffffffu, tot, 16);
    float depth_t = __shfl_sync(0xffffffffu, tot, 12);
    int q = lane >> 2;
    if ((lane & 3) == 0) {
        if (q < 3) rgb[ray * 3 + q] = white_bkgd ? tot + (1.0f - acc_t) : tot;
        else if (q == 3) {
            depth[ray] = depth_t;
            disp[ray] = __fdiv_rn(1.0f, fmaxf(1e-10f, __fdiv_rn(depth_t, acc_t)));
        } else if (q == 4) acc[ray] = acc_t;
        else if (rgb_fg) rgb_fg[ray * 3 + (q - 5)] = tot;
    }
}

// Forward for even S (every production shape: 64, 192): see composite2_ray in composite_parts.cuh.
template <int C>
__global__ void __launch_bounds__(256) composite_fwd2_kernel(
    const float4* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d, int d_stride,
    const float* __restrict__ bc_rgb, const float* __restrict__ noise, int n, int s, int white_bkgd,
    float* __restrict__ rgb, float* __restrict__ disp, float* __restrict__ acc, float* __restrict__ depth,
    float* __restrict__ weights, float* __restrict__ rgb_fg) {
    const int lane = threadIdx.x & 31;
    const int ray = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ray >= n) return;
    float2 z0, w0;
    composite2_ray<C>(raw, z, rays_d, d_stride, bc_rgb, noise, ray, s, white_bkgd, rgb, disp, acc, depth, weights, rgb_fg, nullptr, lane, z0, w0);
}

template <int C>
__global__ void __launch_bounds__(256) composite_bwd_kernel(
    const float4* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d, int d_stride,
    const float* __restrict__ bc_rgb, const float* __restrict__ noise, int n, int s, int white_bkgd,
    const float* __restrict__ g_rgb, const float* __restrict__ g_disp, const float* __restrict__ g_acc,
    const float* __restrict__ g_depth, const float* __restrict__ g_weights, const float* __restrict__ g_rgb_fg,
    float4* __restrict__ d_raw) {
    const int lane = threadIdx.x & 31;
    const int ray = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ray >= n) return;
    const size_t base = (size_t)ray * s;

    float4 rv[C];
    float zv[C], gw[C];
    Sample sm[C];
    float Tv[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        int i = c * 32 + lane;
        bool ok = i < s;
        rv[c] = ok ? ldg_stream4(raw + base + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        zv[c] = ok ? ldg_stream(z + base + i) : 0.f;
        gw[c] = (ok && g_weights) ? ldg_stream(g_weights + base + i) : 0.f;
    }
    const float* dptr = rays_d + (size_t)ray * d_stride;
    float dx = dptr[0], dy = dptr[1], dz = dptr[2];
    float norm = sqrtf(dx * dx + dy * dy + dz * dz);
    float bcr = bc_rgb[ray * 3], bcg = bc_rgb[ray * 3 + 1], bcb = bc_rgb[ray * 3 + 2];
    float gr = g_rgb ? g_rgb[ray * 3] : 0.f, gg = g_rgb ? g_rgb[ray * 3 + 1] : 0.f, gb = g_rgb ? g_rgb[ray * 3 + 2] : 0.f;
    float fr = g_rgb_fg ? g_rgb_fg[ray * 3] : 0.f, fg = g_rgb_fg ? g_rgb_fg[ray * 3 + 1] : 0.f,
          fb = g_rgb_fg ? g_rgb_fg[ray * 3 + 2] : 0.f;

    // forward recompute: alpha, T and the two totals the disparity gradient needs
    float carry = 1.0f, accp = 0.f, depp = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        int i = c * 32 + lane;
        bool valid = i < s, last = i == s - 1;
        float znext = __shfl_down_sync(0xffffffffu, zv[c], 1);
        if (c + 1 < C) {
            float z0 = __shfl_sync(0xffffffffu, zv[c + 1 < C ? c + 1 : c], 0);
            if (lane == 31) znext = z0;
        }
        float nz = (valid && noise) ? noise[base + i] : 0.f;
        sm[c] = make_sample(rv[c].w, nz, zv[c], znext, norm, last, valid);
        float incl = warp_scan_mul(sm[c].q, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        Tv[c] = carry * excl;
        carry *= __shfl_sync(0xffffffffu, incl, 31);
        float w = sm[c].alpha * Tv[c];
        accp += w;
        depp += w * zv[c];
    }
    float acc_t = warp_sum(accp), depth_t = warp_sum(depp);

    // disp = 1 / max(1e-10, depth/acc)
    float gd_tot = g_depth ? g_depth[ray] : 0.f;
    float ga_tot = g_acc ? g_acc[ray] : 0.f;
    if (g_disp) {
        float r = depth_t / acc_t;
        if (r > 1e-10f) {
            float k = -g_disp[ray] / (r * r);          // d disp / d r
            gd_tot += k / acc_t;
            ga_tot += -k * depth_t / (acc_t * acc_t);
        }
    }
    if (white_bkgd) ga_tot -= gr + gg + gb;

    // reverse pass: d alpha_i = G_i T_i - (1/q_i) sum_{k>i} G_k w_k
    float suffix = 0.f;
#pragma unroll
    for (int c = C - 1; c >= 0; --c) {
        int i = c * 32 + lane;
        bool valid = i < s, last = i == s - 1;
        float cr = last ? bcr : sigmoidf_(rv[c].x);
        float cg = last ? bcg : sigmoidf_(rv[c].y);
        float cb = last ? bcb : sigmoidf_(rv[c].z);
        float w = sm[c].alpha * Tv[c];
        float G = gr * cr + gg * cg + gb * cb + gd_tot * zv[c] + ga_tot + gw[c];
        if (!last) G += fr * cr + fg * cg + fb * cb;
        float x = valid ? G * w : 0.f;
        float incl = warp_suffix_add(x, lane);
        float excl = __shfl_down_sync(0xffffffffu, incl, 1);
        if (lane == 31) excl = 0.f;
        float R = suffix + excl;
        suffix += __shfl_sync(0xffffffffu, incl, 0);
        float d_alpha = G * Tv[c] - R / sm[c].q;
        // alpha = 1 - exp(-(relu(s)+1e-6) dist)  =>  d alpha / d s = dist * exp(.) * [s > 0]
        float e = 1.0f - sm[c].alpha;
        float d_sigma = (sm[c].s > 0.f) ? d_alpha * sm[c].dist * e : 0.f;
        float4 o;
        if (last) {
            o.x = o.y = o.z = 0.f;                     // colour of the last sample is bc_rgb
        } else {
            o.x = w * (gr + fr) * cr * (1.0f - cr);
            o.y = w * (gg + fg) * cg * (1.0f - cg);
            o.z = w * (gb + fb) * cb * (1.0f - cb);
        }
        o.w = d_sigma;
        if (valid) d_raw[base + i] = o;
    }
}

int pick_chunks(int s) {
    const int opts[] = {1, 2, 4, 6, 8, 16, 32};
    int need = (s + 31) / 32;
    for (int o : opts)
        if (o >= need) return o;
    return -1;
}

}  // namespace

#define DISPATCH_C(C_, CALL)            \
    switch (C_) {                       \
        case 1: { constexpr int C = 1; CALL; } break;   \
        case 2: { constexpr int C = 2; CALL; } break;   \
        case 4: { constexpr int C = 4; CALL; } break;   \
        case 6: { constexpr int C = 6; CALL; } break;   \
        case 8: { constexpr int C = 8; CALL; } break;   \
        case 16: { constexpr int C = 16; CALL; } break; \
        default: { constexpr int C = 32; CALL; } break; \
    }

extern "C" int inerf_composite_fwd(const float* raw, const float* z, const float* rays_d, int rays_d_stride,
                                   const float* bc_rgb, const float* noise, int n, int s, int white_bkgd, float* rgb,
                                   float* disp, float* acc, float* depth, float* weights, float* rgb_fg, void* stream) {
    if (n < 0 || s <= 0 || s > 1024 || rays_d_stride < 3) return fail(INERF_E_SHAPE, "inerf_composite_fwd: bad n/s/stride (1 <= s <= 1024)");
    if (n == 0) return INERF_OK;
    if (!raw || !z || !rays_d || !bc_rgb || !rgb || !disp || !acc || !depth || !weights)
        return fail(INERF_E_ARG, "inerf_composite_fwd: NULL pointer");
    if ((uintptr_t)raw & 15) return fail(INERF_E_ALIGN, "inerf_composite_fwd: raw must be 16-byte aligned");
    dim3 grid((n + 7) / 8), block(256);
    const bool pairs = (s % 2 == 0) && s <= 512 && !(((uintptr_t)raw & 31) | ((uintptr_t)z & 7) | ((uintptr_t)weights & 7) |
                                                      (noise ? (uintptr_t)noise & 7 : 0));
    if (pairs) {
        const int c2 = (s + 63) / 64;
#define FWD2(C_) composite_fwd2_kernel<C_><<<grid, block, 0, as_stream(stream)>>>((const float4*)raw, z, rays_d, rays_d_stride, bc_rgb, \
                                                                                 noise, n, s, white_bkgd, rgb, disp, acc, depth, weights, rgb_fg)
        if (c2 <= 1) FWD2(1); else if (c2 == 2) FWD2(2); else if (c2 == 3) FWD2(3); else if (c2 == 4) FWD2(4); else FWD2(8);
#undef FWD2
        return check_launch("inerf_composite_fwd");
    }
    int cc = pick_chunks(s);
    DISPATCH_C(cc, (composite_fwd_kernel<C><<<grid, block, 0, as_stream(stream)>>>(
                       (const float4*)raw, z, rays_d, rays_d_stride, bc_rgb, noise, n, s, white_bkgd, rgb, disp, acc,
                       depth, weights, rgb_fg)));
    return check_launch("inerf_composite_fwd");
}

extern "C" int inerf_composite_bwd(const float* raw, const float* z, const float* rays_d, int rays_d_stride,
                                   const float* bc_rgb, const float* noise, int n, int s, int white_bkgd,
                                   const float* g_rgb, const float* g_disp, const float* g_acc, const float* g_depth,
                                   const float* g_weights, const float* g_rgb_fg, float* d_raw, void* stream) {
    if (n < 0 || s <= 0 || s > 1024 || rays_d_stride < 3) return fail(INERF_E_SHAPE, "inerf_composite_bwd: bad n/s/stride (1 <= s <= 1024)");
    if (n == 0) return INERF_OK;
    if (!raw || !z || !rays_d || !bc_rgb || !d_raw) return fail(INERF_E_ARG, "inerf_composite_bwd: NULL pointer");
    if (((uintptr_t)raw | (uintptr_t)d_raw) & 15) return fail(INERF_E_ALIGN, "inerf_composite_bwd: raw/d_raw must be 16-byte aligned");
    int cc = pick_chunks(s);
    dim3 grid((n + 7) / 8), block(256);
    DISPATCH_C(cc, (composite_bwd_kernel<C><<<grid, block, 0, as_stream(stream)>>>(
                       (const float4*)raw, z, rays_d, rays_d_stride, bc_rgb, noise, n, s, white_bkgd, g_rgb, g_disp,
                       g_acc, g_depth, g_weights, g_rgb_fg, (float4*)d_raw)));
    return check_launch("inerf_composite_bwd");
}

// ---------------------------------------------------------------------------------------------
// training loss seed: img_loss + img_loss0 of audio_exp_nerf.py:540-546 and its gradient in one pass
// ---------------------------------------------------------------------------------------------
// loss[0] = mean((rgb - t)^2), loss[1] = mean((rgb0 - t)^2)  (F.mse_loss over all 3n elements);  g = 2 (x - t) / (3n) for both maps.
__global__ void __launch_bounds__(256) mse_pair_kernel(const float* __restrict__ rgb, const float* __restrict__ rgb0,
                                                       const float* __restrict__ tgt, long long m, float inv_m, float* __restrict__ g_rgb,
                                                       float* __restrict__ g_rgb0, float* __restrict__ loss) {
    float s0 = 0.f, s1 = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        const float t = tgt[i], d0 = rgb[i] - t, d1 = rgb0[i] - t;
        s0 = fmaf(d0, d0, s0);
        s1 = fmaf(d1, d1, s1);
        g_rgb[i] = 2.0f * d0 * inv_m;
        g_rgb0[i] = 2.0f * d1 * inv_m;
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    __shared__ float sh[2][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sh[0][w] = s0; sh[1][w] = s1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
        for (int j = 0; j < 8; ++j) t += sh[threadIdx.x][j];
        atomicAdd(loss + threadIdx.x, t * inv_m);
    }
}

extern "C" int inerf_mse_pair(const float* rgb, const float* rgb0, const float* target, int64_t n_elems, float* g_rgb, float* g_rgb0,
                              float* loss2, void* stream) {
    if (n_elems <= 0) return fail(INERF_E_SHAPE, "inerf_mse_pair: need at least one element");
    if (!rgb || !rgb0 || !target || !g_rgb || !g_rgb0 || !loss2) return fail(INERF_E_ARG, "inerf_mse_pair: NULL pointer");
    cudaError_t e = cudaMemsetAsync(loss2, 0, 2 * sizeof(float), as_stream(stream));
    if (e != cudaSuccess) { set_error("inerf_mse_pair: %s", cudaGetErrorString(e)); return (int)e; }
    const int grid = (int)((n_elems + 255) / 256 < 4 * (long long)num_sms() ? (n_elems + 255) / 256 : 4 * (long long)num_sms());
    mse_pair_kernel<<<grid, 256, 0, as_stream(stream)>>>(rgb, rgb0, target, n_elems, 1.0f / (float)n_elems, g_rgb, g_rgb0, loss2);
    return check_launch("inerf_mse_pair");
}
