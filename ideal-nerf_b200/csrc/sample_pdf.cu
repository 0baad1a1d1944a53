// Hierarchical importance sampling: sample_pdf + merge/sort + z_std.
//
// Reference: NeRFs/HeadNeRF/helper.py:269-313 (sample_pdf) and the tail of render_rays,
// NeRFs/HeadNeRF/train/audio_exp_nerf.py:342-347,364.
//
// One warp per ray.  The per-ray CDF lives in shared memory; every lane then inverts it for its
// share of the n_imp samples with a 6-step binary search (torch.searchsorted(right=True)).
//
// Bit-exactness (north_star gate: sample indices equal the reference's in fp32/det mode).  The
// index depends on the last bit of the CDF, so policy INERF_PDF_EXACT_TORCH_CPU reproduces the
// reference arithmetic operation by operation:
//   * weights + 1e-5            fp32 add
//   * torch.sum(weights, -1)    ATen's CPU inner-dim sum order: 8-lane vectors, four vector
//                               accumulators, left-over vectors into accumulator 0, accumulators
//                               folded 1,2,3 -> 0, scalar tail summed from 0.0, then the 8 lanes
//                               added in lane order (oracle/render_oracle.py::torch_cpu_rowsum_f32)
//   * pdf = w / S               IEEE fp32 division
//   * torch.cumsum              fp64 running sum rounded to fp32 per element.  Every pdf_j is a
//                               multiple of 2^-46 for weights in [0,1] and the sum stays below 2, so
//                               all partial sums are exact in fp64 and a parallel fp64 scan returns
//                               the same bits as the sequential one.
//   * the lerp                  fp32 sub/div/mul/add, no FMA contraction.
// u (torch.linspace(0,1,n_imp) or the random draws) is always an input table, never recomputed.
#include <math_constants.h>

#include "common.cuh"
#include "philox.cuh"
#include "sample_parts.cuh"

using namespace inerf;

namespace {

struct PdfArgs {
    const float* bins; int bins_stride; int fused_bins;   // fused_bins: `bins` is z_coarse, mids are formed here
    const float* weights; int w_stride;
    int n, nb, n_imp;
    const float* u; int u_per_ray;
    int policy;
    float* z_samples; long long* inds;
    const float* z_coarse; int s1;
    float* z_merged; float* z_std;
    int sort_n;        // shared floats of the merge (0 when no merge): s1 + imp_p2 + s1 + n_imp
    int imp_p2;        // power of two >= n_imp
    int warp_floats;   // shared floats per warp
};

__device__ __forceinline__ float exact_rowsum(const float* w, int nw, int lane) {
    // ATen CPU sum order, see header comment
    const int nv = nw >> 3, full = nv >> 2;
    const int k = lane >> 3, l = lane & 7;
    float acc = 0.0f;
    for (int it = 0; it < full; ++it) acc = __fadd_rn(acc, w[((it * 4 + k) << 3) + l]);
    if (k == 0)
        for (int v = full * 4; v < nv; ++v) acc = __fadd_rn(acc, w[(v << 3) + l]);
    float a1 = __shfl_sync(0xffffffffu, acc, l + 8);
    float a2 = __shfl_sync(0xffffffffu, acc, l + 16);
    float a3 = __shfl_sync(0xffffffffu, acc, l + 24);
    acc = __fadd_rn(__fadd_rn(__fadd_rn(acc, a1), a2), a3);       // valid on lanes 0..7
    float total = 0.0f;
    for (int j = nv << 3; j < nw; ++j) total = __fadd_rn(total, w[j]);
#pragma unroll
    for (int q = 0; q < 8; ++q) total = __fadd_rn(total, __shfl_sync(0xffffffffu, acc, q));
    return __shfl_sync(0xffffffffu, total, 0);
}

__global__ void __launch_bounds__(256) sample_pdf_kernel(PdfArgs a) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int ray = blockIdx.x * (blockDim.x >> 5) + wib;
    if (ray >= a.n) return;
    const int nb = a.nb, nw = nb - 1;
    float* wsm = smem + (size_t)wib * a.warp_floats;     // nw  : weights+1e-5, then pdf
    float* bsm = wsm + ((nw + 3) & ~3);                  // nb  : bins
    float* csm = bsm + ((nb + 3) & ~3);                  // nb  : cdf
    float* ssm = csm + ((nb + 3) & ~3);                  // sort buffer

    const float* wrow = a.weights + (size_t)ray * a.w_stride;
    const float* brow = a.bins + (size_t)ray * a.bins_stride;
    for (int j = lane; j < nw; j += 32) wsm[j] = __fadd_rn(wrow[j], 1e-5f);
    for (int j = lane; j < nb; j += 32)
        bsm[j] = a.fused_bins ? __fmul_rn(0.5f, __fadd_rn(brow[j + 1], brow[j])) : brow[j];
    __syncwarp();

    float S;
    if (a.policy == INERF_PDF_EXACT_TORCH_CPU) {
        S = exact_rowsum(wsm, nw, lane);
    } else {
        float p = 0.f;
        for (int j = lane; j < nw; j += 32) p += wsm[j];
        S = warp_sum(p);
    }
    // pdf + fp64 prefix: lane owns the contiguous run [lane*E, lane*E+E)
    const int E = (nw + 31) >> 5;
    const int j0 = lane * E, j1 = min(nw, j0 + E);
    double local = 0.0;
    for (int j = j0; j < j1; ++j) {
        float p = __fdiv_rn(wsm[j], S);
        wsm[j] = p;
        local += (double)p;
    }
    double incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    double run = incl - local;                 // exclusive prefix of the lane totals (exact, see header)
    for (int j = j0; j < j1; ++j) {
        run += (double)wsm[j];
        csm[j + 1] = (float)run;
    }
    if (lane == 0) csm[0] = 0.0f;
    __syncwarp();

    // invert the CDF
    const float* urow = a.u + (a.u_per_ray ? (size_t)ray * a.n_imp : 0);
    float sum = 0.f;
    for (int k = lane; k < a.n_imp; k += 32) {
        float u = urow[k];
        int lo = 0, hi = nb;                   // first index with cdf > u   (searchsorted right=True)
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (csm[mid] <= u) lo = mid + 1; else hi = mid;
        }
        int below = max(lo - 1, 0), above = min(lo, nb - 1);
        float cb = csm[below], ca = csm[above];
        float den = __fsub_rn(ca, cb);
        if (den < 1e-5f) den = 1.0f;
        float t = __fdiv_rn(__fsub_rn(u, cb), den);
        float bb = bsm[below], ba = bsm[above];
        float zs = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
        a.z_samples[(size_t)ray * a.n_imp + k] = zs;
        if (a.inds) a.inds[(size_t)ray * a.n_imp + k] = lo;
        if (a.sort_n) ssm[a.s1 + k] = zs;
        sum += zs;
    }

    if (a.z_std) {                              // torch.std(z_samples, -1, unbiased=False)
        float mean = warp_sum(sum) / (float)a.n_imp;
        float sq = 0.f;
        for (int k = lane; k < a.n_imp; k += 32) {
            float d = a.z_samples[(size_t)ray * a.n_imp + k] - mean;   // own writes, same thread
            sq += d * d;
        }
        sq = warp_sum(sq);
        if (lane == 0) a.z_std[ray] = sqrtf(sq / (float)a.n_imp);
    }

    if (a.sort_n) {                             // z_vals = sort(cat([z_vals, z_samples]))  (:347)
        // The coarse depths are already sorted; the samples are sorted when u is (the shared linspace table of det=True: the
        // inverse CDF is monotone), otherwise a bitonic sort of the n_imp samples alone.  Then a rank merge: every element's
        // output slot is its own index plus the number of elements of the OTHER list that precede it (strict for the coarse
        // list, non-strict for the samples, so ties get distinct slots) -- 7-step binary searches instead of sorting s1+n_imp.
        const int tot = a.s1 + a.n_imp, P2 = a.imp_p2;
        float* zsm = ssm + a.s1;                // [P2] samples (padded with +inf)
        float* osm = zsm + P2;                  // [tot] merged row
        const float* zc = a.z_coarse + (size_t)ray * a.s1;
        for (int j = lane; j < a.s1; j += 32) ssm[j] = zc[j];
        for (int j = a.n_imp + lane; j < P2; j += 32) zsm[j] = CUDART_INF_F;
        __syncwarp();
        if (a.u_per_ray) {
            for (int k = 2; k <= P2; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = lane; t < (P2 >> 1); t += 32) {
                        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                        int p = i | j;
                        float x = zsm[i], y = zsm[p];
                        bool up = (i & k) == 0;
                        if ((x > y) == up) { zsm[i] = y; zsm[p] = x; }
                    }
                    __syncwarp();
                }
            }
        }
        for (int j = lane; j < a.s1; j += 32) {             // coarse j: samples strictly below it come first
            const float v = ssm[j];
            int lo = 0, hi = a.n_imp;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (zsm[mid] < v) lo = mid + 1; else hi = mid; }
            osm[j + lo] = v;
        }
        for (int i = lane; i < a.n_imp; i += 32) {          // sample i: coarse depths <= it come first
            const float v = zsm[i];
            int lo = 0, hi = a.s1;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (ssm[mid] <= v) lo = mid + 1; else hi = mid; }
            osm[i + lo] = v;
        }
        __syncwarp();
        float* out = a.z_merged + (size_t)ray * tot;
        for (int j = lane; j < tot; j += 32) out[j] = osm[j];
    }
}


// ---------------------------------------------------------------------------------------------
// Stochastic branch (perturb > 0, helper.py:282-283: u = torch.rand(N, n_imp)) with the draws made IN the kernel.
//
// Parity for this branch is defined on supplied draws (the kernel above); with the reference's own generator out of reach the
// only requirements are u ~ iid U[0,1) per ray and the outputs the renderer consumes: sort(cat(z, z_samples)) and std(z_samples),
// both symmetric in the order of the draws.  So the kernel draws the ORDER STATISTICS of n_imp uniforms directly --
// U_(k) = (E_1 + ... + E_k) / (E_1 + ... + E_{n_imp+1}) with E_i iid Exp(1) (Renyi's representation) -- one warp scan instead of a
// 128-key bitonic sort, no (N, n_imp) tensor of draws in HBM (512 B/ray written by torch.rand and read back), and a plain fp32
// warp-shuffle CDF (policy INERF_PDF_FAST: bit-exactness against the CPU is moot when u is random).  z_samples come out ascending.
// ---------------------------------------------------------------------------------------------
struct PdfRngArgs {
    const float* z_coarse; const float* w_coarse; int n, s1, n_imp;
    const unsigned long long* rng_state; unsigned stream_id;
    float* z_samples;            // may be NULL
    float* z_merged; float* z_std;
    int warp_floats, runs;       // runs = Philox blocks per lane = ceil(ceil(n_imp / 32) / 4)
};

__global__ void __launch_bounds__(256) importance_rng_kernel(PdfRngArgs a) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int ray = blockIdx.x * (blockDim.x >> 5) + wib;
    if (ray >= a.n) return;
    const int s1 = a.s1, nb = s1 - 1, nw = s1 - 2, n_imp = a.n_imp, tot = s1 + n_imp;
    float* zc = smem + (size_t)wib * a.warp_floats;      // [s1]   coarse depths
    float* bsm = zc + ((s1 + 3) & ~3);                   // [nb]   bin mid-points
    float* csm = bsm + ((nb + 3) & ~3);                  // [nb]   cdf
    float* zsm = csm + ((nb + 3) & ~3);                  // [n_imp] uniforms, then samples
    float* osm = zsm + ((n_imp + 3) & ~3);               // [tot]  merged row

    const float* zrow = a.z_coarse + (size_t)ray * s1;
    const float* wrow = a.w_coarse + (size_t)ray * s1 + 1;       // weights[..., 1:-1]
    for (int j = lane; j < s1; j += 32) zc[j] = zrow[j];
    __syncwarp();
    for (int j = lane; j < nb; j += 32) bsm[j] = 0.5f * (zc[j + 1] + zc[j]);

    // cdf: lane owns the contiguous run [lane*E, lane*E+E) of the nw weights
    const int E = (nw + 31) >> 5;
    const int j0 = min(nw, lane * E), j1 = min(nw, j0 + E);
    float local = 0.f;
    for (int j = j0; j < j1; ++j) local += wrow[j] + 1e-5f;
    float incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const float inv_S = 1.0f / __shfl_sync(0xffffffffu, incl, 31);
    float run = incl - local;
    for (int j = j0; j < j1; ++j) {
        run += wrow[j] + 1e-5f;
        csm[j + 1] = run * inv_S;
    }
    if (lane == 0) csm[0] = 0.0f;

    // sorted uniforms: lane owns draws [k0, k1), k0 = lane * R
    const int R = (n_imp + 31) >> 5;
    const int k0 = min(n_imp, lane * R), k1 = min(n_imp, k0 + R);
    const uint64_t base = (uint64_t)ray * (uint64_t)(32 * a.runs + 1);
    float acc = 0.f;
    for (int b = 0; b < a.runs; ++b) {
        const Philox4 q = philox_at(a.rng_state, base + (uint64_t)(lane * a.runs + b), a.stream_id);
        const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = k0 + 4 * b + t;
            if (k < k1) { acc += -__logf(u01_open0(w4[t])); zsm[k] = acc; }
        }
    }
    float eincl = acc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, eincl, o);
        if (lane >= o) eincl += t;
    }
    const float e_tail = -__logf(u01_open0(philox_at(a.rng_state, base + (uint64_t)(32 * a.runs), a.stream_id).x));
    const float inv_T = 1.0f / (__shfl_sync(0xffffffffu, eincl, 31) + e_tail);
    const float e_off = eincl - acc;
    __syncwarp();

    // invert the CDF (searchsorted right=True) for the lane's own draws
    float sum = 0.f;
    for (int k = k0; k < k1; ++k) {
        const float u = fminf((zsm[k] + e_off) * inv_T, 1.0f);
        int lo = 0, hi = nb;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (csm[mid] <= u) lo = mid + 1; else hi = mid;
        }
        const int below = max(lo - 1, 0), above = min(lo, nb - 1);
        const float cb = csm[below], ca = csm[above];
        float den = ca - cb;
        if (den < 1e-5f) den = 1.0f;
        const float t = (u - cb) / den;
        const float bb = bsm[below], ba = bsm[above];
        const float zs = bb + t * (ba - bb);
        zsm[k] = zs;
        sum += zs;
    }
    __syncwarp();
    if (a.z_std) {
        const float mean = warp_sum(sum) / (float)n_imp;
        float sq = 0.f;
        for (int k = k0; k < k1; ++k) { const float d = zsm[k] - mean; sq += d * d; }
        sq = warp_sum(sq);
        if (lane == 0) a.z_std[ray] = sqrtf(sq / (float)n_imp);
    }
    if (a.z_samples) {
        float* zo = a.z_samples + (size_t)ray * n_imp;
        for (int k = lane; k < n_imp; k += 32) zo[k] = zsm[k];
    }
    // rank merge of the two ascending lists (ties: coarse depths first)
    for (int j = lane; j < s1; j += 32) {
        const float v = zc[j];
        int lo = 0, hi = n_imp;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (zsm[mid] < v) lo = mid + 1; else hi = mid; }
        osm[j + lo] = v;
    }
    for (int i = lane; i < n_imp; i += 32) {
        const float v = zsm[i];
        int lo = 0, hi = s1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (zc[mid] <= v) lo = mid + 1; else hi = mid; }
        osm[i + lo] = v;
    }
    __syncwarp();
    float* out = a.z_merged + (size_t)ray * tot;
    if ((tot & 3) == 0 && (a.warp_floats & 3) == 0) {      // rows are 16-byte aligned: 128-bit stores
        for (int j = lane; j < (tot >> 2); j += 32) reinterpret_cast<float4*>(out)[j] = reinterpret_cast<const float4*>(osm)[j];
    } else {
        for (int j = lane; j < tot; j += 32) out[j] = osm[j];
    }
}

// ---------------------------------------------------------------------------------------------
// The same kernel specialised for the reference's configuration (N_samples = 64, N_importance = 128): every loop is unrolled, each
// lane owns 2 coarse depths / 2 pdf bins / 4 draws in registers, the searches are branch-free fixed-depth bisections (4 instructions
// per step, the 4 draws of a lane interleaved), one Philox block per lane (the extra exponential of the spacing construction comes
// from the 4 x 8 low bits the 24-bit uniforms do not use), 128-bit row loads and stores.  The generic kernel above spends 1670 warp
// instructions per ray (issue bound, ncu: 83 % issue-slot utilisation, 364 us per 202 500 rays); this one ~1/3 of that.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) importance_rng_64_128_kernel(PdfRngArgs a) {
    __shared__ __align__(16) float smem[8 * IMP64_WARP_FLOATS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int ray = blockIdx.x * 8 + wib;
    if (ray >= a.n) return;
    const float2 z2 = reinterpret_cast<const float2*>(a.z_coarse + (size_t)ray * 64)[lane];
    const float2 w2 = reinterpret_cast<const float2*>(a.w_coarse + (size_t)ray * 64)[lane];
    importance_rng_64_128_ray(z2, w2, ray, lane, smem + wib * IMP64_WARP_FLOATS, a.rng_state, a.stream_id, a.z_samples, a.z_merged, a.z_std);
}

int launch_pdf(PdfArgs a, void* stream) {
    if (a.n < 0 || a.nb < 2 || a.nb > 1024 || a.n_imp <= 0 || a.n_imp > 1024)
        return fail(INERF_E_SHAPE, "sample_pdf: need 2 <= n_bins <= 1024, 1 <= n_imp <= 1024");
    if (a.n == 0) return INERF_OK;
    if (!a.bins || !a.weights || !a.u || !a.z_samples) return fail(INERF_E_ARG, "sample_pdf: NULL pointer");
    if (a.policy != INERF_PDF_EXACT_TORCH_CPU && a.policy != INERF_PDF_FAST)
        return fail(INERF_E_ARG, "sample_pdf: unknown policy");
    a.sort_n = 0;
    if (a.z_merged) {
        if (!a.z_coarse || a.s1 <= 0 || a.s1 + a.n_imp > 2048) return fail(INERF_E_SHAPE, "sample_pdf: merge needs z_coarse and s1+n_imp <= 2048");
        int p = 2;
        while (p < a.n_imp) p <<= 1;
        a.imp_p2 = p;
        a.sort_n = a.s1 + p + a.s1 + a.n_imp;
    }
    int nw = a.nb - 1;
    a.warp_floats = ((nw + 3) & ~3) + 2 * ((a.nb + 3) & ~3) + a.sort_n;
    int warps = 8;
    while (warps > 1 && (size_t)warps * a.warp_floats * sizeof(float) > 48 * 1024) warps >>= 1;
    size_t smem = (size_t)warps * a.warp_floats * sizeof(float);
    if (smem > 48 * 1024) return fail(INERF_E_SHAPE, "sample_pdf: per-ray working set exceeds 48 KB of shared memory");
    dim3 grid((a.n + warps - 1) / warps), block(warps * 32);
    sample_pdf_kernel<<<grid, block, smem, as_stream(stream)>>>(a);
    return check_launch("inerf_sample_pdf");
}

}  // namespace

extern "C" int inerf_sample_pdf(const float* bins, int bins_stride, const float* weights, int w_stride, int n,
                                int n_bins, int n_imp, const float* u, int u_per_ray, int policy, float* z_samples,
                                int64_t* inds, const float* z_coarse, int s1, float* z_merged, float* z_std,
                                void* stream) {
    PdfArgs a{};
    a.bins = bins; a.bins_stride = bins_stride; a.fused_bins = 0;
    a.weights = weights; a.w_stride = w_stride;
    a.n = n; a.nb = n_bins; a.n_imp = n_imp;
    a.u = u; a.u_per_ray = u_per_ray; a.policy = policy;
    a.z_samples = z_samples; a.inds = (long long*)inds;
    a.z_coarse = z_coarse; a.s1 = s1; a.z_merged = z_merged; a.z_std = z_std;
    if (bins_stride < n_bins || w_stride < n_bins - 1) return fail(INERF_E_SHAPE, "inerf_sample_pdf: row stride smaller than row");
    return launch_pdf(a, stream);
}

extern "C" int inerf_importance_sample(const float* z_coarse, const float* w_coarse, int n, int s1, int n_imp,
                                       const float* u, int u_per_ray, int policy, float* z_samples, int64_t* inds,
                                       float* z_merged, float* z_std, void* stream) {
    if (s1 < 3) return fail(INERF_E_SHAPE, "inerf_importance_sample: need at least 3 coarse samples");
    if (!w_coarse) return fail(INERF_E_ARG, "inerf_importance_sample: NULL pointer");
    PdfArgs a{};
    a.bins = z_coarse; a.bins_stride = s1; a.fused_bins = 1;     // bins = .5*(z[1:]+z[:-1])   (:342)
    a.weights = w_coarse + 1; a.w_stride = s1;                   // weights[..., 1:-1]          (:344)
    a.n = n; a.nb = s1 - 1; a.n_imp = n_imp;
    a.u = u; a.u_per_ray = u_per_ray; a.policy = policy;
    a.z_samples = z_samples; a.inds = (long long*)inds;
    a.z_coarse = z_coarse; a.s1 = s1; a.z_merged = z_merged; a.z_std = z_std;
    return launch_pdf(a, stream);
}

extern "C" int inerf_importance_sample_rng(const float* z_coarse, const float* w_coarse, int n, int s1, int n_imp,
                                           const uint64_t* rng_state, uint32_t stream_id, float* z_samples, float* z_merged,
                                           float* z_std, void* stream) {
    if (n < 0 || s1 < 3 || s1 > 1024 || n_imp <= 0 || n_imp > 1024) return fail(INERF_E_SHAPE, "inerf_importance_sample_rng: need 3 <= s1 <= 1024, 1 <= n_imp <= 1024");
    if (stream_id > 255) return fail(INERF_E_ARG, "inerf_importance_sample_rng: stream_id must be < 256");
    if (n == 0) return INERF_OK;
    if (!z_coarse || !w_coarse || !rng_state || !z_merged) return fail(INERF_E_ARG, "inerf_importance_sample_rng: NULL pointer");
    PdfRngArgs a{};
    a.z_coarse = z_coarse; a.w_coarse = w_coarse; a.n = n; a.s1 = s1; a.n_imp = n_imp;
    a.rng_state = reinterpret_cast<const unsigned long long*>(rng_state); a.stream_id = stream_id;
    a.z_samples = z_samples; a.z_merged = z_merged; a.z_std = z_std;
    if (s1 == 64 && n_imp == 128 && (((uintptr_t)z_coarse | (uintptr_t)w_coarse | (uintptr_t)z_merged | (uintptr_t)z_samples) & 15) == 0) {
        importance_rng_64_128_kernel<<<(n + 7) / 8, 256, 0, as_stream(stream)>>>(a);      // the reference's configuration
        return check_launch("inerf_importance_sample_rng");
    }
    const int nb = s1 - 1;
    a.warp_floats = ((s1 + 3) & ~3) + 2 * ((nb + 3) & ~3) + ((n_imp + 3) & ~3) + ((s1 + n_imp + 3) & ~3);
    a.runs = (((n_imp + 31) >> 5) + 3) >> 2;
    int warps = 8;
    while (warps > 1 && (size_t)warps * a.warp_floats * sizeof(float) > 48 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * a.warp_floats * sizeof(float);
    if (smem > 48 * 1024) return fail(INERF_E_SHAPE, "inerf_importance_sample_rng: per-ray working set exceeds 48 KB of shared memory");
    importance_rng_kernel<<<(n + warps - 1) / warps, warps * 32, smem, as_stream(stream)>>>(a);
    return check_launch("inerf_importance_sample_rng");
}
