// FaceNeRF MLP backward, fp32 mode.
//
// The reference's backward is torch.autograd of models/face_nerf.py:40-80 (called from
// NeRFs/HeadNeRF/train/audio_exp_nerf.py:540-552).  Restated analytically on the folded network
// (SURVEY.md Appendix B), with the activations the fp32 forward kernel stored (mlp_fp32.cu, `save`):
//
//   chain kernel   per 64-point tile, back to front:  delta_l = dH_l * [H_l > 0],  dH_{l-1} = delta_l . W_l
//                  (no dX for layer 0 and for the gamma(p) / gamma(v) / conditioning columns: points are data,
//                  z_samples are detached in the reference).  Writes every delta_l point-major to `deltas` and
//                  accumulates the alpha_linear / rgb_linear gradients.
//   dW kernel      dW_l[n][k] = sum_p delta_l[p][n] * X_l[p][k]  as a split-P SGEMM over 128x128 output tiles
//                  (X_l = stored activations / encodings), db_l = sum_p delta_l[p][n]; fp32 atomics into the
//                  zero-initialised gradient tensors (nn.Linear layout).
//   cond kernel    gradients of the folded conditioning columns: dW[:, cond cols] = db' (x) cond (rank 1) and
//                  d_cond = W[:, cond cols]^T db'  -> d_aud, d_expr (/3, face_nerf.py:49), d_latent.
#include "mlp_common.cuh"

using namespace inerf;

namespace {

constexpr int TM = 64, KC = 16, NTHREADS = 256;

__device__ __forceinline__ int act_idx(int k, int m) { return k * TM + ((((m >> 2) ^ (k >> 2)) & 15) << 2) + (m & 3); }

struct BwdArgs {
    const float* w[INERF_N_PARAMS];
    float* g[INERF_N_PARAMS];       // gradients, nn.Linear layout, accumulated
    const float* acts;              // [P64][SAVE_W]
    float* deltas;                  // [P64][DELTA_W]
    const float* d_raw;             // [P][4]
    long long P;
    int cond_dim, dim_expr;
};

// acc[j][i] = sum_n W[n][wcol + k_j] * in[n][m_i]   (reduction over the layer's OUTPUT features n)
template <int KOUT>
__device__ __forceinline__ void gemm_bwd(const float* __restrict__ W, int ldw, int wcol, int nred,
                                         const float* __restrict__ in, float (&acc)[KOUT / 32][8], float* __restrict__ wsm) {
    constexpr int KT = KOUT / 32, PER = KOUT / 16;
    const int tid = threadIdx.x, tn = tid & 31, tm = tid >> 5;
    const int kcol = tid % KOUT, r0 = (tid / KOUT) * PER;     // loader: PER rows of the 16-row slice, one column
#pragma unroll
    for (int j = 0; j < KT; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
    const int nchunks = nred / KC;
    float wr[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) wr[i] = __ldg(W + (size_t)(r0 + i) * ldw + wcol + kcol);
#pragma unroll
    for (int i = 0; i < PER; ++i) wsm[(r0 + i) * KOUT + kcol] = wr[i];
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        const bool more = c + 1 < nchunks;
        if (more) {
#pragma unroll
            for (int i = 0; i < PER; ++i) wr[i] = __ldg(W + (size_t)((c + 1) * KC + r0 + i) * ldw + wcol + kcol);
        }
        const float* ws = wsm + (c & 1) * (KC * KOUT);
#pragma unroll 4
        for (int nn = 0; nn < KC; ++nn) {
            const int n = c * KC + nn;
            const int sw = (n >> 2) & 15;
            const float4 a0 = *reinterpret_cast<const float4*>(in + n * TM + (((2 * tm) ^ sw) << 2));
            const float4 a1 = *reinterpret_cast<const float4*>(in + n * TM + (((2 * tm + 1) ^ sw) << 2));
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[KT];
            const float4 b0 = *reinterpret_cast<const float4*>(ws + nn * KOUT + 4 * tn);
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            if constexpr (KT == 8) {
                const float4 b1 = *reinterpret_cast<const float4*>(ws + nn * KOUT + 128 + 4 * tn);
                b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
            }
#pragma unroll
            for (int j = 0; j < KT; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(b[j], a[i], acc[j][i]);
        }
        if (more) {
            float* wn = wsm + ((c + 1) & 1) * (KC * KOUT);
#pragma unroll
            for (int i = 0; i < PER; ++i) wn[(r0 + i) * KOUT + kcol] = wr[i];
        }
        __syncthreads();
    }
}

// delta = dH * [H > 0]; write it k-major to smem (next GEMM's operand) and point-major to the global delta store.
template <int KOUT>
__device__ __forceinline__ void mask_store(float (&acc)[KOUT / 32][8], const float* __restrict__ acts_tile, int hcol,
                                           float* __restrict__ deltas_tile, float* __restrict__ out) {
    constexpr int KT = KOUT / 32;
    const int tid = threadIdx.x, tn = tid & 31, tm = tid >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = 8 * tm + i;
#pragma unroll
        for (int jg = 0; jg < KT / 4; ++jg) {
            const int k0 = 4 * tn + 128 * jg;
            const float4 h = *reinterpret_cast<const float4*>(acts_tile + (size_t)m * SAVE_W + hcol + k0);
            float4 d;
            d.x = h.x > 0.f ? acc[4 * jg][i] : 0.f;
            d.y = h.y > 0.f ? acc[4 * jg + 1][i] : 0.f;
            d.z = h.z > 0.f ? acc[4 * jg + 2][i] : 0.f;
            d.w = h.w > 0.f ? acc[4 * jg + 3][i] : 0.f;
            acc[4 * jg][i] = d.x; acc[4 * jg + 1][i] = d.y; acc[4 * jg + 2][i] = d.z; acc[4 * jg + 3][i] = d.w;
            *reinterpret_cast<float4*>(deltas_tile + (size_t)m * DELTA_W + hcol + k0) = d;
        }
    }
#pragma unroll
    for (int j = 0; j < KT; ++j) {
        const int k = 4 * tn + (j & 3) + 128 * (j >> 2);
        const int sw = (k >> 2) & 15;
        *reinterpret_cast<float4*>(out + k * TM + (((2 * tm) ^ sw) << 2)) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
        *reinterpret_cast<float4*>(out + k * TM + (((2 * tm + 1) ^ sw) << 2)) = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NTHREADS, 1) mlp_bwd_chain_kernel(BwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* DA = sm;                       // [256][64]
    float* DB = DA + 256 * TM;            // [256][64]
    float* WS = DB + 256 * TM;            // 2 x [16][256]
    float* DR = WS + 2 * KC * 256;        // [64][4] d_raw of the tile
    const int tid = threadIdx.x, tn = tid & 31, tm = tid >> 5;
    const long long ntiles = (a.P + TM - 1) / TM;
    const int C = a.cond_dim;

    // per-thread partial gradients of the two tiny heads, flushed once at the end
    float g_alpha = 0.f;                  // d alpha_linear.weight[tid]
    float g_rgb[3] = {0.f, 0.f, 0.f};     // d rgb_linear.weight[c][tid % 128] over this thread's half of the points
    float g_hb[4] = {0.f, 0.f, 0.f, 0.f}; // thread 0: d rgb bias (3), d alpha bias

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const float* at = a.acts + (size_t)tile * TM * SAVE_W;
        float* dt = a.deltas + (size_t)tile * TM * DELTA_W;
        if (tid < TM) {
            const long long p = tile * TM + tid;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < a.P) v = reinterpret_cast<const float4*>(a.d_raw)[p];
            *reinterpret_cast<float4*>(DR + tid * 4) = v;
        }
        __syncthreads();
        // ---- rgb_linear: delta_V2 = (W_rgb^T d_rgb) * [v2 > 0] ----------------------------------------------
        {
            const int n = tid & 127, mh = tid >> 7;
            const float w0 = __ldg(a.w[P_RGB_W] + n), w1 = __ldg(a.w[P_RGB_W] + 128 + n), w2 = __ldg(a.w[P_RGB_W] + 256 + n);
            for (int m = mh * 32; m < mh * 32 + 32; ++m) {
                const float4 dr = *reinterpret_cast<const float4*>(DR + m * 4);
                const float v2 = at[(size_t)m * SAVE_W + 2048 + 256 + n];
                const float d = v2 > 0.f ? fmaf(w2, dr.z, fmaf(w1, dr.y, w0 * dr.x)) : 0.f;
                DA[act_idx(n, m)] = d;
                dt[(size_t)m * DELTA_W + 2048 + 256 + n] = d;
                g_rgb[0] = fmaf(dr.x, v2, g_rgb[0]); g_rgb[1] = fmaf(dr.y, v2, g_rgb[1]); g_rgb[2] = fmaf(dr.z, v2, g_rgb[2]);
            }
            // alpha_linear.weight[tid] += sum_m d_sigma[m] * h7[tid][m]
            for (int m = 0; m < TM; ++m) g_alpha = fmaf(DR[m * 4 + 3], at[(size_t)m * SAVE_W + 7 * 256 + tid], g_alpha);
            if (tid == 0)
                for (int m = 0; m < TM; ++m) {
                    g_hb[0] += DR[m * 4]; g_hb[1] += DR[m * 4 + 1]; g_hb[2] += DR[m * 4 + 2]; g_hb[3] += DR[m * 4 + 3];
                }
        }
        __syncthreads();
        // ---- view branch -------------------------------------------------------------------------------------
        {
            float acc[4][8];
            gemm_bwd<128>(a.w[P_VIEWS_W + 4], 128, 0, 128, DA, acc, WS);               // through views_linears.2
            mask_store<128>(acc, at, 2048 + 128, dt, DB);                             // delta_V1
            gemm_bwd<128>(a.w[P_VIEWS_W + 2], 128, 0, 128, DB, acc, WS);               // through views_linears.1
            mask_store<128>(acc, at, 2048, dt, DA);                                   // delta_V0
        }
        float* cur = DB;
        {
            float acc[8][8];
            gemm_bwd<256>(a.w[P_VIEWS_W], 283 + a.dim_expr, 0, 128, DA, acc, WS);      // views_linears.0[:, :256]
#pragma unroll
            for (int j = 0; j < 8; ++j) {                                             // + alpha_linear^T d_sigma
                const float wa = __ldg(a.w[P_ALPHA_W] + 4 * tn + (j & 3) + 128 * (j >> 2));
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(wa, DR[(8 * tm + i) * 4 + 3], acc[j][i]);
            }
            mask_store<256>(acc, at, 7 * 256, dt, DB);                                // delta_7
            // ---- trunk: layers 7 .. 1 ----------------------------------------------------------------------
            float* nxt = DA;
            for (int l = 7; l >= 1; --l) {
                const int ldw = (l == 5) ? 319 + C : 256, wcol = (l == 5) ? 63 + C : 0;
                gemm_bwd<256>(a.w[2 * l], ldw, wcol, 256, cur, acc, WS);
                mask_store<256>(acc, at, (l - 1) * 256, dt, nxt);                      // delta_{l-1}
                float* t = cur; cur = nxt; nxt = t;
            }
        }
        __syncthreads();
    }
    atomicAdd(a.g[P_ALPHA_W] + tid, g_alpha);
    {
        const int n = tid & 127;
        atomicAdd(a.g[P_RGB_W] + n, g_rgb[0]); atomicAdd(a.g[P_RGB_W] + 128 + n, g_rgb[1]); atomicAdd(a.g[P_RGB_W] + 256 + n, g_rgb[2]);
    }
    if (tid == 0) {
        atomicAdd(a.g[P_RGB_B], g_hb[0]); atomicAdd(a.g[P_RGB_B] + 1, g_hb[1]); atomicAdd(a.g[P_RGB_B] + 2, g_hb[2]);
        atomicAdd(a.g[P_ALPHA_B], g_hb[3]);
    }
}

constexpr size_t CHAIN_SMEM = (size_t)(2 * 256 * TM + 2 * KC * 256 + TM * 4) * sizeof(float);

// ---------------------------------------------------------------------------------------------
// dW: split-P SGEMM  dW[n][k] += sum_p delta[p][dcol+n] * acts[p][xcol+k]
// ---------------------------------------------------------------------------------------------
struct DwTile { short dcol, xcol, kvalid, w_index, ldw, wcol, bias_index, row0; };   // one 128x128 output tile
constexpr int MAX_DW_TILES = 48;
struct DwArgs {
    const float* acts; const float* deltas;
    float* g[INERF_N_PARAMS];
    long long ntiles;              // 64-point tiles
    int n_out_tiles, split;
    DwTile t[MAX_DW_TILES];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

static_assert(sizeof(DwArgs) <= 4000, "kernel parameter space");
__global__ void __launch_bounds__(NTHREADS, 1) mlp_bwd_dw_kernel(const __grid_constant__ DwArgs a) {      // table by value: graph-capturable
    extern __shared__ __align__(16) float sm[];
    float* As = sm;                        // 2 x [64 m][128 n]
    float* Bs = As + 2 * TM * 128;         // 2 x [64 m][128 k]
    const DwTile t = a.t[blockIdx.x];
    const int tid = threadIdx.x, tn = tid & 15, tk = tid >> 4;
    const long long per = (a.ntiles + a.split - 1) / a.split;
    const long long t0 = (long long)blockIdx.y * per, t1 = min(a.ntiles, t0 + per);
    float acc[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
    float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int kq = min(128, (t.kvalid + 3) & ~3);         // columns actually stored (multiple of 4 floats)

    auto load = [&](long long tile, int buf) {
        const float* drow = a.deltas + (size_t)tile * TM * DELTA_W + t.dcol;
        const float* xrow = a.acts + (size_t)tile * TM * SAVE_W + t.xcol;
        for (int i = tid; i < TM * 32; i += NTHREADS) {    // 64 rows x 32 float4
            const int m = i >> 5, q = i & 31;
            cp_async16(As + buf * TM * 128 + m * 128 + q * 4, drow + (size_t)m * DELTA_W + q * 4);
            if (q * 4 < kq) cp_async16(Bs + buf * TM * 128 + m * 128 + q * 4, xrow + (size_t)m * SAVE_W + q * 4);
        }
        cp_async_commit();
    };
    for (int i = tid; i < 2 * TM * 128; i += NTHREADS) Bs[i] = 0.f;      // columns >= kvalid stay zero
    __syncthreads();
    if (t0 < t1) load(t0, 0);
    for (long long tile = t0; tile < t1; ++tile) {
        const int buf = (int)((tile - t0) & 1);
        if (tile + 1 < t1) { load(tile + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const float* A = As + buf * TM * 128;
        const float* B = Bs + buf * TM * 128;
#pragma unroll 4
        for (int m = 0; m < TM; ++m) {
            const float4 a0 = *reinterpret_cast<const float4*>(A + m * 128 + 4 * tn);
            const float4 a1 = *reinterpret_cast<const float4*>(A + m * 128 + 64 + 4 * tn);
            const float4 b0 = *reinterpret_cast<const float4*>(B + m * 128 + 4 * tk);
            const float4 b1 = *reinterpret_cast<const float4*>(B + m * 128 + 64 + 4 * tk);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                bsum[j] += av[j];
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(av[j], bv[i], acc[j][i]);
            }
        }
        __syncthreads();
    }
    float* G = a.g[t.w_index];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = 4 * tn + (j & 3) + 64 * (j >> 2);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = 4 * tk + (i & 3) + 64 * (i >> 2);
            if (k < t.kvalid) atomicAdd(G + (size_t)(t.row0 + n) * t.ldw + t.wcol + k, acc[j][i]);
        }
        if (t.bias_index >= 0 && tk == 0) atomicAdd(a.g[t.bias_index] + t.row0 + n, bsum[j]);
    }
}

constexpr size_t DW_SMEM = (size_t)(4 * TM * 128) * sizeof(float);

// ---------------------------------------------------------------------------------------------
// folded conditioning columns
// ---------------------------------------------------------------------------------------------
struct CondBwdArgs {
    const float* w[3]; float* gw[3]; const float* gb[3];   // pts_linears.0, pts_linears.5, views_linears.0
    const float* aud; const float* expr; const float* latent;
    int da, de, dl;
    float* d_cond;                                         // [da+de+dl], zero-initialised
};

__global__ void mlp_bwd_cond_kernel(CondBwdArgs c) {
    __shared__ float cv[1024];
    const int C = c.da + c.de + c.dl;
    for (int i = threadIdx.x; i < C; i += blockDim.x)
        cv[i] = i < c.da ? c.aud[i] : (i < c.da + c.de ? __fdiv_rn(c.expr[i - c.da], 3.0f) : c.latent[i - c.da - c.de]);
    __syncthreads();
    const int job = blockIdx.x;                            // 0: W0, 1: W5, 2: WV0
    const int N = job == 2 ? 128 : 256;
    const int ldw = job == 0 ? 63 + C : (job == 1 ? 319 + C : 283 + c.de);
    const int col0 = job == 2 ? 283 : 63;
    const int j0 = job == 2 ? c.da : 0, nj = job == 2 ? c.de : C;   // slice of the conditioning vector this layer sees
    for (int j = threadIdx.x; j < nj; j += blockDim.x) {
        float dc = 0.f;
        const float cj = cv[j0 + j];
        for (int n = blockIdx.y * (N / gridDim.y); n < (int)(blockIdx.y + 1) * (N / (int)gridDim.y); ++n) {      // gridDim.y slices of the rows
            const float db = c.gb[job][n];
            dc = fmaf(c.w[job][(size_t)n * ldw + col0 + j], db, dc);
            c.gw[job][(size_t)n * ldw + col0 + j] = db * cj;     // rank-1: these columns see the same input at every point
        }
        const int jj = j0 + j;
        if (jj >= c.da && jj < c.da + c.de) dc = __fdiv_rn(dc, 3.0f);   // expr enters as expr/3
        atomicAdd(c.d_cond + jj, dc);
    }
}

}  // namespace

namespace inerf {

int mlp_fp32_bwd_launch(const InerfNetDims* dims, const float* const* params_host, float* const* grads_host,
                        const float* aud, const float* expr, const float* latent, const float* acts, float* deltas,
                        const float* d_raw, long long P, float* d_cond, void* dw_args_dev, cudaStream_t st) {
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaError_t e1 = cudaFuncSetAttribute(mlp_bwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHAIN_SMEM);
        cudaError_t e2 = cudaFuncSetAttribute(mlp_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_SMEM);
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            set_error("mlp_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
            return (int)(e1 != cudaSuccess ? e1 : e2);
        }
        configured_dev = dev;
    }
    const int C = dims->dim_aud + dims->dim_expr + dims->dim_latent, E = dims->dim_expr;
    const long long ntiles = (P + TM - 1) / TM;

    BwdArgs b{};
    for (int i = 0; i < INERF_N_PARAMS; ++i) { b.w[i] = params_host[i]; b.g[i] = grads_host[i]; }
    b.acts = acts; b.deltas = deltas; b.d_raw = d_raw; b.P = P; b.cond_dim = C; b.dim_expr = E;
    const int grid = (int)(ntiles < (long long)num_sms() ? ntiles : (long long)num_sms());
    mlp_bwd_chain_kernel<<<grid, NTHREADS, CHAIN_SMEM, st>>>(b);
    int rc = check_launch("inerf_mlp_bwd[chain]");
    if (rc) return rc;

    DwArgs d{};
    d.acts = acts; d.deltas = deltas; d.ntiles = ntiles;
    for (int i = 0; i < INERF_N_PARAMS; ++i) d.g[i] = grads_host[i];
    int nt = 0;
    auto add = [&](int dcol, int N, int xcol, int K, int w_index, int ldw, int wcol, int bias_index) {
        for (int n0 = 0; n0 < N; n0 += 128)
            for (int k0 = 0; k0 < K; k0 += 128) {
                DwTile& t = d.t[nt++];
                t.dcol = (short)(dcol + n0); t.xcol = (short)(xcol + k0); t.kvalid = (short)((K - k0) < 128 ? (K - k0) : 128);
                t.w_index = (short)w_index; t.ldw = (short)ldw; t.wcol = (short)(wcol + k0);
                t.bias_index = (short)((k0 == 0) ? bias_index : -1);
                t.row0 = (short)n0;
            }
    };
    add(0, 256, SAVE_PE, 63, 0, 63 + C, 0, 1);                                             // pts_linears.0 <- gamma(p)
    for (int l = 1; l < 8; ++l) {
        if (l == 5) {
            add(5 * 256, 256, SAVE_PE, 63, 10, 319 + C, 0, 11);                            // skip layer: gamma(p) columns
            add(5 * 256, 256, 4 * 256, 256, 10, 319 + C, 63 + C, -1);                      //             h4 columns
        } else {
            add(l * 256, 256, (l - 1) * 256, 256, 2 * l, 256, 0, 2 * l + 1);
        }
    }
    add(2048, 128, 7 * 256, 256, P_VIEWS_W, 283 + E, 0, P_VIEWS_W + 1);                    // views_linears.0 <- h7
    add(2048, 128, SAVE_DIR, 27, P_VIEWS_W, 283 + E, 256, -1);                             //                 <- gamma(v)
    add(2048 + 128, 128, 2048, 128, P_VIEWS_W + 2, 128, 0, P_VIEWS_W + 3);
    add(2048 + 256, 128, 2048 + 128, 128, P_VIEWS_W + 4, 128, 0, P_VIEWS_W + 5);
    d.n_out_tiles = nt;
    d.split = (int)((4 * (long long)num_sms() / nt) < 1 ? 1 : (4 * (long long)num_sms() / nt));
    if ((long long)d.split > ntiles) d.split = (int)ntiles;
    (void)dw_args_dev;
    mlp_bwd_dw_kernel<<<dim3(nt, d.split), NTHREADS, DW_SMEM, st>>>(d);
    rc = check_launch("inerf_mlp_bwd[dW]");
    if (rc) return rc;

    if (C > 0) rc = mlp_bwd_cond_launch(dims, params_host, grads_host, aud, expr, latent, d_cond, st);
    return rc;
}

// Gradients of the folded conditioning columns from the bias gradients of pts_linears.0 / .5 and views_linears.0 (both MLP modes).
int mlp_bwd_cond_launch(const InerfNetDims* dims, const float* const* params_host, float* const* grads_host, const float* aud,
                        const float* expr, const float* latent, float* d_cond, cudaStream_t st) {
    const int E = dims->dim_expr;
    CondBwdArgs c{};
    c.w[0] = params_host[0]; c.w[1] = params_host[10]; c.w[2] = params_host[P_VIEWS_W];
    c.gw[0] = grads_host[0]; c.gw[1] = grads_host[10]; c.gw[2] = grads_host[P_VIEWS_W];
    c.gb[0] = grads_host[1]; c.gb[1] = grads_host[11]; c.gb[2] = grads_host[P_VIEWS_W + 1];
    c.aud = aud; c.expr = expr; c.latent = latent;
    c.da = dims->dim_aud; c.de = E; c.dl = dims->dim_latent; c.d_cond = d_cond;
    mlp_bwd_cond_kernel<<<dim3(E > 0 ? 3 : 2, 16), 256, 0, st>>>(c);
    return check_launch("inerf_mlp_bwd[cond]");
}

size_t mlp_fp32_bwd_args_bytes() { return sizeof(DwArgs); }

}  // namespace inerf
