// Per-frame conditioning nets: AudioNet (DeepSpeech window -> audio code) and AudioAttNet (attention over the smoothing window).
//
// Reference: models/audio_net.py:43-69 (AudioNet) and :8-36 (AudioAttNet), called once per frame from
// NeRFs/HeadNeRF/train/audio_exp_nerf.py:241-266.  Tiny (110 k + 3 k MACs per frame) but on the per-frame critical path: on the
// device the audio code feeds inerf_mlp_fold_cond without a host round trip.  The backward kernels are what
// autograd computes for the two modules inside the reference's loss.backward() (audio_exp_nerf.py:263-266 are in the graph, :493 puts
// network.parameters() -- these nets included -- into Adam): the forward is recomputed in shared memory (110 k MACs), then the chain
// rule layer by layer, parameter gradients accumulated with atomicAdd (one CTA per sample of the 8-frame smoothing window).
//
// One CTA per sample; every activation lives in shared memory; weights are read in nn.Conv1d / nn.Linear layout.
#include "common.cuh"

using namespace inerf;

namespace {

__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : 0.02f * x; }

// out[o][t] = b[o] + sum_c sum_k w[o][c][k] * in[c][stride*t + k - 1]   (kernel 3, padding 1), then LeakyReLU(0.02)
__device__ __forceinline__ void conv1d_k3(const float* __restrict__ w, const float* __restrict__ b, const float* in, float* out, int cin,
                                          int cout, int len_in, int stride) {
    const int len_out = (len_in + 2 - 3) / stride + 1;
    for (int i = threadIdx.x; i < cout * len_out; i += blockDim.x) {
        const int o = i / len_out, t = i - o * len_out;
        float acc = b[o];
        const float* wo = w + (size_t)o * cin * 3;
        for (int c = 0; c < cin; ++c)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int p = stride * t + k - 1;
                if (p >= 0 && p < len_in) acc = fmaf(wo[c * 3 + k], in[c * len_in + p], acc);
            }
        out[o * len_out + t] = lrelu(acc);
    }
    __syncthreads();
}

struct AudioNetArgs {
    const float* w[6]; const float* b[6];     // encoder_conv.{0,2,4,6}, encoder_fc1.{0,2}
    const float* x;                            // (n, 16, 29)
    float* y;                                  // (n, dim_aud)
    int dim_aud;
};

__global__ void __launch_bounds__(128) audio_net_kernel(AudioNetArgs a) {
    __shared__ float s0[29 * 16], s1[32 * 8], s2[32 * 4], s3[64 * 2], s4[64], s5[64];
    const float* x = a.x + (size_t)blockIdx.x * 16 * 29;
    // x[:, 8-half_w:8+half_w, :].permute(0, 2, 1): (16 frames, 29 features) -> [29 channels][16]
    for (int i = threadIdx.x; i < 16 * 29; i += blockDim.x) {
        const int t = i / 29, c = i - t * 29;
        s0[c * 16 + t] = x[i];
    }
    __syncthreads();
    conv1d_k3(a.w[0], a.b[0], s0, s1, 29, 32, 16, 2);
    conv1d_k3(a.w[1], a.b[1], s1, s2, 32, 32, 8, 2);
    conv1d_k3(a.w[2], a.b[2], s2, s3, 32, 64, 4, 2);
    conv1d_k3(a.w[3], a.b[3], s3, s4, 64, 64, 2, 2);
    for (int o = threadIdx.x; o < 64; o += blockDim.x) {
        float acc = a.b[4][o];
        for (int c = 0; c < 64; ++c) acc = fmaf(a.w[4][o * 64 + c], s4[c], acc);
        s5[o] = lrelu(acc);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < a.dim_aud; o += blockDim.x) {
        float acc = a.b[5][o];
        for (int c = 0; c < 64; ++c) acc = fmaf(a.w[5][o * 64 + c], s5[c], acc);
        a.y[(size_t)blockIdx.x * a.dim_aud + o] = acc;
    }
}

struct AudioAttArgs {
    const float* w[6]; const float* b[6];     // attentionConvNet.{0,2,4,6,8}, attentionNet.0
    const float* x;                            // (seq_len = 8, dim_feat)
    float* y;                                  // (dim_feat)
    int dim_att, dim_feat;                     // channels the attention looks at (32), width of the codes (64 / 76)
};

__global__ void __launch_bounds__(128) audio_att_kernel(AudioAttArgs a) {
    __shared__ float s0[64 * 8], s1[16 * 8], s2[8 * 8], s3[4 * 8], s4[2 * 8], s5[8], att[8];
    // y = x[..., :dim_att].permute(1, 0): [dim_att channels][8]
    for (int i = threadIdx.x; i < a.dim_att * 8; i += blockDim.x) {
        const int c = i / 8, t = i - c * 8;
        s0[c * 8 + t] = a.x[t * a.dim_feat + c];
    }
    __syncthreads();
    conv1d_k3(a.w[0], a.b[0], s0, s1, a.dim_att, 16, 8, 1);
    conv1d_k3(a.w[1], a.b[1], s1, s2, 16, 8, 8, 1);
    conv1d_k3(a.w[2], a.b[2], s2, s3, 8, 4, 8, 1);
    conv1d_k3(a.w[3], a.b[3], s3, s4, 4, 2, 8, 1);
    conv1d_k3(a.w[4], a.b[4], s4, s5, 2, 1, 8, 1);
    if (threadIdx.x == 0) {                    // Linear(8, 8) + Softmax(dim=1)
        float z[8], m = -3.4e38f;
        for (int o = 0; o < 8; ++o) {
            float acc = a.b[5][o];
            for (int c = 0; c < 8; ++c) acc = fmaf(a.w[5][o * 8 + c], s5[c], acc);
            z[o] = acc;
            m = fmaxf(m, acc);
        }
        float sum = 0.f;
        for (int o = 0; o < 8; ++o) { z[o] = expf(z[o] - m); sum += z[o]; }
        for (int o = 0; o < 8; ++o) att[o] = z[o] / sum;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < a.dim_feat; d += blockDim.x) {      // torch.sum(y * x, dim=0)
        float acc = 0.f;
        for (int t = 0; t < 8; ++t) acc = fmaf(att[t], a.x[t * a.dim_feat + d], acc);
        a.y[d] = acc;
    }
}


// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float lrelu_grad(float y) { return y > 0.f ? 1.f : 0.02f; }     // LeakyReLU(0.02, inplace): sign(out) == sign(in)

// Backward of conv1d_k3 + LeakyReLU.  d_out holds dL/d(out) on entry and is turned into dL/d(pre-activation) in place; gw / gb are
// accumulated with atomicAdd (several CTAs = samples add into the same parameter gradient); d_in (may be NULL) is overwritten.
__device__ __forceinline__ void conv1d_k3_bwd(const float* __restrict__ w, float* __restrict__ gw, float* __restrict__ gb, const float* in,
                                              const float* out, float* d_out, float* d_in, int cin, int cout, int len_in, int stride) {
    const int len_out = (len_in + 2 - 3) / stride + 1;
    for (int i = threadIdx.x; i < cout * len_out; i += blockDim.x) d_out[i] *= lrelu_grad(out[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < cout * cin * 3; i += blockDim.x) {           // gw[o][c][k] += sum_t dpre[o][t] * in[c][stride*t + k - 1]
        const int o = i / (cin * 3), r = i - o * cin * 3, c = r / 3, k = r - c * 3;
        float acc = 0.f;
        for (int t = 0; t < len_out; ++t) {
            const int p = stride * t + k - 1;
            if (p >= 0 && p < len_in) acc = fmaf(d_out[o * len_out + t], in[c * len_in + p], acc);
        }
        atomicAdd(gw + i, acc);
    }
    for (int o = threadIdx.x; o < cout; o += blockDim.x) {
        float acc = 0.f;
        for (int t = 0; t < len_out; ++t) acc += d_out[o * len_out + t];
        atomicAdd(gb + o, acc);
    }
    if (d_in) {
        for (int i = threadIdx.x; i < cin * len_in; i += blockDim.x) {         // d_in[c][p] = sum_o sum_k dpre[o][t] w[o][c][k], stride*t + k - 1 == p
            const int c = i / len_in, p = i - c * len_in;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int q = p + 1 - k;
                if (q >= 0 && q % stride == 0 && q / stride < len_out) {
                    const int t = q / stride;
                    for (int o = 0; o < cout; ++o) acc = fmaf(d_out[o * len_out + t], w[((size_t)o * cin + c) * 3 + k], acc);
                }
            }
            d_in[i] = acc;
        }
    }
    __syncthreads();
}

struct AudioNetBwdArgs {
    const float* w[6]; const float* b[6];
    float* gw[6]; float* gb[6];
    const float* x;      // (n, 16, 29)
    const float* dy;     // (n, dim_aud)
    int dim_aud;
};

__global__ void __launch_bounds__(128) audio_net_bwd_kernel(AudioNetBwdArgs a) {
    __shared__ float s0[29 * 16], s1[32 * 8], s2[32 * 4], s3[64 * 2], s4[64], s5[64];
    __shared__ float d1[32 * 8], d2[32 * 4], d3[64 * 2], d4[64], d5[64];
    const float* x = a.x + (size_t)blockIdx.x * 16 * 29;
    const float* dy = a.dy + (size_t)blockIdx.x * a.dim_aud;
    for (int i = threadIdx.x; i < 16 * 29; i += blockDim.x) {
        const int t = i / 29, c = i - t * 29;
        s0[c * 16 + t] = x[i];
    }
    __syncthreads();
    conv1d_k3(a.w[0], a.b[0], s0, s1, 29, 32, 16, 2);
    conv1d_k3(a.w[1], a.b[1], s1, s2, 32, 32, 8, 2);
    conv1d_k3(a.w[2], a.b[2], s2, s3, 32, 64, 4, 2);
    conv1d_k3(a.w[3], a.b[3], s3, s4, 64, 64, 2, 2);
    for (int o = threadIdx.x; o < 64; o += blockDim.x) {
        float acc = a.b[4][o];
        for (int c = 0; c < 64; ++c) acc = fmaf(a.w[4][o * 64 + c], s4[c], acc);
        s5[o] = lrelu(acc);
    }
    __syncthreads();
    // encoder_fc1.2: y = W5 s5 + b5
    for (int i = threadIdx.x; i < a.dim_aud * 64; i += blockDim.x) atomicAdd(a.gw[5] + i, dy[i >> 6] * s5[i & 63]);
    for (int o = threadIdx.x; o < a.dim_aud; o += blockDim.x) atomicAdd(a.gb[5] + o, dy[o]);
    for (int c = threadIdx.x; c < 64; c += blockDim.x) {
        float acc = 0.f;
        for (int o = 0; o < a.dim_aud; ++o) acc = fmaf(dy[o], a.w[5][o * 64 + c], acc);
        d5[c] = acc * lrelu_grad(s5[c]);                       // through the LeakyReLU after encoder_fc1.0
    }
    __syncthreads();
    // encoder_fc1.0: pre5 = W4 s4 + b4
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) atomicAdd(a.gw[4] + i, d5[i >> 6] * s4[i & 63]);
    for (int o = threadIdx.x; o < 64; o += blockDim.x) atomicAdd(a.gb[4] + o, d5[o]);
    for (int c = threadIdx.x; c < 64; c += blockDim.x) {
        float acc = 0.f;
        for (int o = 0; o < 64; ++o) acc = fmaf(d5[o], a.w[4][o * 64 + c], acc);
        d4[c] = acc;                                           // dL/d(s4) = dL/d(out of conv 4); its LeakyReLU is applied inside conv1d_k3_bwd
    }
    __syncthreads();
    conv1d_k3_bwd(a.w[3], a.gw[3], a.gb[3], s3, s4, d4, d3, 64, 64, 2, 2);
    conv1d_k3_bwd(a.w[2], a.gw[2], a.gb[2], s2, s3, d3, d2, 32, 64, 4, 2);
    conv1d_k3_bwd(a.w[1], a.gw[1], a.gb[1], s1, s2, d2, d1, 32, 32, 8, 2);
    conv1d_k3_bwd(a.w[0], a.gw[0], a.gb[0], s0, s1, d1, nullptr, 29, 32, 16, 2);      // the DeepSpeech features are data: no d_x
}

struct AudioAttBwdArgs {
    const float* w[6]; const float* b[6];
    float* gw[6]; float* gb[6];
    const float* x;      // (8, dim_feat)
    const float* dy;     // (dim_feat)
    float* dx;           // (8, dim_feat)
    int dim_att, dim_feat;
};

__global__ void __launch_bounds__(128) audio_att_bwd_kernel(AudioAttBwdArgs a) {
    __shared__ float s0[64 * 8], s1[16 * 8], s2[8 * 8], s3[4 * 8], s4[2 * 8], s5[8], att[8];
    __shared__ float d0[64 * 8], d1[16 * 8], d2[8 * 8], d3[4 * 8], d4[2 * 8], d5[8], datt[8], dz[8];
    for (int i = threadIdx.x; i < a.dim_att * 8; i += blockDim.x) {
        const int c = i / 8, t = i - c * 8;
        s0[c * 8 + t] = a.x[t * a.dim_feat + c];
    }
    __syncthreads();
    conv1d_k3(a.w[0], a.b[0], s0, s1, a.dim_att, 16, 8, 1);
    conv1d_k3(a.w[1], a.b[1], s1, s2, 16, 8, 8, 1);
    conv1d_k3(a.w[2], a.b[2], s2, s3, 8, 4, 8, 1);
    conv1d_k3(a.w[3], a.b[3], s3, s4, 4, 2, 8, 1);
    conv1d_k3(a.w[4], a.b[4], s4, s5, 2, 1, 8, 1);
    if (threadIdx.x == 0) {
        float z[8], m = -3.4e38f;
        for (int o = 0; o < 8; ++o) {
            float acc = a.b[5][o];
            for (int c = 0; c < 8; ++c) acc = fmaf(a.w[5][o * 8 + c], s5[c], acc);
            z[o] = acc;
            m = fmaxf(m, acc);
        }
        float sum = 0.f;
        for (int o = 0; o < 8; ++o) { z[o] = expf(z[o] - m); sum += z[o]; }
        for (int o = 0; o < 8; ++o) att[o] = z[o] / sum;
    }
    __syncthreads();
    // y[d] = sum_t att[t] x[t][d]:  d_att[t] = sum_d dy[d] x[t][d]  (one warp per pair of t), direct part of dx = att[t] dy[d]
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int t = warp; t < 8; t += 4) {
            float acc = 0.f;
            for (int d = lane; d < a.dim_feat; d += 32) acc = fmaf(a.dy[d], a.x[t * a.dim_feat + d], acc);
            acc = warp_sum(acc);
            if (lane == 0) datt[t] = acc;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {                    // softmax backward, then Linear(8, 8)
        float dot = 0.f;
        for (int o = 0; o < 8; ++o) dot = fmaf(att[o], datt[o], dot);
        for (int o = 0; o < 8; ++o) dz[o] = att[o] * (datt[o] - dot);
        for (int c = 0; c < 8; ++c) {
            float acc = 0.f;
            for (int o = 0; o < 8; ++o) acc = fmaf(dz[o], a.w[5][o * 8 + c], acc);
            d5[c] = acc;                       // dL/d(s5) = dL/d(out of conv 8)
        }
    }
    __syncthreads();
    if (threadIdx.x < 64) atomicAdd(a.gw[5] + threadIdx.x, dz[threadIdx.x >> 3] * s5[threadIdx.x & 7]);
    if (threadIdx.x < 8) atomicAdd(a.gb[5] + threadIdx.x, dz[threadIdx.x]);
    conv1d_k3_bwd(a.w[4], a.gw[4], a.gb[4], s4, s5, d5, d4, 2, 1, 8, 1);
    conv1d_k3_bwd(a.w[3], a.gw[3], a.gb[3], s3, s4, d4, d3, 4, 2, 8, 1);
    conv1d_k3_bwd(a.w[2], a.gw[2], a.gb[2], s2, s3, d3, d2, 8, 4, 8, 1);
    conv1d_k3_bwd(a.w[1], a.gw[1], a.gb[1], s1, s2, d2, d1, 16, 8, 8, 1);
    conv1d_k3_bwd(a.w[0], a.gw[0], a.gb[0], s0, s1, d1, d0, a.dim_att, 16, 8, 1);
    for (int i = threadIdx.x; i < 8 * a.dim_feat; i += blockDim.x) {
        const int t = i / a.dim_feat, d = i - t * a.dim_feat;
        a.dx[i] = att[t] * a.dy[d] + (d < a.dim_att ? d0[d * 8 + t] : 0.f);
    }
}

}  // namespace

extern "C" int inerf_audio_net_fwd(const float* const* params_host, const float* x, int n, int dim_aud, float* y, void* stream) {
    if (n < 0 || dim_aud <= 0 || dim_aud > 1024) return fail(INERF_E_SHAPE, "inerf_audio_net_fwd: bad n / dim_aud");
    if (n == 0) return INERF_OK;
    if (!params_host || !x || !y) return fail(INERF_E_ARG, "inerf_audio_net_fwd: NULL pointer");
    AudioNetArgs a{};
    for (int i = 0; i < 6; ++i) {
        if (!params_host[2 * i] || !params_host[2 * i + 1]) return fail(INERF_E_ARG, "inerf_audio_net_fwd: NULL parameter pointer");
        a.w[i] = params_host[2 * i]; a.b[i] = params_host[2 * i + 1];
    }
    a.x = x; a.y = y; a.dim_aud = dim_aud;
    audio_net_kernel<<<n, 128, 0, as_stream(stream)>>>(a);
    return check_launch("inerf_audio_net_fwd");
}

extern "C" int inerf_audio_att_fwd(const float* const* params_host, const float* x, int seq_len, int dim_feat, int dim_att, float* y,
                                   void* stream) {
    if (seq_len != 8) return fail(INERF_E_UNSUPPORTED, "inerf_audio_att_fwd: seq_len must be 8 (AudioAttNet default, audio_exp_nerf.py:225)");
    if (dim_feat <= 0 || dim_att <= 0 || dim_att > dim_feat || dim_att > 64) return fail(INERF_E_SHAPE, "inerf_audio_att_fwd: bad dims");
    if (!params_host || !x || !y) return fail(INERF_E_ARG, "inerf_audio_att_fwd: NULL pointer");
    AudioAttArgs a{};
    for (int i = 0; i < 6; ++i) {
        if (!params_host[2 * i] || !params_host[2 * i + 1]) return fail(INERF_E_ARG, "inerf_audio_att_fwd: NULL parameter pointer");
        a.w[i] = params_host[2 * i]; a.b[i] = params_host[2 * i + 1];
    }
    a.x = x; a.y = y; a.dim_att = dim_att; a.dim_feat = dim_feat;
    audio_att_kernel<<<1, 128, 0, as_stream(stream)>>>(a);
    return check_launch("inerf_audio_att_fwd");
}

extern "C" int inerf_audio_net_bwd(const float* const* params_host, float* const* grads_host, const float* x, const float* dy, int n,
                                   int dim_aud, void* stream) {
    if (n < 0 || dim_aud <= 0 || dim_aud > 1024) return fail(INERF_E_SHAPE, "inerf_audio_net_bwd: bad n / dim_aud");
    if (n == 0) return INERF_OK;
    if (!params_host || !grads_host || !x || !dy) return fail(INERF_E_ARG, "inerf_audio_net_bwd: NULL pointer");
    AudioNetBwdArgs a{};
    for (int i = 0; i < 6; ++i) {
        if (!params_host[2 * i] || !params_host[2 * i + 1] || !grads_host[2 * i] || !grads_host[2 * i + 1])
            return fail(INERF_E_ARG, "inerf_audio_net_bwd: NULL parameter / gradient pointer");
        a.w[i] = params_host[2 * i]; a.b[i] = params_host[2 * i + 1];
        a.gw[i] = grads_host[2 * i]; a.gb[i] = grads_host[2 * i + 1];
    }
    a.x = x; a.dy = dy; a.dim_aud = dim_aud;
    audio_net_bwd_kernel<<<n, 128, 0, as_stream(stream)>>>(a);
    return check_launch("inerf_audio_net_bwd");
}

extern "C" int inerf_audio_att_bwd(const float* const* params_host, float* const* grads_host, const float* x, const float* dy, int seq_len,
                                   int dim_feat, int dim_att, float* dx, void* stream) {
    if (seq_len != 8) return fail(INERF_E_UNSUPPORTED, "inerf_audio_att_bwd: seq_len must be 8 (AudioAttNet default, audio_exp_nerf.py:225)");
    if (dim_feat <= 0 || dim_att <= 0 || dim_att > dim_feat || dim_att > 64) return fail(INERF_E_SHAPE, "inerf_audio_att_bwd: bad dims");
    if (!params_host || !grads_host || !x || !dy || !dx) return fail(INERF_E_ARG, "inerf_audio_att_bwd: NULL pointer");
    AudioAttBwdArgs a{};
    for (int i = 0; i < 6; ++i) {
        if (!params_host[2 * i] || !params_host[2 * i + 1] || !grads_host[2 * i] || !grads_host[2 * i + 1])
            return fail(INERF_E_ARG, "inerf_audio_att_bwd: NULL parameter / gradient pointer");
        a.w[i] = params_host[2 * i]; a.b[i] = params_host[2 * i + 1];
        a.gw[i] = grads_host[2 * i]; a.gb[i] = grads_host[2 * i + 1];
    }
    a.x = x; a.dy = dy; a.dx = dx; a.dim_att = dim_att; a.dim_feat = dim_feat;
    audio_att_bwd_kernel<<<1, 128, 0, as_stream(stream)>>>(a);
    return check_launch("inerf_audio_att_bwd");
}
