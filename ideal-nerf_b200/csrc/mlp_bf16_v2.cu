// FaceNeRF MLP, bf16 mode, inference kernel v2: CTA PAIRS (tcgen05 cta_group::2) with whole-layer MMAs and the two 128-row slots
// of a CTA taking turns on the tensor pipe.
//
// Reference: models/face_nerf.py:40-80, NeRFs/HeadNeRF/train/audio_exp_nerf.py:332,376-394, NeRFs/HeadNeRF/helper.py:174-204 (same
// folded network as mlp_bf16.cu, whose packed weight blob and bias tiles this kernel reads unchanged).
//
// Why (profiles/r01_mlp_ablation.txt, profiles/r01_umma_rate_probe.txt): in v1 every layer is two N = 128 output halves so that an
// epilogue can overlap the other half; N = 128 MMAs top out at 86-92 % of the tensor pipe next to epilogue traffic, the halves force an
// in-place write-after-read protocol, and the narrow layers expose the MMA -> epilogue -> MMA round trip.  Here:
//   * two CTAs of a cluster form one M = 256 MMA (128 rows from each CTA, one issuing thread in the leader), N = the whole layer width
//     (256 / 128): 100 % issue rate in the probe, half the MMA instructions, B operand split across the pair (each CTA streams and
//     keeps only HALF of every weight stage: same L2 traffic per point as v1 although every stage is now streamed once per slot);
//   * the slots alternate: layer l of slot A (16 MMAs, 2048 cycles), then layer l of slot B while the epilogue of slot A drains its
//     accumulator and rewrites its activations, and so on -- a full layer of cover for every epilogue, no halves, no in-place hazard
//     (a slot's activations are only rewritten after ALL its MMAs of the layer have completed);
//   * cross-CTA protocol: tcgen05.commit multicasts "stage free" / "slot accumulator complete" to both CTAs; the peer's epilogue and
//     positional-encoding warps arrive remotely on the leader's mbarriers; the peer's warp 1 relays "my half of the stage has landed".
#include <cuda_bf16.h>

#include "mlp_common.cuh"
#include "sm100_ptx.cuh"

using namespace inerf;
using namespace sm100;

namespace {

constexpr int NS = 3;                  // weight stages (16 KB each: this CTA's half of a [N x 64] K-block)
constexpr int STAGE_BYTES = 16384;
constexpr int RMAX = 4;                // rays a 128-row slot can touch (s >= 43)
constexpr int NT = 512;

__host__ __device__ constexpr int lay_N(int l) { return l < 8 ? 256 : 128; }
__host__ __device__ constexpr int lay_act_kb(int l) { return l == 0 ? 0 : (l <= 8 ? 4 : 2); }
__host__ __device__ constexpr bool lay_pe(int l) { return l == 0 || l == 5; }
__host__ __device__ constexpr int lay_cnt(int l) { return lay_act_kb(l) + (lay_pe(l) ? 1 : 0); }

constexpr int OFF_ACT = 0;                               // [2 slots][4][16384]
constexpr int OFF_PE = 131072;                           // [2 slots][16384]
constexpr int OFF_W = 163840;                            // [NS][16384]
constexpr int OFF_ONES = OFF_W + NS * STAGE_BYTES;       // 128 x 16 bf16 ones, no-swizzle
constexpr int OFF_BT = OFF_ONES + 4096;                  // [2][4096] bias tiles (this CTA's rows)
constexpr int OFF_SB = OFF_BT + 2 * 4096;                // alpha_linear.bias, rgb_linear.bias
constexpr int OFF_AW = OFF_SB + 16;
constexpr int OFF_RW = OFF_AW + 1024;
constexpr int OFF_DIRB = OFF_RW + 1536;                  // [2][RMAX][128] floats
constexpr int OFF_BAR = OFF_DIRB + 2 * RMAX * 128 * 4;
constexpr int SMEM_V2 = OFF_BAR + 256;
static_assert(SMEM_V2 <= 232448, "shared memory budget");

struct Bars {
    uint64_t wfull[NS], wempty[NS];   // local: my half of the stage landed / both CTAs' MMAs on it complete (multicast commit)
    uint64_t pfull[NS];               // leader: the peer's half landed (relayed)
    uint64_t bfull[2], bempty[2], pbfull[2];   // the same three for the bias-tile ring
    uint64_t cbar[2];                 // per slot: all MMAs of its current layer complete (multicast commit)
    uint64_t ebar[2];                 // leader, per slot: epilogue done in BOTH CTAs (4 + 4 warps)
    uint64_t pe_ready;                // leader: gamma(p) of the next chunk written in both CTAs (4 + 4 warps)
    uint64_t pe_free;                 // multicast commit after the last MMA that reads gamma(p)
    uint64_t dirb_ready, dirb_free;   // local: per-ray view bias (4 PE warps -> 8 epilogue warps and back)
    uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 256, "barrier block");

struct V2Offsets { uint32_t off[11][2][5]; };             // byte offset of the stage image (layer, rank, K-block index) in the v1 blob
__constant__ V2Offsets c_v2;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(local_bar)), "r"(rank)
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded waits: a lost arrival traps after ~4 s instead of hanging the GPU
// (bounded by a retry count, not by clock64: every try_wait already suspends the warp for a hardware time slice, and reading the
    // clock in the loop costs issue slots that the epilogue warp sharing the scheduler needs)
__device__ __forceinline__ void wait_l(uint64_t* bar, uint32_t parity) {
    for (uint32_t n = 0; !mbar_try_wait(bar, parity); ++n)
        if (n > (1u << 24)) __trap();
}
__device__ __forceinline__ void wait_c(uint64_t* bar, uint32_t parity) {
    for (uint32_t n = 0; !mbar_try_wait_cluster(bar, parity); ++n)
        if (n > (1u << 24)) __trap();
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
        : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs once every previously issued MMA has completed
__device__ __forceinline__ void commit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

constexpr uint32_t HI_SW128 = (1024u >> 4) | (1u << 14) | ((uint32_t)SWIZZLE_128B << 29);
constexpr uint32_t HI_NOSWZ = (256u >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t lo_sw128(uint32_t addr) { return ((addr & 0x3FFFF) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t lo_noswz(uint32_t addr) { return ((addr & 0x3FFFF) >> 4) | ((128u >> 4) << 16); }

// One 32-column chunk of an epilogue: (+ per-ray view bias), ReLU, bf16 pack; KIND 1 also accumulates alpha_linear, KIND 3 rgb_linear.
template <int KIND>
__device__ __forceinline__ void epi_chunk(const uint32_t (&r)[32], uint32_t (&pk)[16], const float* __restrict__ dsrc,
                                          const float* __restrict__ aw, const float* __restrict__ rw, float& alpha, float& rgb0,
                                          float& rgb1, float& rgb2) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        float v[4] = {__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])};
        if constexpr (KIND == 2) {
            const float4 d = *reinterpret_cast<const float4*>(dsrc + j);
            v[0] += d.x; v[1] += d.y; v[2] += d.z; v[3] += d.w;
        }
        if constexpr (KIND == 1) {
            const float4 w = *reinterpret_cast<const float4*>(aw + j);
            alpha = fmaf(fmaxf(v[0], 0.f), w.x, alpha); alpha = fmaf(fmaxf(v[1], 0.f), w.y, alpha);
            alpha = fmaf(fmaxf(v[2], 0.f), w.z, alpha); alpha = fmaf(fmaxf(v[3], 0.f), w.w, alpha);
        }
        if constexpr (KIND == 3) {
            const float q[4] = {fmaxf(v[0], 0.f), fmaxf(v[1], 0.f), fmaxf(v[2], 0.f), fmaxf(v[3], 0.f)};
            const float4 w0 = *reinterpret_cast<const float4*>(rw + j);
            const float4 w1 = *reinterpret_cast<const float4*>(rw + 128 + j);
            const float4 w2 = *reinterpret_cast<const float4*>(rw + 256 + j);
            rgb0 = fmaf(q[0], w0.x, rgb0); rgb0 = fmaf(q[1], w0.y, rgb0); rgb0 = fmaf(q[2], w0.z, rgb0); rgb0 = fmaf(q[3], w0.w, rgb0);
            rgb1 = fmaf(q[0], w1.x, rgb1); rgb1 = fmaf(q[1], w1.y, rgb1); rgb1 = fmaf(q[2], w1.z, rgb1); rgb1 = fmaf(q[3], w1.w, rgb1);
            rgb2 = fmaf(q[0], w2.x, rgb2); rgb2 = fmaf(q[1], w2.y, rgb2); rgb2 = fmaf(q[2], w2.z, rgb2); rgb2 = fmaf(q[3], w2.w, rgb2);
        }
        pk[(j >> 1)] = pack_bf16x2_relu(v[0], v[1]);
        pk[(j >> 1) + 1] = pack_bf16x2_relu(v[2], v[3]);
    }
}

struct IssueCtx {
    Bars* bars;
    uint32_t a_lo, pe_lo, w_lo, ones_lo, bt_lo, tmem_base;
    uint32_t stage, wpar, bslot, bpar;
    uint32_t n_layers[2];      // layers issued so far per slot (= epilogues to expect)
    uint32_t iter_ctr;
    long long t_e, t_w;        // profiling (a.trace != NULL): cycles waiting for epilogues / for weight + bias stages
    bool prof;
};

template <int L>
__device__ __forceinline__ void issue_both(IssueCtx& c) {
    constexpr int N = lay_N(L), NKB = lay_act_kb(L), CNT = lay_cnt(L);
    constexpr uint32_t IDESC = umma_idesc_bf16(256, N);
    Bars* bars = c.bars;
    long long q0 = c.prof ? clock64() : 0;
    // both slots' previous epilogues (both CTAs): accumulators drained, activations rewritten
    if (c.n_layers[0] > 0) {
        wait_c(&bars->ebar[0], (c.n_layers[0] - 1) & 1);
        wait_c(&bars->ebar[1], (c.n_layers[1] - 1) & 1);
    }
    if (L == 0) wait_c(&bars->pe_ready, c.iter_ctr & 1);
    if (c.prof) { const long long q1 = clock64(); c.t_e += q1 - q0; q0 = q1; }
    // bias: D = ones[256 x 16] . tile[N x 16]^T, overwrites both accumulators
    wait_l(&bars->bfull[c.bslot], c.bpar);
    wait_c(&bars->pbfull[c.bslot], c.bpar);
    tc_fence_after();
    if (c.prof) c.t_w += clock64() - q0;
    if (elect_one()) {
#pragma unroll
        for (int slot = 0; slot < 2; ++slot)
            umma2(c.tmem_base + slot * 256, c.ones_lo, HI_NOSWZ, c.bt_lo + c.bslot * (4096 >> 4), HI_NOSWZ, IDESC, 0u);
        commit2(&bars->bempty[c.bslot]);
    }
    __syncwarp();
    c.bslot ^= 1;
    if (c.bslot == 0) c.bpar ^= 1;
#pragma unroll
    for (int i = 0; i < CNT; ++i) {
        const bool is_pe = (i == NKB);
        if (c.prof) q0 = clock64();
        wait_l(&bars->wfull[c.stage], c.wpar);
        wait_c(&bars->pfull[c.stage], c.wpar);
        tc_fence_after();
        if (c.prof) c.t_w += clock64() - q0;
        if (elect_one()) {
            const uint32_t b_lo = c.w_lo + c.stage * (STAGE_BYTES >> 4);
#pragma unroll
            for (int slot = 0; slot < 2; ++slot) {
                const uint32_t a_lo = is_pe ? c.pe_lo + slot * (16384 >> 4) : c.a_lo + slot * (65536 >> 4) + i * (16384 >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma2(c.tmem_base + slot * 256, a_lo + 2 * k, HI_SW128, b_lo + 2 * k, HI_SW128, IDESC, 1u);
                if (i == CNT - 1) commit2(&bars->cbar[slot]);      // this slot's layer is complete: its epilogue may start
            }
            commit2(&bars->wempty[c.stage]);
            if (i == CNT - 1 && L == 5) commit2(&bars->pe_free);
        }
        __syncwarp();
        if (++c.stage == NS) { c.stage = 0; c.wpar ^= 1; }
    }
    ++c.n_layers[0];
    ++c.n_layers[1];
}

__global__ void __launch_bounds__(NT, 1) mlp_bf16_v2_kernel(MlpArgs a, int n_rays, int n_pairs) {
    extern __shared__ __align__(1024) uint8_t sm[];
    if ((smem_u32(sm) & 1023u) != 0) __trap();
    Bars* bars = reinterpret_cast<Bars*>(sm + OFF_BAR);
    float* s_sb = reinterpret_cast<float*>(sm + OFF_SB);
    float* s_aw = reinterpret_cast<float*>(sm + OFF_AW);
    float* s_rw = reinterpret_cast<float*>(sm + OFF_RW);
    float* s_dirb = reinterpret_cast<float*>(sm + OFF_DIRB);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const long long pair = blockIdx.x >> 1;
    const long long n_chunks = (a.P + 255) / 256;          // 256 points per CTA per iteration, 512 per pair

    if (tid < 4) s_sb[tid] = a.cond[8 * 256 + 3 * 128 + tid];
    for (int i = tid; i < 256; i += NT) s_aw[i] = a.w[P_ALPHA_W][i];
    for (int i = tid; i < 384; i += NT) s_rw[i] = a.w[P_RGB_W][i];
    for (int i = tid; i < 256; i += NT)
        reinterpret_cast<uint4*>(sm + OFF_ONES)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&bars->wfull[s], 1); mbar_init(&bars->wempty[s], 1); mbar_init(&bars->pfull[s], 1); }
        for (int j = 0; j < 2; ++j) {
            mbar_init(&bars->bfull[j], 1); mbar_init(&bars->bempty[j], 1); mbar_init(&bars->pbfull[j], 1);
            mbar_init(&bars->cbar[j], 1);
            mbar_init(&bars->ebar[j], 8);
        }
        mbar_init(&bars->pe_ready, 8);
        mbar_init(&bars->pe_free, 1);
        mbar_init(&bars->dirb_ready, 4);
        mbar_init(&bars->dirb_free, 8);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc2(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                    // both CTAs' barriers and TMEM exist before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ================= weight + bias producer: this CTA's half of every stage, once per slot =================
        if (lane == 0) {
            const uint8_t* blob = reinterpret_cast<const uint8_t*>(a.packed);
            const uint8_t* tiles = reinterpret_cast<const uint8_t*>(a.cond + 2436);
            uint32_t g = 0, bh = 0;
            for (long long c0 = 2 * pair; c0 < n_chunks; c0 += 2 * n_pairs)
                for (int l = 0; l < 11; ++l) {
                    const uint32_t nh = lay_N(l) >> 1;                     // my rows of the B operand
                    const uint32_t toff = l < 8 ? (2 * l + rank) * 4096u : 65536u + (2 * (l - 8) + rank) * 2048u;
                    {
                        const uint32_t bs = bh & 1;
                        wait_l(&bars->bempty[bs], ((bh >> 1) & 1) ^ 1);
                        mbar_arrive_expect_tx(&bars->bfull[bs], nh * 32u);
                        bulk_g2s(sm + OFF_BT + bs * 4096, tiles + toff, nh * 32u, &bars->bfull[bs]);
                        ++bh;
                        const int cnt = lay_cnt(l);
                        for (int i = 0; i < cnt; ++i, ++g) {
                            const uint32_t stage = g % NS, round = g / NS;
                            wait_l(&bars->wempty[stage], (round & 1) ^ 1);
                            mbar_arrive_expect_tx(&bars->wfull[stage], nh * 128u);
                            bulk_g2s(sm + OFF_W + stage * STAGE_BYTES, blob + c_v2.off[l][rank][i], nh * 128u, &bars->wfull[stage]);
                        }
                    }
                }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ================= MMA issuer (leader; whole warp converged, one elected lane issues) =================
            IssueCtx c;
            c.bars = bars;
            c.a_lo = lo_sw128(smem_u32(sm + OFF_ACT));
            c.pe_lo = lo_sw128(smem_u32(sm + OFF_PE));
            c.w_lo = lo_sw128(smem_u32(sm + OFF_W));
            c.ones_lo = lo_noswz(smem_u32(sm + OFF_ONES));
            c.bt_lo = lo_noswz(smem_u32(sm + OFF_BT));
            c.tmem_base = tmem_base;
            c.stage = 0; c.wpar = 0; c.bslot = 0; c.bpar = 0; c.n_layers[0] = c.n_layers[1] = 0; c.iter_ctr = 0;
            c.t_e = c.t_w = 0; c.prof = a.trace != nullptr;
            const long long t_tot = clock64();
            for (long long c0 = 2 * pair; c0 < n_chunks; c0 += 2 * n_pairs, ++c.iter_ctr) {
                issue_both<0>(c); issue_both<1>(c); issue_both<2>(c); issue_both<3>(c); issue_both<4>(c); issue_both<5>(c);
                issue_both<6>(c); issue_both<7>(c); issue_both<8>(c); issue_both<9>(c); issue_both<10>(c);
            }
            if (c.prof && lane == 0 && blockIdx.x == 0) {
                a.trace[0] = (float)(clock64() - t_tot); a.trace[1] = (float)c.t_e; a.trace[2] = (float)c.t_w; a.trace[3] = (float)c.iter_ctr;
            }
        } else if (lane == 0) {
            // ================= relay (peer): tell the leader when MY half of a bias tile / weight stage has landed =================
            uint32_t g = 0, bh = 0;
            for (long long c0 = 2 * pair; c0 < n_chunks; c0 += 2 * n_pairs)
                for (int l = 0; l < 11; ++l) {
                        wait_l(&bars->bfull[bh & 1], (bh >> 1) & 1);
                        mbar_arrive_remote(&bars->pbfull[bh & 1], 0);
                        ++bh;
                        const int cnt = lay_cnt(l);
                        for (int i = 0; i < cnt; ++i, ++g) {
                            wait_l(&bars->wfull[g % NS], (g / NS) & 1);
                            mbar_arrive_remote(&bars->pfull[g % NS], 0);
                        }
                    }
        }
    } else if (warp >= 4 && warp < 12) {
        // ================= epilogue: one row per thread, whole layer width per pass =================
        const int slot = (warp - 4) >> 2;
        const int row = ((warp & 3) << 5) + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) << 5) << 16) + slot * 256;
        uint8_t* act = sm + OFF_ACT + slot * 65536;
        const uint32_t row_off = (row >> 3) * 1024 + (row & 7) * 128;
        const uint32_t rsw = row & 7;
        uint32_t n_layers = 0, iter_ctr = 0;
        const bool prof = a.trace != nullptr;
        long long te_wait = 0, te_work = 0, te_arr = 0;
        for (long long c0 = 2 * pair; c0 < n_chunks; c0 += 2 * n_pairs, ++iter_ctr) {
            const long long it = c0 + rank;                       // my 256-point chunk (may lie past the end: clamped, not written)
            const long long p0 = it * 256 + slot * 128;
            long long p = p0 + row;
            const bool in_range = p < a.P;
            if (!in_range) p = a.P - 1;
            const int ray_local = (int)(p / a.s - min(p0, a.P - 1) / a.s);
            float alpha = 0.f, rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
            for (int l = 0; l < 11; ++l, ++n_layers) {
                const int N = l < 8 ? 256 : 128;
                long long q0 = prof ? clock64() : 0;
                if (l == 8) wait_l(&bars->dirb_ready, iter_ctr & 1);
                wait_l(&bars->cbar[slot], n_layers & 1);
                __syncwarp();
                tc_fence_after();
                if (prof) { const long long q1 = clock64(); te_wait += q1 - q0; q0 = q1; }
                const int nchunk = N >> 5;                        // 8 or 4
                // NOT unrolled over the chunks: one ~1 KB body per layer kind that stays in the instruction cache (the unrolled form was
                // 9 KB per pass, evicted between passes by the issuer's straight-line schedule: the F2FP block stalled on fetch)
#pragma unroll 1
                for (int c = 0; c < nchunk; ++c) {
                    uint32_t r[32], pk[16];
                    tmem_ld32(t_lane + c * 32, r);
                    tmem_wait_ld();
                    const int f0 = c * 32;
#ifndef V2_EXP
#define V2_EXP 0
#endif
                    if (V2_EXP & 8) epi_chunk<0>(r, pk, nullptr, nullptr, nullptr, alpha, rgb0, rgb1, rgb2);
                    else
                    switch (l) {
                        case 7: epi_chunk<1>(r, pk, nullptr, s_aw + f0, nullptr, alpha, rgb0, rgb1, rgb2); break;
                        case 8: epi_chunk<2>(r, pk, s_dirb + (slot * RMAX + ray_local) * 128 + f0, nullptr, nullptr, alpha, rgb0, rgb1, rgb2); break;
                        case 10: epi_chunk<3>(r, pk, nullptr, nullptr, s_rw + f0, alpha, rgb0, rgb1, rgb2); break;
                        default: epi_chunk<0>(r, pk, nullptr, nullptr, nullptr, alpha, rgb0, rgb1, rgb2); break;
                    }
                    if (l != 10) {                            // every MMA of this slot's layer has completed: rewrite in place
                        uint8_t* kb = act + (f0 >> 6) * 16384 + row_off;
                        const int ch0 = (f0 & 63) >> 3;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<uint4*>(kb + (((ch0 + q) ^ rsw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                    }
                }
                tc_fence_before();
                if (prof) { const long long q1 = clock64(); te_work += q1 - q0; q0 = q1; }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_remote(&bars->ebar[slot], 0);
                    if (l == 8) mbar_arrive(&bars->dirb_free);
                }
                if (prof) te_arr += clock64() - q0;
            }
            if (in_range) {
                float4 o;
                o.x = rgb0 + s_sb[1]; o.y = rgb1 + s_sb[2]; o.z = rgb2 + s_sb[3]; o.w = alpha + s_sb[0];
                reinterpret_cast<float4*>(a.out)[p] = o;
            }
        }
        if (prof && blockIdx.x < 2 && lane == 0 && (warp == 4 || warp == 8)) {
            float* t = a.trace + 8 + (blockIdx.x * 2 + slot) * 4;
            t[0] = (float)te_wait; t[1] = (float)te_work; t[2] = (float)te_arr; t[3] = (float)iter_ctr;
        }
    } else if (warp >= 12) {
        // ================= positional encoding + per-ray view bias producers (as v1) =================
        const int t = tid - 12 * 32;
        float wdir[27];
        {
            const float* wrow = a.w[P_VIEWS_W] + (size_t)t * (283 + a.dim_expr) + 256;
#pragma unroll
            for (int j = 0; j < 27; ++j) wdir[j] = wrow[j];
        }
        uint32_t iter_ctr = 0;
        for (long long c0 = 2 * pair; c0 < n_chunks; c0 += 2 * n_pairs, ++iter_ctr) {
            const long long it = c0 + rank;
            uint32_t pk[2][32];
            if (V2_EXP & 16) {
#pragma unroll
                for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                    for (int j = 0; j < 32; ++j) pk[sl][j] = 0x3c003c00u;
            } else
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                long long p = it * 256 + sl * 128 + t;
                if (p > a.P - 1) p = a.P - 1;
                const long long ray = p / a.s;
                const float* r = a.rays + ray * a.ray_stride;
                const float zz = a.z[p];
                float v[64];
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = __fadd_rn(r[c], __fmul_rn(r[3 + c], zz));
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float sn, cs;
                    sincosf(v[c], &sn, &cs);
                    v[3 + c] = sn;
                    v[6 + c] = cs;
#pragma unroll
                    for (int f = 1; f < 10; ++f) {
                        const float s2 = 2.0f * sn * cs, c2 = fmaf(-2.0f * sn, sn, 1.0f);
                        sn = s2; cs = c2;
                        v[3 + 6 * f + c] = sn;
                        v[6 + 6 * f + c] = cs;
                    }
                }
                v[63] = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) pk[sl][j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
            }
            if (iter_ctr > 0) wait_l(&bars->pe_free, (iter_ctr - 1) & 1);
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                uint8_t* dst = sm + OFF_PE + sl * 16384 + (t >> 3) * 1024 + (t & 7) * 128;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<uint4*>(dst + ((q ^ (t & 7)) << 4)) = make_uint4(pk[sl][4 * q], pk[sl][4 * q + 1], pk[sl][4 * q + 2], pk[sl][4 * q + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&bars->pe_ready, 0);
            {
                float enc[27];
                const int q = lane;
                if (q < 2 * RMAX) {
                    const int sl = q / RMAX, rl = q - sl * RMAX;
                    long long pfirst = it * 256 + sl * 128;
                    if (pfirst > a.P - 1) pfirst = a.P - 1;
                    long long ray = pfirst / a.s + rl;
                    if (ray > n_rays - 1) ray = n_rays - 1;
                    const float* r = a.rays + ray * a.ray_stride + (a.ray_stride - 3);
                    enc[0] = r[0]; enc[1] = r[1]; enc[2] = r[2];
#pragma unroll
                    for (int f = 0; f < 4; ++f)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            float sn, cs;
                            sincosf(enc[c] * (float)(1 << f), &sn, &cs);
                            enc[3 + 6 * f + c] = sn;
                            enc[6 + 6 * f + c] = cs;
                        }
                } else {
#pragma unroll
                    for (int j = 0; j < 27; ++j) enc[j] = 0.f;
                }
                if (iter_ctr > 0) wait_l(&bars->dirb_free, (iter_ctr - 1) & 1);
                for (int q2 = 0; q2 < 2 * RMAX; ++q2) {
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < 27; ++j) acc = fmaf(wdir[j], __shfl_sync(0xffffffffu, enc[j], q2), acc);
                    s_dirb[q2 * 128 + t] = acc;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->dirb_ready);
            }
        }
    }

    // ---- teardown: nobody leaves while the peer may still arrive on / multicast to this CTA --------------------
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc2(tmem_base, 512);
}

}  // namespace

namespace inerf {

int mlp_bf16_v2_launch(const MlpArgs& a, cudaStream_t st) {
    if (a.s < 43) return fail(INERF_E_UNSUPPORTED, "bf16 mode needs at least 43 samples per ray (a 128-row slot may touch at most 4 rays)");
    if ((uintptr_t)a.packed & 15) return fail(INERF_E_ALIGN, "inerf_mlp_fwd: packed weights must be 16-byte aligned");
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        V2Offsets o{};
        uint32_t flat[11][2][5];
        mlp_bf16_stage_offsets(flat);
        for (int l = 0; l < 11; ++l)
            for (int h = 0; h < 2; ++h)
                for (int i = 0; i < 5; ++i) o.off[l][h][i] = flat[l][h][i];
        cudaError_t e = cudaFuncSetAttribute(mlp_bf16_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_V2);
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_v2, &o, sizeof(o));
        if (e != cudaSuccess) { set_error("mlp_bf16_v2: setup: %s", cudaGetErrorString(e)); return (int)e; }
        configured_dev = dev;
    }
    const long long n_chunks = (a.P + 255) / 256;
    long long pairs = (n_chunks + 1) / 2;
    const long long max_pairs = num_sms() / 2;
    if (pairs > max_pairs) pairs = max_pairs;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = SMEM_V2;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mlp_bf16_v2_kernel, a, (int)(a.P / a.s), (int)pairs);
    if (e != cudaSuccess) { set_error("inerf_mlp_fwd[bf16 v2]: %s", cudaGetErrorString(e)); return (int)e; }
    return check_launch("inerf_mlp_fwd[bf16 v2]");
}

}  // namespace inerf
