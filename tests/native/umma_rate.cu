// B200 probe: sustained tcgen05.mma issue rate of the operand-sourcing variants considered for csrc/mlp_bf16.cu
// (DESIGN.md 4.1), with and without epilogue-like shared-memory / tensor-memory traffic beside it.
//
//   variant            A operand         B operand                    CTA pair
//   ss1  N=128|256     smem descriptor   smem descriptor              cta_group::1
//   ts1  N=128|256     tensor memory     smem descriptor              cta_group::1
//   ss2  N=256         smem (128 rows per CTA)   smem (N/2 rows per CTA)   cta_group::2, M = 256
//   ts2  N=256         tensor memory     smem (N/2 rows per CTA)      cta_group::2, M = 256
//
// Every CTA (or pair) issues `iters` x 4 K-blocks x `slots` x 4 MMAs (K = 16 each) from one thread and reports
// cycles per MMA.  Correctness of the ts / cta_group::2 forms is checked on integer data first (exact in bf16/fp32).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../ideal-nerf_b200/csrc/sm100_ptx.cuh"

using namespace sm100;

namespace {

__device__ __forceinline__ void bounded_wait(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 4000000000LL) __trap();
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t* slot, uint32_t ncols) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr, uint32_t ncols) {
    if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int CG, bool TS>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a_desc_or_taddr, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if constexpr (!TS) {
        if constexpr (CG == 1)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                         "l"(a_desc_or_taddr), "l"(bdesc), "r"(idesc), "r"(acc)
                         : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                         "l"(a_desc_or_taddr), "l"(bdesc), "r"(idesc), "r"(acc)
                         : "memory");
    } else {
        const uint32_t ta = (uint32_t)a_desc_or_taddr;
        if constexpr (CG == 1)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                         "r"(ta), "l"(bdesc), "r"(idesc), "r"(acc)
                         : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                         "r"(ta), "l"(bdesc), "r"(idesc), "r"(acc)
                         : "memory");
    }
}

template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
    if constexpr (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                     "h"((uint16_t)3)
                     : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct Args {
    const uint8_t* a_img;   // correctness: [CG][KB=4][128 x 64] swizzled bf16 images (A rows of each CTA)
    const uint8_t* b_img;   // correctness: [CG][KB=4][N/CG x 64] swizzled images (B rows of each CTA)
    float* d;               // correctness: [128*CG][N]
    float* report;          // per CTA: {cycles, n_mma, epilogue loops}
    int N, slots, iters, interfere, check;
    int mimic;              // 0: back-to-back MMAs; 1: + a tcgen05.commit per 8-MMA step (nobody waits); 2: + half/layer commits answered by a responder warp
};

// smem map: A [2 slots][4 kb][16 KB] = 128 KB | B [4 kb][up to 32 KB]... B only one K-block image is kept resident per kb: 4 x N*128
constexpr int OFF_A = 0, OFF_B = 131072, OFF_SCR = OFF_B + 4 * 16384, SMEM = OFF_SCR + 16384 + 1024;
// (N = 256 with cta_group::1 needs 4 x 32 KB of B: the rate runs then alias the 4 K-blocks onto 2 images -- timing only.)

template <int CG, bool TS>
__global__ void __launch_bounds__(384, 1) rate_kernel(Args p) {
    extern __shared__ __align__(1024) uint8_t sm[];      // the only shared allocation: starts 1024-aligned
    if ((smem_u32(sm) & 1023u) != 0) __trap();
    uint64_t& bar_done = *reinterpret_cast<uint64_t*>(sm + OFF_SCR + 16384);
    uint64_t& bar_load = *reinterpret_cast<uint64_t*>(sm + OFF_SCR + 16384 + 8);
    uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(sm + OFF_SCR + 16384 + 16);
    volatile int& stop_flag = *reinterpret_cast<volatile int*>(sm + OFF_SCR + 16384 + 20);
    uint64_t* ring = reinterpret_cast<uint64_t*>(sm + OFF_SCR + 16384 + 32);     // [3] per-step commits, nobody waits
    uint64_t* cbar = ring + 3;                                                   // [2] C0, C2
    uint64_t* ebar = ring + 5;                                                   // [2] E0, E1 (8 responder warps)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
    const int NB = p.N / CG;                       // B rows held by this CTA
    const uint32_t b_img_bytes = (uint32_t)NB * 128u;
    const bool alias_b = !p.check && 4u * b_img_bytes > 65536u;
    const int off_b = p.check ? 65536 : OFF_B;     // correctness runs use one slot of A, so B (up to 128 KB) starts right after it

    if (threadIdx.x == 0) {
        mbar_init(&bar_done, 1);
        mbar_init(&bar_load, 1);
        for (int j = 0; j < 5; ++j) mbar_init(&ring[j], 1);
        mbar_init(&ebar[0], 8);
        mbar_init(&ebar[1], 8);
        fence_mbar_init();
        stop_flag = 0;
    }
    if (warp == 1) tmem_alloc_cg<CG>(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    // ---- operands -------------------------------------------------------------------------------------------
    if (p.check) {
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(&bar_load, 4 * 16384 + 4 * b_img_bytes);
            bulk_g2s(sm + OFF_A, p.a_img + (size_t)rank * 4 * 16384, 4 * 16384, &bar_load);
            bulk_g2s(sm + off_b, p.b_img + (size_t)rank * 4 * b_img_bytes, 4 * b_img_bytes, &bar_load);
        }
        bounded_wait(&bar_load, 0);
    } else {
        for (int i = threadIdx.x; i < (OFF_SCR) / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
        fence_proxy_async_smem();
    }
    __syncthreads();
    if (TS) {
        // A rows of this CTA -> tensor memory columns [256, 384): column 256 + kb*32 + j holds bf16 pair (k = 64 kb + 2j, +1) of row = lane
        if (warp >= 4 && warp < 8) {
            const int row = (warp & 3) * 32 + lane;
            for (int kb = 0; kb < 4; ++kb) {
                uint32_t r[32];
                for (int j = 0; j < 32; ++j) {
                    const uint32_t off = sw128_offset(row, 2 * j);
                    r[j] = *reinterpret_cast<const uint32_t*>(sm + OFF_A + kb * 16384 + off);
                }
                tmem_st32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256 + kb * 32, r);
            }
            tmem_wait_st();
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if constexpr (CG == 2) cluster_sync_all();

    // ---- issue ---------------------------------------------------------------------------------------------------
    if (warp == 0 && lane == 0 && rank == 0) {
        const uint32_t idesc = umma_idesc_bf16(128 * CG, p.N);
        const uint32_t a_base = smem_u32(sm + OFF_A), b_base = smem_u32(sm + off_b);
        const long long t0 = clock64();
        long long n_mma = 0;
        if (p.mimic) {
            // the issue pattern of mlp_bf16_kernel: layers of two output halves (N = 128 each, D = slot*256 + h*128), 4 K-blocks per
            // half, both slots per K-block, a commit per step (weight stage release), C0 / C2 commits per half, and (mimic >= 2) the
            // E0 / E1 waits answered by the responder warps below
            const uint32_t idh = umma_idesc_bf16(128, 128);
            uint32_t g = 0, lc = 0;
            for (int it = 0; it < p.iters; ++it, ++lc)
                for (int h = 0; h < 2; ++h)
                    for (int kb = 0; kb < 4; ++kb, ++g) {
                        if (p.mimic >= 2 && lc > 0 && h == 0 && (kb == 0 || kb == 2)) bounded_wait(&ebar[kb >> 1], (lc - 1) & 1);
                        const uint64_t bd = umma_desc_sw128(b_base + kb * 16384);
                        for (int slot = 0; slot < 2; ++slot) {
                            const uint32_t d = tmem_base + slot * 256 + h * 128;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t a = umma_desc_sw128(a_base + slot * 65536 + kb * 16384) + 2 * k;
                                mma<CG, false>(d, a, bd + 2 * k, idh, (kb | k) ? 1u : 0u);
                                ++n_mma;
                            }
                        }
                        commit<CG>(&ring[g % 3]);
                        if (kb == 3) commit<CG>(&cbar[h]);
                    }
        } else
        for (int it = 0; it < p.iters; ++it)
            for (int kb = 0; kb < 4; ++kb) {
                const uint64_t bd = umma_desc_sw128(b_base + (alias_b ? (kb & 1) : kb) * b_img_bytes);
                for (int slot = 0; slot < p.slots; ++slot) {
                    const uint32_t d = tmem_base + (TS ? 0 : slot * 256);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint64_t a;
                        if constexpr (TS) a = tmem_base + 256 + kb * 32 + k * 8;
                        else a = umma_desc_sw128(a_base + slot * 65536 + kb * 16384) + 2 * k;
                        mma<CG, TS>(d, a, bd + 2 * k, idesc, (it | kb | k) ? 1u : 0u);
                        ++n_mma;
                    }
                }
            }
        commit<CG>(&bar_done);
        bounded_wait(&bar_done, 0);
        const long long t1 = clock64();
        p.report[blockIdx.x * 4 + 0] = (float)(t1 - t0);
        p.report[blockIdx.x * 4 + 1] = (float)n_mma;
        stop_flag = 1;
    } else if (warp >= 4 && warp < 12 && p.mimic >= 2) {
        // responder = the epilogue's barrier protocol (optionally with its TMEM-load / st.shared work: interfere = 1)
        const int row = (warp & 3) * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 3) * 256;
        uint8_t* scr = sm + OFF_SCR + (row >> 3) * 1024 + (row & 7) * 128;
        uint32_t sink = 0;
        long long t_work = 0;
        for (int lc = 0; lc < p.iters; ++lc)
            for (int h = 0; h < 2; ++h) {
                bounded_wait(&cbar[h], lc & 1);
                __syncwarp();
                tc_fence_after();
                const long long w0 = clock64();
                if (p.interfere) {
                    uint32_t packed[64];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t r[32];
                        tmem_ld32(t_lane + h * 128 + c * 32, r);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; ++j) packed[c * 16 + j] = pack_bf16x2_relu(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                    }
                    tc_fence_before();
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        *reinterpret_cast<uint4*>(scr + (((q & 7) ^ (row & 7)) << 4)) =
                            make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                    fence_proxy_async_smem();
                    sink += packed[3];
                }
                t_work += clock64() - w0;
                __syncwarp();
                if (lane == 0) mbar_arrive(&ebar[h]);
            }
        if (lane == 0 && warp == 4) p.report[blockIdx.x * 4 + 3] = (float)t_work / (2.f * p.iters) + (sink == 0x12345u ? 1.f : 0.f);
    } else if (warp >= 4 && warp < 12 && p.interfere) {
        // epilogue-like traffic: tcgen05.ld of 128 fp32 columns per row + 16 x 16-byte swizzled st.shared (interfere = 1)
        // or + tcgen05.st of 64 packed columns (interfere = 2)
        const int row = (warp & 3) * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (TS ? 0 : (warp >> 3) * 256);
        uint8_t* scr = sm + OFF_SCR + (row >> 3) * 1024 + (row & 7) * 128;
        long long loops = 0;
        uint32_t sink = 0;
        if (p.interfere >= 3) {
            // register-only work next to the MMAs: 64 cvt.rn.relu.bf16x2.f32 (interfere = 3) or 64 LOP3 (interfere = 4) per loop
            float v[32];
            for (int j = 0; j < 32; ++j) v[j] = (float)(threadIdx.x * 3 + j) * 0.37f - 5.f;
            while (!stop_flag) {
#pragma unroll
                for (int rep = 0; rep < 4; ++rep)
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        uint32_t q;
                        if (p.interfere == 3) q = pack_bf16x2_relu(v[2 * j], v[2 * j + 1]);
                        else q = __float_as_uint(v[2 * j]) ^ __float_as_uint(v[2 * j + 1]);
                        sink ^= q;
                        v[2 * j] += 1.0f;
                    }
                ++loops;
            }
        } else
        while (!stop_flag) {
            uint32_t packed[64];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld32(t_lane + c * 32, r);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) packed[c * 16 + j] = pack_bf16x2_relu(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            }
            if (p.interfere == 1) {
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    *reinterpret_cast<uint4*>(scr + (((q & 7) ^ (row & 7)) << 4)) =
                        make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                fence_proxy_async_smem();
            } else {
                uint32_t a[32], b[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) { a[j] = packed[j]; b[j] = packed[32 + j]; }
                const uint32_t t_st = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
                tmem_st32(t_st + 448, a);            // columns never read by the ts-mode MMAs of this probe
                tmem_st32(t_st + 480, b);
                tmem_wait_st();
            }
            sink += packed[5];
            ++loops;
        }
        if (lane == 0 && warp == 4) p.report[blockIdx.x * 4 + 2] = (float)loops + (sink == 0x12345u ? 1.f : 0.f);
    }
    if (CG == 2 && rank == 1 && threadIdx.x == 0) {       // the multicast commit reaches the peer too: release its epilogue warps
        bounded_wait(&bar_done, 0);
        stop_flag = 1;
    }
    __syncthreads();

    // ---- correctness read-back ---------------------------------------------------------------------------------------
    if (p.check) {
        if (threadIdx.x == 32) bounded_wait(&bar_done, 0);
        __syncthreads();
        tc_fence_after();
        if (warp >= 4 && warp < 8) {
            const int row = (warp & 3) * 32 + lane;
            for (int c0 = 0; c0 < p.N; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + c0, r);
                tmem_wait_ld();
                for (int j = 0; j < 32; ++j) p.d[(size_t)(rank * 128 + row) * p.N + c0 + j] = __uint_as_float(r[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    if (warp == 1) tmem_dealloc_cg<CG>(tmem_base, 512);
}

uint16_t f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return (uint16_t)(u >> 16);
}

template <int CG, bool TS>
int launch(const Args& a, int grid) {
    cudaFuncSetAttribute(rate_kernel<CG, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<CG, TS>, a);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("   CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}

template <int CG, bool TS>
int check_case(int N) {
    const int M = 128 * CG, K = 256;
    std::vector<float> A((size_t)M * K), B((size_t)N * K), R((size_t)M * N), D((size_t)M * N);
    srand(N * 7 + CG * 3 + TS);
    for (auto& v : A) v = (float)(rand() % 7 - 3);
    for (auto& v : B) v = (float)(rand() % 7 - 3);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int k = 0; k < K; ++k) s += A[(size_t)m * K + k] * B[(size_t)n * K + k];
            R[(size_t)m * N + n] = s;
        }
    const int NB = N / CG;
    std::vector<uint8_t> ai((size_t)CG * 4 * 16384), bi((size_t)CG * 4 * NB * 128);
    for (int r = 0; r < CG; ++r)
        for (int kb = 0; kb < 4; ++kb) {
            for (int m = 0; m < 128; ++m)
                for (int c = 0; c < 64; ++c) {
                    uint16_t h = f2bf(A[(size_t)(r * 128 + m) * K + kb * 64 + c]);
                    memcpy(&ai[((size_t)r * 4 + kb) * 16384 + sw128_offset(m, c)], &h, 2);
                }
            for (int n = 0; n < NB; ++n)
                for (int c = 0; c < 64; ++c) {
                    uint16_t h = f2bf(B[(size_t)(r * NB + n) * K + kb * 64 + c]);
                    memcpy(&bi[((size_t)r * 4 + kb) * NB * 128 + sw128_offset(n, c)], &h, 2);
                }
        }
    uint8_t *da, *db; float *dd, *rep;
    cudaMalloc(&da, ai.size()); cudaMalloc(&db, bi.size()); cudaMalloc(&dd, D.size() * 4); cudaMalloc(&rep, 4096 * 4);
    cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xff, D.size() * 4);
    Args a{da, db, dd, rep, N, 1, 1, 0, 1, 0};
    int rc = launch<CG, TS>(a, CG);
    long bad = 0;
    if (!rc) {
        cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < D.size(); ++i) bad += (D[i] != R[i]);
        if (bad)
            for (size_t i = 0, shown = 0; i < D.size() && shown < 4; ++i)
                if (D[i] != R[i]) { printf("   [m=%zu n=%zu] got %g want %g\n", i / N, i % N, D[i], R[i]); ++shown; }
    }
    printf("check %s cta_group::%d N=%3d: %s (%ld mismatches)\n", TS ? "ts" : "ss", CG, N, (rc || bad) ? "FAIL" : "ok", bad);
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(rep);
    return rc || bad;
}


// ---- MN-major operands: D[m][n] = sum_p A[p][m] * B[p][n] with A, B stored point-major as [128 points][64 features] 128B-swizzled
// images (the layout the forward kernel keeps its activations in) -- the dW GEMM of the training path.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;      // next 64 MN elements
    d |= (uint64_t)(1024 >> 4) << 32;           // next 8 K rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)SWIZZLE_128B << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) mn_kernel(const uint8_t* a_img, const uint8_t* b_img, float* d, int N) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t& bar_load = *reinterpret_cast<uint64_t*>(sm + 98304);
    uint64_t& bar_done = *reinterpret_cast<uint64_t*>(sm + 98304 + 8);
    uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(sm + 98304 + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_done, 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc_cg<1>(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar_load, 2 * 16384 + (N / 64) * 16384);
        bulk_g2s(sm, a_img, 2 * 16384, &bar_load);
        bulk_g2s(sm + 32768, b_img, (N / 64) * 16384, &bar_load);
        bounded_wait(&bar_load, 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16(128, N) | (1u << 15) | (1u << 16);      // A and B MN-major
        for (int k = 0; k < 8; ++k) {
            const uint64_t ad = umma_desc_mn_sw128(smem_u32(sm) + k * 2048, 16384);
            const uint64_t bd = umma_desc_mn_sw128(smem_u32(sm + 32768) + k * 2048, 16384);
            mma<1, false>(tmem_base, ad, bd, idesc, k ? 1u : 0u);
        }
        commit<1>(&bar_done);
    }
    __syncwarp();
    bounded_wait(&bar_done, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) d[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc_cg<1>(tmem_base, 512);
}

int check_mn(int N) {
    const int P = 128, M = 128;
    std::vector<float> A((size_t)P * M), B((size_t)P * N), R((size_t)M * N), D((size_t)M * N);
    srand(N + 99);
    for (auto& v : A) v = (float)(rand() % 7 - 3);
    for (auto& v : B) v = (float)(rand() % 7 - 3);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int p = 0; p < P; ++p) s += A[(size_t)p * M + m] * B[(size_t)p * N + n];
            R[(size_t)m * N + n] = s;
        }
    std::vector<uint8_t> ai((size_t)2 * 16384), bi((size_t)(N / 64) * 16384);
    for (int p = 0; p < P; ++p) {
        for (int m = 0; m < M; ++m) { uint16_t h = f2bf(A[(size_t)p * M + m]); memcpy(&ai[(size_t)(m / 64) * 16384 + sw128_offset(p, m % 64)], &h, 2); }
        for (int n = 0; n < N; ++n) { uint16_t h = f2bf(B[(size_t)p * N + n]); memcpy(&bi[(size_t)(n / 64) * 16384 + sw128_offset(p, n % 64)], &h, 2); }
    }
    uint8_t *da, *db; float* dd;
    cudaMalloc(&da, ai.size()); cudaMalloc(&db, bi.size()); cudaMalloc(&dd, D.size() * 4);
    cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xff, D.size() * 4);
    cudaFuncSetAttribute(mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 64);
    mn_kernel<<<1, 128, 98304 + 64>>>(da, db, dd, N);
    cudaError_t e = cudaDeviceSynchronize();
    long bad = 0;
    if (e != cudaSuccess) { printf("   CUDA error: %s\n", cudaGetErrorString(e)); bad = -1; }
    else {
        cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < D.size(); ++i) bad += (D[i] != R[i]);
        if (bad)
            for (size_t i = 0, shown = 0; i < D.size() && shown < 4; ++i)
                if (D[i] != R[i]) { printf("   [m=%zu n=%zu] got %g want %g\n", i / N, i % N, D[i], R[i]); ++shown; }
    }
    printf("check MN-major A/B (point-major images) N=%3d: %s (%ld mismatches)\n", N, bad ? "FAIL" : "ok", bad);
    cudaFree(da); cudaFree(db); cudaFree(dd);
    return bad != 0;
}

template <int CG, bool TS>
void rate_case(int N, int slots, int interfere, int grid, int mimic = 0) {
    float* rep;
    cudaMalloc(&rep, 4096 * 4);
    cudaMemset(rep, 0, 4096 * 4);
    Args a{nullptr, nullptr, nullptr, rep, N, slots, 200, interfere, 0, mimic};
    if (launch<CG, TS>(a, grid)) { cudaFree(rep); return; }
    std::vector<float> h(4096);
    cudaMemcpy(h.data(), rep, 4096 * 4, cudaMemcpyDeviceToHost);
    double cyc = 0, nm = 0, loops = 0, epi = 0; int cnt = 0;
    for (int b = 0; b < grid; b += CG) { cyc += h[b * 4]; nm += h[b * 4 + 1]; loops += h[b * 4 + 2]; epi += h[b * 4 + 3]; ++cnt; }
    cyc /= cnt; nm /= cnt; loops /= cnt; epi /= cnt;
    if (mimic >= 2) printf("[responder: %.0f cycles of work per half-layer epilogue (128 columns per row)] ", epi);
    const double mac = 128.0 * CG * N * 16, per = cyc / nm;
    if (mimic) printf("mimic=%d ", mimic);
    printf("rate %s cg%d N=%3d slots=%d interfere=%d grid=%3d: %7.1f cyc/MMA  -> %6.0f MAC/clk/SM (%.0f%% of 4096)   epilogue: %.0f rows-of-128col per kcyc per CTA\n",
           TS ? "ts" : "ss", CG, N, slots, interfere, grid, per, mac / per / CG, 100.0 * mac / per / CG / 4096.0,
           loops * 256.0 / (cyc / 1000.0));
    cudaFree(rep);
}

}  // namespace

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IONBF, 0);
    int fails = 0;
    fails += check_case<1, false>(128);
    fails += check_case<1, true>(128);
    fails += check_case<1, true>(256);
    fails += check_case<2, false>(256);
    fails += check_case<2, true>(256);
    fails += check_case<2, false>(128);
    fails += check_mn(256);
    fails += check_mn(64);
    if (argc > 1) { printf(fails ? "UMMA checks: %d FAILED\n" : "UMMA checks passed\n", fails); return fails != 0; }
    const int G = 148;
    for (int inter = 0; inter <= 2; ++inter) {
        if (inter < 2) {
            rate_case<1, false>(128, 2, inter, G);
            rate_case<1, false>(256, 2, inter, G);
            rate_case<2, false>(256, 2, inter, G);
            rate_case<2, false>(128, 2, inter, G);
        }
        rate_case<1, true>(128, 1, inter, G);
        rate_case<1, true>(256, 1, inter, G);
        rate_case<2, true>(256, 1, inter, G);
    }
    rate_case<1, false>(128, 2, 3, G); rate_case<1, false>(128, 2, 4, G);
    rate_case<1, false>(256, 2, 3, G); rate_case<1, false>(256, 2, 4, G);
    rate_case<2, false>(256, 2, 3, G); rate_case<2, false>(256, 2, 4, G);
    rate_case<1, false>(128, 2, 0, G, 1);
    rate_case<1, false>(128, 2, 0, G, 2);
    rate_case<1, false>(128, 2, 1, G, 2);
    printf(fails ? "UMMA RATE: %d check(s) FAILED\n" : "UMMA RATE: all checks passed\n", fails);
    return 0;
}
