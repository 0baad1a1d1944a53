// Stand-alone B200 probe for the tcgen05 building blocks used by csrc/mlp_bf16.cu:
// bulk-copy (TMA engine) -> 128B-swizzled K-major smem operands -> tcgen05.mma (kind::f16, bf16 in,
// fp32 accumulate in TMEM) -> tcgen05.commit -> tcgen05.ld.  Integer-valued inputs make the result
// exact, so any descriptor / layout mistake shows up as a mismatch count, not a tolerance question.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../ideal-nerf_b200/csrc/sm100_ptx.cuh"

using namespace sm100;

struct ProbeArgs {
    const uint8_t* a_img;   // KB blocks of 128x64 bf16, swizzled image (16 KB each)
    const uint8_t* b_img;   // KB blocks of  N x64 bf16, swizzled image
    float* d;               // [128][N]
    int N, KB, tmem_col, a_smem_off;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full, bar_done;
    __shared__ uint32_t tmem_base_slot;
    uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = base + p.a_smem_off;
    uint8_t* sB = sA + p.KB * 16384;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t a_bytes = p.KB * 16384, b_bytes = p.KB * p.N * 128;

    if (threadIdx.x == 0) {
        mbar_init(&bar_full, 1);
        mbar_init(&bar_done, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar_full, a_bytes + b_bytes);
        bulk_g2s(sA, p.a_img, a_bytes, &bar_full);
        bulk_g2s(sB, p.b_img, b_bytes, &bar_full);
        mbar_wait(&bar_full, 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16(128, p.N);
        for (int kb = 0; kb < p.KB; ++kb) {
            const uint64_t ad = umma_desc_sw128(smem_u32(sA + kb * 16384));
            const uint64_t bd = umma_desc_sw128(smem_u32(sB + kb * p.N * 128));
            for (int k = 0; k < 4; ++k)   // 16 bf16 = 32 B per K step: advance the start address field by 2
                umma_bf16(tmem_base + p.tmem_col, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&bar_done);
    }
    __syncwarp();
    mbar_wait(&bar_done, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < p.N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + p.tmem_col + c0, r);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) p.d[(size_t)(warp * 32 + lane) * p.N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static uint16_t f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return (uint16_t)(u >> 16);   // inputs are small integers: exact
}

static int run_case(int N, int KB, int tmem_col, int a_off) {
    const int K = KB * 64;
    std::vector<float> A(128 * K), B((size_t)N * K), D((size_t)128 * N), R((size_t)128 * N, 0.f);
    srand(N * 131 + KB * 7 + tmem_col);
    for (auto& v : A) v = (float)(rand() % 9 - 4);
    for (auto& v : B) v = (float)(rand() % 9 - 4);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int k = 0; k < K; ++k) s += A[m * K + k] * B[(size_t)n * K + k];
            R[(size_t)m * N + n] = s;
        }
    std::vector<uint8_t> ai((size_t)KB * 16384), bi((size_t)KB * N * 128);
    for (int kb = 0; kb < KB; ++kb) {
        for (int m = 0; m < 128; ++m)
            for (int c = 0; c < 64; ++c) {
                uint16_t h = f2bf(A[m * K + kb * 64 + c]);
                memcpy(&ai[(size_t)kb * 16384 + sw128_offset(m, c)], &h, 2);
            }
        for (int n = 0; n < N; ++n)
            for (int c = 0; c < 64; ++c) {
                uint16_t h = f2bf(B[(size_t)n * K + kb * 64 + c]);
                memcpy(&bi[(size_t)kb * N * 128 + sw128_offset(n, c)], &h, 2);
            }
    }
    uint8_t *da, *db;
    float* dd;
    cudaMalloc(&da, ai.size());
    cudaMalloc(&db, bi.size());
    cudaMalloc(&dd, D.size() * 4);
    cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xff, D.size() * 4);
    ProbeArgs p{da, db, dd, N, KB, tmem_col, a_off};
    size_t smem = 1024 + a_off + ai.size() + bi.size();
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("case N=%d KB=%d col=%d off=%d: CUDA error %s\n", N, KB, tmem_col, a_off, cudaGetErrorString(e));
        return 1;
    }
    cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (size_t i = 0; i < D.size(); ++i) bad += (D[i] != R[i]);
    printf("case N=%3d KB=%d tmem_col=%3d a_off=%6d: %ld / %zu mismatches%s\n", N, KB, tmem_col, a_off, bad, D.size(),
           bad ? "  <-- FAIL" : "  ok");
    if (bad) {
        for (int i = 0, shown = 0; i < (int)D.size() && shown < 6; ++i)
            if (D[i] != R[i]) { printf("   [m=%d n=%d] got %g want %g\n", i / N, i % N, D[i], R[i]); ++shown; }
    }
    cudaFree(da); cudaFree(db); cudaFree(dd);
    return bad != 0;
}

int main() {
    int fails = 0;
    fails += run_case(256, 1, 0, 0);
    fails += run_case(128, 1, 0, 0);
    fails += run_case(64, 1, 0, 0);
    fails += run_case(128, 2, 256, 0);
    fails += run_case(256, 4, 256, 16384);
    fails += run_case(64, 2, 448, 65536);
    printf(fails ? "UMMA PROBE: %d case(s) FAILED\n" : "UMMA PROBE: all cases passed\n", fails);
    return fails ? 1 : 0;
}
