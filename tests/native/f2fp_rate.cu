// Throughput of cvt.rn.relu.bf16x2.f32 (SASS F2FP.RELU.BF16.F32.PACK_AB) per SM sub-partition: 1, 2, 4 warps per SMSP, 256 independent
// conversions per warp per repetition, timed with clock64 on an otherwise idle SM.  nvcc -arch=sm_100a -O3 -o f2fp_rate f2fp_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t pack_relu(float lo, float hi) {
    uint32_t r;
    asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t pack_plain(float lo, float hi) {
    uint32_t r;
    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

template <int MODE>
__global__ void k(float* out, long long* cyc, int reps) {
    float v[32];
    for (int j = 0; j < 32; ++j) v[j] = (float)(threadIdx.x * 3 + j) * 0.37f - 5.f;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            uint32_t p;
            if (MODE == 0) p = pack_relu(v[2 * j], v[2 * j + 1]);
            else if (MODE == 1) p = pack_plain(v[2 * j], v[2 * j + 1]);
            else p = __float_as_uint(v[2 * j]) ^ __float_as_uint(v[2 * j + 1]);
            acc ^= p;
            v[2 * j] += 1.0f;          // keep the inputs changing (1 FADD per conversion)
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    const int reps = 1000;
    for (int warps = 4; warps <= 16; warps *= 2) {
        for (int mode = 0; mode < 3; ++mode) {
            if (mode == 0) k<0><<<1, warps * 32>>>(out, cyc, reps);
            else if (mode == 1) k<1><<<1, warps * 32>>>(out, cyc, reps);
            else k<2><<<1, warps * 32>>>(out, cyc, reps);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("warps/SMSP %d  %s: %.2f cycles per warp-instruction-slot (16 conv + 16 FADD per rep per warp): %.1f cycles/rep, %.2f cycles per conversion per SMSP\n",
                   warps / 4, mode == 0 ? "F2FP.RELU" : (mode == 1 ? "F2FP     " : "LOP3     "), (double)c / reps / 32.0, (double)c / reps, (double)c / reps / 16.0 / (warps / 4));
        }
    }
    return 0;
}
