// Throughput of tcgen05.ld (SASS LDTM) per SM: how fast can the epilogue warps read fp32 accumulators out of tensor memory?
// One CTA per SM, W warps; warp w may only touch TMEM lanes 32*(w%4)..+31, so W/4 warps share each lane quarter and walk different columns.
// Every repetition loads 128 columns per warp as 4 x (32x32b.x32) -- the access pattern of the MLP epilogue (csrc/mlp_bf16.cu) -- with a
// tcgen05.wait::ld after each load (MODE 0), after all four (MODE 1), or as 16-bit packed loads (.pack::16b, MODE 2: two adjacent columns
// per register, 32 registers = 64 columns).  The FaceNeRF layer needs 128 rows x 256 columns x 4 B = 128 KB read per 128-point slot per
// layer against 2048 cycles of MMA, i.e. 64 B/clk/SM just to keep up.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../ideal-nerf_b200/csrc -o tmem_ld_rate tmem_ld_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../ideal-nerf_b200/csrc/sm100_ptx.cuh"

using namespace sm100;

__device__ __forceinline__ void tmem_ld32_pack16(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
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

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, long long* cyc, int reps, uint32_t* dump) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = tmem_slot;
    const uint32_t t_lane = base + ((uint32_t)((warp & 3) << 5) << 16);
    const uint32_t col0 = (uint32_t)((warp >> 2) * 128) & 511u;
    // known contents: column c of row (lane) holds (row << 16) | c
    if (warp < 4) {
        for (int c0 = 0; c0 < 512; c0 += 32) {
            uint32_t v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((uint32_t)((warp & 3) * 32 + lane) << 16) | (uint32_t)(c0 + j);
            tmem_st32(t_lane + c0, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (MODE == 1) {
            uint32_t a[32], b[32], c[32], d[32];
            tmem_ld32(t_lane + col0, a); tmem_ld32(t_lane + col0 + 32, b); tmem_ld32(t_lane + col0 + 64, c); tmem_ld32(t_lane + col0 + 96, d);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= a[j] ^ b[j] ^ c[j] ^ d[j];
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t a[32];
                if (MODE == 0) tmem_ld32(t_lane + col0 + q * 32, a);
                else tmem_ld32_pack16(t_lane + ((col0 + q * 64) & 511u), a);          // 64 columns per load
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= a[j];
                if (dump && r == 0 && q == 0 && blockIdx.x == 0 && warp == 0 && lane == 1)
                    for (int j = 0; j < 32; ++j) dump[j] = a[j];
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (lane == 0) cyc[blockIdx.x * 16 + warp] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(base, 512);
}

template <int MODE>
static void run(const char* name, int warps, int reps, uint32_t* out, long long* cyc, uint32_t* dump) {
    cudaMemset(cyc, 0, 148 * 16 * sizeof(long long));
    k<MODE><<<148, warps * 32>>>(out, cyc, reps, dump);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    static long long h[148 * 16];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148 * 16; ++i) mx = h[i] > mx ? h[i] : mx;
    const double cols = (MODE == 2 ? 256.0 : 128.0);      // TMEM columns read per warp per repetition
    const double bytes_tmem = (double)warps * reps * cols * 32 * 4;
    const double regs_bytes = (double)warps * reps * 128.0 * 32 * 4;
    printf("%-34s warps %2d: %7.1f cyc per 128-register repetition per warp-set, %6.1f B/clk/SM of TMEM columns, %6.1f B/clk/SM into registers\n",
           name, warps, (double)mx / reps, bytes_tmem / mx, regs_bytes / mx);
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    uint32_t *out, *dump;
    long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4);
    cudaMalloc(&cyc, 148 * 16 * sizeof(long long));
    cudaMalloc(&dump, 32 * 4);
    const int reps = 2000;
    for (int w : {4, 8, 16}) run<0>("32x32b.x32, wait after each load", w, reps, out, cyc, nullptr);
    for (int w : {4, 8, 16}) run<1>("32x32b.x32, 4 loads then wait", w, reps, out, cyc, nullptr);
    for (int w : {4, 8, 16}) run<2>("32x32b.x32.pack::16b (64 cols/load)", w, reps, out, cyc, w == 4 ? dump : nullptr);
    uint32_t h[32];
    cudaMemcpy(h, dump, sizeof(h), cudaMemcpyDeviceToHost);
    printf("pack::16b, row 1, first load: ");
    for (int j = 0; j < 8; ++j) printf("%08x ", h[j]);
    printf("  (TMEM column c of row r was written as (r << 16) | c)\n");
    return 0;
}
