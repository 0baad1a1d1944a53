// Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) for the
// in-kernel draws of the stochastic branches of render_rays: the stratified jitter of audio_exp_nerf.py:321-328 (torch.rand(N, 64)) and
// the inverse-CDF draws of helper.py:282-283 (torch.rand(N, 128)).  The reference draws from torch's global generator; parity for
// those branches is defined on SUPPLIED draws (pytest=True), so the in-kernel stream only has to be U[0,1), reproducible from
// (seed, offset) and free of HBM traffic.  oracle/philox_ref.py restates the same function in numpy for the bit-exact tests.
#pragma once
#include <stdint.h>

namespace inerf {

struct Philox4 { uint32_t x, y, z, w; };

// counter = (c0, c1, c2, c3), key = (k0, k1); ten rounds
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// The RNG state the entry points take: two uint64 in DEVICE memory, {seed, offset}.  Reading it on the device (instead of passing the
// numbers by value) lets a captured CUDA graph draw fresh numbers on every replay: inerf_rng_advance bumps the offset inside the graph.
// Counter layout: (index lo, index hi, offset lo + stream id, offset hi); `stream_id` separates the draws of different call sites
// (coarse jitter, coarse-pass importance draws, ...) inside one step.
__device__ __forceinline__ Philox4 philox_at(const unsigned long long* __restrict__ state, uint64_t index, uint32_t stream_id) {
    const unsigned long long seed = state[0], off = state[1] + ((unsigned long long)stream_id << 56);
    return philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), (uint32_t)off, (uint32_t)(off >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
}

// 24 random bits -> [0, 1) like torch.rand's float32 path
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }
// (0, 1]: safe argument of a logarithm
__device__ __forceinline__ float u01_open0(uint32_t r) { return (float)((r >> 8) + 1u) * 5.9604644775390625e-08f; }

}  // namespace inerf
