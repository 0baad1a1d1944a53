"""Philox4x32-10 in numpy -- TEST INFRASTRUCTURE ONLY (checker of ideal-nerf_b200/csrc/philox.cuh).

The reference draws its stochastic branches with torch.rand from the global generator
(NeRFs/HeadNeRF/train/audio_exp_nerf.py:321-328, NeRFs/HeadNeRF/helper.py:282-283); the CUDA path draws
them in-kernel with Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11 -- the Random123 library's
generator; also what cuRAND / torch.cuda use).  This restates the published algorithm; it is pinned by
the Random123 known-answer vectors in tests/test_oracle_golden.py.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays of counters; keys may be scalars.  Returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32).copy() for x in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def draws_u01(seed, offset, stream_id, n):
    """The first n draws of stream `stream_id` as the kernels number them: draw i = word (i & 3) of block (i >> 2), counter
    (block lo, block hi, off lo, off hi) with off = offset + (stream_id << 56), key = seed; float = (word >> 8) * 2^-24."""
    blocks = np.arange((n + 3) // 4, dtype=np.uint64)
    off = (int(offset) + (int(stream_id) << 56)) & 0xFFFFFFFFFFFFFFFF
    w = philox4x32_10((blocks & MASK).astype(np.uint32), (blocks >> np.uint64(32)).astype(np.uint32),
                      np.uint32(off & 0xFFFFFFFF), np.uint32(off >> 32), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(w, -1).reshape(-1)[:n]
    return (words >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
