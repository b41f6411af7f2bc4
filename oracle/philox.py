"""numpy replica of the counter-based RNG the kernels use (`csrc/common.cuh`: Philox4x32-10 keyed by
(seed, salt, counter)) and of the noise derived from it.  TEST INFRASTRUCTURE ONLY (see dvae_oracle.py):
lets the tests replay dropout masks and Gumbel-max sampling noise on the host.

Philox4x32-10: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11); constants as published.
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32(seed, salt, ctr):
    """ctr: uint64 array [n] -> uint32 array [n,4]."""
    ctr = np.asarray(ctr, dtype=np.uint64)
    c = [ctr & _MASK, ctr >> np.uint64(32), np.full_like(ctr, np.uint64(salt)), np.full_like(ctr, np.uint64(0x5EED))]
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & _MASK, p1 >> np.uint64(32), p1 & _MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return np.stack(c, axis=1).astype(np.uint32)


def dropout_mask(seed, salt, rows, width, p):
    """Keep-scale mask [rows,width] (0 or 1/(1-p)) exactly as `dropout_scale4` produces it."""
    w4 = (width + 3) // 4
    r = philox4x32(seed, salt, np.arange(rows * w4, dtype=np.uint64)).reshape(rows, w4 * 4)[:, :width]
    u = r.astype(np.float32) * np.float32(2.3283064365386963e-10)
    return np.where(u >= np.float32(p), np.float32(1.0 / (1.0 - p)), np.float32(0.0))


def gumbel_noise(seed, salt, rows, V):
    """Gumbel(0,1) noise [rows,V] as `gumbel4` produces it (up to libm-vs-CUDA logf rounding)."""
    v4 = (V + 3) // 4
    r = philox4x32(seed, salt, np.arange(rows * v4, dtype=np.uint64)).reshape(rows, v4 * 4)[:, :V]
    u = ((r >> np.uint32(8)).astype(np.float64) + 0.5) / 16777216.0
    return -np.log(-np.log(u))
