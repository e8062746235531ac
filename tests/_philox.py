"""numpy Philox4x32-10 (Salmon et al. 2011), the host twin of orca::philox_uniform."""
import numpy as np


def philox_uniform(seed, c0, c1):
    c0 = np.asarray(c0, np.uint64)
    c1 = np.broadcast_to(np.asarray(c1, np.uint64), c0.shape)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    x0, x1, x2, x3 = c0 & mask, c1 & mask, np.zeros_like(c0), np.zeros_like(c0)
    for _ in range(10):
        p0 = M0 * x0
        p1 = M1 * x2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        x0, x1, x2, x3 = (hi1 ^ x1 ^ k0) & mask, lo1, (hi0 ^ x3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return ((x0 >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)
