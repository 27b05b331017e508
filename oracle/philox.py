"""Oracle (test infrastructure): Philox-4x32-10 + Box-Muller in numpy.

The training graph of the reference draws its codes with tf.random_normal (DeepLearning/my_sngan.py:122-124); TensorFlow's
generator is Philox-4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11 -- the
published algorithm restated here; TensorFlow itself is a third-party dependency absent from /root/reference).  TF's exact
stream (its counter layout and its own Box-Muller arithmetic) is NOT reproduced: the reference's codes are unseeded, so only
the distribution is part of the contract.  This restatement pins the CUDA kernel (csrc/elementwise.cu sample_normal_kernel):
the four Philox words bit for bit -- against the algorithm's published known-answer vectors too -- and the normals to fp32
rounding.  Only tests/ may import this module.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: uint32 [..., 4]; key: (k0, k1) -> uint32 [..., 4]."""
    c = np.asarray(counter, dtype=np.uint64).copy()
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[..., 0]
        p1 = M1 * c[..., 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c[..., 1] ^ np.uint64(k0)
        n2 = hi0 ^ c[..., 3] ^ np.uint64(k1)
        c = np.stack([n0, lo1, n2, lo0], axis=-1) & MASK
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def sample_words(n, seed, draw=0):
    """The kernel's counter layout: quadruple q -> counter (q lo, q hi, draw lo, draw hi), key = seed (lo, hi)."""
    nq = (n + 3) // 4
    q = np.arange(nq, dtype=np.uint64)
    ctr = np.stack([q & MASK, q >> np.uint64(32), np.full(nq, draw & 0xFFFFFFFF, np.uint64), np.full(nq, (draw >> 32) & 0xFFFFFFFF, np.uint64)], -1)
    return philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).reshape(-1)[:n]


def sample_normal(n, seed, draw=0):
    """N(0, 1) samples as the kernel forms them: u = ((x >> 8) + 1) 2^-24, z = sqrt(-2 ln u0) (cos, sin)(2 pi u1)."""
    w = sample_words((n + 3) // 4 * 4, seed, draw).reshape(-1, 4).astype(np.float64)
    u = (np.floor(w / 256.0) + 1.0) / 16777216.0
    out = np.empty_like(u)
    for h in range(2):
        r = np.sqrt(-2.0 * np.log(u[:, 2 * h]))
        out[:, 2 * h] = r * np.cos(2.0 * np.pi * u[:, 2 * h + 1])
        out[:, 2 * h + 1] = r * np.sin(2.0 * np.pi * u[:, 2 * h + 1])
    return out.reshape(-1)[:n]
