"""Philox4x32-10 in numpy -- TEST INFRASTRUCTURE ONLY (checks the in-kernel generator of
tempest_b200/csrc/tb_like.cuh against the Random123 known-answer vectors and lets tests
reproduce the device's uniform draws on the CPU)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & MASK for v in (c0, c1, c2, c3))
    a, b = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(a), lo1, hi0 ^ c3 ^ np.uint64(b), lo0
        a, b = (a + W0) & 0xFFFFFFFF, (b + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def uniforms(seed, iteration, purpose, slots, step=0, sub=0, open_interval=False):
    """First uniform of tb::philox_u2 for each walker slot (tb_like.cuh)."""
    slots = np.asarray(slots, dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, ((seed >> 32) ^ iteration) & 0xFFFFFFFF
    r0, r1, _, _ = philox4x32(slots & MASK, slots >> np.uint64(32), np.uint64(step),
                              np.uint64((purpose << 24) | (sub & 0xFFFFFF)), k0, k1)
    bits = ((r1 << np.uint64(32)) | r0) >> np.uint64(11)
    v = bits.astype(np.float64)
    if open_interval:
        v = v + 0.5
    return v * (1.0 / 9007199254740992.0)
