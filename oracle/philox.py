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


# --------------------------------------------------------------------------------------------------
# numpy restatement of the PRODUCTION variate transforms of the fused step kernels
# (tempest_b200/csrc/tb_mcmc_shared.cuh: bm_pair32, gamma_mt, normals_fixed, accept_uniform).
# The device uses MUFU log / sin / cos (abs. error ~2^-21), so equality is to ~1e-6, not bitwise.
RNG_GAMMA, RNG_NORMAL, RNG_ACCEPT = 1, 2, 3


def _block(seed, iteration, slots, step, word3):
    slots = np.asarray(slots, dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, ((seed >> 32) ^ iteration) & 0xFFFFFFFF
    return philox4x32(slots & MASK, slots >> np.uint64(32), np.full(slots.shape, step, dtype=np.uint64),
                      np.full(slots.shape, word3, dtype=np.uint64), k0, k1)


def bm_pair32(a, b):
    """Box-Muller on two 32-bit words with fp32 radius / angle, promoted to fp64 (tb::bm_pair32)."""
    a = np.asarray(a, dtype=np.uint64).astype(np.float32)          # (float)a: round to nearest
    u1 = (a + np.float32(0.5)) * np.float32(2.3283064365386963e-10)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    ang = np.asarray(b, dtype=np.uint64).astype(np.uint32).view(np.int32).astype(np.float32) \
        * np.float32(1.4629180792671596e-9)
    return (r * np.cos(ang)).astype(np.float64), (r * np.sin(ang)).astype(np.float64)


def step_normals(seed, iteration, slots, step, attempt, d):
    """z[n, d] of redraw `attempt` (tb::normals_fixed): ceil(d/4) blocks, four normals each."""
    ncall = (d + 3) // 4
    out = np.empty((len(slots), 4 * ncall))
    for c in range(ncall):
        r0, r1, r2, r3 = _block(seed, iteration, slots, step, (RNG_NORMAL << 24) | ((attempt * ncall + c) & 0xFFFFFF))
        out[:, 4 * c], out[:, 4 * c + 1] = bm_pair32(r0, r1)
        out[:, 4 * c + 2], out[:, 4 * c + 3] = bm_pair32(r2, r3)
    return out[:, :d]


def step_gamma(seed, iteration, slots, step, shape):
    """(standard-gamma variate, accept uniform, margin) per walker (tb::gamma_mt + tb::accept_uniform).
    `margin` is the distance of the deciding comparison from its threshold (tiny margins may flip on the device)."""
    slots = np.asarray(slots, dtype=np.uint64)
    n = len(slots)
    dd = shape - 1.0 / 3.0
    cc = 1.0 / np.sqrt(9.0 * dd)
    g = np.full(n, dd)
    acc = np.zeros(n)
    margin = np.full(n, np.inf)
    todo = np.ones(n, dtype=bool)
    for trial in range(64):
        if not todo.any():
            break
        idx = np.nonzero(todo)[0]
        r0, r1, r2, r3 = _block(seed, iteration, slots[idx], step, (RNG_GAMMA << 24) | trial)
        n0, _ = bm_pair32(r0, r1)
        t = cc * n0
        v1 = 1.0 + t
        v = v1 * v1 * v1
        uu = (r2.astype(np.float64) + 0.5) * 2.3283064365386963e-10
        x2 = n0 * n0
        squeeze = 1.0 - 0.0331 * x2 * x2
        ok = uu < squeeze
        m = np.abs(uu - squeeze)
        small = np.abs(t) < 0.015625
        s = 1.0 / 8.0 - t * (1.0 / 9.0)
        for c in (7.0, 6.0, 5.0, 4.0):
            s = 1.0 / c - t * s
        R = -3.0 * dd * (t * t) ** 2 * s
        thr = 1.0 + R * (1.0 + R * (0.5 + R * (1.0 / 6.0)))
        with np.errstate(invalid="ignore", divide="ignore"):
            exact = np.where(small, uu < thr, np.log(uu) < 0.5 * x2 + dd * (1.0 - v + np.log(np.where(v > 0, v, 1.0))))
            m2 = np.where(small, np.abs(uu - thr), np.abs(np.log(uu) - (0.5 * x2 + dd * (1.0 - v + np.log(np.where(v > 0, v, 1.0))))))
        ok = (ok | exact) & (v1 > 0.0)
        margin[idx] = np.minimum(margin[idx], np.where(ok, np.where(uu < squeeze, m, m2), np.minimum(m, m2)))
        g[idx[ok]] = dd * v[ok]
        acc[idx[ok]] = (r3[ok].astype(np.float64) + 0.5) * 2.3283064365386963e-10
        todo[idx[ok]] = False
    return g, acc, margin
