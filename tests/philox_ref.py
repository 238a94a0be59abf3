"""Host restatement (NumPy) of the device proposal generator of ultranest_b200/csrc/unb_sample.cu:
Philox4x32-10, the 52-bit open-interval uniform, and the two draw recipes (mlfriends.pyx:1105, 1145-1151).
Test infrastructure: the GPU tests compare the device stream with this, word for word."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xffffffff)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over uint64 arrays holding 32-bit words; returns four uint64 arrays."""
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) & MASK for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xffffffff, int(k1) & 0xffffffff
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0 = (k0 + W0) & 0xffffffff
        k1 = (k1 + W1) & 0xffffffff
    return c0, c1, c2, c3


def u52(a, b):
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    return ((a >> np.uint64(6)).astype(np.float64) * 67108864.0
            + (b >> np.uint64(6)).astype(np.float64) + 0.5) * (1.0 / 4503599627370496.0)


def _ctr(offset, m):
    g = np.uint64(offset) + np.arange(m, dtype=np.uint64)
    return g & MASK, g >> np.uint64(32)


def draw_unit_cube(m, d, seed, offset):
    g0, g1 = _ctr(offset, m)
    out = np.empty((m, d))
    for k in range(0, d, 2):
        r = philox4x32_10(g0, g1, np.uint64(k >> 1), np.uint64(1), seed & 0xffffffff, seed >> 32)
        out[:, k] = u52(r[0], r[1])
        if k + 1 < d:
            out[:, k + 1] = u52(r[2], r[3])
    return out


def draw_wrapping_ellipsoid(m, d, seed, offset, center, axes_T, enlarge):
    g0, g1 = _ctr(offset, m)
    z = np.empty((m, d))
    for k in range(0, d, 2):
        r = philox4x32_10(g0, g1, np.uint64(k >> 1), np.uint64(0), seed & 0xffffffff, seed >> 32)
        rad = np.sqrt(-2.0 * np.log(u52(r[0], r[1])))
        ang = 2.0 * np.pi * u52(r[2], r[3])
        z[:, k] = rad * np.cos(ang)
        if k + 1 < d:
            z[:, k + 1] = rad * np.sin(ang)
    r = philox4x32_10(g0, g1, np.uint64(0x7fffffff), np.uint64(0), seed & 0xffffffff, seed >> 32)
    f = enlarge**0.5 * u52(r[0], r[1])**(1.0 / d) / np.sqrt((z**2).sum(axis=1))
    return center + np.dot(z * f.reshape((-1, 1)), axes_T)
