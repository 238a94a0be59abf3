"""CPU model of the fp32 membership pre-filter (ultranest_b200/csrc/unb_scan.cu: k_live_build32,
the refill thresholds of k_inside_any32, tile_filter32) in NumPy float32, checked against the
reference's exact decision `D_ref <= r2` (k-sequential, non-fused fp64) on adversarial pairs:

  * every reference hit is FLAGGED            (acc32 >= thr_lo)   -- no neighbour is ever lost;
  * every CERTAIN neighbour is a reference hit (acc32 >= thr_hi)  -- retiring a slot without the
    exact evaluation never invents a neighbour;

for the scales and offsets at which the host still selects the fp32 filter
(kappa32 (2 |a|^2_max + r2) <= r2 / 16, unb_live_prepare32).  This is the error-budget argument of
DESIGN.md 4.1 put to a numerical test; the GPU tier checks the kernels themselves."""
import numpy as np
import pytest

U32 = 2.0**-24


def kappa32(d):
    return (4.0 * d + 32.0) * U32


def _round_dir(x, up):
    """float64 -> float32 with directed rounding (the kernel's __double2float_ru / _rd)."""
    f = x.astype(np.float32)
    back = f.astype(np.float64)
    if up:
        bump = back < x
        f[bump] = np.nextafter(f[bump], np.float32(np.inf))
    else:
        bump = back > x
        f[bump] = np.nextafter(f[bump], np.float32(-np.inf))
    return f


def _norms(rows):
    """fma chain of the kernels (k ascending); fp64 fma of fp64 inputs is emulated by the plain
    product-sum, whose difference is far below the slack being tested."""
    nb = np.zeros(len(rows))
    for k in range(rows.shape[1]):
        nb = rows[:, k] * rows[:, k] + nb
    return nb


def _exact_d(a, b):
    """The reference's distance (mlfriends.pyx:178-180): d = d + (a-b)*(a-b), k sequential."""
    D = np.zeros(len(a))
    for k in range(a.shape[1]):
        diff = a[:, k] - b[:, k]
        D = D + diff * diff
    return D


def _filter_model(a, b, r2, namax):
    """acc32, thr_lo, thr_hi for the pairs (a[i], b[i])."""
    d = a.shape[1]
    k32 = kappa32(d)
    na, nb = _norms(a), _norms(b)
    h = _round_dir(0.5 * (r2 * (1.0 + k32) - na * (1.0 - k32)), up=True)
    thr_lo = _round_dir(nb * (0.5 * (1.0 - k32)), up=False)
    thr_hi = _round_dir(thr_lo.astype(np.float64) + (1.001 * k32) * ((namax + r2) + nb), up=True)
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    acc = h.copy()
    for k in range(d):
        # fmaf: the product of two floats is exact in float64, one rounding to float32 at the end
        acc = (a32[:, k].astype(np.float64) * b32[:, k].astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
    return acc, thr_lo, thr_hi


CASES = [
    # (d, scale of the cloud, offset of the cloud from the origin, r2 relative to scale^2)
    (20, 1.0, 0.0, 0.8),
    (20, 1.0, 0.0, 1e-3),
    (20, 1e-4, 0.0, 0.5),
    (20, 1e4, 0.0, 0.5),
    (5, 1.0, 3.0, 0.3),        # norms dominated by the offset
    (2, 1.0, 10.0, 0.5),
    (32, 1.0, 1.0, 0.9),
    (100, 1.0, 0.0, 0.7),
    (100, 1.0, 0.5, 0.2),
]


@pytest.mark.parametrize("d,scale,offset,r2rel", CASES)
def test_fp32_filter_never_loses_and_never_invents_a_neighbour(d, scale, offset, r2rel):
    rng = np.random.RandomState(d * 1000 + int(offset * 10))
    n = 40000
    a = (rng.normal(size=(n, d)) / np.sqrt(d) + offset) * scale
    r2 = r2rel * scale * scale
    namax = float(_norms(a).max())
    k32 = kappa32(d)
    if not (k32 * (2.0 * namax + r2) <= r2 / 16.0):
        pytest.skip("the host would select the fp64 filter here (unb_live_prepare32)")
    # partners on a shell around the radius: relative distance from the boundary from 1e-12 to 0.3
    direction = rng.normal(size=(n, d))
    direction /= np.sqrt((direction**2).sum(axis=1, keepdims=True))
    eps = 10.0**rng.uniform(-12, -0.5, size=n) * rng.choice([-1.0, 1.0], size=n)
    eps[: n // 20] = 0.0
    b = a + direction * np.sqrt(r2 * (1.0 + eps)).reshape((-1, 1))
    hit = _exact_d(a, b) <= r2
    acc, thr_lo, thr_hi = _filter_model(a, b, r2, namax)
    flagged = ~(acc < thr_lo)
    sure = acc >= thr_hi
    assert 0.2 < hit.mean() < 0.8
    assert flagged[hit].all(), "a reference neighbour was not flagged"
    assert hit[sure].all(), "a certain neighbour is no reference neighbour"
    # the uncertain shell: (r2 - D)/2 < M + errors, i.e. D within `shell` * r2 of the radius, with
    # M = 1.001 kappa32 (|a|^2_max + r2 + |b|^2) and |b|^2 <= 2 (|a|^2_max + r2); deeper hits are certain
    shell = 8.0 * k32 * 3.0 * (namax + r2) / r2
    assert shell < 1.0
    deep = eps < -shell
    if (deep & hit).any():
        assert sure[deep & hit].all()
    # and generic (volume-uniform) neighbours are almost always certain
    rad = rng.uniform(size=n)**(1.0 / d)
    b2 = a + direction * (np.sqrt(r2) * rad).reshape((-1, 1))
    hit2 = _exact_d(a, b2) <= r2
    acc2, lo2, hi2 = _filter_model(a, b2, r2, namax)
    assert (~(acc2 < lo2))[hit2].all()
    assert hit2[acc2 >= hi2].all()
    # a volume-uniform neighbour falls into the shell with probability 1 - (1 - shell)^(d/2)
    expected_uncertain = 1.0 - (1.0 - shell)**(d / 2.0)
    assert (acc2 >= hi2)[hit2].mean() >= 1.0 - expected_uncertain - 0.002
    if d == 20 and offset == 0.0 and r2rel == 0.8:   # the headline geometry: practically all certain
        assert (acc2 >= hi2)[hit2].mean() > 0.995


def test_far_pairs_are_not_flagged():
    """The false-alarm shell is as thin as the usability condition promises (<= r2/8 in D)."""
    rng = np.random.RandomState(3)
    d, n = 20, 20000
    a = rng.normal(size=(n, d)) / np.sqrt(d)
    r2 = 0.5
    namax = float(_norms(a).max())
    direction = rng.normal(size=(n, d))
    direction /= np.sqrt((direction**2).sum(axis=1, keepdims=True))
    b = a + direction * np.sqrt(r2 * 1.13)
    acc, thr_lo, _ = _filter_model(a, b, r2, namax)
    assert not (~(acc < thr_lo)).any()


def _hi32(x):
    return (x.view(np.int64) >> 32).astype(np.int64)


@pytest.mark.parametrize("d,scale,offset,r2rel", [(20, 1.0, 0.0, 0.8), (20, 1.0, 5.0, 1e-6), (5, 1e-3, 0.0, 0.3),
                                                 (32, 1e3, 1.0, 1e-4), (2, 1.0, 100.0, 1e-8)])
def test_fp64_filter_never_loses_a_neighbour(d, scale, offset, r2rel):
    """The fp64 filter of the scan kernels (k_live_set_h, cs_init, cs_flag: kappa = (8d+64) 2^-53,
    flagged when hi32(acc) >= hi32(thr) - 1) keeps every reference hit, down to radii ten orders of
    magnitude below the norms (where the fp32 filter is not used)."""
    rng = np.random.RandomState(d + int(offset))
    n = 40000
    kappa = (8.0 * d + 64.0) * 2.0**-53
    a = (rng.normal(size=(n, d)) / np.sqrt(d) + offset) * scale
    r2 = r2rel * scale * scale
    direction = rng.normal(size=(n, d))
    direction /= np.sqrt((direction**2).sum(axis=1, keepdims=True))
    eps = 10.0**rng.uniform(-16, -1, size=n) * rng.choice([-1.0, 1.0], size=n)
    eps[: n // 10] = 0.0
    b = a + direction * np.sqrt(r2 * (1.0 + eps)).reshape((-1, 1))
    hit = _exact_d(a, b) <= r2
    na, nb = _norms(a), _norms(b)
    h = 0.5 * (r2 * (1.0 + kappa) - na * (1.0 - kappa))
    thr = 0.5 * (nb * (1.0 - kappa))
    acc = h.copy()
    for k in range(d):
        acc = a[:, k] * b[:, k] + acc
    flagged = _hi32(acc) >= _hi32(thr) - 1
    assert 0.1 < hit.mean() < 0.9
    assert flagged[hit].all()


@pytest.mark.parametrize("d", [2, 5, 20, 32, 50, 100, 143])
@pytest.mark.parametrize("cond", [1e0, 1e6, 1e12])
def test_ellipsoid_filter_band_covers_the_einsum_value(d, cond):
    """Prep kernels (k_prep_reg, k_prep_tile): the fused filter value on the FOLDED matrix
    (S_jj = A_jj, S_jk = A_jk + A_kj above the diagonal) and the reference's einsum value
    (acc += (d_j A_jk) d_k, mlfriends.pyx:910) differ by less than the band half-width
    tol = 2 (d^2 + 2d + 8) u |A|_F |delta|^2, so deciding rows outside the band by the filter and
    rows inside it by the exact order gives the reference's mask -- also for ill-conditioned A."""
    from oracle import cport
    rng = np.random.RandomState(d + int(np.log10(cond)))
    n = 3000
    q, _ = np.linalg.qr(rng.normal(size=(d, d)))
    eig = np.logspace(0, np.log10(cond), d)
    cov = (q * eig) @ q.T
    A = np.linalg.inv(cov)                      # not exactly symmetric, like the reference's invcov
    ctr = rng.uniform(0.4, 0.6, size=d)
    pts = ctr + rng.normal(size=(n, d)) * np.sqrt(eig.mean())
    _, r_ref = cport.inside_ellipsoid(pts, ctr, A, 1.0, return_r=True)
    delta = pts - ctr
    S = np.triu(A, 1) + np.triu(A.T, 1) + np.diag(np.diag(A))
    r_fast = np.einsum('ij,ij->i', delta @ S.T, delta)     # sum_j d_j sum_{k>=j} S_jk d_k
    tol = 2.0 * (d * d + 2 * d + 8) * 2.0**-53 * np.sqrt((A * A).sum()) * (delta**2).sum(axis=1)
    assert (np.abs(r_fast - r_ref) <= 0.5 * tol).all()
    # the band is a sliver of the value range unless A is nearly singular
    if cond <= 1e6:
        assert np.median(tol / np.abs(r_ref)) < 1e-6
