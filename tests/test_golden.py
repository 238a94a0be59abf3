"""Golden vectors generated from the unmodified reference (tests/golden/make_golden.py):
the reference's own fixtures (eggboxregion.txt, clusters2.txt) and a 20-D region.

CPU half: pins the oracle to the golden vectors (no oracle/_ref needed).
GPU half: the CUDA path, through the C ABI, against the same vectors."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import cport

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


def load(name):
    return np.load(os.path.join(GOLD, name))


@pytest.fixture(scope="module")
def region20():
    g = load("ref_region_d20.npz")
    n, d, m = 1000, 20, 20000
    u, _ = mg.correlated_live(n, d, 1)
    assert mg.sha(u) == str(g["sha_u"]), "live set did not regenerate bit for bit"
    assert (u == g["u"]).all()
    crng = np.random.RandomState(3)
    z = mg.exact_ball(crng, m, d)
    cand_a = g["ell_center"] + mg.exact_matmul(z * float(g["enlarge"])**0.5,
                                              np.ascontiguousarray(g["ell_axes_T"]))
    lo, hi = u.min(axis=0), u.max(axis=0)
    cand_b = u[crng.randint(n, size=m)] + crng.normal(size=(m, d)) * 0.008
    tb = -1.2 + crng.uniform(size=(m, d)) * 2.4
    assert mg.sha(cand_a) == str(g["sha_cand_a"])
    assert mg.sha(cand_b) == str(g["sha_cand_b"])
    assert mg.sha(tb) == str(g["sha_tb"])
    return dict(g=g, u=u, cand_a=cand_a, cand_b=cand_b, tb=tb, m=m)


def _bits(packed, m):
    return np.unpackbits(packed)[:m].astype(bool)


# --------------------------------------------------------------------------- CPU: oracle
def test_oracle_eggboxregion_known_answer():
    g = load("ref_eggboxregion.npz")
    u = g["u"]
    assert (cport.transform_scaling(u, g["mean"], g["std"]) == g["unormed"]).all()
    for seed in range(10):
        np.random.seed(seed)
        r = 0
        for _ in range(30):
            r = max(r, cport.maxradiussq_selected(g["unormed"], cport.draw_selection(np.random, len(u))))
        assert r == g["maxr"][seed]
        assert 1e-10 < r < 6e-10   # the reference's own window (test_clustering.py:86-90)
    assert 14 < int(g["nclusters"]) < 20


def test_oracle_clusters2():
    g = load("ref_clusters2.npz")
    assert (cport.subtract_nearby(g["u"], float(g["maxr"])) == g["subtracted"]).all()


def test_oracle_region_d20(region20):
    g = region20["g"]
    u, m = region20["u"], region20["m"]
    rs = np.random.RandomState(2)
    for r in range(12):
        sel = cport.draw_selection(rs, len(u))
        assert cport.maxradiussq_selected(g["unormed"], sel) == g["maxd_rounds"][r]
        ctr, cov = cport.bounding_ellipsoid(u[sel])
        f = cport.enlargement_f(u, sel, ctr, np.linalg.inv(cov))
        assert abs(f - g["f_rounds"][r]) <= 1e-12 * f   # LAPACK inv is host specific
    for name in ("a", "b"):
        cand = region20["cand_" + name]
        ell = cport.inside_ellipsoid(cand, g["ell_center"], g["ell_invcov"], float(g["enlarge"]))
        assert (ell == _bits(g["ell_" + name], m)).all()
        for xf in (lambda p: np.dot(p - g["ctr"], g["T"]),
                   lambda p: cport.transform_affine(p, g["ctr"], g["T"])):
            mask = cport.region_inside(cand, g["unormed"], xf, float(g["maxradiussq"]),
                                       g["ell_center"], g["ell_invcov"], float(g["enlarge"]))
            assert (mask == _bits(g["mask_" + name], m)).all()
        assert 0 < _bits(g["mask_" + name], m).sum()
    idx = cport.find_nearby(g["unormed"], region20["tb"], float(g["maxradiussq"]) * 1.5)
    assert (idx == g["idx_tb"].astype(np.int64)).all()
    assert (cport.subtract_nearby(u, 0.02)[:50] == g["subtracted_head"]).all()
    ids = (np.arange(len(u)) % 3).astype(np.int64)
    assert cport.mean_pair_distance(g["unormed"], ids) == float(g["mean_pair_distance"])


# --------------------------------------------------------------------------- GPU: product
@pytest.mark.gpu
def test_gpu_eggboxregion_known_answer():
    from ultranest_b200 import mlfriends as m
    g = load("ref_eggboxregion.npz")
    u = g["u"]
    layer = m.ScalingLayer()
    layer.optimize(u, u)
    assert (layer.mean == g["mean"]).all() and (layer.std == g["std"]).all()
    for seed in range(10):
        np.random.seed(seed)
        region = m.MLFriends(u, layer)
        assert (region.unormed == g["unormed"]).all()
        maxr = region.compute_maxradiussq(nbootstraps=30)
        assert maxr == g["maxr"][seed]
        assert 1e-10 < maxr < 6e-10
    nclusters, ids, _ = m.update_clusters(u, u, maxr)
    assert nclusters == int(g["nclusters"])
    assert (ids == g["clusterids"]).all()


@pytest.mark.gpu
def test_gpu_clusters2():
    from ultranest_b200 import mlfriends as m
    g = load("ref_clusters2.npz")
    nclusters, ids, over = m.update_clusters(g["u"], g["u"], float(g["maxr"]))
    assert nclusters == int(g["nclusters"]) and (ids == g["clusterids"]).all()
    assert (over == g["overlapped"]).all()
    assert (m.subtract_nearby(g["u"], float(g["maxr"])) == g["subtracted"]).all()


@pytest.mark.gpu
def test_gpu_region_d20(region20):
    from ultranest_b200 import mlfriends as m
    from ultranest_b200 import _native
    g = region20["g"]
    u, mm = region20["u"], region20["m"]
    layer = m.AffineLayer(ctr=g["ctr"], T=g["T"], invT=g["invT"])
    layer.set_clusterids(npoints=len(u))
    region = m.MLFriends(u, layer)
    # our defined-order transform vs the reference's np.dot: ulp-level, never bitwise (fact 6)
    np.testing.assert_allclose(region.unormed, g["unormed"], rtol=0, atol=1e-13)
    # identical t-space inputs -> identical radius, round by round
    region.unormed = g["unormed"].copy()
    r, f = region.compute_enlargement(nbootstraps=12, rng=np.random.RandomState(2))
    assert r == g["maxd_rounds"].max()
    assert abs(f - g["f_rounds"].max()) <= 1e-12 * f
    eng = _native.get_engine()
    sel = np.zeros((12, len(u)), dtype=bool)
    rs = np.random.RandomState(2)
    for i in range(12):
        sel[i, rs.randint(len(u), size=len(u))] = True
    maxd_r, _ = eng.region_bootstrap(g["unormed"], sel)
    assert (maxd_r == g["maxd_rounds"]).all()
    # membership: bit-exact masks against the reference's own inside() output
    region.maxradiussq = float(g["maxradiussq"])
    region.enlarge = float(g["enlarge"])
    region.ellipsoid_center = g["ell_center"]
    region.ellipsoid_invcov = g["ell_invcov"]
    for name in ("a", "b"):
        cand = region20["cand_" + name]
        assert (region.inside_ellipsoid(cand) == _bits(g["ell_" + name], mm)).all()
        assert (region.inside(cand) == _bits(g["mask_" + name], mm)).all()
        mask, idx = eng.region_inside(cand, want_index=True)
        ell = _bits(g["ell_" + name], mm)
        assert (idx[ell] == g["idx_" + name].astype(np.int64)[ell]).all()
    out = np.empty(mm, dtype=np.int64)
    m.find_nearby(g["unormed"], region20["tb"], float(g["maxradiussq"]) * 1.5, out)
    assert (out == g["idx_tb"].astype(np.int64)).all()
    assert (m.subtract_nearby(u, 0.02)[:50] == g["subtracted_head"]).all()
    ids = (np.arange(len(u)) % 3).astype(np.int64)
    np.testing.assert_allclose(m.compute_mean_pair_distance(g["unormed"], ids),
                               float(g["mean_pair_distance"]), rtol=1e-12)
