"""CPU tier: the host-side pieces of the step-sampler mirror (direction proposals, unit-cube line
intersection, move diagnostics) consume the RNG and round exactly like the reference's
(ultranest/stepfuncs.pyx:348-533, popstepsampler.py:26-94).  Needs oracle/_ref."""
import types

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def ref():
    oracle.reference()
    import ultranest.popstepsampler as rp
    import ultranest.stepfuncs as rs
    return rs, rp


def _region(seed, n, d):
    rng = np.random.RandomState(seed)
    u = rng.uniform(0.2, 0.8, size=(n, d))
    A = rng.normal(size=(d, d))
    layer = types.SimpleNamespace(axes=A / np.linalg.norm(A, axis=1, keepdims=True),
                                  transform=lambda x: np.dot(x - 0.5, A))
    return types.SimpleNamespace(u=u, transformLayer=layer, maxradiussq=0.3)


GENERATORS = ["generate_cube_oriented_direction", "generate_cube_oriented_direction_scaled",
              "generate_random_direction", "generate_region_oriented_direction",
              "generate_region_random_direction", "generate_differential_direction",
              "generate_mixture_random_direction"]


@pytest.mark.parametrize("name", GENERATORS)
@pytest.mark.parametrize("scale", [1, 0.3, 1.7])
def test_direction_generators_bitexact_and_rng_aligned(ref, name, scale):
    from ultranest_b200 import stepfuncs as ours
    rs, _ = ref
    region = _region(3, 50, 7)
    ui = region.u[:33]
    np.random.seed(12)
    want = getattr(rs, name)(ui, region, scale=scale)
    tail_want = np.random.uniform()
    np.random.seed(12)
    got = getattr(ours, name)(ui, region, scale=scale)
    tail_got = np.random.uniform()
    np.testing.assert_array_equal(got, want)
    assert tail_got == tail_want   # same number of draws consumed


def test_line_intersection_and_diagnostics(ref):
    from ultranest_b200 import popstepsampler as ours
    _, rp = ref
    rng = np.random.RandomState(4)
    origin = rng.uniform(size=(200, 6))
    direction = rng.normal(size=(200, 6))
    direction[::7, 2] = 0.0    # axis-parallel slabs give NaN/inf intermediates
    a = ours.unitcube_line_intersection(origin, direction)
    b = rp.unitcube_line_intersection(origin, direction)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    region = _region(5, 40, 6)
    fa, (ma, ra) = ours.diagnose_move_distances(region, origin, origin + 0.01 * direction)
    fb, (mb, rb) = rp.diagnose_move_distances(region, origin, origin + 0.01 * direction)
    np.testing.assert_array_equal(fa, fb)
    np.testing.assert_array_equal(ma, mb)
    assert ra == rb
