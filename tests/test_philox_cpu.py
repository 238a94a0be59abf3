"""CPU: the host restatement of the device generator reproduces the published Philox4x32-10
known-answer vectors (Random123 `kat_vectors`), so comparing the device stream with it
(tests/test_gpu_device_rng.py) pins the device generator to the published algorithm."""
import numpy as np

import philox_ref as pr


def _one(ctr, key):
    r = pr.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1])
    return tuple(int(x[0]) for x in r)


def test_philox4x32_10_known_answers():
    assert _one((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert _one((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert _one((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_u52_is_open_interval_and_exact():
    lo = pr.u52(np.array([0]), np.array([0]))[0]
    hi = pr.u52(np.array([0xffffffff]), np.array([0xffffffff]))[0]
    assert 0.0 < lo < 1e-15 and 1.0 - 1e-15 < hi < 1.0
    x = pr.draw_unit_cube(2000, 5, 12345, 7)
    assert x.shape == (2000, 5) and (x > 0).all() and (x < 1).all()
    assert abs(x.mean() - 0.5) < 0.02
    # counter-based: a call is a window into one stream
    y = pr.draw_unit_cube(1000, 5, 12345, 1007)
    assert (x[1000:] == y).all()
