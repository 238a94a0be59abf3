#!/usr/bin/env python
"""Flip-count study: `region.inside()` of the CUDA path against the REFERENCE's own
`MLFriends.inside` / `RobustEllipsoidRegion.inside` (oracle/_ref, i.e. its np.einsum ellipsoid,
its np.dot layer transform and its Cython find_nearby) on >= 10^7 proposals per configuration.

    python tests/flip_study.py [--rows 10485760] [--out profiles/r02_flip_study.jsonl]

The device transforms proposals in a defined order, the reference with OpenBLAS dgemm (SURVEY
fact 6), so a decision could differ only where a pair distance sits within ~1e-16 relative of
the radius.  This counts how often that happens (expected: 0) and, for every flip, how far the
pair distance is from the radius.  Proposal mix per configuration: wrapping-ellipsoid draws
(accepting regime) and draws from the ellipsoid inflated by 30 % (rejections by both stages).
Test infrastructure (imports oracle/): the product never runs this.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle  # noqa: E402

_STATE = {}


def _init():
    try:
        import threadpoolctl
        _STATE["limit"] = threadpoolctl.threadpool_limits(1)
    except Exception:  # noqa: BLE001
        pass


def _ref_inside(i):
    return _STATE["region"].inside(_STATE["chunks"][i])


def draw(region, m, seed, inflate):
    """mlfriends.pyx:1145-1154 with the enlargement scaled by `inflate`, cut to the unit cube."""
    rng = np.random.RandomState(seed)
    d = region.u.shape[1]
    out = np.empty((m, d))
    filled = 0
    while filled < m:
        ns = int((m - filled) * 1.2) + 16
        z = rng.normal(size=(ns, d))
        z /= ((z**2).sum(axis=1)**0.5).reshape((ns, 1))
        uu = z * (region.enlarge * inflate)**0.5 * rng.uniform(size=(ns, 1))**(1. / d)
        w = region.ellipsoid_center + np.dot(uu, region.ellipsoid_axes_T)
        w = w[np.logical_and(w > 0, w < 1).all(axis=1)]
        take = min(len(w), m - filled)
        out[filled:filled + take] = w[:take]
        filled += take
    return out


def build(mod, cls_name, u, nboot=30):
    layer = mod.AffineLayer()
    layer.optimize(u, u)
    region = getattr(mod, cls_name)(u, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(nbootstraps=nboot, rng=np.random.RandomState(2))
    region.create_ellipsoid()
    return region


def study(name, cls_name, n, d, rows, chunk, cores, scale=0.05):
    from ultranest_b200 import mlfriends as ours
    import ultranest.mlfriends as ref
    u = bench.make_live(n, d, seed=1) if scale == 0.05 else 0.5 + (bench.make_live(n, d, seed=1) - 0.5) * (scale / 0.05)
    reg_o = build(ours, cls_name, u)
    reg_r = build(ref, cls_name, u)
    same_region = bool(reg_o.maxradiussq == reg_r.maxradiussq and reg_o.enlarge == reg_r.enlarge
                       and (reg_o.unormed == reg_r.unormed).all()
                       and (reg_o.ellipsoid_invcov == reg_r.ellipsoid_invcov).all())
    _STATE["region"] = reg_r
    done = flips = acc_o = 0
    t_ours = t_ref = 0.0
    flip_rows = []
    seed = 100
    ctx = mp.get_context("fork")
    while done < rows:
        m = min(chunk, rows - done)
        pts = np.vstack([draw(reg_r, m // 2, seed, 1.0), draw(reg_r, m - m // 2, seed + 1, 1.3)])
        seed += 2
        t0 = time.perf_counter()
        got = reg_o.inside(pts)
        t_ours += time.perf_counter() - t0
        _STATE["chunks"] = np.array_split(pts, cores)
        t0 = time.perf_counter()
        with ctx.Pool(cores, initializer=_init) as pool:   # forked per chunk: inherits the rows
            want = np.concatenate(pool.map(_ref_inside, range(cores), chunksize=1))
        t_ref += time.perf_counter() - t0
        bad = np.flatnonzero(got != want)
        flips += len(bad)
        for j in bad[:20]:
            t = reg_r.transformLayer.transform(pts[j:j + 1])
            dist = ((reg_r.unormed - t)**2).sum(axis=1).min()
            flip_rows.append({"ours": bool(got[j]), "reference": bool(want[j]),
                              "min_dist_over_r2_minus_1": float(dist / reg_r.maxradiussq - 1.0)})
        acc_o += int(got.sum())
        done += m
    return {"config": name, "region": cls_name, "n_live": n, "ndim": d, "proposals": done,
            "accepted_fraction": acc_o / float(done), "flips": flips, "flip_details": flip_rows,
            "region_state_identical_to_reference": same_region,
            "maxradiussq": reg_o.maxradiussq, "enlarge": reg_o.enlarge,
            "ours_s": t_ours, "reference_s": t_ref, "reference_cores": cores}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10 * (1 << 20))
    ap.add_argument("--chunk", type=int, default=1 << 20)
    ap.add_argument("--out", default=None)
    ap.add_argument("--configs", default="all")
    args = ap.parse_args()
    oracle.reference()
    cores = len(os.sched_getaffinity(0))
    cases = [("configs[1] N=4000 d=20 MLFriends", "MLFriends", 4000, 20, args.rows),
             ("configs[4] N=4000 d=5 MLFriends", "MLFriends", 4000, 5, args.rows),
             ("configs[4] N=4000 d=100 MLFriends", "MLFriends", 4000, 100, args.rows),
             ("configs[3] N=8000 d=50 RobustEllipsoidRegion", "RobustEllipsoidRegion", 8000, 50, args.rows)]
    if args.configs != "all":
        keep = set(int(i) for i in args.configs.split(","))
        cases = [c for i, c in enumerate(cases) if i in keep]
    lines = []
    for name, cls_name, n, d, rows in cases:
        rec = study(name, cls_name, n, d, rows, args.chunk, cores)
        lines.append(rec)
        print(json.dumps(rec), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            for rec in lines:
                f.write(json.dumps(rec) + "\n")
    return 1 if any(r["flips"] for r in lines) else 0


if __name__ == "__main__":
    sys.exit(main())
