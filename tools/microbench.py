#!/usr/bin/env python
"""Membership-test microbench (BASELINE.json configs[4]): M proposals x N_live=4000 x d in
{5,20,100}, accepting (A: wrapping-ellipsoid draws) and rejecting (R: uniform in the t-space
bounding box, full scans) regimes; plus the ellipsoid-only filter at N=8000, d=50 (configs[3])
and the bootstrapped radius (configs[1,2]).  Device-resident timings with CUDA events.

    python tools/microbench.py [--m 1048576] [--quick]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ultranest_b200 import _native  # noqa: E402
from ultranest_b200 import mlfriends as m  # noqa: E402


def timed(fn, stream, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=1 << 20)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    eng = _native.get_engine()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sh = ctypes.c_void_p(stream.cuda_stream)
    out = []
    dims = [20] if args.quick else [5, 20, 100]
    for d in dims:
        M = args.m if d <= 20 else args.m // 8
        u = bench.make_live(4000, d, seed=1)
        layer = m.AffineLayer()
        layer.optimize(u, u)
        region = m.MLFriends(u, layer)
        t0 = time.perf_counter()
        region.maxradiussq, region.enlarge = region.compute_enlargement(30, rng=np.random.RandomState(2))
        t_boot = time.perf_counter() - t0
        region.create_ellipsoid()
        region._bind()
        cand = bench.make_candidates(region, M, 3)
        rng = np.random.RandomState(4)
        r = region.maxradiussq**0.5
        tbox = rng.uniform(region.bbox_lo - r, region.bbox_hi + r, size=(M, d))
        for regime, tpts in (("A", region.transformLayer.transform(cand)), ("R", tbox)):
            t_dev = torch.from_numpy(np.ascontiguousarray(tpts)).cuda()
            idx = torch.empty(M, dtype=torch.int64, device="cuda")
            mask = torch.empty(M, dtype=torch.uint8, device="cuda")
            ms_find = timed(lambda: eng.call("unb_region_find_nearby_dev", t_dev.data_ptr(), M,
                                             idx.data_ptr(), None, sh), stream)
            ms_any = timed(lambda: eng.call("unb_region_find_nearby_dev", t_dev.data_ptr(), M,
                                            None, mask.data_ptr(), sh), stream)
            ih = idx.cpu().numpy()
            assert ((ih >= 0) == mask.cpu().numpy().astype(bool)).all()
            visited = np.where(ih >= 0, ih + 1, 4000).astype(np.float64)
            out.append(dict(case="membership", d=d, n_live=4000, m=M, regime=regime,
                            accept=float((ih >= 0).mean()), ms_find_first=ms_find, ms_any=ms_any,
                            points_per_s_any=M / ms_any * 1e3, points_per_s_find=M / ms_find * 1e3,
                            ref_pairdims_per_s_any=visited.sum() * d / ms_any * 1e3,
                            fullscan_pairdims_per_s_find=(4000.0 * M * d / ms_find * 1e3) if regime == "R" else None,
                            hbm_gbs_any=(M * (8 * d + 1) + 4000 * d * 8) / ms_any / 1e6))
        if d == 20:   # launch-size sweep of the membership kernel: slope = steady state, intercept = drain
            t_all = torch.from_numpy(np.ascontiguousarray(region.transformLayer.transform(cand))).cuda()
            mask = torch.empty(M, dtype=torch.uint8, device="cuda")
            for msub in (4096, 65536, 262144, M):
                ms = timed(lambda: eng.call("unb_region_find_nearby_dev", t_all.data_ptr(), msub,
                                            None, mask.data_ptr(), sh), stream)
                out.append(dict(case="any_size_sweep", d=d, m=msub, ms=ms,
                                tile_units=eng.stat(_native.STAT_TILE_VISITS)))
        if d == 20:   # the integrator's per-iteration pattern (integrator.py:1855, 2749-2758)
            u0 = region.u.copy()
            t0 = time.perf_counter()
            nit = 200
            for i in range(nit):
                worst = i % len(u0)
                unew = u0[(i * 7 + 3) % len(u0)] + 1e-4
                region.u[worst] = unew
                region.unormed[worst] = region.transformLayer.transform(unew)
                region.ellipsoid_center = np.mean(region.u, axis=0)
                ok = region.inside(region.u)
            per_it = (time.perf_counter() - t0) / nit
            out.append(dict(case="integrator_iteration_inside_live", d=d, n_live=4000,
                            us_per_iteration=per_it * 1e6, all_inside=bool(ok.all())))
            for ndraw in (128, 4096, 65536):
                for meth in region.sampling_methods:
                    np.random.seed(1)
                    meth(nsamples=ndraw)
                    t0 = time.perf_counter()
                    for _ in range(10):
                        got = meth(nsamples=ndraw)
                    out.append(dict(case="sample", method=meth.__name__, ndraw=ndraw,
                                    us_per_call=(time.perf_counter() - t0) / 10 * 1e6,
                                    returned=int(len(got))))
        # fused inside (u-space proposals, mask only)
        p_dev = torch.from_numpy(cand).cuda()
        mask = torch.empty(M, dtype=torch.uint8, device="cuda")
        ms_inside = timed(lambda: eng.call("unb_region_inside_dev", p_dev.data_ptr(), M,
                                           mask.data_ptr(), sh), stream)
        out.append(dict(case="inside_fused", d=d, n_live=4000, m=M, ms=ms_inside,
                        points_per_s=M / ms_inside * 1e3, bootstrap30_s=t_boot))
    if not args.quick:
        # configs[3]: ellipsoid-only region, N=8000, d=50
        d, n, M = 50, 8000, args.m // 4
        u = bench.make_live(n, d, seed=1)
        layer = m.AffineLayer()
        layer.optimize(u, u)
        region = m.RobustEllipsoidRegion(u, layer)
        t0 = time.perf_counter()
        region.maxradiussq, region.enlarge = region.compute_enlargement(30, rng=np.random.RandomState(2))
        t_boot = time.perf_counter() - t0
        region.create_ellipsoid()
        cand = bench.make_candidates(region, M, 3)
        t0 = time.perf_counter()
        mask = region.inside(cand)
        t_host = time.perf_counter() - t0
        out.append(dict(case="robust_ellipsoid_inside_hostapi", d=d, n_live=n, m=M, s=t_host,
                        points_per_s=M / t_host, accept=float(mask.mean()), bootstrap30_s=t_boot))
        c_dev = torch.from_numpy(cand).cuda()
        mk = torch.empty(M, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: eng.call("unb_region_inside_ellipsoid_dev", c_dev.data_ptr(), M,
                                    mk.data_ptr(), sh), stream)
        assert (mk.cpu().numpy().astype(bool) == mask).all()
        out.append(dict(case="robust_ellipsoid_inside_device", d=d, n_live=n, m=M, ms=ms,
                        points_per_s=M / ms * 1e3, hbm_gbs=M * (8 * d + 1) / ms / 1e6,
                        dp_instr_per_s=3.0 * d * d * M / ms * 1e3))
        # configs[2]: eggbox-like multimodal live set d=10, N=2000: rebuild time
        rng = np.random.RandomState(5)
        centres = rng.randint(0, 5, size=(2000, 10)) * 0.2 + 0.1
        u = centres + rng.normal(size=(2000, 10)) * 0.01
        layer = m.AffineLayer()
        layer.optimize(u, u)
        region = m.MLFriends(u, layer)
        t0 = time.perf_counter()
        r2, f = region.compute_enlargement(30, rng=np.random.RandomState(2))
        out.append(dict(case="bootstrap_multimodal", d=10, n_live=2000, rounds=30,
                        s=time.perf_counter() - t0, r2=r2, f=f))
    for row in out:
        print(json.dumps(row))


if __name__ == "__main__":
    main()
