#!/usr/bin/env python
"""Whole-run comparison: the reference's UNMODIFIED ReactiveNestedSampler (oracle/_ref) on the same
seeded problem with (a) its own Cython region module on the host, (b) ultranest_b200 installed
behind the same names, (c) additionally the likelihood on the device and the fused
``_refill_samples``, (d) additionally the exact fast ``MultiCounter`` (ultranest_b200.netiter).
Each arm runs in its own process; the runs must be the same run
(niter, ncall, logZ), so the wall-clock ratio is the end-to-end effect of the drop-in.

    python tests/run_compare.py [--ndim 20] [--nlive 4000] [--max-ncalls 8000]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def arm(mode, ndim, nlive, max_ncalls, sigma):
    sys.path.insert(0, ROOT)
    import numpy as np
    import oracle
    oracle.reference()
    if mode != "reference":
        import ultranest_b200
        ultranest_b200.install(force=True)
    from ultranest import ReactiveNestedSampler
    if mode == "full":   # + the exact fast MultiCounter
        import ultranest_b200
        ultranest_b200.install(force=True, netiter=True)

    def numpy_loglike(theta):
        return -0.5 * (((theta - 0.5) / sigma)**2).sum(axis=1) - 0.5 * np.log(2 * np.pi * sigma**2) * ndim

    loglike, transform = numpy_loglike, (lambda x: x)
    if mode in ("device", "full"):
        from ultranest_b200.likelihoods import GaussianLogLike
        from ultranest_b200.transforms import IdentityTransform
        loglike, transform = GaussianLogLike(0.5, sigma), IdentityTransform()
    np.random.seed(7)
    sampler = ReactiveNestedSampler(["p%d" % i for i in range(ndim)], loglike, transform=transform,
                                    log_dir=None, vectorized=True)
    stats = None
    if mode in ("device", "full"):
        from ultranest_b200 import refill
        stats = refill.attach(sampler)
    t0 = time.perf_counter()
    res = sampler.run(min_num_live_points=nlive, max_ncalls=max_ncalls, viz_callback=False,
                      show_status=False)
    wall = time.perf_counter() - t0
    out = dict(mode=mode, wall_s=wall, niter=int(res["niter"]), ncall=int(res["ncall"]),
               ncall_region=int(sampler.ncall_region), logz=float(res["logz"]),
               region=type(sampler.region).__module__,
               iterations_per_s=res["niter"] / wall, region_proposals_per_s=sampler.ncall_region / wall)
    if stats is not None:
        out["refill"] = stats
    print("RESULT " + json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ndim", type=int, default=20)
    ap.add_argument("--nlive", type=int, default=4000)
    ap.add_argument("--max-ncalls", type=int, default=8000)
    ap.add_argument("--sigma", type=float, default=0.05)
    ap.add_argument("--arm", default=None)
    ap.add_argument("--arms", default="reference,ours,device,full")
    args = ap.parse_args()
    if args.arm:
        return arm(args.arm, args.ndim, args.nlive, args.max_ncalls, args.sigma)
    rows = []
    for mode in args.arms.split(","):
        cmd = [sys.executable, os.path.abspath(__file__), "--arm", mode, "--ndim", str(args.ndim),
               "--nlive", str(args.nlive), "--max-ncalls", str(args.max_ncalls), "--sigma", str(args.sigma)]
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1") if mode == "reference" else os.environ
        out = subprocess.run(cmd, capture_output=True, text=True, env=env)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT ")]
        if not line:
            sys.stderr.write(out.stdout[-2000:] + out.stderr[-4000:])
            raise SystemExit("arm %s failed" % mode)
        rows.append(json.loads(line[0][7:]))
        print(json.dumps(rows[-1]))
    if len(rows) > 1:
        ref = rows[0]
        for r in rows[1:]:
            same = (r["niter"], r["ncall"]) == (ref["niter"], ref["ncall"])
            print(json.dumps(dict(case="whole_run", ndim=args.ndim, nlive=args.nlive, mode=r["mode"],
                                  speedup_vs_reference=ref["wall_s"] / r["wall_s"], same_run=same,
                                  logz_rel_diff=abs(r["logz"] - ref["logz"]) / abs(ref["logz"]))))


if __name__ == "__main__":
    main()
