#!/usr/bin/env python
"""Headline benchmark: proposed-points/sec through MLFriends.inside() + loglike, N_live=4000, d=20.

    python bench.py --gpus N --steps K --warmup W            (ours; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path)

A step = one pass of the hot path over one batch of synthetic proposals per GPU:
``mask = region.inside(c); L = loglike(c[mask])``.  Workload = BASELINE.json
configs[1]: a 20-D correlated live set of 4000 points (reference helper recipe,
tests/test_run.py:14-19), AffineLayer, MLFriends with a 30-round bootstrapped radius, proposals
drawn with the reference's wrapping-ellipsoid recipe (mlfriends.pyx:1145-1154), Gaussian
likelihood (docs/gauss.py:25-27).

Printed (rank 0, one JSON line): ``value`` = device-resident throughput (inputs already in HBM),
``e2e`` = the same metric through the C-ABI host call (pinned host buffers, H2D/D2H inside the
timed region), ``roofline`` for the dominant kernel (first-neighbour scan), ``cpu_baseline`` =
the unmodified reference (oracle/_ref) on this box's host cores over a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_LIVE = 4000
NDIM = 20
SIGMA = 0.01
NBOOT = 30
METRIC = "proposed-points/sec through MLFriends.inside + loglike, N_live=4000 d=20"
# the same workload description on both arms (ours and --impl reference)
WORKLOAD = ("BASELINE configs[1]: 20-D correlated Gaussian live set, N_live=4000, AffineLayer, "
            "MLFriends.inside() + Gaussian loglike, wrapping-ellipsoid proposals (accepting regime)")


def workload_config(world, batch):
    """The `config` object: identical on both arms (ours and --impl reference) by construction.
    The reference arm times a bounded sample of this workload (see its `cpu_baseline.sample`)."""
    return {"workload": WORKLOAD, "n_live": N_LIVE, "ndim": NDIM, "nbootstraps": NBOOT,
            "rows_per_step_per_gpu": int(batch),
            "l2_policy": "inputs larger than L2 (%.0f MB of proposals per step per GPU)" % (batch * NDIM * 8 / 1e6),
            "parallelism": "proposal rows sharded over %d GPU(s), no data-path collective; bootstrap "
                           "rounds sharded with one allreduce(MAX) per rebuild" % world}


def multimodal_live(n=2000, d=10, seed=5):
    """Eggbox-like live set (BASELINE configs[2] shape: N=2000, d=10): points around a lattice of
    modes, as a seeded eggbox run leaves them at a rebuild with many clusters."""
    rng = np.random.RandomState(seed)
    centres = rng.randint(0, 5, size=(n, d)) * 0.2 + 0.1
    return centres + rng.normal(size=(n, d)) * 0.01


def time_rebuild(mod, u, reps=3):
    """Median wall time of the 30-round bootstrapped radius + enlargement."""
    layer = mod.AffineLayer()
    layer.optimize(u, u)
    region = mod.MLFriends(u, layer)
    times, res = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        res = region.compute_enlargement(nbootstraps=NBOOT, rng=np.random.RandomState(2))
        times.append(time.perf_counter() - t0)
    return region, float(np.median(times)) * 1e3, res


# --------------------------------------------------------------------------------------
# workload (pure NumPy; identical for both arms)
# --------------------------------------------------------------------------------------
def make_live(n=N_LIVE, d=NDIM, seed=1):
    rng = np.random.RandomState(seed)
    z = rng.normal(size=(n, d))
    z /= ((z**2).sum(axis=1)**0.5).reshape((n, 1))
    z *= rng.uniform(size=(n, 1))**(1. / d)
    C = 0.5 * np.ones((d, d)) + 0.5 * np.eye(d)
    L = np.linalg.cholesky(C)
    return 0.5 + 0.05 * np.dot(z, L.T)


def build_region(mod, u):
    """mod = ultranest_b200.mlfriends (ours) or ultranest.mlfriends (reference)."""
    layer = mod.AffineLayer()
    layer.optimize(u, u)
    region = mod.MLFriends(u, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(
        nbootstraps=NBOOT, rng=np.random.RandomState(2))
    region.create_ellipsoid()
    return region


def make_candidates(region, m, seed):
    """The reference's sample_from_wrapping_ellipsoid draw (mlfriends.pyx:1145-1154), topped up
    to exactly m rows inside the unit cube."""
    rng = np.random.RandomState(seed)
    d = region.u.shape[1]
    out = np.empty((m, d))
    filled = 0
    while filled < m:
        ns = int((m - filled) * 1.1) + 16
        z = rng.normal(size=(ns, d))
        z /= ((z**2).sum(axis=1)**0.5).reshape((ns, 1))
        uu = z * region.enlarge**0.5 * rng.uniform(size=(ns, 1))**(1. / d)
        w = region.ellipsoid_center + np.dot(uu, region.ellipsoid_axes_T)
        w = w[np.logical_and(w > 0, w < 1).all(axis=1)]
        take = min(len(w), m - filled)
        out[filled:filled + take] = w[:take]
        filled += take
    return out


def numpy_loglike(theta):
    centers = 0.5
    return -0.5 * (((theta - centers) / SIGMA)**2).sum(axis=1) - 0.5 * np.log(2 * np.pi * SIGMA**2) * NDIM


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# reference (CPU) arm
# --------------------------------------------------------------------------------------
_REF_STATE = {}


def _ref_init():
    """Worker start-up: one BLAS thread per worker (the workers are the parallelism)."""
    try:
        import threadpoolctl
        _REF_STATE["blas_limit"] = threadpoolctl.threadpool_limits(1)
    except Exception:  # noqa: BLE001
        pass


def _ref_worker(i):
    """One worker's share of a step: the reference's own calls on its rows (inherited by fork,
    nothing is pickled but the index)."""
    region = _REF_STATE["region"]
    chunk = _REF_STATE["chunks"][i]
    m = region.inside(chunk)
    like = numpy_loglike(chunk[m])
    return int(m.sum()), float(like.sum()) if len(like) else 0.0


def reference_setup():
    """Region + candidates with the reference implementation (oracle/_ref when built here,
    else the oracle port)."""
    import oracle
    kind = "reference"
    try:
        oracle.reference()
        import ultranest.mlfriends as refmod
    except Exception:  # noqa: BLE001
        refmod = None
        kind = "port"
    u = make_live()
    if refmod is not None:
        region = build_region(refmod, u)
    else:
        region = _PortRegion(u)
    _REF_STATE["region"] = region
    return region, kind


class _PortRegion(object):
    """Oracle-port stand-in with the same inside() contract (only when oracle/_ref is absent)."""

    def __init__(self, u):
        from oracle import cport
        self.cport = cport
        self.u = u
        d = u.shape[1]
        self.ctr = np.mean(u, axis=0)
        cov = np.cov(u, rowvar=0) * (d + 2)
        w, v = np.linalg.eigh(cov)
        self.T = v * w**-0.5
        self.unormed = np.dot(u - self.ctr, self.T)
        self.maxradiussq, self.enlarge = cport.compute_enlargement(u, self.unormed, NBOOT, np.random.RandomState(2))
        self.ellipsoid_center, ecov = cport.bounding_ellipsoid(u)
        self.ellipsoid_invcov = np.linalg.inv(ecov)
        l, vv = np.linalg.eigh(self.ellipsoid_invcov)
        self.ellipsoid_axes_T = np.dot(vv, np.diag(1. / np.sqrt(l))).transpose()

    def inside(self, pts):
        return self.cport.region_inside(pts, self.unormed, lambda p: np.dot(p - self.ctr, self.T),
                                        self.maxradiussq, self.ellipsoid_center,
                                        self.ellipsoid_invcov, self.enlarge)


def reference_throughput(steps, warmup, rows_per_core=65536, cores=None):
    """Times `steps` bounded samples through region.inside + loglike on all host cores
    (row-sharded over forked workers, which is what the reference's `mpiexec -np K` does)."""
    import multiprocessing as mp
    region, kind = reference_setup()
    cores = cores or len(os.sched_getaffinity(0))
    sample_rows = rows_per_core * cores
    cand = make_candidates(region, sample_rows, 3)
    _REF_STATE["chunks"] = np.array_split(cand, cores)
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores, initializer=_ref_init) as pool:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, range(cores), chunksize=1)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    accepted = sum(r[0] for r in res)
    total = float(np.sum(times))
    value = sample_rows * len(times) / total
    return {"value": value, "unit": "points/s", "cores": cores, "kind": kind,
            "sample": "%d wrapping-ellipsoid proposals per step (%d per core), %d steps, "
                      "accept fraction %.3f" % (sample_rows, rows_per_core, len(times),
                                                accepted / float(sample_rows)),
            "ms_per_step": 1e3 * total / len(times), "rows": sample_rows}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    base = reference_throughput(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, args.batch),
        "rows_per_step": base["rows"],
        "cpu_baseline": {"value": base["value"], "unit": "points/s", "cores": base["cores"],
                         "kind": base["kind"], "sample": base["sample"]},
        "e2e": {"value": base["value"], "unit": "points/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:   # the region rebuild on one host core, for the rebuild numbers of our arm
        import ultranest.mlfriends as refmod
        rebuild = {}
        for name, u in (("configs[1] N=4000 d=20", make_live()), ("configs[2] eggbox-like N=2000 d=10", multimodal_live())):
            _, ms, res = time_rebuild(refmod, u, reps=1)
            rebuild[name] = {"rebuild_ms": ms, "r2": float(res[0]), "f": float(res[1]), "cores": 1}
        line["rebuild"] = rebuild
    except Exception as exc:  # noqa: BLE001
        line["rebuild"] = {"unavailable": str(exc)}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------
# ours
# --------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes
    import torch
    from ultranest_b200 import _native
    from ultranest_b200 import mlfriends as ours
    from ultranest_b200.likelihoods import GaussianLogLike

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = _native.get_engine()
    M = args.batch
    K, W = args.steps, max(args.warmup, 3)

    # ---- region + proposals
    u = make_live()
    t0 = time.perf_counter()
    region = build_region(ours, u)
    rebuild_s = time.perf_counter() - t0
    cand = make_candidates(region, M, 3 + rank)
    loglike = GaussianLogLike(0.5, SIGMA)
    kind, lparams = loglike.device_spec(NDIM)
    region._bind()

    # an explicit (non-default) stream: the kernels are launched on it and the CUDA events that
    # time them are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sh = ctypes.c_void_p(stream.cuda_stream)
    assert sh.value, "need a non-default stream handle"
    pts_dev = torch.from_numpy(cand).cuda()
    mask_dev = torch.empty(M, dtype=torch.uint8, device="cuda")
    like_dev = torch.empty(M, dtype=torch.float64, device="cuda")
    lp = _native.as_f64(lparams)

    def dev_step(first=False):
        eng.call("unb_region_inside_loglike_dev", pts_dev.data_ptr(), M, mask_dev.data_ptr(),
                 like_dev.data_ptr(), int(kind), lp.ctypes.data if first else None, sh)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- (1) device-resident value
    dev_step(first=True)
    for _ in range(W):
        dev_step()
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    launches0 = eng.stat(_native.STAT_KERNEL_LAUNCHES)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        dev_step()
    e1.record(stream)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.stat(_native.STAT_KERNEL_LAUNCHES) - launches0
    accepted = int(mask_dev.sum().item())

    # ---- (2) dominant kernel alone (first-neighbour scan over t-space proposals)
    tcand = region.transformLayer.transform(cand)
    t_dev = torch.from_numpy(tcand).cuda()
    idx_dev = torch.empty(M, dtype=torch.int64, device="cuda")
    mask2_dev = torch.empty(M, dtype=torch.uint8, device="cuda")
    # ordered first-index scan once (untimed): tells how many live points the reference's
    # early-exit loop visits for these proposals
    eng.call("unb_region_find_nearby_dev", t_dev.data_ptr(), M, idx_dev.data_ptr(), None, sh)

    def scan_step():   # the membership kernel inside() runs (mask only)
        eng.call("unb_region_find_nearby_dev", t_dev.data_ptr(), M, None, mask2_dev.data_ptr(), sh)

    for _ in range(W):
        scan_step()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    for _ in range(K):
        scan_step()
    k1.record(stream)
    torch.cuda.synchronize()
    scan_ms = k0.elapsed_time(k1) / K
    tile_units = eng.stat(_native.STAT_TILE_VISITS)      # of the last launch
    fp64_peak = eng.fp64_peak()                          # lane-FMAs/s, measured on this device
    fp32_peak = eng.fp32_peak()
    # same launch with the fp32 pre-filter disabled (pure fp64 filter), for the record
    eng.set_option(_native.OPT_FILTER_FP32, 0)
    for _ in range(2):
        scan_step()
    torch.cuda.synchronize()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0.record(stream)
    for _ in range(K):
        scan_step()
    q1.record(stream)
    torch.cuda.synchronize()
    scan64_ms = q0.elapsed_time(q1) / K
    tile_units64 = eng.stat(_native.STAT_TILE_VISITS)
    eng.set_option(_native.OPT_FILTER_FP32, 1)
    idx_host = idx_dev.cpu().numpy()
    assert ((idx_host >= 0) == mask2_dev.cpu().numpy().astype(bool)).all()
    # pair-dimensions the reference's early-exit loop evaluates for these proposals
    scanned = np.where(idx_host >= 0, idx_host + 1, N_LIVE).astype(np.float64).sum()

    # ---- (3) end to end through the C ABI with pinned host buffers
    pin_pts = torch.empty((M, NDIM), dtype=torch.float64).pin_memory()
    pin_pts.numpy()[...] = cand
    pin_mask = torch.empty(M, dtype=torch.uint8).pin_memory()
    pin_like = torch.empty(M, dtype=torch.float64).pin_memory()
    np_pts, np_mask, np_like = pin_pts.numpy(), pin_mask.numpy().view(bool), pin_like.numpy()

    def e2e_step():
        eng.region_inside_loglike(np_pts, kind, lparams, mask_out=np_mask, like_out=np_like)

    for _ in range(W):
        e2e_step()
    barrier()
    h2d0, d2h0 = eng.stat(_native.STAT_H2D_BYTES), eng.stat(_native.STAT_D2H_BYTES)
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    h2d = (eng.stat(_native.STAT_H2D_BYTES) - h2d0) // K
    d2h = (eng.stat(_native.STAT_D2H_BYTES) - d2h0) // K
    # the same call with ordinary (pageable) NumPy buffers, as an UltraNest user would pass them
    def pageable_step():
        return eng.region_inside_loglike(cand, kind, lparams)

    for _ in range(2):
        pageable_step()
    t0 = time.perf_counter()
    for _ in range(max(K // 2, 1)):
        pageable_step()
    pageable_ms = 1e3 * (time.perf_counter() - t0) / max(K // 2, 1)
    # fused refill chain (membership -> transform -> loglike -> logl > Lmin; refill.py), same buffers
    finite = np_like[np.isfinite(np_like)]
    Lmin = float(np.median(finite)) if len(finite) else 0.0
    rf = eng.region_refill(np_pts, 2, False, None, None, kind, lparams, Lmin)
    t0 = time.perf_counter()
    for _ in range(max(K // 2, 1)):
        rf = eng.region_refill(np_pts, 2, False, None, None, kind, lparams, Lmin)
    refill_ms = 1e3 * (time.perf_counter() - t0) / max(K // 2, 1)
    refill_ok = bool(((rf[0] & 1).astype(bool) == np_mask).all()) and rf[2][2] == int((np_like > Lmin).sum())
    # throughput mode: the proposals are DRAWN ON THE DEVICE (Philox; not the reference's random
    # stream), filtered by the region, likelihood + `> Lmin` cut fused, and only the accepted rows
    # and their likelihoods come back (pinned buffers) -- no H2D at all
    pin_rows = torch.empty((M, NDIM), dtype=torch.float64).pin_memory()
    pin_like2 = torch.empty(M, dtype=torch.float64).pin_memory()
    n_acc = ctypes.c_int64(0)
    sdesc, skeep = _native.make_sample_desc(_native.SAMPLE_WRAPPING_ELLIPSOID, 20261017 + rank, 0,
                                            region.ellipsoid_axes_T, kind, lparams, Lmin)

    def rng_step(i):
        sdesc.offset = i * M
        eng.call("unb_region_sample", ctypes.addressof(sdesc), M, pin_rows.data_ptr(),
                 pin_like2.data_ptr(), ctypes.byref(n_acc), None)

    for i in range(W):
        rng_step(i)
    barrier()
    d2h1 = eng.stat(_native.STAT_D2H_BYTES)
    h2d1 = eng.stat(_native.STAT_H2D_BYTES)
    t0 = time.perf_counter()
    for i in range(K):
        rng_step(W + i)
    torch.cuda.synchronize()
    rng_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    rng_d2h = (eng.stat(_native.STAT_D2H_BYTES) - d2h1) // K
    rng_h2d = (eng.stat(_native.STAT_H2D_BYTES) - h2d1) // K
    rng_rows = int(n_acc.value)
    rng_ok = bool(rng_rows > 0 and (pin_like2.numpy()[:rng_rows] > Lmin).all()
                  and region.inside(pin_rows.numpy()[:min(rng_rows, 4096)]).all())
    del skeep
    clk = clocks.stop()
    e2e_ok = bool((np_mask.view(np.uint8) == mask_dev.cpu().numpy()).all()) if world == 1 else True

    # ---- (4) region rebuild: 30 bootstrap rounds, un-sharded and (N > 1) sharded over the ranks
    # with ONE NCCL allreduce(MAX) (integrator.py:388-404); device-timed collective, max over ranks
    from ultranest_b200 import distributed as D
    rebuild = {}
    for name, ulive in (("configs[1] N=4000 d=20", u), ("configs[2] eggbox-like N=2000 d=10", multimodal_live())):
        reg_r, ms_1, res_1 = time_rebuild(ours, ulive)
        rec = {"rebuild_ms_1rank": ms_1, "r2": float(res_1[0]), "f": float(res_1[1])}
        if dist is not None:
            D.enable()
            try:
                times, colls = [], []
                for _ in range(4):
                    barrier()
                    t0 = time.perf_counter()
                    res_s = reg_r.compute_enlargement(nbootstraps=NBOOT, rng=np.random.RandomState(2))
                    times.append(max_over_ranks(1e3 * (time.perf_counter() - t0)))
                    colls.append(max_over_ranks(D.last_timings.get("collective_us", 0.0)))
                rec.update({"rebuild_ms_sharded": float(np.median(times[1:])),
                            "collective_us": float(np.median(colls[1:])),
                            "identical_to_1rank": bool(res_s == res_1), "ranks": world})
            finally:
                D.disable()
        rebuild[name] = rec

    if dist is not None:
        dist.barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the scan kernel (HBM roofline as BASELINE.json prescribes; see DESIGN.md
    # for why the fp64 pipe, not HBM, is the binding ceiling of this kernel)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # DRAM traffic per launch of the dominant kernel from the committed ncu capture (profiles/)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as f:
            tr = json.load(f)
        for name, rec in tr.items():
            if name.startswith("k_inside_any"):
                traffic = rec["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    alg_bytes = M * (8.0 * NDIM + 8.0) + N_LIVE * NDIM * 8.0
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": "k_inside_any32<20,2> (any-neighbour membership scan: fp32 pre-filter, certain hits retire at once, uncertain pairs decided in exact fp64)",
        "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
        "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
        "traffic": traffic, "traffic_note": "dram__bytes_read+write per launch, ncu --set full capture "
                                              "of this command (profiles/kernel_traffic.json); "
                                              "algorithmic bytes per launch = %.0f" % alg_bytes,
        "ms_per_launch": scan_ms,
        # compute roofline: this kernel is bound by the fp64 pipe, not by HBM
        "compute": {"filter": "fp32 pre-filter + certain-neighbour level + exact fp64 decisions for the uncertain shell",
                    "peak_ffma_per_s": fp32_peak, "peak_dfma_per_s": fp64_peak,
                    "peak_source": "measured on this device (unb_fp32_peak / unb_fp64_peak, FMA chains)",
                    "executed_ffma_per_s": tile_units * 32.0 * 64.0 * NDIM / (scan_ms * 1e-3),
                    "frac_of_fp32_peak": tile_units * 32.0 * 64.0 * NDIM / (scan_ms * 1e-3) / fp32_peak,
                    "useful_pair_dims_per_s": scanned * NDIM / (scan_ms * 1e-3),
                    "fp64_filter_variant": {
                        "ms_per_launch": scan64_ms,
                        "executed_dfma_per_s": tile_units64 * 32.0 * 64.0 * NDIM / (scan64_ms * 1e-3),
                        "frac_of_fp64_peak": tile_units64 * 32.0 * 64.0 * NDIM / (scan64_ms * 1e-3) / fp64_peak},
                    "note": "executed = filter FMAs actually issued (warp-tiles x lanes); useful = "
                            "pair-dimensions the reference's early-exit loop evaluates for the same "
                            "proposals (3 DP instructions each there)"},
    }
    line = {
        "metric": METRIC, "value": world * M * K / (dev_ms * 1e-3), "unit": "points/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(world, M),
        "details": {"accept_fraction": accepted / float(M), "region_rebuild_s": rebuild_s,
                    "maxradiussq": region.maxradiussq, "enlarge": region.enlarge},
        "rebuild": rebuild,
        "e2e": {"value": world * M * K / (e2e_ms * 1e-3), "unit": "points/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms / K, "matches_device_path": e2e_ok,
                "pageable_buffers_ms_per_step": pageable_ms,
                "pageable_buffers_value": M / (pageable_ms * 1e-3),
                "call": "unb_region_inside_loglike (pinned host buffers, chunked double-buffered)",
                "refill_call": {"ms_per_step": refill_ms, "value": M / (refill_ms * 1e-3),
                                "matches": refill_ok, "accepted_rows": rf[2][2],
                                "call": "unb_region_refill (integrator.py:1773-1805 as one pipeline)"},
                "device_rng_call": {"value": world * M * K / (rng_ms * 1e-3), "ms_per_step": rng_ms / K,
                                    "h2d_bytes_per_step": int(rng_h2d), "d2h_bytes_per_step": int(rng_d2h),
                                    "accepted_rows_last_step": rng_rows, "checks_ok": rng_ok,
                                    "parity": "NON-PARITY random stream: proposals drawn on the device "
                                              "(Philox4x32-10), statistically the reference's "
                                              "sample_from_wrapping_ellipsoid; same region filter, "
                                              "likelihood and logl > Lmin cut (Lmin = median)",
                                    "call": "unb_region_sample (opt-in: region.device_rng = True)"}},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "clocks": clk,
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            base = reference_throughput(steps=3, warmup=1, rows_per_core=32768)
            line["cpu_baseline"] = {"value": base["value"], "unit": "points/s",
                                    "cores": base["cores"], "kind": base["kind"],
                                    "sample": base["sample"]}
        except Exception as exc:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "points/s", "cores": 0, "kind": "port",
                                    "sample": "failed: %s" % exc}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------
# the other BASELINE.json configurations (not the driver's headline line): --config 2|3|4
# --------------------------------------------------------------------------------------
def run_config(args):
    """One JSON line for BASELINE.json configs[2], [3] or [4] at N GPUs (torchrun for N > 1):
    rows / rounds sharded over the ranks, device-timed `value`, warmed host-buffer `e2e`."""
    import ctypes
    import torch
    from ultranest_b200 import _native
    from ultranest_b200 import distributed as D
    from ultranest_b200 import mlfriends as ours
    from ultranest_b200.likelihoods import RosenbrockLogLike

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = _native.get_engine()
    K, W = args.steps, max(args.warmup, 3)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sh = ctypes.c_void_p(stream.cuda_stream)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def dev_timed(fn):
        for _ in range(W):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(K):
            fn()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / K

    def host_timed(fn):
        for _ in range(W):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            fn()
        torch.cuda.synchronize()
        return max_over_ranks(1e3 * (time.perf_counter() - t0)) / K

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64).pin_memory()
        t.numpy()[...] = a
        return t

    clocks = ClockSampler(local_rank)
    clocks.start()
    line = {"n_gpus": world, "steps": K, "warmup": W, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    if args.config == 2:
        # eggbox-like 10-D, N_live=2000: the 30-round bootstrapped radius + enlargement, rounds
        # sharded over the ranks, ONE NCCL allreduce(MAX) per rebuild (integrator.py:388-404)
        u = multimodal_live()
        region, ms_1, res_1 = time_rebuild(ours, u, reps=5)
        rec = {"rebuild_ms_1rank": ms_1, "r2": float(res_1[0]), "f": float(res_1[1])}
        ms = ms_1
        if dist is not None:
            D.enable()
            times, colls = [], []
            for _ in range(W + K):
                barrier()
                t0 = time.perf_counter()
                res_s = region.compute_enlargement(nbootstraps=NBOOT, rng=np.random.RandomState(2))
                times.append(max_over_ranks(1e3 * (time.perf_counter() - t0)))
                colls.append(max_over_ranks(D.last_timings.get("collective_us", 0.0)))
            D.disable()
            ms = float(np.median(times[W:]))
            rec.update({"rebuild_ms_sharded": ms, "collective_us": float(np.median(colls[W:])),
                        "identical_to_1rank": bool(res_s == res_1)})
        line.update({"metric": "region rebuilds/sec (30-round bootstrapped max-radius + enlargement), eggbox-like N_live=2000 d=10",
                     "value": 1e3 / ms, "unit": "rebuilds/s", "ms_per_step": ms,
                     "config": {"workload": "BASELINE configs[2]: eggbox-like 10-D live set, N_live=2000, "
                                            "bootstrap rounds sharded over %d GPU(s), one allreduce(MAX)" % world,
                                "n_live": 2000, "ndim": 10, "nbootstraps": NBOOT},
                     "rebuild": rec, "gpu_launches": int(eng.stat(_native.STAT_KERNEL_LAUNCHES))})
    elif args.config == 3:
        # 50-D, N_live=8000 RobustEllipsoidRegion: Mahalanobis filter + Rosenbrock batch likelihood
        n, d, M = 8000, 50, args.batch // 4
        u = make_live(n, d, seed=1)
        layer = ours.AffineLayer()
        layer.optimize(u, u)
        region = ours.RobustEllipsoidRegion(u, layer)
        region.maxradiussq, region.enlarge = region.compute_enlargement(nbootstraps=NBOOT, rng=np.random.RandomState(2))
        region.create_ellipsoid()
        cand = make_candidates(region, M, 3 + rank)
        loglike = RosenbrockLogLike()
        mask0 = region.inside(cand)
        c_dev = torch.from_numpy(cand).cuda()
        m_dev = torch.empty(M, dtype=torch.uint8, device="cuda")
        l_dev = torch.empty(M, dtype=torch.float64, device="cuda")
        launches0 = eng.stat(_native.STAT_KERNEL_LAUNCHES)

        def dev_step():
            eng.call("unb_region_inside_ellipsoid_dev", c_dev.data_ptr(), M, m_dev.data_ptr(), sh)
            eng.call("unb_loglike_rosenbrock_dev", c_dev.data_ptr(), d, M, l_dev.data_ptr(), m_dev.data_ptr(), sh)

        have_dev_like = hasattr(eng.lib, "unb_loglike_rosenbrock_dev")
        if not have_dev_like:
            def dev_step():   # noqa: F811 -- filter alone (the likelihood has no _dev entry point)
                eng.call("unb_region_inside_ellipsoid_dev", c_dev.data_ptr(), M, m_dev.data_ptr(), sh)
        dev_ms = dev_timed(dev_step)
        launches = (eng.stat(_native.STAT_KERNEL_LAUNCHES) - launches0) // (K + W)
        pin = pinned(cand)
        h2d0, d2h0 = eng.stat(_native.STAT_H2D_BYTES), eng.stat(_native.STAT_D2H_BYTES)

        def e2e_step():
            m = region.inside(pin.numpy())
            return loglike(pin.numpy()[m] * 20 - 10) if args.with_loglike else m

        e2e_ms = host_timed(e2e_step)
        steps_run = K + W
        alg = M * (8.0 * d + 1.0)
        line.update({"metric": "proposed-points/sec through RobustEllipsoidRegion.inside (Mahalanobis filter), N_live=8000 d=50",
                     "value": world * M / (dev_ms * 1e-3), "unit": "points/s", "ms_per_step": dev_ms,
                     "config": {"workload": "BASELINE configs[3]: 50-D, N_live=8000, RobustEllipsoidRegion "
                                            "(ellipsoid filter alone), wrapping-ellipsoid proposals; rows sharded over %d GPU(s)" % world,
                                "n_live": n, "ndim": d, "rows_per_step_per_gpu": M,
                                "l2_policy": "inputs larger than L2 (%.0f MB per step per GPU)" % (M * d * 8 / 1e6),
                                "device_step": "filter + Rosenbrock likelihood" if have_dev_like else "filter"},
                     "e2e": {"value": world * M / (e2e_ms * 1e-3), "unit": "points/s", "ms_per_step": e2e_ms,
                             "h2d_bytes_per_step": int((eng.stat(_native.STAT_H2D_BYTES) - h2d0) // steps_run),
                             "d2h_bytes_per_step": int((eng.stat(_native.STAT_D2H_BYTES) - d2h0) // steps_run),
                             "call": "RobustEllipsoidRegion.inside(pinned host rows)" + (" + RosenbrockLogLike" if args.with_loglike else "")},
                     "roofline": {"bound": "hbm", "kernel": "k_prep_tile (ellipsoid filter, d=50)",
                                  "achieved": alg / (dev_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                  "frac": alg / (dev_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                                  "note": "whole device step; algorithmic bytes = 8d+1 per proposal"},
                     "details": {"accept_fraction": float(mask0.mean()), "enlarge": region.enlarge},
                     "gpu_launches": int(launches)})
    else:
        # membership microbench: M proposals x N_live=4000 x d in {5, 20, 100}; accepting (A) and
        # full-scan (R) regimes; device-resident and through the host API (pinned t-space rows)
        cases = []
        for d in (5, 20, 100):
            M = args.batch if d <= 20 else args.batch // 8
            u = make_live(N_LIVE, d, seed=1)
            layer = ours.AffineLayer()
            layer.optimize(u, u)
            region = ours.MLFriends(u, layer)
            region.maxradiussq, region.enlarge = region.compute_enlargement(nbootstraps=NBOOT, rng=np.random.RandomState(2))
            region.create_ellipsoid()
            region._bind()
            cand = make_candidates(region, M, 3 + rank)
            rng = np.random.RandomState(4 + rank)
            r = region.maxradiussq**0.5
            tbox = rng.uniform(region.bbox_lo - r, region.bbox_hi + r, size=(M, d))
            for regime, tpts in (("A", region.transformLayer.transform(cand)), ("R", tbox)):
                tpts = np.ascontiguousarray(tpts)
                t_dev = torch.from_numpy(tpts).cuda()
                mask = torch.empty(M, dtype=torch.uint8, device="cuda")
                ms_any = dev_timed(lambda: eng.call("unb_region_find_nearby_dev", t_dev.data_ptr(), M,
                                                    None, mask.data_ptr(), sh))
                tile_units = eng.stat(_native.STAT_TILE_VISITS)
                pin = pinned(tpts)
                ms_host = host_timed(lambda: eng.region_has_neighbour(pin.numpy()))
                alg = M * (8.0 * d + 1.0) + N_LIVE * d * 8.0
                cases.append({"ndim": d, "regime": regime, "rows_per_gpu": M,
                              "accept": float(mask.float().mean().item()),
                              "ms_device": ms_any, "points_per_s": world * M / (ms_any * 1e-3),
                              "hbm_gbs": alg / (ms_any * 1e-3) / 1e9, "hbm_frac": alg / (ms_any * 1e-3) / 1e9 / hbm_peak,
                              "filter_lane_fma_per_s": tile_units * 32.0 * 64.0 * d / (ms_any * 1e-3),
                              "ms_host_api": ms_host, "points_per_s_host_api": world * M / (ms_host * 1e-3)})
        head = [c for c in cases if c["ndim"] == 20 and c["regime"] == "R"][0]
        line.update({"metric": "membership-test candidates/sec, 10^6 candidates x 4000 live x d in {5,20,100}",
                     "value": head["points_per_s"], "unit": "points/s", "ms_per_step": head["ms_device"],
                     "config": {"workload": "BASELINE configs[4]: membership microbench, N_live=4000, d in {5,20,100}, "
                                            "accepting (A) and full-scan (R) regimes; `value` = d=20 full-scan; rows sharded over %d GPU(s)" % world,
                                "n_live": N_LIVE, "l2_policy": "inputs larger than L2 except d=5 (42 MB)"},
                     "cases": cases, "hbm_peak_gbs": hbm_peak,
                     "gpu_launches": int(eng.stat(_native.STAT_KERNEL_LAUNCHES))})
    line["clocks"] = clocks.stop()
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="proposals per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4],
                    help="BASELINE.json configs[k]; 1 (default) is the headline the driver runs")
    ap.add_argument("--with-loglike", action="store_true", help="--config 3: add the likelihood to the e2e step")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config != 1:
        return run_config(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
