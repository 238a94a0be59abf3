"""Multi-GPU sharding of the region path: one process per GPU (``torchrun``), replicated region
state, and exactly one collective per region rebuild.

Where the path shards:

* **proposals** (``inside`` / ``find_nearby`` / ellipsoid / loglike) are independent rows: each
  rank takes a contiguous slice, no data-path collective; :func:`allgather_rows` re-unites masks
  when a caller needs the full vector.
* **bootstrap rounds** are independent: rank ``r`` evaluates rounds ``[lo_r, hi_r)`` and ONE
  ``all_reduce(MAX)`` over a 5-double device buffer ``[r2, f, failure, tag, -tag]`` replaces the
  reference's pickled ``gather`` + ``bcast`` (``integrator.py:395-404``).  The per-round results
  are folded into that buffer on the device (``unb_region_bootstrap_fold_dev``) and reduced in
  place by NCCL; the host reads 40 bytes once, after the collective.  For parity with the
  single-process oracle every rank draws the selection masks of ALL rounds from its own,
  identically seeded host stream (torchrun replicas of one seeded sampler); the tag pair proves
  that they agree.  (The reference's MPI mode re-seeds every rank, ``integrator.py:1239-1251``,
  and is therefore not comparable with its own 1-process run)

Backend: ``nccl`` on GPUs (NVLink 5 / NVSwitch), ``gloo`` in the CPU-only tests.  The payload is
40 bytes, so the collective is latency-bound; there is nothing to overlap or fuse.
"""
import numpy as np

_enabled = False
_group = None
# {"collective_us": ...} of the most recent reduce_enlargement on this rank (CUDA events around the
# all_reduce on the current stream; NCCL only) -- read by bench.py and the NCCL tests
last_timings = {}


def _dist():
    import torch.distributed as dist
    return dist


def enable(group=None):
    """Shard region work over the (already initialised) default process group."""
    global _enabled, _group
    dist = _dist()
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _enabled = True
    _group = group


def disable():
    global _enabled, _group
    _enabled = False
    _group = None


def world_size():
    if not _enabled:
        return 1
    return _dist().get_world_size(_group)


def rank():
    if not _enabled:
        return 0
    return _dist().get_rank(_group)


def shard_bounds(n, world, r):
    """Contiguous balanced slice ``[lo, hi)`` of ``n`` items for rank ``r`` of ``world``
    (first ``n % world`` ranks get one extra item, like ``np.array_split``)."""
    base, extra = divmod(int(n), int(world))
    lo = r * base + min(r, extra)
    hi = lo + base + (1 if r < extra else 0)
    return lo, hi


def _device():
    import torch
    dist = _dist()
    if dist.get_backend(_group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_array(arr, src=0):
    """Broadcast a NumPy array (same shape/dtype on every rank) from ``src``; returns the array."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(_device())
    _dist().broadcast(t, src=src, group=_group)
    return t.cpu().numpy()


def allreduce_max(values):
    """Element-wise max over ranks of a small float64 vector (ONE collective)."""
    import torch
    t = torch.tensor(np.asarray(values, dtype=np.float64), dtype=torch.float64, device=_device())
    dist = _dist()
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=_group)
    return t.cpu().numpy()


def allgather_rows(local, total_rows):
    """Concatenate per-rank row slices (made with :func:`shard_bounds`) in rank order."""
    import torch
    dist = _dist()
    world = dist.get_world_size(_group)
    local = np.ascontiguousarray(local)
    per = -(-int(total_rows) // world)   # ceil: equal-size buffers for all_gather
    pad_shape = (per,) + local.shape[1:]
    buf = np.zeros(pad_shape, dtype=local.dtype)
    buf[:len(local)] = local
    as_u8 = local.dtype == np.bool_
    send = torch.from_numpy(buf.view(np.uint8) if as_u8 else buf).to(_device())
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=_group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(total_rows, world, r)
        part = recv[r].cpu().numpy()[:hi - lo]
        parts.append(part.view(np.bool_) if as_u8 else part)
    return np.concatenate(parts, axis=0)


def _mask_tag(selected):
    """A checksum of the selection masks that is exact in a float64 (< 2^52): equal on all ranks
    iff they drew the same rounds.  Rides in the same reduction as ``[tag, -tag]``."""
    import zlib
    sel = np.ascontiguousarray(selected, dtype=np.uint8)
    crc = zlib.crc32(sel.tobytes()) & 0xffffffff
    return float(crc * 1024 + (sel.shape[0] % 1024))


def _host_rounds(u, selected, lo, hi, minvol):
    """The d x d host algebra of rounds ``[lo, hi)`` (``bounding_ellipsoid`` + ``inv``,
    mlfriends.pyx:1057-1058).  Returns ``(ctrs, invcovs, stop, failure)``: rounds from ``stop`` on
    are not evaluated because the algebra of round ``stop`` failed (``failure`` is the exception)."""
    from .mlfriends import bounding_ellipsoid
    nrounds = selected.shape[0]
    ndim = u.shape[1]
    active = ~(selected.all(axis=1) | ~selected.any(axis=1))
    ctrs = np.zeros((nrounds, ndim))
    invcovs = np.zeros((nrounds, ndim, ndim))
    for r in range(lo, hi):
        if not active[r]:
            continue
        try:
            ctr, cov = bounding_ellipsoid(u[selected[r], :], minvol=minvol)
            invcovs[r] = np.linalg.inv(cov)
            ctrs[r] = ctr
        except (np.linalg.LinAlgError, FloatingPointError, AssertionError, Warning) as exc:
            return ctrs, invcovs, r, exc
    return ctrs, invcovs, hi, None


def reduce_enlargement(u, unormed, selected, minvol=0., compute_rounds=None, masks="replicated",
                       timings=None):
    """Sharded ``compute_enlargement``: this rank's slice of the rounds on its GPU and ONE
    ``all_reduce(MAX)`` of ``[r2, f, failed, tag, -tag]`` -- the reference's ``comm.gather`` +
    ``np.max`` + ``comm.bcast`` (integrator.py:395-404) as a single exchange.  Returns
    ``(r2, f)``, identical on all ranks and identical to the single-process result (max is
    order-independent).

    ``masks="replicated"`` (default): every rank drew ``selected`` itself from an identically
    seeded stream (torchrun replicas of one seeded sampler); the tag pair in the same reduction
    proves it, a mismatch raises.  ``masks="broadcast"``: rank 0's masks are shipped first (a
    second collective) -- for callers whose ranks do not share a random stream.

    On NCCL a rank's slice first goes through the enlargement screen
    (``mlfriends._bootstrap_rounds_screened``: device moments, the reference's NumPy algebra only
    for the rounds that can decide the maximum); when the screen does not apply (fewer than three
    local rounds, few points, ``minvol``) the exact per-round algebra runs on the host and the
    per-round results never leave the device: the library folds them into the 5-double buffer
    that the collective reduces in place (``unb_region_bootstrap_fold_dev``).
    ``compute_rounds(u, unormed, selected, lo, hi, minvol) -> (maxd_r, f_r, active, failure)``
    replaces the device path (the CPU tests inject the oracle here; gloo has no device buffer).
    Any exception on one rank still reaches the collective (failed flag), so no rank is left
    waiting; it is re-raised afterwards on the rank that had it, ``LinAlgError`` on the others.
    """
    import torch
    dist = _dist()
    world, me = world_size(), rank()
    if masks == "broadcast":
        selected = broadcast_array(np.asarray(selected, dtype=np.uint8)).astype(bool)
    elif masks != "replicated":
        raise ValueError("masks must be 'replicated' or 'broadcast'")
    selected = np.asarray(selected, dtype=bool)
    nrounds = selected.shape[0]
    lo, hi = shard_bounds(nrounds, world, me)
    tag = _mask_tag(selected)
    dev = _device()
    on_device = compute_rounds is None and dev.type == "cuda"
    local_exc = None
    message = None
    buf = None
    try:
        screened = None
        if on_device and minvol == 0 and hi > lo:
            # the enlargement screen (mlfriends._bootstrap_rounds_screened) applies to a rank's
            # slice as well: the global maximum is some rank's local maximum, and that one is
            # computed with the reference's exact algebra.  It needs the screened f on the host to
            # pick its candidates, so this rank's three numbers travel to the device as 40 bytes.
            from .mlfriends import _bootstrap_rounds_screened
            active = ~(selected.all(axis=1) | ~selected.any(axis=1))
            screened = _bootstrap_rounds_screened(u, unormed, selected, lo, hi, active)
        if screened is not None:
            maxd_r, f_r, active, _ = screened
            rounds = [r for r in range(lo, hi) if active[r]]
            buf = torch.tensor([max(maxd_r[r] for r in rounds), max(f_r[r] for r in rounds), 0.0,
                                tag, -tag], dtype=torch.float64, device=dev)
        elif on_device:
            from .mlfriends import _engine
            ctrs, invcovs, stop, failure = _host_rounds(u, selected, lo, hi, minvol)
            if failure is not None:
                local_exc, message = failure, str(failure)
            buf = torch.empty(5, dtype=torch.float64, device=dev)
            # the fold kernel must precede the collective in stream order: torch's current stream
            # (handle 0 = the legacy default stream, spelled cudaStreamLegacy = 1 for the library,
            # whose own streams do not synchronise with it)
            stream = torch.cuda.current_stream(dev).cuda_stream or 1
            _engine().region_bootstrap_fold_dev(unormed, selected, u, ctrs, invcovs, lo, stop,
                                                failure is not None, tag, buf.data_ptr(), stream)
        else:
            if compute_rounds is None:
                from .mlfriends import _bootstrap_rounds as compute_rounds
            maxd, maxf, failed = 0.0, 0.0, 0.0
            if hi > lo:
                maxd_r, f_r, active, failure = compute_rounds(u, unormed, selected, lo, hi, minvol)
                if failure is not None:
                    failed, local_exc, message = 1.0, failure[1], str(failure[1])
                for r in range(lo, hi):
                    if not active[r]:
                        continue
                    if failure is not None and r >= failure[0]:
                        break
                    maxd = max(maxd, maxd_r[r])
                    f = f_r[r]
                    if not np.isfinite(f) or not f > 0:
                        failed, message = 1.0, "Distances are not positive"
                        break
                    maxf = max(maxf, f)
            buf = torch.tensor([maxd, maxf, failed, tag, -tag], dtype=torch.float64, device=dev)
    except Exception as exc:  # noqa: BLE001 -- every rank must still reach the collective
        local_exc, message = exc, str(exc)
        buf = torch.tensor([0.0, 0.0, 1.0, tag, -tag], dtype=torch.float64, device=dev)
    # NaN under MAX is implementation-defined in NCCL, so failure travels as its own flag
    # (the reference ships NaN through gather/bcast, integrator.py:391-393, 406-411)
    if timings is None:
        timings = last_timings
    if dev.type == "cuda":
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=_group)
        e1.record()
        r2, f, anyfail, tag_hi, tag_lo = buf.cpu().numpy()
        timings["collective_us"] = 1e3 * e0.elapsed_time(e1)
    else:
        dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=_group)
        r2, f, anyfail, tag_hi, tag_lo = buf.cpu().numpy()
    if tag_hi != -tag_lo:
        raise RuntimeError("the ranks hold different bootstrap selection masks (not replicas of "
                           "one seeded stream); use masks='broadcast'")
    if local_exc is not None:
        raise local_exc
    if anyfail > 0:
        raise np.linalg.LinAlgError(message or "compute_enlargement failed on another rank")
    assert r2 > 0, (r2, u, unormed)
    assert f > 0, (f, u, unormed)
    return float(r2), float(f)


def allgather_varrows(local):
    """Concatenate per-rank row blocks of DIFFERENT lengths in rank order (two collectives: the
    counts, then the rows padded to the longest block).  Every rank gets the same array."""
    import torch
    dist = _dist()
    world = dist.get_world_size(_group)
    local = np.ascontiguousarray(local)
    dev = _device()
    n_local = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n_local, group=_group)
    counts = [int(c.item()) for c in counts]
    longest = max(max(counts), 1)
    buf = np.zeros((longest,) + local.shape[1:], dtype=local.dtype)
    buf[:len(local)] = local
    send = torch.from_numpy(buf).to(dev)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=_group)
    return np.concatenate([recv[r].cpu().numpy()[:counts[r]] for r in range(world)], axis=0)


def sharded_sample_device(region, nsamples, loglike=None, Lmin=None, method=None, seed=None):
    """Throughput-mode proposals (``MLFriends.sample_device``: drawn on the device, NOT the
    reference's random stream) with the DRAWS split over the ranks: every rank draws
    ``nsamples / world`` proposals from its own range of the generator's counter space, filters
    them on its GPU, and the accepted rows are re-united in rank order -- the reference's MPI mode
    does the same with its per-rank streams (gather + bcast of the accepted points,
    integrator.py:1916-1928).  Every rank returns the same ``rows`` (and ``logl``), so the ranks
    stay replicas of one sampler while the proposal work is divided by the world size."""
    world, me = world_size(), rank()
    lo, hi = shard_bounds(int(nsamples), world, me)
    out = region.sample_device(hi - lo, method=method, seed=seed, loglike=loglike, Lmin=Lmin)
    rows, like = out if loglike is not None else (out, None)
    if world == 1:
        return out
    rows = allgather_varrows(rows)
    if like is not None:
        return rows, allgather_varrows(like)
    return rows


def sharded_inside(region, pts):
    """``region.inside(pts)`` with the rows split over the ranks; every rank gets the full mask."""
    world, me = world_size(), rank()
    lo, hi = shard_bounds(len(pts), world, me)
    local = region.inside(pts[lo:hi]) if hi > lo else np.zeros(0, dtype=bool)
    if world == 1:
        return local
    return allgather_rows(local, len(pts))
