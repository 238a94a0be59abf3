"""Multi-GPU sharding of the region path: one process per GPU (``torchrun``), replicated region
state, and exactly one collective per region rebuild.

Where the path shards:

* **proposals** (``inside`` / ``find_nearby`` / ellipsoid / loglike) are independent rows: each
  rank takes a contiguous slice, no data-path collective; :func:`allgather_rows` re-unites masks
  when a caller needs the full vector.
* **bootstrap rounds** are independent: rank ``r`` evaluates rounds ``[lo_r, hi_r)`` and ONE
  ``all_reduce(MAX)`` over a 3-double buffer ``[r2, f, failure]`` replaces the reference's
  pickled ``gather`` + ``bcast`` (``integrator.py:395-404``).  For parity with the
  single-process oracle the selection masks of ALL rounds come from rank 0's host stream
  (the reference's MPI mode re-seeds every rank, ``integrator.py:1239-1251``, and is therefore
  not comparable with its own 1-process run).

Backend: ``nccl`` on GPUs (NVLink 5 / NVSwitch), ``gloo`` in the CPU-only tests.  The payload is
24 bytes, so the collective is latency-bound; there is nothing to overlap or fuse.
"""
import numpy as np

_enabled = False
_group = None


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


def reduce_enlargement(u, unormed, selected, minvol=0., compute_rounds=None):
    """Sharded ``compute_enlargement``: masks from rank 0, this rank's slice of rounds on its GPU,
    one ``all_reduce(MAX)`` of ``[r2, f, failed]``.  Returns ``(r2, f)``, identical on all ranks
    and identical to the single-process result (max is order-independent).

    ``compute_rounds(u, unormed, selected, lo, hi, minvol) -> (maxd_r, f_r, active, failure)``
    defaults to the device implementation; the CPU tests inject the oracle here.
    """
    if compute_rounds is None:
        from .mlfriends import _bootstrap_rounds as compute_rounds
    world, me = world_size(), rank()
    selected = broadcast_array(np.asarray(selected, dtype=np.uint8)).astype(bool)
    nrounds = selected.shape[0]
    lo, hi = shard_bounds(nrounds, world, me)
    maxd, maxf, failed = 0.0, 0.0, 0.0
    message = None
    if hi > lo:
        maxd_r, f_r, active, failure = compute_rounds(u, unormed, selected, lo, hi, minvol)
        if failure is not None:
            failed, message = 1.0, str(failure[1])
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
    # NaN under MAX is implementation-defined in NCCL, so failure travels as its own flag
    # (the reference ships NaN through gather/bcast, integrator.py:391-393, 406-411)
    r2, f, anyfail = allreduce_max([maxd, maxf, failed])
    if anyfail > 0:
        raise np.linalg.LinAlgError(message or "compute_enlargement failed on another rank")
    assert r2 > 0, (r2, u, unormed)
    assert f > 0, (f, u, unormed)
    return float(r2), float(f)


def sharded_inside(region, pts):
    """``region.inside(pts)`` with the rows split over the ranks; every rank gets the full mask."""
    world, me = world_size(), rank()
    lo, hi = shard_bounds(len(pts), world, me)
    local = region.inside(pts[lo:hi]) if hi > lo else np.zeros(0, dtype=bool)
    if world == 1:
        return local
    return allgather_rows(local, len(pts))
