"""world_size=2 gloo tests (CPU) of the multi-GPU host logic: round slicing, the single
allreduce(MAX) with its failure flag and mask tag, the optional mask broadcast, exception safety
(no rank may be left waiting in the collective), row all-gather.  The per-rank round evaluation is
injected from the oracle, so no GPU is needed; the result must equal the single-process oracle.
The CUDA path under NCCL is covered by tests/test_gpu_distributed_nccl.py."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from oracle import cport


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_rounds(u, unormed, selected, lo, hi, minvol):
    nrounds = selected.shape[0]
    active = ~(selected.all(axis=1) | ~selected.any(axis=1))
    maxd, f = np.zeros(nrounds), np.zeros(nrounds)
    for r in range(lo, hi):
        if not active[r]:
            continue
        maxd[r] = cport.maxradiussq_selected(unormed, selected[r])
        ctr, cov = cport.bounding_ellipsoid(u[selected[r]])
        f[r] = cport.enlargement_f(u, selected[r], ctr, np.linalg.inv(cov))
    return maxd, f, active, None


def _failing_rounds(u, unormed, selected, lo, hi, minvol):
    maxd, f, active, _ = _oracle_rounds(u, unormed, selected, lo, hi, minvol)
    if lo > 0:   # only the second rank fails
        return maxd, f, active, (lo, np.linalg.LinAlgError("singular matrix"))
    return maxd, f, active, None


def _raising_rounds(u, unormed, selected, lo, hi, minvol):
    if lo > 0:   # an unexpected error (not a numerical failure) on the second rank only
        raise MemoryError("simulated device allocation failure")
    return _oracle_rounds(u, unormed, selected, lo, hi, minvol)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ultranest_b200 import distributed as D
    D.enable()
    try:
        assert D.world_size() == world and D.rank() == rank
        rng = np.random.RandomState(3)
        u = rng.uniform(0.3, 0.7, size=(300, 4))
        ctr = u.mean(axis=0)
        w, v = np.linalg.eigh(np.cov(u, rowvar=0) * 6)
        unormed = np.dot(u - ctr, v * w**-0.5)
        # rank 1 deliberately holds DIFFERENT masks: rank 0's must win
        sel = np.zeros((7, 300), dtype=bool)
        srng = np.random.RandomState(10 + rank)
        for r in range(7):
            sel[r, srng.randint(300, size=300)] = True
        r2, f = D.reduce_enlargement(u, unormed, sel, compute_rounds=_oracle_rounds, masks="broadcast")
        out.put(("enl", rank, r2, f))
        try:
            D.reduce_enlargement(u, unormed, sel, compute_rounds=_failing_rounds, masks="broadcast")
            out.put(("fail", rank, "no error"))
        except np.linalg.LinAlgError:
            out.put(("fail", rank, "raised"))
        # replicated masks (the default): different masks are detected by the tag pair that rides
        # in the same reduction ...
        try:
            D.reduce_enlargement(u, unormed, sel, compute_rounds=_oracle_rounds)
            out.put(("tag", rank, "no error"))
        except RuntimeError:
            out.put(("tag", rank, "raised"))
        # ... identical masks give the single-process result with ONE collective
        sel0 = np.zeros((7, 300), dtype=bool)
        srng0 = np.random.RandomState(10)
        for r in range(7):
            sel0[r, srng0.randint(300, size=300)] = True
        r2, f = D.reduce_enlargement(u, unormed, sel0, compute_rounds=_oracle_rounds)
        out.put(("enl", rank, r2, f))
        # an exception on one rank must not leave the other waiting in the collective
        try:
            D.reduce_enlargement(u, unormed, sel0, compute_rounds=_raising_rounds)
            out.put(("exc", rank, "no error"))
        except MemoryError:
            out.put(("exc", rank, "MemoryError"))
        except np.linalg.LinAlgError:
            out.put(("exc", rank, "LinAlgError"))
        rows = np.arange(11 * 3, dtype=np.float64).reshape(11, 3)
        lo, hi = D.shard_bounds(11, world, rank)
        got = D.allgather_rows(rows[lo:hi], 11)
        mask = (np.arange(11) % 3 == 0)
        gotm = D.allgather_rows(mask[lo:hi], 11)
        ragged = np.arange((rank + 2) * 3, dtype=np.float64).reshape(rank + 2, 3) + 100 * rank
        gotr = D.allgather_varrows(ragged)
        wantr = np.concatenate([np.arange((r + 2) * 3, dtype=np.float64).reshape(r + 2, 3) + 100 * r
                                for r in range(world)])
        empty = D.allgather_varrows(np.zeros((0, 3)) if rank == 0 else ragged)
        out.put(("gather", rank, bool((got == rows).all() and (gotm == mask).all() and gotm.dtype == bool
                                      and (gotr == wantr).all() and len(empty) == 3)))
    finally:
        D.disable()
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    from ultranest_b200.distributed import shard_bounds
    for n in range(0, 40):
        for world in range(1, 9):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_rank_bootstrap_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=100) for _ in range(6 * world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    # single-process oracle with rank 0's masks
    rng = np.random.RandomState(3)
    u = rng.uniform(0.3, 0.7, size=(300, 4))
    ctr = u.mean(axis=0)
    w, v = np.linalg.eigh(np.cov(u, rowvar=0) * 6)
    unormed = np.dot(u - ctr, v * w**-0.5)
    sel = np.zeros((7, 300), dtype=bool)
    srng = np.random.RandomState(10)
    for r in range(7):
        sel[r, srng.randint(300, size=300)] = True
    maxd, f, active, _ = _oracle_rounds(u, unormed, sel, 0, 7, 0.)
    want = (maxd.max(), f.max())
    enl = [r for r in results if r[0] == "enl"]
    assert len(enl) == 4 and all((r[2], r[3]) == want for r in enl)
    assert sorted(r[2] for r in results if r[0] == "fail") == ["raised", "raised"]
    assert sorted(r[2] for r in results if r[0] == "tag") == ["raised", "raised"]
    assert sorted((r[1], r[2]) for r in results if r[0] == "exc") == [(0, "LinAlgError"), (1, "MemoryError")]
    assert all(r[2] for r in results if r[0] == "gather")
