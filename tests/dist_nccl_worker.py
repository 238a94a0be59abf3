"""torchrun worker of tests/test_gpu_distributed_nccl.py: the multi-GPU region path on REAL GPUs
under NCCL, checked against the single-process CUDA result and the CPU oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_nccl_worker.py

Every rank prints one JSON line; a failed check raises (non-zero exit of the launcher)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def multimodal_live(n=2000, d=10, seed=5):
    """Eggbox-like live set (BASELINE configs[2] shape): points around a 5^d lattice of modes."""
    rng = np.random.RandomState(seed)
    centres = rng.randint(0, 5, size=(n, d)) * 0.2 + 0.1
    return centres + rng.normal(size=(n, d)) * 0.01


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import cport
    from ultranest_b200 import distributed as D
    from ultranest_b200 import mlfriends as m
    import bench
    me, world = dist.get_rank(), dist.get_world_size()
    report = {"rank": me, "world": world, "cases": []}
    for name, u, nboot in (("configs[2] eggbox-like N=2000 d=10", multimodal_live(), 30),
                           ("configs[1] N=4000 d=20", bench.make_live(), 30),
                           ("odd split N=700 d=5, 7 rounds", bench.make_live(700, 5, seed=4), 7)):
        layer = m.AffineLayer()
        layer.optimize(u, u)
        region = m.MLFriends(u, layer)
        single = region.compute_enlargement(nbootstraps=nboot, rng=np.random.RandomState(2))
        want = cport.compute_enlargement(u, region.unormed, nboot, np.random.RandomState(2))
        assert single == want, ("single-process CUDA vs oracle", name, single, want)
        D.enable()
        try:
            sharded = region.compute_enlargement(nbootstraps=nboot, rng=np.random.RandomState(2))
            coll_us = D.last_timings.get("collective_us")
            assert sharded == single, ("sharded vs single", name, sharded, single)
            # ranks that are NOT replicas of one stream are detected inside the same collective
            try:
                region.compute_enlargement(nbootstraps=nboot, rng=np.random.RandomState(100 + me))
                raise AssertionError("mask mismatch went unnoticed")
            except RuntimeError as exc:
                assert "different bootstrap selection masks" in str(exc)
            region.maxradiussq, region.enlarge = single
            region.create_ellipsoid()
            pts = bench.make_candidates(region, 40000, 7)
            pts[::5] = np.random.RandomState(8).uniform(0.05, 0.95, size=pts[::5].shape)
            full = region.inside(pts)
            got = D.sharded_inside(region, pts)
            assert (got == full).all(), ("sharded_inside", name)
            # throughput mode with the draws split over the ranks: every rank ends with the same
            # accepted rows, each rank's block comes from its own counter range and is a member
            region.device_seed = 77
            rows = D.sharded_sample_device(region, 40000, method="sample_from_wrapping_ellipsoid")
            digest = float(rows.sum())
            same = torch.tensor([digest, -digest], dtype=torch.float64, device="cuda")
            dist.all_reduce(same, op=dist.ReduceOp.MAX)
            assert same[0].item() == -same[1].item(), "ranks disagree on the gathered rows"
            assert len(rows) > 1000 and region.inside(rows).all()
            assert len(np.unique(rows[:, 0])) == len(rows), "ranks drew overlapping proposals"
            sub = slice(0, 6000)
            want_mask = cport.region_inside(
                pts[sub], region.unormed, lambda p: cport.transform_affine(p, layer.ctr, layer.T),
                region.maxradiussq, region.ellipsoid_center, region.ellipsoid_invcov, region.enlarge)
            assert (got[sub] == want_mask).all(), ("sharded_inside vs oracle", name)
        finally:
            D.disable()
        report["cases"].append({"case": name, "r2": single[0], "f": single[1],
                                "collective_us": coll_us, "inside_accept": float(full.mean())})
    report["status"] = "ok"
    sys.stdout.write("\n" + json.dumps(report) + "\n")
    sys.stdout.flush()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
