"""torchrun check of the multi-GPU path on real GPUs: sharded bootstrap == single-process result,
sharded inside() == local inside().   torchrun --nproc-per-node 2 tools/dist_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ultranest_b200 import distributed as D
    from ultranest_b200 import mlfriends as m
    import bench
    u = bench.make_live(2000, 10, seed=1)
    layer = m.AffineLayer()
    layer.optimize(u, u)
    region = m.MLFriends(u, layer)
    single = region.compute_enlargement(nbootstraps=30, rng=np.random.RandomState(2))
    D.enable()
    sharded = region.compute_enlargement(nbootstraps=30, rng=np.random.RandomState(2))
    assert sharded == single, (sharded, single)
    region.maxradiussq, region.enlarge = single
    region.create_ellipsoid()
    pts = bench.make_candidates(region, 50000, 7)
    full = region.inside(pts)
    got = D.sharded_inside(region, pts)
    assert (got == full).all()
    print("rank %d/%d ok: r2=%.9g f=%.9g inside=%d" % (D.rank(), D.world_size(), single[0], single[1], full.sum()))
    D.disable()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
