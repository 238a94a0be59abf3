"""One membership launch in the FULL-SCAN regime (no proposal has a neighbour) for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:k_inside_any32 -s 1 -c 1 -o out python tools/ncu_fullscan.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ultranest_b200 import _native  # noqa: E402
from ultranest_b200 import mlfriends as m  # noqa: E402

M = 1 << 20
eng = _native.get_engine()
region = bench.build_region(m, bench.make_live())
region._bind()
rng = np.random.RandomState(4)
r = region.maxradiussq**0.5
tbox = np.ascontiguousarray(rng.uniform(region.bbox_lo - r, region.bbox_hi + r, size=(M, bench.NDIM)))
t_dev = torch.from_numpy(tbox).cuda()
mask = torch.empty(M, dtype=torch.uint8, device="cuda")
for _ in range(3):
    eng.call("unb_region_find_nearby_dev", t_dev.data_ptr(), M, None, mask.data_ptr(), None)
eng.synchronize()
print("accepted", int(mask.sum().item()), "tile units", eng.stat(_native.STAT_TILE_VISITS))
