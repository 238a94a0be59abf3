"""A few steps of the device-resident hot path for an ncu launch list / full capture:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/launch_list.py
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ultranest_b200 import _native  # noqa: E402
from ultranest_b200 import mlfriends as m  # noqa: E402
from ultranest_b200.likelihoods import GaussianLogLike  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = _native.get_engine()
u = bench.make_live()
region = bench.build_region(m, u)
cand = bench.make_candidates(region, M, 3)
kind, lparams = GaussianLogLike(0.5, bench.SIGMA).device_spec(bench.NDIM)
region._bind()
pts = torch.from_numpy(cand).cuda()
mask = torch.empty(M, dtype=torch.uint8, device="cuda")
like = torch.empty(M, dtype=torch.float64, device="cuda")
lp = _native.as_f64(lparams)
for i in range(steps):
    eng.call("unb_region_inside_loglike_dev", pts.data_ptr(), M, mask.data_ptr(), like.data_ptr(),
             int(kind), lp.ctypes.data if i == 0 else None, None)
eng.synchronize()
rows, lk = region.sample_device(M, method="sample_from_wrapping_ellipsoid", seed=1,
                                loglike=GaussianLogLike(0.5, bench.SIGMA), Lmin=float(np.median(like.cpu().numpy())))
print("accepted", int(mask.sum().item()), "sampled", len(rows))
