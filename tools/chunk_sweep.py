"""e2e step time of the host-buffer call (pinned buffers) as a function of the pipeline chunk."""
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
from ultranest_b200.likelihoods import GaussianLogLike  # noqa: E402

M = 1 << 20
eng = _native.get_engine()
region = bench.build_region(m, bench.make_live())
cand = bench.make_candidates(region, M, 3)
kind, lparams = GaussianLogLike(0.5, bench.SIGMA).device_spec(bench.NDIM)
region._bind()
pin = torch.empty((M, bench.NDIM), dtype=torch.float64).pin_memory(); pin.numpy()[...] = cand
pm = torch.empty(M, dtype=torch.uint8).pin_memory(); pl = torch.empty(M, dtype=torch.float64).pin_memory()
np_pts, np_mask, np_like = pin.numpy(), pm.numpy().view(bool), pl.numpy()
for chunk in (1 << 15, 1 << 16, 1 << 17, 3 << 16, 1 << 18, 1 << 19, 1 << 20):
    eng.set_option(_native.OPT_CHUNK_ROWS, chunk)
    for _ in range(3):
        eng.region_inside_loglike(np_pts, kind, lparams, mask_out=np_mask, like_out=np_like)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(10):
            eng.region_inside_loglike(np_pts, kind, lparams, mask_out=np_mask, like_out=np_like)
        ts.append((time.perf_counter() - t0) / 10 * 1e3)
    print(json.dumps({"chunk_rows": chunk, "ms_per_step_min": min(ts), "ms_per_step_median": float(np.median(ts))}))
