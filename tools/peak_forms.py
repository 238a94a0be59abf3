#!/usr/bin/env python
"""FMA issue rates of this device by instruction form (the compute-roofline denominators)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultranest_b200 import _native  # noqa: E402

eng = _native.get_engine()
names = {0: "FFMA (scalar)", 1: "FFMA2 (packed pairs)", 2: "FFMA2 (scalar multiplicand)"}
out = {"lane_fma_per_s": {names[f]: eng.fp32_peak_form(f) for f in (0, 1, 2)},
       "fp32_peak": eng.fp32_peak(), "fp64_peak": eng.fp64_peak()}
print(json.dumps(out))
