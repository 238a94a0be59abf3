#!/usr/bin/env python
"""cProfile of one arm of tests/run_compare.py (default: the drop-in arm) -- where a whole
ReactiveNestedSampler run spends its time once the region is on the device.

    python tests/profile_run.py [reference|ours|device]
"""
import cProfile
import io
import os
import pstats
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import run_compare
pr = cProfile.Profile()
pr.enable()
run_compare.arm(sys.argv[1] if len(sys.argv) > 1 else 'ours', 20, 4000, 8000, 0.05)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue()[:9000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(25)
print(s.getvalue()[:6000])
