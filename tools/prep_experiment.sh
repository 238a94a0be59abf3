#!/bin/bash
# A/B of the prep kernel's occupancy target (run on the GPU box; nvcc is in the image)
for minb in 1 6 8 10; do
  touch ultranest_b200/csrc/unb_region.cu
  python - <<PY
from ultranest_b200 import build
build.build(extra_flags=["-DUNB_PREP_MINB=$minb"])
PY
  echo "MINB=$minb"
  python tools/microbench.py --quick 2>&1 | grep inside_fused | cut -c1-120
done
