#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_coclust.py -m gpu -q -x --timeout=200 2>&1 | tail -3
MCB_COVER_KERNEL=slow timeout 300 python -m pytest tests/test_gpu_coclust.py -m gpu -q -x --timeout=200 2>&1 | tail -2
timeout 300 python - <<'PY'
import sys, json
sys.path.insert(0, '.')
import numpy as np
import bench
from metacell_b200 import _lib
class A: seed=1234
ctx=_lib.Context(0)
r=bench.bench_coclust(ctx, A)
print("coclust resamples/s", round(r["value"],1), "ms", round(r["ms"],1), r["stages_ms_rank0"], r.get("config"))
PY
