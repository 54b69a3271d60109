#!/bin/bash
# parity of every cover kernel variant, then latency vs concurrency for the main ones
timeout 600 python -m pytest tests/test_gpu_coclust.py -m gpu -q -x --timeout=300 2>&1 | tail -3
for k in auto warp16 fast; do
  echo "== latency MCB_COVER_KERNEL=$k"
  MCB_COVER_KERNEL=$k timeout 300 python tools/cover_latency.py 2>&1 | tail -8
done
