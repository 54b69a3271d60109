#!/bin/bash
# parity tests, then the bench at a reduced and at the full C2 size
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s --timeout=300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --cells 20000 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_20k.log 2>&1
echo "bench20k exit $?" >> gpurun_out/bench_20k.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c2.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_c2.log
tail -4 gpurun_out/pytest_gpu.log; tail -c 1500 gpurun_out/bench_20k.log; tail -c 3000 gpurun_out/bench_c2.log
