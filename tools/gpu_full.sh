#!/bin/bash
# full GPU validation: all gpu tests, smoke, full bench, then ncu evidence.  usage: tools/gpu_full.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/${TAG}_smoke.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.log 2>&1; echo "ref exit $?"
bash tools/gpu_profile.sh ${TAG} list
tail -3 gpurun_out/${TAG}_pytest_gpu.log; tail -2 gpurun_out/${TAG}_smoke.log
