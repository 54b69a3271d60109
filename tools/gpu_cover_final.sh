#!/bin/bash
# single-GPU evidence pass after a change of the cover kernel: all GPU tests, smoke, the bench line, one ncu capture of the
# cover kernel on C3 (500 resamples).  usage: tools/gpu_cover_final.sh <tag>
TAG=${1:-r2x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=400 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
CMD3="python tools/c3_cover.py 500 1"
$CMD3 > gpurun_out/${TAG}_plain3.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:graph_cover_incr" -c 1 -o gpurun_out/${TAG}_cover -f $CMD3 > gpurun_out/${TAG}_ncu_cover.log 2>&1; echo "cover exit $?"
tail -3 gpurun_out/${TAG}_pytest_gpu.log; tail -1 gpurun_out/${TAG}_smoke.log; tail -1 gpurun_out/${TAG}_plain3.log
