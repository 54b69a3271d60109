#!/bin/bash
# final single-GPU evidence pass of a round: tests, smoke, bench (both arms), then the ncu captures, one after the other.
# usage: tools/gpu_final.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=400 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref exit $?"
python tools/stage_times.py > gpurun_out/${TAG}_stage_times.json 2>&1; echo "stage_times exit $?"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-coclust --no-parity"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "list exit $?"
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:tc_cor2_kernelILi2 -s 1 -c 1 -o gpurun_out/${TAG}_prof_tc -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "gemm exit $?"
CMD2="python tools/stage_times.py --passes 2"
$CMD2 > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:probe_blocked|prepare_blocked|mutual_kernel|prune_kernel|tc_sample_hist|scatter_candidates|topk_select|feat_stats|feat_rows" -s 9 -c 9 -o gpurun_out/${TAG}_stages -f $CMD2 > gpurun_out/${TAG}_ncu_stages.log 2>&1; echo "stages exit $?"
CMD3="python tools/c3_cover.py 500 2"
$CMD3 > gpurun_out/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:graph_cover_batch" -s 1 -c 1 -o gpurun_out/${TAG}_cover -f $CMD3 > gpurun_out/${TAG}_ncu_cover.log 2>&1; echo "cover exit $?"
tail -3 gpurun_out/${TAG}_pytest_gpu.log; tail -1 gpurun_out/${TAG}_smoke.log
