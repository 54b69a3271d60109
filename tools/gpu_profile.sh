#!/bin/bash
# ncu evidence for the bench command, one ncu pass per gpurun call:
#   tools/gpu_profile.sh <tag> list   launch list (per-kernel durations -> shares of the step)
#   tools/gpu_profile.sh <tag> full   one --set full capture of the dominant kernel (tc_cor2_kernel<SYM>)
#   tools/gpu_profile.sh <tag> cover  one --set full capture of the graph-cover kernel (co-clustering)
TAG=${1:-r1}
MODE=${2:-list}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-coclust"
if [ "$MODE" = "list" ]; then
  $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
  echo "list exit $?"
elif [ "$MODE" = "full" ]; then
  $CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:tc_cor2_kernelILi2 -s 1 -c 1 \
      -o gpurun_out/${TAG}_prof_tc -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
  echo "full exit $?"
else
  CMD2="python tools/coclust_small.py"
  $CMD2 > gpurun_out/${TAG}_plain3.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:graph_cover -c 1 \
      -o gpurun_out/${TAG}_prof_cover -f $CMD2 > gpurun_out/${TAG}_ncu_cover.log 2>&1
  echo "cover exit $?"
fi
ls -la gpurun_out | tail -6
