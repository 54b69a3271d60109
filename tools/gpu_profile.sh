#!/bin/bash
# ncu evidence for the bench command: launch list (shares) + one full capture of the dominant kernel.
# usage: tools/gpu_profile.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-coclust"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "list exit $?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:tc_cor2_kernelILi2 -s 1 -c 1 \
    -o gpurun_out/${TAG}_prof_tc -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full exit $?"
ls -la gpurun_out | tail -8
tail -c 600 gpurun_out/${TAG}_plain.log
