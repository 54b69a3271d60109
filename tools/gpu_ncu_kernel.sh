#!/bin/bash
# one --set full ncu pass over kernels of the bench step matching a regex.
# usage: tools/gpu_ncu_kernel.sh <tag> <kernel regex> [skip] [count]
TAG=$1; KRE=$2; SKIP=${3:-1}; CNT=${4:-1}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-coclust"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$KRE" -s $SKIP -c $CNT \
    -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/${TAG}_ncu.log
