#!/bin/bash
# one ncu --set full capture of a kernel_bench case: gpu_ncu.sh <tag> <kernel regex> <--only value> [extra kernel_bench args]
set -u
mkdir -p gpurun_out
TAG=$1; RX=$2; ONLY=$3; shift 3
KB="python scripts/kernel_bench.py --impls auto --iters 3 --only $ONLY $*"
$KB > gpurun_out/prof_${TAG}.plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$RX -s 3 -c 1 -f -o gpurun_out/prof_${TAG} $KB > gpurun_out/prof_${TAG}.ncu.log 2>&1
echo "ncu $TAG rc=$?"
tail -3 gpurun_out/prof_${TAG}.ncu.log
