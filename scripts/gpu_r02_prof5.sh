#!/bin/bash
# ncu captures at config 5 (K = 200, column-window kernels), 1 GPU.  Every command first runs plain (exit code checked).
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TAG=r02c5
KB="python scripts/kernel_bench.py --impls auto --iters 3 --K 200"
run_ncu () {  # name, kernel regex, --only value, skip count
  $KB --only $3 > $OUT/prof_${TAG}_$1.plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -f -o $OUT/prof_${TAG}_$1 $KB --only $3 > $OUT/prof_${TAG}_$1.ncu.log 2>&1
  echo "ncu $1 rc=$?" | tee -a $OUT/prof_${TAG}_summary.log
}
rm -f $OUT/prof_${TAG}_summary.log
run_ncu gradplain grad_kernel grad_contig 3
run_ncu synth synth_kernel synth 3
BENCH="python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-cudnn-autotune"
$BENCH > $OUT/prof_${TAG}_bench.plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_${TAG}.csv $BENCH > $OUT/prof_${TAG}_bench.ncu.log 2>&1
echo "ncu launches rc=$?" | tee -a $OUT/prof_${TAG}_summary.log
ls -la $OUT/*${TAG}* | tee -a $OUT/prof_${TAG}_summary.log
