#!/bin/bash
# round 2, third session: tile-size probe of the synthesis kernel (ADIL_SYNTH_MAX_TP)
set -u
mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/p_summary.log
for cfg in "50 64" "50 48" "50 32" "64 48" "100 32" "100 16" "200 16"; do
  set -- $cfg
  echo "== synth K=$1 ADIL_SYNTH_MAX_TP=$2" | tee -a $OUT/p_summary.log
  ADIL_SYNTH_MAX_TP=$2 python scripts/kernel_bench.py --impls auto --only synth --iters 20 --K $1 2>&1 | grep -E "^auto|rror" | tee -a $OUT/p_summary.log
done
