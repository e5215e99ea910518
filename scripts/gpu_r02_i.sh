#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
KB="python scripts/kernel_bench.py --impls auto --only grad_dict_step_partials,grad_partials --iters 15"
for v in "base" "ADIL_GRAD_MMA_ORDER=1" "ADIL_GRAD_MAX_TP=48" "ADIL_GRAD_MMA_ORDER=1 ADIL_GRAD_MAX_TP=48"; do
  if [ "$v" = "base" ]; then $KB > $OUT/i_kb.log 2>&1; else env $v $KB > $OUT/i_kb.log 2>&1; fi
  echo "== $v" | tee -a $OUT/i_summary.log; grep -E "^auto" $OUT/i_kb.log | tee -a $OUT/i_summary.log
done
for K in 64 100; do
  for v in "base" "ADIL_GRAD_MMA_ORDER=1"; do
    if [ "$v" = "base" ]; then $KB --K $K > $OUT/i_kb.log 2>&1; else env $v $KB --K $K > $OUT/i_kb.log 2>&1; fi
    echo "== K=$K $v" | tee -a $OUT/i_summary.log; grep -E "^auto" $OUT/i_kb.log | tee -a $OUT/i_summary.log
  done
done
ADIL_GRAD_MMA_ORDER=1 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "grad or fused or host_index or random_shapes or partial" > $OUT/i_pytest_order1.log 2>&1; echo "pytest order1 rc=$?" | tee -a $OUT/i_summary.log
tail -3 $OUT/i_pytest_order1.log | tee -a $OUT/i_summary.log
python -m pytest tests/test_adil_gpu.py -m gpu -q -k "whole_set or sadil" > $OUT/i_pytest_new.log 2>&1; echo "pytest new rc=$?" | tee -a $OUT/i_summary.log
tail -3 $OUT/i_pytest_new.log | tee -a $OUT/i_summary.log
