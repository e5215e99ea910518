#!/bin/bash
# round 2, third session: GPU tests of sadil_updated / learn_coding_vectors; tile-size knob on the column-window fused step
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests/test_adil_gpu.py -m gpu -q -k "sadil or coder_on_a_fixed or full_batch" > $OUT/n_pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/n_summary.log
tail -4 $OUT/n_pytest.log | tee -a $OUT/n_summary.log
for cfg in "136 48" "136 32" "256 16" "176 64" "176 48" "176 32"; do
  set -- $cfg
  echo "== K=$1 ADIL_GRAD_MAX_TP=$2" | tee -a $OUT/n_summary.log
  ADIL_GRAD_MAX_TP=$2 python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig --iters 20 --K $1 2>&1 | grep -E "^auto|rror" | tee -a $OUT/n_summary.log
done
