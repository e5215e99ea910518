#!/bin/bash
# round 2, third session: two alternating dv accumulators in the single-window kernels too? (ADIL_GRAD_DV2=1: column windows
# only, =2: wherever tensor memory has the columns); GPU tests of the coder on a fixed dictionary
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests/test_adil_gpu.py -m gpu -q -k "coder_on_a_fixed or full_batch or sadil" > $OUT/l_pytest_lcv.log 2>&1; echo "pytest lcv rc=$?" | tee $OUT/l_summary.log
tail -5 $OUT/l_pytest_lcv.log | tee -a $OUT/l_summary.log
ADIL_GRAD_DV2=2 python -m pytest tests/test_kernels_gpu.py -m gpu -q > $OUT/l_pytest.log 2>&1; echo "pytest dv2=2 rc=$?" | tee -a $OUT/l_summary.log
tail -3 $OUT/l_pytest.log | tee -a $OUT/l_summary.log
for K in 50 64 100 128; do
  for dv in 1 2; do
    echo "== K=$K ADIL_GRAD_DV2=$dv" | tee -a $OUT/l_summary.log
    ADIL_GRAD_DV2=$dv python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig,grad_contig --iters 20 --K $K 2>&1 | grep -E "^auto|rror" | tee -a $OUT/l_summary.log
  done
done
