#!/bin/bash
# round 2, third session: both column windows of K > 128 in one launch (A/B against the two serial launches), K = 64 knobs
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > $OUT/j_pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/j_summary.log
tail -3 $OUT/j_pytest.log | tee -a $OUT/j_summary.log
for K in 200 256 136; do
  for mode in serial paired; do
    echo "== K=$K windows=$mode" | tee -a $OUT/j_summary.log
    ADIL_GRAD_WINDOWS=$mode python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig,grad_contig --iters 20 --K $K 2>&1 | grep -E "^auto|rror" | tee -a $OUT/j_summary.log
  done
done
echo "== K=64 default" | tee -a $OUT/j_summary.log
python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig --iters 20 --K 64 2>&1 | grep -E "^auto|rror" | tee -a $OUT/j_summary.log
echo "== K=64 ADIL_GRAD_NPF=3" | tee -a $OUT/j_summary.log
ADIL_GRAD_NPF=3 python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig --iters 20 --K 64 2>&1 | grep -E "^auto|rror" | tee -a $OUT/j_summary.log
echo "== K=64 ADIL_GRAD_MAX_TP=48" | tee -a $OUT/j_summary.log
ADIL_GRAD_MAX_TP=48 python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig --iters 20 --K 64 2>&1 | grep -E "^auto|rror" | tee -a $OUT/j_summary.log
