#!/bin/bash
# round 2, third session: column windows of K > 128 -- paired launch with two dv accumulators, full dictionary rows in the
# raw stages (A/B: ADIL_GRAD_FULLROW=0, ADIL_GRAD_WINDOWS=serial); ncu capture of the K = 200 fused step
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -q > $OUT/k_pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/k_summary.log
tail -3 $OUT/k_pytest.log | tee -a $OUT/k_summary.log
ADIL_GRAD_FULLROW=0 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "200 or 256 or 136 or 176" > $OUT/k_pytest_rows.log 2>&1; echo "pytest (per-row copies) rc=$?" | tee -a $OUT/k_summary.log
tail -1 $OUT/k_pytest_rows.log | tee -a $OUT/k_summary.log
for K in 200 256 136; do
  for fr in 0 1; do
    echo "== K=$K paired fullrow=$fr" | tee -a $OUT/k_summary.log
    ADIL_GRAD_FULLROW=$fr python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig,grad_contig --iters 20 --K $K 2>&1 | grep -E "^auto|rror" | tee -a $OUT/k_summary.log
  done
done
KB="python scripts/kernel_bench.py --impls auto --iters 3 --K 200"
ncu --set full --clock-control none --import-source on -k regex:grad_kernel -s 3 -c 1 -f -o $OUT/prof_r02c_grad_k200 $KB --only grad_dict_step_contig > $OUT/prof_r02c_grad_k200.ncu.log 2>&1
echo "ncu rc=$?" | tee -a $OUT/k_summary.log
