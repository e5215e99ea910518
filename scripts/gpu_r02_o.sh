#!/bin/bash
# round 2, third session: tile cap of the fused column windows (TP = 32 / 16); tile-size probes of the other kernels
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_adil_gpu.py -m gpu -q > $OUT/o_pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/o_summary.log
tail -3 $OUT/o_pytest.log | tee -a $OUT/o_summary.log
for K in 136 176 200 256; do
  echo "== K=$K (default plan)" | tee -a $OUT/o_summary.log
  python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig,grad_contig --iters 20 --K $K 2>&1 | grep -E "^auto|rror" | tee -a $OUT/o_summary.log
done
for cfg in "200 32" "200 48" "136 32" "256 32" "256 16"; do
  set -- $cfg
  echo "== plain K=$1 ADIL_GRAD_MAX_TP=$2" | tee -a $OUT/o_summary.log
  ADIL_GRAD_MAX_TP=$2 python scripts/kernel_bench.py --impls auto --only grad_contig --iters 20 --K $1 2>&1 | grep -E "^auto|rror" | tee -a $OUT/o_summary.log
done
for cfg in "64 32" "50 32" "100 16" "128 16"; do
  set -- $cfg
  echo "== single window K=$1 ADIL_GRAD_MAX_TP=$2" | tee -a $OUT/o_summary.log
  ADIL_GRAD_MAX_TP=$2 python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig,grad_contig --iters 20 --K $1 2>&1 | grep -E "^auto|rror" | tee -a $OUT/o_summary.log
done
