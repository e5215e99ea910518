#!/bin/bash
# round 2, third session: synthesis kernel, register path of the code rows (K > ~100) with 128-bit loads
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -q > $OUT/r_pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/r_summary.log
tail -3 $OUT/r_pytest.log | tee -a $OUT/r_summary.log
for K in 50 64 100 128 200 224; do
  echo "== synth K=$K" | tee -a $OUT/r_summary.log
  python scripts/kernel_bench.py --impls auto --only synth,synth_contig --iters 20 --K $K 2>&1 | grep -E "^auto|rror" | tee -a $OUT/r_summary.log
done
