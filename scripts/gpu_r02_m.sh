#!/bin/bash
# round 2, third session: plain dD output of a column window as direct 128-bit stores of the epilogue warps
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/m_smoke.log 2>&1; echo "smoke rc=$?" | tee $OUT/m_summary.log; tail -1 $OUT/m_smoke.log | tee -a $OUT/m_summary.log
python -m pytest tests/test_kernels_gpu.py tests/test_adil_gpu.py -m gpu -q > $OUT/m_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/m_summary.log
tail -3 $OUT/m_pytest.log | tee -a $OUT/m_summary.log
for K in 200 256 136; do
  echo "== K=$K" | tee -a $OUT/m_summary.log
  python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig,grad_contig,grad --iters 20 --K $K 2>&1 | grep -E "^auto|rror" | tee -a $OUT/m_summary.log
done
