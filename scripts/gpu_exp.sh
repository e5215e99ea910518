#!/bin/bash
# A/B: baseline .so vs new build, interleaved (box-to-box and run-to-run noise is ~1 us), + stamps + parity tests
set -u
mkdir -p gpurun_out
KBN="python scripts/kernel_bench.py --impls auto --only grad_dict_step_partials,grad_partials,synth,grad_dict_step_contig,grad_contig,synth_contig --iters 20"
for rep in 1 2; do
echo "== new";  $KBN 2>&1 | grep -E "^auto|rror"
done
if [ "${FULL:-0}" = "1" ]; then for K in 64 100 200; do echo "== new K=$K"; $KBN --K $K 2>&1 | grep -E "^auto|rror"; done; fi
ONLY=grad_dict_step_contig,synth bash scripts/gpu_tim.sh
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/exp_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/exp_pytest.log
