#!/bin/bash
set -u
mkdir -p gpurun_out
KB="python scripts/kernel_bench.py --impls auto --only ${ONLY:-grad_dict_step_partials} --iters 6 ${KBARGS:-}"
ADIL_B200_LIB=$PWD/scripts/libadil_b200_timing.so $KB > gpurun_out/exp_tim.log 2>&1
grep "grad CTA0" gpurun_out/exp_tim.log | tail -4
grep "synth CTA0" gpurun_out/exp_tim.log | tail -3
grep "chain" gpurun_out/exp_tim.log | tail -2
