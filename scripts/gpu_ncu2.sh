#!/bin/bash
set -u
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = "cur" ]; then L=$PWD/dl_attack_on_imagenet_b200/libadil_b200.so; else L=$PWD/scripts/libadil_b200_$v.so; fi
  export ADIL_B200_LIB=$L
  KB="python scripts/kernel_bench.py --impls auto --iters 3 --only grad_dict_step_contig"
  $KB > gpurun_out/prof_var_$v.plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:grad_kernel -s 3 -c 1 -f -o gpurun_out/prof_var_$v $KB > gpurun_out/prof_var_$v.ncu.log 2>&1
  echo "ncu $v rc=$?"
done
