#!/bin/bash
set -u
KB="python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig,grad_contig,synth --iters 20"
for v in "$@"; do
  if [ "$v" = "cur" ]; then L=$PWD/dl_attack_on_imagenet_b200/libadil_b200.so; else L=$PWD/scripts/libadil_b200_$v.so; fi
  echo "== $v"; ADIL_B200_LIB=$L $KB 2>&1 | grep -E "^auto|rror"
done
