#!/bin/bash
set -u
KB="python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig,grad_contig --iters 20"
for v in cur v1 v2; do
  if [ "$v" = "cur" ]; then L=$PWD/dl_attack_on_imagenet_b200/libadil_b200.so; else L=$PWD/scripts/libadil_b200_$v.so; fi
  for K in 50 64 100; do echo "== $v K=$K"; ADIL_B200_LIB=$L $KB --K $K 2>&1 | grep -E "^auto|rror"; done
done
