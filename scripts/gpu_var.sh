#!/bin/bash
set -u
KB="python scripts/kernel_bench.py --impls auto --only synth,synth_contig --iters 30"
for v in cur v1 cur v1; do
  if [ "$v" = "cur" ]; then L=$PWD/dl_attack_on_imagenet_b200/libadil_b200.so; else L=$PWD/scripts/libadil_b200_$v.so; fi
  for K in 50 100; do echo "== $v K=$K"; ADIL_B200_LIB=$L $KB --K $K 2>&1 | grep -E "^auto|rror"; done
done
