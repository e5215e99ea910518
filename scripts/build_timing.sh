#!/bin/bash
# Debug build (never the product library): ADIL_DEBUG_DEFS=-DADIL_CHAIN (default: global-timer stamps of the hand-off
# chain and of the prologue / tail) or -DADIL_TIMING (per-phase clock64 counters, synthesis kernel).
set -e
cd "$(dirname "$0")/.."
S=dl_attack_on_imagenet_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared ${ADIL_DEBUG_DEFS:--DADIL_CHAIN} \
  -Iinclude -I$S -o scripts/libadil_b200_timing.so $S/adil_api.cu $S/adil_fma.cu $S/adil_steps.cu $S/adil_tc.cu
echo scripts/libadil_b200_timing.so
