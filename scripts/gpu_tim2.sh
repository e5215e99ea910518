set -u
for K in 64 100; do echo "== K=$K"; ONLY=grad_dict_step_contig KBARGS="--K $K" bash scripts/gpu_tim.sh 2>&1 | grep "grad CTA" | tail -3 | cut -c1-600; done
