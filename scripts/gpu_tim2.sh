set -u
for K in 50 52 51; do echo "== K=$K"; ONLY=grad_dict_step_partials KBARGS="--K $K" bash scripts/gpu_tim.sh 2>&1 | grep "grad CTA0" | tail -2 | sed 's/.*| E: /E: /'; done
