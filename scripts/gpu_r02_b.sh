#!/bin/bash
# GPU pass (1 GPU): whole GPU test suite, kernel timings, bench.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TAG=${1:-b}
python -m pytest tests -m gpu -q --deselect "tests/test_imagenet_gpu.py::test_fooling_rate_within_half_a_point_of_the_reference[densenet121]" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_summary.log
tail -15 $OUT/${TAG}_pytest.log | tee -a $OUT/${TAG}_summary.log
for early in 1 0; do
  ADIL_GRAD_EARLY=$early python scripts/kernel_bench.py --impls auto --only grad,grad_dict_step,grad_dict_step_partials,grad_partials --iters 15 > $OUT/${TAG}_kb_grad_early$early.log 2>&1
  echo "== ADIL_GRAD_EARLY=$early" | tee -a $OUT/${TAG}_summary.log; grep -E "^auto|code_step|launches" $OUT/${TAG}_kb_grad_early$early.log | tee -a $OUT/${TAG}_summary.log
done
python scripts/kernel_bench.py --impls auto --only synth --iters 15 > $OUT/${TAG}_kb_synth.log 2>&1; grep -E "^auto" $OUT/${TAG}_kb_synth.log | tee -a $OUT/${TAG}_summary.log
python scripts/kernel_bench.py --impls auto --K 100 --iters 10 > $OUT/${TAG}_kb_k100.log 2>&1; echo "== K=100" | tee -a $OUT/${TAG}_summary.log; grep -E "^auto|code_step|launches" $OUT/${TAG}_kb_k100.log | tee -a $OUT/${TAG}_summary.log
python scripts/kernel_bench.py --impls auto,fma --K 200 --iters 5 --only synth > $OUT/${TAG}_kb_k200.log 2>&1; echo "== K=200 synth" | tee -a $OUT/${TAG}_summary.log; grep -E "^auto|^fma" $OUT/${TAG}_kb_k200.log | tee -a $OUT/${TAG}_summary.log
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?" | tee -a $OUT/${TAG}_summary.log
python - $TAG <<'PY' | tee -a gpurun_out/${TAG}_summary.log
import json, sys
try:
    d = json.loads(open("gpurun_out/%s_bench.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "variants", d["e2e_variants"])
    print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
    print("cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], "per_rank", d["per_rank"]["step_ms"])
except Exception as e:
    print("bench parse failed", e)
PY
