#!/bin/bash
# First GPU pass of round 2 (1 GPU): smoke, the whole GPU test suite, kernel A/B runs of the new prologue knobs, bench.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/a_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/a_summary.log
python -m pytest tests -m gpu -q -x --deselect "tests/test_imagenet_gpu.py::test_fooling_rate_within_half_a_point_of_the_reference[densenet121]" > $OUT/a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/a_summary.log
tail -5 $OUT/a_pytest.log | tee -a $OUT/a_summary.log
for early in 1 0; do
  ADIL_GRAD_EARLY=$early python scripts/kernel_bench.py --impls auto --only grad,grad_dict_step,grad_dict_step_partials,grad_partials --iters 15 > $OUT/a_kb_grad_early$early.log 2>&1
  echo "== ADIL_GRAD_EARLY=$early" | tee -a $OUT/a_summary.log; grep -E "^auto|code_step|launches" $OUT/a_kb_grad_early$early.log | tee -a $OUT/a_summary.log
done
for ex in 0 1 2 3; do
  ADIL_SYNTH_EARLY_X=$ex python scripts/kernel_bench.py --impls auto --only synth --iters 15 > $OUT/a_kb_synth_x$ex.log 2>&1
  echo "== ADIL_SYNTH_EARLY_X=$ex" | tee -a $OUT/a_summary.log; grep -E "^auto" $OUT/a_kb_synth_x$ex.log | tee -a $OUT/a_summary.log
done
python scripts/kernel_bench.py --impls auto --K 100 --iters 10 > $OUT/a_kb_k100.log 2>&1; echo "== K=100" | tee -a $OUT/a_summary.log; grep -E "^auto|code_step|launches" $OUT/a_kb_k100.log | tee -a $OUT/a_summary.log
python bench.py --steps 10 --warmup 3 > $OUT/a_bench.json 2> $OUT/a_bench.err; echo "bench rc=$?" | tee -a $OUT/a_summary.log
python - <<'PY' | tee -a gpurun_out/a_summary.log
import json
try:
    d = json.loads(open("gpurun_out/a_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "variants", d["e2e_variants"])
    print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
    print("cpu", d["cpu_baseline"], "launches", d["gpu_launches"])
except Exception as e:
    print("bench parse failed", e)
PY
