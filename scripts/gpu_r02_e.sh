#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests -m gpu -q > $OUT/e_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/e_summary.log
tail -25 $OUT/e_pytest.log | tee -a $OUT/e_summary.log
python scripts/kernel_bench.py --impls auto --K 200 --iters 8 --only synth,grad,grad_dict_step,dict_step > $OUT/e_kb_k200.log 2>&1; echo "== K=200" | tee -a $OUT/e_summary.log; grep -E "^auto|code_step|launches" $OUT/e_kb_k200.log | tee -a $OUT/e_summary.log
python scripts/kernel_bench.py --impls auto --K 64 --iters 8 --only synth,grad,grad_dict_step_partials > $OUT/e_kb_k64.log 2>&1; echo "== K=64" | tee -a $OUT/e_summary.log; grep -E "^auto" $OUT/e_kb_k64.log | tee -a $OUT/e_summary.log
python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/e_bench_cfg5.json 2> $OUT/e_bench_cfg5.err; echo "bench cfg5 rc=$?" | tee -a $OUT/e_summary.log
python - <<'PY' | tee -a gpurun_out/e_summary.log
import json
try:
    d = json.loads(open("gpurun_out/e_bench_cfg5.json").read().strip().splitlines()[-1])
    print("cfg5 slice: value", d["value"], "e2e", d["e2e"]["value"], "variants", d["e2e_variants"])
    print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/e_bench_cfg5.err").read()[-2000:])
PY
