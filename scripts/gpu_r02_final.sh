#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/parity_r02.json
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/z_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/z_summary.log; tail -1 $OUT/z_smoke.log | tee -a $OUT/z_summary.log
python -m pytest tests -m gpu -q > $OUT/z_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/z_summary.log
tail -6 $OUT/z_pytest.log | tee -a $OUT/z_summary.log
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/z_bench_ref.json 2> $OUT/z_bench_ref.err; echo "ref rc=$?" | tee -a $OUT/z_summary.log
python bench.py --steps 20 --warmup 5 > $OUT/z_bench.json 2> $OUT/z_bench.err; echo "bench rc=$?" | tee -a $OUT/z_summary.log
python - <<'PY' | tee -a gpurun_out/z_summary.log
import json
r = json.loads(open("gpurun_out/z_bench_ref.json").read().strip().splitlines()[-1])
d = json.loads(open("gpurun_out/z_bench.json").read().strip().splitlines()[-1])
print("reference arm:", r["value"], r["cpu_baseline"]["cores"], "same config:", r["config"] == d["config"])
print("value", d["value"], "e2e", d["e2e"]["value"], "variants", d["e2e_variants"])
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["traffic"], "kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
PY
