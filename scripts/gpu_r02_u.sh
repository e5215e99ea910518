#!/bin/bash
# round 2, third session: opt-in experiment -- the dictionary in the persisting L2 set-aside (ADIL_L2_PERSIST_D=1)
set -u
mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/u_summary.log
for v in 0 1 0 1; do
  ADIL_L2_PERSIST_D=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/u_bench_$v.json 2> $OUT/u_bench_$v.err; echo "bench persist=$v rc=$?" | tee -a $OUT/u_summary.log
  python - $v <<'PY' | tee -a gpurun_out/u_summary.log
import json, sys
d = json.loads(open("gpurun_out/u_bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("   value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "product", round(d["e2e_variants"]["product_default_resident_cached_labels"], 1),
      "kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
PY
done
ADIL_L2_PERSIST_D=1 python -m pytest tests/test_adil_gpu.py -m gpu -q -k "fit_gd or fit_alter or cuda_graph" > $OUT/u_pytest.log 2>&1; echo "pytest persist=1 rc=$?" | tee -a $OUT/u_summary.log
tail -2 $OUT/u_pytest.log | tee -a $OUT/u_summary.log
