#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python scripts/transfer_sweep.py --images-per-gpu 512 --eval-images-per-gpu 200 > $OUT/g_sweep_w1.log 2>&1; echo "sweep rc=$?" | tee -a $OUT/g_summary.log
tail -3 $OUT/g_sweep_w1.log | cut -c1-1500 | tee -a $OUT/g_summary.log
for cfg in 3 4 1; do
  python bench.py --config $cfg --steps 10 --warmup 3 > $OUT/g_bench_cfg$cfg.json 2> $OUT/g_bench_cfg$cfg.err; echo "bench cfg$cfg rc=$?" | tee -a $OUT/g_summary.log
  python - $cfg <<'PY' | tee -a gpurun_out/g_summary.log
import json, sys
try:
    d = json.loads(open("gpurun_out/g_bench_cfg%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("cfg", sys.argv[1], d["config"]["model"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "product", round(d["e2e_variants"]["product_default_resident_cached_labels"], 1), "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
    print("   kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/g_bench_cfg%s.err" % sys.argv[1]).read()[-2000:])
PY
done
