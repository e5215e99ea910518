#!/bin/bash
# round 2, third session: synthesis kernel exports the code rows from its register path (no gather launch beyond ~100 atoms)
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_adil_gpu.py -m gpu -q > $OUT/s_pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/s_summary.log
tail -3 $OUT/s_pytest.log | tee -a $OUT/s_summary.log
for K in 50 128 200; do
  echo "== synth K=$K" | tee -a $OUT/s_summary.log
  python scripts/kernel_bench.py --impls auto --only synth,synth_contig --iters 20 --K $K 2>&1 | grep -E "^auto|rror" | tee -a $OUT/s_summary.log
done
python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/s_bench_cfg5.json 2> $OUT/s_bench_cfg5.err; echo "bench cfg5 rc=$?" | tee -a $OUT/s_summary.log
python - <<'PY' | tee -a gpurun_out/s_summary.log
import json
d = json.loads(open("gpurun_out/s_bench_cfg5.json").read().strip().splitlines()[-1])
print("cfg5 value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"], "roofline", round(d["roofline"]["frac"], 3),
      "kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
PY
