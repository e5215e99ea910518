#!/bin/bash
# round 2, third session: one slab-reduction launch for both column windows
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_adil_gpu.py -m gpu -q > $OUT/t_pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/t_summary.log
tail -3 $OUT/t_pytest.log | tee -a $OUT/t_summary.log
for K in 200 136; do
  echo "== K=$K" | tee -a $OUT/t_summary.log
  python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig,grad_contig,grad --iters 20 --K $K 2>&1 | grep -E "^auto|rror" | tee -a $OUT/t_summary.log
done
python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/t_bench_cfg5.json 2> $OUT/t_bench_cfg5.err; echo "bench cfg5 rc=$?" | tee -a $OUT/t_summary.log
python - <<'PY' | tee -a gpurun_out/t_summary.log
import json
d = json.loads(open("gpurun_out/t_bench_cfg5.json").read().strip().splitlines()[-1])
print("cfg5 value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"], "roofline", round(d["roofline"]["frac"], 3),
      "kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
PY
