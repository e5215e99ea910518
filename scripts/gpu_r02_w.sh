#!/bin/bash
# round 2, third session: synthesis kernel with the code-row path as a template parameter -- kernel tests + config-2 bench line
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > $OUT/x_pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/x_summary.log
tail -2 $OUT/x_pytest.log | tee -a $OUT/x_summary.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/x_bench_cfg2.json 2> $OUT/x_bench_cfg2.err; echo "bench rc=$?" | tee -a $OUT/x_summary.log
python - <<'PY' | tee -a gpurun_out/x_summary.log
import json
d = json.loads(open("gpurun_out/x_bench_cfg2.json").read().strip().splitlines()[-1])
print("cfg2 value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"], "roofline", round(d["roofline"]["frac"], 3),
      "kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
PY
