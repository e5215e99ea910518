#!/bin/bash
# round 2, third session: A/B of the synthesis kernel in-step at config 2 -- current library vs the build before the
# 128-bit register path (scripts/libadil_b200_presynth.so, built from e45a179's adil_tc.cu) -- then the whole GPU suite
set -u
mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/v_summary.log
for v in cur pre cur pre; do
  if [ $v = pre ]; then export ADIL_B200_LIB=$PWD/scripts/libadil_b200_presynth.so; else unset ADIL_B200_LIB; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/v_bench_$v.json 2> $OUT/v_bench_$v.err; echo "bench lib=$v rc=$?" | tee -a $OUT/v_summary.log
  python - $v <<'PY' | tee -a gpurun_out/v_summary.log
import json, sys
d = json.loads(open("gpurun_out/v_bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("   value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
PY
done
unset ADIL_B200_LIB
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/v_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/v_summary.log; tail -1 $OUT/v_smoke.log | tee -a $OUT/v_summary.log
python -m pytest tests -m gpu -q > $OUT/v_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/v_summary.log
tail -3 $OUT/v_pytest.log | tee -a $OUT/v_summary.log
