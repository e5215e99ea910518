#!/bin/bash
# final-state validation of round 2 (third session): smoke, every GPU test, both bench arms at cfg2, bench lines of the
# other BASELINE configs on one GPU, kernel table at every K
set -u
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/paritw_r02.json $OUT/w_summary.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/w_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/w_summary.log; tail -1 $OUT/w_smoke.log | tee -a $OUT/w_summary.log
python -m pytest tests -m gpu -q > $OUT/w_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/w_summary.log
tail -6 $OUT/w_pytest.log | tee -a $OUT/w_summary.log
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/w_bench_ref.json 2> $OUT/w_bench_ref.err; echo "ref rc=$?" | tee -a $OUT/w_summary.log
python bench.py --steps 20 --warmup 5 > $OUT/w_bench_cfg2.json 2> $OUT/w_bench_cfg2.err; echo "bench rc=$?" | tee -a $OUT/w_summary.log
for cfg in 1 3 4 5; do
  python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > $OUT/w_bench_cfg$cfg.json 2> $OUT/w_bench_cfg$cfg.err; echo "bench cfg$cfg rc=$?" | tee -a $OUT/w_summary.log
done
python - <<'PY' | tee -a gpurun_out/w_summary.log
import json
r = json.loads(open("gpurun_out/w_bench_ref.json").read().strip().splitlines()[-1])
for cfg in (2, 1, 3, 4, 5):
    try:
        d = json.loads(open("gpurun_out/w_bench_cfg%d.json" % cfg).read().strip().splitlines()[-1])
        print("cfg", cfg, d["config"].get("workload", "")[:60], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
              "product", round(d["e2e_variants"]["product_default_resident_cached_labels"], 1), "roofline", round(d["roofline"]["frac"], 3),
              "kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
        if cfg == 2: print("reference arm:", r["value"], r["cpu_baseline"]["cores"], "same config:", r["config"] == d["config"], "clocks", d.get("clocks"))
    except Exception as e:
        print("cfg", cfg, "parse failed", e)
PY
for K in 50 64 100 128 200; do echo "== K=$K" | tee -a $OUT/w_summary.log; python scripts/kernel_bench.py --impls auto --only synth,grad_dict_step_contig,grad_contig --iters 20 --K $K 2>&1 | grep -E "^auto|rror" | tee -a $OUT/w_summary.log; done
echo "== K=200, one dv accumulator (ADIL_GRAD_DV2=0)" | tee -a $OUT/w_summary.log
ADIL_GRAD_DV2=0 python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig,grad_contig --iters 20 --K 200 2>&1 | grep -E "^auto|rror" | tee -a $OUT/w_summary.log
echo "== K=200, two serial launches (ADIL_GRAD_WINDOWS=serial)" | tee -a $OUT/w_summary.log
ADIL_GRAD_WINDOWS=serial python scripts/kernel_bench.py --impls auto --only grad_dict_step_contig,grad_contig --iters 20 --K 200 2>&1 | grep -E "^auto|rror" | tee -a $OUT/w_summary.log
