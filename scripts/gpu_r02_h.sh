#!/bin/bash
# 8-GPU pass: distributed parity at 8 ranks (fused peer kernel), BASELINE configs[1..4] at 8 GPUs, transfer sweep.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29631 scripts/dist_parity.py > $OUT/h_dist_parity_w$N.log 2>&1; echo "dist_parity rc=$?" | tee -a $OUT/h_summary.log
grep -E '"backend"|"pass"|m_rel|v_abs|D_abs|D_frac|replicas|peer|sharded_step|replicated|all_reduce_us|K=' $OUT/h_dist_parity_w$N.log | tee -a $OUT/h_summary.log
for cfg in 2 3 4 5; do
  steps=10; warm=3; if [ $cfg = 2 ]; then steps=20; warm=5; fi
  timeout 600 $TR --master-port 2964$cfg bench.py --gpus $N --config $cfg --steps $steps --warmup $warm > $OUT/h_bench_cfg${cfg}_n$N.json 2> $OUT/h_bench_cfg${cfg}_n$N.err; echo "bench cfg$cfg rc=$?" | tee -a $OUT/h_summary.log
  python - $cfg $N <<'PY' | tee -a gpurun_out/h_summary.log
import json, sys
try:
    d = json.loads(open("gpurun_out/h_bench_cfg%s_n%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
    print("cfg", sys.argv[1], d["config"]["model"], "N", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "product", round(d["e2e_variants"]["product_default_resident_cached_labels"], 1))
    print("   kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
    pr = d["per_rank"]
    print("   per-rank median step ms", [round(r["median"], 2) for r in pr["step_ms"]], "max", [round(r["max"], 2) for r in pr["step_ms"]], "clk", pr["sm_mhz_median"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/h_bench_cfg%s_n%s.err" % (sys.argv[1], sys.argv[2])).read()[-2500:])
PY
done
timeout 600 $TR --master-port 29651 scripts/transfer_sweep.py > $OUT/h_sweep_w$N.log 2>&1; echo "sweep rc=$?" | tee -a $OUT/h_summary.log
tail -2 $OUT/h_sweep_w$N.log | cut -c1-1800 | tee -a $OUT/h_summary.log
