#!/bin/bash
# 8-GPU refresh after the kernel revision of the second session: distributed parity at 8 ranks + BASELINE configs[1] (cfg2)
set -u
mkdir -p gpurun_out
OUT=gpurun_out
N=${1:-8}
rm -f $OUT/h2_summary.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29631 scripts/dist_parity.py > $OUT/h2_dist_parity_w$N.log 2>&1; echo "dist_parity rc=$?" | tee -a $OUT/h2_summary.log
grep -E '"backend"|"pass"|m_rel|v_abs|D_abs|D_frac|replicas|peer_step|sharded_step|replicated|K=' $OUT/h2_dist_parity_w$N.log | tee -a $OUT/h2_summary.log
timeout 400 $TR --master-port 29642 bench.py --gpus $N --config 2 --steps 20 --warmup 5 > $OUT/h2_bench_cfg2_n$N.json 2> $OUT/h2_bench_cfg2_n$N.err; echo "bench cfg2 rc=$?" | tee -a $OUT/h2_summary.log
python - $N <<'PY' | tee -a gpurun_out/h2_summary.log
import json, sys
try:
    d = json.loads(open("gpurun_out/h2_bench_cfg2_n%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("cfg 2 N", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "product", round(d["e2e_variants"]["product_default_resident_cached_labels"], 1))
    print("   kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
    pr = d["per_rank"]
    print("   per-rank median step ms", [round(r["median"], 2) for r in pr["step_ms"]], "clk", pr["sm_mhz_median"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/h2_bench_cfg2_n%s.err" % sys.argv[1]).read()[-2500:])
PY
