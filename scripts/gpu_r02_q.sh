#!/bin/bash
# round 2, third session, 2 GPUs: the product distributed fit at K = 200 (column-window kernels on the multi-GPU path:
# plain contractions + fused peer dictionary step) against the single-GPU fused step, and the bench line of config 5
set -u
mkdir -p gpurun_out
OUT=gpurun_out
N=${1:-2}
rm -f $OUT/q_summary.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
ADIL_PARITY_K=200 timeout 400 $TR --master-port 29631 scripts/dist_parity.py > $OUT/q_dist_parity_k200_w$N.log 2>&1; echo "dist_parity K=200 rc=$?" | tee -a $OUT/q_summary.log
grep -E '"backend"|"pass"|m_rel|v_abs|D_abs|D_frac|replicas' $OUT/q_dist_parity_k200_w$N.log | head -20 | tee -a $OUT/q_summary.log
timeout 400 $TR --master-port 29642 bench.py --gpus $N --config 5 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/q_bench_cfg5_n$N.json 2> $OUT/q_bench_cfg5_n$N.err; echo "bench cfg5 rc=$?" | tee -a $OUT/q_summary.log
python - $N <<'PY' | tee -a gpurun_out/q_summary.log
import json, sys
try:
    d = json.loads(open("gpurun_out/q_bench_cfg5_n%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("cfg 5 N", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "roofline", d["roofline"])
    print("   kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/q_bench_cfg5_n%s.err" % sys.argv[1]).read()[-2500:])
PY
rm -rf gpurun_out/trained_dicts trained_dicts   # (the K = 200 dictionary files are 120 MB each: beyond what gpurun copies back)
ls -la gpurun_out | head -30
