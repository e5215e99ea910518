#!/bin/bash
# 2-GPU pass: multi-GPU parity of the product distributed fit + collective timings, bench at N=2 (both arms of the driver's launch).
set -u
mkdir -p gpurun_out
OUT=gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/dist_parity.py > $OUT/c_dist_parity_w$N.log 2>&1; echo "dist_parity rc=$?" | tee -a $OUT/c_summary.log
tail -40 $OUT/c_dist_parity_w$N.log | tee -a $OUT/c_summary.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/c_bench_n$N.json 2> $OUT/c_bench_n$N.err; echo "bench rc=$?" | tee -a $OUT/c_summary.log
python - $N <<'PY' | tee -a gpurun_out/c_summary.log
import json, sys
try:
    d = json.loads(open("gpurun_out/c_bench_n%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "variants", d["e2e_variants"])
    print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
    print("per_rank", d["per_rank"])
except Exception as e:
    print("bench parse failed", e)
    print(open("gpurun_out/c_bench_n%s.err" % sys.argv[1]).read()[-3000:])
PY
