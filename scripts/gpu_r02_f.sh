#!/bin/bash
# multi-GPU pass: fused peer-memory dictionary step vs the NCCL chain (parity + timing), bench at N ranks
set -u
mkdir -p gpurun_out
OUT=gpurun_out
N=${1:-2}
for mode in auto nccl; do
  ADIL_DICT_STEP=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 scripts/dist_parity.py > $OUT/f_dist_parity_w${N}_$mode.log 2>&1; echo "dist_parity[$mode] rc=$?" | tee -a $OUT/f_summary.log
  grep -E "backend|multicast|multimem|\"pass\"|m_rel|v_abs|D_abs|D_frac|replicas|peer|sharded_step|replicated|K=" $OUT/f_dist_parity_w${N}_$mode.log | tee -a $OUT/f_summary.log
  tail -5 $OUT/f_dist_parity_w${N}_$mode.log | grep -v "^ \|^}" | tee -a $OUT/f_summary.log
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/f_bench_n$N.json 2> $OUT/f_bench_n$N.err; echo "bench rc=$?" | tee -a $OUT/f_summary.log
python - $N <<'PY' | tee -a gpurun_out/f_summary.log
import json, sys
try:
    d = json.loads(open("gpurun_out/f_bench_n%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "variants", d["e2e_variants"])
    print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "kernels", {k: (round(v["ms"] * 1e3, 1), round(v.get("frac", 0), 3)) for k, v in d["kernels"].items()})
    print("per_rank", d["per_rank"]["step_ms"], d["impl_notes"]["parallelism"])
except Exception as e:
    print("bench parse failed", e)
    print(open("gpurun_out/f_bench_n%s.err" % sys.argv[1]).read()[-3000:])
PY
