"""Isolated timing of the ADiL kernels at BASELINE shapes (CUDA events, L2 flushed between launches)."""
import sys, os, json, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dl_attack_on_imagenet_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=100)
ap.add_argument("--K", type=int, default=50)
ap.add_argument("--N", type=int, default=1024)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--impls", default="fma,auto")
ap.add_argument("--only", default="")
ap.add_argument("--no-flush", action="store_true", help="back-to-back launches (L2 keeps data and code)")
ap.add_argument("--device-index", action="store_true", help="pass the batch indices as a device array (default: CPU tensor -> kernel parameters)")
args = ap.parse_args()
P = 3 * 224 * 224
B, K, N = args.B, args.K, args.N
dev = torch.device("cuda")
torch.manual_seed(0)
D2 = (-1 + 2 * torch.rand(P, K, device=dev))
m = torch.zeros_like(D2); s = torch.zeros_like(D2)
v = torch.rand(N, K, device=dev) * 1e-3
x = torch.rand(B, P, device=dev)
g = torch.randn(B, P, device=dev) * 1e-3
idx = torch.randperm(N, device=dev)[:B]
if not args.device_index:
    idx = idx.cpu()
out = torch.empty(B, P, device=dev)
dD = torch.empty(P, K, device=dev)
dvb = torch.empty(B, K, device=dev)
vb = v[idx.to(dev)].contiguous()   # the block adil_synth leaves behind (codes_out)
vb_out = torch.empty(B, K, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)  # 256 MB
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
PEAK = 6554.2

def timeit(fn, iters):
    ts = []
    for i in range(iters + 3):
        if not args.no_flush: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    ts.sort()
    timeit.mean = sum(ts[1:-1]) / max(1, len(ts) - 2)   # (CUDA events tick in 1.024 us steps: the trimmed mean resolves finer)
    return ts[len(ts) // 2], ts[0]

res = {}
for impl in args.impls.split(","):
    ops.set_impl({"fma": ops.IMPL_FMA, "auto": ops.IMPL_AUTO, "tc": ops.IMPL_TC}[impl])
    cases = {
        "synth": (lambda: ops.synth(D2, v, idx, x=x, mean=MEAN, std=STD, flags=ops.SYNTH_NORMALIZE, out=out, codes_out=vb_out), 4.0 * P * (2 * B + K) + 4.0 * B * K),
        "synth_contig": (lambda: ops.synth(D2, vb, None, x=x, mean=MEAN, std=STD, flags=ops.SYNTH_NORMALIZE, out=out), 4.0 * P * (2 * B + K) + 4.0 * B * K),
        "grad_dict_step_contig": (lambda: ops.grad_dict_step(D2, m, s, g, vb, None, ops.adamw_params(3, 0.01), STD, keep_partials=True), 4.0 * P * (B + 6 * K) + 8.0 * B * K),
        "grad_contig": (lambda: ops.grad(g, D2, vb, None, STD, dD2=dD, keep_partials=True), 4.0 * P * (B + 2 * K) + 8.0 * B * K),
        "grad": (lambda: ops.grad(g, D2, v, idx, STD, dD2=dD, dvb=dvb), 4.0 * P * (B + 2 * K) + 8.0 * B * K),
        "grad_dict_step": (lambda: ops.grad_dict_step(D2, m, s, g, v, idx, ops.adamw_params(3, 0.01), STD, dvb=dvb), 4.0 * P * (B + 6 * K) + 8.0 * B * K),
        "grad_dict_step_partials": (lambda: ops.grad_dict_step(D2, m, s, g, v, idx, ops.adamw_params(3, 0.01), STD, keep_partials=True), 4.0 * P * (B + 6 * K) + 8.0 * B * K),
        "grad_partials": (lambda: ops.grad(g, D2, v, idx, STD, dD2=dD, keep_partials=True), 4.0 * P * (B + 2 * K) + 8.0 * B * K),
        "dict_step": (lambda: ops.dict_step(D2, m, s, dD, ops.adamw_params(3, 0.01)), 28.0 * P * K),
    }
    for name, (fn, nbytes) in cases.items():
        if args.only and name not in args.only.split(","): continue
        med, best = timeit(fn, args.iters)
        res[f"{impl}.{name}"] = {"ms_median": med, "ms_best": best, "GBps": nbytes / med / 1e6, "frac": nbytes / med / 1e6 / PEAK}
        print(f"{impl:5s} {name:15s} median {med*1e3:8.1f} us  mean {timeit.mean*1e3:8.2f} us  best {best*1e3:8.1f} us  {nbytes/med/1e6:7.0f} GB/s  {100*nbytes/med/1e6/PEAK:5.1f}% of measured HBM peak", flush=True)
ops.set_impl(ops.IMPL_AUTO)
vv = torch.rand(N, K, device=dev) * 1e-2; mv = torch.zeros_like(vv); sv = torch.zeros_like(vv)
idx_d = idx.to(dev)
med, best = timeit(lambda: ops.code_step(vv, mv, sv, dvb, idx_d, ops.adamw_params(3, 0.01), ops.ROWS_L1BALL, 8 / 255), args.iters)
print(f"code_step N={N} K={K}: median {med*1e3:.1f} us best {best*1e3:.1f} us")
res["code_step"] = {"ms_median": med, "ms_best": best}
if K > 128:
    print(json.dumps(res)); sys.exit(0)
part = ops.grad_dict_step(D2, m, s, g, v, idx, ops.adamw_params(3, 0.01), STD, keep_partials=True)
med, best = timeit(lambda: ops.code_step(vv, mv, sv, part, idx_d, ops.adamw_params(3, 0.01), ops.ROWS_L1BALL, 8 / 255), args.iters)
print(f"code_step (reduces {part.nslabs} partial slabs itself) N={N} K={K}: median {med*1e3:.1f} us best {best*1e3:.1f} us")
res["code_step_partials"] = {"ms_median": med, "ms_best": best}
def pair():
    p_ = ops.grad_dict_step(D2, m, s, g, v, idx, ops.adamw_params(3, 0.01), STD, keep_partials=True)
    ops.code_step(vv, mv, sv, p_, idx_d, ops.adamw_params(3, 0.01), ops.ROWS_L1BALL, 8 / 255)
med, best = timeit(pair, args.iters)
print(f"grad_dict_step + code_step (2 launches): median {med*1e3:.1f} us best {best*1e3:.1f} us")
res["grad_dict_step+code_step"] = {"ms_median": med, "ms_best": best}
def pair_old():
    ops.grad_dict_step(D2, m, s, g, v, idx, ops.adamw_params(3, 0.01), STD, dvb=dvb)
    ops.code_step(vv, mv, sv, dvb, idx_d, ops.adamw_params(3, 0.01), ops.ROWS_L1BALL, 8 / 255)
med, best = timeit(pair_old, args.iters)
print(f"grad_dict_step + reduce + code_step (3 launches): median {med*1e3:.1f} us best {best*1e3:.1f} us")
res["grad_dict_step+reduce+code_step"] = {"ms_median": med, "ms_best": best}
print(json.dumps(res))
