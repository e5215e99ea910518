"""Kernel time vs problem size (fixed B, K): separates fixed per-launch cost from per-tile cost."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dl_attack_on_imagenet_b200 import ops
B, K, N = 100, 50, 256
dev = torch.device("cuda")
flush = torch.empty(64 * 1024 * 1024, device=dev)
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
def timeit(fn, iters=7):
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2] * 1e3
for hw in (64 * 148 // 3 * 3 // 3, 12544, 25088, 50176, 100352):
    hw = hw // 64 * 64
    P = 3 * hw
    D2 = torch.rand(P, K, device=dev); m = torch.zeros_like(D2); s = torch.zeros_like(D2)
    v = torch.rand(N, K, device=dev) * 1e-3; x = torch.rand(B, P, device=dev); g = torch.randn(B, P, device=dev) * 1e-3
    idx = torch.randperm(N, device=dev)[:B]; out = torch.empty(B, P, device=dev); dvb = torch.empty(B, K, device=dev)
    ts = timeit(lambda: ops.synth(D2, v, idx, x=x, mean=MEAN, std=STD, flags=ops.SYNTH_NORMALIZE, out=out))
    tg = timeit(lambda: ops.grad_dict_step(D2, m, s, g, v, idx, ops.adamw_params(3, 0.01), STD, dvb=dvb))
    tg2 = timeit(lambda: ops.grad_dict_step(D2, m, s, g, v, idx, ops.adamw_params(3, 0.01), STD, want_dv=False))
    dD = torch.empty(P, K, device=dev)
    tg3 = timeit(lambda: ops.grad(g, D2, v, idx, STD, dD2=dD, want_dv=False))
    tg4 = timeit(lambda: ops.grad(g, D2, v, idx, STD, dvb=dvb, want_dD=False))
    print(f"      fused no-dv {tg2:7.1f}   unfused dD only {tg3:7.1f}   dv only {tg4:7.1f}")
    print(f"P={P:7d} tiles64={P//64:5d} ({P/64/148:5.2f}/SM)  synth {ts:7.1f} us   grad_dict_step {tg:7.1f} us", flush=True)
