"""Cold (L2 flushed between launches) vs warm (back-to-back) timing of the two kernels at a small and the full size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dl_attack_on_imagenet_b200 import ops
B, K, N = 100, 50, 256
dev = torch.device("cuda")
flush = torch.empty(64 * 1024 * 1024, device=dev)
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
def timeit(fn, cold, iters=9):
    ts = []
    for i in range(iters + 3):
        if cold: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2] * 1e3
for hw in (3136, 50176):
    P = 3 * hw
    D2 = torch.rand(P, K, device=dev); m = torch.zeros_like(D2); s = torch.zeros_like(D2)
    v = torch.rand(N, K, device=dev) * 1e-3; x = torch.rand(B, P, device=dev); g = torch.randn(B, P, device=dev) * 1e-3
    idx = torch.randperm(N, device=dev)[:B]; out = torch.empty(B, P, device=dev); dvb = torch.empty(B, K, device=dev)
    fs = lambda: ops.synth(D2, v, idx, x=x, mean=MEAN, std=STD, flags=ops.SYNTH_NORMALIZE, out=out)
    fg = lambda: ops.grad_dict_step(D2, m, s, g, v, idx, ops.adamw_params(3, 0.01), STD, dvb=dvb)
    fn = lambda: ops.project_rows(v, ops.ROWS_NONE, 0.0)
    print(f"P={P}: synth cold {timeit(fs, True):6.1f} warm {timeit(fs, False):6.1f} | grad_dict_step cold {timeit(fg, True):6.1f} warm {timeit(fg, False):6.1f} | tiny kernel cold {timeit(fn, True):5.1f} warm {timeit(fn, False):5.1f}")
