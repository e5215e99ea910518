import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dl_attack_on_imagenet_b200 import ops
K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B, P, N = 100, 3 * 224 * 224, 128
gen = torch.Generator(device="cuda").manual_seed(K)
D2 = (-1 + 2 * torch.rand(P, K, device="cuda", generator=gen))
v = torch.rand(N, K, device="cuda", generator=gen) * (8 / 255 / K)
idx = torch.randperm(N, device="cuda", generator=gen)[:B]
ref = (v[idx].double() @ D2.double().t())
TP = int(sys.argv[2]) if len(sys.argv) > 2 else 48
for trial in range(4):
    _, d = ops.synth(D2, v, idx, delta_out=torch.empty(B, P, device="cuda"), want_out=False)
    err = (d.double() - ref).abs()
    bad = (err > 1e-7).nonzero()
    print("trial", trial, "max err", err.max().item(), "n bad", bad.shape[0])
    if bad.shape[0]:
        bs = bad[:, 0].unique(); ps = bad[:, 1].unique()
        tiles = (ps // TP).unique()
        print("  n bad images", bs.numel(), " bad pixels n", ps.numel(), " tiles", tiles[:16].tolist(), " its", (tiles // 148).unique().tolist(), " local px", (ps % TP).unique().tolist())
        p = ps[0].item(); b = bad[0, 0].item()
        print("  sample", b, p, d[b, p].item(), ref[b, p].item())
