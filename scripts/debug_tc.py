"""GPU debug helper for the tcgen05 grad kernel (not part of the test-suite)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dl_attack_on_imagenet_b200 import ops

torch.manual_seed(0)
def run(B, P, K, tag):
    g = torch.randn(B, P, device='cuda')
    D2 = torch.randn(P, K, device='cuda')
    v = torch.randn(B, K, device='cuda')
    ops.set_impl(ops.IMPL_FMA)
    dD_f, dv_f = ops.grad(g, D2, v, None, None)
    ops.set_impl(ops.IMPL_TC)
    dD = torch.full((P, K), 7.0, device='cuda'); dv = torch.full((B, K), 7.0, device='cuda')
    torch.cuda.synchronize(); t0 = time.time()
    ops.grad(g, D2, v, None, None, dD2=dD, dvb=dv)
    torch.cuda.synchronize(); t1 = time.time()
    ref_dD = (g.double().t() @ v.double()).float(); ref_dv = (g.double() @ D2.double()).float()
    print(f"[{tag}] B={B} P={P} K={K} time {t1-t0:.3f}s  dD err {(dD-ref_dD).abs().max():.3e} (ref max {ref_dD.abs().max():.3e}, fma err {(dD_f-ref_dD).abs().max():.3e})"
          f"  dv err {(dv-ref_dv).abs().max():.3e} (ref max {ref_dv.abs().max():.3e})")
    print("   dD: #sevens", int((dD == 7).sum()), "#zeros", int((dD == 0).sum()), "of", dD.numel(), " dv: #sevens", int((dv == 7).sum()), "#zeros", int((dv == 0).sum()))
    if (dD - ref_dD).abs().max() > 1e-3:
        print("   dD[0:4,0:6]\n", dD[:4, :6].cpu(), "\n   ref\n", ref_dD[:4, :6].cpu())
        # which rows / cols are right?
        ok = (dD - ref_dD).abs() < 1e-3
        print("   rows all-ok:", ok.all(1).nonzero().flatten()[:16].tolist(), " cols all-ok:", ok.all(0).nonzero().flatten().tolist())
    if (dv - ref_dv).abs().max() > 1e-2:
        print("   dv[0:4,0:6]\n", dv[:4, :6].cpu(), "\n   ref\n", ref_dv[:4, :6].cpu())
        ok = (dv - ref_dv).abs() < 1e-2
        print("   rows all-ok:", ok.all(1).nonzero().flatten().tolist(), " cols all-ok:", ok.all(0).nonzero().flatten().tolist())

for (B, P, K) in [(8, 128, 16), (8, 128, 8), (16, 256, 16), (4, 192, 6), (100, 1280, 50)]:
    try:
        run(B, P, K, "tc")
    except Exception as e:
        print("EXC", B, P, K, repr(e)[:300])
        break
