"""PyTorch-eager cost of the ADiL side of one minibatch step on the same GPU -- the op sequence the reference runs
(adil.py:24-35,154,186-188; demo_dL_attack.py:22-25; utils.py:21-41) written with plain torch ops -- next to the fused
kernels of this repo, both timed with CUDA events and a flushed L2.  The classifier is replaced by a given gradient
`g` (it is identical on both sides).  This is the honest "reference on a B200" bar for the ADiL part of the step
(SURVEY section 8d); it is a measurement script, not a product path."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dl_attack_on_imagenet_b200 import ops

B, K, N, P = 100, 50, 1024, 3 * 224 * 224
EPS = 8 / 255
dev = torch.device("cuda")
torch.manual_seed(0)
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
mean_t = torch.tensor(MEAN, device=dev).view(1, 3, 1, 1)
std_t = torch.tensor(STD, device=dev).view(1, 3, 1, 1)
x_all = torch.rand(N, 3, 224, 224, device=dev)
g = torch.randn(B, 3, 224, 224, device=dev) * 1e-3
idx_cpu = torch.randperm(N)[:B]            # the reference indexes v with a CPU LongTensor (adil.py:25,168)
flush = torch.empty(64 * 1024 * 1024, device=dev)


def project_onto_l1_ball(x, eps):          # utils.py:21-41, verbatim algorithm (sort, cumsum, rho.cpu() sync)
    original_shape = x.shape
    x = x.view(x.shape[0], -1)
    mask = (torch.norm(x, p=1, dim=1) < eps).float().unsqueeze(1)
    mu, _ = torch.sort(torch.abs(x), dim=1, descending=True)
    cumsum = torch.cumsum(mu, dim=1)
    arange = torch.arange(1, x.shape[1] + 1, device=x.device)
    rho, _ = torch.max((mu * arange > (cumsum - eps)) * arange, dim=1)
    theta = (cumsum[torch.arange(x.shape[0]), rho.cpu() - 1] - eps) / rho
    proj = (torch.abs(x) - theta.unsqueeze(1)).clamp(min=0)
    x = mask * x + (1 - mask) * proj * torch.sign(x)
    return x.view(original_shape)


def timeit(fn, iters=10):
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


# ---- eager (reference op sequence) -------------------------------------------------------------------------
d = torch.nn.Parameter(-1 + 2 * torch.rand(3, 224, 224, K, device=dev))
v = torch.nn.Parameter(project_onto_l1_ball(torch.rand(N, K, device=dev), EPS))
opt = torch.optim.AdamW([d, v], lr=0.01)
xb = x_all[idx_cpu.to(dev)]
state = {}

def eager_forward():
    dv = torch.tensordot(v[idx_cpu, :], d, dims=([1], [3]))           # adil.py:25
    state["xin"] = ((xb + dv) - mean_t) / std_t                      # adil.py:26 + Normalize (demo:22-25)

def eager_backward():
    opt.zero_grad(set_to_none=True)
    state["xin"].backward(g)                                         # autograd of the above (adil.py:185)

def eager_update():
    opt.step()                                                       # adil.py:186
    with torch.no_grad():
        v.copy_(project_onto_l1_ball(v, EPS))                        # adil.py:29-31,187
        d.copy_(torch.clamp(d, min=-1, max=1))                       # adil.py:33-35,188

def eager_step():
    eager_forward(); eager_backward(); eager_update()

eager_forward()
res = {"eager_forward_us": timeit(eager_forward)}
def fb():
    eager_forward(); eager_backward()
res["eager_forward_backward_us"] = timeit(fb)
res["eager_step_us"] = timeit(eager_step)

# ---- fused kernels of this repo ------------------------------------------------------------------------------
D2 = (-1 + 2 * torch.rand(P, K, device=dev)); mD = torch.zeros_like(D2); sD = torch.zeros_like(D2)
V = torch.rand(N, K, device=dev) * 1e-3; mV = torch.zeros_like(V); sV = torch.zeros_like(V)
x2 = x_all.view(N, P); g2 = g.view(B, P).contiguous(); idx = idx_cpu.to(dev)
out = torch.empty(B, P, device=dev); dvb = torch.empty(B, K, device=dev)
f_synth = lambda: ops.synth(D2, V, idx, x=x2, x_index=idx, mean=MEAN, std=STD, flags=ops.SYNTH_NORMALIZE, out=out)
f_grad = lambda: ops.grad_dict_step(D2, mD, sD, g2, V, idx, ops.adamw_params(3, 0.01), STD, dvb=dvb)
f_code = lambda: ops.code_step(V, mV, sV, dvb, idx, ops.adamw_params(3, 0.01), ops.ROWS_L1BALL, EPS)
def fused_step():
    f_synth(); f_grad(); f_code()
res["fused_synth_us"] = timeit(f_synth)
res["fused_grad_dict_step_us"] = timeit(f_grad)
res["fused_code_step_us"] = timeit(f_code)
res["fused_step_us"] = timeit(fused_step)
res["speedup_step"] = res["eager_step_us"] / res["fused_step_us"]
print(json.dumps(res, indent=1))
