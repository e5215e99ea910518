"""TEST INFRASTRUCTURE -- CPU oracle for the ADiL attack-learning hot path.

This file is a *restatement* (flattened, explicit formulas, torch-CPU fp32 / optional fp64) of the
algorithm in the reference `attacks/attacks_classes/adil.py` + `attacks/utils.py`, and of the four
function-level drivers of `attacks/attacks_classes/adil_regularized.py` (`sadil` :200-312, `adil` :31-197,
`sadil_updated` :315-501, `learn_coding_vectors` :508-628).  It is the
checker for the CUDA kernels; it is NOT a product path.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.

Parity pin: PINNED.  `oracle/make_golden.py` (and `make_golden_imagenet.py`, `make_golden_adil_fb.py`,
`make_golden_sadil_updated.py`, `make_golden_lcv.py`) ran the UNMODIFIED reference (through
`oracle/ref_shim.py`) in the build container and committed its outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks every function here against those fixtures (bit-exact on
CPU for the projections / AdamW / whole-fit trajectories) and against the known-answer vectors
of SURVEY.md section 4.

Notation: C,H,W image dims; P = C*H*W; K atoms; N images; B minibatch.
`D2` is the dictionary viewed as [P, K] (atoms innermost: adil.py:148 creates [C,H,W,K]).
"""
import math

import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

# synth flags (mirrors include/adil_b200.h)
F_NORMALIZE = 1
F_CLAMP_DELTA = 2
F_CLAMP01 = 4

# projection modes (mirrors include/adil_b200.h)
ROWS_NONE, ROWS_L1BALL, ROWS_L2BALL, ROWS_SOFTSHRINK = 0, 1, 2, 3
ATOMS_NONE, ATOMS_CLAMP1, ATOMS_L2BALL, ATOMS_L2SPHERE, ATOMS_L1BALL = 0, 1, 2, 3, 4


# ------------------------------------------------------------------------------------------------
# projections  (attacks/utils.py:17-57,159-161 ; adil.py:625-642)
# ------------------------------------------------------------------------------------------------
def project_rows_l1(v, radius):
    """Row-wise Euclidean projection onto {||.||_1 <= radius}  (utils.py:21-41, Duchi et al. 2008).

    Rows strictly inside the ball are left untouched (utils.py:33, strict '<')."""
    shape = v.shape
    x = v.reshape(shape[0], -1)
    n, k = x.shape
    a = x.abs()
    inside = (torch.norm(x, p=1, dim=1) < radius)                      # utils.py:33 (strict)
    mu = torch.sort(a, dim=1, descending=True).values                  # utils.py:34
    csum = torch.cumsum(mu, dim=1)                                     # utils.py:35
    j = torch.arange(1, k + 1, device=x.device)
    ok = (mu * j > (csum - radius))                                    # utils.py:37
    rho = (ok * j).max(dim=1).values                                   # largest j with ok (>=1 when the row is outside)
    theta = (csum[torch.arange(n, device=x.device), rho - 1] - radius) / rho   # utils.py:38 (rho==0 -> index -1, value unused)
    proj = (a - theta.unsqueeze(1)).clamp(min=0) * torch.sign(x)       # utils.py:39-40
    m = inside.float().unsqueeze(1)
    out = m * x + (1 - m) * proj                                       # utils.py:40 (arithmetic blend, keeps NaN semantics)
    return out.reshape(shape)


def project_rows_l2(v, radius):
    """adil.py:626-629: radius * v / max(||v||_2, radius) row-wise."""
    nrm = torch.norm(v, p='fro', dim=1, keepdim=True)
    return radius * torch.div(v, torch.maximum(nrm, radius * torch.ones_like(nrm)))


def softshrink(v, lam):
    """utils.py:159-161 (nn.Softshrink): sign(v) * max(|v| - lam, 0)."""
    return torch.where(v > lam, v - lam, torch.where(v < -lam, v + lam, torch.zeros_like(v)))


def project_rows(v, mode, radius):
    if mode == ROWS_NONE:
        return v.clone()
    if mode == ROWS_L1BALL:
        return project_rows_l1(v, radius)
    if mode == ROWS_L2BALL:
        return project_rows_l2(v, radius)
    if mode == ROWS_SOFTSHRINK:
        return softshrink(v, radius)
    raise ValueError(mode)


def project_atoms(D, mode):
    """Per-atom projection of D[..., K]  (utils.py:44-57 ; adil.py:33-35,635-642).  Returns a new tensor."""
    K = D.shape[-1]
    D2 = D.reshape(-1, K).clone()
    if mode == ATOMS_NONE:
        pass
    elif mode == ATOMS_CLAMP1:
        D2 = D2.clamp(min=-1, max=1)                                   # adil.py:35,642
    elif mode in (ATOMS_L2BALL, ATOMS_L2SPHERE):
        for k in range(K):                                             # utils.py:47-54 (one atom at a time)
            col = D.reshape(-1, K)[:, k].reshape(D.shape[:-1])
            nrm = torch.norm(col, p='fro')
            den = nrm if mode == ATOMS_L2SPHERE else torch.maximum(nrm, torch.ones_like(nrm))
            D2[:, k] = torch.div(col, den).reshape(-1)
    elif mode == ATOMS_L1BALL:
        for k in range(K):                                             # utils.py:56: rows of the [C, H*W... ] view
            col = D.reshape(-1, K)[:, k].reshape(D.shape[:-1])
            D2[:, k] = project_rows_l1(col, 1).reshape(-1)
    else:
        raise ValueError(mode)
    return D2.reshape(D.shape)


def clamp_image(x, max_val=1, min_val=0):
    return torch.clamp(x, min=min_val, max=max_val)                    # utils.py:17-18


# ------------------------------------------------------------------------------------------------
# synthesis / normalisation / backward contractions  (adil.py:24-27 ; demo_dL_attack.py:16-25)
# ------------------------------------------------------------------------------------------------
def channel_vec(vals, C, hw, dtype):
    t = torch.as_tensor(vals, dtype=dtype)
    return t.reshape(C, 1).expand(C, hw).reshape(-1)                   # [P], channel-major like NCHW


def synth(x, D2, v, v_index, mean=None, std=None, eps=0.0, flags=0, x_index=None, hw=None):
    """delta = v[v_index] . D2^T ; out = f(x + delta).  x: [B,P] or [N,P] with x_index; returns (out, delta)."""
    vb = v[v_index] if v_index is not None else v
    delta = vb @ D2.t()                                                # adil.py:25 (tensordot -> mm)
    if flags & F_CLAMP_DELTA:
        delta = delta.clamp(min=-eps, max=eps)                         # adil.py:482
    if x is None:
        out = delta.clone()
    else:
        xb = x[x_index] if x_index is not None else x
        out = xb + delta                                               # adil.py:26
    if flags & F_CLAMP01:
        out = out.clamp(min=0, max=1)                                  # adil.py:484,567,623
    if flags & F_NORMALIZE:
        C = len(mean)
        hw = hw if hw is not None else D2.shape[0] // C
        mvec = channel_vec(mean, C, hw, out.dtype)
        svec = channel_vec(std, C, hw, out.dtype)
        out = (out - mvec) / svec                                      # demo_dL_attack.py:22-25 (true division)
    return out, delta


def grad(g, D2, vb, std=None, hw=None):
    """Backward of synth w.r.t. D2 and the batch codes.  g: [B,P] gradient w.r.t. the (normalised) classifier
    input.  Returns (dD2 [P,K], dvb [B,K])."""
    if std is not None:
        C = len(std)
        hw = hw if hw is not None else D2.shape[0] // C
        gx = g / channel_vec(std, C, hw, g.dtype)                      # Normalize backward: div by std
    else:
        gx = g
    dD2 = gx.t() @ vb                                                  # contraction over B
    dvb = gx @ D2                                                      # contraction over P
    return dD2, dvb


# ------------------------------------------------------------------------------------------------
# AdamW  (torch/optim/adam.py single-tensor path; adil.py:154,186,250-251,588)
# ------------------------------------------------------------------------------------------------
def adamw_scalars(t, lr, beta1=0.9, beta2=0.999):
    """Python-float (fp64) scalars exactly as torch computes them for step count t (1-based)."""
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    return lr / bc1, bc2 ** 0.5


def adamw_step_(p, g, m, s, t, lr, beta1=0.9, beta2=0.999, eps=1e-8, wd=1e-2):
    """In-place AdamW step (decoupled weight decay), op order of torch.optim.adam._single_tensor_adam."""
    step_size, bc2_sqrt = adamw_scalars(t, lr, beta1, beta2)
    p.mul_(1 - lr * wd)
    m.lerp_(g, 1 - beta1)
    s.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    denom = (s.sqrt() / bc2_sqrt).add_(eps)
    p.addcdiv_(m, denom, value=-step_size)
    return p


# ------------------------------------------------------------------------------------------------
# one learning step on explicit state (SURVEY.md section 8(a'))
# ------------------------------------------------------------------------------------------------
class State(object):
    """Flattened learnables + optimizer state."""

    def __init__(self, D, v):
        self.shape = tuple(D.shape)                                    # [C,H,W,K]
        self.K = D.shape[-1]
        self.D2 = D.reshape(-1, self.K).clone()
        self.v = v.clone()
        self.mD = torch.zeros_like(self.D2)
        self.sD = torch.zeros_like(self.D2)
        self.mv = torch.zeros_like(self.v)
        self.sv = torch.zeros_like(self.v)
        self.tD = 0
        self.tv = 0

    def D(self):
        return self.D2.reshape(self.shape)


def dict_step_(st, dD2, lr, atoms_mode=ATOMS_CLAMP1, wd=1e-2):
    st.tD += 1
    adamw_step_(st.D2, dD2, st.mD, st.sD, st.tD, lr, wd=wd)
    st.D2.copy_(project_atoms(st.D2.reshape(st.shape), atoms_mode).reshape(-1, st.K))


def code_step_(st, dvb, v_index, lr, radius, rows_mode=ROWS_L1BALL, wd=1e-2):
    """AdamW on ALL rows of v (dense grad, zero outside the batch: adil.py:154,186) then row projection."""
    gV = torch.zeros_like(st.v)
    gV.index_put_((torch.as_tensor(v_index, dtype=torch.long).to(st.v.device),), dvb, accumulate=True)
    st.tv += 1
    adamw_step_(st.v, gV, st.mv, st.sv, st.tv, lr, wd=wd)
    st.v.copy_(project_rows(st.v, rows_mode, radius))


def joint_step_(st, g, v_index, lr, radius, std=None):
    """'gd' method step given the classifier input-gradient g [B,P]  (adil.py:185-188)."""
    vb = st.v[torch.as_tensor(v_index, dtype=torch.long)]
    dD2, dvb = grad(g, st.D2, vb, std)
    dict_step_(st, dD2, lr)
    code_step_(st, dvb, v_index, lr, radius)
    return dD2, dvb


# ------------------------------------------------------------------------------------------------
# losses  (adil.py:103-112,136)
# ------------------------------------------------------------------------------------------------
def f_loss(outputs, labels, kappa, targeted=False):
    onehot = torch.eye(outputs.shape[1], device=outputs.device)[labels]
    i = ((1 - onehot) * outputs).max(dim=1).values                     # label slot zeroed, not -inf (adil.py:106)
    j = torch.masked_select(outputs, onehot.bool())
    return torch.clamp(i - j, min=-kappa) if targeted else torch.clamp(j - i, min=-kappa)


def attack_loss(outputs, labels, loss, kappa, targeted, reduction):
    coeff = 1.0 if targeted else -1.0
    if loss == 'ce':
        return coeff * torch.nn.functional.cross_entropy(outputs, labels, reduction=reduction)
    if loss == 'logits':
        return f_loss(outputs, labels, kappa, targeted).sum()
    raise ValueError(loss)


# ------------------------------------------------------------------------------------------------
# drivers (flattened restatements of adil.py:114-210, 212-332, 569-623, 508-567, 460-506)
# ------------------------------------------------------------------------------------------------
def split_normalize(model):
    """If model = Sequential(Normalize-like, net) return (net, mean, std) else (model, None, None)."""
    if isinstance(model, torch.nn.Sequential) and len(model) >= 1:
        first = model[0]
        if hasattr(first, 'mean') and hasattr(first, 'std') and not list(first.parameters()):
            rest = model[1] if len(model) == 2 else torch.nn.Sequential(*list(model)[1:])
            return rest, [float(a) for a in first.mean.reshape(-1)], [float(a) for a in first.std.reshape(-1)]
    return model, None, None


def classifier_grad(model, xin, labels, loss, kappa, targeted, reduction):
    """Loss value and d loss / d xin through the (frozen) classifier."""
    xin = xin.detach().requires_grad_(True)
    out = model(xin)
    val = attack_loss(out, labels, loss, kappa, targeted, reduction)
    (g,) = torch.autograd.grad(val, xin)
    return val.detach(), g, out.detach()


def init_state(nc, nx, ny, n_img, n_atoms, eps, norm='linf', v_zero=False, device='cpu'):
    """adil.py:145-150 / 242-246: same RNG draws, same order (D first, then v)."""
    if norm == 'l2':
        D = project_atoms(torch.randn(nc, nx, ny, n_atoms, device=device), ATOMS_L2BALL)
    else:
        D = -1 + 2 * torch.rand(nc, nx, ny, n_atoms, device=device)
    v0 = torch.zeros(n_img, n_atoms, device=device) if v_zero else torch.rand(n_img, n_atoms, device=device)
    v = project_rows_l2(v0, eps) if norm == 'l2' else project_rows_l1(v0, eps)
    return State(D, v)


def learn_dictionary_a(model, dataset, eps, steps, n_atoms, batch_size, step_size=0.01, norm='linf', loss='ce',
                       kappa=50, targeted=False, state=None, val=None, fused_normalize=False, val_coder=True,
                       on_step=None):
    """Joint ('gd') fit.  Returns (state, loss_all, fooling_rate_all, val_fooling_rate).

    `val` (optional dataset of (x, y)) reproduces the per-epoch validation coder of adil.py:198-205, including
    the CPU-RNG draws its shuffling DataLoader makes.

    `dataset[i]` -> (x, y) with `.indexed` False, (i, x, y) with True (imagenet_loading.py:8-18).
    With `fused_normalize` the leading Normalize module is peeled off `model` and applied by `synth` / `grad`
    (what the CUDA path does); otherwise the model is called on x+delta exactly like adil.py:26.

    `val_coder=False` iterates the validation loader (same CPU-RNG draws) but skips the 100-iteration coder, like the
    stubbed reference runs of oracle/make_golden_imagenet.py.  `on_step(st, index, x, xin, g, loss, phase)` is called
    with phase 'before' (state untouched, classifier gradient known) and 'after' (state stepped) on every minibatch."""
    dataset.indexed = False
    n_img = len(dataset)
    x0, _ = next(iter(dataset))
    nc, nx, ny = x0.shape
    P = nc * nx * ny
    bs = n_img if batch_size is None else batch_size
    dataset.indexed = True
    loader = torch.utils.data.DataLoader(dataset, batch_size=bs, shuffle=True, num_workers=0)
    val_loader = None if val is None else torch.utils.data.DataLoader(val, batch_size=bs, shuffle=True, num_workers=0)
    st = state if state is not None else init_state(nc, nx, ny, n_img, n_atoms, eps, norm)
    net, mean, std = split_normalize(model) if fused_normalize else (model, None, None)
    flags = F_NORMALIZE if mean is not None else 0
    loss_all, fool_all = [], []
    for it in range(int(steps)):
        loss_full = 0.0
        fooled = 0
        for index, x, _ in loader:
            with torch.no_grad():
                label = model(x).argmax(dim=-1)                        # adil.py:172
            xin, _ = synth(x.reshape(len(index), P), st.D2, st.v, index, mean, std, eps, flags)
            lval, g, out = classifier_grad(net, xin.reshape(x.shape), label, loss, kappa, targeted, 'sum')
            fooled += int((out.argmax(dim=-1) != label).sum())
            if on_step is not None:
                on_step(st, index, x, xin, g, lval, 'before')
            joint_step_(st, g.reshape(len(index), P), index, step_size, eps, std)
            if on_step is not None:
                on_step(st, index, x, xin, g, lval, 'after')
            loss_full = loss_full + lval
        loss_all.append(float(loss_full) / n_img)
        fool_all.append(fooled / n_img)
        val_fool = None
        if val_loader is not None:
            val_fool = 0
            for xv, _ in val_loader:
                if val_coder:
                    val_fool = val_fool + coder_adamw(model, xv, st.D(), eps, norm, loss, kappa, targeted,
                                                      mode='train', fused_normalize=fused_normalize)
            val_fool = val_fool / len(val)
        if it > 1 and abs(loss_all[it] - loss_all[it - 1]) < 1e-6:     # adil.py:207
            break
    return st, loss_all, fool_all, val_fool


def learn_dictionary_b(model, dataset, eps, steps, steps_inner, n_atoms, batch_size, step_size=0.01, norm='linf',
                       loss='ce', kappa=50, targeted=False, state=None, val=None, fused_normalize=False):
    """Alternating ('alter') fit, adil.py:212-332: v-epochs (lr) then D-epochs (2*lr), independent AdamW counters."""
    dataset.indexed = False
    n_img = len(dataset)
    x0, _ = next(iter(dataset))
    nc, nx, ny = x0.shape
    P = nc * nx * ny
    bs = n_img if batch_size is None else batch_size
    dataset.indexed = True
    loader = torch.utils.data.DataLoader(dataset, batch_size=bs, shuffle=True, num_workers=0)
    val_loader = None if val is None else torch.utils.data.DataLoader(val, batch_size=bs, shuffle=True, num_workers=0)
    st = state if state is not None else init_state(nc, nx, ny, n_img, n_atoms, eps, norm, v_zero=True)
    net, mean, std = split_normalize(model) if fused_normalize else (model, None, None)
    flags = F_NORMALIZE if mean is not None else 0
    loss_all, fool_all = [], []

    def one_batch(index, x, which):
        with torch.no_grad():
            label = model(x).argmax(dim=-1)
        xin, _ = synth(x.reshape(len(index), P), st.D2, st.v, index, mean, std, eps, flags)
        lval, g, out = classifier_grad(net, xin.reshape(x.shape), label, loss, kappa, targeted, 'sum')
        vb = st.v[index]
        dD2, dvb = grad(g.reshape(len(index), P), st.D2, vb, std)
        if which == 'v':
            code_step_(st, dvb, index, step_size, eps)
        else:
            dict_step_(st, dD2, 2 * step_size)
        return lval, int((out.argmax(dim=-1) != label).sum())

    for it in range(int(steps // steps_inner)):
        for _ in range(steps_inner):
            for index, x, _ in loader:
                one_batch(index, x, 'v')
        for _ in range(steps_inner):
            fooled = 0
            for index, x, _ in loader:
                lval, f = one_batch(index, x, 'd')
                fooled += f
            loss_full = lval                                            # adil.py:313-314: only the last batch's loss
        loss_all.append(float(loss_full) / n_img)
        fool_all.append(fooled / n_img)
        val_fool = None
        if val_loader is not None:
            val_fool = 0
            for xv, _ in val_loader:
                val_fool = val_fool + coder_adamw(model, xv, st.D(), eps, norm, loss, kappa, targeted, mode='train',
                                                  fused_normalize=fused_normalize)
            val_fool = val_fool / len(val)
        if it > 1 and abs(loss_all[it] - loss_all[it - 1]) < 1e-6:
            break
    return st, loss_all, fool_all, val_fool


def coder_adamw(model, images, D, eps, norm='linf', loss='ce', kappa=50, targeted=False, iters=100, mode='train',
                fused_normalize=False):
    """forward_supervised_AdamW, adil.py:569-623: v-only AdamW(lr=1e-2) with D frozen, CE mean reduction."""
    n = images.shape[0]
    K = D.shape[-1]
    P = images[0].numel()
    D2 = D.reshape(-1, K)
    st = State(D, torch.zeros(n, K))
    net, mean, std = split_normalize(model) if fused_normalize else (model, None, None)
    flags = F_NORMALIZE if mean is not None else 0
    idx = torch.arange(n)
    labels = None
    for _ in range(iters):
        with torch.no_grad():
            labels = model(images).argmax(dim=-1)
        xin, _ = synth(images.reshape(n, P), D2, st.v, idx, mean, std, eps, flags)
        _, g, _ = classifier_grad(net, xin.reshape(images.shape), labels, loss, kappa, targeted, 'mean')
        _, dvb = grad(g.reshape(n, P), D2, st.v, std)
        v_old = st.v.clone()
        code_step_(st, dvb, idx, 1e-2, eps)                            # update_v is always the l1 ball (adil.py:29-31)
        if (st.v - v_old).abs().max() < 1e-6:
            break
    vproj = project_rows_l2(st.v, eps) if norm == 'l2' else project_rows_l1(st.v, eps)   # adil.py:617
    dv = (vproj @ D2.t()).reshape(images.shape)
    if mode == 'train':
        with torch.no_grad():
            return (model(images + dv).argmax(-1) != labels).sum()
    return torch.clamp(images + dv, min=0, max=1)


def ddrague(model, images, D, eps, steps_inference=30, loss='ce', kappa=50, targeted=False):
    """forward_supervised_DDrague, adil.py:508-567: optimise z (image-shaped), delta = D D^+ z."""
    n = images.shape[0]
    K = D.shape[-1]
    P = images[0].numel()
    D2 = D.reshape(-1, K)
    gram = D2.t() @ D2                                                 # adil.py:523
    pinv2 = (gram.inverse() @ D2.t()).t().contiguous()                 # [P,K]: d_drg viewed like D2 (adil.py:524-525)
    z = torch.zeros(n, P)
    mz, sz = torch.zeros_like(z), torch.zeros_like(z)
    for t in range(1, int(steps_inference) + 1):
        with torch.no_grad():
            labels = model(images).argmax(dim=-1)
        v = z @ pinv2                                                  # adil.py:542
        dv = v @ D2.t()                                                # adil.py:543
        _, g, _ = classifier_grad(model, (images.reshape(n, P) + dv).reshape(images.shape), labels, loss, kappa,
                                  targeted, 'mean')
        gv = g.reshape(n, P) @ D2
        gz = gv @ pinv2.t()
        z_old = z.clone()
        adamw_step_(z, gz, mz, sz, t, 1e-2)
        z.clamp_(min=-eps, max=eps)                                    # adil.py:555
        if (z - z_old).abs().max() < 1e-6:
            break
    dv = (z @ pinv2) @ D2.t()
    return torch.clamp(images + dv.reshape(images.shape), min=0, max=1)


def sample_sphere(n_samples, n_atoms, eps, norm='linf'):
    """adil.py:644-655 (same RNG draws)."""
    if norm == 'l2':
        var = 2 * torch.rand(n_samples, n_atoms) - 1
        return eps * torch.div(var, torch.norm(var, p='fro', dim=1, keepdim=True))
    m = torch.distributions.uniform.Uniform(torch.tensor([eps]), torch.tensor([2 * eps]))
    raw = m.sample(sample_shape=[n_samples, n_atoms])[:, :, 0]
    return project_rows_l1(raw, eps)


def unsupervised(model, images, D, eps, trials=10, norm='linf'):
    """forward_unsupervised, adil.py:460-506.  Returns (adv_best, dv_norm_inf of the last trial)."""
    n = images.shape[0]
    K = D.shape[-1]
    P = images[0].numel()
    D2 = D.reshape(-1, K)
    flag = torch.zeros(n, dtype=torch.bool)
    best_fool = float('inf') * torch.ones(n)
    best_nofool = float('inf') * torch.ones(n)
    adv_best = images.clone()
    dv_norm_inf = []
    for _ in range(int(trials)):
        v = sample_sphere(n, K, eps, norm)
        rows = [synth(images[i:i + 1].reshape(1, P), D2, v[i:i + 1], None, eps=eps, flags=F_CLAMP_DELTA | F_CLAMP01)
                for i in range(n)]                                     # adil.py:480-484: one sample at a time
        adv = torch.cat([r[0] for r in rows])
        delta = torch.cat([r[1] for r in rows])
        dv_norm_inf = [float(r.abs().max()) for r in delta]
        adv = adv.reshape(images.shape)
        with torch.no_grad():
            adv_labels = model(adv).argmax(dim=1)
            pre_labels = model(images).argmax(dim=1)
        fooling = adv_labels != pre_labels
        mse = ((images - adv) ** 2).sum(dim=[1, 2, 3])
        for i in range(n):
            if not flag[i] and fooling[i]:
                flag[i] = True
                best_fool[i] = mse[i]                                  # adil.py:494-495 (stored, never compared)
                adv_best[i] = adv[i]
            elif (flag[i] and fooling[i]) or (not flag[i] and not fooling[i]):
                if mse[i] < best_nofool[i]:
                    best_nofool[i] = mse[i]
                    adv_best[i] = adv[i]
    return adv_best, dv_norm_inf


# ------------------------------------------------------------------------------------------------
# synthetic data / models shared by tests, smoke and bench (no reference code involved)
# ------------------------------------------------------------------------------------------------
class IndexedTensorDataset(torch.utils.data.Dataset):
    """Synthetic stand-in for imagenet_loading.Subset_I: `.indexed` toggles (x,y) <-> (item,x,y)."""

    def __init__(self, images, labels, indexed=False):
        self.images, self.labels, self.indexed = images, labels, indexed

    def __len__(self):
        return len(self.images)

    def __getitem__(self, item):
        if self.indexed:
            return item, self.images[item], self.labels[item]
        return self.images[item], self.labels[item]


class Normalize(torch.nn.Module):
    """Same arithmetic as demo_dL_attack.py:16-25: (input - mean) / std per channel."""

    def __init__(self, mean=IMAGENET_MEAN, std=IMAGENET_STD):
        super().__init__()
        self.register_buffer('mean', torch.tensor(mean, dtype=torch.float32))
        self.register_buffer('std', torch.tensor(std, dtype=torch.float32))

    def forward(self, x):
        return (x - self.mean.reshape(1, -1, 1, 1)) / self.std.reshape(1, -1, 1, 1)


def tiny_classifier(seed=0, n_classes=10, width=8):
    """Small conv net used where the test only needs *a* differentiable classifier."""
    g = torch.Generator().manual_seed(seed)
    net = torch.nn.Sequential(
        torch.nn.Conv2d(3, width, 3, padding=1), torch.nn.Tanh(),
        torch.nn.Conv2d(width, width, 3, padding=1, stride=2), torch.nn.Tanh(),
        torch.nn.AdaptiveAvgPool2d(2), torch.nn.Flatten(), torch.nn.Linear(4 * width, n_classes))
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.5 if p.dim() > 1 else 0.1))
    return torch.nn.Sequential(Normalize(), net).eval()


# ------------------------------------------------------------------------------------------------
# regularised variant (attacks/attacks_classes/adil_regularized.py): SADiL, the stochastic forward-backward scheme
# ------------------------------------------------------------------------------------------------
def get_target(model, x, y, targeted):
    """utils.py:164-174: second most probable class of the clean image (targeted) or the given label."""
    with torch.no_grad():
        if targeted:
            return model(x).sort().indices[:, -2]
        return y


def penalised_grad(model, x, D2, vb, target, coeff, l2_fool):
    """Gradients of coeff * CE_sum(model(x + D vb), target) + 0.5 * l2_fool * ||D vb||^2 w.r.t. D2 and vb
    (adil_regularized.py:273-279,291-297): gx = d CE / d(x + dv) + l2_fool * dv ; dD2 = gx^T vb ; dvb = gx D2."""
    n, P = x.shape[0], D2.shape[0]
    dv = vb @ D2.t()
    xin = (x.reshape(n, P) + dv).reshape(x.shape).detach().requires_grad_(True)
    loss = coeff * torch.nn.functional.cross_entropy(model(xin), target, reduction='sum')
    (g,) = torch.autograd.grad(loss, xin)
    gx = g.reshape(n, P) + l2_fool * dv
    return gx.t() @ vb, gx @ D2, float(loss) + 0.5 * l2_fool * float((dv ** 2).sum())


def sadil_loss(model, loader, slices, D2, v, coeff, l2_fool, lambda_coding, targeted):
    """loss_all of adil_regularized.py:245-254 (loss-only pass over the whole set)."""
    total = 0.0
    with torch.no_grad():
        for i, (x, y) in enumerate(loader):
            n = x.shape[0]
            dv = v[slices[i]] @ D2.t()
            out = model((x.reshape(n, -1) + dv).reshape(x.shape))
            total += (coeff * torch.nn.functional.cross_entropy(out, get_target(model, x, y, targeted), reduction='sum')
                      + 0.5 * l2_fool * torch.sum(dv ** 2)).item()
    return total + (lambda_coding * torch.sum(torch.abs(v))).item()


def sadil(model, dataset, targeted=True, nepochs=10, batchsize=1, lambda_coding=1., l2_fool=1., stepsize=1., n_atom=5,
          dict_set='l2ball', D0=None):
    """SADiL, adil_regularized.py:200-312: per minibatch a gradient step on D followed by the per-atom projection
    (constraint_dict), then -- with the new D -- a proximal gradient step (soft threshold stepsize * lambda) on the
    code rows of the batch.  Returns (D [C,H,W,K], v [N,K], loss list)."""
    nimg = len(dataset)
    x0, _ = next(iter(dataset))
    nc, nx, ny = x0.shape
    P = nc * nx * ny
    loader = torch.utils.data.DataLoader(dataset, batch_size=batchsize, shuffle=False)
    coeff = 1. if targeted else -1.
    slices = [list(range(i, min(i + batchsize, nimg))) for i in range(0, nimg, batchsize)]   # utils.py:153-156
    mode = {'l2ball': ATOMS_L2BALL, 'l2sphere': ATOMS_L2SPHERE}.get(dict_set, ATOMS_L1BALL)
    D = project_atoms(torch.randn(3, nx, ny, n_atom), mode) if D0 is None else D0.clone()
    D2 = D.reshape(P, n_atom).clone()
    v = torch.zeros(nimg, n_atom)
    # Reference behaviour (adil_regularized.py:287-304): `v` becomes a leaf that requires grad at the first V-step and
    # its `.grad` is never zeroed, so every later backward -- the D-step's (v takes part in D v) and the V-step's --
    # ACCUMULATES into it; the V-step then uses the accumulated rows `grad_v[ind]`.
    gv_acc = torch.zeros(nimg, n_atom)
    v_has_grad = False
    loss = [sadil_loss(model, loader, slices, D2, v, coeff, l2_fool, lambda_coding, targeted)]
    for _ in range(int(nepochs)):
        for i, (x, y) in enumerate(loader):
            ind = slices[i]
            target = get_target(model, x, y, targeted)
            dD2, dvb, _ = penalised_grad(model, x, D2, v[ind], target, coeff, l2_fool)        # D-step
            if v_has_grad:
                gv_acc[ind] += dvb
            D2 = project_atoms((D2 - stepsize * dD2).reshape(nc, nx, ny, n_atom), mode).reshape(P, n_atom)
            _, dvb, _ = penalised_grad(model, x, D2, v[ind], target, coeff, l2_fool)          # V-step, new D
            v_has_grad = True
            gv_acc[ind] += dvb
            v[ind] = softshrink(v[ind] - stepsize * gv_acc[ind], stepsize * lambda_coding)
        loss.append(sadil_loss(model, loader, slices, D2, v, coeff, l2_fool, lambda_coding, targeted))
        if abs(loss[-1] - loss[-2]) < 1e-6:
            break
    return D2.reshape(nc, nx, ny, n_atom), v, loss


# ------------------------------------------------------------------------------------------------
# regularised variant, full batch: ADiL with backtracking line search (adil_regularized.py:31-197)
# ------------------------------------------------------------------------------------------------
def adil_fb(model, dataset, targeted=True, niter=10, lambda_coding=1., l2_fool=1., batchsize=None, step_size=.1, n_atom=10,
            dict_set='l2ball', D0=None, learn_dictionary=True, trace=None):
    """Full-batch forward-backward scheme with the backtracking of Bonettini et al. (adil_regularized.py:31-197).

    Per iteration: the penalised loss and its gradients over the WHOLE set (:109-120); from the third iteration on a
    Barzilai-Borwein-like Lipschitz estimate ||delta grad|| / ||delta (v, D)|| (:126-130) and step 0.9 / L (:140);
    v <- soft threshold(v - step grad_v, step * lambda), D <- constraint_dict(D - step grad_D) (:143-147); then the
    line search over the segment (v_old, D_old) -> (v, D): the first i in 0..50 with
    loss(v_old + 0.5^i dv, D_old + 0.5^i dD) <= loss_old + 0.5 * 0.5^i * h is accepted (:158-188); none: stop (:189-192).
    `learn_dictionary=False` is the reference's `dictionary is not None` mode (D fixed, :103-105,119,145).
    `trace` (a list) receives the accepted line-search index of every iteration (51: stopped).
    Returns (D [C,H,W,K], v [N,K], loss_all [niter] with NaN for iterations that never ran)."""
    import numpy as np
    nimg = len(dataset)
    x0, _ = next(iter(dataset))
    nc, nx, ny = x0.shape
    P = nc * nx * ny
    if batchsize is None:
        batchsize = nimg
    delta, gamma, beta = .5, 1., .5
    lipschitz = .9 / step_size
    coeff = 1. if targeted else -1.
    slices = [list(range(i, min(i + batchsize, nimg))) for i in range(0, nimg, batchsize)]   # utils.py:153-156
    loader = torch.utils.data.DataLoader(dataset, batch_size=batchsize, shuffle=False)
    mode = {'l2ball': ATOMS_L2BALL, 'l2sphere': ATOMS_L2SPHERE}.get(dict_set, ATOMS_L1BALL)
    D = project_atoms(torch.randn(3, nx, ny, n_atom), mode) if D0 is None else D0.clone()
    D2 = D.reshape(P, n_atom).clone()
    v = torch.zeros(nimg, n_atom)

    def smooth_loss(D2_, v_, want_grad):
        """sum over the batches of coeff * CE_sum + 0.5 * l2_fool * ||D v||^2 (:112-117); optionally its gradients"""
        total = torch.zeros(())
        gD = torch.zeros_like(D2_) if want_grad else None
        gv = torch.zeros_like(v_) if want_grad else None
        for i, (x, y) in enumerate(loader):
            ind = slices[i]
            target = get_target(model, x, y, targeted)
            if want_grad:
                # (same accumulation order as the reference's fp32 tensors: (loss + coeff * CE) + 0.5 * l2 * ||dv||^2)
                n = x.shape[0]
                vb = v_[ind]
                dv = vb @ D2_.t()
                xin = (x.reshape(n, P) + dv).reshape(x.shape).detach().requires_grad_(True)
                ce = coeff * torch.nn.functional.cross_entropy(model(xin), target, reduction='sum')
                (g,) = torch.autograd.grad(ce, xin)
                gx = g.reshape(n, P) + l2_fool * dv
                gD += gx.t() @ vb
                gv[ind] = gx @ D2_
                total = total + ce.detach() + .5 * l2_fool * torch.sum(dv ** 2)
            else:
                with torch.no_grad():
                    n = x.shape[0]
                    dv = v_[ind] @ D2_.t()
                    out = model((x.reshape(n, -1) + dv).reshape(x.shape))
                    total = total + coeff * torch.nn.functional.cross_entropy(out, target, reduction='sum') \
                        + .5 * l2_fool * torch.sum(dv ** 2)
        return total, gD, gv

    D_old, v_old = torch.zeros_like(D2), torch.zeros_like(v)
    gD_old, gv_old = torch.zeros_like(D2), torch.zeros_like(v)
    loss_all = np.nan * np.ones(int(niter))
    loss_ns_old = torch.zeros(())
    stop = False
    for it in range(int(niter)):
        if stop:
            continue
        loss_ns = lambda_coding * torch.sum(torch.abs(v))
        loss_s, gD, gv = smooth_loss(D2, v, True)
        if not learn_dictionary:
            gD = torch.zeros_like(D2)
        loss_full = loss_s + loss_ns
        if it > 1:
            num = torch.sqrt(torch.norm(gv - gv_old) ** 2 + torch.norm(gD - gD_old) ** 2)
            lipschitz = num / torch.sqrt(torch.norm(v - v_old) ** 2 + torch.norm(D2 - D_old) ** 2)
        D_old, v_old, gv_old, gD_old = D2.clone(), v.clone(), gv.clone(), gD.clone()
        loss_old = loss_full
        step = .9 / lipschitz
        v = softshrink(v - step * gv, step * lambda_coding)
        if learn_dictionary:
            D2 = project_atoms((D2 - step * gD).reshape(nc, nx, ny, n_atom), mode).reshape(P, n_atom)
        d_v, d_d = v - v_old, D2 - D_old
        h = torch.sum(d_d * gD) + torch.sum(d_v * gv) + .5 * (gamma / step) * (torch.norm(d_d) ** 2 + torch.norm(d_v) ** 2) \
            + loss_ns - loss_ns_old
        i = 0
        while True:
            new_v, new_D = v_old + (delta ** i) * d_v, D_old + (delta ** i) * d_d
            loss_ns = lambda_coding * torch.sum(torch.abs(new_v))
            loss_s, _, _ = smooth_loss(new_D, new_v, False)
            loss_full = loss_s + loss_ns
            if loss_full <= loss_old + beta * (delta ** i) * h:
                v, D2 = new_v, new_D
                loss_ns_old = loss_ns
                break
            i += 1
            if i > 50:
                stop = True
                break
        if trace is not None:
            trace.append(i)
        loss_all[it] = float(loss_full)
    return D2.reshape(nc, nx, ny, n_atom), v, loss_all


# ------------------------------------------------------------------------------------------------
# regularised variant: the coder on a fixed dictionary (learn_coding_vectors, adil_regularized.py:508-628)
# ------------------------------------------------------------------------------------------------
def learn_coding_vectors(model, dataset, D, targeted=True, niter=10, lambda_l1=1., lambda_l2=1., batch_size=None,
                         step_size=.1, trace=None):
    """Codes of `dataset` on the FIXED dictionary D [C,H,W,K] for the penalised objective
    coeff * CE_sum(x + D v) + 0.5 * lambda_l2 * ||D v||^2 + lambda_l1 * ||v||_1  (adil_regularized.py:508-628): full-batch
    proximal gradient from v = 0 (:538) with a backtracking line search over the segment v_old -> prox step (:573-622).

    Per iteration: loss and gradient over the whole set (:553-565); v <- soft threshold(v - step grad, step * lambda_l1)
    (:574-577); h = <dv, grad> + 0.5 / step * ||dv||^2 + lambda_l1 (|v|_1 - |v_old|_1) (:583-584); then the first i in
    0..10 with loss(v_old + 0.9^i dv) <= loss_old + 0.5 * 0.9^i * h ends the search (:589-613) -- the shortened point is
    KEPT only if its loss is below the full step's (`loss_cur > loss_full`, :607-610: then the step size shrinks by
    0.9^i too), otherwise the full prox step stays and its loss is recorded; no i <= 10 qualifies: the last point tried
    is taken (:615-620).  Stops when the recorded loss decreased by less than 1e-6 (:626).
    `trace` (a list) receives (index at which the search ended, kept the shortened point) per iteration.
    Returns (v [N,K], loss_all list starting with NaN like the reference's, final step size)."""
    import numpy as np
    nimg = len(dataset)
    nc, nx, ny, n_atom = D.shape
    P = nc * nx * ny
    D2 = D.reshape(P, n_atom)
    delta, gamma, beta = .9, 1, .5
    batch_size = nimg if batch_size is None else batch_size
    coeff = 1. if targeted else -1.
    slices = [list(range(i, min(i + batch_size, nimg))) for i in range(0, nimg, batch_size)]   # utils.py:153-156
    loader = torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=False)
    step_size = torch.as_tensor(step_size, dtype=torch.float32)   # (the reference's default is a tensor, :509)
    v = torch.zeros(nimg, n_atom)

    def smooth_loss(v_, want_grad):
        """(:553-559 / :597-603) sum over the batches, accumulated as the reference's fp32 tensor expression"""
        total = 0
        gv = torch.zeros_like(v_) if want_grad else None
        for i, (x, y) in enumerate(loader):
            ind = slices[i]
            n = x.shape[0]
            target = get_target(model, x, y, targeted)
            dv = v_[ind] @ D2.t()
            if want_grad:
                xin = (x.reshape(n, P) + dv).reshape(x.shape).detach().requires_grad_(True)
                ce = coeff * torch.nn.functional.cross_entropy(model(xin), target, reduction='sum')
                (g,) = torch.autograd.grad(ce, xin)
                gv[ind] = (g.reshape(n, P) + lambda_l2 * dv) @ D2
                total = total + ce.detach() + .5 * lambda_l2 * torch.sum(dv ** 2)
            else:
                with torch.no_grad():
                    out = model((x.reshape(n, P) + dv).reshape(x.shape))
                    total = total + coeff * torch.nn.functional.cross_entropy(out, target, reduction='sum') \
                        + .5 * lambda_l2 * torch.sum(dv ** 2)
        return total, gv

    loss_all = [np.nan]
    for _ in range(int(niter)):
        loss_s, grad_v = smooth_loss(v, True)
        loss_old = (loss_s + lambda_l1 * torch.sum(torch.abs(v))).item()
        v_old = v.clone()
        v = softshrink(v - step_size * grad_v, float(step_size * lambda_l1))
        d_v = v - v_old
        h = torch.sum((v - v_old) * grad_v) + .5 * (gamma / step_size) * (torch.norm(v - v_old, 'fro') ** 2) \
            + lambda_l1 * torch.sum(torch.abs(v)) - lambda_l1 * torch.sum(torch.abs(v_old))
        index_i, kept = 0, False
        while True:
            new_v = v_old + (delta ** index_i) * d_v
            loss_s, _ = smooth_loss(new_v, False)
            loss_full = (loss_s + lambda_l1 * torch.sum(torch.abs(new_v))).item()
            if index_i == 0:
                loss_cur = loss_full
            if loss_full <= loss_old + beta * (delta ** index_i) * h:
                if loss_cur > loss_full:
                    v = new_v
                    step_size = step_size * delta ** index_i
                    loss_all.append(loss_full)
                    kept = True
                else:
                    loss_all.append(loss_cur)
                break
            index_i += 1
            if index_i > 10:
                v = new_v
                loss_all.append(loss_full)
                kept = True
                break
        if trace is not None:
            trace.append((index_i, kept))
        if loss_all[-2] - loss_all[-1] < 1e-6:
            break
    return v, loss_all, float(step_size)


# ------------------------------------------------------------------------------------------------
# regularised variant: SADiL "updated" (sadil_updated, adil_regularized.py:315-501)
# ------------------------------------------------------------------------------------------------
def sadil_updated(model, dataset, targeted=True, nepochs=10, batchsize=1, lambda_coding=1., l2_fool=1., stepsize=1.,
                  n_atom=5, dict_set='l2ball', D0=None, trace=None):
    """Per epoch: a proximal step on the code rows of every minibatch, then ONE projected gradient step on D with the
    gradient accumulated over the epoch, each followed by a backtracking test (factor 0.5, at most 5 halvings) that only
    adapts the step sizes -- the full steps are always kept (adil_regularized.py:444-448,486-492).

    What the reference's autograd bookkeeping amounts to (restated explicitly here):
      * `v` is one tensor updated in place whose `.grad` is never zeroed (:392,398): every backward -- the V-step's (:398)
        and the D-step's (:446, v takes part in D v) -- accumulates into it, and the V-step uses the accumulated rows;
      * `D` requires grad from the first D-step backward of its lifetime on (:439), so from the SECOND minibatch on the
        V-step backward (old codes) adds to `D.grad` as well as the D-step backward (new codes); a new D tensor (after
        an accepted step, :460) starts without gradient, a skipped step (:451-452) keeps accumulating;
      * the backtracking loop of the V-step measures the l1 term WITHOUT lambda (:433) and its first-order term has
        |v_cur|_1 - |v[ind]|_1 = 0 (:420-421).
    `trace` (a list) receives per epoch (i_max of the V-steps, halvings of the D-step or -1 when it was skipped).
    Returns (D [C,H,W,K], v [N,K], loss list, (stepsize_v, stepsize_D))."""
    nimg = len(dataset)
    x0, _ = next(iter(dataset))
    nc, nx, ny = x0.shape
    P = nc * nx * ny
    delta, beta = 0.5, 0.5
    loader = torch.utils.data.DataLoader(dataset, batch_size=batchsize, shuffle=False)
    coeff = 1. if targeted else -1.
    slices = [list(range(i, min(i + batchsize, nimg))) for i in range(0, nimg, batchsize)]   # utils.py:153-156
    mode = {'l2ball': ATOMS_L2BALL, 'l2sphere': ATOMS_L2SPHERE}.get(dict_set, ATOMS_L1BALL)
    D = project_atoms(torch.randn(3, nx, ny, n_atom), mode) if D0 is None else D0.clone()
    D2 = D.reshape(P, n_atom).clone()
    v = torch.zeros(nimg, n_atom)
    stepsize_D = stepsize_v = stepsize

    def smooth(x, target, vb, D2_, want_grad):
        """coeff * CE_sum + 0.5 * l2_fool * ||D vb||^2 as the reference's fp32 tensor (:395-396) [, d/dD2, d/dvb]"""
        n = x.shape[0]
        dv = vb @ D2_.t()
        if not want_grad:
            with torch.no_grad():
                out = model((x.reshape(n, P) + dv).reshape(x.shape))
                return coeff * torch.nn.functional.cross_entropy(out, target, reduction='sum') + .5 * l2_fool * torch.sum(dv ** 2)
        xin = (x.reshape(n, P) + dv).reshape(x.shape).detach().requires_grad_(True)
        ce = coeff * torch.nn.functional.cross_entropy(model(xin), target, reduction='sum')
        (g,) = torch.autograd.grad(ce, xin)
        gx = g.reshape(n, P) + l2_fool * dv
        return ce.detach() + .5 * l2_fool * torch.sum(dv ** 2), gx.t() @ vb, gx @ D2_

    def loss_all(v_, D2_):
        """:362-373"""
        total = 0
        for i, (x, y) in enumerate(loader):
            total += smooth(x, get_target(model, x, y, targeted), v_[slices[i]], D2_, False).item()
        return total + (lambda_coding * torch.sum(torch.abs(v_))).item()

    loss = [loss_all(v, D2)]
    gv_acc = torch.zeros(nimg, n_atom)            # v.grad
    gD_acc = torch.zeros(P, n_atom)               # D.grad of the current D tensor
    D_has_grad = False
    for _ in range(int(nepochs)):
        i_max = 0
        for bi, (x, y) in enumerate(loader):
            ind = slices[bi]
            target = get_target(model, x, y, targeted)
            # ---------- V-step (:391-448) ----------
            loss_s, dD2, dvb = smooth(x, target, v[ind], D2, True)
            gv_acc[ind] += dvb
            if D_has_grad:
                gD_acc += dD2
            g_rows = gv_acc[ind]
            v_old = v[ind].clone()
            loss_batch_old = (loss_s + lambda_coding * torch.sum(torch.abs(v[ind]))).item()
            v[ind] = softshrink(v[ind] - stepsize_v * g_rows, stepsize_v * lambda_coding)
            loss_s = smooth(x, target, v[ind], D2, False)
            v_cur = v[ind].clone()
            loss_batch_cur = (loss_s + lambda_coding * torch.sum(torch.abs(v[ind]))).item()
            loss_batch_cur_0 = loss_batch_cur
            delta_h = (torch.sum(torch.mul(g_rows, (v_cur - v_old))) + 1 / 2 / stepsize_v * torch.norm(v_cur - v_old) ** 2
                       + (torch.sum(torch.abs(v_cur)) - torch.sum(torch.abs(v[ind])))).item()
            i = 0
            while loss_batch_cur > loss_batch_old + delta_h * beta and i < 5:
                i += 1
                v[ind] = (delta ** i) * v_cur + (1 - delta ** i) * v_old
                loss_s = smooth(x, target, v[ind], D2, False)
                loss_batch_cur = (loss_s + torch.sum(torch.abs(v[ind]))).item()
                delta_h = delta_h * delta
            v[ind] = v_cur                         # (both branches of :444-448 keep the full step)
            if not loss_batch_cur_0 <= loss_batch_cur:
                i_max = max(i, i_max)
            # ---------- gradient of the D-step, new codes (:450-461) ----------
            _, dD2, dvb = smooth(x, target, v[ind], D2, True)
            D_has_grad = True
            gD_acc += dD2
            gv_acc[ind] += dvb
        stepsize_v = max(stepsize_v * (delta ** i_max), 1e-5)
        grad_D = gD_acc
        if torch.max(torch.abs(grad_D)).item() < 1e-4:
            if trace is not None:
                trace.append((i_max, -1))
            continue
        D_old = D2.clone()
        loss_i_old = loss_all(v, D_old)
        D2 = project_atoms((D2 - stepsize_D * grad_D).reshape(nc, nx, ny, n_atom), mode).reshape(P, n_atom)
        D_cur = D2.clone()
        loss_i_cur = loss_all(v, D_cur)
        loss_i_cur_0 = loss_i_cur
        delta_h_D = torch.sum(torch.mul(grad_D, (D_cur - D_old))) + 1 / 2 / stepsize_D * torch.norm(D_cur - D_old) ** 2
        i = 0
        while loss_i_cur > loss_i_old + delta_h_D * beta and i < 5:
            i += 1
            loss_i_cur = loss_all(v, (delta ** i) * D_cur + (1 - delta ** i) * D_old)
            delta_h_D = delta_h_D * delta
        D2 = D_cur                                 # (:486-492: the full step is kept either way)
        if loss_i_cur_0 <= loss_i_cur:
            loss.append(loss_i_cur_0)
        else:
            stepsize_D = max(stepsize_D * delta ** i, 1e-6)
            loss.append(loss_i_cur)
        gD_acc = torch.zeros(P, n_atom)            # a new D tensor: no gradient yet
        D_has_grad = False
        if trace is not None:
            trace.append((i_max, i))
        if abs(loss[-1] - loss[-2]) < 1e-6:
            break
    return D2.reshape(nc, nx, ny, n_atom), v, loss, (stepsize_v, stepsize_D)
