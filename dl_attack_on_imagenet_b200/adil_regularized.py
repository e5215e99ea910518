"""Regularised ADiL variant on the B200 kernels: SADiL, the stochastic forward-backward scheme of the reference's
attacks/attacks_classes/adil_regularized.py:200-312 (the penalised objective  coeff * CE(x + D v) + 0.5 * l2_fool *
||D v||^2 + lambdaCoding * ||v||_1  with D constrained per atom).  Same function name and arguments as the reference;
the arithmetic runs in the CUDA kernels behind the C ABI (include/adil_b200.h) -- there is no CPU path:

    synthesis + perturbation output   adil_synth (delta_out)              adil_regularized.py:272,290
    penalised backward contractions   adil_grad (delta, l2_coef)          loss_smooth.backward(), :277,295
    D <- constraint_dict(D - step g)  adil_dict_step_atoms (hp = NULL)    :283-285
    v <- prox_l1(v - step g)          adil_code_prox_step (SOFTSHRINK)    :304
    loss-only pass                    adil_synth + classifier forward     loss_all, :245-254 (and every line search)
"""
import torch

from . import ops
from .adil import split_normalize
from .utils import get_target  # noqa: F401  (re-exported like the reference's `from attacks.utils import *`)

_ATOMS = {'l2ball': ops.ATOMS_L2BALL, 'l2sphere': ops.ATOMS_L2SPHERE}


def _device_of(model):
    dev = next(model.parameters()).device
    if dev.type != 'cuda':
        raise RuntimeError("sadil (B200) needs the classifier on a CUDA device; there is no CPU fallback")
    return dev


def penalised_loss(model, batches, D2, v, coeff, l2_fool, lambdaCoding, targeted):
    """loss_all of adil_regularized.py:245-254: one loss-only pass (synthesis + classifier forward, no backward) over
    `batches` = [(x [n,C,H,W] on the device, y, rows of v)] -- also the evaluation every line search repeats."""
    total = torch.zeros((), device=D2.device, dtype=torch.float64)
    with torch.no_grad():
        for x, y, rows in batches:
            n = x.shape[0]
            delta = torch.empty(n, D2.shape[0], device=D2.device)
            adv, _ = ops.synth(D2, v, rows, x=x.view(n, -1), delta_out=delta)
            target = get_target(x, y, targeted, model)
            ce = torch.nn.functional.cross_entropy(model(adv.view_as(x)), target, reduction='sum')
            total += (coeff * ce + 0.5 * l2_fool * delta.square().sum()).double()
    return total.item() + (lambdaCoding * v.abs().sum()).item()


def sadil(dataset, model, targeted=True, nepochs=1e3, batchsize=1, lambdaCoding=1., l2_fool=1., stepsize=1., n_atom=5,
          dict_set='l2ball', device=None, model_file=None, dictionary=None):
    """SADiL (adil_regularized.py:200-312).  Returns (D [C,H,W,K], v [N,K], loss list); `dictionary` optionally gives
    the initial D (the reference draws randn and projects it)."""
    dev = _device_of(model)
    model = model.eval()
    net, mean, std = split_normalize(model)
    flags = ops.SYNTH_NORMALIZE if mean is not None else 0
    nimg = len(dataset)
    x0, _ = next(iter(dataset))
    nc, nx, ny = x0.shape
    P = nc * nx * ny
    K = n_atom
    atoms_mode = _ATOMS.get(dict_set, ops.ATOMS_L1BALL)
    coeff = 1. if targeted else -1.
    loader = torch.utils.data.DataLoader(dataset, batch_size=batchsize, shuffle=False)
    batches = []
    start = 0
    for x, y in loader:                                         # the set stays resident in HBM (shuffle=False: fixed slices)
        n = x.shape[0]
        batches.append((x.to(dev).float().contiguous(), y.to(dev), torch.arange(start, start + n, device=dev)))
        start += n
    if dictionary is None:
        D = ops.project_atoms(torch.randn(3, nx, ny, K, device=dev), atoms_mode)   # adil_regularized.py:241-242
    else:
        D = dictionary.to(dev).float().contiguous().clone()
    D2 = D.view(P, K)
    v = torch.zeros(nimg, K, device=dev)
    dD2 = torch.empty(P, K, device=dev)
    loss = [penalised_loss(model, batches, D2, v, coeff, l2_fool, lambdaCoding, targeted)]

    def input_grad(x, target, rows):
        """Synthesis (classifier input + perturbation), d [coeff * CE_sum] / d input through the frozen classifier."""
        n = x.shape[0]
        delta = torch.empty(n, P, device=dev)
        xin, _ = ops.synth(D2, v, rows, x=x.view(n, -1), mean=mean, std=std, flags=flags, delta_out=delta, n_channels=nc)
        xin = xin.view_as(x).requires_grad_(True)
        ce = coeff * torch.nn.functional.cross_entropy(net(xin), target, reduction='sum')
        (g,) = torch.autograd.grad(ce, xin)
        return g.contiguous().view(n, P), delta

    # Reference behaviour kept (adil_regularized.py:287-304): `v` becomes a leaf that requires grad at the first V-step
    # and its `.grad` is never zeroed, so every later backward -- the D-step's too, v takes part in D v -- accumulates
    # into it, and the V-step uses the accumulated rows grad_v[ind].
    gv_acc = torch.zeros(nimg, K, device=dev)
    v_has_grad = False
    for _ in range(int(nepochs)):
        for x, y, rows in batches:
            target = get_target(x, y, targeted, model)
            # ---------- D-step (adil_regularized.py:264-285) ----------
            g, delta = input_grad(x, target, rows)
            _, dvb = ops.grad(g, D2, v, rows, std, want_dv=v_has_grad, dD2=dD2, delta=delta, l2_coef=l2_fool)
            if v_has_grad:
                gv_acc[rows] += dvb
            if atoms_mode == ops.ATOMS_L1BALL:
                ops.dict_step_atoms(D2, dD2, ops.ATOMS_NONE, step=stepsize)
                ops.project_atoms(D, ops.ATOMS_L1BALL)
            else:
                ops.dict_step_atoms(D2, dD2, atoms_mode, step=stepsize)
            # ---------- V-step with the new dictionary (adil_regularized.py:287-304) ----------
            g, delta = input_grad(x, target, rows)
            _, dvb = ops.grad(g, D2, v, rows, std, want_dD=False, delta=delta, l2_coef=l2_fool)
            v_has_grad = True
            gv_acc[rows] += dvb
            ops.code_prox_step(v, gv_acc[rows].contiguous(), rows, stepsize, ops.ROWS_SOFTSHRINK, stepsize * lambdaCoding)
        loss.append(penalised_loss(model, batches, D2, v, coeff, l2_fool, lambdaCoding, targeted))
        if abs(loss[-1] - loss[-2]) < 1e-6:
            break
    if model_file is not None:
        torch.save([D, loss], model_file)                        # adil_regularized.py:310
    return D, v, loss
