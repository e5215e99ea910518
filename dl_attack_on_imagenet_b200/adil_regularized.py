"""Regularised ADiL variants on the B200 kernels: `sadil`, the stochastic forward-backward scheme of the reference's
attacks/attacks_classes/adil_regularized.py:200-312, `adil`, its full-batch scheme with backtracking line search
(:31-197), `sadil_updated` (:315-501) and `learn_coding_vectors`, the coder on a fixed dictionary (:508-628), for the
penalised objective  coeff * CE(x + D v) + 0.5 * l2_fool * ||D v||^2 + lambdaCoding * ||v||_1  with
D constrained per atom.  Same function names and arguments as the reference; the arithmetic of the path runs in the
CUDA kernels behind the C ABI (include/adil_b200.h) -- there is no CPU path:

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


def _resident_batches(dataset, batchsize, dev):
    """[(x [n,C,H,W], y [n], rows of v)] of the reference's DataLoader(dataset, batch_size, shuffle=False) with its
    get_slices index lists (utils.py:153-156), moved to the device once: the set stays resident in HBM."""
    batches, start = [], 0
    for x, y in torch.utils.data.DataLoader(dataset, batch_size=batchsize, shuffle=False):
        n = x.shape[0]
        batches.append((x.to(dev).float().contiguous(), y.to(dev), torch.arange(start, start + n, device=dev)))
        start += n
    return batches


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


def _smooth_loss(model, batches, D2, v, coeff, l2_fool, targeted):
    """sum over the batches of coeff * CE_sum + 0.5 * l2_fool * ||D v||^2 as an fp32 device scalar, accumulated in the
    reference's order (adil_regularized.py:112-117,166-174): the loss-only pass of the line search"""
    total = torch.zeros((), device=D2.device)
    with torch.no_grad():
        for x, y, rows in batches:
            n = x.shape[0]
            pert = torch.empty(n, D2.shape[0], device=D2.device)
            adv, _ = ops.synth(D2, v, rows, x=x.view(n, -1), delta_out=pert)
            target = get_target(x, y, targeted, model)
            total = total + coeff * torch.nn.functional.cross_entropy(model(adv.view_as(x)), target, reduction='sum') \
                + .5 * l2_fool * pert.square().sum()
    return total


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
    batches = _resident_batches(dataset, batchsize, dev)
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


def adil(dataset, model, targeted=True, niter=1e3, lambdaCoding=1., l2_fool=1., batchsize=None, step_size=.1, n_atom=10,
         dict_set='l2ball', device=None, dictionary=None, trace=None):
    """ADiL, full batch, with the backtracking of Bonettini et al. (adil_regularized.py:31-197).  Returns
    (D [C,H,W,K], v [N,K], loss_all [niter], NaN where an iteration never ran).

    Per iteration: penalised loss + gradients over the whole set (adil_synth with the perturbation output, the classifier's
    input gradient, the l2-penalised backward contractions accumulating dD over the batches), the Lipschitz estimate from
    the third iteration on (:126-130), v <- soft threshold, D <- constraint_dict(D - step grad) (adil_code_prox_step,
    adil_dict_step_atoms), and the line search over the segment old -> new with loss-only passes (`penalised_loss`).
    `dictionary` given: D is fixed and only the codes are learned, like the reference.  `trace` (a list) receives the
    accepted line-search index of every iteration (51: no decrease found, the iteration stops -- :189-192)."""
    import numpy as np
    dev = _device_of(model)
    model = model.eval()
    net, mean, std = split_normalize(model)
    flags = ops.SYNTH_NORMALIZE if mean is not None else 0
    nimg = len(dataset)
    x0, _ = next(iter(dataset))
    nc, nx, ny = x0.shape
    P = nc * nx * ny
    learn_D = dictionary is None
    if batchsize is None:
        batchsize = nimg
    delta_ls, gamma, beta = .5, 1., .5
    lipschitz = .9 / step_size
    coeff = 1. if targeted else -1.
    atoms_mode = _ATOMS.get(dict_set, ops.ATOMS_L1BALL)
    batches = _resident_batches(dataset, batchsize, dev)
    if learn_D:
        K = n_atom
        D = ops.project_atoms(torch.randn(3, nx, ny, K, device=dev), atoms_mode)   # adil_regularized.py:83-84
    else:
        D = dictionary.to(dev).float().contiguous().clone()
        K = D.shape[-1]
    D2 = D.view(P, K)
    v = torch.zeros(nimg, K, device=dev)
    all_rows = torch.arange(nimg, device=dev)

    def loss_and_grads():
        """loss_smooth (:109-117) with its gradients w.r.t. v (all rows) and D (summed over the batches)"""
        gD = torch.zeros(P, K, device=dev)
        gv = torch.zeros(nimg, K, device=dev)
        total = torch.zeros((), device=dev)
        for i, (x, y, rows) in enumerate(batches):
            n = x.shape[0]
            target = get_target(x, y, targeted, model)
            pert = torch.empty(n, P, device=dev)
            xin, _ = ops.synth(D2, v, rows, x=x.view(n, -1), mean=mean, std=std, flags=flags, delta_out=pert, n_channels=nc)
            xin = xin.view_as(x).requires_grad_(True)
            ce = coeff * torch.nn.functional.cross_entropy(net(xin), target, reduction='sum')
            (g,) = torch.autograd.grad(ce, xin)
            _, dvb = ops.grad(g.contiguous().view(n, P), D2, v, rows, std, want_dD=learn_D, dD2=gD if learn_D else None,
                              accumulate=learn_D and i > 0, delta=pert, l2_coef=l2_fool)
            gv[rows] = dvb
            total = total + ce.detach() + .5 * l2_fool * pert.square().sum()
        return total, gD, gv

    D_old, v_old = torch.zeros_like(D2), torch.zeros_like(v)
    gD_old, gv_old = torch.zeros_like(D2), torch.zeros_like(v)
    loss_all = np.nan * np.ones(int(niter))
    loss_ns_old = torch.zeros((), device=dev)
    stop = False
    for it in range(int(niter)):
        if stop:
            continue
        loss_ns = lambdaCoding * v.abs().sum()
        loss_s, gD, gv = loss_and_grads()
        loss_full = loss_s + loss_ns
        if it > 1:                                                                          # :126-130
            num = torch.sqrt(torch.norm(gv - gv_old) ** 2 + torch.norm(gD - gD_old) ** 2)
            lipschitz = num / torch.sqrt(torch.norm(v - v_old) ** 2 + torch.norm(D2 - D_old) ** 2)
        D_old.copy_(D2); v_old.copy_(v); gv_old.copy_(gv); gD_old.copy_(gD)
        loss_old = loss_full
        step = float(.9 / lipschitz)                                                        # :140
        ops.code_prox_step(v, gv, all_rows, step, ops.ROWS_SOFTSHRINK, step * lambdaCoding)  # :143
        if learn_D:                                                                         # :144-147
            if atoms_mode == ops.ATOMS_L1BALL:
                ops.dict_step_atoms(D2, gD, ops.ATOMS_NONE, step=step)
                ops.project_atoms(D, ops.ATOMS_L1BALL)
            else:
                ops.dict_step_atoms(D2, gD, atoms_mode, step=step)
        d_v, d_d = v - v_old, D2 - D_old
        h = (d_d * gD).sum() + (d_v * gv).sum() + .5 * (gamma / step) * (torch.norm(d_d) ** 2 + torch.norm(d_v) ** 2) \
            + loss_ns - loss_ns_old                                                          # :154-156
        i = 0
        new_v, new_D2 = torch.empty_like(v), torch.empty_like(D2)
        while True:                                                                          # :158-192
            torch.add(v_old, d_v, alpha=delta_ls ** i, out=new_v)
            torch.add(D_old, d_d, alpha=delta_ls ** i, out=new_D2)
            loss_ns = lambdaCoding * new_v.abs().sum()
            loss_full = _smooth_loss(model, batches, new_D2, new_v, coeff, l2_fool, targeted) + loss_ns
            if bool(loss_full <= loss_old + beta * (delta_ls ** i) * h):
                v.copy_(new_v)
                D2.copy_(new_D2)
                loss_ns_old = loss_ns
                break
            i += 1
            if i > 50:
                stop = True
                break
        if trace is not None:
            trace.append(i)
        loss_all[it] = float(loss_full)
    return D, v, loss_all


def learn_coding_vectors(dataset, model, targeted=True, niter=1e2, lambda_l1=1., lambda_l2=1., batch_size=None,
                         step_size=.1, n_atom=10, dict_set='l2ball', device=None, dictionary=None, verbose=False, trace=None):
    """The coder of the regularised variant (adil_regularized.py:508-628): codes of `dataset` on the FIXED `dictionary`
    [C,H,W,K] for  coeff * CE(x + D v) + 0.5 * lambda_l2 * ||D v||^2 + lambda_l1 * ||v||_1 , full-batch proximal gradient
    from v = 0 with the reference's line search over the segment v_old -> prox step (factor 0.9, at most 10 shortenings;
    a shortened point is kept -- and the step size shrunk with it -- only when its loss is below the full step's).
    Returns v [N,K] like the reference; `trace` (a list) receives (index the search ended at, shortened point kept,
    recorded loss) per iteration.

    On the kernels: adil_synth with the perturbation output, the classifier's input gradient, the l2-penalised code-gradient
    contraction (adil_grad, dv only), adil_code_prox_step (soft threshold) and loss-only passes for the search."""
    import numpy as np
    dev = _device_of(model)
    model = model.eval()
    net, mean, std = split_normalize(model)
    flags = ops.SYNTH_NORMALIZE if mean is not None else 0
    nimg = len(dataset)
    D = dictionary.to(dev).float().contiguous()
    nc, nx, ny, K = D.shape
    P = nc * nx * ny
    D2 = D.view(P, K)
    delta_ls, gamma, beta = .9, 1, .5
    batch_size = nimg if batch_size is None else batch_size
    coeff = 1. if targeted else -1.
    batches = _resident_batches(dataset, batch_size, dev)
    targets = [get_target(x, y, targeted, model) for x, y, _ in batches]   # (clean images, fixed classifier: loop invariant)
    # the step size lives as an fp32 scalar and is updated with the reference's expressions (its default is a tensor, :509)
    step_t = torch.as_tensor(step_size, dtype=torch.float32).cpu()
    v = torch.zeros(nimg, K, device=dev)
    v_old = torch.zeros_like(v)
    all_rows = torch.arange(nimg, device=dev)

    def loss_and_grad():
        """loss_smooth (:553-559) and its gradient w.r.t. every code row (:565)"""
        gv = torch.zeros(nimg, K, device=dev)
        total = torch.zeros((), device=dev)
        for (x, _, rows), target in zip(batches, targets):
            n = x.shape[0]
            pert = torch.empty(n, P, device=dev)
            xin, _ = ops.synth(D2, v, rows, x=x.view(n, -1), mean=mean, std=std, flags=flags, delta_out=pert, n_channels=nc)
            xin = xin.view_as(x).requires_grad_(True)
            ce = coeff * torch.nn.functional.cross_entropy(net(xin), target, reduction='sum')
            (g,) = torch.autograd.grad(ce, xin)
            _, dvb = ops.grad(g.contiguous().view(n, P), D2, v, rows, std, want_dD=False, delta=pert, l2_coef=lambda_l2)
            gv[rows] = dvb
            total = total + ce.detach() + .5 * lambda_l2 * pert.square().sum()
        return total, gv

    def loss_only(v_):
        total = torch.zeros((), device=dev)
        with torch.no_grad():
            for (x, _, rows), target in zip(batches, targets):
                n = x.shape[0]
                pert = torch.empty(n, P, device=dev)
                adv, _ = ops.synth(D2, v_, rows, x=x.view(n, -1), delta_out=pert)
                total = total + coeff * torch.nn.functional.cross_entropy(model(adv.view_as(x)), target, reduction='sum') \
                    + .5 * lambda_l2 * pert.square().sum()
        return total

    loss_all = [np.nan]
    for _ in range(int(niter)):
        loss_s, gv = loss_and_grad()
        loss_old = (loss_s + lambda_l1 * v.abs().sum()).item()
        v_old.copy_(v)
        step = float(step_t)
        ops.code_prox_step(v, gv, all_rows, step, ops.ROWS_SOFTSHRINK, float(step_t * lambda_l1))          # :574-577
        d_v = v - v_old
        h = (d_v * gv).sum() + .5 * float(gamma / step_t) * (torch.norm(d_v) ** 2) \
            + lambda_l1 * v.abs().sum() - lambda_l1 * v_old.abs().sum()                                    # :583-584
        index_i, kept = 0, False
        new_v = torch.empty_like(v)
        while True:                                                                                        # :589-622
            torch.add(v_old, d_v, alpha=delta_ls ** index_i, out=new_v)
            loss_full = (loss_only(new_v) + lambda_l1 * new_v.abs().sum()).item()
            if index_i == 0:
                loss_cur = loss_full
            if bool(loss_full <= loss_old + beta * (delta_ls ** index_i) * h):
                if loss_cur > loss_full:
                    v.copy_(new_v)
                    step_t = step_t * delta_ls ** index_i
                    loss_all.append(loss_full)
                    kept = True
                else:
                    loss_all.append(loss_cur)
                break
            index_i += 1
            if index_i > 10:                                    # no sufficient decrease found: the last point tried is taken
                v.copy_(new_v)
                loss_all.append(loss_full)
                kept = True
                break
        if trace is not None:
            trace.append((index_i, kept, loss_all[-1]))
        if verbose:
            print("learn_coding_vectors: loss %.6f step %.4g (search ended at %d)" % (loss_all[-1], float(step_t), index_i))
        if loss_all[-2] - loss_all[-1] < 1e-6:
            break
    return v


def sadil_updated(dataset, model, targeted=True, nepochs=1e3, batchsize=1, lambdaCoding=1., l2_fool=1., stepsize=1.,
                  n_atom=5, dict_set='l2ball', device=None, model_file=None, dictionary=None, trace=None):
    """SADiL, "updated" (adil_regularized.py:315-501): per epoch a proximal step on the code rows of every minibatch and ONE
    projected gradient step on D with the gradient accumulated over the epoch; both steps are followed by a backtracking
    test (factor 0.5, at most 5 halvings) that adapts the step sizes while the full steps are kept (:444-448,486-492).
    Returns (D [C,H,W,K], v [N,K]) like the reference and saves [D, label, pred, v, loss] to `model_file` (:499);
    `dictionary` optionally gives the initial D (the reference draws randn and projects it); `trace` (a list) receives
    per epoch (largest number of halvings of the V-steps, halvings of the D-step or -1 when it was skipped, loss).

    The reference's autograd bookkeeping is kept (restated explicitly in the test suite's CPU checker, which reproduces the
    reference bit for bit): the
    gradient of v accumulates over every backward of the run, the gradient of D over every backward since the first
    D-step backward of the current D.  On the kernels: adil_synth (perturbation output), the classifier's input
    gradient, the l2-penalised contractions with dD ACCUMULATED in place (adil_grad, ADIL_GRAD_ACCUMULATE_DD),
    adil_code_prox_step, adil_dict_step_atoms and loss-only passes."""
    dev = _device_of(model)
    model = model.eval()
    net, mean, std = split_normalize(model)
    flags = ops.SYNTH_NORMALIZE if mean is not None else 0
    nimg = len(dataset)
    x0, _ = next(iter(dataset))
    nc, nx, ny = x0.shape
    P = nc * nx * ny
    K = n_atom
    delta_ls, beta = .5, .5
    atoms_mode = _ATOMS.get(dict_set, ops.ATOMS_L1BALL)
    coeff = 1. if targeted else -1.
    batches = _resident_batches(dataset, batchsize, dev)
    targets = [get_target(x, y, targeted, model) for x, y, _ in batches]   # (clean images, fixed classifier: loop invariant)
    if dictionary is None:
        D = ops.project_atoms(torch.randn(3, nx, ny, K, device=dev), atoms_mode)   # :357-358
    else:
        D = dictionary.to(dev).float().contiguous().clone()
        K = D.shape[-1]
    D2 = D.view(P, K)
    v = torch.zeros(nimg, K, device=dev)
    stepsize_D = stepsize_v = stepsize

    def smooth_and_grads(x, target, rows, want_dD):
        """loss_smooth (:395-396) as an fp32 device scalar; its code gradient; dD accumulated into gD_acc"""
        n = x.shape[0]
        pert = torch.empty(n, P, device=dev)
        xin, _ = ops.synth(D2, v, rows, x=x.view(n, -1), mean=mean, std=std, flags=flags, delta_out=pert, n_channels=nc)
        xin = xin.view_as(x).requires_grad_(True)
        ce = coeff * torch.nn.functional.cross_entropy(net(xin), target, reduction='sum')
        (g,) = torch.autograd.grad(ce, xin)
        _, dvb = ops.grad(g.contiguous().view(n, P), D2, v, rows, std, want_dD=want_dD, dD2=gD_acc if want_dD else None,
                          accumulate=want_dD, delta=pert, l2_coef=l2_fool)
        return ce.detach() + .5 * l2_fool * pert.square().sum(), dvb

    def smooth_only(x, target, rows, D2_):
        n = x.shape[0]
        with torch.no_grad():
            pert = torch.empty(n, P, device=dev)
            adv, _ = ops.synth(D2_, v, rows, x=x.view(n, -1), delta_out=pert)
            return coeff * torch.nn.functional.cross_entropy(model(adv.view_as(x)), target, reduction='sum') \
                + .5 * l2_fool * pert.square().sum()

    def loss_all(D2_):                                                                       # :362-373
        total = 0
        for (x, _, rows), target in zip(batches, targets):
            total += smooth_only(x, target, rows, D2_).item()
        return total + (lambdaCoding * v.abs().sum()).item()

    loss = [loss_all(D2)]
    label, pred = [], []
    gv_acc = torch.zeros(nimg, K, device=dev)       # v.grad of the reference: never zeroed
    gD_acc = torch.zeros(P, K, device=dev)          # D.grad of the current D tensor
    D_has_grad = False
    D_old, D_try = torch.empty_like(D2), torch.empty_like(D2)
    for epoch in range(int(nepochs)):
        i_max = 0
        for (x, y, rows), target in zip(batches, targets):
            if epoch == 0:                                                                   # :387-389
                label += y.tolist()
                with torch.no_grad():
                    pred += model(x).sort().indices[:, -1].tolist()
            # ---------- V-step (:391-448) ----------
            loss_s, dvb = smooth_and_grads(x, target, rows, D_has_grad)
            gv_acc[rows] += dvb
            g_rows = gv_acc[rows].contiguous()
            v_old = v[rows].clone()
            loss_batch_old = (loss_s + lambdaCoding * v_old.abs().sum()).item()
            ops.code_prox_step(v, g_rows, rows, stepsize_v, ops.ROWS_SOFTSHRINK, stepsize_v * lambdaCoding)   # :412-416
            v_cur = v[rows].clone()
            loss_batch_cur = (smooth_only(x, target, rows, D2) + lambdaCoding * v_cur.abs().sum()).item()
            loss_batch_cur_0 = loss_batch_cur
            d = v_cur - v_old
            delta_h = ((g_rows * d).sum() + 1 / 2 / stepsize_v * torch.norm(d) ** 2).item()   # (:420-421: the l1 terms cancel)
            i = 0
            while loss_batch_cur > loss_batch_old + delta_h * beta and i < 5:
                i += 1
                v[rows] = (delta_ls ** i) * v_cur + (1 - delta_ls ** i) * v_old
                loss_batch_cur = (smooth_only(x, target, rows, D2) + v[rows].abs().sum()).item()   # (:433: no lambda)
                delta_h = delta_h * delta_ls
            v[rows] = v_cur                                                                  # :444-448 keep the full step
            if not loss_batch_cur_0 <= loss_batch_cur:
                i_max = max(i, i_max)
            # ---------- gradient of the D-step at the new codes (:450-461) ----------
            _, dvb = smooth_and_grads(x, target, rows, True)
            D_has_grad = True
            gv_acc[rows] += dvb
        stepsize_v = max(stepsize_v * (delta_ls ** i_max), 1e-5)                              # :463
        if gD_acc.abs().max().item() < 1e-4:                                                 # :466-467
            if trace is not None:
                trace.append((i_max, -1, loss[-1]))
            continue
        D_old.copy_(D2)
        loss_i_old = loss_all(D_old)
        if atoms_mode == ops.ATOMS_L1BALL:                                                   # :472-474
            ops.dict_step_atoms(D2, gD_acc, ops.ATOMS_NONE, step=stepsize_D)
            ops.project_atoms(D, ops.ATOMS_L1BALL)
        else:
            ops.dict_step_atoms(D2, gD_acc, atoms_mode, step=stepsize_D)
        loss_i_cur = loss_all(D2)
        loss_i_cur_0 = loss_i_cur
        dd = D2 - D_old
        delta_h_D = ((gD_acc * dd).sum() + 1 / 2 / stepsize_D * torch.norm(dd) ** 2).item()
        i = 0
        while loss_i_cur > loss_i_old + delta_h_D * beta and i < 5:                          # :481-485
            i += 1
            torch.add(D_old, dd, alpha=delta_ls ** i, out=D_try)
            loss_i_cur = loss_all(D_try)
            delta_h_D = delta_h_D * delta_ls
        if loss_i_cur_0 <= loss_i_cur:                                                       # :486-492: D stays the full step
            loss.append(loss_i_cur_0)
        else:
            stepsize_D = max(stepsize_D * delta_ls ** i, 1e-6)
            loss.append(loss_i_cur)
        gD_acc.zero_()                                 # the reference's new D tensor has no gradient yet
        D_has_grad = False
        if trace is not None:
            trace.append((i_max, i, loss[-1]))
        if abs(loss[-1] - loss[-2]) < 1e-6:
            break
    if model_file is not None:
        torch.save([D, label, pred, v, loss], model_file)                                    # :499
    return D, v
