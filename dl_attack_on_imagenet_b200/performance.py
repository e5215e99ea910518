"""Evaluation of crafted images: device-side mirror of the reference's performance.py for the ADiL path (SURVEY
section 8(f) row 3: what main.py / performance.py call per evaluated batch, and the transfer sweep of BASELINE
configs[4]).  Same names, argument meaning and return types as the reference; the per-image error reductions run in
one fused CUDA pass (adil_image_errors), the classifier passes stay in PyTorch / cuDNN.  No CPU fallback."""
import torch

from . import ops


def _reduce(values, reduction):
    if reduction == 'sum':
        return values.sum().item()
    if reduction == 'mean':
        return values.mean().item()
    return None  # (the reference returns None for any other reduction as well)


def compute_fooling_rate(model, adversary, clean, reduction='sum'):
    """Number ('sum') or fraction ('mean') of images whose predicted label changes (performance.py:238-246)."""
    net = model.eval()
    with torch.no_grad():
        changed = net(clean).argmax(dim=1).ne(net(adversary).argmax(dim=1))
    return _reduce(changed.float(), reduction)


def _errors(adversary, clean):
    if not adversary.is_cuda:
        raise RuntimeError("performance metrics run on the CUDA device of the attack; there is no CPU fallback")
    return ops.image_errors(adversary.float(), clean.float())


def compute_rmse(adversary, clean, reduction='sum'):
    """Relative squared error ||adv - clean||^2 / ||clean||^2 per image, summed or averaged (performance.py:249-257)."""
    err2, ref2, _ = _errors(adversary, clean)
    return _reduce(err2 / ref2, reduction)


def compute_mse(adversary, clean, reduction='sum'):
    """Squared error ||adv - clean||^2 per image, summed or averaged (performance.py:260-266)."""
    err2, _, _ = _errors(adversary, clean)
    return _reduce(err2, reduction)


def batch_metrics(model, adversary, clean):
    """fooling count, sum of relative errors and sum of squared errors of one batch as DEVICE scalars: the three
    reference metrics with one error pass and no host synchronisation (used by the loops below)."""
    with torch.no_grad():
        fooled = (model(clean).argmax(dim=1) != model(adversary).argmax(dim=1)).sum()
    upper, lower, _ = _errors(adversary, clean)
    return fooled, (upper / lower).sum(), upper.sum()


def performance(attack, model, data, device=None):
    """performance.py:154-177: attack every correctly classified image of `data`, report fooling rate, rmse, mse."""
    device = attack.device
    model = model.eval()
    num_samples = torch.zeros((), dtype=torch.long, device=device)
    acc = torch.zeros(3, dtype=torch.float64, device=device)
    for x, y in data:
        x, y = x.to(device=device), y.to(device=device)
        with torch.no_grad():
            ind = model(x).argmax(dim=-1) == y
        x, y = x[ind], y[ind]
        num_samples += ind.sum()
        if x.shape[0] == 0:
            continue
        adversary = attack(x, y)
        if isinstance(adversary, tuple):  # forward_unsupervised returns (adv_best, dv_norm_inf), adil.py:506
            adversary = adversary[0]
        f, r, m = batch_metrics(model, adversary, x)
        acc += torch.stack([f.double(), r.double(), m.double()])
    n = num_samples.item()
    vals = (acc / max(n, 1)).tolist()
    return {"fooling_rate": vals[0], "rmse": vals[1], "mse": vals[2]}


def empty_transfer_performance(model_transfer):
    """NaN entries for an attack family without a trained attack (performance.py:198-202)."""
    nan = float('nan')
    return {name: dict(fooling_rate=nan, rmse=nan, mse=nan) for name in model_transfer}


def get_transfer_performance_aux(attack, model_transfer, data, device=None, errors_fn=None):
    """performance.py:205-232: craft the adversarial images once per batch, evaluate them on every model.  Sharded
    evaluation: when torch.distributed is initialised each rank passes its own shard of `data` and the counters are
    summed (images are independent; no other exchange).  `errors_fn(adv, clean) -> (err2, ref2, linf)` defaults to the
    CUDA kernel; the multi-process host test injects a host implementation."""
    device = attack.device
    _errors = errors_fn if errors_fn is not None else globals()['_errors']
    names = list(model_transfer.keys())
    acc = torch.zeros(len(names), 3, dtype=torch.float64, device=device)
    count = torch.zeros((), dtype=torch.float64, device=device)
    for x, y in data:
        x, y = x.to(device=device), y.to(device=device)
        adversary = attack(x, y)
        if isinstance(adversary, tuple):
            adversary = adversary[0]
        upper, lower, _ = _errors(adversary, x)
        r, m = (upper / lower).sum().double(), upper.sum().double()
        count += x.shape[0]
        for i, name in enumerate(names):
            model = model_transfer[name].to(device=device).eval()
            with torch.no_grad():
                f = (model(x).argmax(dim=1) != model(adversary).argmax(dim=1)).sum().double()
            acc[i] += torch.stack([f, r, m])
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(acc)
            dist.all_reduce(count)
    except ImportError:
        pass
    vals = (acc / count.clamp(min=1)).tolist()
    return {name: {'fooling_rate': vals[i][0], 'rmse': vals[i][1], 'mse': vals[i][2]} for i, name in enumerate(names)}


def get_transfer_performance(atks, models, data, device=None, errors_fn=None):
    """Transfer performance of the first attack of every family on every model (performance.py:183-195)."""
    return {family: (get_transfer_performance_aux(members[0], models, data=data, device=device, errors_fn=errors_fn)
                     if len(members) > 0
                     else empty_transfer_performance(models))
            for family, members in atks.items()}
