"""Device-side mirrors of the attack math helpers in the reference's attacks/utils.py that sit on the ADiL hot
path.  Same names and argument meaning; the arithmetic runs in the CUDA kernels (no CPU path)."""
import torch

from . import ops


def clamp_image(image, max_val=1, min_val=0):
    """utils.py:17-18."""
    return torch.clamp(image, max=max_val, min=min_val)


def project_onto_l1_ball(x, eps):
    """Euclidean projection of every row of x onto the l1 ball of radius eps (utils.py:21-41).  Returns a new
    tensor of the same shape; rows are x.view(x.shape[0], -1)."""
    out = x.detach().clone().contiguous()
    ops.project_rows(out.view(out.shape[0], -1), ops.ROWS_L1BALL, eps)
    return out


def constraint_dict(d, constr_set='l2ball'):
    """Per-atom projection of d[C,H,W,K] (utils.py:44-57): 'l2sphere', 'l2ball', anything else = the l1 ball of
    radius 1 applied to every channel row of every atom (utils.py:55-56).  In place, like the reference."""
    mode = {'l2sphere': ops.ATOMS_L2SPHERE, 'l2ball': ops.ATOMS_L2BALL}.get(constr_set, ops.ATOMS_L1BALL)
    if not d.is_contiguous():
        raise ValueError("constraint_dict needs a contiguous dictionary")
    ops.project_atoms(d, mode)
    return d


class _SoftShrink(object):
    def __init__(self, lambd):
        self.lambd = lambd

    def __call__(self, v):
        out = v.detach().clone().contiguous()
        ops.project_rows(out.view(out.shape[0], -1) if out.dim() > 1 else out.view(1, -1), ops.ROWS_SOFTSHRINK,
                         self.lambd)
        return out


def get_prox_l1(param):
    """Soft thresholding operator (utils.py:159-161)."""
    return _SoftShrink(param)


def get_target(img, label, targeted, classifier):
    """Second most probable class of the clean image (targeted) or the given label (utils.py:164-174)."""
    with torch.no_grad():
        if targeted:
            return classifier(img).sort().indices[:, -2]
        return label


class QuickAttackDataset(torch.utils.data.Dataset):
    """utils.py:177-186."""

    def __init__(self, images, labels):
        self.images = images
        self.labels = labels

    def __len__(self):
        return len(self.images)

    def __getitem__(self, item):
        return self.images[item], self.labels[item]
