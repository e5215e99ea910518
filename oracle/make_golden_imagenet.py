"""TEST INFRASTRUCTURE -- generates tests/golden/adil_reference_imagenet.npz by running the UNMODIFIED reference
(`/root/reference`, imported through oracle/ref_shim.py) at real image size (3x224x224) on random-init
torchvision classifiers, CPU, same seeds as SURVEY.md section 8(d):

    cfg1        BASELINE.json configs[0]: ResNet-18, 32 images, 10 atoms, l_inf 8/255, batch 32, 20 iterations
    fr_<model>  fooling-rate cases: 200 images (one image = 0.5 points), 10 atoms, batch 100, <steps> epochs on
                resnet18 / vgg11 / densenet121

    python oracle/make_golden_imagenet.py [case ...]        # build container only; ~1 h of CPU for all cases
    python oracle/make_golden_imagenet.py fr_vgg11+floor    # the reference's own sensitivity floor of a case: the same
                                                            # run with an ulp-level change of the Normalize layer

Stored per case (all small): per-epoch loss and training fooling rate (adil.py:194-195), checksums of the initial
state (so that a test can prove it regenerated the same D0 / v0), the final codes v, and the final dictionary /
perturbation on 256 fixed pixels; for cfg1 also per-step checksums of D and v.  The per-epoch validation coder
(adil.py:198-205, 100 classifier iterations per epoch) is stubbed out as SURVEY.md section 8(c) prescribes; the
validation DataLoader is still iterated, so the CPU-RNG draws of the run are the reference's.
"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle.adil_oracle import IndexedTensorDataset  # noqa: E402
from dl_attack_on_imagenet_b200.data import build_classifier, synthetic_images  # noqa: E402  (synthetic data helpers)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "adil_reference_imagenet.npz")
EPS = 8.0 / 255.0
P = 3 * 224 * 224
NPIX = 256

CASES = {
    # name: (model, n_img, n_atoms, batch, steps)
    "cfg1": ("resnet18", 32, 10, 32, 20),
    "fr_resnet18": ("resnet18", 200, 10, 100, 26),
    "fr_vgg11": ("vgg11", 200, 10, 100, 8),
    "fr_densenet121": ("densenet121", 200, 10, 100, 10),
}


def pixel_subset():
    return torch.randperm(P, generator=torch.Generator().manual_seed(5))[:NPIX]


def checksum(t):
    t = t.detach().double()
    return np.asarray([t.sum().item(), t.abs().sum().item()], dtype=np.float64)


class NormalizeUlp(torch.nn.Module):
    """Normalize with the same mean / std, evaluated as input * (1/std) - mean * (1/std) instead of (input - mean) / std:
    an ulp-level change of the classifier input.  Running the UNMODIFIED reference on a model that starts with this
    layer measures the reference's own sensitivity floor for a case."""

    def __init__(self, base):
        super().__init__()
        self.register_buffer('mean', base.mean.clone())
        self.register_buffer('std', base.std.clone())

    def forward(self, x):
        r = (1.0 / self.std).reshape(1, -1, 1, 1)
        return x * r - self.mean.reshape(1, -1, 1, 1) * r


def run_floor(ref, name, out):
    """`<case>_loss_ulp`, `<case>_fool_ulp`: the reference's trajectory when only the Normalize layer's rounding changes."""
    arch, n_img, K, B, steps = CASES[name]
    torch.set_num_threads(os.cpu_count() or 1)
    model = build_classifier(arch, seed=0)
    model = torch.nn.Sequential(NormalizeUlp(model[0]), model[1]).eval()
    x, y = synthetic_images(n_img, seed=1)
    xv, yv = synthetic_images(1, seed=2)
    tr, va = IndexedTensorDataset(x, y), IndexedTensorDataset(xv, yv)
    ref.ADIL.forward_supervised_AdamW = lambda self, images, labels, d, model='train': torch.zeros((), dtype=torch.long)
    t0 = time.time()
    torch.manual_seed(1234)
    atk = ref.ADIL(model, eps=EPS, steps=steps, norm='linf', n_atoms=K, batch_size=B, data_train=tr, data_val=va,
                   model_name=name + "_ulp", step_size=0.01, loss='ce', method='gd')
    D, v, loss_all, fool_all, _ = torch.load(atk.model_file, weights_only=False)
    out[name + "_loss_ulp"] = np.asarray(loss_all, dtype=np.float64)
    out[name + "_fool_ulp"] = np.asarray(fool_all, dtype=np.float64)
    pix = pixel_subset()
    out[name + "_Dv_sub_ulp"] = (v @ D.reshape(P, K)[pix].t()).numpy()
    print("%s (ulp-perturbed Normalize): %d epochs in %.0f s, fooling %s" % (
        name, len(loss_all), time.time() - t0, np.round(np.asarray(fool_all), 4).tolist()), flush=True)


def run_case(ref, name, out):
    if name.endswith("+floor"):
        return run_floor(ref, name[:-len("+floor")], out)
    arch, n_img, K, B, steps = CASES[name]
    torch.set_num_threads(os.cpu_count() or 1)
    model = build_classifier(arch, seed=0)                       # manual_seed(0) + torchvision init, Sequential(Normalize, net)
    x, y = synthetic_images(n_img, seed=1)
    xv, yv = synthetic_images(1, seed=2)
    tr, va = IndexedTensorDataset(x, y), IndexedTensorDataset(xv, yv)
    per_step = []
    if name == "cfg1":
        orig = ref.Attack_dict_model.update_d

        def update_d(self):                                       # last call of every minibatch step (adil.py:188)
            orig(self)
            per_step.append(np.concatenate([checksum(self.d.data), checksum(self.v.data)]))
        ref.Attack_dict_model.update_d = update_d
    init = {}
    orig_init = ref.Attack_dict_model.__init__

    def adm_init(self, d, v, eps):                                # record the initial state the reference drew
        init["D0"], init["v0"] = checksum(d), checksum(v)
        orig_init(self, d, v, eps)
    ref.Attack_dict_model.__init__ = adm_init
    ref.ADIL.forward_supervised_AdamW = lambda self, images, labels, d, model='train': torch.zeros((), dtype=torch.long)
    t0 = time.time()
    torch.manual_seed(1234)
    atk = ref.ADIL(model, eps=EPS, steps=steps, norm='linf', n_atoms=K, batch_size=B, data_train=tr, data_val=va,
                   model_name=name, step_size=0.01, loss='ce', method='gd')
    wall = time.time() - t0
    ref.Attack_dict_model.__init__ = orig_init
    if name == "cfg1":
        ref.Attack_dict_model.update_d = orig
    D, v, loss_all, fool_all, _ = torch.load(atk.model_file, weights_only=False)
    D2 = D.reshape(P, K)
    pix = pixel_subset()
    out[name + "_loss"] = np.asarray(loss_all, dtype=np.float64)
    out[name + "_fool"] = np.asarray(fool_all, dtype=np.float64)
    out[name + "_init"] = np.concatenate([init["D0"], init["v0"]])
    out[name + "_v"] = v.numpy()
    out[name + "_D_sub"] = D2[pix].numpy()
    out[name + "_Dv_sub"] = (v @ D2[pix].t()).numpy()
    out[name + "_final"] = np.concatenate([checksum(D), checksum(v)])
    out[name + "_meta"] = np.asarray([n_img, K, B, steps, wall], dtype=np.float64)
    if per_step:
        out[name + "_steps"] = np.stack(per_step)
    print("%s: %d epochs in %.0f s, loss %.6f -> %.6f, fooling %.4f -> %.4f" % (
        name, len(loss_all), wall, loss_all[0], loss_all[-1], fool_all[0], fool_all[-1]), flush=True)


def main():
    names = sys.argv[1:] or list(CASES)
    ref = ref_shim.load_reference()
    out = dict(np.load(OUT, allow_pickle=False)) if os.path.exists(OUT) else {}
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    os.makedirs("trained_dicts", exist_ok=True)
    try:
        for name in names:
            run_case(ref, name, out)
            out["meta_torch_version"] = np.asarray(torch.__version__)
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            np.savez_compressed(OUT, **out)                         # saved after every case: the run takes a while
    finally:
        os.chdir(cwd)
    print("wrote %s (%d arrays, %.1f KB)" % (OUT, len(out), os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
