"""TEST INFRASTRUCTURE -- generates tests/golden/sadil_updated_reference_golden.npz: the reference's own `sadil_updated()`
(attacks/attacks_classes/adil_regularized.py:315-501, UNMODIFIED, imported through oracle/ref_shim.py) on the tiny
seeded problem of oracle/make_golden.py.

Run in the build container only (the reference tree does not travel to the GPU box):

    python oracle/make_golden_sadil_updated.py

The function returns (D, v) and saves [D, label, pred, v, loss]; its final step sizes are read from the frame's locals
when it returns (a `sys.setprofile` observer -- the source is not touched).  Stored per case: the initial dictionary the
reference drew, D, v, the loss list, the labels / predictions it records and the final step sizes.
"""
import importlib
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle.adil_oracle import tiny_classifier  # noqa: E402
from oracle.make_golden import QuickDataset, tiny_data, H, W  # noqa: E402
from oracle.make_golden_lcv import run_observed  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "sadil_updated_reference_golden.npz")

CASES = (("su_untargeted", dict(targeted=False, nepochs=5, batchsize=4, lambdaCoding=0.01, l2_fool=0.5, stepsize=0.05,
                                n_atom=6, dict_set='l2ball')),
         ("su_targeted", dict(targeted=True, nepochs=5, batchsize=3, lambdaCoding=0.02, l2_fool=2.0, stepsize=0.02,
                              n_atom=5, dict_set='l2sphere')),
         # steps far too long: both backtracking tests fire and the step sizes shrink
         ("su_backtrack", dict(targeted=False, nepochs=5, batchsize=4, lambdaCoding=0.05, l2_fool=0.5, stepsize=2.0,
                               n_atom=6, dict_set='l2ball')))


def main():
    torch.set_num_threads(1)
    ref_shim.load_reference()
    ru = ref_shim.load_reference_utils()
    reg = importlib.import_module("attacks.attacks_classes.adil_regularized")
    xtr, ytr, _, _ = tiny_data()
    model = tiny_classifier()
    out = {}
    for tag, kw in CASES:
        torch.manual_seed(4321)
        state = torch.get_rng_state()
        D0 = ru.constraint_dict(torch.randn(3, H, W, kw["n_atom"]), constr_set=kw["dict_set"])   # the draw sadil_updated() makes
        torch.set_rng_state(state)
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "su.bin")
            (d, v), loc = run_observed(reg.sadil_updated, QuickDataset(xtr, ytr), model, device="cpu", model_file=path, **kw)
            saved = torch.load(path, weights_only=False)
        out[tag + "_D0"] = D0.numpy()
        out[tag + "_D"] = d.detach().numpy()
        out[tag + "_v"] = v.detach().numpy()
        out[tag + "_loss"] = np.asarray(saved[4], dtype=np.float64)
        out[tag + "_label"] = np.asarray(saved[1], dtype=np.int64)
        out[tag + "_pred"] = np.asarray(saved[2], dtype=np.int64)
        out[tag + "_steps"] = np.asarray([float(loc["stepsize_v"]), float(loc["stepsize_D"])], dtype=np.float64)
        print(tag, "loss", np.asarray(saved[4]), "steps", out[tag + "_steps"])
    out["meta_torch_version"] = np.asarray(torch.__version__)
    np.savez_compressed(OUT, **out)
    print("wrote %s (%d arrays, %.1f KB)" % (OUT, len(out), os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
