"""TEST INFRASTRUCTURE -- generates tests/golden/adil_fb_reference_golden.npz: the reference's own full-batch ADiL with
backtracking line search (`adil()` of attacks/attacks_classes/adil_regularized.py:31-197, UNMODIFIED, imported through
oracle/ref_shim.py) on the tiny seeded problem of oracle/make_golden.py.

Run in the build container only (the reference tree does not travel to the GPU box):

    python oracle/make_golden_adil_fb.py

Stored per case: the initial dictionary the reference drew (so that the oracle and the kernels start from the same
point), its outputs (D, v, loss per iteration) and the inputs are those of make_golden.tiny_data().
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle.adil_oracle import tiny_classifier  # noqa: E402
from oracle.make_golden import QuickDataset, tiny_data, H, W  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "adil_fb_reference_golden.npz")

CASES = (("fb_untargeted", dict(targeted=False, niter=8, lambdaCoding=0.01, l2_fool=0.5, batchsize=4, step_size=0.05,
                                n_atom=6, dict_set='l2ball')),
         ("fb_targeted", dict(targeted=True, niter=8, lambdaCoding=0.02, l2_fool=2.0, batchsize=None, step_size=0.02,
                              n_atom=5, dict_set='l2sphere')),
         # a first step far too long: the line search backtracks (accepted indices 4, 5, then 0) before the
         # Lipschitz estimate takes over
         ("fb_backtrack", dict(targeted=False, niter=8, lambdaCoding=0.05, l2_fool=0.5, batchsize=4, step_size=10.0,
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
        D0 = ru.constraint_dict(torch.randn(3, H, W, kw["n_atom"]), constr_set=kw["dict_set"])   # the draw adil() makes
        torch.set_rng_state(state)
        d, v, loss_all = reg.adil(QuickDataset(xtr, ytr), model, device=torch.device("cpu"), **kw)
        out[tag + "_D0"] = D0.numpy()
        out[tag + "_D"] = d.detach().numpy()
        out[tag + "_v"] = v.detach().numpy()
        out[tag + "_loss"] = np.asarray(loss_all, dtype=np.float64)
        print(tag, "loss", np.asarray(loss_all))
    out["meta_torch_version"] = np.asarray(torch.__version__)
    np.savez_compressed(OUT, **out)
    print("wrote %s (%d arrays, %.1f KB)" % (OUT, len(out), os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
