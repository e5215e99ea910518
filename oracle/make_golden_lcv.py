"""TEST INFRASTRUCTURE -- generates tests/golden/lcv_reference_golden.npz: the reference's own coder on a fixed
dictionary (`learn_coding_vectors()` of attacks/attacks_classes/adil_regularized.py:508-628, UNMODIFIED, imported
through oracle/ref_shim.py) on the tiny seeded problem of oracle/make_golden.py.

Run in the build container only (the reference tree does not travel to the GPU box):

    python oracle/make_golden_lcv.py

The function returns the codes only; its per-iteration losses and final step size are read from the frame's locals
when it returns (a `sys.setprofile` observer -- the source is not touched).  Stored per case: the dictionary, the codes,
the recorded losses (the first entry is the reference's NaN) and the final step size.
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

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "lcv_reference_golden.npz")

CASES = (("lcv_untargeted", dict(targeted=False, niter=8, lambda_l1=0.01, lambda_l2=0.5, batch_size=4, step_size=0.05,
                                 n_atom=6)),
         ("lcv_targeted", dict(targeted=True, niter=8, lambda_l1=0.02, lambda_l2=2.0, batch_size=None, step_size=0.02,
                               n_atom=5)),
         # a step far too long: no index up to 10 qualifies, the last point tried is taken (:615-620)
         ("lcv_backtrack", dict(targeted=False, niter=8, lambda_l1=0.05, lambda_l2=0.5, batch_size=4, step_size=10.0,
                                n_atom=6)),
         # a step a few times too long: the search ends at indices 5, 0, 3, 2, 5, 0, 2, 0 and the step size follows
         ("lcv_linesearch", dict(targeted=False, niter=8, lambda_l1=0.05, lambda_l2=0.5, batch_size=4, step_size=1.0,
                                 n_atom=6)))


def run_observed(fn, *args, **kw):
    """fn(*args, **kw) and the locals of its frame at return"""
    seen = {}

    def prof(frame, event, arg):
        if event == "return" and frame.f_code is fn.__code__:
            seen.update(frame.f_locals)

    sys.setprofile(prof)
    try:
        out = fn(*args, **kw)
    finally:
        sys.setprofile(None)
    return out, seen


def main():
    torch.set_num_threads(1)
    ref_shim.load_reference()
    ru = ref_shim.load_reference_utils()
    reg = importlib.import_module("attacks.attacks_classes.adil_regularized")
    xtr, ytr, _, _ = tiny_data()
    model = tiny_classifier()
    out = {}
    for tag, kw in CASES:
        torch.manual_seed(2468)
        D = ru.constraint_dict(torch.randn(3, H, W, kw["n_atom"]), constr_set='l2ball')
        kw = dict(kw, step_size=torch.tensor(kw["step_size"]))          # (a tensor, like the reference's default)
        v, loc = run_observed(reg.learn_coding_vectors, QuickDataset(xtr, ytr), model, device=torch.device("cpu"),
                              dictionary=D, **kw)
        out[tag + "_D"] = D.numpy()
        out[tag + "_v"] = v.detach().numpy()
        out[tag + "_loss"] = np.asarray(loc["loss_all"], dtype=np.float64)
        out[tag + "_step"] = np.asarray(float(loc["step_size"]), dtype=np.float64)
        print(tag, "loss", np.asarray(loc["loss_all"]), "final step", float(loc["step_size"]), "nnz", int((v != 0).sum()))
    out["meta_torch_version"] = np.asarray(torch.__version__)
    np.savez_compressed(OUT, **out)
    print("wrote %s (%d arrays, %.1f KB)" % (OUT, len(out), os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
