"""TEST INFRASTRUCTURE -- loader for the UNMODIFIED reference ADIL (build container only).

Imports `/root/reference/attacks/attacks_classes/adil.py` without touching it, by
pre-seeding `sys.modules` with the third-party modules that are absent in this image
(`torchattacks`, `hostlist`) and with bare package objects for `attacks` /
`attacks.attacks_classes`, so that `attacks/__init__.py` (which pulls in `fast_uap.py`,
un-importable on torch >= 2.x, fast_uap.py:12) never executes.  The SLURM environment
variables `env_setting.py:10-16` reads at import time are faked.

`/root/reference` exists only in the build container, never on the GPU box: this module is
used by `oracle/make_golden.py` (fixture generation) and by CPU tests that are skipped when
the reference tree is missing.  Nothing in the product package imports it.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ADIL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "attacks", "attacks_classes", "adil.py"))


def _fake_torchattacks():
    import torch

    class Attack(object):
        # Minimal stand-in for torchattacks.attack.Attack (3.x API): adil.py:38,68,104,109 only use
        # .model, .device, ._targeted and __call__ -> forward.
        def __init__(self, name, model):
            self.attack = name
            self.model = model
            self.model_name = str(model).split("(")[0]
            self.device = next(model.parameters()).device
            self._targeted = False
            self._attack_mode = "default"
            self._return_type = "float"
            self._supported_mode = ["default"]

        def forward(self, *inputs):
            raise NotImplementedError

        def __call__(self, *inputs, **kwargs):
            self.model.eval()
            return self.forward(*inputs, **kwargs)

    pkg = types.ModuleType("torchattacks")
    sub = types.ModuleType("torchattacks.attack")
    sub.Attack = Attack
    pkg.attack = sub
    pkg.Attack = Attack
    return pkg, sub


def load_reference():
    """Return the reference module `attacks.attacks_classes.adil` (unmodified source)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "attacks.attacks_classes.adil" in sys.modules:
        return sys.modules["attacks.attacks_classes.adil"]
    if "torchattacks" not in sys.modules:
        try:
            import torchattacks  # noqa: F401  (prefer the real one when installed)
        except Exception:
            pkg, sub = _fake_torchattacks()
            sys.modules["torchattacks"] = pkg
            sys.modules["torchattacks.attack"] = sub
    if "hostlist" not in sys.modules:
        hl = types.ModuleType("hostlist")
        hl.expand_hostlist = lambda s: [s]
        sys.modules["hostlist"] = hl
    for key, val in (("SLURM_JOB_NODELIST", "127.0.0.1"), ("SLURM_STEP_GPUS", "0"), ("SLURM_NTASKS", "1"),
                     ("SLURM_JOB_NUM_NODES", "1"), ("SLURM_PROCID", "0"), ("SLURM_LOCALID", "0")):
        os.environ.setdefault(key, val)
    for name, sub in (("attacks", "attacks"), ("attacks.attacks_classes", "attacks/attacks_classes")):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = [os.path.join(REFERENCE_ROOT, sub)]
            sys.modules[name] = mod
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    return importlib.import_module("attacks.attacks_classes.adil")


def load_reference_utils():
    load_reference()
    import importlib
    return importlib.import_module("attacks.utils")
