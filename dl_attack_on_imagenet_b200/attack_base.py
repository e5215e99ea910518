"""Minimal stand-in for `torchattacks.attack.Attack` (the reference's base class, adil.py:38, utils.py:4).

Only the contract ADIL relies on is provided: `.model`, `.device`, `._targeted` and `__call__ -> forward` with
the model switched to eval mode.  When the real torchattacks package is importable it is used instead.
"""
try:  # pragma: no cover - torchattacks is not part of this image
    from torchattacks.attack import Attack  # noqa: F401
except Exception:

    class Attack(object):
        def __init__(self, name, model):
            self.attack = name
            self.model = model
            self.model_name = str(model).split("(")[0]
            self.device = next(model.parameters()).device
            self._targeted = False
            self._attack_mode = 'default'
            self._return_type = 'float'
            self._supported_mode = ['default']

        def forward(self, *inputs):
            raise NotImplementedError

        def set_mode_targeted_by_function(self, *_a, **_k):
            self._targeted = True

        def __call__(self, *inputs, **kwargs):
            self.model.eval()
            return self.forward(*inputs, **kwargs)
