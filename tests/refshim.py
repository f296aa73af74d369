"""Import the reference's model.py unmodified (only possible where /root/reference exists, i.e. in
the build container).  `timm` is not installed; model.py:4 needs three symbols from
timm.models.layers, provided here as an in-memory shim with timm's published semantics:
DropPath (identity in eval; per-sample Bernoulli(keep)/keep in training), to_2tuple,
trunc_normal_ (== torch.nn.init.trunc_normal_, truncation at absolute +-2)."""
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_DIR = os.environ.get("UFORMER_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "model.py"))


def _install_timm_shim():
    if "timm" in sys.modules:
        return

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1.0 - self.drop_prob
            shape = (x.shape[0],) + (1,) * (x.ndim - 1)
            return x * x.new_empty(shape).bernoulli_(keep).div_(keep)

    def to_2tuple(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.DropPath = DropPath
    layers.to_2tuple = to_2tuple
    layers.trunc_normal_ = nn.init.trunc_normal_
    timm.models = models
    models.layers = layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def import_reference_model():
    """Returns the reference `model` module (fresh import on first call)."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_DIR)
    _install_timm_shim()
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import importlib
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("model")
