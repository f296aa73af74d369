"""uformer_b200 — a Blackwell-native (sm_100a) LeWin-block engine that drops into Uformer.

    import model                      # the reference's model.py
    import uformer_b200
    uformer_b200.install(model)       # model.Uformer(...) now builds on the native engine

`install` rebinds the three classes the reference resolves by global name at construction time
(LeWinTransformerBlock, WindowAttention, LeFF — model.py:1027-1028, :882, :893) and patches the
def-time-bound `dowsample=` / `upsample=` defaults of `Uformer.__init__` (model.py:1076); it does NOT
rebind the `Downsample` / `Upsample` globals, because the originals' `super(Downsample, self)` calls
(model.py:732, :758) resolve that global.  See INTEGRATION.md.
"""
from .modules import (Downsample, DropPath, EngineUnavailable, LeFF, LeWinTransformerBlock, LinearProjection,  # noqa: F401
                      Upsample, WindowAttention, set_residual_precision)
from .inference import expand2square, restore_image  # noqa: F401
from .network import GraphedForward, InputProj, LeWinStage, OutputProj, Uformer  # noqa: F401
from .training import CharbonnierLoss, FlatAdamW, FlatArena, GradReducer, TrainStep  # noqa: F401

__all__ = ["install", "uninstall", "LeWinTransformerBlock", "WindowAttention", "LeFF", "Downsample", "Upsample", "Uformer",
           "EngineUnavailable", "restore_image", "expand2square", "set_residual_precision", "TrainStep", "CharbonnierLoss", "FlatAdamW", "FlatArena", "GradReducer"]

_SAVED = {}


def install(model_module):
    """Patch the reference's `model` module in place; returns it.  Idempotent."""
    if id(model_module) in _SAVED:
        return model_module
    init = model_module.Uformer.__init__
    names = init.__code__.co_varnames[:init.__code__.co_argcount]
    defaults = list(init.__defaults__)
    off = len(names) - len(defaults)
    saved = dict(LeWinTransformerBlock=model_module.LeWinTransformerBlock, WindowAttention=model_module.WindowAttention,
                 LeFF=model_module.LeFF, defaults=init.__defaults__)
    defaults[names.index("dowsample") - off] = Downsample
    defaults[names.index("upsample") - off] = Upsample
    init.__defaults__ = tuple(defaults)
    model_module.LeWinTransformerBlock = LeWinTransformerBlock
    model_module.WindowAttention = WindowAttention
    model_module.LeFF = LeFF
    _SAVED[id(model_module)] = saved
    return model_module


def uninstall(model_module):
    saved = _SAVED.pop(id(model_module), None)
    if saved is None:
        return
    model_module.Uformer.__init__.__defaults__ = saved["defaults"]
    for k in ("LeWinTransformerBlock", "WindowAttention", "LeFF"):
        setattr(model_module, k, saved[k])
