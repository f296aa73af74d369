"""autograd glue of the training path: native forward, recompute-from-input backward.

``NativeFn.apply(run_native, restate, n_act, *tensors)`` runs ``run_native(*activations)`` (the sm_100a
kernels, no autograd graph) and saves only its inputs.  ``backward`` re-materialises the op with
``restate(*activations)`` (uformer_b200/restated.py — torch statements under bf16 autocast, so the
GEMMs are cuBLAS bf16 with fp32 accumulation and LayerNorm/softmax run in fp32 like the reference's
autocast training step, train/train_denoise.py:178-180) and differentiates that.  The parameters are
passed as explicit inputs so their gradients accumulate into ``.grad`` (the flat gradient arena of
uformer_b200/training.py) through the normal autograd accumulation.

SURVEY §7.1-9: "backward kernels (recompute-from-block-input)" — the recompute structure is final;
the restated statements are replaced by hand-written kernels next round.
"""
from __future__ import annotations

import torch

Tensor = torch.Tensor


class NativeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, run_native, restate, n_act, *tensors):
        ctx.restate, ctx.n_act = restate, n_act
        acts = tensors[:n_act]
        with torch.no_grad():
            out = run_native(*acts)
        ctx.save_for_backward(*tensors)
        return out

    @staticmethod
    def backward(ctx, g):
        tensors = ctx.saved_tensors
        n_act = ctx.n_act
        need = ctx.needs_input_grad[3:]
        dev = g.device.type
        amp = autocast_enabled(dev)
        acts = []
        for i, t in enumerate(tensors[:n_act]):
            t = t.detach()
            if not amp and t.dtype == torch.bfloat16:
                t = t.float()                                # no autocast (CPU test runs): plain fp32 statements
            if need[i] and torch.is_floating_point(t):
                t.requires_grad_(True)
            acts.append(t)
        params = tensors[n_act:]
        with torch.enable_grad(), torch.autocast(device_type=dev, dtype=torch.bfloat16, enabled=amp):
            out = ctx.restate(*acts)
        wrt = [a for a in acts if a.requires_grad] + [p for i, p in enumerate(params) if need[n_act + i]]
        grads = torch.autograd.grad(out, wrt, g.to(out.dtype), allow_unused=True)
        it = iter(grads)
        res = [next(it) if a.requires_grad else None for a in acts]
        res += [next(it) if need[n_act + i] else None for i in range(len(params))]
        res = [r if r is None or r.dtype == t.dtype else r.to(t.dtype) for r, t in zip(res, tensors)]   # match input dtypes
        return (None, None, None, *res)


_AUTOCAST = {"cuda": True, "cpu": False}


def autocast_enabled(device_type: str) -> bool:
    """bf16 autocast for the backward recompute: on for CUDA (training dtype of BASELINE configs[2]); off on CPU, where
    only the test suite calls the restated statements and wants fp32 gradients to compare with the reference's."""
    return _AUTOCAST.get(device_type, False)


def trainable_tensors(mod):
    """Trainable parameter tensors of `mod`, also for nn.DataParallel replicas (train/train_denoise.py:83): torch's
    `_replicate_for_data_parallel` empties `_parameters` in every replica and parks the broadcast copies (autograd
    non-leaf tensors that route gradients back to the source module) in `_former_parameters` — `mod.parameters()` is
    empty there, which would silently drop every weight gradient."""
    out, seen = [], set()
    for m in mod.modules():
        for p in list(m._parameters.values()) + list(getattr(m, "_former_parameters", {}).values()):
            if p is not None and p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                out.append(p)
    return out


def wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def apply(run_native, restate, acts, params):
    """acts: list of activation tensors (inputs of both closures); params: parameters read by the closures."""
    return NativeFn.apply(run_native, restate, len(acts), *acts, *params)
