"""Training step of BASELINE configs[2]: fwd + bwd + gradient all-reduce + AdamW, one process per GPU.

Reference being replaced (train/train_denoise.py:77-83, :164-185): ``optim.AdamW(model.parameters(), lr,
betas=(0.9, 0.999), eps=1e-8, weight_decay)``, ``nn.DataParallel`` (scatter the batch, replicate the weights and
reduce gradients onto GPU 0 every iteration, all from Python threads of one process), ``CharbonnierLoss`` under
fp16 autocast with a GradScaler.  The B200 design:

* ``FlatArena``   — every trainable parameter lives in ONE flat fp32 buffer and every gradient in another, laid out
                    in reverse execution order (output projection first), so buckets of the gradient arena
                    become final in the order backward produces them;
* ``GradReducer`` — bucketed NCCL sum all-reduce of the gradient arena over NVLink/NVSwitch, launched from
                    post-accumulate hooks while backward is still running (no gradient copy: NCCL reduces the
                    arena in place); replaces DataParallel's per-iteration replicate + reduce;
* ``FlatAdamW``   — ONE native kernel launch over the arenas (lw_adamw_step) that also applies the 1/world averaging
                    and zeroes the gradient arena for the next step;
* ``CharbonnierLoss`` — loss and d(loss)/d(restored) in one native pass (lw_charbonnier_fwd_bwd);
* bf16 (no loss scaling needed: bf16 has fp32's exponent range, so the reference's GradScaler has no counterpart).

``TrainStep`` wires them together.  Host logic (arena layout, bucketing, hook-driven all-reduce) is device
agnostic and covered by world_size-2 gloo tests on CPU; the optimizer and loss kernels exist only for B200 and
raise ``EngineUnavailable`` anywhere else.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from . import modules, ops
from .prepack import prepack
from ._lib import EngineUnavailable

Tensor = torch.Tensor
_ALIGN = 64                     # elements (256 B): every parameter view starts on a sector boundary


# ------------------------------------------------------------------------------------------------------------------
def execution_ordered_parameters(net: nn.Module):
    """Trainable parameters in forward execution order.  For the engine's Uformer (and the reference's, which uses
    the same attribute names) that is input_proj, encoder stages interleaved with dowsample_i, the bottleneck,
    upsample_j / decoder stages, output_proj (model.py:1269-1305); any other module: registration order."""
    names = ["input_proj"]
    for i in range(4):
        names += [f"encoderlayer_{i}", f"dowsample_{i}"]
    names.append("conv")
    for j in range(4):
        names += [f"upsample_{j}", f"decoderlayer_{j}"]
    names.append("output_proj")
    if all(hasattr(net, n) for n in names):
        seen, out = set(), []
        for n in names:
            for p in getattr(net, n).parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    out.append(p)
        rest = [p for p in net.parameters() if id(p) not in seen]
        out = out + rest
    else:
        out = list(net.parameters())
    return [p for p in out if p.requires_grad]


class FlatArena:
    """Re-homes `params` into one flat fp32 buffer (`flat`) and their gradients into another (`grad`); afterwards
    ``p.data`` and ``p.grad`` are views, so autograd accumulates straight into the arena and the optimizer /
    all-reduce see one contiguous range.  Build it AFTER moving the module to its device."""

    def __init__(self, params):
        self.params = list(params)
        if not self.params:
            raise ValueError("FlatArena needs at least one parameter")
        dev = self.params[0].device
        self.offsets = []
        off = 0
        for p in self.params:
            if p.device != dev or p.dtype != torch.float32:
                raise ValueError("FlatArena expects fp32 parameters on one device")
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                n = p.numel()
                self.flat[o:o + n].view_as(p).copy_(p.data)
                p.data = self.flat[o:o + n].view_as(p)
                p.grad = self.grad[o:o + n].view_as(p)

    def zero_grad(self):
        self.grad.zero_()
        for p, o in zip(self.params, self.offsets):       # a caller may have set .grad = None
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view_as(p)


# ------------------------------------------------------------------------------------------------------------------
class GradReducer:
    """Sum all-reduce of the gradient arena in buckets, overlapped with backward.

    Buckets are contiguous arena ranges of >= bucket_bytes (the arena is in reverse execution order, so bucket 0
    holds the parameters whose gradients arrive first).  A post-accumulate-grad hook per parameter counts arrivals;
    when a bucket is complete its range is all-reduced asynchronously (NCCL: on the communicator's own stream, so
    it overlaps the rest of backward).  `finish()` launches whatever did not fire (unused parameters) and waits.
    The division by world size is folded into the optimizer kernel (FlatAdamW.step(grad_scale=1/world))."""

    def __init__(self, arena: FlatArena, process_group=None, bucket_bytes: int = 32 << 20, enabled: bool = True):
        self.arena = arena
        self.group = process_group
        self.world = dist.get_world_size(process_group) if enabled and dist.is_available() and dist.is_initialized() else 1
        self.buckets = []                # (lo, hi, [param indices])
        lo, members = 0, []
        for i, (p, o) in enumerate(zip(arena.params, arena.offsets)):
            members.append(i)
            hi = o + (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            if (hi - lo) * 4 >= bucket_bytes:
                self.buckets.append((lo, hi, members))
                lo, members = hi, []
        if members:
            self.buckets.append((lo, arena.numel, members))
        self._bucket_of = {}
        for b, (_, _, mem) in enumerate(self.buckets):
            for i in mem:
                self._bucket_of[i] = b
        self._pending = [0] * len(self.buckets)
        self._fired = [False] * len(self.buckets)
        self._works = []
        self.launch_order = []           # bucket ids in the order they were reduced (introspection / tests)
        self._hooks = []
        if self.world > 1:
            for i, p in enumerate(arena.params):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))
        self.begin()

    def _make_hook(self, i):
        def hook(_param):
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0 and not self._fired[b]:
                self._launch(b)
        return hook

    def _launch(self, b):
        lo, hi, _ = self.buckets[b]
        self._fired[b] = True
        self.launch_order.append(b)
        self._works.append(dist.all_reduce(self.arena.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def begin(self):
        """Call before each backward."""
        self._pending = [len(mem) for (_, _, mem) in self.buckets]
        self._fired = [False] * len(self.buckets)
        self._works = []
        self.launch_order = []

    def finish(self):
        """Call after backward: reduce the buckets that never completed, then make the current stream wait for all."""
        if self.world == 1:
            return
        for b in range(len(self.buckets)):
            if not self._fired[b]:
                self._launch(b)
        for w in self._works:
            w.wait()
        self._works = []

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


# ------------------------------------------------------------------------------------------------------------------
class FlatAdamW:
    """torch.optim.AdamW arithmetic (train/train_denoise.py:77) as one native launch over a FlatArena.  `lr` is a plain
    attribute: the reference's epoch-level schedule (train_denoise.py:88-98, control plane) sets it between steps."""

    def __init__(self, arena: FlatArena, lr: float = 2e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.02):
        self.arena = arena
        self.lr, self.betas, self.eps, self.weight_decay = lr, tuple(betas), eps, weight_decay
        self.exp_avg = torch.zeros_like(arena.flat)
        self.exp_avg_sq = torch.zeros_like(arena.flat)
        self.steps = 0

    def step(self, grad_scale: float = 1.0, zero_grad: bool = True):
        a = self.arena
        self.steps += 1
        ops.adamw_step(a.flat, a.grad, self.exp_avg, self.exp_avg_sq, step=self.steps, lr=self.lr, beta1=self.betas[0],
                       beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay, grad_scale=grad_scale, zero_grad=zero_grad)
        modules.invalidate_packed()          # the kernel wrote the weights behind torch's version counters

    def state_dict(self):
        return dict(exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, steps=self.steps, lr=self.lr, betas=self.betas, eps=self.eps,
                    weight_decay=self.weight_decay)

    def load_state_dict(self, st):
        self.exp_avg.copy_(st["exp_avg"])
        self.exp_avg_sq.copy_(st["exp_avg_sq"])
        self.steps = int(st["steps"])
        self.lr, self.betas, self.eps, self.weight_decay = st["lr"], tuple(st["betas"]), st["eps"], st["weight_decay"]


# ------------------------------------------------------------------------------------------------------------------
class _CharbonnierFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, eps):
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        loss, grad = ops.charbonnier(x.contiguous(), y.contiguous(), eps, need)
        ctx.grad = grad
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        gx = ctx.grad * g if ctx.grad is not None else None
        return (gx if ctx.needs_input_grad[0] else None, -gx if ctx.needs_input_grad[1] else None, None)


class CharbonnierLoss(nn.Module):
    """losses.py:41-52: mean(sqrt((x - y)^2 + eps^2)); forward and backward are one native pass."""

    def __init__(self, eps: float = 1e-3):
        super().__init__()
        self.eps = eps

    def forward(self, x, y):
        return _CharbonnierFn.apply(x.float(), y.float(), float(self.eps))


def mixup(rgb_gt: Tensor, rgb_noisy: Tensor, generator=None):
    """MixUp_AUG.aug (utils/dataset_utils.py:34-49): convex combination of each pair with a random partner,
    lam ~ Beta(1.2, 1.2) per sample (used after epoch 5, train_denoise.py:176-177)."""
    bs = rgb_gt.size(0)
    idx = torch.randperm(bs, generator=generator).to(rgb_gt.device)
    lam = torch.distributions.beta.Beta(torch.tensor([1.2]), torch.tensor([1.2])).rsample((bs, 1)).view(-1, 1, 1, 1).to(rgb_gt.device)
    return lam * rgb_gt + (1 - lam) * rgb_gt[idx], lam * rgb_noisy + (1 - lam) * rgb_noisy[idx]


# ------------------------------------------------------------------------------------------------------------------
class TrainStep:
    """One data-parallel training step (train/train_denoise.py:171-185) on this rank's GPU.

        step = TrainStep(net)                  # net: uformer_b200.Uformer (or the reference's, after install()) on cuda
        loss = step(noisy, clean)              # device scalar; no host sync inside
    """

    def __init__(self, net: nn.Module, lr: float = 2e-4, weight_decay: float = 0.02, betas=(0.9, 0.999), eps: float = 1e-8,
                 process_group=None, bucket_bytes: int = 32 << 20, criterion: nn.Module | None = None, batched_repack: bool = True,
                 data_parallel: bool = True):
        """data_parallel=False: a purely local step even inside an initialised process group (no all-reduce, no broadcast)."""
        self.net = net
        self.batched_repack = batched_repack
        params = execution_ordered_parameters(net)[::-1]                 # reverse execution order
        self.arena = FlatArena(params)
        self.reducer = GradReducer(self.arena, process_group, bucket_bytes, enabled=data_parallel)
        self.optimizer = FlatAdamW(self.arena, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.criterion = criterion if criterion is not None else CharbonnierLoss()
        self.world = self.reducer.world
        self.sync_from_rank0()

    def sync_from_rank0(self):
        """Make every rank start from rank 0's weights and optimizer state (ranks that initialise differently — rank-dependent
        seeds, a partial checkpoint load — would otherwise silently train diverged replicas: the update is applied locally on
        every rank from the all-reduced gradient, weights are never re-broadcast).  nn.DataParallel did this every iteration by
        replicating GPU 0's module (train/train_denoise.py:83)."""
        if self.world > 1:
            group = self.reducer.group
            dist.broadcast(self.arena.flat, src=0, group=group)
            dist.broadcast(self.optimizer.exp_avg, src=0, group=group)
            dist.broadcast(self.optimizer.exp_avg_sq, src=0, group=group)
            steps = torch.tensor([self.optimizer.steps], dtype=torch.int64, device=self.arena.flat.device)
            dist.broadcast(steps, src=0, group=group)
            self.optimizer.steps = int(steps.item())
            modules.invalidate_packed()

    def __call__(self, input_: Tensor, target: Tensor) -> Tensor:
        self.net.train()
        self.reducer.begin()
        restored = self.net(input_)
        loss = self.criterion(restored, target)
        loss.backward()
        self.reducer.finish()
        self.optimizer.step(grad_scale=1.0 / self.world, zero_grad=True)
        if self.batched_repack:
            prepack(self.net)                # one permutation per (stage, weight kind) instead of ~6 launches per weight
        return loss.detach()

    # ---- checkpoint / resume (the reference saves {'epoch', 'state_dict', 'optimizer'}, train_denoise.py:222-236) ----
    def state_dict(self):
        """Model weights under the reference's key names (parameters are views of the arena, so this is a snapshot of
        it in checkpoint layout) + the optimizer moments in arena layout."""
        return dict(state_dict={k: v.detach().clone() for k, v in self.net.state_dict().items()},
                    optimizer={k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in self.optimizer.state_dict().items()})

    def load_state_dict(self, st):
        """`st["state_dict"]` may be a reference checkpoint's (keys with the DataParallel `module.` prefix are accepted, like
        utils/model_utils.py:23-33 does); `st["optimizer"]` (optional) is this class's arena-layout moment dict — the reference's
        torch.optim.AdamW per-parameter state is a different layout and is not loadable here (resume then restarts the moments)."""
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in st["state_dict"].items()}
        with torch.no_grad():
            own = self.net.state_dict()
            missing = set(own) ^ set(sd)
            if missing:
                raise KeyError(f"checkpoint / model key mismatch: {sorted(missing)[:5]}")
            for k, v in sd.items():
                own[k].copy_(v)                                      # in place: the arena views stay intact
        if st.get("optimizer") is not None and "exp_avg" in st["optimizer"]:
            self.optimizer.load_state_dict(st["optimizer"])
        self.arena.zero_grad()
        modules.invalidate_packed()
        self.sync_from_rank0()


__all__ = ["FlatArena", "GradReducer", "FlatAdamW", "CharbonnierLoss", "TrainStep", "mixup",
           "execution_ordered_parameters", "EngineUnavailable"]
