"""CPU: the explicit LeWin-block backward (uformer_b200/block_bwd.py, the specification of the native backward kernels)
against torch autograd through the restated forward — every parameter gradient and the input gradient, with shift,
modulator, stochastic depth and the input mask; fp32 for the math, bf16 to exercise the GPU's dtype flow on CPU."""
import pytest
import torch

import kernel_model as KM
import uformer_b200 as U
from uformer_b200 import autograd as AG
from uformer_b200 import restated as R
from uformer_b200.block_bwd import lewin_block_bwd
from helpers import load_golden, rel_l2
from paramgen import randomize_state

CASES = [(32, 1, 16, 0, False, False, False), (64, 2, 24, 4, True, True, False), (32, 2, 16, 4, True, False, True),
         (64, 4, 16, 0, False, True, True), (128, 4, 8, 0, True, False, False)]


def _setup(dim, heads, H, shift, modu, dp, msk, seed=21):
    blk = U.LeWinTransformerBlock(dim, (max(H, 16), max(H, 16)), heads, win_size=8, shift_size=shift, modulator=modu)
    blk.load_state_dict(randomize_state(blk.state_dict(), seed))
    B = 3
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, H * H, dim, generator=g)
    gout = torch.randn(B, H * H, dim, generator=g)
    s1 = torch.tensor([1.25, 0.0, 1.25]).view(B, 1, 1) if dp else None
    s2 = torch.tensor([0.0, 1.25, 1.25]).view(B, 1, 1) if dp else None
    mask = (torch.rand(B, 1, H, H, generator=g) > 0.5).float() if msk else None
    return blk, x, gout, s1, s2, mask


@pytest.mark.parametrize("dim,heads,H,shift,modu,dp,msk", CASES)
def test_explicit_backward_equals_autograd_fp32(dim, heads, H, shift, modu, dp, msk):
    blk, x, gout, s1, s2, mask = _setup(dim, heads, H, shift, modu, dp, msk)
    xa = x.clone().requires_grad_(True)
    R.lewin_block(blk, xa, mask, s1, s2).backward(gout)
    want = {k: p.grad.clone() for k, p in blk.named_parameters()}
    dx, grads = lewin_block_bwd(blk, x, gout, mask, s1, s2, cd=torch.float32)
    assert rel_l2(dx, xa.grad) < 1e-4
    assert set(grads) == set(want)
    for k, v in want.items():
        assert grads[k].shape == v.shape and grads[k].dtype == torch.float32, k
        assert rel_l2(grads[k], v) < 1e-4, k


@pytest.mark.parametrize("dim,heads,H,shift,modu,dp,msk", CASES[:3])
def test_explicit_backward_bf16_dtype_flow(dim, heads, H, shift, modu, dp, msk):
    """The dtype flow the GPU runs (bf16 operands, fp32 statistics and accumulators), executed on CPU."""
    blk, x, gout, s1, s2, mask = _setup(dim, heads, H, shift, modu, dp, msk)
    xb = x.to(torch.bfloat16)
    xa = xb.float().requires_grad_(True)
    R.lewin_block(blk, xa, mask, s1, s2).backward(gout.to(torch.bfloat16).float())
    dx, grads = lewin_block_bwd(blk, xb, gout.to(torch.bfloat16), mask, s1, s2, cd=torch.bfloat16)
    assert dx.dtype == torch.bfloat16 and rel_l2(dx.float(), xa.grad) < 3e-2
    for k, p in blk.named_parameters():
        assert grads[k].dtype == torch.float32 and torch.isfinite(grads[k]).all(), k
        assert rel_l2(grads[k], p.grad) < 6e-2, (k, rel_l2(grads[k], p.grad))


def test_explicit_backward_through_the_module_wrapper_and_train_step():
    """BlockFn wiring: with the explicit backward selected, TrainStep reproduces the reference's golden gradients and the
    autograd-recompute path's gradients (kernel model standing in for the native forward)."""
    from uformer_b200.training import TrainStep
    g = load_golden("train_t2_128")

    def grads_with(explicit):
        net = U.Uformer(**g["cfg"])
        net.load_state_dict(randomize_state(net.state_dict(), g["seed"]), strict=True)
        AG.use_explicit_block_backward(explicit)
        try:
            with KM.patched():
                step = TrainStep(net, lr=2e-4, weight_decay=0.0)
                net.train()
                loss = step.criterion(net(g["x"]), g["target"])
                loss.backward()
                out = {k: p.grad.clone() for k, p in net.named_parameters()}
                step.optimizer.step(zero_grad=True)
                l2 = step(g["x"], g["target"]).item()
        finally:
            AG.use_explicit_block_backward(False)
        return out, loss.item(), l2
    ge, le, le2 = grads_with(True)
    ga, la, la2 = grads_with(False)
    assert le == la and le2 < le
    got = torch.cat([ge[k].reshape(-1)[::g["grads"][k]["stride"]] for k in ge])
    want = torch.cat([g["grads"][k]["sample"] for k in ge])
    assert rel_l2(got, want) < 5e-2
    for k in ge:
        assert rel_l2(ge[k], ga[k]) < 2e-2, k
