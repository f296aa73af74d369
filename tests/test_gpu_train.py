"""GPU: the training step (BASELINE configs[2]) — native forward under autograd, recompute backward, native
Charbonnier loss and AdamW — against the gradients the unmodified reference produced (tests/golden/train_t2_128.pt,
fp32 CPU) and against torch's own optimizer / autograd on the same device.

Tolerances: the forward runs in bf16 kernels and backward under bf16 autocast while the golden is fp32, so gradients
are compared at 5e-2 relative L2 over the concatenated per-parameter samples (north_star states 1e-2 for forward
outputs only; measured gradient noise of the reference itself under bf16 autocast is ~1e-2)."""
import pytest
import torch

from helpers import load_golden, rel_l2
from paramgen import randomize_state

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_charbonnier_native_vs_torch():
    from uformer_b200.training import CharbonnierLoss
    torch.manual_seed(0)
    for shape in [(2, 3, 64, 64), (1, 3, 17, 19), (3, 1, 5, 7)]:          # incl. sizes that are not multiples of 4
        x = torch.rand(*shape, device=DEV, requires_grad=True)
        y = torch.rand(*shape, device=DEV)
        loss = CharbonnierLoss(1e-3)(x, y)
        loss.backward()
        xr = x.detach().clone().requires_grad_(True)
        ref = torch.sqrt((xr - y) ** 2 + 1e-6).mean()
        ref.backward()
        assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
        assert rel_l2(x.grad, xr.grad) < 1e-5


def test_adamw_native_vs_torch():
    import torch.nn as nn
    from uformer_b200.training import FlatAdamW, FlatArena
    torch.manual_seed(1)
    net = nn.Sequential(nn.Linear(37, 53), nn.LayerNorm(53), nn.Linear(53, 11)).to(DEV)
    ref = nn.Sequential(nn.Linear(37, 53), nn.LayerNorm(53), nn.Linear(53, 11)).to(DEV)
    ref.load_state_dict(net.state_dict())
    arena = FlatArena(list(net.parameters())[::-1])
    opt = FlatAdamW(arena, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.02)
    topt = torch.optim.AdamW(ref.parameters(), lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.02)
    for it in range(4):
        x = torch.randn(8, 37, device=DEV)
        net(x).pow(2).mean().backward()
        arena.grad.mul_(2.0)                         # pretend a 2-rank sum all-reduce happened: step() averages
        opt.step(grad_scale=0.5, zero_grad=True)
        assert arena.grad.abs().sum().item() == 0.0  # zeroed in the same pass
        topt.zero_grad()
        ref(x).pow(2).mean().backward()
        topt.step()
        for a, b in zip(net.parameters(), ref.parameters()):
            assert rel_l2(a.detach(), b.detach()) < 1e-5, it


def _engine_net(g):
    import uformer_b200 as U
    net = U.Uformer(**g["cfg"])
    net.load_state_dict(randomize_state(net.state_dict(), g["seed"]), strict=True)
    return net.to(DEV).train()


def test_train_gradients_vs_reference_golden():
    from uformer_b200.training import CharbonnierLoss
    g = load_golden("train_t2_128")
    net = _engine_net(g)
    restored = net(g["x"].to(DEV))
    assert restored.requires_grad
    loss = CharbonnierLoss()(restored, g["target"].to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - g["loss"]) < 2e-2 * abs(g["loss"])
    got, want, per = [], [], {}
    for k, p in net.named_parameters():
        ref = g["grads"][k]
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        s = p.grad.detach().float().cpu().reshape(-1)[::ref["stride"]]
        got.append(s)
        want.append(ref["sample"])
        per[k] = rel_l2(s, ref["sample"])
    e = rel_l2(torch.cat(got), torch.cat(want))
    worst = sorted(per.items(), key=lambda kv: -kv[1])[:5]
    print(f"train grads: rel-L2 over all sampled gradients {e:.3e}; worst params {worst}")
    assert e < 5e-2, (e, worst)


def test_train_step_updates_weights_and_reduces_loss():
    """Five TrainStep iterations on one batch: the loss must fall, which also proves the packed operand images are
    rebuilt after the native optimizer wrote the arena (a stale image would freeze the forward)."""
    from uformer_b200.training import TrainStep
    g = load_golden("train_t2_128")
    net = _engine_net(g)
    step = TrainStep(net, lr=2e-4, weight_decay=0.0)
    x, t = g["x"].to(DEV), g["target"].to(DEV)
    losses = [step(x, t).item() for _ in range(5)]
    print("losses:", losses)
    assert all(torch.isfinite(torch.tensor(losses)))
    assert len(set(round(v, 6) for v in losses)) == 5
    assert losses[-1] < losses[0]


def test_block_training_mode_stochastic_depth():
    """Training-mode block with DropPath (model.py:887, :986-987): native kernels + per-sample branch scaling against the
    CPU oracle fed the same per-sample factors (drawn from the same RNG state in the reference's order)."""
    import uformer_b200 as U
    from oracle import lewin_oracle as O
    blk = U.LeWinTransformerBlock(64, (16, 16), 2, win_size=8, shift_size=4, modulator=True, drop_path=0.5)
    st = randomize_state(blk.state_dict(), 17)
    blk.load_state_dict(st)
    blk = blk.to(DEV).train()
    x = torch.randn(6, 256, 64, device=DEV).to(torch.bfloat16)
    torch.manual_seed(5)
    with torch.no_grad():
        y = blk(x).float()
    torch.manual_seed(5)
    s1, s2 = blk.drop_path.draw(6, x.device), blk.drop_path.draw(6, x.device)
    assert 0 < int((s1 == 0).sum() + (s2 == 0).sum()) < 12           # some branches dropped, some kept
    ref = O.lewin_block(x.float().cpu(), st, "", 2, 8, 4, drop_scales=(s1.cpu(), s2.cpu()))
    e = rel_l2(y.cpu(), ref)
    print(f"droppath block vs oracle: {e:.3e}")
    assert e < 1e-2
