"""Generate golden input/output vectors from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/model.py behind the timm shim (tests/refshim.py), builds reference modules
with the seeded synthetic weights of tests/paramgen.py, runs the reference forward in fp32 on CPU
and stores inputs + outputs as fp32 .pt files in this directory.  Weights are NOT stored: they are
re-derived from the seed (paramgen.randomize_state is deterministic), the fixture keeps a checksum.
The reference cannot travel to the GPU box, these files can.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from refshim import import_reference_model  # noqa: E402
from paramgen import randomize_state  # noqa: E402


def state_checksum(state):
    s = 0.0
    for k in sorted(state):
        if torch.is_floating_point(state[k]):
            s += float(state[k].double().abs().sum())
    return s


def load_random(mod, seed):
    st = randomize_state(mod.state_dict(), seed)
    mod.load_state_dict(st, strict=True)
    mod.eval()
    return st


def ws16_fixtures(m):
    """BASELINE configs[4], win_size 16: WindowAttention on 256-token windows (head_dim 16 / 32 / 64) and two LeWin blocks with
    16x16 windows (shift 8 + modulator; shift 0 at C = 256).  Own RNG, so the older fixtures stay reproducible."""
    out = {}
    gen = torch.Generator().manual_seed(1616)
    for name, dim, heads, nwin in [("wattn_ws16_c32_h1", 32, 1, 4), ("wattn_ws16_c64_h4_hd16", 64, 4, 3), ("wattn_ws16_c128_h2_hd64", 128, 2, 2)]:
        mod = m.WindowAttention(dim, win_size=(16, 16), num_heads=heads)
        st = load_random(mod, 21)
        x = torch.randn(nwin, 256, dim, generator=gen)
        mask = torch.where(torch.rand(nwin if nwin < 3 else 1, 256, 256, generator=gen) > 0.7, torch.tensor(-100.0), torch.tensor(0.0))
        out[name] = dict(kind="wattn", ws=16, dim=dim, heads=heads, seed=21, x=x, y=mod(x), mask=mask, y_mask=mod(x, mask=mask),
                         checksum=state_checksum(st))
    for name, dim, heads, H, shift, modu, B in [("block_ws16_c64_s8_mod", 64, 2, 32, 8, True, 2), ("block_ws16_c256_s0", 256, 8, 16, 0, False, 1)]:
        mod = m.LeWinTransformerBlock(dim, (32, 32), heads, win_size=16, shift_size=shift, modulator=modu)
        st = load_random(mod, 22)
        x = torch.randn(B, H * H, dim, generator=gen)
        out[name] = dict(kind="block", ws=16, dim=dim, heads=heads, H=H, shift=shift, modulator=modu, seed=22, x=x, y=mod(x),
                         checksum=state_checksum(st))
    return out


def save_all(out, only=None):
    for k, v in out.items():
        if only and not k.startswith(only):
            continue
        path = os.path.join(HERE, k + ".pt")
        torch.save(v, path)
        print("%-28s %8.1f KB" % (k, os.path.getsize(path) / 1024))


def main():
    torch.manual_seed(1234)
    torch.set_grad_enabled(False)
    m = import_reference_model()
    if os.environ.get("GOLDEN_ONLY") == "ws16":          # only the 16x16-window fixtures (skips the slow whole-model forwards)
        save_all(ws16_fixtures(m))
        return
    out = {}

    # ---- module level: WindowAttention (model.py:452-546), with and without mask ----
    for name, dim, heads, nwin in [("wattn_c32_h1", 32, 1, 6), ("wattn_c128_h4", 128, 4, 8), ("wattn_c64_h4_hd16", 64, 4, 4)]:
        mod = m.WindowAttention(dim, win_size=(8, 8), num_heads=heads)
        st = load_random(mod, 11)
        x = torch.randn(nwin, 64, dim)
        mask = torch.where(torch.rand(2, 64, 64) > 0.7, torch.tensor(-100.0), torch.tensor(0.0))
        out[name] = dict(kind="wattn", dim=dim, heads=heads, seed=11, x=x, y=mod(x), mask=mask,
                         y_mask=mod(x, mask=mask), checksum=state_checksum(st))

    # ---- LeFF (model.py:654-699) ----
    for name, dim, B, H in [("leff_c32", 32, 2, 16), ("leff_c128", 128, 1, 24)]:
        mod = m.LeFF(dim, 4 * dim)
        st = load_random(mod, 12)
        x = torch.randn(B, H * H, dim)
        out[name] = dict(kind="leff", dim=dim, seed=12, x=x, y=mod(x), checksum=state_checksum(st))

    # ---- Downsample / Upsample (model.py:730-778) ----
    mod = m.Downsample(32, 64)
    st = load_random(mod, 13)
    x = torch.randn(2, 16 * 16, 32)
    out["down_32_64"] = dict(kind="down", cin=32, cout=64, seed=13, x=x, y=mod(x), checksum=state_checksum(st))
    mod = m.Upsample(64, 16)
    st = load_random(mod, 14)
    x = torch.randn(2, 8 * 8, 64)
    out["up_64_16"] = dict(kind="up", cin=64, cout=16, seed=14, x=x, y=mod(x), checksum=state_checksum(st))

    # ---- LeWinTransformerBlock (model.py:850-1008): shift 0/4, modulator on/off ----
    for name, dim, heads, H, shift, modu in [("block_c32_s0", 32, 1, 16, 0, False), ("block_c64_s4_mod", 64, 2, 24, 4, True),
                                            ("block_c128_s4", 128, 4, 16, 4, False), ("block_c32_s0_mod_hd16", 32, 2, 16, 0, True)]:
        mod = m.LeWinTransformerBlock(dim, (H, H), heads, win_size=8, shift_size=shift, modulator=modu)
        st = load_random(mod, 15)
        x = torch.randn(2, H * H, dim)
        out[name] = dict(kind="block", dim=dim, heads=heads, H=H, shift=shift, modulator=modu, seed=15,
                         x=x, y=mod(x), checksum=state_checksum(st))

    # ---- whole model, BASELINE config #1: Uformer-T-like, 128x128, batch 1 (depths=[1]*9) ----
    cfg = dict(img_size=128, embed_dim=16, depths=[1] * 9, win_size=8, token_projection="linear", token_mlp="leff", modulator=True)
    net = m.Uformer(**cfg)
    st = load_random(net, 1234)
    x = torch.rand(1, 3, 128, 128)
    out["uformer_t1_128"] = dict(kind="model", cfg=cfg, seed=1234, x=x, y=net(x), checksum=state_checksum(st))
    # real Uformer_T depths ([2]*9) at 128x128: exercises shifted blocks at every stage
    cfg2 = dict(img_size=128, embed_dim=16, depths=[2] * 9, win_size=8, token_projection="linear", token_mlp="leff", modulator=True)
    net = m.Uformer(**cfg2)
    st = load_random(net, 77)
    x = torch.rand(2, 3, 128, 128)
    out["uformer_t2_128"] = dict(kind="model", cfg=cfg2, seed=77, x=x, y=net(x), checksum=state_checksum(st))
    # arbitrary-resolution path: model built for 128, run at 256 (whole-image forward)
    x = torch.rand(1, 3, 256, 256)
    out["uformer_t2_128_at256"] = dict(kind="model", cfg=cfg2, seed=77, x=x, y=net(x), checksum=state_checksum(st))
    # embed_dim=32 (head_dim 32 as in Uformer-S/B), 2 blocks per stage, 128x128
    cfg3 = dict(img_size=128, embed_dim=32, depths=[2, 2, 2, 2, 2, 2, 2, 2, 2], win_size=8, token_projection="linear", token_mlp="leff", modulator=True)
    net = m.Uformer(**cfg3)
    st = load_random(net, 5)
    x = torch.rand(1, 3, 128, 128)
    out["uformer_s2_128"] = dict(kind="model", cfg=cfg3, seed=5, x=x, y=net(x), checksum=state_checksum(st))

    # ---- the flagship: Uformer-B 256x256 (BASELINE configs[1]; utils/model_utils.py:76-78), one image ----
    # (a) the bench's weights (seed 1234, O(1) activations through all 40 blocks): after 40 blocks even the reference's
    #     own bf16-autocast forward is ~1.3e-2 away from its fp32 forward, so that error is recorded as the yardstick;
    # (b) the same weights scaled by 0.5 (residual branch ~0.1, like a trained denoiser): plain 1e-2 bound.
    cfgb = dict(img_size=256, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff", depths=[1, 2, 8, 8, 2, 8, 8, 2, 1],
                modulator=True, dd_in=3)
    x = torch.rand(1, 3, 256, 256)
    for name, gain in [("uformer_b_256", 1.0), ("uformer_b_256_g05", 0.5)]:
        net = m.Uformer(**cfgb)
        st = randomize_state(net.state_dict(), 1234, gain)
        net.load_state_dict(st, strict=True)
        net.eval()
        y = net(x)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            yb = net(x).float()
        rl2 = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())   # noqa: E731
        out[name] = dict(kind="model", cfg=cfgb, seed=1234, gain=gain, x=x, y=y, checksum=state_checksum(st),
                         ref_bf16=dict(full=rl2(yb, y), resid=rl2(yb - x, y - x)))
        print(name, "reference bf16-autocast vs fp32:", out[name]["ref_bf16"])

    # ---- BASELINE configs[3]: Uformer-B built for 256x256 run on a 512x512 image in ONE whole-image forward (the reference has
    # no tiling: test/test_sidd.py:79-108 pads to a square multiple of 128 and calls the model once).  Weights: the
    # gain-0.5 set (trained-denoiser-like residual branch).  The input is re-derivable: torch.rand under seed 512 (not stored).
    net = m.Uformer(**cfgb)
    st = randomize_state(net.state_dict(), 1234, 0.5)
    net.load_state_dict(st, strict=True)
    net.eval()
    x512 = torch.rand(1, 3, 512, 512, generator=torch.Generator().manual_seed(512))
    y512 = net(x512)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        yb = net(x512).float()
    out["uformer_b_512_g05"] = dict(kind="model", cfg=cfgb, seed=1234, gain=0.5, x_seed=512, x_shape=(1, 3, 512, 512), y=y512,
                                    checksum=state_checksum(st), ref_bf16=dict(full=rl2(yb, y512), resid=rl2(yb - x512, y512 - x512)))
    print("uformer_b_512_g05 reference bf16-autocast vs fp32:", out["uformer_b_512_g05"]["ref_bf16"])

    # ---- BASELINE configs[4] point: head_dim 64 (embed_dim 64, 2 heads) WindowAttention, own RNG so older fixtures stay reproducible ----
    gen = torch.Generator().manual_seed(6464)
    mod = m.WindowAttention(128, win_size=(8, 8), num_heads=2)
    st = load_random(mod, 16)
    xw = torch.randn(6, 64, 128, generator=gen)
    maskw = torch.where(torch.rand(3, 64, 64, generator=gen) > 0.7, torch.tensor(-100.0), torch.tensor(0.0))
    out["wattn_c128_h2_hd64"] = dict(kind="wattn", dim=128, heads=2, seed=16, x=xw, y=mod(xw), mask=maskw, y_mask=mod(xw, mask=maskw),
                                     checksum=state_checksum(st))

    out.update(ws16_fixtures(m))
    save_all(out, os.environ.get("GOLDEN_ONLY"))


if __name__ == "__main__":
    main()
