"""Executable model of the C-ABI kernel contracts (include/lewin_b200.h) in torch.  TEST INFRASTRUCTURE ONLY.

Each function has the signature of its namesake in uformer_b200/ops.py and computes what the header says the kernel
computes, FROM THE PACKED OPERAND IMAGES the modules hand over (so it decodes the UMMA tile images, the per-head
[q|k|v] row order, the folded attention scale, the tap-major conv weights, ...).  ``patched()`` swaps these in for
the native entry points so the CPU suite can drive the complete host path — modules, packing, caches, the stage
wiring, autograd wrappers, arena, optimizer — end to end without a GPU.  It is a checker like oracle/: nothing under
uformer_b200/ imports it, and the product still raises EngineUnavailable on CPU.
"""
import contextlib
import math

import torch
import torch.nn.functional as F

from uformer_b200 import _lib, ops, packing

BF = torch.bfloat16


def _q(t):
    """bf16 rounding of a value the kernel stores to HBM / feeds to the tensor core."""
    return t.to(BF).float()


def _relidx(ws=8):
    t = torch.arange(ws * ws)
    y, x = t // ws, t % ws
    return (y[:, None] - y[None, :] + ws - 1) * (2 * ws - 1) + (x[:, None] - x[None, :] + ws - 1)


def _partition(x, ws=8):
    B, H, W, C = x.shape
    return x.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def _reverse(w, H, W, ws=8):
    C = w.shape[-1]
    B = w.shape[0] // ((H // ws) * (W // ws))
    return w.view(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)


def _shift_mask(H, W, shift, ws=8):
    r, c = torch.arange(H), torch.arange(W)
    reg = 3 * ((r >= H - ws).long() + (r >= H - shift).long())[:, None] + ((c >= W - ws).long() + (c >= W - shift).long())[None, :]
    rw = _partition(reg[None, :, :, None].float(), ws).squeeze(-1)
    return torch.where(rw[:, None, :] != rw[:, :, None], -100.0, 0.0)


def _wmsa_tma_eligible(x, p, shift, windowed, x_b):
    """lw_wmsa_fwd's choice of the TMA-gather kernel (include/lewin_b200.h, lw_wmsa_args.wqkv_fold_img)."""
    C = x.shape[-1]
    return ("wqkv_fold_img" in p and not windowed and (p.get("modulator") is None or "wmod_fold_img" in p) and p.get("ln_w") is not None and shift % 4 == 0
            and (x.dtype == BF or x_b is not None) and bool(_lib.load().lw_wmsa_tma_supported(C, p["head_dim"])))


def wmsa(x, p, *, H, W, shift, windowed, resid, mask=None, out=None, out_dtype=None, bf16_copy=False, x_b=None, win=8):
    assert x.dtype in (BF, torch.float32) and x.is_contiguous()
    assert resid is None or resid.dtype == x.dtype
    C = x.shape[-1]
    hd = p["head_dim"]
    heads = C // hd
    assert win in (8, 16) and (win == 8 or _lib.load().lw_wmsa16_supported(C, hd))      # lw_wmsa_args.win_size
    N = win * win
    fold = win == 8 and _wmsa_tma_eligible(x, p, shift, windowed, x_b)
    if fold:
        # TMA-gather kernel: the raw bf16 tokens are the GEMM operand; LayerNorm statistics come from that same bf16 tile
        # and are applied to the accumulator (rstd*acc - rstd*mean*cs + bf)
        assert x_b is None or (x_b.dtype == BF and x_b.shape == x.shape)
        xf = (x if x.dtype == BF else x_b).float()
    else:
        xf = x.float()
        if p.get("ln_w") is not None:
            xf = F.layer_norm(xf, (C,), p["ln_w"], p["ln_b"], p.get("ln_eps", 1e-5))
    if windowed:
        xw = xf
    else:
        B = x.shape[0]
        m = xf.view(B, H, W, C)
        if shift:
            m = torch.roll(m, (-shift, -shift), (1, 2))
        xw = _partition(m, win)
    if p.get("modulator") is not None and not fold:
        xw = xw + p["modulator"]
    xw = _q(xw)                                                   # A operand is bf16
    nW = xw.shape[0]
    if fold:
        wg = packing.unpack_kmajor(p["wqkv_fold_img"], heads * 3 * hd, C, 3 * hd, "nk")
        mean = xw.mean(-1, keepdim=True)
        rstd = torch.rsqrt(((xw - mean) ** 2).mean(-1, keepdim=True) + p.get("ln_eps", 1e-5))
        acc = xw @ wg.t()
        if p.get("modulator") is not None:
            # one-hot k-block: sigma (bf16) on the diagonal x the bf16 image of (m W^T)^T, accumulated with the projection
            mwq = packing.unpack_kmajor(p["wmod_fold_img"], heads * 3 * hd, 64, 3 * hd, "nk")       # columns: quarter-major positions
            mw = torch.empty_like(mwq)
            mw[:, packing.quarter_major_positions()] = mwq                                           # -> natural positions
            acc = acc + _q(1.0 / rstd) * mw.t()[None]
        qkv = (rstd * acc - (rstd * mean) * p["cs_qkv"] + p["bqkv_fold"]).view(nW, 64, heads, 3, hd)
    else:
        wcat = packing.unpack_kmajor(p["wqkv_img"], heads * 3 * hd, C, 3 * hd, "nk")        # [head][q|k|v][hd] rows, q pre-scaled
        qkv = (xw @ wcat.t() + p["bqkv"]).view(nW, N, heads, 3, hd)
    q, k, v = (_q(qkv[:, :, :, i].permute(0, 2, 1, 3)) for i in range(3))
    s = q @ k.transpose(-2, -1) + p["relpos"][:, _relidx(win).reshape(-1)].view(heads, N, N)[None]
    if not windowed and shift:
        sm = _shift_mask(H, W, shift, win)
        s = (s.view(-1, sm.shape[0], heads, N, N) + sm[None, :, None]).view(nW, heads, N, N)
    if mask is not None:
        mk = mask.float()
        s = s + mk[torch.arange(nW) % mk.shape[0]][:, None]
    o = _q(torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(nW, N, C)
    wp = packing.unpack_kmajor(p["wproj_img"], C, C, min(C, 128), "nk")
    y = o @ wp.t() + p["bproj"]
    if not windowed:
        y = _reverse(y, H, W, win)
        if shift:
            y = torch.roll(y, (shift, shift), (1, 2))
        y = y.reshape(x.shape)
    y = _q(y)                                                    # the branch is rounded to bf16 in the staging tile
    if resid is not None:
        y = y + resid.float()
    odt = out.dtype if out is not None else (out_dtype or BF)
    yb = y.to(BF) if bf16_copy else None
    y = y.to(odt)
    if out is not None:
        out.copy_(y)
        y = out
    return (y, yb) if bf16_copy else y


def leff(x, p, *, B, H, W, resid, out=None, out_dtype=None, bf16_copy=False):
    assert x.dtype == BF
    assert not bf16_copy or "w1f_img" in p
    C, hid = x.shape[-1], p["hidden"]
    xf = x.float().reshape(B * H * W, C)
    if "w1f_img" in p:
        # single-kernel contract (lw_leff_fwd): LayerNorm folded into linear1; the raw bf16 x is the GEMM operand
        assert _lib.load().lw_leff_fused_supported(C, hid)
        sl = p["slice"]
        assert sl == _lib.load().lw_leff_slice(C)
        w1g = packing.unpack_kmajor_sw(p["w1f_img"], hid, C, 64, 2 * min(C, 64))      # linear1: 64-row units for every C
        acc = xf @ w1g.t()
        if p["has_ln"]:
            mean = xf.mean(1, keepdim=True)
            rstd = torch.rsqrt(((xf - mean) ** 2).mean(1, keepdim=True) + p.get("ln_eps", 1e-5))
            pre = rstd * acc - (rstd * mean) * p["cs"] + p["b1f"]
        else:
            pre = acc + p["b1f"]
        q16 = lambda t: t.to(torch.float16).float()                # the hidden map is fp16 on chip (E1 -> conv -> GEMM-2)
        h1 = q16(F.gelu(q16(pre)))
        assert p["w2f_img"].dtype == torch.float16 and p["taps"].dtype == torch.float16
        w2 = packing.unpack_kmajor_sw(p["w2f_img"], C, hid, C, 2 * sl)
        t = p["taps"].float().permute(1, 0, 2).reshape(10, hid)       # [NS][10][sl] -> (10, hidden)
        wd_t, bd_t = t[:9], t[9]
    else:
        if p.get("ln_w") is not None:
            xf = F.layer_norm(xf, (C,), p["ln_w"], p["ln_b"], p.get("ln_eps", 1e-5))
        nch1 = min(hid, 256 if C == 256 else 128)                     # lw_nch_ares
        assert nch1 == _lib.load().lw_nch_ares(C, hid)
        w1 = packing.unpack_kmajor(p["w1_img"], hid, C, nch1, "nk")
        q16 = lambda t: t.to(torch.float16).float()
        h1 = q16(F.gelu(q16(_q(xf) @ w1.t() + p["b1"])))              # fp16 round trip through HBM
        assert p["w2_img"].dtype == torch.float16 and p["taps16"].dtype == torch.float16
        w2 = packing.unpack_kmajor(p["w2_img"], C, hid, min(C, 128), "kn")
        wd_t, bd_t = p["taps16"].float()[:9], p["taps16"].float()[9]
    m = h1.view(B, H, W, hid).permute(0, 3, 1, 2)
    wd = wd_t.t().reshape(hid, 1, 3, 3)                           # taps (9, hidden), tap = ky*3+kx
    qh = lambda t: t.to(torch.float16).float()                      # both LeFF paths keep the hidden map in half precision
    h2 = qh(F.gelu(qh(F.conv2d(m, wd, bd_t, padding=1, groups=hid)))).permute(0, 2, 3, 1).reshape(B * H * W, hid)
    y = _q(h2 @ w2.t() + p["b2"]).view(x.shape)                  # the branch is rounded to bf16 before the residual add
    if resid is not None:
        y = y + resid.float()
    odt = out.dtype if out is not None else (out_dtype or BF)
    yb = y.to(BF).contiguous() if bf16_copy else None
    y = y.to(odt)
    if out is not None:
        out.copy_(y)
        y = out
    return (y, yb) if bf16_copy else y


def downsample(x, p, *, B, H, W):
    Cin, Cout = x.shape[-1], p["cout"]
    wk = packing.unpack_kmajor(p["w_img"], Cout, 16 * Cin, min(Cout, 128), "kn")         # K index = (ky*4+kx)*Cin + ci
    w = wk.view(Cout, 4, 4, Cin).permute(0, 3, 1, 2)
    o = F.conv2d(x.float().reshape(B, H, W, Cin).permute(0, 3, 1, 2), w, p["bias"], stride=2, padding=1)
    return o.flatten(2).transpose(1, 2).contiguous().to(BF)


def upsample(x, p, *, B, H, W, out=None):
    Cin, Cout = x.shape[-1], p["cout"]
    nch = _lib.load().lw_nch_ares(Cin, 4 * Cout)
    wn = packing.unpack_kmajor(p["w_img"], 4 * Cout, Cin, nch, "nk")                      # N index = (dy*2+dx)*Cout + co
    y = (x.float().view(B, H, W, Cin) @ wn.t()).view(B, H, W, 2, 2, Cout) + p["bias"]
    y = y.permute(0, 1, 3, 2, 4, 5).reshape(B, 4 * H * W, Cout).to(BF)
    if out is None:
        return y
    out[:, :, :Cout] = y
    return out


def input_proj(img, w, b):
    return F.leaky_relu(F.conv2d(img.float(), w, b, padding=1), 0.01).flatten(2).transpose(1, 2).contiguous().to(BF)


def output_proj(tok, w, b, img, H, W):
    B, _, C = tok.shape
    y = F.conv2d(tok.float().transpose(1, 2).reshape(B, C, H, W), w, b, padding=1)
    return y if img is None else img.float() + y


def charbonnier(x, y, eps, need_grad):
    d = x - y
    r = torch.sqrt(d * d + eps * eps)
    return r.mean().view(1), (d / r / x.numel() if need_grad else None)


def adamw_step(p, g, m, v, *, step, lr, beta1, beta2, eps, weight_decay, grad_scale, zero_grad):
    gg = g * grad_scale
    p.mul_(1 - lr * weight_decay)
    m.mul_(beta1).add_(gg, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gg, gg, value=1 - beta2)
    denom = v.sqrt() / math.sqrt(1 - beta2 ** step) + eps
    p.addcdiv_(m, denom, value=-lr / (1 - beta1 ** step))
    if zero_grad:
        g.zero_()


_NAMES = ["wmsa", "leff", "downsample", "upsample", "input_proj", "output_proj", "charbonnier", "adamw_step"]


@contextlib.contextmanager
def patched():
    """Swap the native entry points of uformer_b200.ops (and the device check) for the models above."""
    saved = {n: getattr(ops, n) for n in _NAMES}
    saved_req = _lib.require_device
    calls = {n: 0 for n in _NAMES}

    def counted(n):
        fn = globals()[n]

        def w(*a, **k):
            calls[n] += 1
            return fn(*a, **k)
        return w
    try:
        for n in _NAMES:
            setattr(ops, n, counted(n))
        _lib.require_device = lambda device: None
        yield calls
    finally:
        for n in _NAMES:
            setattr(ops, n, saved[n])
        _lib.require_device = saved_req
