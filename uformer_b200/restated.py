"""Differentiable restatement of the engine's forward ops, used ONLY to build the backward pass.

Round-1 training design: every forward runs in the native sm_100a kernels and keeps
no autograd graph; the only activation a LeWin block saves is its bf16 input.  ``backward``
re-materialises the block with the device-agnostic torch statements below (cuBLAS GEMMs under bf16
autocast on the GPU) and differentiates that — activation checkpointing with a native forward.  The
hand-written backward kernels that replace these statements one by one are the next round's work;
this file is then the parity target they are checked against.

This is NOT a forward fallback: nothing here is reachable from a module's ``forward`` result, the
modules still raise ``EngineUnavailable`` without a B200, and ``oracle/`` is not imported.  The CPU
tests call these functions directly to pin the gradient math against the reference's own autograd
(tests/test_train_cpu.py).

Each function cites the reference statement it follows (``/root/reference/model.py``).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def window_partition(x: Tensor, ws: int) -> Tensor:
    """(B, H, W, C) -> (B*nW, ws*ws, C), window order row-major over (wy, wx), batch-major (model.py:704-715)."""
    B, H, W, C = x.shape
    return x.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def window_reverse(win: Tensor, ws: int, H: int, W: int) -> Tensor:
    """Inverse of window_partition (model.py:717-726): (B*nW, ws*ws, C) -> (B, H, W, C)."""
    C = win.shape[-1]
    B = win.shape[0] // ((H // ws) * (W // ws))
    return win.view(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)


_SHIFT_MASKS = {}


def shift_attn_mask(H: int, W: int, ws: int, shift: int, device) -> Tensor:
    """{0,-100} SW-MSA mask (model.py:924-942) in closed form: in rolled coordinates a token at (r, c) has region id
    3*reg(r)+reg(c), reg(p) = (p >= L-ws) + (p >= L-shift); logits between different regions get -100."""
    key = (H, W, ws, shift, str(device))
    m = _SHIFT_MASKS.get(key)
    if m is None:
        r = torch.arange(H, device=device)
        c = torch.arange(W, device=device)
        reg = 3 * ((r >= H - ws).long() + (r >= H - shift).long())[:, None] + ((c >= W - ws).long() + (c >= W - shift).long())[None, :]
        rw = window_partition(reg[None, :, :, None].float(), ws).squeeze(-1)          # (nW, N)
        m = torch.where(rw[:, None, :] != rw[:, :, None], -100.0, 0.0)
        _SHIFT_MASKS[key] = m
    return m


def input_mask_to_attn_mask(mask: Tensor, H: int, W: int, ws: int) -> Tensor:
    """(B,1,h,w) input mask -> additive (B*nW, N, N) mask (model.py:914-921)."""
    m = window_partition(F.interpolate(mask.float(), size=(H, W)).permute(0, 2, 3, 1), ws).squeeze(-1)
    am = m.unsqueeze(2) * m.unsqueeze(1)
    return torch.where(am != 0, -100.0, 0.0)


def window_attention(attn, xw: Tensor, mask: Tensor | None = None) -> Tensor:
    """WindowAttention.forward (model.py:494-522) with LinearProjection (model.py:431-442) on windows (B_, N, C)."""
    B_, N, C = xw.shape
    h = attn.num_heads
    hd = C // h
    q = F.linear(xw, attn.qkv.to_q.weight, attn.qkv.to_q.bias).reshape(B_, N, h, hd).permute(0, 2, 1, 3)
    kv = F.linear(xw, attn.qkv.to_kv.weight, attn.qkv.to_kv.bias).reshape(B_, N, 2, h, hd).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    s = (q * attn.scale) @ k.transpose(-2, -1)                                           # (B_, h, N, N)
    bias = attn.relative_position_bias_table[attn.relative_position_index.reshape(-1)].view(N, N, h).permute(2, 0, 1)
    s = s + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        s = (s.view(B_ // nW, nW, h, N, N) + mask[None, :, None]).view(B_, h, N, N)
    p = torch.softmax(s, dim=-1)
    o = (p.to(v.dtype) @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(o, attn.proj.weight, attn.proj.bias)


def leff(mlp, x: Tensor) -> Tensor:
    """LeFF.forward (model.py:666-685): Linear+GELU -> (B,4C,H,W) depthwise 3x3 (zero pad on the post-GELU map)+GELU -> Linear."""
    B, L, C = x.shape
    H = int(math.isqrt(L))
    h1 = F.gelu(F.linear(x, mlp.linear1[0].weight, mlp.linear1[0].bias))
    hid = h1.shape[-1]
    m = h1.view(B, H, H, hid).permute(0, 3, 1, 2)
    dw = mlp.dwconv[0]
    h2 = F.gelu(F.conv2d(m, dw.weight, dw.bias, padding=1, groups=hid)).permute(0, 2, 3, 1).reshape(B, L, hid)
    return F.linear(h2, mlp.linear2[0].weight, mlp.linear2[0].bias)


def lewin_block(blk, x: Tensor, mask: Tensor | None = None, scale1: Tensor | None = None, scale2: Tensor | None = None) -> Tensor:
    """LeWinTransformerBlock.forward (model.py:908-989).  scale1/scale2: per-sample stochastic-depth factors
    (B,1,1) in {0, 1/keep} for the two residual branches (timm DropPath, model.py:986-987); None = identity."""
    B, L, C = x.shape
    H = W = int(math.isqrt(L))
    ws, shift = blk.win_size, blk.shift_size
    amask = None if mask is None else input_mask_to_attn_mask(mask.to(x.device), H, W, ws)
    if shift > 0:
        sm = shift_attn_mask(H, W, ws, shift, x.device)
        amask = sm if amask is None else (amask.view(B, -1, ws * ws, ws * ws) + sm[None]).view(-1, ws * ws, ws * ws)
    y = F.layer_norm(x, (C,), blk.norm1.weight, blk.norm1.bias, blk.norm1.eps).view(B, H, W, C)
    if shift > 0:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
    xw = window_partition(y, ws)
    if blk.modulator is not None:
        xw = xw + blk.modulator.weight.to(xw.dtype)
    aw = window_attention(blk.attn, xw, amask)
    y = window_reverse(aw, ws, H, W)
    if shift > 0:
        y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
    a = y.reshape(B, L, C)
    x1 = x + (a if scale1 is None else a * scale1.to(a.dtype))
    f = leff(blk.mlp, F.layer_norm(x1, (C,), blk.norm2.weight, blk.norm2.bias, blk.norm2.eps))
    return x1 + (f if scale2 is None else f * scale2.to(f.dtype))


def downsample(mod, x: Tensor) -> Tensor:
    """Downsample.forward (model.py:739-746)."""
    B, L, C = x.shape
    H = int(math.isqrt(L))
    cv = mod.conv[0]
    o = F.conv2d(x.transpose(1, 2).reshape(B, C, H, H), cv.weight, cv.bias, stride=2, padding=1)
    return o.flatten(2).transpose(1, 2)


def upsample(mod, x: Tensor) -> Tensor:
    """Upsample.forward (model.py:765-771)."""
    B, L, C = x.shape
    H = int(math.isqrt(L))
    dc = mod.deconv[0]
    o = F.conv_transpose2d(x.transpose(1, 2).reshape(B, C, H, H), dc.weight, dc.bias, stride=2)
    return o.flatten(2).transpose(1, 2)


def input_proj(mod, img: Tensor) -> Tensor:
    """InputProj.forward (model.py:800-805): conv3x3 + LeakyReLU(0.01), NCHW -> tokens."""
    cv = mod.proj[0]
    return F.leaky_relu(F.conv2d(img, cv.weight, cv.bias, padding=1), 0.01).flatten(2).transpose(1, 2)


def output_proj(mod, tok: Tensor, residual: Tensor | None = None) -> Tensor:
    """OutputProj.forward (model.py:834-842) + the global residual of Uformer.forward (model.py:1305)."""
    B, L, C = tok.shape
    H = int(math.isqrt(L))
    cv = mod.proj[0]
    y = F.conv2d(tok.transpose(1, 2).reshape(B, C, H, H), cv.weight, cv.bias, padding=1).float()
    return y if residual is None else residual.float() + y


def uformer(net, x: Tensor, mask: Tensor | None = None, scales=None) -> Tensor:
    """Uformer.forward (model.py:1269-1305) over the stage table of network.Uformer.  `scales` (optional): iterator
    yielding the (scale1, scale2) pair of each block in execution order."""
    def stage(st, y):
        for blk in st.blocks:
            s1, s2 = (None, None) if scales is None else next(scales)
            y = lewin_block(blk, y, mask, s1, s2)
        return y
    y = input_proj(net.input_proj, x)
    skips = []
    for i in range(4):
        y = stage(getattr(net, f"encoderlayer_{i}"), y)
        skips.append(y)
        y = downsample(getattr(net, f"dowsample_{i}"), y)
    y = stage(net.conv, y)
    for j in range(4):
        y = torch.cat([upsample(getattr(net, f"upsample_{j}"), y), skips[3 - j]], -1)
        y = stage(getattr(net, f"decoderlayer_{j}"), y)
    return output_proj(net.output_proj, y, x if net.dd_in == 3 else None)
