"""Explicit backward of one LeWin block, recomputed from the block input — the per-kernel specification of the
native backward (next round) and an opt-in replacement for autograd-over-`restated.lewin_block` today.

`lewin_block_bwd(blk, x, g, ...)` returns d loss/d x and the gradient of every block parameter, given the block input
`x` and the output gradient `g`.  It re-materialises only what backward needs and keeps the attention half in window-major
layout (one gather in, one scatter out).  Dtypes are explicit (no autocast): GEMM operands in `cd` (bf16 on the GPU),
LayerNorm / softmax statistics, bias and weight-gradient accumulators in fp32 — the same split the reference's autocast
training step makes (train/train_denoise.py:178-180).  The numbered steps are the planned kernels:

  recompute   R1 LN1 + roll/partition gather + modulator          (model.py:952-969)
              R2 q,k,v projections, S = q k^T*scale + bias + mask, P = softmax(S), O = P v   (model.py:494-517)
              R3 proj + reverse/unroll + first residual            (model.py:518-520, 975-986)
              R4 LN2, linear1 (+ pre-activation), GELU, depthwise conv (+ pre-activation), GELU   (model.py:666-680, 987)
  backward    B1 linear2: dW2, db2, dh2;  GELU' * ;  depthwise conv: dwd, dbd, dh1;  GELU' *
              B2 linear1: dW1, db1, dz;  LN2 backward;  + g  -> dx1
              B3 gather dx1 to windows;  proj: dWp, dbp, dO
              B4 attention core: dP, dv, softmax backward, bias-table scatter, dq, dk
              B5 q/kv projections: dWq, dWkv, biases, dxw;  modulator sum;  scatter;  LN1 backward;  + dx1 -> dx

Enabled with `uformer_b200.autograd.use_explicit_block_backward(True)` (or UFORMER_B200_EXPLICIT_BWD=1); the default
stays the autograd recompute that was validated on the B200 this round.  tests/test_block_bwd_cpu.py pins every output
against torch autograd through the restated forward (fp32) and runs the bf16 dtype flow on CPU.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import restated

Tensor = torch.Tensor
aten = torch.ops.aten


def _sum_rows(t: Tensor) -> Tensor:
    """fp32 column sums over all leading dims (bias gradients)."""
    return t.reshape(-1, t.shape[-1]).sum(0, dtype=torch.float32)


def _wgrad(dy: Tensor, x: Tensor) -> Tensor:
    """dW (N, K) = dy^T x over all leading dims, fp32 result (the GEMM itself runs in the operands' dtype)."""
    return (dy.reshape(-1, dy.shape[-1]).t() @ x.reshape(-1, x.shape[-1])).float()


def lewin_block_bwd(blk, x: Tensor, g: Tensor, mask: Tensor | None = None, scale1: Tensor | None = None,
                    scale2: Tensor | None = None, cd: torch.dtype = torch.bfloat16, need_dx: bool = True):
    """x, g: (B, L, C).  Returns (dx or None, {parameter name (as in blk.named_parameters()): fp32 gradient})."""
    B, L, C = x.shape
    H = W = int(math.isqrt(L))
    ws, shift = blk.win_size, blk.shift_size
    N = ws * ws
    attn, mlp = blk.attn, blk.mlp
    h = attn.num_heads
    hd = C // h
    f32 = torch.float32
    grads = {}

    def part(t):                                    # (B, L, C') token map -> (B_, N, C') windows of the rolled map
        t = t.view(B, H, W, t.shape[-1])
        if shift > 0:
            t = torch.roll(t, shifts=(-shift, -shift), dims=(1, 2))
        return restated.window_partition(t, ws)

    def unpart(t):                                  # inverse of part
        t = restated.window_reverse(t, ws, H, W)
        if shift > 0:
            t = torch.roll(t, shifts=(shift, shift), dims=(1, 2))
        return t.reshape(B, L, t.shape[-1])

    amask = None if mask is None else restated.input_mask_to_attn_mask(mask.to(x.device), H, W, ws)
    if shift > 0:
        sm = restated.shift_attn_mask(H, W, ws, shift, x.device)
        amask = sm if amask is None else (amask.view(B, -1, N, N) + sm[None]).view(-1, N, N)

    # ---------------- recompute ----------------
    # R1
    g1w, b1w = blk.norm1.weight.float(), blk.norm1.bias.float()
    xf = x.float()
    xn, mu1, rstd1 = aten.native_layer_norm(xf, [C], g1w, b1w, blk.norm1.eps)
    xw = part(xn)
    if blk.modulator is not None:
        xw = xw + blk.modulator.weight.float()
    xw = xw.to(cd)                                                                     # (B_, N, C)
    B_ = xw.shape[0]
    # R2
    wq, wkv, wp = attn.qkv.to_q.weight.to(cd), attn.qkv.to_kv.weight.to(cd), attn.proj.weight.to(cd)
    q = F.linear(xw, wq, None if attn.qkv.to_q.bias is None else attn.qkv.to_q.bias.to(cd))
    kv = F.linear(xw, wkv, None if attn.qkv.to_kv.bias is None else attn.qkv.to_kv.bias.to(cd))
    qs = (q * attn.scale).view(B_, N, h, hd).transpose(1, 2)                           # (B_, h, N, hd), scaled
    k = kv[..., :C].reshape(B_, N, h, hd).transpose(1, 2)
    v = kv[..., C:].reshape(B_, N, h, hd).transpose(1, 2)
    idx = attn.relative_position_index.reshape(-1)
    bias = attn.relative_position_bias_table.float()[idx].view(N, N, h).permute(2, 0, 1)
    s = (qs @ k.transpose(-2, -1)).float() + bias[None]
    if amask is not None:
        nW = amask.shape[0]
        s = (s.view(B_ // nW, nW, h, N, N) + amask[None, :, None]).view(B_, h, N, N)
    p = torch.softmax(s, dim=-1)                                                       # fp32
    pc = p.to(cd)
    o = (pc @ v).transpose(1, 2).reshape(B_, N, C)                                     # (B_, N, C)
    # R3
    a = unpart(F.linear(o, wp, attn.proj.bias.to(cd)))
    x1f = (xf + (a.float() if scale1 is None else a.float() * scale1)).to(cd).float()   # x1 crosses HBM in `cd`
    # R4
    g2w, b2w = blk.norm2.weight.float(), blk.norm2.bias.float()
    z, mu2, rstd2 = aten.native_layer_norm(x1f, [C], g2w, b2w, blk.norm2.eps)
    z = z.to(cd)
    l1, dw, l2 = mlp.linear1[0], mlp.dwconv[0], mlp.linear2[0]
    w1, w2, wd = l1.weight.to(cd), l2.weight.to(cd), dw.weight.to(cd)
    hid = w1.shape[0]
    h1p = F.linear(z, w1, l1.bias.to(cd))                                              # (B, L, 4C) pre-activation
    h1 = F.gelu(h1p)
    h1m = h1.view(B, H, W, hid).permute(0, 3, 1, 2)                                    # NCHW view, channels-last memory
    cpre = F.conv2d(h1m, wd, dw.bias.to(cd), padding=1, groups=hid)                    # pre-activation of the 2nd GELU
    h2 = F.gelu(cpre).permute(0, 2, 3, 1).reshape(B, L, hid)

    # ---------------- backward ----------------
    gc = g.to(cd)
    # B1
    df = gc if scale2 is None else (g.float() * scale2).to(cd)
    grads["mlp.linear2.0.weight"] = _wgrad(df, h2)
    grads["mlp.linear2.0.bias"] = _sum_rows(df)
    dh2 = (df @ w2).view(B, H, W, hid).permute(0, 3, 1, 2)
    dc = aten.gelu_backward(dh2, cpre)
    dh1m, dwd, dbd = aten.convolution_backward(dc, h1m, wd, [hid], [1, 1], [1, 1], [1, 1], False, [0, 0], hid, [True, True, True])
    grads["mlp.dwconv.0.weight"] = dwd.float()
    grads["mlp.dwconv.0.bias"] = dbd.float()
    dh1p = aten.gelu_backward(dh1m.permute(0, 2, 3, 1).reshape(B, L, hid), h1p)
    # B2
    grads["mlp.linear1.0.weight"] = _wgrad(dh1p, z)
    grads["mlp.linear1.0.bias"] = _sum_rows(dh1p)
    dz = (dh1p @ w1).float()
    dx1_ln, dg2, db2 = aten.native_layer_norm_backward(dz, x1f, [C], mu2, rstd2, g2w, b2w, [True, True, True])
    grads["norm2.weight"], grads["norm2.bias"] = dg2, db2
    dx1 = g.float() + dx1_ln                                                           # fp32 (B, L, C)
    # B3
    da = dx1 if scale1 is None else dx1 * scale1
    daw = part(da.to(cd))                                                              # (B_, N, C)
    grads["attn.proj.weight"] = _wgrad(daw, o)
    grads["attn.proj.bias"] = _sum_rows(daw)
    do = (daw @ wp).view(B_, N, h, hd).transpose(1, 2)                                 # (B_, h, N, hd)
    # B4
    dp = (do @ v.transpose(-2, -1)).float()
    dv = pc.transpose(-2, -1) @ do
    ds = aten._softmax_backward_data(dp, p, -1, f32)                                   # fp32 (B_, h, N, N)
    dtab = torch.zeros_like(attn.relative_position_bias_table, dtype=f32)
    dtab.index_add_(0, idx, ds.sum(0).permute(1, 2, 0).reshape(N * N, h))
    grads["attn.relative_position_bias_table"] = dtab
    dsc = ds.to(cd)
    dq = ((dsc @ k) * attn.scale).transpose(1, 2).reshape(B_, N, C)
    dk = (dsc.transpose(-2, -1) @ qs).transpose(1, 2).reshape(B_, N, C)
    dkv = torch.cat([dk, dv.transpose(1, 2).reshape(B_, N, C)], -1)
    # B5
    grads["attn.qkv.to_q.weight"] = _wgrad(dq, xw)
    grads["attn.qkv.to_kv.weight"] = _wgrad(dkv, xw)
    if attn.qkv.to_q.bias is not None:
        grads["attn.qkv.to_q.bias"] = _sum_rows(dq)
    if attn.qkv.to_kv.bias is not None:
        grads["attn.qkv.to_kv.bias"] = _sum_rows(dkv)
    dxw = (dq @ wq + dkv @ wkv).float()                                                # (B_, N, C)
    if blk.modulator is not None:
        grads["modulator.weight"] = dxw.sum(0)
    dxn = unpart(dxw)
    dx_ln, dg1, db1 = aten.native_layer_norm_backward(dxn, xf, [C], mu1, rstd1, g1w, b1w, [need_dx, True, True])
    grads["norm1.weight"], grads["norm1.bias"] = dg1, db1
    dx = (dx1 + dx_ln).to(x.dtype) if need_dx else None
    return dx, grads
