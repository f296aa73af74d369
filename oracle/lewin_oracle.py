"""CPU restatement of the reference's LeWin-block hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``uformer_b200/`` imports it and the product path has no CPU fallback.

Every function restates (does not copy) the arithmetic of one reference function, cited as
``model.py:<lines>`` relative to ``/root/reference``.  The restatement is functional (plain tensors
and a state-dict of weights, no nn.Module), uses closed forms where the reference builds index
tensors procedurally (relative-position index, shift mask, window gather), and works in whatever
dtype the caller passes (fp32 for parity, fp64 to pin the restatement itself).

Parity pinning: the reference repo ships no golden vectors or known-answer tests for this path
(SURVEY.md §8c).  The oracle is therefore pinned against outputs of the reference itself,
generated in the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference/model.py`` unmodified behind a 3-symbol timm shim) and committed under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every fixture, and
``tests/test_oracle_vs_reference.py`` re-checks live whenever ``/root/reference`` is present.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# FAST = True swaps the explicit gather / nine-tap / per-tap statements below for the equivalent library
# calls (view+permute, grouped conv2d, conv2d, conv_transpose2d) — the same ATen ops the reference's
# own CPU forward dispatches to.  Used ONLY by bench.py's CPU-baseline legs so the timed CPU port is not
# handicapped by the didactic formulation; tests/test_oracle_golden.py checks FAST == explicit.
FAST = False


# --------------------------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------------------------
def layer_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm(dim) as used at model.py:881,888 (eps 1e-5, affine, biased variance)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * weight + bias


def gelu_erf(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (model.py:658,660)."""
    if FAST:
        return F.gelu(x)
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def relative_position_index(ws: int) -> Tensor:
    """Closed form of the buffer built at model.py:467-477:
    idx[i, j] = (yi - yj + ws-1) * (2ws-1) + (xi - xj + ws-1), token i = (yi, xi) row-major."""
    t = torch.arange(ws * ws)
    y, x = t // ws, t % ws
    return (y[:, None] - y[None, :] + ws - 1) * (2 * ws - 1) + (x[:, None] - x[None, :] + ws - 1)


def relative_position_bias(table: Tensor, ws: int) -> Tensor:
    """bias[h, i, j] = table[idx[i, j], h]   (model.py:500-503). table: ((2ws-1)^2, heads)."""
    idx = relative_position_index(ws)
    return table[idx.reshape(-1)].reshape(ws * ws, ws * ws, -1).permute(2, 0, 1)


def shift_attn_mask(H: int, W: int, ws: int, shift: int, dtype=torch.float32) -> Tensor:
    """Closed form of the SW-MSA mask built at model.py:924-942.

    In rolled coordinates a token at (r, c) belongs to region 3*reg(r, H) + reg(c, W) with
    reg(p, L) = (p >= L - ws) + (p >= L - shift); two tokens of one window may attend to each
    other iff their regions agree, otherwise the logit gets -100 (finite, not -inf).
    Returns (nW, ws*ws, ws*ws)."""
    r = torch.arange(H)
    c = torch.arange(W)
    reg_r = (r >= H - ws).long() + (r >= H - shift).long()
    reg_c = (c >= W - ws).long() + (c >= W - shift).long()
    region = 3 * reg_r[:, None] + reg_c[None, :]                        # (H, W)
    rw = window_partition(region[None, :, :, None].to(dtype), ws).reshape(-1, ws * ws)
    diff = rw[:, None, :] - rw[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def window_partition(x: Tensor, ws: int) -> Tensor:
    """(B, H, W, C) -> (B*nW, ws, ws, C); window order row-major over (wy, wx), batch-major
    (model.py:704-715, dilation branch unused).  Stated as an explicit gather."""
    B, H, W, C = x.shape
    nwy, nwx = H // ws, W // ws
    if FAST:
        return x.reshape(B, nwy, ws, nwx, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B * nwy * nwx, ws, ws, C)
    wy = torch.arange(nwy)[:, None, None, None]
    wx = torch.arange(nwx)[None, :, None, None]
    iy = torch.arange(ws)[None, None, :, None]
    ix = torch.arange(ws)[None, None, None, :]
    rows = (wy * ws + iy).expand(nwy, nwx, ws, ws)
    cols = (wx * ws + ix).expand(nwy, nwx, ws, ws)
    out = x[:, rows, cols, :]                                           # (B, nwy, nwx, ws, ws, C)
    return out.reshape(B * nwy * nwx, ws, ws, C)


def window_reverse(win: Tensor, ws: int, H: int, W: int) -> Tensor:
    """Inverse of window_partition (model.py:717-726)."""
    nwy, nwx = H // ws, W // ws
    B = win.shape[0] // (nwy * nwx)
    C = win.shape[-1]
    if FAST:
        return win.reshape(B, nwy, nwx, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    out = win.new_zeros(B, H, W, C)
    v = win.reshape(B, nwy, nwx, ws, ws, C)
    for wy in range(nwy):
        for wx in range(nwx):
            out[:, wy * ws:(wy + 1) * ws, wx * ws:(wx + 1) * ws, :] = v[:, wy, wx]
    return out


# --------------------------------------------------------------------------------------------
# W-MSA
# --------------------------------------------------------------------------------------------
def window_attention(x: Tensor, p: Dict[str, Tensor], prefix: str, heads: int, ws: int,
                     mask: Optional[Tensor] = None) -> Tensor:
    """WindowAttention.forward (model.py:494-522) with LinearProjection (model.py:431-442).

    x: (B_, N, C) windows.  q/k/v head h = channels [h*hd, (h+1)*hd) of to_q / to_kv[:C] /
    to_kv[C:].  S = (q * hd^-0.5) k^T + bias (+ mask[w % nW]); softmax over keys; O = P v;
    out = O Wp^T + bp."""
    B_, N, C = x.shape
    hd = C // heads
    wq, bq = p[prefix + "qkv.to_q.weight"], p[prefix + "qkv.to_q.bias"]
    wkv, bkv = p[prefix + "qkv.to_kv.weight"], p[prefix + "qkv.to_kv.bias"]
    q = x @ wq.t() + bq
    kv = x @ wkv.t() + bkv
    k, v = kv[..., :C], kv[..., C:]

    def split(t):
        return t.reshape(B_, N, heads, hd).permute(0, 2, 1, 3)           # (B_, h, N, hd)

    q, k, v = split(q) * (hd ** -0.5), split(k), split(v)
    s = q @ k.transpose(-2, -1)                                          # (B_, h, N, N)
    s = s + relative_position_bias(p[prefix + "relative_position_bias_table"], ws).unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        s = (s.reshape(B_ // nW, nW, heads, N, N) + mask[None, :, None]).reshape(B_, heads, N, N)
    pm = torch.softmax(s, dim=-1)
    o = (pm @ v).permute(0, 2, 1, 3).reshape(B_, N, C)
    return o @ p[prefix + "proj.weight"].t() + p[prefix + "proj.bias"]


# --------------------------------------------------------------------------------------------
# LeFF
# --------------------------------------------------------------------------------------------
def leff(x: Tensor, p: Dict[str, Tensor], prefix: str) -> Tensor:
    """LeFF.forward (model.py:666-685): h1 = GELU(x W1^T + b1); h2 = GELU(dwconv3x3(h1) + bd) with
    zero padding applied to h1 (post-GELU), out = h2 W2^T + b2.  The depthwise conv is stated as
    nine shifted multiply-adds on the (B, H, W, 4C) token map."""
    B, L, C = x.shape
    H = int(math.isqrt(L))
    h1 = gelu_erf(x @ p[prefix + "linear1.0.weight"].t() + p[prefix + "linear1.0.bias"])
    hid = h1.shape[-1]
    if FAST:
        m = h1.reshape(B, H, H, hid).permute(0, 3, 1, 2)
        c = F.conv2d(m, p[prefix + "dwconv.0.weight"], p[prefix + "dwconv.0.bias"], padding=1, groups=hid)
        h2 = F.gelu(c).permute(0, 2, 3, 1).reshape(B, L, hid)
        return h2 @ p[prefix + "linear2.0.weight"].t() + p[prefix + "linear2.0.bias"]
    m = h1.reshape(B, H, H, hid)
    mp = F.pad(m, (0, 0, 1, 1, 1, 1))                                   # zero-pad W and H by 1
    wd = p[prefix + "dwconv.0.weight"].reshape(hid, 3, 3)
    acc = torch.zeros_like(m)
    for ky in range(3):
        for kx in range(3):
            acc = acc + mp[:, ky:ky + H, kx:kx + H, :] * wd[:, ky, kx]
    h2 = gelu_erf(acc + p[prefix + "dwconv.0.bias"]).reshape(B, L, hid)
    return h2 @ p[prefix + "linear2.0.weight"].t() + p[prefix + "linear2.0.bias"]


# --------------------------------------------------------------------------------------------
# Down / Up sampling
# --------------------------------------------------------------------------------------------
def downsample(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """Downsample.forward (model.py:739-746): Conv2d(k4, s2, p1) on the token map.
    out[b, (y, x), co] = bias[co] + sum_{ky,kx,ci} in[b, 2y-1+ky, 2x-1+kx, ci] * W[co, ci, ky, kx]."""
    B, L, C = x.shape
    H = int(math.isqrt(L))
    if FAST:
        o = F.conv2d(x.reshape(B, H, H, C).permute(0, 3, 1, 2), weight, bias, stride=2, padding=1)
        return o.flatten(2).transpose(1, 2)
    m = F.pad(x.reshape(B, H, H, C), (0, 0, 1, 1, 1, 1))
    Ho = H // 2
    out = x.new_zeros(B, Ho, Ho, weight.shape[0])
    for ky in range(4):
        for kx in range(4):
            tap = m[:, ky:ky + 2 * Ho:2, kx:kx + 2 * Ho:2, :]            # (B, Ho, Ho, Cin)
            out = out + tap @ weight[:, :, ky, kx].t()
    return (out + bias).reshape(B, Ho * Ho, -1)


def upsample(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """Upsample.forward (model.py:765-771): ConvTranspose2d(k2, s2); every output pixel has one tap:
    out[b, 2y+dy, 2x+dx, co] = bias[co] + sum_ci in[b, y, x, ci] * W[ci, co, dy, dx]."""
    B, L, C = x.shape
    H = int(math.isqrt(L))
    Cout = weight.shape[1]
    if FAST:
        o = F.conv_transpose2d(x.reshape(B, H, H, C).permute(0, 3, 1, 2), weight, bias, stride=2)
        return o.flatten(2).transpose(1, 2)
    m = x.reshape(B, H, H, C)
    out = x.new_zeros(B, 2 * H, 2 * H, Cout)
    for dy in range(2):
        for dx in range(2):
            out[:, dy::2, dx::2, :] = m @ weight[:, :, dy, dx] + bias
    return out.reshape(B, 4 * L, Cout)


# --------------------------------------------------------------------------------------------
# LeWin block
# --------------------------------------------------------------------------------------------
def lewin_block(x: Tensor, p: Dict[str, Tensor], prefix: str, heads: int, ws: int, shift: int,
                input_mask: Optional[Tensor] = None, drop_scales=None) -> Tensor:
    """LeWinTransformerBlock.forward (model.py:908-989).  Eval mode by default (DropPath = identity).

    x1 = x + DP1(reverse(WMSA(partition(roll(LN1(x), -s)) + modulator), +s)); out = x1 + DP2(LeFF(LN2(x1))).
    ``input_mask`` is the optional (B,1,h,w) mask of model.py:914-921 (nearest-resized).
    ``drop_scales`` = (s1, s2), two (B,1,1) tensors in {0, 1/keep}: the per-sample factors timm's DropPath
    (model.py:887, applied at :986-987) multiplies the two branches with in training mode — the caller draws them
    (``bernoulli_(keep).div_(keep)``, first branch first) so that oracle and engine see the same stochastic depth."""
    B, L, C = x.shape
    H = W = int(math.isqrt(L))
    mask = None
    if input_mask is not None:
        im = F.interpolate(input_mask, size=(H, W)).permute(0, 2, 3, 1)
        iw = window_partition(im, ws).reshape(-1, ws * ws)
        am = iw[:, :, None] * iw[:, None, :]
        mask = torch.where(am != 0, torch.full_like(am, -100.0), torch.zeros_like(am))
    if shift > 0:
        sm = shift_attn_mask(H, W, ws, shift, x.dtype)
        mask = sm if mask is None else mask + sm
    y = layer_norm(x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"]).reshape(B, H, W, C)
    if shift > 0:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
    win = window_partition(y, ws).reshape(-1, ws * ws, C)
    if (prefix + "modulator.weight") in p:
        win = win + p[prefix + "modulator.weight"]
    a = window_attention(win, p, prefix + "attn.", heads, ws, mask)
    y = window_reverse(a.reshape(-1, ws, ws, C), ws, H, W)
    if shift > 0:
        y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
    y = y.reshape(B, L, C)
    if drop_scales is not None:
        y = y * drop_scales[0].to(y.dtype)
    x1 = x + y
    z = leff(layer_norm(x1, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"]), p, prefix + "mlp.")
    if drop_scales is not None:
        z = z * drop_scales[1].to(z.dtype)
    return x1 + z


# --------------------------------------------------------------------------------------------
# whole network (caller side: model.py:1013-1066 BasicUformerLayer, :1069-1305 Uformer)
# --------------------------------------------------------------------------------------------
def stage_plan(img_size: int, embed_dim: int, depths: Sequence[int], num_heads: Sequence[int],
               win_size: int = 8, shift_flag: bool = True) -> List[dict]:
    """Static description of the 9 stages: name, dim, resolution, heads, per-block shift and
    window size, including the construction-time clamp of model.py:863-865 (which uses the
    *constructor* img_size, not the runtime map size)."""
    E = embed_dim
    names = ["encoderlayer_0", "encoderlayer_1", "encoderlayer_2", "encoderlayer_3", "conv",
             "decoderlayer_0", "decoderlayer_1", "decoderlayer_2", "decoderlayer_3"]
    dims = [E, 2 * E, 4 * E, 8 * E, 16 * E, 16 * E, 8 * E, 4 * E, 2 * E]
    res = [img_size // (2 ** i) for i in (0, 1, 2, 3, 4, 3, 2, 1, 0)]
    plan = []
    for i, n in enumerate(names):
        blocks = []
        for b in range(depths[i]):
            shift = 0 if (b % 2 == 0 or not shift_flag) else win_size // 2
            ws = win_size
            if res[i] <= win_size:
                shift, ws = 0, res[i]
            blocks.append(dict(shift=shift, ws=ws))
        plan.append(dict(name=n, dim=dims[i], res=res[i], heads=num_heads[i], blocks=blocks))
    return plan


def uformer_forward(x: Tensor, p: Dict[str, Tensor], img_size: int, embed_dim: int,
                    depths: Sequence[int], num_heads: Sequence[int] = (1, 2, 4, 8, 16, 16, 8, 4, 2),
                    win_size: int = 8, shift_flag: bool = True, dd_in: int = 3,
                    input_mask: Optional[Tensor] = None, taps: Optional[dict] = None) -> Tensor:
    """Uformer.forward (model.py:1269-1305) in eval mode on an NCHW image batch."""
    plan = stage_plan(img_size, embed_dim, depths, num_heads, win_size, shift_flag)
    B, _, H, W = x.shape
    # InputProj (model.py:781-812): conv3x3 + LeakyReLU(0.01), NCHW -> tokens
    y = F.leaky_relu(F.conv2d(x, p["input_proj.proj.0.weight"], p["input_proj.proj.0.bias"], padding=1), 0.01)
    y = y.flatten(2).transpose(1, 2)

    def run_stage(t, st):
        for bi, blk in enumerate(st["blocks"]):
            t = lewin_block(t, p, f"{st['name']}.blocks.{bi}.", st["heads"], blk["ws"], blk["shift"], input_mask)
        if taps is not None:
            taps[st["name"]] = t
        return t

    skips = []
    for i in range(4):
        y = run_stage(y, plan[i])
        skips.append(y)
        y = downsample(y, p[f"dowsample_{i}.conv.0.weight"], p[f"dowsample_{i}.conv.0.bias"])
    y = run_stage(y, plan[4])
    for j in range(4):
        up = upsample(y, p[f"upsample_{j}.deconv.0.weight"], p[f"upsample_{j}.deconv.0.bias"])
        y = torch.cat([up, skips[3 - j]], dim=-1)
        y = run_stage(y, plan[5 + j])
    # OutputProj (model.py:815-846) + global residual (model.py:1305)
    Bt, L, C = y.shape
    Hs = int(math.isqrt(L))
    m = y.transpose(1, 2).reshape(Bt, C, Hs, Hs)
    out = F.conv2d(m, p["output_proj.proj.0.weight"], p["output_proj.proj.0.bias"], padding=1)
    return x + out if dd_in == 3 else out
