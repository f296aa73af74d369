"""Functional layer: torch tensors in, native sm_100a kernels (C ABI via ctypes) out.

All activations are bf16 CUDA tensors in the reference's token layout (B, H*W, C).  Launches go on
torch's current stream, never synchronise and never allocate inside the library (outputs are
torch.empty'd here), so a whole forward is CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

Tensor = torch.Tensor

# bench.py instrumentation: number of native kernel launches, and (when PROFILE is a list) one
# (label, algorithmic flops, start event, end event) record per launch on the launching stream.
LAUNCH_COUNT = 0
PROFILE = None


def _launch(label: str, flops: float, fn, rc_name: str, device=None):
    """fn(stream) -> return code.  The launch runs with `device` (the device of the op's tensors) current and on THAT
    device's current stream: a tensor on cuda:1 while cuda:0 is current must not launch on cuda:0 with cuda:1 pointers."""
    global LAUNCH_COUNT
    LAUNCH_COUNT += 1
    if device is not None and device.type == "cuda" and device.index is not None and device.index != torch.cuda.current_device():
        with torch.cuda.device(device):
            return _launch(label, flops, fn, rc_name, None)
    stream = torch.cuda.current_stream().cuda_stream
    if PROFILE is None:
        _lib.check(fn(stream), rc_name)
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    _lib.check(fn(stream), rc_name)
    e.record()
    PROFILE.append((label, flops, s, e))


def _ptr(t):
    return None if t is None else t.data_ptr()


def _check_act(x: Tensor, name: str):
    if x.dtype != torch.bfloat16:
        raise TypeError(f"{name} must be bfloat16 (got {x.dtype}); use the module wrappers for dtype conversion")
    if not x.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    _lib.require_device(x.device)


def wmsa(x: Tensor, p: dict, *, H: int, W: int, shift: int, windowed: bool, resid: Tensor | None,
         mask: Tensor | None = None, out: Tensor | None = None, out_dtype=None, bf16_copy: bool = False, x_b: Tensor | None = None,
         win: int = 8):
    """Fused (LN) + (roll/partition) + W-MSA + proj + (reverse/unroll) + (residual).
    p: packed parameter dict from modules._pack_attention (+ optional ln_w/ln_b/modulator).
    fp32 residual-stream mode: x (and resid, which must then be x's dtype) may be fp32; out_dtype=torch.float32 writes the
    sum in fp32; bf16_copy=True additionally returns a bf16 copy of the output (the GEMM operand of the LeFF kernel that
    follows): the return value is then the pair (out, out_bf16).
    When p carries the LayerNorm-folded projection ("wqkv_fold_img", packing.pack_qkv_fold) the library takes the persistent
    TMA-gather kernel (csrc/wmsa_tma.cuh) for token-map inputs; an fp32 x then needs `x_b`, its bf16 copy (the previous
    kernel's bf16_copy output), as the gather source."""
    if x.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError(f"x must be bfloat16 or float32 (got {x.dtype})")
    if not x.is_contiguous():
        raise ValueError("x must be contiguous")
    if resid is not None and resid.dtype != x.dtype:
        raise TypeError("resid must have x's dtype")
    _lib.require_device(x.device)
    Cc = x.shape[-1]
    if windowed:
        n_windows = x.shape[0]
    else:
        n_windows = x.shape[0] * (H // win) * (W // win)
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype or torch.bfloat16, device=x.device)
    out_b = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if bf16_copy else None
    a = _lib.WmsaArgs()
    a.x, a.out, a.resid = _ptr(x), _ptr(out), _ptr(resid)
    a.ln_w, a.ln_b, a.modulator = _ptr(p.get("ln_w")), _ptr(p.get("ln_b")), _ptr(p.get("modulator"))
    a.wqkv_img, a.bqkv, a.wproj_img, a.bproj, a.relpos = (_ptr(p["wqkv_img"]), _ptr(p["bqkv"]), _ptr(p["wproj_img"]),
                                                          _ptr(p["bproj"]), _ptr(p["relpos"]))
    if mask is not None:
        mask = mask.to(device=x.device, dtype=torch.float32).contiguous()
        a.mask, a.n_mask_windows = _ptr(mask), mask.shape[0]
    a.n_windows, a.H, a.W, a.C, a.head_dim = n_windows, H, W, Cc, p["head_dim"]
    a.shift, a.windowed, a.ln_eps = shift, int(windowed), p.get("ln_eps", 1e-5)
    a.x_fp32, a.out_fp32, a.out_b = int(x.dtype == torch.float32), int(out.dtype == torch.float32), _ptr(out_b)
    a.win_size = win
    if "wqkv_fold_img" in p and not windowed and win == 8:
        if x_b is not None and (x_b.dtype != torch.bfloat16 or x_b.shape != x.shape or not x_b.is_contiguous()):
            raise ValueError("x_b must be a contiguous bfloat16 tensor of x's shape")
        a.wqkv_fold_img, a.bqkv_fold, a.cs_qkv, a.x_b = _ptr(p["wqkv_fold_img"]), _ptr(p["bqkv_fold"]), _ptr(p["cs_qkv"]), _ptr(x_b)
        a.wmod_fold_img = _ptr(p.get("wmod_fold_img"))
    ntok = n_windows * win * win
    _launch(f"wmsa_C{Cc}_T{ntok}" if win == 8 else f"wmsa{win}_C{Cc}_T{ntok}", 2.0 * ntok * (4 * Cc * Cc + 2 * win * win * Cc), lambda st: _lib.load().lw_wmsa_fwd(C.byref(a), st), "lw_wmsa_fwd", x.device)
    return (out, out_b) if bf16_copy else out


def _row_stride(t: Tensor, name: str) -> int:
    """Row stride (elements) of a (B, L, C) / (rows, C) activation that may be a column slice of a wider buffer."""
    if t.stride(-1) != 1 or (t.dim() == 3 and t.shape[0] > 1 and t.stride(0) != t.shape[1] * t.stride(1)):
        raise ValueError(f"{name}: rows must be contiguous and uniformly strided")
    return t.stride(-2)


def leff(x: Tensor, p: dict, *, B: int, H: int, W: int, resid: Tensor | None, out: Tensor | None = None, out_dtype=None,
         bf16_copy: bool = False):
    """(LN) + Linear1 + GELU + dwconv3x3 + GELU + Linear2 (+ residual).
    p from packing.pack_leff_fused (key "w1f_img"): ONE launch (lw_leff_fwd), the hidden map stays on chip; x / resid / out
    may be column slices of wider buffers, resid / out may be fp32.  Otherwise (C = 512) the two-kernel path: the hidden
    map makes one bf16 round trip through HBM/L2 between lw_leff1_fwd and lw_leff2_fwd.
    bf16_copy=True (fused kernel only): additionally returns a contiguous bf16 copy of the output — the TMA source of the
    W-MSA kernel that follows when `out` is the fp32 residual stream; the return value is then (out, out_bf16)."""
    if "w1f_img" not in p:
        if bf16_copy:
            raise ValueError("bf16_copy needs the fused LeFF kernel")
        return _leff_two_kernels(x, p, B=B, H=H, W=W, resid=resid, out=out, out_dtype=out_dtype)
    if x.dtype != torch.bfloat16:
        raise TypeError(f"x must be bfloat16 (got {x.dtype})")
    _lib.require_device(x.device)
    Cc, hidden = x.shape[-1], p["hidden"]
    n_tokens = B * H * W
    if out is None:
        out = torch.empty((B, H * W, Cc) if x.dim() == 3 else (n_tokens, Cc), dtype=out_dtype or torch.bfloat16, device=x.device)
    for t, nm in ((resid, "resid"), (out, "out")):
        if t is not None and t.dtype not in (torch.bfloat16, torch.float32):
            raise TypeError(f"{nm} must be bfloat16 or float32")
    a = _lib.LeffArgs()
    a.x, a.out, a.resid = _ptr(x), _ptr(out), _ptr(resid)
    a.w1_img, a.b1f, a.cs, a.taps, a.w2_img, a.b2 = (_ptr(p["w1f_img"]), _ptr(p["b1f"]), _ptr(p["cs"]), _ptr(p["taps"]), _ptr(p["w2f_img"]),
                                                     _ptr(p["b2"]))
    a.B, a.H, a.W, a.C, a.hidden = B, H, W, Cc, hidden
    a.x_stride, a.out_stride = _row_stride(x, "x"), _row_stride(out, "out")
    a.resid_stride = _row_stride(resid, "resid") if resid is not None else 0
    a.resid_fp32 = int(resid is not None and resid.dtype == torch.float32)
    a.out_fp32 = int(out.dtype == torch.float32)
    a.has_ln, a.ln_eps = int(p["has_ln"]), p.get("ln_eps", 1e-5)
    out_b = torch.empty(out.shape, dtype=torch.bfloat16, device=x.device) if bf16_copy else None
    a.out_b = _ptr(out_b)
    lib = _lib.load()
    _launch(f"leff_C{Cc}_T{n_tokens}", 2.0 * n_tokens * (2 * hidden * Cc + 9 * hidden), lambda st: lib.lw_leff_fwd(C.byref(a), st), "lw_leff_fwd", x.device)
    return (out, out_b) if bf16_copy else out


def _leff_two_kernels(x: Tensor, p: dict, *, B: int, H: int, W: int, resid: Tensor | None, out: Tensor | None = None, out_dtype=None) -> Tensor:
    _check_act(x, "x")
    Cc, hidden = x.shape[-1], p["hidden"]
    n_tokens = B * H * W
    h1 = torch.empty((n_tokens, hidden), dtype=torch.float16, device=x.device)      # the hidden map is half precision between the two kernels
    a = _lib.Leff1Args()
    a.x, a.h1, a.ln_w, a.ln_b = _ptr(x), _ptr(h1), _ptr(p.get("ln_w")), _ptr(p.get("ln_b"))
    a.w1_img, a.b1 = _ptr(p["w1_img"]), _ptr(p["b1"])
    a.n_tokens, a.C, a.hidden, a.ln_eps = n_tokens, Cc, hidden, p.get("ln_eps", 1e-5)
    lib = _lib.load()
    _launch(f"leff1_C{Cc}_T{n_tokens}", 2.0 * n_tokens * Cc * hidden, lambda st: lib.lw_leff1_fwd(C.byref(a), st), "lw_leff1_fwd", x.device)
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype or torch.bfloat16, device=x.device)
    if not out.is_contiguous() or (resid is not None and not resid.is_contiguous()):
        raise ValueError("two-kernel LeFF: out / resid must be contiguous")
    b = _lib.Leff2Args()
    b.h1, b.out, b.resid = _ptr(h1), _ptr(out), _ptr(resid)
    b.resid_fp32, b.out_fp32 = int(resid is not None and resid.dtype == torch.float32), int(out.dtype == torch.float32)
    b.taps, b.w2_img, b.b2 = _ptr(p["taps16"]), _ptr(p["w2_img"]), _ptr(p["b2"])
    b.B, b.H, b.W, b.C, b.hidden = B, H, W, Cc, hidden
    _launch(f"leff2_C{Cc}_T{n_tokens}", 2.0 * n_tokens * (hidden * Cc + 9 * hidden), lambda st: lib.lw_leff2_fwd(C.byref(b), st), "lw_leff2_fwd", x.device)
    return out


def downsample(x: Tensor, p: dict, *, B: int, H: int, W: int) -> Tensor:
    """x may be a column slice of a wider (B, HW, S) buffer (the encoder skip living in the skip-concat buffer)."""
    if x.dtype != torch.bfloat16:
        raise TypeError(f"x must be bfloat16 (got {x.dtype})")
    _lib.require_device(x.device)
    Cin, Cout = x.shape[-1], p["cout"]
    out = torch.empty((B, (H // 2) * (W // 2), Cout), dtype=torch.bfloat16, device=x.device)
    a = _lib.DownArgs()
    a.x, a.out, a.w_img, a.bias = _ptr(x), _ptr(out), _ptr(p["w_img"]), _ptr(p["bias"])
    a.B, a.H, a.W, a.Cin, a.Cout, a.x_stride = B, H, W, Cin, Cout, _row_stride(x, "x")
    _launch(f"down_C{Cin}", 2.0 * B * (H // 2) * (W // 2) * 16 * Cin * Cout, lambda st: _lib.load().lw_downsample_fwd(C.byref(a), st),
            "lw_downsample_fwd", x.device)
    return out


def upsample(x: Tensor, p: dict, *, B: int, H: int, W: int, out: Tensor | None = None) -> Tensor:
    """out may be a wider (B, 4HW, S) buffer (skip-concat fusion): the first Cout channels are written."""
    _check_act(x, "x")
    Cin, Cout = x.shape[-1], p["cout"]
    if out is None:
        out = torch.empty((B, 4 * H * W, Cout), dtype=torch.bfloat16, device=x.device)
    a = _lib.UpArgs()
    a.x, a.out, a.w_img, a.bias = _ptr(x), _ptr(out), _ptr(p["w_img"]), _ptr(p["bias"])
    a.B, a.H, a.W, a.Cin, a.Cout, a.out_stride = B, H, W, Cin, Cout, out.shape[-1]
    _launch(f"up_C{Cin}", 2.0 * B * H * W * Cin * 4 * Cout, lambda st: _lib.load().lw_upsample_fwd(C.byref(a), st), "lw_upsample_fwd", x.device)
    return out


def input_proj(img: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """NCHW fp32 image -> bf16 tokens (B, H*W, E): conv3x3 + LeakyReLU(0.01)."""
    _lib.require_device(img.device)
    img = img.float().contiguous()
    B, Cin, H, W = img.shape
    E = w.shape[0]
    tok = torch.empty((B, H * W, E), dtype=torch.bfloat16, device=img.device)
    _launch("input_proj", 2.0 * B * H * W * 9 * Cin * E,
            lambda st: _lib.load().lw_input_proj_fwd(_ptr(img), _ptr(w), _ptr(b), _ptr(tok), B, Cin, H, W, E, st), "lw_input_proj_fwd", img.device)
    return tok


def output_proj(tok: Tensor, w: Tensor, b: Tensor, img: Tensor | None, H: int, W: int) -> Tensor:
    """bf16 tokens (B, H*W, Cin) -> NCHW fp32 conv3x3 output (+ img residual)."""
    _check_act(tok, "tokens")
    B, _, Cin = tok.shape
    Cout = w.shape[0]
    out = torch.empty((B, Cout, H, W), dtype=torch.float32, device=tok.device)
    if img is not None:
        img = img.float().contiguous()
    _launch("output_proj", 2.0 * B * H * W * 9 * Cin * Cout,
            lambda st: _lib.load().lw_output_proj_fwd(_ptr(tok), _ptr(w), _ptr(b), _ptr(img), _ptr(out), B, Cin, H, W, Cout, st),
            "lw_output_proj_fwd", tok.device)
    return out


def charbonnier(x: Tensor, y: Tensor, eps: float, need_grad: bool):
    """(loss (1,) fp32, d loss / d x or None) of mean(sqrt((x-y)^2 + eps^2)) in one pass (two tiny launches)."""
    _lib.require_device(x.device)
    if x.dtype != torch.float32 or y.dtype != torch.float32 or x.shape != y.shape or not x.is_contiguous() or not y.is_contiguous():
        raise TypeError("charbonnier expects two contiguous fp32 tensors of one shape")
    grad = torch.empty_like(x) if need_grad else None
    loss = torch.empty(1, dtype=torch.float32, device=x.device)
    partial = torch.empty(_lib.CHARBONNIER_PARTIALS, dtype=torch.float32, device=x.device)
    n = x.numel()
    _launch("charbonnier", 8.0 * n, lambda st: _lib.load().lw_charbonnier_fwd_bwd(_ptr(x), _ptr(y), _ptr(grad), _ptr(loss), _ptr(partial), n, eps,
                                                                             st), "lw_charbonnier_fwd_bwd", x.device)
    return loss, grad


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, *, step: int, lr: float, beta1: float, beta2: float, eps: float,
               weight_decay: float, grad_scale: float, zero_grad: bool):
    """One AdamW update over flat fp32 arenas, in place (p, m, v; g zeroed when zero_grad)."""
    _lib.require_device(p.device)
    for t in (p, g, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != p.numel():
            raise TypeError("adamw_step expects four contiguous fp32 tensors of one size")
    a = _lib.AdamWArgs()
    a.p, a.g, a.m, a.v, a.n = _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel()
    a.step, a.zero_grad = step, int(zero_grad)
    a.lr, a.beta1, a.beta2, a.eps, a.weight_decay, a.grad_scale = lr, beta1, beta2, eps, weight_decay, grad_scale
    _launch("adamw", 12.0 * p.numel(), lambda st: _lib.load().lw_adamw_step(C.byref(a), st), "lw_adamw_step", p.device)
