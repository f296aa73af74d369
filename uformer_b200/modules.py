"""Drop-in nn.Module surface of the reference's LeWin block engine (model.py, L0 in SURVEY §1).

Each class keeps the reference's constructor signature, attributes, ``flops()`` and — crucially —
the exact child-module tree, so ``state_dict()`` keys/shapes are identical (strict checkpoint
loading, utils/model_utils.py:23-33) and ``Uformer._init_weights`` (model.py:1249-1256) still finds
real nn.Linear / nn.LayerNorm children.  Only ``forward`` differs: it packs the parameters once
(cached, invalidated on in-place updates) and calls the native sm_100a kernels.  When autograd is
recording (training, BASELINE configs[2]) the same native forward runs inside
``autograd.NativeFn``, which saves only the op's input and builds the backward by recomputing the op
from it (uformer_b200/restated.py).  There is NO CPU path: calling forward on a non-CUDA tensor
raises ``EngineUnavailable``.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn

from . import _lib, autograd, ops, packing, restated
from ._lib import EngineUnavailable  # noqa: F401  (re-export)


def _to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class DropPath(nn.Module):
    """Stochastic depth (timm semantics, un-vendored dependency of model.py:4,887): identity in eval; in training a
    per-sample Bernoulli(keep)/keep factor on the residual branch.  LeWinTransformerBlock does not call this module
    (its two branches live inside fused kernels); it draws the same factors with `draw` and applies them itself."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        return x * x.new_empty(shape).bernoulli_(keep).div_(keep)

    def draw(self, batch: int, device):
        """(batch,1,1) fp32 factor in {0, 1/keep}, or None when the branch is kept as is."""
        if self.drop_prob == 0.0 or not self.training:
            return None
        keep = 1.0 - self.drop_prob
        return torch.empty((batch, 1, 1), dtype=torch.float32, device=device).bernoulli_(keep).div_(keep)


_WEIGHTS_EPOCH = [0]


def invalidate_packed():
    """Declare every cached operand image stale.  torch bumps a tensor's version on in-place torch ops, which the
    caches already watch; a native optimizer step writes the parameter arena through raw pointers, so
    uformer_b200.training.FlatAdamW calls this after each step."""
    _WEIGHTS_EPOCH[0] += 1


class _PackCache:
    """Packed-parameter cache: recomputed when any source tensor is replaced or modified in place.  The (key, value)
    pair is one attribute read/written atomically: nn.DataParallel replicas (train/train_denoise.py:83) share this
    object across threads and devices, and a replica must never pick up another replica's value."""

    def __init__(self):
        self._entry = (None, None)

    @staticmethod
    def key_for(tensors):
        return (_WEIGHTS_EPOCH[0],) + tuple((t.data_ptr(), t._version, str(t.device)) for t in tensors if t is not None)

    def get(self, tensors, build):
        key = self.key_for(tensors)
        k, v = self._entry
        if k != key:
            with torch.no_grad():
                v = build()
            self._entry = (key, v)
        return v

    def put(self, tensors, value):
        """Install a value built elsewhere (uformer_b200/prepack.py batches the packing of same-shaped modules)."""
        self._entry = (self.key_for(tensors), value)


def _as_bf16(x):
    """Module-boundary dtype policy: compute is bf16 (fp32 accumulate); other float dtypes are cast."""
    if x.dtype == torch.bfloat16:
        return x.contiguous(), None
    if not torch.is_floating_point(x):
        raise TypeError("expected a floating point tensor")
    return x.to(torch.bfloat16).contiguous(), x.dtype


def _as_bf16_rows(x):
    """Like _as_bf16, but a bf16 column slice of a wider buffer (unit channel stride, uniform row stride) passes through
    without a copy: the kernels that accept it take the row stride."""
    if x.dtype == torch.bfloat16 and x.dim() == 3 and x.stride(-1) == 1 and (x.shape[0] == 1 or x.stride(0) == x.shape[1] * x.stride(1)):
        return x, None
    return _as_bf16(x)


def _run(mod, native, restate, acts):
    """Native forward of `mod`; under autograd the same call is wrapped so that backward recomputes it from `acts`."""
    _lib.require_device(acts[0].device)         # device check first: no CPU fallback, fail loudly
    if torch.is_grad_enabled():                 # (inference runs under no_grad: skip the parameter walk entirely)
        params = autograd.trainable_tensors(mod)     # replica-aware (nn.DataParallel empties _parameters)
        if autograd.wants_grad(*acts, *params):
            return autograd.apply(native, restate, acts, params)
    return native(*acts)


# -------------------------------------------------------------------------------------------------
class LinearProjection(nn.Module):
    """Parameter container with the reference's names (model.py:421-447): to_q (C->C), to_kv (C->2C)."""

    def __init__(self, dim, heads=8, dim_head=64, dropout=0., bias=True):
        super().__init__()
        inner_dim = dim_head * heads
        self.heads = heads
        self.to_q = nn.Linear(dim, inner_dim, bias=bias)
        self.to_kv = nn.Linear(dim, inner_dim * 2, bias=bias)
        self.dim = dim
        self.inner_dim = inner_dim

    def flops(self, q_L, kv_L=None):
        kv_L = kv_L or q_L
        return q_L * self.dim * self.inner_dim + kv_L * self.dim * self.inner_dim * 2


class WindowAttention(nn.Module):
    """model.py:452-546.  forward(x (B_, N, C), attn_kv=None, mask=None (nW, N, N)) -> (B_, N, C)."""

    def __init__(self, dim, win_size, num_heads, token_projection='linear', qkv_bias=True, qk_scale=None,
                 attn_drop=0., proj_drop=0.):
        super().__init__()
        if token_projection != 'linear':
            raise NotImplementedError("uformer_b200 implements token_projection='linear' (every get_arch config)")
        self.dim = dim
        self.win_size = win_size
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        n_rel = (2 * win_size[0] - 1) * (2 * win_size[1] - 1)
        self.relative_position_bias_table = nn.Parameter(torch.zeros(n_rel, num_heads))
        t = torch.arange(win_size[0] * win_size[1])
        ty, tx = t // win_size[1], t % win_size[1]
        idx = (ty[:, None] - ty[None, :] + win_size[0] - 1) * (2 * win_size[1] - 1) + (tx[:, None] - tx[None, :] + win_size[1] - 1)
        self.register_buffer("relative_position_index", idx)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.qkv = LinearProjection(dim, num_heads, dim // num_heads, bias=qkv_bias)
        self.token_projection = token_projection
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.softmax = nn.Softmax(dim=-1)
        self._cache = _PackCache()
        self._cache_ln = _PackCache()

    # ---- packing -------------------------------------------------------------------------------
    def pack_sources(self):
        q, kv, pr = self.qkv.to_q, self.qkv.to_kv, self.proj
        return [q.weight, q.bias, kv.weight, kv.bias, pr.weight, pr.bias, self.relative_position_bias_table]

    def packed(self):
        srcs = self.pack_sources()

        def build():
            wq, bq, wkv, bkv, wp, bp, table = srcs
            C = self.dim
            bq_ = bq if bq is not None else torch.zeros(C, device=wq.device)
            bkv_ = bkv if bkv is not None else torch.zeros(2 * C, device=wq.device)
            wimg, bqkv = packing.pack_qkv(wq, bq_, wkv, bkv_, self.num_heads, float(self.scale))
            return dict(wqkv_img=wimg, bqkv=bqkv, wproj_img=packing.pack_kmajor(wp, min(C, 128), "nk"),
                        bproj=bp.float().contiguous(), relpos=packing.pack_relpos(table), head_dim=C // self.num_heads)
        return self._cache.get(srcs, build)

    def tma_gather(self) -> bool:
        """True when the persistent TMA-gather W-MSA kernel (csrc/wmsa_tma.cuh) is built for this shape; the block then also
        packs the LayerNorm-folded projection.  UFORMER_B200_WMSA=classic keeps every launch on wmsa_kernel (A/B switch)."""
        if os.environ.get("UFORMER_B200_WMSA", "tma") == "classic" or tuple(self.win_size) != (8, 8):
            return False
        return bool(_lib.load().lw_wmsa_tma_supported(self.dim, self.dim // self.num_heads))

    def packed_fold(self, norm: nn.LayerNorm, modulator: nn.Embedding | None = None):
        """The projection with `norm` (the block's norm1, model.py:953) folded in, plus the image of the block's window
        modulator pushed through the projection: packing.pack_qkv_fold."""
        q, kv = self.qkv.to_q, self.qkv.to_kv
        srcs = [q.weight, q.bias, kv.weight, kv.bias, norm.weight, norm.bias] + ([modulator.weight] if modulator is not None else [])

        def build():
            C = self.dim
            bq_ = q.bias if q.bias is not None else torch.zeros(C, device=q.weight.device)
            bkv_ = kv.bias if kv.bias is not None else torch.zeros(2 * C, device=q.weight.device)
            r = packing.pack_qkv_fold(q.weight, bq_, kv.weight, bkv_, self.num_heads, float(self.scale), norm.weight, norm.bias,
                                      None if modulator is None else modulator.weight)
            d = dict(wqkv_fold_img=r[0], bqkv_fold=r[1], cs_qkv=r[2])
            if modulator is not None:
                d["wmod_fold_img"] = r[3]
            return d
        return self._cache_ln.get(srcs, build)

    def _check_supported(self):
        if tuple(self.win_size) not in ((8, 8), (16, 16)):
            raise NotImplementedError(f"uformer_b200 kernels are built for 8x8 and 16x16 windows (got {self.win_size})")
        hd = self.dim // self.num_heads
        if tuple(self.win_size) == (16, 16):
            if not _lib.load().lw_wmsa16_supported(self.dim, hd):
                raise NotImplementedError(f"unsupported with 16x16 windows (dim={self.dim}, heads={self.num_heads})")
        elif hd not in (16, 32, 64) or self.dim % hd or self.dim > 512 or (hd in (16, 64) and self.dim > 256):
            raise NotImplementedError(f"unsupported (dim={self.dim}, heads={self.num_heads})")
        if self.attn_drop.p > 0 and self.training or self.proj_drop.p > 0 and self.training:
            raise NotImplementedError("attention/projection dropout is not implemented (p=0 in every Uformer config)")

    def forward(self, x, attn_kv=None, mask=None):
        if attn_kv is not None:
            raise NotImplementedError("cross-attention keys (attn_kv) are never used by Uformer configs")
        self._check_supported()
        _lib.require_device(x.device)
        xb, back = _as_bf16(x)
        acts = [xb] if mask is None else [xb, mask.to(device=x.device, dtype=torch.float32)]
        out = _run(self, lambda t, m=None: ops.wmsa(t, self.packed(), H=0, W=0, shift=0, windowed=True, resid=None, mask=m, win=self.win_size[0]),
                   lambda t, m=None: restated.window_attention(self, t, m), acts)
        return out if back is None else out.to(back)

    def extra_repr(self) -> str:
        return f'dim={self.dim}, win_size={self.win_size}, num_heads={self.num_heads}'

    def flops(self, H, W):
        N = self.win_size[0] * self.win_size[1]
        nW = H * W / N
        hd = self.dim // self.num_heads
        return (self.qkv.flops(H * W, H * W) + nW * self.num_heads * N * hd * N + nW * self.num_heads * N * N * hd
                + nW * N * self.dim * self.dim)


# -------------------------------------------------------------------------------------------------
class LeFF(nn.Module):
    """model.py:654-699.  forward(x (B, HW, C)) -> (B, HW, C)."""

    def __init__(self, dim=32, hidden_dim=128, act_layer=nn.GELU, drop=0., use_eca=False):
        super().__init__()
        if act_layer is not nn.GELU or use_eca:
            raise NotImplementedError("uformer_b200 LeFF: exact GELU, no ECA (as in every Uformer config)")
        self.linear1 = nn.Sequential(nn.Linear(dim, hidden_dim), act_layer())
        self.dwconv = nn.Sequential(nn.Conv2d(hidden_dim, hidden_dim, groups=hidden_dim, kernel_size=3, stride=1, padding=1),
                                    act_layer())
        self.linear2 = nn.Sequential(nn.Linear(hidden_dim, dim))
        self.dim = dim
        self.hidden_dim = hidden_dim
        self.eca = nn.Identity()
        self._cache = _PackCache()
        self._cache_ln = _PackCache()

    def pack_sources(self):
        l1, dw, l2 = self.linear1[0], self.dwconv[0], self.linear2[0]
        return [l1.weight, l1.bias, dw.weight, dw.bias, l2.weight, l2.bias]

    def fused(self) -> bool:
        """True when the single-kernel LeFF (lw_leff_fwd) covers this shape; wider blocks use the two-kernel pair."""
        if os.environ.get("UFORMER_B200_LEFF", "fused") == "split":        # A/B switch for measurements (tools/, bench.py)
            return False
        return bool(_lib.load().lw_leff_fused_supported(self.dim, self.hidden_dim))

    def packed(self, norm: nn.LayerNorm | None = None):
        """Operand dict for ops.leff.  `norm` is the LayerNorm applied in front (the block's norm2, model.py:987): the fused
        kernel folds it into linear1 (packing.pack_leff_fused); the two-kernel path passes its affine through."""
        srcs = self.pack_sources() + ([norm.weight, norm.bias] if norm is not None else [])
        cache = self._cache if norm is None else self._cache_ln

        def build():
            w1, b1, wdw, bdw, w2, b2 = srcs[:6]
            eps = norm.eps if norm is not None else 1e-5
            if self.fused():
                d = packing.pack_leff_fused(w1, b1, None if norm is None else norm.weight, None if norm is None else norm.bias,
                                            wdw, bdw, w2, b2, _lib.load().lw_leff_slice(self.dim))
                d["ln_eps"] = eps
                return d
            d = dict(w1_img=packing.pack_kmajor(w1, _lib.load().lw_nch_ares(self.dim, self.hidden_dim), "nk"), b1=b1.float().contiguous(),
                     taps16=packing.pack_dwconv16(wdw, bdw), w2_img=packing.pack_kmajor(w2, min(self.dim, 128), "kn", torch.float16),
                     b2=b2.float().contiguous(), hidden=self.hidden_dim, ln_eps=eps)
            if norm is not None:
                d.update(ln_w=norm.weight.float().contiguous(), ln_b=norm.bias.float().contiguous())
            return d
        return cache.get(srcs, build)

    def _check_supported(self):
        if self.dim % 16 or self.dim > 512 or self.hidden_dim % 64 or self.dim not in (16, 32, 64, 128, 256, 512):
            raise NotImplementedError(f"unsupported LeFF dims ({self.dim}, {self.hidden_dim})")

    def forward(self, x):
        self._check_supported()
        _lib.require_device(x.device)
        B, L, _ = x.shape
        H = int(math.sqrt(L))
        xb, back = _as_bf16(x)
        out = _run(self, lambda t: ops.leff(t, self.packed(), B=B, H=H, W=H, resid=None), lambda t: restated.leff(self, t), [xb])
        return out if back is None else out.to(back)

    def flops(self, H, W):
        return H * W * self.dim * self.hidden_dim + H * W * self.hidden_dim * 3 * 3 + H * W * self.hidden_dim * self.dim


# -------------------------------------------------------------------------------------------------
class Downsample(nn.Module):
    """model.py:730-753: tokens -> Conv2d(k4,s2,p1) -> tokens."""

    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_channel, out_channel, kernel_size=4, stride=2, padding=1))
        self.in_channel = in_channel
        self.out_channel = out_channel
        self._cache = _PackCache()

    def packed(self):
        cv = self.conv[0]
        return self._cache.get([cv.weight, cv.bias], lambda: dict(
            w_img=packing.pack_downsample(cv.weight, min(self.out_channel, 128)), bias=cv.bias.float().contiguous(),
            cout=self.out_channel))

    def forward(self, x):
        _lib.require_device(x.device)
        B, L, _ = x.shape
        H = int(math.sqrt(L))
        xb, back = _as_bf16_rows(x)
        out = _run(self, lambda t: ops.downsample(t, self.packed(), B=B, H=H, W=H), lambda t: restated.downsample(self, t), [xb])
        return out if back is None else out.to(back)

    def flops(self, H, W):
        return H / 2 * W / 2 * self.in_channel * self.out_channel * 4 * 4


class Upsample(nn.Module):
    """model.py:756-778: tokens -> ConvTranspose2d(k2,s2) -> tokens."""

    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.deconv = nn.Sequential(nn.ConvTranspose2d(in_channel, out_channel, kernel_size=2, stride=2))
        self.in_channel = in_channel
        self.out_channel = out_channel
        self._cache = _PackCache()

    def packed(self):
        dc = self.deconv[0]
        return self._cache.get([dc.weight, dc.bias], lambda: dict(
            w_img=packing.pack_upsample(dc.weight, _lib.load().lw_nch_ares(self.in_channel, 4 * self.out_channel)), bias=dc.bias.float().contiguous(),
            cout=self.out_channel))

    def forward(self, x, out=None):
        _lib.require_device(x.device)
        B, L, _ = x.shape
        H = int(math.sqrt(L))
        xb, back = _as_bf16(x)
        if out is not None and torch.is_grad_enabled() and autograd.wants_grad(xb, *autograd.trainable_tensors(self)):
            raise ValueError("Upsample(out=...) writes in place and cannot be recorded by autograd; call it without `out`")
        res = _run(self, lambda t: ops.upsample(t, self.packed(), B=B, H=H, W=H, out=out), lambda t: restated.upsample(self, t), [xb])
        return res if back is None or out is not None else res.to(back)

    def flops(self, H, W):
        # same expression as the reference (model.py:773-778), which over-counts 4x
        return H * 2 * W * 2 * self.in_channel * self.out_channel * 2 * 2


# -------------------------------------------------------------------------------------------------
class LeWinTransformerBlock(nn.Module):
    """model.py:850-1008.  forward(x (B, HW, C), mask=None) -> (B, HW, C): two kernel launches for C <= 256 — fused W-MSA
    (LN1 .. first residual; 8x8 or 16x16 windows) and fused LeFF (LN2 .. second residual) — three at C = 512, where LeFF is
    split into part 1 (LN2 + Linear1 + GELU) and part 2 (dwconv .. second residual)."""

    def __init__(self, dim, input_resolution, num_heads, win_size=8, shift_size=0, mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 token_projection='linear', token_mlp='leff', modulator=False, cross_modulator=False):
        super().__init__()
        if token_mlp != 'leff' or cross_modulator or norm_layer is not nn.LayerNorm:
            raise NotImplementedError("uformer_b200 block: token_mlp='leff', LayerNorm, no cross-modulator "
                                      "(every get_arch config, utils/model_utils.py:56-81)")
        self.dim = dim
        self.input_resolution = input_resolution
        self.num_heads = num_heads
        self.win_size = win_size
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        self.token_mlp = token_mlp
        if min(self.input_resolution) <= self.win_size:          # construction-time clamp, model.py:863-865
            self.shift_size = 0
            self.win_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.win_size, "shift_size must in 0-win_size"
        self.modulator = nn.Embedding(win_size * win_size, dim) if modulator else None
        self.cross_modulator = None
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, win_size=_to_2tuple(self.win_size), num_heads=num_heads, qkv_bias=qkv_bias,
                                    qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop, token_projection=token_projection)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = LeFF(dim, int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self._cache = _PackCache()
        self.residual_fp32 = default_residual_fp32()      # see set_residual_precision()

    def extra_repr(self) -> str:
        return (f"dim={self.dim}, input_resolution={self.input_resolution}, num_heads={self.num_heads}, "
                f"win_size={self.win_size}, shift_size={self.shift_size}, mlp_ratio={self.mlp_ratio},modulator={self.modulator}")

    def packed(self):
        srcs = [self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias,
                self.modulator.weight if self.modulator is not None else None]

        def build():
            return dict(ln1_w=self.norm1.weight.float().contiguous(), ln1_b=self.norm1.bias.float().contiguous(),
                        ln2_w=self.norm2.weight.float().contiguous(), ln2_b=self.norm2.bias.float().contiguous(),
                        modulator=None if self.modulator is None else self.modulator.weight.float().contiguous())
        return self._cache.get(srcs, build)

    def _attn_operands(self):
        """Operand dict of ops.wmsa for this block: attention images + norm1 affine + modulator, plus the LayerNorm-folded
        projection (and modulator image) when the TMA-gather kernel is built for the block's shape."""
        pk = self.packed()
        pa = dict(self.attn.packed(), ln_w=pk["ln1_w"], ln_b=pk["ln1_b"], modulator=pk["modulator"], ln_eps=self.norm1.eps)
        if self.attn.tma_gather():
            pa.update(self.attn.packed_fold(self.norm1, self.modulator))
        return pa

    @staticmethod
    def input_mask_to_attn_mask(mask, H, W, ws):
        """(B,1,h,w) input mask -> additive (B*nW, N, N) mask, model.py:914-921 (host side, rarely used)."""
        m = torch.nn.functional.interpolate(mask.float(), size=(H, W)).permute(0, 2, 3, 1)
        B = m.shape[0]
        m = m.view(B, H // ws, ws, W // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
        am = m.unsqueeze(2) * m.unsqueeze(1)
        return torch.where(am != 0, torch.full_like(am, -100.0), torch.zeros_like(am))

    def forward(self, x, mask=None, out=None, out_dtype=None, x_b=None, want_b=False):
        """`out` (optional, bf16, same shape as x): write the block output there (used by the stage scheduler).
        `out_dtype` (fp32 residual-stream mode only): dtype of the returned tensor (fp32 inside a stage, bf16 at its end).
        `x_b` / `want_b` (fp32 residual-stream mode, used by the stage scheduler): a bf16 copy of an fp32 `x` — the TMA gather
        source of the W-MSA kernel — and whether to return one of the output as well: the result is then (out, out_bf16)."""
        B, L, C = x.shape
        H = W = int(math.sqrt(L))
        ws = self.win_size
        if H * W != L or ws not in (8, 16) or H % ws:
            raise NotImplementedError(f"uformer_b200 block needs a square token map with side % win_size == 0 and 8x8 or 16x16 "
                                      f"windows (L={L}, win_size={self.win_size})")
        self.attn._check_supported()
        self.mlp._check_supported()
        if mask is not None and self.shift_size > 0 and B > 1:
            # same behaviour as the reference: model.py:942 adds the (B*nW,N,N) input mask to the (nW,N,N) shift mask,
            # which cannot broadcast for B > 1 (its RuntimeError); never silently extend the semantics
            raise RuntimeError(f"input mask with a shifted block needs batch 1 (got {B}): the reference's mask sum at "
                               "model.py:942 does not broadcast (B*nW,N,N) + (nW,N,N)")
        _lib.require_device(x.device)
        stochastic = self.training and isinstance(self.drop_path, DropPath) and self.drop_path.drop_prob > 0.0
        if self.residual_fp32 and not stochastic and not (torch.is_grad_enabled() and autograd.wants_grad(x, *autograd.trainable_tensors(self))):
            return self._forward_fp32_residual(x, B, H, W, mask, out, out_dtype, x_b, want_b)      # inference (nothing to differentiate)
        if want_b:
            raise ValueError("want_b is an fp32 residual-stream option (inference)")
        xb, back = _as_bf16(x)
        # stochastic depth (model.py:986-987): the two per-sample factors, drawn in the reference's order
        dp = self.drop_path if isinstance(self.drop_path, DropPath) else None
        s1 = dp.draw(B, x.device) if dp is not None else None
        s2 = dp.draw(B, x.device) if dp is not None else None
        acts = [xb] + ([s1, s2] if s1 is not None else [])
        if mask is not None:
            acts.append(mask.to(device=x.device, dtype=torch.float32))
        has_dp, has_mask = s1 is not None, mask is not None
        dst = out if back is None else None

        def split(rest):
            sc = rest[:2] if has_dp else (None, None)
            return sc[0], sc[1], (rest[-1] if has_mask else None)

        def native(t, *rest):
            a1, a2, m = split(rest)
            pa = self._attn_operands()
            pm = self.mlp.packed(self.norm2)
            amask = None if m is None else self.input_mask_to_attn_mask(m, H, W, ws)
            if a1 is None:
                x1 = ops.wmsa(t, pa, H=H, W=W, shift=self.shift_size, windowed=False, resid=t, mask=amask, win=ws)
                return ops.leff(x1, pm, B=B, H=H, W=W, resid=x1, out=dst)
            br = ops.wmsa(t, pa, H=H, W=W, shift=self.shift_size, windowed=False, resid=None, mask=amask, win=ws)
            x1 = torch.addcmul(t.float(), br.float(), a1).to(torch.bfloat16)
            br = ops.leff(x1, pm, B=B, H=H, W=W, resid=None)
            y = torch.addcmul(x1.float(), br.float(), a2).to(torch.bfloat16)
            return y if dst is None else dst.copy_(y)

        def restate(t, *rest):
            a1, a2, m = split(rest)
            return restated.lewin_block(self, t, m, a1, a2)

        res = _run(self, native, restate, acts)
        return res if back is None else res.to(back)

    @torch.no_grad()
    def _forward_fp32_residual(self, x, B, H, W, mask=None, out=None, out_dtype=None, x_b=None, want_b=False):
        """Precision mode (set_residual_precision): the residual stream x -> x1 -> out stays fp32 in HBM, inside the kernels:
        W-MSA reads the fp32 stream (LayerNorm in fp32, bf16 GEMM operand), adds its branch in fp32 and writes x1 in fp32 plus
        a bf16 copy (the LeFF kernel's GEMM operand); LeFF adds its branch to the fp32 x1 and writes fp32 (bf16 at the end of
        a stage).  Removes the 80 bf16 roundings of the residual stream that dominate the flagship model's parity error
        (DESIGN §2).  Inference only; `x` may be bf16 (first block of a stage) or fp32."""
        pa = self._attn_operands()
        pm = self.mlp.packed(self.norm2)
        if x.dtype not in (torch.bfloat16, torch.float32):
            x = x.float()
        x = x.contiguous()
        amask = None if mask is None else self.input_mask_to_attn_mask(mask, H, W, self.win_size)
        x1, x1b = ops.wmsa(x, pa, H=H, W=W, shift=self.shift_size, windowed=False, resid=x, mask=amask, out_dtype=torch.float32, bf16_copy=True,
                           x_b=x_b if x.dtype == torch.float32 else None, win=self.win_size)
        odt = out.dtype if out is not None else (out_dtype or torch.float32)
        if want_b and not (self.mlp.fused() and odt == torch.float32):
            raise ValueError("want_b needs the fused LeFF kernel and an fp32 output")
        if out is not None and not self.mlp.fused() and not out.is_contiguous():
            return out.copy_(ops.leff(x1b, pm, B=B, H=H, W=W, resid=x1, out_dtype=odt))
        return ops.leff(x1b, pm, B=B, H=H, W=W, resid=x1, out=out, out_dtype=odt, bf16_copy=want_b)

    def wants_bf16_copy(self) -> bool:
        """Whether this block's W-MSA would gather an fp32 input through its bf16 copy (TMA path; see forward's x_b)."""
        return self.attn.tma_gather() and self.mlp.fused()

    def flops(self):
        H, W = self.input_resolution
        return self.dim * H * W + self.attn.flops(H, W) + self.dim * H * W + self.mlp.flops(H, W)


def residual_mode() -> str:
    """UFORMER_B200_RESIDUAL: "auto" (default), "fp32" or "bf16" — precision of the residual stream between the kernels of a
    stage at inference.  bf16 rounds the residual twice per block (it costs the 40-block flagship ~4e-3 of parity error,
    DESIGN §2); fp32 keeps it in fp32 in HBM (more bytes per token); auto = fp32 in stages of >= 4 blocks (where the roundings
    accumulate and the kernels are compute-bound), bf16 in the 1-2-block full-resolution stages (which are HBM-bound)."""
    m = os.environ.get("UFORMER_B200_RESIDUAL", "auto")
    if m not in ("auto", "fp32", "bf16"):
        raise ValueError(f"UFORMER_B200_RESIDUAL={m!r}: expected auto, fp32 or bf16")
    return m


def default_residual_fp32(depth=None) -> bool:
    """Whether a block constructed now (inside a stage of `depth` blocks; None: standalone) carries an fp32 residual stream."""
    m = residual_mode()
    return m == "fp32" or (m == "auto" and (depth is None or depth >= 4))


def set_residual_precision(net: nn.Module, dtype=torch.float32):
    """Select how LeWin blocks under `net` carry the residual stream between kernels at inference: torch.float32 (default:
    fp32 in HBM inside each stage, read / written by the kernels themselves, see LeWinTransformerBlock._forward_fp32_residual)
    or torch.bfloat16 (faster, noisier).  Returns the number of blocks switched."""
    if dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("residual precision must be torch.float32 or torch.bfloat16")
    n = 0
    for m in net.modules():
        if isinstance(m, LeWinTransformerBlock):
            m.residual_fp32 = dtype == torch.float32
            n += 1
    return n
