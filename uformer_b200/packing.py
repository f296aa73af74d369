"""Host-side packing of module parameters into the operand images the sm_100a kernels stream.

A GEMM weight W (N, K) (``nn.Linear`` layout, K contiguous = "K-major") is cut into chunks of
``nch`` output rows x 64 input channels.  Each chunk is stored as the exact byte image of a UMMA
K-major SWIZZLE_128B shared-memory tile (rows of 128 B, the 16-byte column index XOR-ed with
``row % 8``), so the kernel's producer warp can land it with a single linear ``cp.async.bulk`` and
the tensor core can consume it without any further shuffling.  K is zero-padded to a multiple of 64.

These functions are pure index permutations (plus the bf16 cast and, for q, the attention scale);
they run on whatever device the parameters live on and are unit-tested on CPU.
"""
from __future__ import annotations

import torch

Tensor = torch.Tensor


_PERMS = {}


def _swizzle128_perm(nch: int, device) -> Tensor:
    """perm[n, k] = element offset inside a (nch x 64) bf16 SWIZZLE_128B tile image (cached per (nch, device): a
    training step re-packs ~200 weights after every optimizer update)."""
    key = (nch, str(device))
    perm = _PERMS.get(key)
    if perm is None:
        n = torch.arange(nch)[:, None]
        k = torch.arange(64)[None, :]
        chunk = (k // 8) ^ (n % 8)
        perm = (n * 64 + chunk * 8 + (k % 8)).to(device)
        _PERMS[key] = perm
    return perm


def pack_kmajor(w: Tensor, nch: int, order: str = "nk", dtype=torch.bfloat16) -> Tensor:
    """(N, K) -> bf16 images.  order "nk": [N/nch][KB][nch*64] (A-resident kernels iterate N outer),
    order "kn": [KB][N/nch][nch*64] (A-streamed kernels iterate K outer)."""
    N, K = w.shape
    assert N % nch == 0, (N, nch)
    KB = (K + 63) // 64
    wp = w.to(dtype)                                  # round first: the permutation then moves half the bytes (bf16, or fp16 for
                                                      # the LeFF linear2 operand, which meets a half-precision hidden map)
    if KB * 64 != K:
        wp = torch.nn.functional.pad(wp, (0, KB * 64 - K))
    perm = _swizzle128_perm(nch, w.device).reshape(-1)
    t = wp.reshape(N // nch, nch, KB, 64).permute(0, 2, 1, 3).reshape(N // nch, KB, nch * 64)
    img = torch.empty_like(t)
    img[:, :, perm] = t
    if order == "kn":
        img = img.permute(1, 0, 2)
    return img.contiguous()


def unpack_kmajor(img: Tensor, N: int, K: int, nch: int, order: str = "nk") -> Tensor:
    """Inverse of pack_kmajor (testing aid)."""
    KB = (K + 63) // 64
    if order == "kn":
        img = img.permute(1, 0, 2)
    perm = _swizzle128_perm(nch, img.device).reshape(-1)
    t = img.float()[:, :, perm].reshape(N // nch, KB, nch, 64).permute(0, 2, 1, 3).reshape(N, KB * 64)
    return t[:, :K]


def _swizzle_perm(nch: int, sw: int, device) -> Tensor:
    """perm[n, k] = element offset inside a (nch rows x sw bytes) bf16 K-major tile image with SWIZZLE_<sw>B (sw in 32/64/128):
    the 16-byte chunk index of byte offset `lin = n*sw + 2k` is XOR-ed with bits [7..] of lin (csrc/umma.cuh swz<SW>)."""
    key = (nch, sw, str(device))
    perm = _PERMS.get(key)
    if perm is None:
        n = torch.arange(nch)[:, None]
        k = torch.arange(sw // 2)[None, :]
        lin = n * sw + 2 * k
        lin = lin ^ (((lin >> 7) & (sw // 16 - 1)) << 4)
        perm = (lin // 2).to(device)
        _PERMS[key] = perm
    return perm


def pack_kmajor_sw(w: Tensor, nch: int, sw: int, dtype=torch.bfloat16) -> Tensor:
    """(N, K) -> 16-bit images [N/nch][KB][nch * sw/2] with rows of `sw` bytes (sw/2 channels per k-block, K % (sw/2) == 0);
    dtype bfloat16 (default) or float16 (operands of a kind::f16 UMMA whose other operand is produced in half precision)."""
    N, K = w.shape
    cb = sw // 2
    assert N % nch == 0 and K % cb == 0, (N, K, nch, sw)
    KB = K // cb
    t = w.to(dtype).reshape(N // nch, nch, KB, cb).permute(0, 2, 1, 3).reshape(N // nch, KB, nch * cb)
    img = torch.empty_like(t)
    img[:, :, _swizzle_perm(nch, sw, w.device).reshape(-1)] = t
    return img.contiguous()


def unpack_kmajor_sw(img: Tensor, N: int, K: int, nch: int, sw: int) -> Tensor:
    """Inverse of pack_kmajor_sw (testing aid)."""
    cb = sw // 2
    KB = K // cb
    t = img.float()[:, :, _swizzle_perm(nch, sw, img.device).reshape(-1)]
    return t.reshape(N // nch, KB, nch, cb).permute(0, 2, 1, 3).reshape(N, K)


def pack_leff_taps(wdw: Tensor, bdw: Tensor, sl: int) -> Tensor:
    """Depthwise Conv2d(groups=hidden) weight (hidden,1,3,3) + bias -> [hidden/sl][10][sl] fp16: per hidden slice the 9 taps
    (tap = ky*3+kx) then the bias, one contiguous bulk-copy chunk per slice.  fp16: the fused kernel runs the depthwise
    conv in packed half precision (the hidden map only exists on chip, in fp16 — 3 more mantissa bits than the bf16 map
    the two-kernel path stores)."""
    hid = wdw.shape[0]
    t = torch.cat([wdw.float().reshape(hid, 9).t(), bdw.float()[None, :]], 0)            # (10, hidden)
    return t.reshape(10, hid // sl, sl).permute(1, 0, 2).contiguous().to(torch.float16)


def pack_leff_fused(w1: Tensor, b1: Tensor, ln_w, ln_b, wdw: Tensor, bdw: Tensor, w2: Tensor, b2: Tensor, sl: int = 64) -> dict:
    """Operands of the single-kernel LeFF (csrc/leff_fused.cuh, lw_leff_fwd); `sl` = lw_leff_slice(C) hidden channels per slice
    (taps and linear2 are cut with it; linear1 is always cut in 64-row units: at sl = 32 GEMM-1 runs on pairs of slices).
    LayerNorm (norm2, model.py:987) is folded into linear1 (model.py:671):
      LN(x) W1^T + b1 = rstd*(x W1g^T) - rstd*mean*cs + b1f  with  W1g = W1 diag(gamma) rounded to bf16,
    cs = row sums of that bf16 matrix (so the mean term cancels exactly against what the tensor core accumulates) and
    b1f = b1 + W1 beta.  Without LayerNorm (LeFF standalone) W1g = W1, b1f = b1."""
    hid, C = w1.shape
    w1f, b1f = w1.float(), b1.float()
    if ln_w is not None:
        b1f = b1f + (w1f * ln_b.float()[None, :]).sum(1)       # same reduction as prepack.py (bit-identical images)
        w1f = w1f * ln_w.float()[None, :]
    w1g = w1f.to(torch.bfloat16)
    return dict(w1f_img=pack_kmajor_sw(w1g, 64, 2 * min(C, 64)), b1f=b1f.contiguous(), cs=w1g.float().sum(1).contiguous(),
                taps=pack_leff_taps(wdw, bdw, sl), w2f_img=pack_kmajor_sw(w2, C, 2 * sl, torch.float16), b2=b2.float().contiguous(), hidden=hid,
                has_ln=ln_w is not None, slice=sl)


def pack_qkv(wq: Tensor, bq: Tensor, wkv: Tensor, bkv: Tensor, heads: int, scale: float):
    """LinearProjection weights (model.py:421-447) -> per-head [q_h; k_h; v_h] row blocks.
    The attention scale (model.py:497, q * hd^-0.5) is folded into the q rows and q bias."""
    C = wq.shape[0]
    hd = C // heads
    w3 = torch.cat([wq.float() * scale, wkv.float()], 0)                          # (3C, C): q (scaled) | k | v
    wcat = w3.view(3, heads, hd, C).permute(1, 0, 2, 3).reshape(heads * 3 * hd, C)
    b3 = torch.cat([bq.float() * scale, bkv.float()], 0)
    bias = b3.view(3, heads, hd).permute(1, 0, 2).reshape(-1)
    img = pack_kmajor(wcat, 3 * hd, "nk")                       # [heads][KB][3hd*64]
    return img, bias.contiguous()                               # bias (heads*3*hd,)


def quarter_major_positions() -> Tensor:
    """Row order of a window inside the TMA-gather W-MSA kernel (csrc/wmsa_tma.cuh): row k of a window holds the natural
    (row-major 8x8) position nat[k]; the four 4x4 quarters follow each other, each row-major."""
    k = torch.arange(64)
    y = ((k >> 5) & 1) * 4 + ((k >> 2) & 3)
    x = ((k >> 4) & 1) * 4 + (k & 3)
    return y * 8 + x


def pack_qkv_fold(wq: Tensor, bq: Tensor, wkv: Tensor, bkv: Tensor, heads: int, scale: float, ln_w: Tensor, ln_b: Tensor,
                  modulator: Tensor | None = None):
    """pack_qkv with the LayerNorm in front (norm1, model.py:953) folded into the projection, for the TMA-gather W-MSA kernel
    (csrc/wmsa_tma.cuh), whose A operand is the raw token tile:
      LN(x) W^T + b = rstd*(x Wg^T) - rstd*mean*cs + bf,  Wg = W diag(gamma) rounded to bf16, cs = row sums of that bf16
    matrix (the mean term then cancels exactly against what the tensor core accumulates), bf = b + W beta.
    Returns (image, bf, cs), rows in pack_qkv's per-head [q_h; k_h; v_h] order, q rows pre-scaled.
    With a window modulator (model.py:966-969, `LN(x) + m[pos]` before the projection) a fourth value: the image of
    (m W^T)^T per head, [heads][3hd rows][64 positions in the kernel's quarter-major order] — the B operand of the one-hot
    k-block that adds the per-position term on the tensor core."""
    C = wq.shape[0]
    hd = C // heads
    w3 = torch.cat([wq.float() * scale, wkv.float()], 0)                          # (3C, C)
    b3 = torch.cat([bq.float() * scale, bkv.float()], 0)
    bf = b3 + (w3 * ln_b.float()[None, :]).sum(1)              # same reductions as prepack.py (bit-identical images)
    wg = (w3 * ln_w.float()[None, :]).to(torch.bfloat16)
    cs = wg.float().sum(1)

    def per_head(t):
        return t.view(3, heads, hd, *t.shape[1:]).transpose(0, 1).reshape(heads * 3 * hd, *t.shape[1:])
    out = (pack_kmajor(per_head(wg), 3 * hd, "nk"), per_head(bf).contiguous(), per_head(cs).contiguous())
    if modulator is None:
        return out
    mw = w3 @ modulator.float().t()                                               # (3C, 64): row n, natural position
    mw = per_head(mw)[:, quarter_major_positions().to(mw.device)]                 # columns in the kernel's window order
    return out + (pack_kmajor(mw, 3 * hd, "nk"),)


def pack_relpos(table: Tensor) -> Tensor:
    """relative_position_bias_table ((2ws-1)^2, heads) -> (heads, (2ws-1)^2) fp32 contiguous."""
    return table.float().t().contiguous()


def pack_dwconv(w: Tensor, b: Tensor):
    """Conv2d(groups=hidden) weight (hidden,1,3,3) -> taps (9, hidden) fp32; bias fp32."""
    hid = w.shape[0]
    return w.float().reshape(hid, 9).t().contiguous(), b.float().contiguous()


def pack_dwconv16(w: Tensor, b: Tensor) -> Tensor:
    """Two-kernel LeFF (lw_leff2_fwd): (10, hidden) fp16 — the 9 taps (tap = ky*3+kx), then the bias row."""
    wd, bd = pack_dwconv(w, b)
    return torch.cat([wd, bd[None, :]], 0).to(torch.float16).contiguous()


def pack_downsample(w: Tensor, nch: int) -> Tensor:
    """Conv2d(Cin->Cout,k4,s2,p1) weight (Cout,Cin,4,4) -> GEMM weight (Cout, K=tap*Cin+ci)."""
    Cout, Cin = w.shape[:2]
    wk = w.float().permute(0, 2, 3, 1).reshape(Cout, 16 * Cin)
    return pack_kmajor(wk, nch, "kn")


def pack_upsample(w: Tensor, nch: int) -> Tensor:
    """ConvTranspose2d(Cin->Cout,k2,s2) weight (Cin,Cout,2,2) -> GEMM weight (N=(dy*2+dx)*Cout+co, Cin)."""
    Cin, Cout = w.shape[:2]
    wn = w.float().permute(2, 3, 1, 0).reshape(4 * Cout, Cin)
    return pack_kmajor(wn, nch, "nk")
