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


def pack_kmajor(w: Tensor, nch: int, order: str = "nk") -> Tensor:
    """(N, K) -> bf16 images.  order "nk": [N/nch][KB][nch*64] (A-resident kernels iterate N outer),
    order "kn": [KB][N/nch][nch*64] (A-streamed kernels iterate K outer)."""
    N, K = w.shape
    assert N % nch == 0, (N, nch)
    KB = (K + 63) // 64
    wp = w.to(torch.bfloat16)                         # round first: the permutation then moves half the bytes
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


def pack_relpos(table: Tensor) -> Tensor:
    """relative_position_bias_table ((2ws-1)^2, heads) -> (heads, (2ws-1)^2) fp32 contiguous."""
    return table.float().t().contiguous()


def pack_dwconv(w: Tensor, b: Tensor):
    """Conv2d(groups=hidden) weight (hidden,1,3,3) -> taps (9, hidden) fp32; bias fp32."""
    hid = w.shape[0]
    return w.float().reshape(hid, 9).t().contiguous(), b.float().contiguous()


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
