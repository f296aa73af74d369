"""Arbitrary-resolution inference (BASELINE configs[3]): the host logic of the reference's test scripts
(test/test_sidd.py:79-108, same in test_dnd.py / test_gopro_hide.py / test_realblur.py) around the engine's forward.

The LeWin kernels are resolution-agnostic (any square token map whose side is a multiple of 8 at every stage, i.e. an
image side that is a multiple of 128 for the 4-level U), so a model built with img_size=256 runs 512x512 unchanged;
other sizes are zero-padded to the next square multiple of `factor` and the valid region is cut out of the result —
exactly what the reference does (its `mask` is only used for that crop, never fed to the model)."""
from __future__ import annotations

import math

import torch

Tensor = torch.Tensor


def expand2square(timg: Tensor, factor: float = 16.0):
    """test/test_sidd.py:79-92: centre `timg` (B,C,h,w) in a zero (B,C,X,X) canvas, X = ceil(max(h,w)/factor)*factor;
    also returns the {0,1} mask (B,1,X,X) of the valid region.  (The reference handles B=1, C=3; any B, C here.)"""
    B, C, h, w = timg.shape
    X = int(math.ceil(max(h, w) / float(factor)) * factor)
    img = torch.zeros(B, C, X, X, dtype=timg.dtype, device=timg.device)
    mask = torch.zeros(B, 1, X, X, dtype=timg.dtype, device=timg.device)
    y0, x0 = (X - h) // 2, (X - w) // 2
    img[:, :, y0:y0 + h, x0:x0 + w] = timg
    mask[:, :, y0:y0 + h, x0:x0 + w] = 1
    return img, mask


def restore_image(net, noisy: Tensor, factor: int = 128, clamp: bool = True) -> Tensor:
    """test/test_sidd.py:101-108: pad to a square multiple of `factor`, run the network, cut the valid region out,
    clamp to [0,1].  `net` is the engine's Uformer (or the reference's after install()) on a B200."""
    B, C, h, w = noisy.shape
    padded, _ = expand2square(noisy, factor)
    X = padded.shape[-1]
    with torch.no_grad():
        restored = net(padded)
    y0, x0 = (X - h) // 2, (X - w) // 2
    out = restored[:, :, y0:y0 + h, x0:x0 + w]
    return out.clamp(0, 1) if clamp else out
