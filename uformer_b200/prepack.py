"""Batched re-packing of a whole network's operand images.

A training step invalidates every packed weight image (the optimizer rewrites the parameter arena), and building
them lazily costs ~6 tiny torch launches per weight x ~200 weights = 26 ms per Uformer-B step, twice the native
forward (profiles/r01d_train_probe.json).  Blocks of one stage have identical shapes, so their weights are stacked
and permuted together: one `pack_kmajor` per (stage, weight kind) instead of one per block.  The images are
bit-identical to the lazily built ones (tests/test_host_logic.py) and are installed into the modules' pack caches.
"""
from __future__ import annotations

from collections import defaultdict

import torch

from . import _lib, packing
from .modules import Downsample, LeWinTransformerBlock, Upsample


@torch.no_grad()
def prepack(net: torch.nn.Module) -> int:
    """Rebuild and cache the packed parameters of every engine module under `net`; returns the number of modules."""
    groups = defaultdict(list)
    n = 0
    for m in net.modules():
        if isinstance(m, LeWinTransformerBlock):
            a = m.attn
            key = (m.dim, m.num_heads, m.mlp.hidden_dim, float(a.scale), a.qkv.to_q.bias is not None, a.qkv.to_kv.bias is not None,
                   str(a.proj.weight.device), m.modulator is None)
            groups[key].append(m)
        elif isinstance(m, (Downsample, Upsample)):
            m.packed()                              # four of each, all different shapes: nothing to batch
            n += 1
    for blks in groups.values():
        _pack_blocks(blks)
        n += len(blks)
    return n


def _pack_blocks(blks):
    b0 = blks[0]
    nb, C, heads, hid = len(blks), b0.dim, b0.num_heads, b0.mlp.hidden_dim
    hd = C // heads
    dev = b0.attn.proj.weight.device
    scale = float(b0.attn.scale)

    def stack(get):
        return torch.stack([get(b).detach().float() for b in blks], 0)

    # ---- attention: per-head [q*scale | k | v] row blocks (packing.pack_qkv), proj, bias table ----
    zeros_c, zeros_2c = torch.zeros(C, device=dev), torch.zeros(2 * C, device=dev)
    wq = stack(lambda b: b.attn.qkv.to_q.weight)                                        # (nb, C, C)
    wkv = stack(lambda b: b.attn.qkv.to_kv.weight)                                      # (nb, 2C, C)
    bq = stack(lambda b: b.attn.qkv.to_q.bias if b.attn.qkv.to_q.bias is not None else zeros_c)
    bkv = stack(lambda b: b.attn.qkv.to_kv.bias if b.attn.qkv.to_kv.bias is not None else zeros_2c)
    w3 = torch.cat([wq * scale, wkv], 1)                                                # (nb, 3C, C): q | k | v
    wcat = w3.view(nb, 3, heads, hd, C).permute(0, 2, 1, 3, 4).reshape(nb * heads * 3 * hd, C)
    wqkv_img = packing.pack_kmajor(wcat, 3 * hd, "nk").view(nb, heads, -1, 3 * hd * 64)
    bqkv = torch.cat([bq * scale, bkv], 1).view(nb, 3, heads, hd).permute(0, 2, 1, 3).reshape(nb, -1).contiguous()
    # LayerNorm-folded projection for the TMA-gather W-MSA kernel (packing.pack_qkv_fold), blocks without a modulator
    fold = b0.attn.tma_gather()
    has_mod = b0.modulator is not None                             # (groups are uniform in modulator presence)
    if fold:
        gamma1, beta1 = stack(lambda b: b.norm1.weight), stack(lambda b: b.norm1.bias)
        b3 = torch.cat([bq * scale, bkv], 1)                                            # (nb, 3C)
        bf = b3 + (w3 * beta1[:, None, :]).sum(2)
        wg = (w3 * gamma1[:, None, :]).to(torch.bfloat16)
        cs = wg.float().sum(2)

        def per_head(t):
            return t.view(nb, 3, heads, hd, *t.shape[2:]).transpose(1, 2).reshape(nb, heads * 3 * hd, *t.shape[2:])
        wfold_img = packing.pack_kmajor(per_head(wg).reshape(nb * heads * 3 * hd, C), 3 * hd, "nk").view(nb, heads, -1, 3 * hd * 64)
        bfold, csq = per_head(bf).contiguous(), per_head(cs).contiguous()
        if has_mod:                                 # (m W^T)^T per head, positions in the kernel's window order
            mod = stack(lambda b: b.modulator.weight)                                   # (nb, 64, C)
            mw = per_head(torch.matmul(w3, mod.transpose(1, 2)))[:, :, packing.quarter_major_positions().to(dev)]
            wmod_img = packing.pack_kmajor(mw.reshape(nb * heads * 3 * hd, 64), 3 * hd, "nk").view(nb, heads, -1, 3 * hd * 64)
    nchp = min(C, 128)
    wproj_img = packing.pack_kmajor(stack(lambda b: b.attn.proj.weight).view(nb * C, C), nchp, "nk").view(nb, C // nchp, -1, nchp * 64)
    relpos = stack(lambda b: b.attn.relative_position_bias_table).transpose(1, 2).contiguous()          # (nb, heads, 225)
    # ---- LeFF.  Fused kernel (C <= 128): LayerNorm-folded linear1 in 64-row slices + row sums + folded bias, linear2 k-blocks
    # (packing.pack_leff_fused).  Two-kernel path: linear1 (A-resident, "nk"), linear2 (A-streamed, "kn"). ----
    lib = _lib.load()
    fused = bool(lib.lw_leff_fused_supported(C, hid))
    w1 = stack(lambda b: b.mlp.linear1[0].weight)                                       # (nb, hid, C)
    b1 = stack(lambda b: b.mlp.linear1[0].bias)
    wd = stack(lambda b: b.mlp.dwconv[0].weight).view(nb, hid, 9).transpose(1, 2).contiguous()          # (nb, 9, hid)
    bdw = stack(lambda b: b.mlp.dwconv[0].bias)                                         # (nb, hid)
    w2 = stack(lambda b: b.mlp.linear2[0].weight)                                       # (nb, C, hid)
    if fused:
        gamma, beta = stack(lambda b: b.norm2.weight), stack(lambda b: b.norm2.bias)
        b1f = (b1 + (w1 * beta[:, None, :]).sum(2)).contiguous()
        w1g = (w1 * gamma[:, None, :]).to(torch.bfloat16)
        cs = w1g.float().sum(2).contiguous()
        sl = lib.lw_leff_slice(C)
        w1f_img = packing.pack_kmajor_sw(w1g.view(nb * hid, C), 64, 2 * min(C, 64)).view(nb, hid // 64, -1, 64 * min(C, 64))
        w2f_img = packing.pack_kmajor_sw(w2.view(nb * C, hid), C, 2 * sl, torch.float16).view(nb, 1, hid // sl, C * sl)
        taps = torch.cat([wd, bdw[:, None, :]], 1).view(nb, 10, hid // sl, sl).permute(0, 2, 1, 3).contiguous().to(torch.float16)   # (nb, NS, 10, sl)
    else:
        nch1 = lib.lw_nch_ares(C, hid)
        w1_img = packing.pack_kmajor(w1.view(nb * hid, C), nch1, "nk").view(nb, hid // nch1, -1, nch1 * 64)
        w2_nk = packing.pack_kmajor(w2.view(nb * C, hid), nchp, "nk", torch.float16).view(nb, C // nchp, -1, nchp * 64)
        taps16 = torch.cat([wd, bdw[:, None, :]], 1).to(torch.float16).contiguous()      # (nb, 10, hid)
        w2_img = w2_nk.permute(0, 2, 1, 3).contiguous()                                 # per block [KB][C/nch][nch*64]

    for i, b in enumerate(blks):
        a, m = b.attn, b.mlp
        a._cache.put(a.pack_sources(), dict(wqkv_img=wqkv_img[i], bqkv=bqkv[i], wproj_img=wproj_img[i],
                                            bproj=a.proj.bias.detach().float().contiguous(), relpos=relpos[i], head_dim=hd))
        if fold:
            q, kv = a.qkv.to_q, a.qkv.to_kv
            d = dict(wqkv_fold_img=wfold_img[i], bqkv_fold=bfold[i], cs_qkv=csq[i])
            if has_mod:
                d["wmod_fold_img"] = wmod_img[i]
            a._cache_ln.put([q.weight, q.bias, kv.weight, kv.bias, b.norm1.weight, b.norm1.bias] + ([b.modulator.weight] if has_mod else []), d)
        bd = m.dwconv[0].bias.detach().float().contiguous()
        b2 = m.linear2[0].bias.detach().float().contiguous()
        if fused:
            d = dict(w1f_img=w1f_img[i], b1f=b1f[i], cs=cs[i], taps=taps[i], w2f_img=w2f_img[i], b2=b2, hidden=hid, has_ln=True,
                     slice=sl, ln_eps=b.norm2.eps)
        else:
            d = dict(w1_img=w1_img[i], b1=m.linear1[0].bias.detach().float().contiguous(), taps16=taps16[i], w2_img=w2_img[i], b2=b2,
                     hidden=hid, ln_eps=b.norm2.eps, ln_w=b.norm2.weight.detach().float().contiguous(),
                     ln_b=b.norm2.bias.detach().float().contiguous())
        m._cache_ln.put(m.pack_sources() + [b.norm2.weight, b.norm2.bias], d)
        b.packed()                                  # LN1 affine / modulator: fp32 views of the parameters, no launches
