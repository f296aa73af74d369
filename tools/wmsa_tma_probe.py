"""GPU probe of the persistent TMA-gather W-MSA kernel (csrc/wmsa_tma.cuh): every shape in its own subprocess, parity against
the classic kernel (same launch API, UFORMER_B200_WMSA=classic), the CPU contract model and the oracle block; fp32
residual-stream and explicit-mask variants; time against the classic kernel.
    python tools/wmsa_tma_probe.py                       # all shapes
    python tools/wmsa_tma_probe.py C heads H B shift     # one shape (child mode)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MODULATOR = os.environ.get("PROBE_MODULATOR", "0") == "1"          # probe the modulated (decoder-type) blocks instead
SHAPES = [(32, 1, 16, 1, 0), (32, 1, 16, 2, 4), (64, 2, 24, 1, 4), (128, 4, 16, 2, 4), (256, 8, 16, 1, 4), (256, 8, 16, 2, 0), (16, 1, 8, 1, 0),
          (16, 1, 24, 1, 4), (64, 4, 16, 1, 4), (256, 16, 16, 1, 4), (32, 1, 8, 3, 0),
          (32, 1, 256, 32, 4), (64, 2, 128, 32, 4), (64, 2, 256, 32, 4), (128, 4, 64, 32, 4), (128, 4, 128, 32, 4), (256, 8, 32, 32, 4), (256, 8, 64, 32, 4)]


def child(C, heads, H, B, shift):
    import torch
    import uformer_b200 as U
    from uformer_b200 import ops
    import kernel_model as KM
    from oracle import lewin_oracle as O
    from paramgen import randomize_state
    torch.manual_seed(C + H + shift)
    dev = "cuda:0"
    blk = U.LeWinTransformerBlock(C, (max(H, 16),) * 2, heads, win_size=8, shift_size=shift, modulator=MODULATOR).eval()
    st = randomize_state(blk.state_dict(), 9)
    blk.load_state_dict(st)
    x = torch.randn(B, H * H, C).to(torch.bfloat16)
    res = dict(C=C, heads=heads, H=H, B=B, shift=shift, modulator=MODULATOR)
    small = B * H * H <= 20000
    if small:
        os.environ["UFORMER_B200_WMSA"] = "tma"
        pa_cpu = blk._attn_operands()
        assert "wqkv_fold_img" in pa_cpu
        y_model = KM.wmsa(x, pa_cpu, H=H, W=H, shift=shift, windowed=False, resid=x)
        ref_blk = O.lewin_block(x.float(), st, "", heads, 8, shift)
    blk = blk.to(dev)
    xd = x.to(dev)

    def rel(a, b):
        return ((a.float() - b.float()).norm() / b.float().norm()).item()

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n * 1e3

    with torch.no_grad():
        os.environ["UFORMER_B200_WMSA"] = "classic"
        pc = blk._attn_operands()
        assert "wqkv_fold_img" not in pc
        y_c = ops.wmsa(xd, pc, H=H, W=H, shift=shift, windowed=False, resid=xd)
        torch.cuda.synchronize()
        os.environ["UFORMER_B200_WMSA"] = "tma"
        pt = blk._attn_operands()
        assert "wqkv_fold_img" in pt
        y_t = ops.wmsa(xd, pt, H=H, W=H, shift=shift, windowed=False, resid=xd)
        torch.cuda.synchronize()
        res["finite"] = bool(torch.isfinite(y_t.float()).all())
        br_c = y_c.float() - xd.float()
        res["tma_vs_classic_branch_rel_l2"] = ((y_t.float() - y_c.float()).norm() / br_c.norm()).item()
        res["tma_vs_classic_rel_l2"] = rel(y_t, y_c)
        if small:
            res["tma_vs_model_branch_rel_l2"] = ((y_t.float().cpu() - y_model.float()).norm() / (y_model.float() - x.float()).norm()).item()
            res["tma_vs_model_maxabs"] = (y_t.float().cpu() - y_model.float()).abs().max().item()
            yb = blk(xd)
            res["block_rel_l2_vs_oracle"] = rel(yb.cpu(), ref_blk)
            res["block_rel_max_vs_oracle"] = ((yb.float().cpu() - ref_blk).abs().max() / ref_blk.abs().max()).item()
        # fp32 residual stream: fp32 x + its bf16 copy as the gather source
        x32 = xd.float()
        y32, y32b = ops.wmsa(x32, pt, H=H, W=H, shift=shift, windowed=False, resid=x32, out_dtype=torch.float32, bf16_copy=True, x_b=xd)
        torch.cuda.synchronize()
        res["fp32_maxdiff_vs_bf16"] = (y32 - y_t.float()).abs().max().item()
        res["fp32_copy_equal"] = bool(torch.equal(y32b, y32.to(torch.bfloat16)))
        # explicit additive mask (the input-mask path): allowed with shift only for batch 1
        if shift == 0 or B == 1:
            nw = B * (H // 8) ** 2
            m = torch.where(torch.rand(nw, 64, 64, device=dev) < 0.2, -100.0, 0.0)
            ym_c = ops.wmsa(xd, pc, H=H, W=H, shift=shift, windowed=False, resid=xd, mask=m)
            ym_t = ops.wmsa(xd, pt, H=H, W=H, shift=shift, windowed=False, resid=xd, mask=m)
            torch.cuda.synchronize()
            res["mask_tma_vs_classic_branch_rel_l2"] = ((ym_t.float() - ym_c.float()).norm() / (ym_c.float() - xd.float()).norm()).item()
        res["tma_us"] = timeit(lambda: ops.wmsa(xd, pt, H=H, W=H, shift=shift, windowed=False, resid=xd))
        res["classic_us"] = timeit(lambda: ops.wmsa(xd, pc, H=H, W=H, shift=shift, windowed=False, resid=xd))
        res["tma_fp32_us"] = timeit(lambda: ops.wmsa(x32, pt, H=H, W=H, shift=shift, windowed=False, resid=x32, out_dtype=torch.float32, bf16_copy=True, x_b=xd))
        res["classic_fp32_us"] = timeit(lambda: ops.wmsa(x32, pc, H=H, W=H, shift=shift, windowed=False, resid=x32, out_dtype=torch.float32, bf16_copy=True))
        gb = 4.0 * B * H * H * C / 1e9
        res["tma_GBps"] = gb / (res["tma_us"] * 1e-6)
    print("RESULT " + json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) == 6:
        child(*map(int, sys.argv[1:]))
    else:
        big = len(sys.argv) == 2 and sys.argv[1] == "big"          # timing shapes only (A/B of build variants via UFORMER_B200_LIB)
        for sh in (SHAPES[11:] if big else SHAPES):
            try:
                out = subprocess.run([sys.executable, __file__, *map(str, sh)], capture_output=True, text=True, timeout=150)
                lines = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT ")]
                print(lines[0] if lines else f"FAIL {sh} rc={out.returncode}: {out.stderr[-800:]}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"TIMEOUT {sh}", flush=True)
