"""GPU probe of the single-kernel LeFF (lw_leff_fwd): every shape in its own subprocess (a trap / sticky CUDA error in one
shape must not hide the others), parity against the CPU contract model + oracle, and timing against the two-kernel path.
    python tools/leff_fused_probe.py            # all shapes
    python tools/leff_fused_probe.py C H B      # one shape (child mode)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SHAPES = [(32, 16, 1), (32, 8, 3), (64, 24, 1), (128, 16, 2), (16, 40, 1), (256, 16, 2), (256, 8, 1), (64, 64, 2), (128, 64, 8), (256, 64, 8), (32, 256, 4),
          (64, 256, 2), (128, 128, 8), (256, 64, 32)]


def child(C, H, B):
    import torch
    import uformer_b200 as U
    from uformer_b200 import ops
    from oracle import lewin_oracle as O
    from paramgen import randomize_state
    torch.manual_seed(C + H)
    dev = "cuda:0"
    blk = U.LeWinTransformerBlock(C, (max(H, 16),) * 2, max(1, C // 32), win_size=8, shift_size=0).eval()
    os.environ.pop("UFORMER_B200_LEFF", None)
    st = randomize_state(blk.state_dict(), 9)
    blk.load_state_dict(st)
    blk = blk.to(dev)
    x = torch.randn(B, H * H, C).to(torch.bfloat16)
    xd = x.to(dev)
    res = dict(C=C, H=H, B=B)
    pm = blk.mlp.packed(blk.norm2)
    assert "w1f_img" in pm
    with torch.no_grad():
        y = ops.leff(xd, pm, B=B, H=H, W=H, resid=xd)
        torch.cuda.synchronize()
        if B * H * H <= 70000:
            z = O.layer_norm(x.float(), st["norm2.weight"], st["norm2.bias"])
            ref = x.float() + O.leff(z, st, "mlp.")
            e = ((y.float().cpu() - ref).norm() / ref.norm()).item()
            res["rel_l2_vs_oracle"] = e
            res["branch_rel_l2"] = ((y.float().cpu() - x.float() - (ref - x.float())).norm() / (ref - x.float()).norm()).item()
        # fp32 residual / fp32 out / strided out
        wide = torch.zeros(B, H * H, 2 * C, dtype=torch.bfloat16, device=dev)
        ops.leff(xd, pm, B=B, H=H, W=H, resid=xd, out=wide[:, :, C:])
        res["strided_out_equal"] = bool(torch.equal(wide[:, :, C:], y)) and bool((wide[:, :, :C] == 0).all())
        wide[:, :, :C] = xd
        y2 = ops.leff(wide[:, :, :C], pm, B=B, H=H, W=H, resid=wide[:, :, :C])
        res["strided_in_equal"] = bool(torch.equal(y2, y))
        y32 = ops.leff(xd, pm, B=B, H=H, W=H, resid=xd.float(), out_dtype=torch.float32)
        res["fp32_maxdiff_vs_bf16"] = (y32 - y.float()).abs().max().item()
        # timing: fused vs split
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
        res["fused_us"] = timeit(lambda: ops.leff(xd, pm, B=B, H=H, W=H, resid=xd))
        os.environ["UFORMER_B200_LEFF"] = "split"
        U.modules.invalidate_packed()
        ps = blk.mlp.packed(blk.norm2)
        assert "w1_img" in ps
        ys = ops.leff(xd, ps, B=B, H=H, W=H, resid=xd)
        res["split_us"] = timeit(lambda: ops.leff(xd, ps, B=B, H=H, W=H, resid=xd))
        res["fused_vs_split_rel_l2"] = ((y.float() - ys.float()).norm() / ys.float().norm()).item()
    print("RESULT " + json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) == 4:
        child(*map(int, sys.argv[1:]))
    else:
        for C, H, B in SHAPES:
            try:
                out = subprocess.run([sys.executable, __file__, str(C), str(H), str(B)], capture_output=True, text=True, timeout=120)
                lines = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT ")]
                print(lines[0] if lines else f"FAIL C={C} H={H} B={B} rc={out.returncode}: {out.stderr[-600:]}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"TIMEOUT C={C} H={H} B={B}", flush=True)
