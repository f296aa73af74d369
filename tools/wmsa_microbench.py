"""BASELINE configs[4]: W-MSA microbench — WindowAttention alone over win_size x head_dim x heads, achieved HBM GB/s
and tensor TFLOP/s against the measured peaks (MEASURED_PEAKS.json).

    python tools/wmsa_microbench.py > gpurun_out/wmsa_microbench.json         (on a B200)

Shapes follow how Uformer wires the module: stage i has dim = E*2^i, heads = 2^i, so head_dim = E and
the module sees B_ = 32 * (256 / 2^i / ws)^2 windows of N = ws^2 tokens.  Algorithmic work per call:
FLOPs = 2*B_*N*(4*dim^2 + 2*N*dim), bytes = 2*B_*N*dim*2 (bf16 in + out).  Timing: CUDA events, 256 MB L2 flush between
iterations, median of 20 after 3 warm-ups.  Combinations the kernels do not implement yet are listed as unsupported."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import uformer_b200 as U  # noqa: E402
from paramgen import randomize_state  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm, tf = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops", 1590.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    iters = int(os.environ.get("WMSA_MB_ITERS", 23))         # 1 under ncu (one profiled launch per point)
    for ws in (8, 16):
        for hd in (16, 32, 64):
            for heads in (1, 2, 4, 8):
                dim, N = hd * heads, ws * ws
                side = 256 // heads                                     # heads = 2^i  ->  resolution 256 / 2^i
                B_ = 32 * (side // ws) ** 2
                row = dict(win_size=ws, head_dim=hd, heads=heads, dim=dim, windows=B_)
                att = U.WindowAttention(dim, (ws, ws), heads)
                try:
                    att._check_supported()
                except NotImplementedError as exc:
                    row["unsupported"] = str(exc)
                    rows.append(row)
                    continue
                att.load_state_dict(randomize_state(att.state_dict(), 3))
                att = att.to(dev).eval()
                x = torch.randn(B_, N, dim, device=dev).to(torch.bfloat16)
                times = []
                with torch.no_grad():
                    for it in range(iters):
                        flush.zero_()
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        att(x)
                        b.record()
                        torch.cuda.synchronize()
                        if it >= 3 or iters < 4:           # (3 warm-ups unless the run is cut short for a profiler)
                            times.append(a.elapsed_time(b))
                ms = sorted(times)[len(times) // 2]
                flops = 2.0 * B_ * N * (4 * dim * dim + 2 * N * dim)
                byts = 2.0 * B_ * N * dim * 2
                row.update(ms=round(ms, 4), tflops=round(flops / ms / 1e9, 2), tensor_frac=round(flops / ms / 1e9 / tf, 4),
                           gbs=round(byts / ms / 1e6, 1), hbm_frac=round(byts / ms / 1e6 / hbm, 4))
                rows.append(row)
    print(json.dumps(dict(peaks=dict(hbm_gbs=hbm, bf16_tflops=tf), rows=rows), indent=1))


if __name__ == "__main__":
    main()
