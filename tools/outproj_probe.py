"""OutputProj kernel probe at the bench shape (Uformer-B 256x256, batch 32: 2.1 M pixels x 64 channels -> 3): parity against
torch's fp32 conv2d of the same bf16 tokens and CUDA-event timing with a 256 MB L2 flush between launches.
    python tools/outproj_probe.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from uformer_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False          # the reference conv must be real fp32
torch.backends.cuda.matmul.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H = W = 256
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for cin in (64, 32):
    tok = torch.randn(B, H * W, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(3, cin, 3, 3, device=dev) * 0.05
    b = torch.randn(3, device=dev)
    img = torch.rand(B, 3, H, W, device=dev)
    y = ops.output_proj(tok, w, b, img, H, W)
    ref = F.conv2d(tok[:4].float().transpose(1, 2).reshape(4, cin, H, W), w, b, padding=1) + img[:4]
    err = (y[:4] - ref).abs().max().item() / ref.abs().max().item()
    times = []
    for it in range(13):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.output_proj(tok, w, b, img, H, W)
        e.record()
        torch.cuda.synchronize()
        if it >= 3:
            times.append(s.elapsed_time(e))
    ms = sorted(times)[len(times) // 2]
    byts = tok.numel() * 2 + 2 * img.numel() * 4
    print(f"output_proj Cin={cin} B={B}: max-abs/max-abs {err:.2e}  {ms * 1e3:.1f} us  {byts / ms / 1e6:.0f} GB/s algorithmic")
