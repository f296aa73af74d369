"""Run every Uformer-B stage shape (batch 32, 256x256 input) one block at a time with timing; log progressively."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import uformer_b200 as U
from uformer_b200 import ops
from paramgen import randomize_state

log = open(os.path.join(ROOT, "gpurun_out", "stage_probe.log"), "w")
def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True); log.write(s + "\n"); log.flush(); os.fsync(log.fileno())

B = int(os.environ.get("PB", 32))
dev = torch.device("cuda:0")
stages = [("enc0", 32, 1, 256, False), ("enc1", 64, 2, 128, False), ("enc2", 128, 4, 64, False), ("enc3", 256, 8, 32, False),
          ("bott", 512, 16, 16, False), ("dec0", 512, 16, 32, True), ("dec1", 256, 8, 64, True), ("dec2", 128, 4, 128, True), ("dec3", 64, 2, 256, True)]
P("start", time.strftime("%X"))
for name, C, h, H, modu in stages:
    for shift in (0, 4):
        blk = U.LeWinTransformerBlock(C, (H, H), h, win_size=8, shift_size=shift, modulator=modu).eval()
        blk.load_state_dict(randomize_state(blk.state_dict(), 1))
        blk = blk.to(dev)
        x = torch.randn(B, H * H, C, device=dev).to(torch.bfloat16)
        P(name, "C", C, "H", H, "shift", shift, "launch...")
        for rep in range(2):
            ops.PROFILE = []
            with torch.no_grad():
                y = blk(x)
            torch.cuda.synchronize()
            rec, ops.PROFILE = ops.PROFILE, None
        msg = " ".join(f"{l}={s.elapsed_time(e):.3f}ms({f / s.elapsed_time(e) / 1e9:.0f}TF)" for l, f, s, e in rec)
        P("   ", msg, "finite", bool(torch.isfinite(y.float()).all()))
        del blk, x, y
# down / up
for i, (cin, H) in enumerate([(32, 256), (64, 128), (128, 64), (256, 32)]):
    m = U.Downsample(cin, 2 * cin).to(dev).eval()
    x = torch.randn(B, H * H, cin, device=dev).to(torch.bfloat16)
    ops.PROFILE = []
    with torch.no_grad():
        m(x); m(x)
    torch.cuda.synchronize()
    rec, ops.PROFILE = ops.PROFILE, None
    P("down", cin, H, f"{rec[-1][2].elapsed_time(rec[-1][3]):.3f}ms")
for cin, cout, H in [(512, 256, 16), (512, 128, 32), (256, 64, 64), (128, 32, 128)]:
    m = U.Upsample(cin, cout).to(dev).eval()
    x = torch.randn(B, H * H, cin, device=dev).to(torch.bfloat16)
    ops.PROFILE = []
    with torch.no_grad():
        m(x); m(x)
    torch.cuda.synchronize()
    rec, ops.PROFILE = ops.PROFILE, None
    P("up", cin, cout, H, f"{rec[-1][2].elapsed_time(rec[-1][3]):.3f}ms")
P("done", time.strftime("%X"))
