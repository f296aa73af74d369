"""Run one LeWin block of a given stage shape a few times (target for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import uformer_b200 as U
from paramgen import randomize_state
C, heads, H, modu, B, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
dev = torch.device("cuda:0")
blk = U.LeWinTransformerBlock(C, (H, H), heads, win_size=8, shift_size=4, modulator=bool(modu)).eval()
blk.load_state_dict(randomize_state(blk.state_dict(), 1))
blk = blk.to(dev)
x = torch.randn(B, H * H, C, device=dev).to(torch.bfloat16)
with torch.no_grad():
    for _ in range(reps):
        y = blk(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()          # ncu --profile-from-start off: only this block forward is profiled
    y = blk(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok", float(y.float().abs().mean()))
