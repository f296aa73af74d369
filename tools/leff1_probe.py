import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import uformer_b200 as U
from uformer_b200 import ops
from paramgen import randomize_state
dev = torch.device("cuda:0")
for C, H in [(256, 64), (512, 32), (128, 128), (64, 256)]:
    m = U.LeFF(C, 4 * C).eval(); m.load_state_dict(randomize_state(m.state_dict(), 1)); m = m.to(dev)
    x = torch.randn(32, H * H, C, device=dev).to(torch.bfloat16)
    ops.PROFILE = []
    with torch.no_grad():
        for _ in range(3): m(x)
    torch.cuda.synchronize()
    rec, ops.PROFILE = ops.PROFILE, None
    print(os.environ.get("LW_DEBUG", "0"), " ".join(f"{l}={s.elapsed_time(e):.3f}ms" for l, f, s, e in rec[-2:]), flush=True)
