"""Per-phase clock64 timeline of CTA 0 (worker thread 0) of the TMA-gather W-MSA kernel, first three tiles.
Needs a -DLW_TRACE build:  UFORMER_B200_LIB=uformer_b200/lib/liblewin_b200_trace.so python tools/wmsa_tma_trace.py C heads H B
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import uformer_b200 as U
from uformer_b200 import ops
from paramgen import randomize_state

C, heads, H, B = map(int, sys.argv[1:5])
dev = torch.device("cuda:0")
blk = U.LeWinTransformerBlock(C, (H, H), heads, win_size=8, shift_size=4).eval()
blk.load_state_dict(randomize_state(blk.state_dict(), 1))
blk = blk.to(dev)
x = torch.randn(B, H * H, C, device=dev).to(torch.bfloat16)
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
pa = blk._attn_operands()
assert "wqkv_fold_img" in pa
with torch.no_grad():
    ops.wmsa(x, pa, H=H, W=H, shift=4, windowed=False, resid=x)
    torch.cuda.synchronize()
    os.environ["LW_TRACE_PTR"] = str(buf.data_ptr())
    os.environ["LW_DEBUG"] = "16"
    ops.wmsa(x, pa, H=H, W=H, shift=4, windowed=False, resid=x)
    torch.cuda.synchronize()
t = buf.cpu().tolist()
t = t[:t.index(-1)]
d = [v - t[0] for v in t]
nh = min(2, heads)
nc = max(1, C // 128)
per_tile = 2 + 5 * nh + 1 + nc + 1
print(f"C={C} heads={heads} H={H} B={B}: cycles since the first event; {len(d)} events, {per_tile} per tile")
for it in range(len(d) // per_tile):
    e = d[it * per_tile:(it + 1) * per_tile]
    print(f"tile {it}: start {e[0]}  stats done {e[1]}")
    k = 2
    for h in range(nh):
        print(f"   head {h}: qkv_full {e[k]}  staged {e[k+1]}  s_full {e[k+2]}  p_ready {e[k+3]}  o_full {e[k+4]}")
        k += 5
    print(f"   heads done {e[k]}  d_full {e[k+1:k+1+nc]}  tile end {e[k+1+nc]}")
