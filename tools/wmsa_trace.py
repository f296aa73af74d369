import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import uformer_b200 as U
from paramgen import randomize_state
dev = torch.device("cuda:0")
C, heads, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
blk = U.LeWinTransformerBlock(C, (H, H), heads, win_size=8, shift_size=4, modulator=True).eval()
blk.load_state_dict(randomize_state(blk.state_dict(), 1)); blk = blk.to(dev)
x = torch.randn(32, H * H, C, device=dev).to(torch.bfloat16)
buf = torch.zeros(2048, dtype=torch.int64, device=dev)
with torch.no_grad():
    blk(x)
torch.cuda.synchronize()
os.environ["LW_TRACE_PTR"] = str(buf.data_ptr()); os.environ["LW_DEBUG"] = "16"
from uformer_b200 import ops
pk = blk.packed()
pa = dict(blk.attn.packed(), ln_w=pk["ln1_w"], ln_b=pk["ln1_b"], modulator=pk["modulator"], ln_eps=1e-5)
with torch.no_grad():
    ops.wmsa(x, pa, H=H, W=H, shift=4, windowed=False, resid=x)
torch.cuda.synchronize()
t = buf.cpu().tolist(); n = t.index(-1); t = t[:n]; t0 = t[0]; d = [v - t0 for v in t]
print("after LN staging/arrive:", d[0])
k = 1
for h in range(min(4, heads)):
    print(f"head {h}: qkv_full {d[k]}, qkv staged {d[k+1]}, s_full {d[k+2]}, p_ready {d[k+3]}, o_full {d[k+4]}, O done {d[k+5]}")
    k += 6
print("all heads done:", d[k], " proj chunks done:", d[k + 1:])
