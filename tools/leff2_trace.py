import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import uformer_b200 as U
from uformer_b200 import ops
from paramgen import randomize_state
dev = torch.device("cuda:0")
C, H = int(sys.argv[1]), int(sys.argv[2])
os.environ['UFORMER_B200_LEFF'] = 'split'          # this tool traces the two-kernel path (C = 512, or forced)
m = U.LeFF(C, 4 * C).eval(); m.load_state_dict(randomize_state(m.state_dict(), 1)); m = m.to(dev)
x = torch.randn(32, H * H, C, device=dev).to(torch.bfloat16)
buf = torch.zeros(2048, dtype=torch.int64, device=dev)
with torch.no_grad():
    m(x); m(x)
torch.cuda.synchronize()
os.environ["LW_TRACE_PTR"] = str(buf.data_ptr()); os.environ["LW_DEBUG"] = "16"
# only the leff2 launch should trace: leff1 also honours dbg&16 but writes a different layout -> run leff2 alone via ops
p = m.packed()
import ctypes as Cx
h1 = torch.randn(32 * H * H, 4 * C, device=dev).to(torch.float16)
out = torch.empty_like(x)
from uformer_b200 import _lib
b = _lib.Leff2Args()
b.h1, b.out, b.resid = h1.data_ptr(), out.data_ptr(), x.data_ptr()
b.taps, b.w2_img, b.b2 = p["taps16"].data_ptr(), p["w2_img"].data_ptr(), p["b2"].data_ptr()
b.B, b.H, b.W, b.C, b.hidden = 32, H, H, C, 4 * C
_lib.check(_lib.load().lw_leff2_fwd(Cx.byref(b), torch.cuda.current_stream().cuda_stream), "leff2")
torch.cuda.synchronize()
t = buf.cpu().tolist(); n = t.index(-1); t = t[:n]; t0 = t[0]
d = [v - t0 for v in t]
print("start:", d[0])
k = 1
for kb in range(min(6, 4 * C // 64)):
    print(f"slice {kb}: after wait+bar {d[k]}, prefetch issued {d[k+1]}, taps->regs {d[k+2]}, conv done {d[k+3]}, a_empty ok {d[k+4]}, gelu+sts+arrive {d[k+5]}")
    k += 6
print("all slices done:", d[k], " d_full:", d[k + 1], " epilogue sub-chunks:", d[k + 2:])
