"""Per-role clock64 timeline of CTA 0 of the fused LeFF kernel (needs the -DLW_TRACE build of the library):
    nvcc ... -DLW_TRACE -o uformer_b200/lib/liblewin_b200_trace.so ; UFORMER_B200_LIB=<that> python tools/leff_fused_trace.py C H B"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("UFORMER_B200_LIB", os.path.join(ROOT, "uformer_b200", "lib", "liblewin_b200_trace.so"))
import torch
import uformer_b200 as U
from uformer_b200 import ops
from paramgen import randomize_state
C, H, B = map(int, sys.argv[1:4])
dev = "cuda:0"
blk = U.LeWinTransformerBlock(C, (max(H, 16),) * 2, max(1, C // 32), win_size=8).eval()
blk.load_state_dict(randomize_state(blk.state_dict(), 9)); blk = blk.to(dev)
pm = blk.mlp.packed(blk.norm2)
x = torch.randn(B, H * H, C, device=dev).to(torch.bfloat16)
buf = torch.zeros(7 * 512, dtype=torch.int64, device=dev)
with torch.no_grad():
    ops.leff(x, pm, B=B, H=H, W=H, resid=x); ops.leff(x, pm, B=B, H=H, W=H, resid=x)
    torch.cuda.synchronize()
    os.environ["LW_TRACE_PTR"] = str(buf.data_ptr()); os.environ["LW_DEBUG"] = "16"
    ops.leff(x, pm, B=B, H=H, W=H, resid=x)
    torch.cuda.synchronize()
t = buf.cpu().view(7, 512)
nz = t[t > 0]
t0 = int(nz.min())
NS = 4 * C // (64 if C <= 128 else 32)
def row(r, n): return [int(v) - t0 if v > 0 else None for v in t[r, :n].tolist()]
nsl = min(3 * NS, 120)
e1a, e1b, cv, iss = row(0, 2 * nsl), row(1, 2 * nsl), row(2, 4 * nsl), row(3, 2 * nsl)
print(f"C={C} NS={NS}; cycles relative to the first event; per slice k: G1 issue | E1a start-end | E1b start-end | conv: halo ok, reads done, a2 ok, end | G2 issue")
for k in range(nsl):
    print(f"k={k:3d} (tile {k // NS}, j={k % NS}): G1 {iss[2*k]} | E1a {e1a[2*k]}-{e1a[2*k+1]} | E1b {e1b[2*k]}-{e1b[2*k+1]} | conv {cv[4*k]} {cv[4*k+1]} {cv[4*k+2]} {cv[4*k+3]} | G2 {iss[2*k+1]}")
print("stats (x_full ok, done) per tile:", row(4, 8))
print("x load issued per tile:", row(5, 4))
print("E2 (start, d2_full ok, end) per tile:", row(6, 9))
