import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import uformer_b200 as U
from paramgen import randomize_state
dev = torch.device("cuda:0")
C, H = int(sys.argv[1]), int(sys.argv[2])
m = U.LeFF(C, 4 * C).eval(); m.load_state_dict(randomize_state(m.state_dict(), 1)); m = m.to(dev)
x = torch.randn(32, H * H, C, device=dev).to(torch.bfloat16)
buf = torch.zeros(2048, dtype=torch.int64, device=dev)
with torch.no_grad():
    m(x); m(x)
torch.cuda.synchronize()
os.environ["LW_TRACE_PTR"] = str(buf.data_ptr()); os.environ["LW_DEBUG"] = str(16 + int(sys.argv[3]))
with torch.no_grad():
    m(x)
torch.cuda.synchronize()
t = buf.cpu().tolist()
iss = t[:512]; n = iss.index(-1); iss = iss[:n]
wk = t[512:]; n = wk.index(-1); wk = wk[:n]
t0 = iss[0]
print("issuer (cycles since start): start, a_ready, then per N-chunk: d_empty, [acquire, release]*KB")
print([v - t0 for v in iss])
print("worker0: start, then per chunk: d_full, (after sub-chunk)*")
print([v - t0 for v in wk])
fine = t[1024:]; n = fine.index(-1); fine = fine[:n]
print("chunk 1 fine (per sub-chunk: start, after tmem ld, after math+stsm, after bar1, after phase B) :", [v - t0 for v in fine])
