import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench
import uformer_b200 as U
dev = torch.device("cuda:0")
net = bench.build_engine(dev)
x = torch.rand(32, 3, 256, 256, device=dev)
g = U.GraphedForward(net, x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3): g(g.x)
evs = []
for _ in range(10):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g(g.x); e.record(); evs.append((s, e))
torch.cuda.synchronize()
ms = sum(s.elapsed_time(e) for s, e in evs) / 10
print("budget", os.environ.get("UFORMER_B200_L2_BUDGET_MB", "off"), "graph replay ms/step %.2f -> %.0f img/s" % (ms, 32 / ms * 1e3), flush=True)
