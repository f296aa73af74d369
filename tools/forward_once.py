"""ncu target: build Uformer-B, warm up (weight packing, allocator), then ONE forward between cudaProfilerStart/Stop.
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/forward_once.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
dev = torch.device("cuda:0")
net = bench.build_engine(dev)
x = torch.rand(int(os.environ.get("PB", 32)), 3, 256, 256, device=dev)
for _ in range(2):
    net(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
net(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
