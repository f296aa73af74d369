"""ncu target: build Uformer-B, warm up (weight packing, allocator), then ONE forward between cudaProfilerStart/Stop.
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/forward_once.py
Writes the launch labels of that forward (uformer_b200.ops labels, in launch order) to $LABELS_OUT so that the ncu rows can be
matched to bench.py's kernel classes (tools/kernel_metrics.py)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from uformer_b200 import ops
dev = torch.device("cuda:0")
net = bench.build_engine(dev)
x = torch.rand(int(os.environ.get("PB", 32)), 3, 256, 256, device=dev)
for _ in range(2):
    net(x)
torch.cuda.synchronize()
ops.PROFILE = []
torch.cuda.profiler.start()
net(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
labels = [r[0] for r in ops.PROFILE]
ops.PROFILE = None
if os.environ.get("LABELS_OUT"):
    json.dump(labels, open(os.environ["LABELS_OUT"], "w"))
print("done", len(labels), "launches")
