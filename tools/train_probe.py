"""Phase breakdown of one Uformer-B training step on a B200 (CUDA events) + the top device kernels of a step
(torch.profiler).  Usage: python tools/train_probe.py [batch] > gpurun_out/train_probe.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import uformer_b200  # noqa: E402
from bench import UFORMER_B  # noqa: E402
from paramgen import randomize_state  # noqa: E402
from uformer_b200 import modules  # noqa: E402
from uformer_b200.training import TrainStep  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
net = uformer_b200.Uformer(**UFORMER_B, drop_path_rate=0.1)
net.load_state_dict(randomize_state(net.state_dict(), 1234), strict=True)
net = net.to(dev)
ts = TrainStep(net)
torch.manual_seed(0)
clean = torch.rand(B, 3, 256, 256, device=dev)
noisy = (clean + 0.1 * torch.randn_like(clean)).clamp(0, 1)
for _ in range(3):
    ts(noisy, clean)
torch.cuda.synchronize()


def ev():
    return torch.cuda.Event(enable_timing=True)


phases = {}
for it in range(3):
    e = [ev() for _ in range(6)]
    net.train()
    e[0].record()
    modules.invalidate_packed()
    for m in net.modules():                     # rebuild every operand image (what the first forward after a step pays)
        if hasattr(m, "packed"):
            m.packed()
    e[1].record()
    restored = net(noisy)
    e[2].record()
    loss = ts.criterion(restored, clean)
    e[3].record()
    loss.backward()
    e[4].record()
    ts.optimizer.step(grad_scale=1.0, zero_grad=True)
    e[5].record()
    torch.cuda.synchronize()
    for k, (a, b) in zip(["repack", "forward_native", "loss", "backward_recompute", "adamw"], zip(e, e[1:])):
        phases.setdefault(k, []).append(a.elapsed_time(b))
out = {"batch": B, "phases_ms": {k: round(min(v), 3) for k, v in phases.items()}}
with torch.no_grad():
    net.eval()
    a, b = ev(), ev()
    a.record()
    net(noisy)
    b.record()
    torch.cuda.synchronize()
    out["forward_eval_ms"] = round(a.elapsed_time(b), 3)
try:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ts(noisy, clean)
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:25]
    out["top_kernels"] = [{"name": r.key[:90], "calls": r.count, "ms": round(r.device_time_total / 1e3, 3)} for r in rows]
    out["device_ms_total"] = round(sum(r.device_time_total for r in prof.key_averages()) / 1e3, 3)
except Exception as exc:                         # pragma: no cover
    out["profiler_error"] = repr(exc)
print(json.dumps(out))
