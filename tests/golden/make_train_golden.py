"""Generate the training-step golden vector from the UNMODIFIED reference (build container only).

    python tests/golden/make_train_golden.py

Reference path: model.py Uformer (train mode, drop_path_rate=0) -> losses.py CharbonnierLoss -> loss.backward(), fp32 on
CPU, seeded synthetic weights (tests/paramgen.py).  Stored: input, target, loss, the restored image and — because the
full gradient set is 25 MB — for every parameter its gradient's L2 norm and a strided sample (<= 1024 elements,
`flat[::stride]`).  tests/test_train_cpu.py checks the restated backward math against it on CPU;
tests/test_gpu_train.py checks the native-forward / recompute-backward path on the B200.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from refshim import REFERENCE_DIR, import_reference_model  # noqa: E402
from paramgen import randomize_state  # noqa: E402

CFG = dict(img_size=128, embed_dim=16, depths=[2] * 9, win_size=8, token_projection="linear", token_mlp="leff", modulator=True,
           drop_path_rate=0.0)
SEED = 4321


def sample(t, cap=1024):
    flat = t.reshape(-1)
    stride = max(1, (flat.numel() + cap - 1) // cap)
    return stride, flat[::stride].clone()


def main():
    m = import_reference_model()
    sys.path.insert(0, REFERENCE_DIR)
    from losses import CharbonnierLoss
    torch.manual_seed(SEED)
    net = m.Uformer(**CFG)
    st = randomize_state(net.state_dict(), SEED)
    net.load_state_dict(st, strict=True)
    net.train()
    clean = torch.rand(2, 3, 128, 128)
    noisy = (clean + 0.1 * torch.randn_like(clean)).clamp(0, 1)
    restored = net(noisy)
    loss = CharbonnierLoss()(restored, clean)
    loss.backward()
    grads = {}
    for k, p in net.named_parameters():
        stride, smp = sample(p.grad)
        grads[k] = dict(norm=float(p.grad.double().norm()), stride=stride, sample=smp)
    out = dict(kind="train", cfg=CFG, seed=SEED, x=noisy, target=clean, loss=float(loss), y=restored.detach(), grads=grads)
    path = os.path.join(HERE, "train_t2_128.pt")
    torch.save(out, path)
    print("train_t2_128 %8.1f KB  loss=%.6f" % (os.path.getsize(path) / 1024, float(loss)))


if __name__ == "__main__":
    main()
