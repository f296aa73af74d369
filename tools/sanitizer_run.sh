mkdir -p gpurun_out
cat > /tmp/mini.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import uformer_b200 as U
from paramgen import randomize_state
from oracle import lewin_oracle as O
torch.manual_seed(0)
from uformer_b200 import _lib
_lib.load().lw_set_max_ctas(1)          # persistent kernels: one CTA walks every tile (multi-tile barrier phases under the sanitizer)
for C, heads, H, shift, modu in [(32, 1, 16, 4, False), (128, 4, 16, 0, True), (256, 8, 16, 4, False), (128, 4, 16, 4, False)]:
    blk = U.LeWinTransformerBlock(C, (16, 16), heads, win_size=8, shift_size=shift, modulator=modu).eval()
    st = randomize_state(blk.state_dict(), 3); blk.load_state_dict(st); blk = blk.cuda()
    x = torch.randn(1, H * H, C).to(torch.bfloat16)
    with torch.no_grad():
        y = blk(x.cuda()).float().cpu()
    ref = O.lewin_block(x.float(), st, "", heads, 8, shift)
    print(C, float((y - ref).norm() / ref.norm()))
d = U.Downsample(32, 64).cuda().eval(); u = U.Upsample(64, 32).cuda().eval()
with torch.no_grad():
    z = u(d(torch.randn(1, 256, 32, device="cuda").to(torch.bfloat16)))
torch.cuda.synchronize(); print("ok", z.shape)
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/mini.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r02_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/mini.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r02_sanitizer_racecheck.log
