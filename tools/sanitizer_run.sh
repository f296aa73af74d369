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
# 16 x 16 windows (wmsa16_kernel): shifted + modulated, two heads
blk = U.LeWinTransformerBlock(32, (32, 32), 2, win_size=16, shift_size=8, modulator=True).eval()
st = randomize_state(blk.state_dict(), 4); blk.load_state_dict(st); blk = blk.cuda()
x = torch.randn(1, 32 * 32, 32).to(torch.bfloat16)
with torch.no_grad():
    y = blk(x.cuda()).float().cpu()
print("ws16", float((y - O.lewin_block(x.float(), st, "", 2, 16, 8)).norm() / y.norm()))
# C = 512 block: wmsa_kernel + the two-kernel LeFF with the N range of its single tile split over CTAs (ares nsplit)
blk = U.LeWinTransformerBlock(512, (16, 16), 16, win_size=8, shift_size=4).eval()
st = randomize_state(blk.state_dict(), 5); blk.load_state_dict(st); blk = blk.cuda()
x = torch.randn(1, 256, 512).to(torch.bfloat16)
with torch.no_grad():
    y = blk(x.cuda()).float().cpu()
print("C512", float((y - O.lewin_block(x.float(), st, "", 16, 8, 4)).norm() / y.norm()))
# OutputProj on the tensor core (TMA halo tile, persistent, several tiles through one CTA)
from uformer_b200 import ops
tok = torch.randn(1, 24 * 40, 64, device="cuda").to(torch.bfloat16)
o = ops.output_proj(tok, torch.randn(3, 64, 3, 3, device="cuda") * 0.05, torch.randn(3, device="cuda"), torch.rand(1, 3, 24, 40, device="cuda"), 24, 40)
print("outproj", float(o.abs().max()))
d = U.Downsample(32, 64).cuda().eval(); u = U.Upsample(64, 32).cuda().eval()
with torch.no_grad():
    z = u(d(torch.randn(1, 256, 32, device="cuda").to(torch.bfloat16)))
torch.cuda.synchronize(); print("ok", z.shape)
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/mini.py > gpurun_out/${TAG:-r02}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/${TAG:-r02}_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/mini.py > gpurun_out/${TAG:-r02}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/${TAG:-r02}_sanitizer_racecheck.log
