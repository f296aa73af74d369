"""CPU, build container only: re-check the oracle live against the unmodified reference on fresh random
inputs (skipped where /root/reference does not exist, e.g. on the GPU box)."""
import pytest
import torch

from refshim import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not mounted")


def test_block_live():
    from oracle import lewin_oracle as O
    from paramgen import randomize_state
    from refshim import import_reference_model
    m = import_reference_model()
    torch.manual_seed(7)
    for dim, heads, H, shift, modu in [(32, 1, 16, 4, True), (64, 4, 32, 4, False)]:
        blk = m.LeWinTransformerBlock(dim, (H, H), heads, win_size=8, shift_size=shift, modulator=modu).eval()
        st = randomize_state(blk.state_dict(), 21)
        blk.load_state_dict(st)
        x = torch.randn(2, H * H, dim)
        with torch.no_grad():
            ref = blk(x)
        got = O.lewin_block(x, st, "", heads, 8, shift)
        assert (got - ref).abs().max() < 1e-4 * ref.abs().max()


def test_input_mask_path_live():
    """The optional input-mask branch (model.py:914-921) with batch 1."""
    from oracle import lewin_oracle as O
    from paramgen import randomize_state
    from refshim import import_reference_model
    m = import_reference_model()
    torch.manual_seed(8)
    blk = m.LeWinTransformerBlock(32, (16, 16), 2, win_size=8, shift_size=0).eval()
    st = randomize_state(blk.state_dict(), 22)
    blk.load_state_dict(st)
    x = torch.randn(1, 256, 32)
    mask = (torch.rand(1, 1, 16, 16) > 0.5).float()
    with torch.no_grad():
        ref = blk(x, mask)
    got = O.lewin_block(x, st, "", 2, 8, 0, input_mask=mask)
    assert (got - ref).abs().max() < 1e-4 * ref.abs().max()


def test_install_builds_reference_uformer_on_engine():
    import uformer_b200
    from refshim import import_reference_model
    m = import_reference_model()
    cfg = dict(img_size=128, embed_dim=16, depths=[1] * 9, win_size=8, token_projection="linear", token_mlp="leff", modulator=True)
    ref_state = m.Uformer(**cfg).state_dict()
    uformer_b200.install(m)
    try:
        net = m.Uformer(**cfg)
        assert type(net.encoderlayer_0.blocks[0]) is uformer_b200.LeWinTransformerBlock
        assert type(net.dowsample_0) is uformer_b200.Downsample and type(net.upsample_3) is uformer_b200.Upsample
        net.load_state_dict(ref_state, strict=True)
        with pytest.raises(uformer_b200.EngineUnavailable):
            net(torch.rand(1, 3, 128, 128))       # no CPU fallback
    finally:
        uformer_b200.uninstall(m)
    assert m.Uformer(**cfg).encoderlayer_0.blocks[0].__class__.__module__ == "model"


def test_uformer_win16_live():
    """A whole reference Uformer built with win_size = 16 (16x16 windows, shift 8, the clamp of model.py:863-865 at the 16x16- and
    8x8-token stages) against the oracle — the pin behind tests/test_host_path_cpu.py::test_uformer_with_16x16_windows_..."""
    from oracle import lewin_oracle as O
    from paramgen import randomize_state
    from refshim import import_reference_model
    m = import_reference_model()
    cfg = dict(img_size=128, embed_dim=16, depths=[2] * 9, win_size=16, token_projection="linear", token_mlp="leff", modulator=False)
    net = m.Uformer(**cfg).eval()
    st = randomize_state(net.state_dict(), 31)
    net.load_state_dict(st)
    x = torch.rand(1, 3, 128, 128, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = net(x)
    got = O.uformer_forward(x.double(), {k: (v.double() if torch.is_floating_point(v) else v) for k, v in st.items()}, 128, 16, [2] * 9, win_size=16)
    assert ((got.float() - ref).norm() / ref.norm()).item() < 1e-5
