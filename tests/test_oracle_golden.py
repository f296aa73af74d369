"""CPU: pin the oracle restatement against the golden vectors generated from the unmodified reference
(tests/golden/make_golden.py).  Also proves our modules' state-dict keys/shapes equal the reference's:
the fixture stores a checksum of the reference state after seeded randomisation over sorted keys."""
import pytest
import torch

from helpers import TOL_FP32, build_module, golden_names, load_golden, oracle_run, rel_l2, rel_max, state_checksum

ALL = [n for k in ("wattn", "leff", "down", "up", "block", "model") for n in golden_names(k)]


@pytest.mark.parametrize("name", ALL)
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    _, st = build_module(g)
    assert abs(state_checksum(st) - g["checksum"]) <= 1e-6 * g["checksum"], "state-dict keys/shapes differ from the reference"
    y = oracle_run(g, st, g["x"], torch.float64)
    assert y.shape == g["y"].shape
    assert rel_l2(y, g["y"]) < 1e-5 and rel_max(y, g["y"]) < 1e-4      # fp64 oracle vs fp32 reference
    y32 = oracle_run(g, st, g["x"], torch.float32)
    assert rel_l2(y32, g["y"]) < TOL_FP32 * 1e-1
    if g["kind"] == "wattn":
        ym = oracle_run(g, st, g["x"], torch.float64, mask=g["mask"])
        assert rel_l2(ym, g["y_mask"]) < 1e-5


def test_shift_mask_closed_form():
    """shift_attn_mask must equal the procedural construction of model.py:924-942 (restated with slices here)."""
    from oracle import lewin_oracle as O
    for H, ws, s in [(16, 8, 4), (32, 8, 4), (24, 8, 3)]:
        m = torch.zeros(1, H, H, 1)
        cnt = 0
        for hs in (slice(0, -ws), slice(-ws, -s), slice(-s, None)):
            for wsl in (slice(0, -ws), slice(-ws, -s), slice(-s, None)):
                m[:, hs, wsl, :] = cnt
                cnt += 1
        mw = O.window_partition(m, ws).reshape(-1, ws * ws)
        d = mw.unsqueeze(1) - mw.unsqueeze(2)
        ref = torch.where(d != 0, torch.tensor(-100.0), torch.tensor(0.0))
        assert torch.equal(O.shift_attn_mask(H, H, ws, s), ref)


def test_window_roundtrip():
    from oracle import lewin_oracle as O
    x = torch.randn(2, 16, 24, 5)
    w = O.window_partition(x, 8)
    assert w.shape == (2 * 2 * 3, 8, 8, 5)
    assert torch.equal(O.window_reverse(w, 8, 16, 24), x)
    # window order: batch-major, then row-major over (wy, wx)
    assert torch.equal(w[4], x[0, 8:16, 8:16])


def test_fast_mode_equals_explicit():
    """oracle.FAST (library-op formulation used for the timed CPU baseline) == explicit restatement."""
    from oracle import lewin_oracle as O
    g = load_golden("uformer_t2_128")
    _, st = build_module(g)
    y0 = oracle_run(g, st, g["x"])
    O.FAST = True
    try:
        y1 = oracle_run(g, st, g["x"])
    finally:
        O.FAST = False
    assert rel_l2(y1, y0) < 1e-5
