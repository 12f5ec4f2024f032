"""Row-streaming pass A (csrc/cab_pass_a_stream.cu) against the 16x16-tile pass A (same C-ABI call, selected by a debug dump
request) and against the oracle, on shapes that exercise its scheduling: strips narrower than 26 px, heights that are not
multiples of 4, pieces cut in the middle of a strip, more work than one wave, one-frame and one-block problems."""
import os
import sys

import pytest
import torch

import golden_io as gio

sys.path.insert(0, os.path.join(gio.ROOT, "oracle"))
import shiftnet_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda:0"

SHAPES = [(1, 4, 26), (1, 3, 7), (2, 9, 33), (1, 40, 52), (3, 150, 170), (2, 96, 640), (5, 61, 27)]


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(DEV).half()


@pytest.fixture(scope="module")
def env():
    sd, spec = gio.synthetic_checkpoint("gshift_deblur2")
    eng = gio.pkg("host.engine").Engine(spec, sd, DEV)          # default: the 16x16-tile pass A
    eng_s = gio.pkg("host.engine").Engine(spec, {}, DEV)        # every pass A on the row-streaming kernel
    eng_s.sd, eng_s.pass_a_stream = eng.sd, True
    return sd, spec, eng, eng_s


@pytest.mark.parametrize("shape", SHAPES, ids=[f"T{t}_{h}x{w}" for t, h, w in SHAPES])
def test_stream_pass_a_vs_tile_pass_a_and_oracle(env, shape):
    sd, spec, eng, eng_s = env
    L = gio.pkg("host.lib")
    T, H, W = shape
    g = torch.Generator().manual_seed(T * 1000 + H)
    x = 0.5 * torch.randn(T, spec.c1, H, W, generator=g)
    blk = "stage1.encoder_level2"
    for which, p, mode in (("cab2_fwd", blk + ".encoder_level1.0", L.MODE_CAB2_FWD), ("cab2_rev", blk + ".encoder_level1_1.0", L.MODE_CAB2_REV),
                           ("cab1", blk + ".encoder_level1.1", L.MODE_CAB1)):
        xd = _nhwc(x)
        out_s = eng_s.gated_cab(p, xd, mode)
        out_t = eng.gated_cab(p, xd, mode)
        torch.cuda.synchronize()
        a, b = out_s.float().cpu(), out_t.float().cpu()
        assert torch.isfinite(a).all()
        rel = ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()
        if which == "cab1":
            ref = O.cab1(sd, p, x, False)
        else:
            ref = O.cab2(sd, p, O.channel_shift(x, which == "cab2_rev", True), spec.c1, False)
        ro = ((a.permute(0, 3, 1, 2) - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
        print(f"[stream] {shape} {which}: vs tile kernel {rel:.2e} ; vs oracle {ro:.2e}")
        assert rel < 1e-3 and ro < 3e-3, (shape, which, rel, ro)


def test_stream_pass_a_is_deterministic(env):
    sd, spec, eng, eng_s = env
    L = gio.pkg("host.lib")
    g = torch.Generator().manual_seed(9)
    xd = _nhwc(0.5 * torch.randn(4, 64, 180, 320, generator=g))
    p = "stage1.encoder_level2.encoder_level1.0"
    a = eng_s.gated_cab(p, xd, L.MODE_CAB2_FWD).clone()
    for _ in range(3):
        assert torch.equal(eng_s.gated_cab(p, xd, L.MODE_CAB2_FWD), a)


def test_stream_shift_block_vs_oracle(env):
    """A whole Encoder_shift_block with every pass A on the streaming kernel (fused LayerNorm producers feeding it)."""
    sd, spec, eng, eng_s = env
    g = torch.Generator().manual_seed(12)
    x = 0.5 * torch.randn(3, 64, 150, 170, generator=g)
    out = eng_s.shift_block("stage1.decoder_level1", _nhwc(x)).float().cpu().permute(0, 3, 1, 2)
    ref = O.shift_block(sd, "stage1.decoder_level1", x, O.ARCHS["gshift_deblur2"])
    r = ((out - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    print(f"[stream] shift block vs oracle: {r:.2e}")
    assert r < 5e-3
