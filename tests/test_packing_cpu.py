"""Host-side weight packing and engine host logic (no GPU): every layout the kernels read is decoded back from the packed blob
produced by shift-net_b200/host/packing.py and compared with the reference-shaped weights it came from."""
import re

import numpy as np
import pytest
import torch

import golden_io as gio

P = gio.pkg("host.packing")


def _precfg_offsets(shift):
    """Byte offsets of PreCfg<KC1> (csrc/cab_pass_a_pre.cu) parsed from the source, so a change on either side fails here."""
    src = open(gio.ROOT + "/shift-net_b200/csrc/cab_pass_a_pre.cu").read()
    assert re.search(r"OFF_C1 = 2 \* CIN \* 4;", src) and re.search(r"OFF_W1 = OFF_C1 \+ \(SHIFT \? 9 \* HC \* 2 : 0\);", src)
    assert re.search(r"OFF_DA = OFF_W1 \+ W1_BYTES;", src) and re.search(r"W1_BYTES = KC1 \* N \* 16;", src)
    assert re.search(r"DA_BYTES = 9 \* 2 \* C \* 2, DB_BYTES = 25 \* C \* 2, W2_BYTES = KC2 \* N \* 16;", src)
    C, cin = 64, 96 if shift else 64
    off_c1 = 2 * cin * 4
    off_w1 = off_c1 + (9 * 32 * 2 if shift else 0)
    off_da = off_w1 + (cin // 8) * 128 * 16
    off_db = off_da + 9 * 128 * 2
    off_w2 = off_db + 25 * 64 * 2
    return off_c1, off_w1, off_da, off_db, off_w2, off_w2 + 8 * 128 * 16


@pytest.mark.parametrize("shift", [True, False])
@pytest.mark.parametrize("arch", ["gshift_deblur2", "gshift_denoise2"])
def test_pass_a_blob_decodes_to_the_reference_weights(arch, shift):
    sd, spec = gio.synthetic_checkpoint(arch)
    k = 1 if spec.denoise else 0
    p = "stage1.decoder_level1.encoder_level1." + ("0" if shift else "1")
    blob = P.pack_cab_pass_a(sd, p, 64, shift, k)
    off_c1, off_w1, off_da, off_db, off_w2, end = _precfg_offsets(shift)
    assert blob.dtype == torch.uint8 and blob.numel() == end
    raw = blob.numpy().tobytes()
    cin = 96 if shift else 64
    ln = np.frombuffer(raw[:off_c1], dtype=np.float32)
    assert np.array_equal(ln[:cin], sd[p + ".norm.weight"].numpy()) and np.array_equal(ln[cin:], sd[p + ".norm.bias"].numpy())
    if shift:
        c1 = torch.from_numpy(np.frombuffer(raw[off_c1:off_w1], dtype=np.float16).copy()).view(9, 32)
        assert torch.equal(c1, sd[p + ".conv1.weight"].view(32, 9).t().half())
    # first 1x1 in the k-chunk planar UMMA layout [K/8][N][8]: element (n, k) at chunk k//8, row n, lane k%8
    w1 = torch.from_numpy(np.frombuffer(raw[off_w1:off_da], dtype=np.float16).copy()).view(cin // 8, 128, 8)
    assert torch.equal(w1.permute(1, 0, 2).reshape(128, cin), sd[p + ".body.0.weight"].flatten(1).half())
    # dw3x3 taps [9][2C] with RepConv2's identity on the centre tap
    da = torch.from_numpy(np.frombuffer(raw[off_da:off_db], dtype=np.float16).copy()).view(9, 128)
    ref = sd[p + ".body.1.conv_2.weight"].view(128, 9).t().clone()
    ref[4] += 1.0
    assert torch.equal(da, ref.half())
    # merged 5x5 (+3x3, +identity) taps [25][C]
    db = torch.from_numpy(np.frombuffer(raw[off_db:off_w2], dtype=np.float16).copy()).view(25, 64)
    w5 = sd[p + f".body.{3 + k}.conv_1.weight"].clone()
    w5[:, :, 1:4, 1:4] += sd[p + f".body.{3 + k}.conv_2.weight"]
    w5[:, :, 2, 2] += 1.0
    assert torch.equal(db, w5.view(64, 25).t().half())
    w2 = torch.from_numpy(np.frombuffer(raw[off_w2:end], dtype=np.float16).copy()).view(8, 128, 8)
    assert torch.equal(w2.permute(1, 0, 2).reshape(128, 64), sd[p + f".body.{4 + k}.weight"].flatten(1).half())


def test_merged_repconv_taps_equal_the_three_branch_sum():
    """RepConv = dw5x5(x) + dw3x3(x) + x (gshift_deblur2.py:159-168) as ONE 5x5 per channel: the packed taps applied as a
    depthwise conv reproduce the three-branch sum."""
    import torch.nn.functional as F
    sd, _ = gio.synthetic_checkpoint("gshift_deblur2")
    p = "stage1.decoder_level1.encoder_level1.1"
    w5, w3 = sd[p + ".body.3.conv_1.weight"], sd[p + ".body.3.conv_2.weight"]
    x = torch.randn(2, 64, 12, 13, generator=torch.Generator().manual_seed(1))
    ref = F.conv2d(x, w5, padding=2, groups=64) + F.conv2d(x, w3, padding=1, groups=64) + x
    m = w5.clone()
    m[:, :, 1:4, 1:4] += w3
    m[:, :, 2, 2] += 1.0
    assert torch.allclose(F.conv2d(x, m, padding=2, groups=64), ref, atol=1e-5)


def test_group_conv5_fragments_decode_to_the_merged_grouped_conv():
    """pack_group_conv5 (csrc/generic_cab.cu group_conv5): per-lane mma B fragments [group][13 k-steps][32 lanes][b0 (tap 2k) | b1
    (tap 2k+1)] x 2 input channels; rebuilding W[o][i][tap] from them gives conv_1 + zero-padded conv_2 (the identity is added
    by the kernel from the centre pixel)."""
    g = torch.Generator().manual_seed(2)
    w5, w3 = torch.randn(80, 8, 5, 5, generator=g), torch.randn(80, 8, 3, 3, generator=g)
    f = P.pack_group_conv5(w5, w3).float().view(10, 13, 32, 2, 2)     # group, kstep, lane, tsel, e
    W = torch.zeros(80, 8, 26)
    for lane in range(32):
        n, tig = lane >> 2, lane & 3
        for ts in range(2):
            for e in range(2):
                for ks in range(13):
                    W[torch.arange(10) * 8 + n, 2 * tig + e, 2 * ks + ts] = f[:, ks, lane, ts, e]
    ref = w5.clone()
    ref[:, :, 1:4, 1:4] += w3
    assert torch.equal(W[:, :, :25], ref.reshape(80, 8, 25).half().float()) and W[:, :, 25].abs().max() == 0


def test_engine_host_staging_of_a_device_free_checkpoint():
    """Engine._to_host: a CPU fp16/fp32 checkpoint becomes fp32 host tensors under the same keys and shapes (the GPU box runs the
    same code on CUDA tensors through one concatenated copy per dtype)."""
    Engine = gio.pkg("host.engine").Engine
    sd, _ = gio.synthetic_checkpoint("gshift_deblur2")
    half = {k: v.half() for k, v in sd.items()}
    out = Engine._to_host(half)
    assert set(out) == set(sd)
    for k in list(sd)[:50]:
        assert out[k].dtype == torch.float32 and out[k].shape == sd[k].shape and torch.equal(out[k], sd[k].half().float())


def test_fold_pack_has_the_mid_attention_for_denoise_only():
    sd, spec = gio.synthetic_checkpoint("gshift_denoise2")
    p = "stage1.decoder_level1.encoder_level1.1"
    d = P.pack_cab_fold(sd, p, 1)
    assert d["bias3"] is not None and d["mid_du0"].shape == (16, 64) and d["mid_du2"].shape == (64, 16) and d["w2"].shape == (128, 64)
    sd2, _ = gio.synthetic_checkpoint("gshift_deblur2")
    d2 = P.pack_cab_fold(sd2, p, 0)
    assert d2["bias3"] is None and "mid_du0" not in d2 and d2["w3"].shape == (64, 64)


@pytest.mark.parametrize("cin", [80, 120])
def test_layernorm_folding_around_the_first_1x1(cin):
    """pack_ln_pw_tc (csrc/ln_pw_tc.cu): rstd * (W' x - mu rowsum(W')) + W beta with the packed (fp16-rounded) W' reproduces
    W . LayerNorm(x) (gshift_deblur1.py:19-28,209,250) -- also for pixels whose mean is large against their spread, where the mean
    term only cancels because rowsum is taken over the ROUNDED weights."""
    g = torch.Generator().manual_seed(cin)
    w1 = torch.randn(160, cin, generator=g) / cin ** 0.5
    gamma, beta = 1 + 0.1 * torch.randn(cin, generator=g), 0.1 * torch.randn(cin, generator=g)
    wfold, wvec = P.pack_ln_pw_tc(w1, gamma, beta)
    assert wfold.shape == (16, 160, 8) and wfold.dtype == torch.float16 and wvec.shape == (320,)
    wp = wfold.float().permute(1, 0, 2).reshape(160, 128)
    assert wp[:, cin:].abs().max() == 0
    x = (torch.randn(64, cin, generator=g) * torch.rand(64, 1, generator=g) + 3.0 * torch.randn(64, 1, generator=g)).half().float()
    mu = x.mean(1, keepdim=True)
    rstd = ((x - mu).pow(2).mean(1, keepdim=True) + 1e-6).rsqrt()
    got = rstd * (x @ wp[:, :cin].t() - mu * wvec[:160]) + wvec[160:]
    ref = (((x - mu) * rstd) * gamma + beta) @ w1.t()
    assert ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item() < 1e-3
