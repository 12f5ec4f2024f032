"""GPU parity suite (-m gpu): every CUDA op and the whole net, called through the C-ABI, against the CPU oracle
(same seeded inputs) and against the committed reference goldens.  Tolerances: activations are stored in fp16
(2^-11 relative per stage, fp32 accumulation), so single ops must agree to ~1e-3 relative RMS and the whole net to
>= 60 dB PSNR against the fp32 reference (the reference's own fp16 path reaches 66-68 dB)."""
import dataclasses
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import golden_io as gio

sys.path.insert(0, os.path.join(gio.ROOT, "oracle"))
import shiftnet_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda:0"


def to_nhwc(x, cp=None):
    """(T,C,H,W) fp32 cpu -> (T,H,W,Cp) fp16 cuda, zero padded channels."""
    T, C, H, W = x.shape
    cp = cp or (C + 7) // 8 * 8          # storage convention: channels padded to one 16-byte vector
    out = torch.zeros(T, H, W, cp, dtype=torch.float16, device=DEV)
    out[..., :C] = x.permute(0, 2, 3, 1).to(DEV).half()
    return out


def from_nhwc(t, c):
    return t[..., :c].permute(0, 3, 1, 2).float().cpu()


def rel_rms(a, ref):
    return ((a - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-12)).item()


def check(a, ref, tol, what):
    assert a.shape == ref.shape, (what, a.shape, ref.shape)
    assert torch.isfinite(a).all(), what
    r = rel_rms(a, ref)
    m = (a - ref).abs().max().item()
    print(f"[parity] {what}: rel_rms={r:.2e} max_abs={m:.2e} ref_rms={ref.pow(2).mean().sqrt().item():.3f}")
    assert r < tol, (what, r, m)


CUDA_ARCHS = ["gshift_deblur2", "gshift_denoise2", "gshift_deblur1", "gshift_denoise1"]


def _make_env(arch):
    sd, spec = gio.synthetic_checkpoint(arch)
    eng = gio.pkg("host.engine").Engine(spec, {k: v.to(DEV) for k, v in sd.items()}, DEV)
    return sd, spec, eng, gio.load_golden(arch)


@pytest.fixture(scope="module")
def env():
    return _make_env("gshift_deblur2")


@pytest.fixture(scope="module", params=CUDA_ARCHS)
def aenv(request):
    return _make_env(request.param)


# ------------------------------------------------------------------------------------------------ dense conv
CONV_CASES = [
    # name, src channels, cout, k, stride, pad, bias, prelu, residual, pixel_shuffle
    ("3x3_14_14", [14], 14, 3, 1, 1, False, False, False, False),
    ("3x3_s2_bias_14_18", [14], 18, 3, 2, 1, True, False, False, False),
    ("1x1_22_18", [22], 18, 1, 1, 0, False, False, False, False),
    ("2x2_s2_prelu_14_64", [14], 64, 2, 2, 0, False, True, False, False),
    ("3x3_s2_bias_64_64", [64], 64, 3, 2, 1, True, False, False, False),
    ("3x3_cat3_bias_14", [14, 14, 14], 14, 3, 1, 1, True, False, False, False),
    ("3x3_res_14_14", [14], 14, 3, 1, 1, False, False, True, False),
    ("3x3_shuffle_prelu_64_56", [64], 56, 3, 1, 1, True, True, False, True),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_mma(env, case):
    _, _, eng, _ = env
    name, srcs_c, cout, k, stride, pad, bias, prelu, residual, shuffle = case
    g = torch.Generator().manual_seed(5)
    T, H, W = 2, 20, 28                       # not multiples of the 16x16 tile
    xs = [torch.randn(T, c, H, W, generator=g) for c in srcs_c]
    w = torch.randn(cout, sum(srcs_c), k, k, generator=g) / (sum(srcs_c) * k * k) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1 if bias else None
    key = "test." + name
    eng.sd[key + ".weight"] = w.to(DEV)
    if bias:
        eng.sd[key + ".bias"] = b.to(DEV)
    eng.sd[key + ".slope"] = torch.tensor([0.2], device=DEV)
    xq = [x.half().float() for x in xs]       # the kernel sees fp16-rounded inputs
    ref = F.conv2d(torch.cat(xq, 1), w.half().float(), b, stride=stride, padding=pad)
    if prelu:
        ref = F.prelu(ref, torch.tensor([0.2]))
    res_t = None
    if residual:
        r = torch.randn(ref.shape, generator=g)
        res_t = to_nhwc(r)
        ref = ref + r.half().float()
    if shuffle:
        ref = F.pixel_shuffle(ref, 2)
    out = eng.conv(key, [to_nhwc(x) for x in xs], srcs_c, cout, stride=stride, pad=pad,
                   prelu_key=key + ".slope" if prelu else None, residual=res_t, pixel_shuffle=shuffle)
    torch.cuda.synchronize()
    cr = cout // 4 if shuffle else cout
    got = from_nhwc(out, cr)
    check(got, ref, 2e-3, "conv " + name)
    if out.shape[-1] > cr:
        assert out[..., cr:].abs().max().item() == 0.0, "padding channels must stay zero"


# ------------------------------------------------------------------------------------------------ fused dense CAB body
DENSE_CASES = [
    # name, channels, T, H, W, bias      (sizes straddle the 30-wide / 30- or 14-row tiles, incl. a single partial tile)
    ("c14_ragged", 14, 2, 37, 45, False),
    ("c14_tiny", 14, 1, 5, 7, False),
    ("c14_exact_tiles", 14, 1, 60, 60, False),
    ("c14_bias", 14, 1, 33, 31, True),
    ("c18_ragged", 18, 2, 23, 41, False),
    ("c22_bias", 22, 1, 29, 64, True),
    ("c24_full_width", 24, 1, 16, 35, False),
]


@pytest.mark.parametrize("small", [False, True], ids=["tile30", "tile14"])
@pytest.mark.parametrize("case", DENSE_CASES, ids=[c[0] for c in DENSE_CASES])
def test_cab_dense_body(env, case, small):
    """gsn_cab_dense (conv3x3 -> PReLU -> conv3x3 + channel sums, one kernel) against torch fp32 on the fp16-rounded
    input, and against the two-launch gsn_conv_mma path."""
    _, _, eng, _ = env
    name, c, T, H, W, bias = case
    if small and c > 16:
        pytest.skip("the 14-row tile switch only changes the 16-channel instance")
    g = torch.Generator().manual_seed(11)
    x = torch.randn(T, c, H, W, generator=g)
    key = "testdense." + name
    w1 = torch.randn(c, c, 3, 3, generator=g) / (c * 9) ** 0.5
    w2 = torch.randn(c, c, 3, 3, generator=g) / (c * 9) ** 0.5
    eng.sd[key + ".body.0.weight"], eng.sd[key + ".body.2.weight"] = w1.to(DEV), w2.to(DEV)
    eng.sd[key + ".body.1.weight"] = torch.tensor([0.25], device=DEV)
    b1 = b2 = None
    if bias:
        b1, b2 = torch.randn(c, generator=g) * 0.1, torch.randn(c, generator=g) * 0.1
        eng.sd[key + ".body.0.bias"], eng.sd[key + ".body.2.bias"] = b1.to(DEV), b2.to(DEV)
    xq = x.half().float()
    mid = F.prelu(F.conv2d(xq, w1.half().float(), b1, padding=1), torch.tensor([0.25]))
    ref = F.conv2d(mid.half().float(), w2.half().float(), b2, padding=1)     # the kernel keeps the intermediate in fp16
    xn = to_nhwc(x)
    os.environ["GSN_CAB_DENSE_SMALL"] = "1" if small else "0"
    try:
        r, partial = eng.cab_body_fused(key, xn, c)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("GSN_CAB_DENSE_SMALL", None)
    got = from_nhwc(r, c)
    check(got, ref, 2e-3, "cab_dense " + name)
    cp = xn.shape[3]
    if cp > c:
        assert r[..., c:].abs().max().item() == 0.0, "padding channels must stay zero"
    sums = partial.sum(1)[:, :c].cpu()
    ref_s = r[..., :c].float().sum((1, 2)).cpu()
    assert torch.allclose(sums, ref_s, rtol=2e-3, atol=2e-2 * (H * W) ** 0.5), (sums - ref_s).abs().max()
    # the two-launch path must agree to fp16 rounding
    r1 = eng.conv(key + ".body.0", [xn], [c], c, prelu_key=key + ".body.1.weight")
    r2 = eng.conv(key + ".body.2", [r1], [c], c)
    torch.cuda.synchronize()
    check(got, from_nhwc(r2, c), 2e-3, "cab_dense vs conv_mma " + name)


def test_conv_in_out(env):
    sd, spec, eng, _ = env
    P, L = gio.pkg("host.packing"), gio.pkg("host.lib")
    import ctypes as C
    g = torch.Generator().manual_seed(6)
    T, H, W = 3, 20, 36
    x = torch.rand(T, 3, H, W, generator=g)
    for dt, tdt in ((L.DTYPE_F16, torch.float16), (L.DTYPE_F32, torch.float32)):
        xd = x.to(DEV, tdt).contiguous()
        wi, bi = eng._up(P.pack_conv_in(eng.sd["feat_extract.0.weight"], eng.sd["feat_extract.0.bias"], 16))   # packed on the host
        f0 = torch.empty(T, H, W, 16, dtype=torch.float16, device=DEV)
        L.check(eng.lib.gsn_conv_in(xd.data_ptr(), dt, T, 3, H, W, wi.data_ptr(), bi.data_ptr(), 16, f0.data_ptr(), eng._stream()))
        ref = F.conv2d(xd.float().cpu(), sd["feat_extract.0.weight"], sd["feat_extract.0.bias"], padding=1)
        check(from_nhwc(f0, 14), ref, 1.5e-3, f"conv_in dtype={dt}")
        assert f0[..., 14:].abs().max().item() == 0.0
        # conv_out: 5x5 14->3 + residual
        feat = torch.randn(T, 14, H, W, generator=g)
        wo = eng._up(P.pack_conv_out(eng.sd["conv_last.weight"], 16))
        out = torch.empty(T, 3, H, W, dtype=tdt, device=DEV)
        fd = to_nhwc(feat)
        L.check(eng.lib.gsn_conv_out(fd.data_ptr(), 16, 5, wo.data_ptr(), xd.data_ptr(), 3, dt, T, H, W, out.data_ptr(), eng._stream()))
        ref = F.conv2d(feat.half().float(), sd["conv_last.weight"], None, padding=2) + xd.float().cpu()
        check(out.float().cpu(), ref, 1.5e-3, f"conv_out dtype={dt}")


def test_upsample_add(env):
    _, _, eng, _ = env
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 18, 10, 14, generator=g)
    skip = torch.randn(2, 18, 20, 28, generator=g)
    out = torch.empty(2, 20, 28, 24, dtype=torch.float16, device=DEV)
    xd, sk = to_nhwc(x), to_nhwc(skip)      # keep the device tensors alive across the async launch (18 -> 24 channels)
    gio.pkg("host.lib").check(eng.lib.gsn_upsample2x_add(xd.data_ptr(), sk.data_ptr(), out.data_ptr(), 2, 10, 14, 24, eng._stream()))
    torch.cuda.synchronize()
    ref = F.interpolate(x.half().float(), scale_factor=2, mode="bilinear", align_corners=False) + skip.half().float()
    check(from_nhwc(out, 18), ref, 1e-3, "upsample2x_add")


# ------------------------------------------------------------------------------------------------ blocks
def test_cab(aenv):
    sd, spec, eng, gold = aenv
    x = gio.module_input("cab", spec)
    out = from_nhwc(eng.cab("feat_extract.1", to_nhwc(x), spec.n0), spec.n0)
    check(out, O.cab(sd, "feat_extract.1", x), 3e-3, "cab vs oracle")
    check(out, torch.from_numpy(gold["cab"]), 3e-3, "cab vs reference golden")


@pytest.mark.parametrize("which", ["cab2_fwd", "cab2_rev", "cab1"])
@pytest.mark.parametrize("circular", [True, False])
def test_gated_cab(aenv, which, circular):
    sd, spec, eng, gold = aenv
    L = gio.pkg("host.lib")
    x = gio.module_input("shift", spec)
    blk = "stage1.decoder_level1"
    eng2 = gio.pkg("host.engine").Engine(dataclasses.replace(spec, circular=circular), {}, DEV)
    eng2.sd = eng.sd
    if which == "cab1":
        p, mode = blk + ".encoder_level1.1", L.MODE_CAB1
        ref = O.cab1(sd, p, x, spec.denoise)
    else:
        rev = which == "cab2_rev"
        p = blk + (".encoder_level1_1.0" if rev else ".encoder_level1.0")
        mode = L.MODE_CAB2_REV if rev else L.MODE_CAB2_FWD
        ref = O.cab2(sd, p, O.channel_shift(x, rev, circular), spec.c1, spec.denoise)
    out = from_nhwc(eng2.gated_cab(p, to_nhwc(x), mode), spec.c1)
    check(out, ref, 3e-3, f"{spec.name} {which} circular={circular} vs oracle")
    if circular == spec.circular:
        check(out, torch.from_numpy(gold[which]), 3e-3, f"{which} vs reference golden")


@pytest.mark.parametrize("rev", [False, True])
def test_gated_cab_roll_halo_equals_wrap_over_one_more_frame(aenv, rev):
    """GSN_ROLL_HALO (T-sharded clips, include/shiftnet_b200.h): CAB2 on the n own frames of an (n+1)-frame buffer, the roll reaching
    frame n by wrapping, must equal -- bit for bit -- the first n frames of the wrapping CAB2 over all n+1 frames.  All four nets
    (the C=64 fused kernels and the C=80 TMA/tcgen05 kernels build their tensor maps over n+1 frames)."""
    sd, spec, eng, _ = aenv
    L = gio.pkg("host.lib")
    g = torch.Generator().manual_seed(11)
    n = 3
    full = to_nhwc(0.5 * torch.randn(n + 1, spec.c1, 40, 56, generator=g))
    p = "stage1.decoder_level1" + (".encoder_level1_1.0" if rev else ".encoder_level1.0")
    mode = L.MODE_CAB2_REV if rev else L.MODE_CAB2_FWD
    try:
        eng._circ_override = L.ROLL_WRAP
        want = eng.gated_cab(p, full, mode)[:n].clone()
        eng._circ_override = L.ROLL_HALO
        have = eng.gated_cab(p, full[:n], mode)
    finally:
        eng._circ_override = None
    assert have.shape == want.shape and torch.equal(have, want)


def test_gated_cab_ragged_sizes(env):
    """H, W not multiples of the tile, T=1 and T=2 (wrap onto itself / neighbour is the only other frame)."""
    sd, spec, eng, _ = env
    L = gio.pkg("host.lib")
    blk = "stage1.encoder_level2"
    for T, H, W in ((1, 12, 20), (2, 9, 33), (3, 26, 18)):
        g = torch.Generator().manual_seed(T * 100 + H)
        x = 0.5 * torch.randn(T, spec.c1, H, W, generator=g)
        for rev in (False, True):
            p = blk + (".encoder_level1_1.0" if rev else ".encoder_level1.0")
            ref = O.cab2(sd, p, O.channel_shift(x, rev, True), spec.c1, False)
            out = from_nhwc(eng.gated_cab(p, to_nhwc(x), L.MODE_CAB2_REV if rev else L.MODE_CAB2_FWD), spec.c1)
            check(out, ref, 3e-3, f"cab2 rev={rev} T={T} {H}x{W}")
        ref = O.cab1(sd, blk + ".encoder_level1.1", x, False)
        out = from_nhwc(eng.gated_cab(blk + ".encoder_level1.1", to_nhwc(x), L.MODE_CAB1), spec.c1)
        check(out, ref, 3e-3, f"cab1 T={T} {H}x{W}")


@pytest.mark.parametrize("arch", ["gshift_deblur2", "gshift_denoise2"])
def test_fused_block_many_tiles_vs_oracle(arch):
    """The fused C=64 block on a clip with more tiles than SMs (persistent loops, cross-tile pipelining, ragged borders):
    every CAB against the oracle, and the fused LayerNorm producers (shift_conv1_ln, pass-B epilogue) against the un-fused
    chain gsn_shift_conv1 -> gsn_ln_planar."""
    sd, spec, eng, _ = _make_env(arch)
    L = gio.pkg("host.lib")
    blk = "stage1.encoder_level2"
    g = torch.Generator().manual_seed(5)
    x = 0.5 * torch.randn(3, spec.c1, 150, 170, generator=g)         # 10 x 11 x 3 = 330 tiles of 16x16
    for which, p, mode in (("cab2_fwd", blk + ".encoder_level1.0", L.MODE_CAB2_FWD), ("cab2_rev", blk + ".encoder_level1_1.0", L.MODE_CAB2_REV),
                           ("cab1", blk + ".encoder_level1.1", L.MODE_CAB1)):
        a = from_nhwc(eng.gated_cab(p, to_nhwc(x), mode), spec.c1)
        if which == "cab1":
            ref = O.cab1(sd, p, x, spec.denoise)
        else:
            ref = O.cab2(sd, p, O.channel_shift(x, which == "cab2_rev", spec.circular), spec.c1, spec.denoise)
        check(a, ref, 3e-3, f"{arch} {which}: many tiles vs oracle")
    eng_planar = gio.pkg("host.engine").Engine(spec, {}, DEV)
    eng_planar.sd = eng.sd
    eng_planar.ln_fuse = False
    xb = to_nhwc(x)
    a = from_nhwc(eng.shift_block("stage1.decoder_level1", xb), spec.c1)
    b = from_nhwc(eng_planar.shift_block("stage1.decoder_level1", xb), spec.c1)
    check(a, b, 5e-4, f"{arch} shift block: fused LayerNorm producers vs ln_planar")
    check(a, O.shift_block(sd, "stage1.decoder_level1", x, O.ARCHS[arch]), 5e-3, f"{arch} shift block (many tiles) vs oracle")


def test_shift_block(aenv):
    sd, spec, eng, gold = aenv
    x = gio.module_input("shift", spec)
    out = from_nhwc(eng.shift_block("stage1.decoder_level1", to_nhwc(x)), spec.c1)
    check(out, O.shift_block(sd, "stage1.decoder_level1", x, O.ARCHS[spec.name]), 5e-3, "shift block vs oracle")
    check(out, torch.from_numpy(gold["block"]), 5e-3, "shift block vs reference golden")


def test_tfr_unet(aenv):
    sd, spec, eng, gold = aenv
    x = gio.module_input("tfr", spec)
    out = from_nhwc(eng.tfr_unet("orb1", to_nhwc(x)), spec.n0)
    check(out, O.tfr_unet(sd, "orb1", x, O.ARCHS[spec.name]), 5e-3, "TFR_UNet vs oracle")
    check(out, torch.from_numpy(gold["tfr"]), 5e-3, "TFR_UNet vs reference golden")


def test_stage1(aenv):
    sd, spec, eng, gold = aenv
    x = gio.module_input("stage1", spec)
    out = from_nhwc(eng.stage1("stage1", to_nhwc(x)), spec.n0)
    check(out, torch.from_numpy(gold["stage1"]), 1e-2, "stage1 (Encoder2) vs reference golden")


def _net(sd, dtype=torch.float16, arch="gshift_deblur2"):
    import importlib
    GShiftNet = importlib.import_module("basicsr.models.archs." + arch).GShiftNet
    net = GShiftNet(future_frames=2, past_frames=2)
    net.load_state_dict(sd)
    net = net.to(DEV).eval()
    return net.half() if dtype == torch.float16 else net


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_full_forward_golden(aenv, dtype):
    """The drop-in call of inference/test_deblur_small.py:84-89,134 (test_denoise_small.py:83-88,162) on the golden clip."""
    sd, spec, _, gold = aenv
    x, nm = gio.clip_input(spec)
    net = _net(sd, dtype, spec.name)
    out = net(x.to(DEV, dtype), nm.to(DEV, dtype)) if spec.denoise else net(x.to(DEV, dtype))
    assert out.dtype == dtype and tuple(out.shape) == (2, 3, 32, 40)
    ref = torch.from_numpy(gold["full"])
    p = O.psnr(out.float().cpu(), ref)
    print(f"[parity] {spec.name} full forward {dtype}: PSNR vs reference fp32 = {p:.2f} dB")
    assert p >= 60.0          # SURVEY.md section 8(c) contract; measured 68-84 dB


def test_full_forward_k1_config_vs_oracle(env):
    """BASELINE config K1 (1x8x3x256x256): PSNR vs the fp32 oracle and the relative PSNR-vs-GT drift (<= 1e-3)."""
    sd, spec, _, _ = env
    gt, x = gio.pkg("host.synth").synthetic_clip(8, 256, 256)
    ref = O.gshiftnet_forward(sd, O.ARCHS[spec.name], x)
    out = _net(sd)(x.to(DEV).half()).float().cpu()
    p = O.psnr(out, ref)
    pg_ref, pg_out = O.psnr(ref.clamp(0, 1), gt[0, 2:-2]), O.psnr(out.clamp(0, 1), gt[0, 2:-2])
    drift = abs(pg_out - pg_ref) / pg_ref
    print(f"[parity] K1 256x256 T=8: PSNR(cuda, oracle)={p:.2f} dB ; PSNR-vs-GT ref={pg_ref:.4f} ours={pg_out:.4f} drift={drift:.2e}")
    assert p >= 60.0 and drift <= 1e-3


def test_full_size_cyclic_frame_equivariance(env):
    """Size-independent property at 720p: with the circular temporal roll of Ours-s every op is equivariant to a cyclic
    shift of the frames, so forward(roll(x, 1))[i] == forward(x)[i-1] bit-exactly on the overlapping output frames."""
    sd, spec, _, _ = env
    net = _net(sd)
    g = torch.Generator().manual_seed(3)
    T, H, W = 7, 720, 1280
    x = torch.rand(1, T, 3, H, W, generator=g).to(DEV).half()
    a = net(x)
    b = net(torch.roll(x, 1, dims=1))
    torch.cuda.synchronize()
    assert torch.isfinite(a).all()
    # a[i] is frame i+2 of x ; b[i] is frame i+2 of roll(x) = frame i+1 of x  => b[i+1] == a[i]
    assert torch.equal(b[1:], a[:-1])
    # and the run is deterministic
    assert torch.equal(net(x), a)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_device_io_u8_clip_and_psnr(dtype):
    """gsn_u8_to_clip is bit-exact against the reference's numpy2tensor arithmetic; gsn_psnr_sse reproduces the float64 PSNR of
    clamp(out,0,1)*255 (float32, unrounded) against the uint8 ground truth (ragged size, values outside [0,1])."""
    infer = gio.pkg("host.infer")
    dio = infer.DeviceIO(torch.device(DEV))
    g = np.random.default_rng(3)
    T, H, W = 3, 37, 53
    frames = [g.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(T)]
    frames[0][:2] = 255; frames[0][2:4] = 0
    clip = dio.clip_from_u8(dio.upload_u8(frames), dtype)
    ref = torch.from_numpy(np.stack(frames)).permute(0, 3, 1, 2).float().mul_(1.0 / 255).to(dtype)
    assert clip.shape == (1, T, 3, H, W) and torch.equal(clip[0].cpu(), ref)
    out = (ref.float() + 0.05 * torch.from_numpy(g.standard_normal(ref.shape).astype(np.float32))).to(dtype)   # some values < 0 and > 1
    gts = [g.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(T)]
    got = dio.psnr(out.to(DEV), dio.upload_u8(gts))
    imgs = (out.float().clamp(0, 1.0).permute(0, 2, 3, 1).numpy() * 255)
    for e in range(T):
        want = infer.psnr_255(imgs[e], gts[e])
        assert abs(got[e] - want) <= 1e-9 * abs(want), (e, got[e], want)
    same = dio.psnr(torch.from_numpy(np.stack(gts)).permute(0, 3, 1, 2).float().div(255).to(DEV), dio.upload_u8(gts))
    assert all(p > 120 or p == float("inf") for p in same)     # float32 x/255*255 is exact or off by one ulp


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_device_ssim_matches_scipy(dtype):
    """gsn_ssim against the reference's ssim_calculate (scipy gaussian_filter over the (3,H,W) array, inference/test_deblur_small.py:
    25-49) to 1e-6, on ragged sizes incl. one narrower than the 13-tap window, values outside [0,1]."""
    infer = gio.pkg("host.infer")
    dio = infer.DeviceIO(torch.device(DEV))
    g = np.random.default_rng(5)
    for T, H, W in ((3, 37, 53), (2, 64, 96), (1, 9, 7)):
        gts = [g.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(T)]
        ref = torch.from_numpy(np.stack(gts)).permute(0, 3, 1, 2).float().div(255)
        out = (ref + 0.08 * torch.from_numpy(g.standard_normal(ref.shape).astype(np.float32))).to(dtype)
        got = dio.ssim(out.to(DEV), dio.upload_u8(gts))
        imgs = (out.float().clamp(0, 1.0).permute(0, 2, 3, 1).numpy() * 255)
        for e in range(T):
            want = infer.ssim_calculate(imgs[e], gts[e])
            assert abs(got[e] - want) <= 1e-6, (T, H, W, e, got[e], want)


def test_denoise_entry_point_device_io_matches_host_io(tmp_path):
    """The denoise entry point with the device I/O path (uint8 H2D, PSNR and SSIM on the GPU) and with --cpu_io print the same
    metrics (same seeded host noise on both paths)."""
    import subprocess
    outs = []
    for extra in ([], ["--cpu_io"]):
        r = subprocess.run([sys.executable, os.path.join(gio.ROOT, "inference", "test_denoise_small.py"), "--synthetic", "1",
                            "--sigma", "30", "--synthetic_frames", "9", "--synthetic_h", "96", "--synthetic_w", "128",
                            "--result_path", str(tmp_path)] + extra, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([l for l in r.stdout.splitlines() if l.startswith("# ")])
    assert outs[0] == outs[1] and len(outs[0]) == 3, outs


def test_inference_entry_point_device_io_matches_host_io(tmp_path):
    """The deblur entry point with the device I/O path (default) and with --cpu_io print the same metrics."""
    import subprocess
    outs = []
    for extra in ([], ["--cpu_io"]):
        r = subprocess.run([sys.executable, os.path.join(gio.ROOT, "inference", "test_deblur_small.py"), "--synthetic", "1",
                            "--one_len", "4", "--synthetic_frames", "8", "--result_path", str(tmp_path)] + extra,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([l for l in r.stdout.splitlines() if l.startswith("# ")])
    assert outs[0] == outs[1] and len(outs[0]) == 2, outs


def test_denoise_entry_point_synthetic(tmp_path):
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(gio.ROOT, "inference", "test_denoise_small.py"), "--synthetic", "1",
                        "--sigma", "30", "--synthetic_frames", "9", "--synthetic_h", "96", "--synthetic_w", "128",
                        "--result_path", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert sum(l.startswith("# Total AVG-PSNR=") for l in r.stdout.splitlines()) == 2, r.stdout[-2000:]


def test_inference_entry_point_synthetic(tmp_path):
    """The reference's CLI (inference/test_deblur_small.py) end to end on generated videos, single rank."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(gio.ROOT, "inference", "test_deblur_small.py"), "--synthetic", "2",
                        "--one_len", "4", "--synthetic_frames", "12", "--result_path", str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("# ")]
    assert len(lines) == 3 and lines[-1].startswith("# Total AVG-PSNR="), r.stdout[-2000:]
    assert any(f.startswith("inference_log_") for f in os.listdir(tmp_path))
