"""GPU parity, part 2 (-m gpu):

* BIT-EXACT tests of the integer/index part of the path -- the grouped spatial shift gather and the half-channel temporal roll
  (gshift_deblur2.py:465-519) -- through the kernels that carry them (gsn_shift_conv1 with a delta conv1, pass B with a zero
  weight, gsn_roll_copy), against the oracle's index maps (which tests/test_oracle_cpu.py pins to the reference by sha256);
* scale-placement tests with NON-UNIFORM channel-attention scales (grouped RepConv of Ours+ denoise);
* the whole net against the oracle at the sizes bench.py measures: BASELINE configs K2 (Ours-s 720p T=20), K3 (Ours+ 720p T=52),
  one K4 tile (denoise, 68 x 272 x 448) and a K5-shaped Ours+ 1080p clip.  The checker there is the oracle run in fp32 on the GPU
  (TF32 off), itself pinned to the CPU oracle at K1 in this file.  Contract: PSNR(ours, oracle fp32)
  >= 60 dB on [0,1] outputs and |PSNR(ours,gt) - PSNR(oracle,gt)| / PSNR(oracle,gt) <= 1e-3.
"""
import ctypes as C
import importlib
import os
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

import golden_io as gio

sys.path.insert(0, os.path.join(gio.ROOT, "oracle"))
import shiftnet_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda:0"
L = gio.pkg("host.lib")
P = gio.pkg("host.packing")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def nhwc16(x):
    """(T,C,H,W) fp32 cpu -> (T,H,W,C) fp16 cuda (C % 8 == 0 here)."""
    return x.permute(0, 2, 3, 1).contiguous().to(DEV).half()


def int_valued(T, Cc, H, W, seed):
    """fp16-exact, all-distinct-ish values: small integers / 8 (any index mix-up changes the result)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-1024, 1024, (T, Cc, H, W), generator=g).float() / 8.0


# ------------------------------------------------------------------------------------------------ exact index maps
@pytest.mark.parametrize("Cc", [64, 80])
@pytest.mark.parametrize("circular", [True, False])
@pytest.mark.parametrize("shape", [(1, 12, 20), (2, 9, 33), (5, 37, 50)])
def test_shift_gather_index_map_bit_exact(Cc, circular, shape):
    """gsn_shift_conv1 with conv1 = delta (centre tap 1, others 0) must return spatial_shift2(neighbour half) EXACTLY:
    24 offsets, zero fill, the circular / clamped neighbour frame (gshift_deblur2.py:465-519, gshift_deblur1.py:504-528)."""
    lib = L.load()
    T, H, W = shape
    x = int_valued(T, Cc, H, W, 100 + Cc + T)
    xd = nhwc16(x)
    wc1 = torch.zeros(9, Cc // 2, dtype=torch.float16, device=DEV)
    wc1[4] = 1.0
    for rev in (False, True):
        out = torch.empty(T, H, W, Cc // 2, dtype=torch.float16, device=DEV)
        L.check(lib.gsn_shift_conv1(xd.data_ptr(), T, H, W, Cc, L.MODE_CAB2_REV if rev else L.MODE_CAB2_FWD, int(circular),
                                    wc1.data_ptr(), out.data_ptr(), _stream()), "shift_conv1")
        torch.cuda.synchronize()
        _, hw = O.temporal_roll(x, rev, circular)
        ref = O.spatial_shift(hw, Cc)
        got = out.permute(0, 3, 1, 2).float().cpu()
        assert torch.equal(got, ref), f"C={Cc} rev={rev} circular={circular} {shape}: {(got != ref).sum().item()} elements differ"


@pytest.mark.parametrize("Cc", [64, 80])
@pytest.mark.parametrize("circular", [True, False])
def test_temporal_roll_shortcut_bit_exact(Cc, circular):
    """Pass B with a zero effective weight returns its shortcut: the rolled stream for the CAB2 modes (gshift_deblur2.py:253,257),
    x itself for CAB1 -- bit-exact against the oracle's temporal_roll, C=64 (tcgen05 streaming kernel) and C=80."""
    lib = L.load()
    for T, H, W in ((1, 8, 16), (2, 9, 33), (5, 40, 52)):
        x = int_valued(T, Cc, H, W, 7 + T)
        xd = nhwc16(x)
        z = torch.randn(T, H, W, Cc, device=DEV).half()
        weff = torch.zeros(T, Cc * Cc, dtype=torch.float16, device=DEV)
        beff = torch.zeros(T, Cc, dtype=torch.float32, device=DEV)
        for mode, rev in ((L.MODE_CAB2_FWD, False), (L.MODE_CAB2_REV, True), (L.MODE_CAB1, None)):
            out = torch.empty(T, H, W, Cc, dtype=torch.float16, device=DEV)
            b = L.CabPassB()
            b.T, b.H, b.W, b.C, b.mode, b.circular = T, H, W, Cc, mode, int(circular)
            b.x, b.z, b.weff, b.beff, b.out = xd.data_ptr(), z.data_ptr(), weff.data_ptr(), beff.data_ptr(), out.data_ptr()
            L.check(lib.gsn_cab_pass_b(C.byref(b), _stream()), "cab_pass_b")
            torch.cuda.synchronize()
            ref = x if rev is None else O.temporal_roll(x, rev, circular)[0]
            got = out.permute(0, 3, 1, 2).float().cpu()
            assert torch.equal(got, ref), f"C={Cc} mode={mode} circular={circular} T={T}: {(got != ref).sum().item()} differ"


def test_roll_copy_bit_exact():
    """Shift_CAB.channel_shift of Ours+ denoise (gshift_denoise1.py:167-179): clamped roll, real channels 24 of 24 / 80 of 80."""
    lib = L.load()
    for Cc, T, H, W in ((24, 3, 10, 12), (80, 4, 9, 16), (24, 1, 6, 8)):
        x = int_valued(T, Cc, H, W, Cc + T)
        xd = nhwc16(x)
        for rev in (False, True):
            y = torch.empty_like(xd)
            L.check(lib.gsn_roll_copy(xd.data_ptr(), y.data_ptr(), T, H, W, Cc, Cc, int(rev), _stream()), "roll_copy")
            torch.cuda.synchronize()
            assert torch.equal(y.permute(0, 3, 1, 2).float().cpu(), O.temporal_roll(x, rev, False)[0])


# ------------------------------------------------------------------------------------------------ scale placement
def test_group_conv5_nonuniform_scale():
    """gsn_group_conv5 with a per-channel scale spanning 0.05..0.95 must equal RepConv(s * g) (gshift_denoise1.py:190-191,
    224-225): conv5x5(s g) + conv3x3(s g) + s g with groups of 8 -- the scale sits on the INPUT of the grouped conv."""
    lib = L.load()
    g = torch.Generator().manual_seed(21)
    T, Cc, H, W = 2, 80, 37, 45
    x = torch.randn(T, Cc, H, W, generator=g)
    w5 = torch.randn(Cc, 8, 5, 5, generator=g) / 200 ** 0.5
    w3 = torch.randn(Cc, 8, 3, 3, generator=g) / 72 ** 0.5
    s = 0.05 + 0.9 * torch.rand(T, Cc, generator=g)
    wfrag = P.pack_group_conv5(w5, w3).to(DEV)
    xd = nhwc16(x)
    sd_ = s.to(DEV).contiguous()
    for scale in (None, sd_):
        out = torch.empty(T, H, W, Cc, dtype=torch.float16, device=DEV)
        L.check(lib.gsn_group_conv5(xd.data_ptr(), T, H, W, Cc, wfrag.data_ptr(), scale.data_ptr() if scale is not None else None,
                                    out.data_ptr(), _stream()), "group_conv5")
        torch.cuda.synchronize()
        xin = x.half().float()
        if scale is not None:
            xin = (xin * s.view(T, Cc, 1, 1)).half().float()          # the kernel rounds the scaled tile to fp16
        wm = w5.clone()
        wm[:, :, 1:4, 1:4] += w3                                        # host/packing.py merges the two tap sets in fp32 -> fp16
        wm = wm.half().float()
        ref = F.conv2d(xin, wm, None, padding=2, groups=Cc // 8) + xin
        got = out.permute(0, 3, 1, 2).float().cpu()
        r = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
        print(f"[parity] group_conv5 scale={'yes' if scale is not None else 'no'}: rel_rms={r:.2e}")
        assert r < 1.5e-3
        if scale is not None:     # and the old (wrong) placement must be distinguishable at this tolerance
            wrong = (F.conv2d(x.half().float(), wm, None, padding=2, groups=Cc // 8) + x.half().float()) * s.view(T, Cc, 1, 1)
            assert ((wrong - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item() > 0.05


@pytest.mark.parametrize("mode_name", ["cab1", "cab2_fwd", "cab2_rev"])
def test_ln_pw_tcgen05_vs_mma_sync_and_torch(mode_name):
    """gsn_ln_pw_tc (TMA + tcgen05, LayerNorm folded around the GEMM) against gsn_ln_pw (mma.sync, LayerNorm in shared memory) and
    against torch fp32: W . LayerNorm([rolled stream | conv1(shifted half)]) for C = 80, clamped roll, ragged pixel count."""
    lib = L.load()
    mode = {"cab1": L.MODE_CAB1, "cab2_fwd": L.MODE_CAB2_FWD, "cab2_rev": L.MODE_CAB2_REV}[mode_name]
    g = torch.Generator().manual_seed(31)
    T, Cc, H, W = 3, 80, 37, 45
    cin = Cc if mode == L.MODE_CAB1 else Cc + Cc // 2
    x = 0.5 * torch.randn(T, Cc, H, W, generator=g) + 0.3
    hwp = 0.5 * torch.randn(T, Cc // 2, H, W, generator=g)
    w1 = torch.randn(2 * Cc, cin, generator=g) / cin ** 0.5
    gamma, beta = 1 + 0.1 * torch.randn(cin, generator=g), 0.1 * torch.randn(cin, generator=g)
    xd, hd = nhwc16(x), nhwc16(hwp)
    wfold, wvec = (t.to(DEV) for t in P.pack_ln_pw_tc(w1, gamma, beta))
    kpad = (cin + 15) // 16 * 16
    w1z = torch.zeros(2 * Cc, kpad)
    w1z[:, :cin] = w1
    w1p = P.planar_chunks(w1z).contiguous().to(DEV)
    ln = torch.cat((gamma, beta)).to(DEV)
    outs = []
    for tc in (True, False):
        ga = torch.empty(T, H, W, Cc, dtype=torch.float16, device=DEV)
        gb = torch.empty_like(ga)
        hp = hd.data_ptr() if mode != L.MODE_CAB1 else None
        if tc:
            L.check(lib.gsn_ln_pw_tc(xd.data_ptr(), hp, T, H, W, Cc, mode, 0, wfold.data_ptr(), wvec.data_ptr(), ga.data_ptr(), gb.data_ptr(), _stream()), "ln_pw_tc")
        else:
            L.check(lib.gsn_ln_pw(xd.data_ptr(), hp, T, H, W, Cc, mode, 0, ln.data_ptr(), w1p.data_ptr(), ga.data_ptr(), gb.data_ptr(), _stream()), "ln_pw")
        torch.cuda.synchronize()
        outs.append(torch.cat((ga, gb), -1).permute(0, 3, 1, 2).float().cpu())
    xq, hq = x.half().float(), hwp.half().float()
    if mode == L.MODE_CAB1:
        a = xq
    else:
        a = torch.cat((O.temporal_roll(xq, mode == L.MODE_CAB2_REV, False)[0], hq), 1)
    ref = F.conv2d(O.layer_norm2d(a, gamma, beta), w1.view(2 * Cc, cin, 1, 1))
    for name, o in zip(("tcgen05", "mma.sync"), outs):
        r = ((o - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
        print(f"[parity] ln_pw {mode_name} {name}: rel_rms={r:.2e}")
        assert torch.isfinite(o).all() and r < 2e-3, (name, r)


@pytest.mark.parametrize("case", [(36, 40, True, False), (48, 48, False, True), (36, 40, False, True)], ids=["c36_prelu_bias", "c48_sums", "c36_sums"])
def test_conv3x3_tcgen05_vs_torch_and_mma_sync(case):
    """gsn_conv3x3_tc (implicit GEMM on tcgen05, taps as descriptor offsets) against torch fp32 and against gsn_conv_mma, on ragged sizes
    (partial tiles in x and y, more tiles than SMs), with bias + PReLU and with the per-tile channel sums."""
    c, cp, prelu, sums = case
    sd, spec = gio.synthetic_checkpoint("gshift_deblur1")
    eng = gio.pkg("host.engine").Engine(spec, {}, DEV)
    g = torch.Generator().manual_seed(c)
    for T, H, W in ((2, 37, 45), (3, 90, 160), (1, 5, 7)):
        x = torch.randn(T, c, H, W, generator=g)
        w = torch.randn(c, c, 3, 3, generator=g) / (c * 9) ** 0.5
        key = f"testc3.{c}.{H}"
        eng.sd[key + ".weight"] = w
        b = None
        if prelu:
            b = 0.1 * torch.randn(c, generator=g)
            eng.sd[key + ".bias"] = b
            eng.sd[key + ".slope"] = torch.tensor([0.2])
        xd = torch.zeros(T, H, W, cp, dtype=torch.float16, device=DEV)
        xd[..., :c] = x.permute(0, 2, 3, 1).to(DEV).half()
        ref = F.conv2d(x.half().float(), w.half().float(), b, padding=1)
        if prelu:
            ref = F.prelu(ref, torch.tensor([0.2]))
        res = eng.conv3x3_tc(key, xd, c, prelu_key=key + ".slope" if prelu else None, want_sums=sums)
        out, partial = res if sums else (res, None)
        ref2 = eng.conv(key, [xd], [c], c, prelu_key=key + ".slope" if prelu else None)
        torch.cuda.synchronize()
        got = out[..., :c].permute(0, 3, 1, 2).float().cpu()
        r = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
        r2 = ((got - ref2[..., :c].permute(0, 3, 1, 2).float().cpu()).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
        print(f"[parity] conv3x3_tc c={c} {T}x{H}x{W}: vs torch {r:.2e} ; vs conv_mma {r2:.2e}")
        assert torch.isfinite(out).all() and r < 2e-3 and r2 < 1e-3
        if cp > c:
            assert out[..., c:].abs().max().item() == 0.0, "padding channels must stay zero"
        if sums:
            s_got = partial.sum(1)[:, :c].cpu()
            s_ref = out[..., :c].float().sum((1, 2)).cpu()
            assert torch.allclose(s_got, s_ref, rtol=2e-3, atol=2e-2 * (H * W) ** 0.5), (s_got - s_ref).abs().max()


def test_pw_gate_tcgen05_vs_mma_sync_and_torch():
    """gsn_pw_gate_tc (second 1x1 + SimpleGate2 + channel sums on TMA + tcgen05, half-scale weights + tanh) against gsn_cab_pass_a2
    (mma.sync) and torch fp32, C = 80, pixel counts that are not multiples of the 128-pixel tile."""
    lib = L.load()
    g = torch.Generator().manual_seed(41)
    Cc = 80
    for T, H, W in ((2, 37, 45), (3, 16, 24), (1, 90, 160)):
        u = 0.7 * torch.randn(T, Cc, H, W, generator=g)
        w2 = torch.randn(2 * Cc, Cc, generator=g) / Cc ** 0.5
        ud = nhwc16(u)
        w2p = P.planar_chunks(w2).contiguous().to(DEV)
        w2h = P.planar_chunks(0.5 * w2).contiguous().to(DEV)
        ntl = lib.gsn_cab_tiles_linear(H * W)
        outs = []
        for tc in (True, False):
            z = torch.empty(T, H, W, Cc, dtype=torch.float16, device=DEV)
            part = torch.empty(T, ntl, Cc, dtype=torch.float32, device=DEV)
            if tc:
                L.check(lib.gsn_pw_gate_tc(ud.data_ptr(), w2h.data_ptr(), z.data_ptr(), part.data_ptr(), T, H, W, Cc, _stream()), "pw_gate_tc")
            else:
                L.check(lib.gsn_cab_pass_a2(ud.data_ptr(), w2p.data_ptr(), z.data_ptr(), part.data_ptr(), T, H, W, Cc, 0, _stream()), "cab_pass_a2")
            torch.cuda.synchronize()
            outs.append((z.permute(0, 3, 1, 2).float().cpu(), part.sum(1).cpu()))
        v = F.conv2d(u.half().float(), w2.half().float().view(2 * Cc, Cc, 1, 1))
        ref = v[:, :Cc] * torch.sigmoid(v[:, Cc:])
        for name, (o, sm) in zip(("tcgen05", "mma.sync"), outs):
            r = ((o - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
            print(f"[parity] pw_gate {T}x{H}x{W} {name}: rel_rms={r:.2e}")
            assert torch.isfinite(o).all() and r < 2e-3, (name, r)
            assert torch.allclose(sm, o.sum((2, 3)), rtol=2e-3, atol=2e-2 * (H * W) ** 0.5), (name, (sm - o.sum((2, 3))).abs().max())


def test_conv_in_noise_map_strided_vs_cat():
    """gsn_conv_in_nm reads the noise map through its strides (expand()ed (1,T,1,H,W) view of one scalar, as
    inference/test_denoise_small.py:162 passes it) -- same result as the conv over torch.cat((x, noise_map), 1)."""
    lib = L.load()
    g = torch.Generator().manual_seed(3)
    T, H, W = 3, 20, 36
    x = torch.rand(T, 3, H, W, generator=g)
    w = torch.randn(14, 4, 3, 3, generator=g) / 6.0
    b = torch.randn(14, generator=g) * 0.1
    wi, bi = P.pack_conv_in(w, b, 16)
    wi, bi = wi.to(DEV), bi.to(DEV)
    for tdt, dt in ((torch.float16, L.DTYPE_F16), (torch.float32, L.DTYPE_F32)):
        xd = x.to(DEV, tdt).contiguous()
        std = torch.full((1, 1, 1, 1, 1), 30.0 / 255, device=DEV, dtype=tdt)
        dense = (torch.rand(1, T, 1, H, W, generator=g) * 0.2).to(DEV, tdt)
        for nm5 in (std.expand(1, T, 1, H, W), dense):
            nm = nm5[0]
            st = nm.stride()
            f0 = torch.empty(T, H, W, 16, dtype=torch.float16, device=DEV)
            L.check(lib.gsn_conv_in_nm(xd.data_ptr(), dt, T, 3, H, W, nm.data_ptr(), st[0], st[2], st[3], wi.data_ptr(), bi.data_ptr(),
                                       16, f0.data_ptr(), _stream()), "conv_in_nm")
            torch.cuda.synchronize()
            ref = F.conv2d(torch.cat((xd, nm.expand(T, 1, H, W)), 1).float().cpu(), w, b, padding=1)
            got = f0[..., :14].permute(0, 3, 1, 2).float().cpu()
            r = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
            assert r < 1.5e-3, r
            assert f0[..., 14:].abs().max().item() == 0.0


def test_plus_archs_reject_sizes_not_multiple_of_8():
    """Ours+ halves the resolution three times in stage 1: H % 8 == 4 makes the reference raise in SkipUpSample
    (gshift_deblur1.py:352); the engine raises too instead of reading a mismatched skip tensor."""
    sd, spec = gio.synthetic_checkpoint("gshift_deblur1")
    net = importlib.import_module("basicsr.models.archs.gshift_deblur1").GShiftNet(future_frames=2, past_frames=2)
    net.load_state_dict(sd)
    net = net.half().to(DEV).eval()
    with pytest.raises(ValueError):
        net(torch.rand(1, 6, 3, 36, 40, device=DEV).half())
    out = net(torch.rand(1, 6, 3, 40, 48, device=DEV).half())
    assert tuple(out.shape) == (2, 3, 40, 48)


# ------------------------------------------------------------------------------------------------ parity at benchmark sizes
def cuda_oracle(sd, spec, x, nm=None):
    """The oracle (oracle/shiftnet_oracle.py is device-agnostic torch code) in fp32 on the GPU, TF32 off."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        sdd = {k: v.to(DEV, torch.float32) for k, v in sd.items()}
        out = O.gshiftnet_forward(sdd, O.ARCHS[spec.name], x.to(DEV, torch.float32), None if nm is None else nm.to(DEV, torch.float32))
        torch.cuda.synchronize()
        return out.cpu()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        torch.cuda.empty_cache()


def _ours(sd, spec, x, nm=None):
    net = importlib.import_module("basicsr.models.archs." + spec.name).GShiftNet(future_frames=2, past_frames=2)
    net.load_state_dict(sd)
    net = net.half().to(DEV).eval()
    out = net(x.to(DEV).half(), nm.to(DEV).half()) if spec.denoise else net(x.to(DEV).half())
    out = out.float().cpu()
    del net
    torch.cuda.empty_cache()
    return out


def _contract(out, ref, gt, x, what):
    """>= 60 dB against the fp32 oracle, relative PSNR-vs-GT drift <= 1e-3.  Also reported (and bounded):
    the error relative to what the network actually computes, ||out - ref|| / ||ref - x||."""
    p = O.psnr(out, ref)
    gtc = gt[0, 2:-2]
    pg_ref, pg_out = O.psnr(ref.clamp(0, 1), gtc), O.psnr(out.clamp(0, 1), gtc)
    drift = abs(pg_out - pg_ref) / pg_ref
    resid = ((out - ref).pow(2).mean().sqrt() / (ref - x[0, 2:-2, :3]).pow(2).mean().sqrt()).item()
    print(f"[parity-at-size] {what}: PSNR(cuda, oracle fp32) = {p:.2f} dB ; PSNR-vs-GT oracle {pg_ref:.4f} ours {pg_out:.4f} "
          f"drift {drift:.2e} ; error / net residual = {resid:.2e} ; max|diff| = {(out - ref).abs().max().item():.2e}")
    assert torch.isfinite(out).all()
    # the residual-relative bound is looser than it looks for the denoise nets: their output stays close to the input, so the
    # denominator is small (measured 3e-3 / 8e-3 at 75 dB) -- it is a report, the contract is the two conditions before it
    assert p >= 60.0 and drift <= 1e-3 and resid <= 2e-2, (what, p, drift, resid)


def test_cuda_oracle_pinned_to_cpu_oracle_at_k1():
    """The checker of the size tests: the same oracle code in fp32 on the GPU against the CPU oracle on BASELINE config K1."""
    sd, spec = gio.synthetic_checkpoint("gshift_deblur2")
    gt, x = gio.pkg("host.synth").synthetic_clip(8, 256, 256)
    nthr = torch.get_num_threads()
    torch.set_num_threads(min(16, os.cpu_count() or 1))      # the many small convs oversubscribe a 128-thread host badly
    try:
        cpu = O.gshiftnet_forward(sd, O.ARCHS[spec.name], x)
    finally:
        torch.set_num_threads(nthr)
    gpu = cuda_oracle(sd, spec, x)
    p = O.psnr(gpu, cpu)
    print(f"[parity-at-size] oracle fp32 on CUDA vs on CPU at K1: {p:.2f} dB, max|diff| {(gpu - cpu).abs().max().item():.2e}")
    assert p >= 100.0
    _contract(_ours(sd, spec, x), cpu, gt, x, "K1 Ours-s 256x256 T=8 (vs CPU oracle)")


SIZE_CASES = [
    # name, arch, T, H, W
    ("K2_ours_s_720p_T20", "gshift_deblur2", 20, 720, 1280),
    ("K3_ours_plus_720p_T52", "gshift_deblur1", 52, 720, 1280),
    ("K4tile_denoise_s_272x448_T68", "gshift_denoise2", 68, 272, 448),
    ("K4tile_denoise_plus_272x448_T20", "gshift_denoise1", 20, 272, 448),
    ("K5shape_ours_plus_1080p_T12", "gshift_deblur1", 12, 1080, 1920),
]


@pytest.mark.parametrize("case", SIZE_CASES, ids=[c[0] for c in SIZE_CASES])
def test_full_forward_at_benchmark_size_vs_oracle(case):
    name, arch, T, H, W = case
    sd, spec = gio.synthetic_checkpoint(arch)
    synth = gio.pkg("host.synth")
    if spec.denoise:
        gt, x, nm = synth.synthetic_clip(T, H, W, denoise_sigma=30)
    else:
        (gt, x), nm = synth.synthetic_clip(T, H, W), None
    out = _ours(sd, spec, x, nm)
    ref = cuda_oracle(sd, spec, x, nm)
    _contract(out, ref, gt, x, name)


# ------------------------------------------------------------------------------------------------ T-sharded single clip
def test_tshard_single_rank_equals_plain_forward():
    """host/tshard.py with a ring of ONE rank: the halo behind the clip is the clip's own wrap-around frame, so the T-sharded code
    path (halo buffers, n+1-frame CAB2 launches, prefix views for CAB1, local crop) must reproduce net(x) bit-exactly."""
    sd, spec = gio.synthetic_checkpoint("gshift_deblur2")
    net = importlib.import_module("basicsr.models.archs.gshift_deblur2").GShiftNet(future_frames=2, past_frames=2)
    net.load_state_dict(sd)
    net = net.half().to(DEV).eval()
    _, x = gio.pkg("host.synth").synthetic_clip(7, 96, 128)
    x = x.half().to(DEV)
    ts = gio.pkg("host.tshard").TShard(0, 1, 7)
    a = net.forward_tsharded(x, ts)
    b = net(x)
    assert ts.exchanges == 48 and torch.equal(a, b)


@pytest.mark.parametrize("arch", ["gshift_deblur2", "gshift_denoise2", "gshift_deblur1", "gshift_denoise1"])
def test_tshard_two_ranks_nccl_bit_exact(arch):
    """One clip across 2 GPUs (torchrun, NCCL send/recv halo exchange): gathered output == single-GPU forward, bit for bit -- the
    wrapping roll of Ours-s deblur (ring of ranks) and the clamped roll of the denoise / Ours+ nets (open chain)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(gio.ROOT, "scripts", "tshard_check.py"), "9", "96", "128", arch],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "BIT-EXACT" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.parametrize("arch", ["gshift_denoise2", "gshift_deblur1", "gshift_denoise1"])
def test_tshard_single_rank_clamped_equals_plain_forward(arch):
    """A chain of ONE rank for the clamped-roll nets: both clip ends are its own, no halo is used, and the T-sharded code path must
    reproduce the plain forward bit-exactly."""
    sd, spec = gio.synthetic_checkpoint(arch)
    net = importlib.import_module("basicsr.models.archs." + arch).GShiftNet(future_frames=2, past_frames=2)
    net.load_state_dict(sd)
    net = net.half().to(DEV).eval()
    synth = gio.pkg("host.synth")
    if spec.denoise:
        _, x, nm = synth.synthetic_clip(7, 96, 128, denoise_sigma=30)
        nm = nm.half().to(DEV)
    else:
        (_, x), nm = synth.synthetic_clip(7, 96, 128), None
    x = x.half().to(DEV)
    ts = gio.pkg("host.tshard").TShard(0, 1, 7)
    a = net.forward_tsharded(x, ts, nm)
    b = net(x, nm) if nm is not None else net(x)
    assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------ multi-GPU entry point
def test_inference_entry_point_two_ranks_nccl(tmp_path):
    """inference/test_deblur_small.py under torchrun on 2 GPUs: units shard across ranks, ONE NCCL all_gather of the
    (video, frame, psnr, ssim) records, rank 0 prints the same totals as a single-rank run."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    script = os.path.join(gio.ROOT, "inference", "test_deblur_small.py")
    common = ["--synthetic", "3", "--one_len", "4", "--synthetic_frames", "12"]
    env = dict(os.environ, NCCL_DEBUG="WARN")
    one = subprocess.run([sys.executable, script] + common + ["--result_path", str(tmp_path / "n1")],
                         capture_output=True, text=True, timeout=900, env=env)
    assert one.returncode == 0, one.stderr[-2000:]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", script] + common + ["--result_path", str(tmp_path / "n2")],
                         capture_output=True, text=True, timeout=900, env=env)
    assert two.returncode == 0, two.stderr[-3000:]
    l1 = [l for l in one.stdout.splitlines() if l.startswith("# ")]
    l2 = [l for l in two.stdout.splitlines() if l.startswith("# ")]
    assert l1 == l2 and l1[-1].startswith("# Total AVG-PSNR="), (l1, l2)
    ranks = {l.split("]")[0] for l in two.stdout.splitlines() if l.startswith("> [rank ")}
    assert ranks == {"> [rank 0", "> [rank 1"}, ranks
    out_dir = os.path.join(gio.ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "infer_two_ranks.log"), "w") as f:
            f.write(two.stdout + "\n--- stderr ---\n" + two.stderr[-4000:])
