"""CPU oracle for the Shift-Net forward hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch *functional* restatement (fp32, torch CPU ops) of the
reference's ``GShiftNet.forward`` for the four arch variants.  It is imported only by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs -- never by the product path (``shift-net_b200/``).

Parity status: PINNED.  The reference ships no golden vectors, so the
oracle is pinned against outputs of the reference's own classes, loaded by file path in
the build container and run on seeded synthetic checkpoints/clips; those outputs are
committed under ``tests/golden/`` together with ``tests/golden/make_golden.py``.

Everything here takes a *state_dict with the reference's key names* (``sd``) plus a key
prefix, so the same synthetic checkpoint drives reference, oracle and CUDA path.

Reference citations are ``file:line`` relative to ``/root/reference/basicsr/models/archs``.
Abbreviations: d2 = gshift_deblur2.py (Ours-s deblur), d1 = gshift_deblur1.py (Ours+ deblur),
n2 = gshift_denoise2.py (Ours-s denoise), n1 = gshift_denoise1.py (Ours+ denoise).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# arch table
# --------------------------------------------------------------------------------------


@dataclass(frozen=True)
class ArchSpec:
    name: str
    denoise: bool          # 4-ch input, mid CALayer2, biased last 1x1, PReLU'd DownSample, ...
    plus: bool             # "Ours+" topology (3-level stage1, 8 pairs per block, grouped RepConv)
    n0: int                # full-res width            (d2:709 -> 14, d1:738 -> 24)
    c1: int                # stage-1 width             (d2:704 -> 64, d1:733 -> 80)
    unet_step: int         # TFR_UNet width increment  (d2:657 -> 4,  d1:684 -> 12)
    n_orb: int             # TFR_UNets actually called per stage (d2:731-746 -> 3, d1:762-781 -> 5)
    pairs: int             # (shift, CAB2, CAB1) pairs per Encoder_shift_block (d2:521-530 -> 4, d1:530-547 -> 8)
    circular: bool         # temporal roll wraps (d2:504-505) or is clamped (d1:513,517)


ARCHS = {
    "gshift_deblur2": ArchSpec("gshift_deblur2", False, False, 14, 64, 4, 3, 4, True),
    "gshift_deblur1": ArchSpec("gshift_deblur1", False, True, 24, 80, 12, 5, 8, False),
    "gshift_denoise2": ArchSpec("gshift_denoise2", True, False, 14, 64, 4, 3, 4, False),
    "gshift_denoise1": ArchSpec("gshift_denoise1", True, True, 24, 80, 12, 5, 8, False),
}

# --------------------------------------------------------------------------------------
# index maps of the grouped spatial-temporal shift (exact / integer part of the path)
# --------------------------------------------------------------------------------------

_OUTER = [(8, 8), (8, 4), (8, 0), (8, -4), (8, -8), (-8, 8), (-8, 4), (-8, 0), (-8, -4), (-8, -8),
          (4, 8), (4, -8), (0, 8), (0, -8), (-4, 8), (-4, -8)]
_INNER = [(4, 4), (4, 0), (4, -4), (0, 4), (0, -4), (-4, 4), (-4, 0), (-4, -4)]


def shift_offsets(c_full: int):
    """Per-channel (dy, dx) of ``spatial_shift2`` for a block of width ``c_full``.

    d2:465-498: ``number = c_full // 16`` (d2:449), ``n2 = (number-1)//2`` channels for each of
    the 16 outer-ring offsets, then ``n1 = number - 2*n2`` channels for each of the 8 inner-ring
    offsets; ``out[h, w] = in[h-dy, w-dx]`` with zero fill.  Returns a list of c_full//2 pairs.
    """
    number = c_full // 2 // 8
    n2 = (number - 1) // 2
    n1 = number - 2 * n2
    offs = []
    for o in _OUTER:
        offs += [o] * n2
    for o in _INNER:
        offs += [o] * n1
    assert len(offs) == 8 * number
    return offs


def spatial_shift(hw: torch.Tensor, c_full: int) -> torch.Tensor:
    """Grouped spatial shift with zero fill (d2:465-498), table-driven."""
    offs = shift_offsets(c_full)
    T, Ch, H, W = hw.shape
    assert Ch == len(offs)
    pad = F.pad(hw, (8, 8, 8, 8))
    out = torch.empty_like(hw)
    for c, (dy, dx) in enumerate(offs):
        out[:, c] = pad[:, c, 8 - dy:8 - dy + H, 8 - dx:8 - dx + W]
    return out


def temporal_roll(x: torch.Tensor, reverse: bool, circular: bool):
    """Half-channel +-1 frame roll (d2:499-512 circular; d1:504-519 / n2:491-506 clamped).

    Returns ``(y, hw)``: the rolled residual stream and the C/2 channels that came from the
    neighbour frame (which then go through ``spatial_shift``).
    fwd:  y[t, c] = x[t-1, c+C/2] (c <  C/2) ; x[t, c-C/2] (c >= C/2) ; hw = y[:, :C/2]
    rev:  y[t, c] = x[t, c+C/2]   (c <  C/2) ; x[t+1, c-C/2] (c >= C/2) ; hw = y[:, C/2:]
    clamped variants keep frame 0 (fwd) / frame T-1 (rev) *un-swapped*.
    """
    T, C, H, W = x.shape
    h = C // 2
    lo, hi = x[:, :h], x[:, h:]
    if not reverse:
        y = torch.cat((torch.roll(hi, 1, 0), lo), dim=1)
        if not circular:
            y = torch.cat((x[0:1], y[1:]), dim=0)
        hw = y[:, :h]
    else:
        y = torch.cat((hi, torch.roll(lo, -1, 0)), dim=1)
        if not circular:
            y = torch.cat((y[:-1], x[-1:]), dim=0)
        hw = y[:, h:]
    return y, hw


def channel_shift(x: torch.Tensor, reverse: bool, circular: bool) -> torch.Tensor:
    """``Encoder_shift_block.channel_shift`` (d2:499-519): (T,C,h,w) -> (T,3C/2,h,w)."""
    y, hw = temporal_roll(x, reverse, circular)
    return torch.cat((y, spatial_shift(hw, x.shape[1])), dim=1)


# --------------------------------------------------------------------------------------
# small functional pieces
# --------------------------------------------------------------------------------------


def _conv(sd, p, x, stride=1, pad=None, groups=None):
    w = sd[p + ".weight"]
    b = sd.get(p + ".bias")
    if pad is None:
        pad = w.shape[-1] // 2
    if groups is None:
        groups = x.shape[1] // w.shape[1]
    return F.conv2d(x, w, b, stride=stride, padding=pad, groups=groups)


def layer_norm2d(x, w, b, eps=1e-6):
    """Per-pixel LayerNorm across channels, biased variance (d2:19-28)."""
    mu = x.mean(1, keepdim=True)
    var = (x - mu).pow(2).mean(1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def channel_attention(sd, p, x):
    """CALayer / CALayer2 (d2:54-89): x * sigmoid(W2 relu(W1 mean_hw(x)))."""
    y = x.mean((2, 3), keepdim=True)
    y = torch.relu(_conv(sd, p + ".conv_du.0", y))
    y = torch.sigmoid(_conv(sd, p + ".conv_du.2", y))
    return x * y


def cab(sd, p, x):
    """Dense-3x3 channel-attention block ``CAB`` (d2:143-158); PReLU is ``body.1``."""
    r = _conv(sd, p + ".body.0", x)
    r = F.prelu(r, sd[p + ".body.1.weight"])
    r = _conv(sd, p + ".body.2", r)
    return channel_attention(sd, p + ".CA", r) + x


def shift_cab(sd, p, x, reverse):
    """``Shift_CAB`` of Ours+ denoise (n1:157-186): clamped temporal roll only, then CAB body."""
    y, _ = temporal_roll(x, reverse, circular=False)
    return cab(sd, p, y)


def _gated_body(sd, p, a, mid_ca):
    """The shared NAF-style body of CAB1/CAB2 after the LayerNorm (d2:193-204 / n2:188-199).

    Sequential indices (deblur): 0 1x1 | 1 RepConv2 | 2 gate | 3 RepConv | 4 1x1 | 5 sigmoid-gate |
    6 CALayer2 | 7 1x1.  Denoise inserts a CALayer2 at index 3 and shifts the rest by one.
    """
    k = 1 if mid_ca else 0
    u = _conv(sd, p + ".body.0", a)
    u = _conv(sd, p + ".body.1.conv_2", u) + u                    # RepConv2 (d2:169-177)
    h = u.shape[1] // 2
    g = u[:, :h] * u[:, h:]                                       # SimpleGate (d2:178-181)
    if mid_ca:
        g = channel_attention(sd, p + ".body.3", g)
    rp = p + f".body.{3 + k}"
    g = _conv(sd, rp + ".conv_1", g) + _conv(sd, rp + ".conv_2", g) + g   # RepConv (d2:159-168)
    v = _conv(sd, p + f".body.{4 + k}", g)
    z = v[:, :h] * torch.sigmoid(v[:, h:])                        # SimpleGate2 (d2:182-185)
    z = channel_attention(sd, p + f".body.{6 + k}", z)
    return _conv(sd, p + f".body.{7 + k}", z)


def cab1(sd, p, x, mid_ca):
    """``CAB1`` (d2:186-214): x + body(LN(x)) * beta."""
    a = layer_norm2d(x, sd[p + ".norm.weight"], sd[p + ".norm.bias"])
    return x + _gated_body(sd, p, a, mid_ca) * sd[p + ".beta"]


def cab2(sd, p, x_in, c, mid_ca):
    """``CAB2`` (d2:215-258): input is cat(rolled stream y (C), shifted half (C/2))."""
    shortcut, hw = x_in[:, :c], x_in[:, c:]
    hw = _conv(sd, p + ".conv1", hw)                              # dw3x3 on the shifted half (d2:226,254)
    a = layer_norm2d(torch.cat((shortcut, hw), 1), sd[p + ".norm.weight"], sd[p + ".norm.bias"])
    return shortcut + _gated_body(sd, p, a, mid_ca) * sd[p + ".beta"]


_PAIR_NAMES = ["encoder_level1"] + [f"encoder_level1_{i}" for i in range(1, 8)]


def shift_block(sd, p, x, spec: ArchSpec):
    """``Encoder_shift_block.forward`` (d2:521-530, d1:530-547): alternating fwd/rev pairs."""
    c = x.shape[1]
    for i in range(spec.pairs):
        q = f"{p}.{_PAIR_NAMES[i]}"
        x = channel_shift(x, reverse=bool(i & 1), circular=spec.circular)
        x = cab2(sd, q + ".0", x, c, spec.denoise)
        x = cab1(sd, q + ".1", x, spec.denoise)
    return x


def _seq_cabs(sd, p, x):
    i = 0
    while f"{p}.{i}.body.0.weight" in sd:
        x = cab(sd, f"{p}.{i}", x)
        i += 1
    return x


def downsample(sd, p, x, spec: ArchSpec):
    """``DownSample`` (d2:333-343 biased conv; n2:326-335 unbiased conv + PReLU)."""
    if spec.denoise:
        return F.prelu(_conv(sd, p + ".down.0", x, stride=2, pad=1), sd[p + ".down.1.weight"])
    return _conv(sd, p + ".down", x, stride=2, pad=1)


def skip_upsample(sd, p, x, skip):
    """``SkipUpSample`` (d2:344-353): bilinear x2 (align_corners=False) -> 1x1 -> + skip."""
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    return _conv(sd, p + ".up.1", x) + skip


def tfr_unet(sd, p, x, spec: ArchSpec):
    """``TFR_UNet.forward`` (d2:682-695)."""
    enc1 = _seq_cabs(sd, p + ".encoder_level1", x)
    enc2 = _seq_cabs(sd, p + ".encoder_level2", downsample(sd, p + ".down12", enc1, spec))
    enc3 = _seq_cabs(sd, p + ".encoder_level3", downsample(sd, p + ".down23", enc2, spec))
    dec3 = _seq_cabs(sd, p + ".decoder_level3", enc3)
    x = skip_upsample(sd, p + ".up32", dec3, cab(sd, p + ".skip_attn2", enc2))
    dec2 = _seq_cabs(sd, p + ".decoder_level2", x)
    x = skip_upsample(sd, p + ".up21", dec2, cab(sd, p + ".skip_attn1", enc1))
    return _seq_cabs(sd, p + ".decoder_level1", x)


def pixel_shuffle_pack(sd, p, x):
    """``PixelShufflePack`` (d2:259-281): 3x3 conv -> pixel_shuffle(2)."""
    return F.pixel_shuffle(_conv(sd, p + ".upsample_conv", x), 2)


def stage1(sd, p, x, spec: ArchSpec):
    """``Encoder2.forward``: d2:587-613 (Ours-s), d1:614-643 (Ours+), n1:640-671 (Ours+ denoise)."""
    x = cab(sd, p + ".concat", x)
    shortcut = x
    if spec.plus and spec.denoise:
        x = shift_cab(sd, p + ".encoder_level0", x, False)
        x = shift_cab(sd, p + ".encoder_level0_1", x, True)
    x = F.prelu(_conv(sd, p + ".down01.0", x, stride=2, pad=0), sd[p + ".down01.1.weight"])
    if not spec.plus:
        enc11 = x
        for n in ("encoder_level1", "encoder_level1_1", "encoder_level1_2"):
            enc11 = shift_block(sd, f"{p}.{n}", enc11, spec)
        y = downsample(sd, p + ".down12", enc11, spec)
        for n in ("encoder_level2", "encoder_level2_1", "encoder_level2_2",
                  "decoder_level2", "decoder_level2_1", "decoder_level2_2"):
            y = shift_block(sd, f"{p}.{n}", y, spec)
    else:
        if spec.denoise:
            enc11 = shift_cab(sd, p + ".encoder_level1", x, False)
            enc11 = shift_cab(sd, p + ".encoder_level1_1", enc11, True)
        else:
            enc11 = cab(sd, p + ".encoder_level1_1", cab(sd, p + ".encoder_level1", x))
        y = downsample(sd, p + ".down12", enc11, spec)
        enc22 = cab(sd, p + ".encoder_level2_1", cab(sd, p + ".encoder_level2", y))
        y = downsample(sd, p + ".down23", enc22, spec)
        y = cab(sd, p + ".encoder_level3_1", cab(sd, p + ".encoder_level3", y))
        y = shift_block(sd, p + ".decoder_level3", y, spec)
        y = shift_block(sd, p + ".decoder_level3_1", y, spec)
        y = skip_upsample(sd, p + ".up32", y, cab(sd, p + ".skip_attn2", enc22))
        y = shift_block(sd, p + ".decoder_level2", y, spec)
        y = shift_block(sd, p + ".decoder_level2_1", y, spec)
    y = skip_upsample(sd, p + ".up21", y, cab(sd, p + ".skip_attn1", enc11))
    for n in ("decoder_level1", "decoder_level1_1", "decoder_level1_2"):
        y = shift_block(sd, f"{p}.{n}", y, spec)
    up = pixel_shuffle_pack(sd, p + ".upsample0", y)
    sk = cab(sd, p + ".skip_conv", shortcut)
    if spec.plus or spec.denoise:
        out = _conv(sd, p + ".conv_hr0", torch.cat((up, sk), 1))          # d1:640, n2:607
    else:
        out = _conv(sd, p + ".conv_hr0", F.prelu(up, sd[p + ".act.weight"])) + sk   # d2:611
    return cab(sd, p + ".out_conv", out)


def gshiftnet_forward(sd, spec: ArchSpec, x, noise_map=None, past=2, future=2):
    """``GShiftNet.forward`` (d2:748-756 ; n2:744-753): (1,T,3,H,W) -> (T-past-future,3,H,W)."""
    assert x.shape[0] == 1
    x = x[0]
    T = x.shape[0]
    inp = torch.cat((x, noise_map[0]), 1) if spec.denoise else x
    x0 = cab(sd, "feat_extract.1", _conv(sd, "feat_extract.0", inp))
    # stage0 (d2:731-737 ; n2:729-734)
    f = x0
    for i in range(1, spec.n_orb + 1):
        f = tfr_unet(sd, f"orb{i}", f, spec)
    if not spec.denoise:
        f = f + x0
    sam0, sam = f, _conv(sd, "conv_trans", f)
    dec = stage1(sd, "stage1", sam, spec)
    s = slice(past, T - future)
    # stage2 (d2:738-746 uses sam0 ; n2:735-742 uses sam, PReLU, no shortcut)
    third = sam[s] if spec.denoise else sam0[s]
    y = _conv(sd, "rconcat", torch.cat((x0[s], third, dec[s]), 1))
    if spec.denoise:
        y = F.prelu(y, sd["lrelu.weight"])
    r = y
    for i in range(1, spec.n_orb + 1):
        r = tfr_unet(sd, f"rorb{i}", r, spec)
    if not spec.denoise:
        r = r + y
    return _conv(sd, "conv_last", r) + x[s]


# --------------------------------------------------------------------------------------
# metrics used by the inference scripts (inference/test_deblur_small.py:139-143)
# --------------------------------------------------------------------------------------


def psnr(a: torch.Tensor, b: torch.Tensor, data_range=1.0) -> float:
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    if mse == 0:
        return float("inf")
    import math
    return 10.0 * math.log10(data_range ** 2 / mse)
