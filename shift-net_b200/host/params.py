"""Parameter tree of the four GShiftNet variants.

The drop-in contract is ``net.load_state_dict(torch.load(p)['params'])``
with the *reference's* key names and shapes (inference/test_deblur_small.py:85).  The
classes below therefore only *own parameters* under the reference's attribute names --
they carry no forward logic (the forward lives in ``engine.py`` and runs on our CUDA
kernels).  Constructor citations are to /root/reference/basicsr/models/archs/.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class ConvP(nn.Module):
    """Weight (+bias) holder with nn.Conv2d's shapes and default init distribution."""

    def __init__(self, cin, cout, k, bias, groups=1):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin // groups, k, k))
        self.bias = nn.Parameter(torch.empty(cout)) if bias else None
        bound = 1.0 / math.sqrt(cin // groups * k * k)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if bias:
                self.bias.uniform_(-bound, bound)


class Empty(nn.Module):
    """Parameter-free placeholder keeping nn.Sequential indices aligned with the reference."""


def seq(*mods):
    return nn.Sequential(*mods)


class ChannelAttnP(nn.Module):
    """CALayer / CALayer2 (gshift_deblur2.py:54-89): conv_du = [1x1, ReLU, 1x1, Sigmoid]."""

    def __init__(self, c, reduction):
        super().__init__()
        self.conv_du = seq(ConvP(c, c // reduction, 1, False), Empty(), ConvP(c // reduction, c, 1, False), Empty())


class CABP(nn.Module):
    """CAB / Shift_CAB (gshift_deblur2.py:143-158, gshift_denoise1.py:157-186)."""

    def __init__(self, c, reduction, act):
        super().__init__()
        self.CA = ChannelAttnP(c, reduction)
        self.body = seq(ConvP(c, c, 3, False), act, ConvP(c, c, 3, False))


class RepConvP(nn.Module):
    def __init__(self, c, group_size):
        super().__init__()
        g = c // group_size
        self.conv_1 = ConvP(c, c, 5, False, groups=g)
        self.conv_2 = ConvP(c, c, 3, False, groups=g)


class RepConv2P(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv_2 = ConvP(c, c, 3, False, groups=c)


class LayerNormP(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


class GatedCABP(nn.Module):
    """CAB1 (add=0) / CAB2 (add=C/2)  (gshift_deblur2.py:186-258, gshift_denoise2.py:183-251)."""

    def __init__(self, c, add, ca_reduction, group_size, denoise):
        super().__init__()
        if add:
            self.conv1 = ConvP(add, add, 3, False, groups=add)
        self.norm = LayerNormP(c + add)
        body = [ConvP(c + add, 2 * c, 1, False), RepConv2P(2 * c), Empty()]
        if denoise:
            body.append(ChannelAttnP(c, ca_reduction))
        body += [RepConvP(c, group_size), ConvP(c, 2 * c, 1, False), Empty(),
                 ChannelAttnP(c, ca_reduction), ConvP(c, c, 1, denoise)]
        self.body = seq(*body)
        self.beta = nn.Parameter(torch.zeros(1, c, 1, 1))


_PAIRS = ["encoder_level1"] + [f"encoder_level1_{i}" for i in range(1, 8)]


class ShiftBlockP(nn.Module):
    """Encoder_shift_block (gshift_deblur2.py:443-458, gshift_deblur1.py:440-463)."""

    def __init__(self, c, pairs, ca_reduction, group_size, denoise):
        super().__init__()
        for n in _PAIRS[:pairs]:
            setattr(self, n, seq(GatedCABP(c, c // 2, ca_reduction, group_size, denoise),
                                 GatedCABP(c, 0, ca_reduction, group_size, denoise)))


class DownSampleP(nn.Module):
    """DownSample (gshift_deblur2.py:333-343 ; gshift_denoise2.py:326-335)."""

    def __init__(self, c, s, denoise):
        super().__init__()
        self.down = seq(ConvP(c, c + s, 3, False), nn.PReLU()) if denoise else ConvP(c, c + s, 3, True)


class SkipUpSampleP(nn.Module):
    """SkipUpSample (gshift_deblur2.py:344-353)."""

    def __init__(self, c, s):
        super().__init__()
        self.up = seq(Empty(), ConvP(c + s, c, 1, False))


class PixelShufflePackP(nn.Module):
    """PixelShufflePack (gshift_deblur2.py:259-281)."""

    def __init__(self, cin, cout, scale):
        super().__init__()
        self.upsample_conv = ConvP(cin, cout * scale * scale, 3, True)


class TFRUNetP(nn.Module):
    """TFR_UNet (gshift_deblur2.py:654-681)."""

    def __init__(self, n0, step, red, denoise):
        super().__init__()
        act = nn.PReLU()
        mk = lambda c, n: seq(*[CABP(c, red(c), act) for _ in range(n)])
        self.encoder_level1 = mk(n0, 1)
        self.encoder_level2 = mk(n0 + step, 3)
        self.encoder_level3 = mk(n0 + 2 * step, 3)
        self.down12 = DownSampleP(n0, step, denoise)
        self.down23 = DownSampleP(n0 + step, step, denoise)
        self.decoder_level1 = mk(n0, 1)
        self.decoder_level2 = mk(n0 + step, 3)
        self.decoder_level3 = mk(n0 + 2 * step, 3)
        self.skip_attn1 = CABP(n0, red(n0), act)
        self.skip_attn2 = CABP(n0 + step, red(n0 + step), act)
        self.up21 = SkipUpSampleP(n0, step)
        self.up32 = SkipUpSampleP(n0 + step, step)


class Stage1P(nn.Module):
    """Encoder2 (gshift_deblur2.py:531-577, gshift_deblur1.py:548-604, gshift_denoise1.py:573-633)."""

    def __init__(self, spec, red):
        super().__init__()
        c, n0 = spec.c1, spec.n0
        act = nn.PReLU()
        self.act = act
        gs = 8 if spec.plus else 1
        blk = lambda: ShiftBlockP(c, spec.pairs, red(c), gs, spec.denoise)
        cabc = lambda w: CABP(w, red(w), act)
        if not spec.plus:
            for n in ("encoder_level1", "encoder_level1_1", "encoder_level1_2",
                      "encoder_level2", "encoder_level2_1", "encoder_level2_2"):
                setattr(self, n, blk())
        else:
            if spec.denoise:
                self.encoder_level0 = cabc(n0)
                self.encoder_level0_1 = cabc(n0)
            for n in ("encoder_level1", "encoder_level1_1", "encoder_level2", "encoder_level2_1",
                      "encoder_level3", "encoder_level3_1"):
                setattr(self, n, cabc(c))
        self.concat = cabc(n0)
        self.down01 = seq(ConvP(n0, c, 2, False), nn.PReLU())
        self.down12 = DownSampleP(c, 0, spec.denoise)
        if spec.plus:
            self.down23 = DownSampleP(c, 0, spec.denoise)
        names = ["decoder_level1", "decoder_level1_1", "decoder_level1_2", "decoder_level2", "decoder_level2_1"]
        names += ["decoder_level3", "decoder_level3_1"] if spec.plus else ["decoder_level2_2"]
        for n in names:
            setattr(self, n, blk())
        self.skip_attn1 = cabc(c)
        if spec.plus:
            self.skip_attn2 = cabc(c)
        self.upsample0 = PixelShufflePackP(c, n0, 2)
        self.skip_conv = cabc(n0)
        self.out_conv = cabc(n0)
        small_deblur = not spec.plus and not spec.denoise
        self.conv_hr0 = ConvP(n0 if small_deblur else 2 * n0, n0, 3, not small_deblur)
        self.up21 = SkipUpSampleP(c, 0)
        if spec.plus:
            self.up32 = SkipUpSampleP(c, 0)


def build_param_tree(net: nn.Module, spec) -> None:
    """Attach the reference-named parameter tree to ``net`` (GShiftNet.__init__ of each arch:
    gshift_deblur2.py:701-730, gshift_deblur1.py:728-760, gshift_denoise2.py:697-727,
    gshift_denoise1.py:758-787)."""
    n0 = spec.n0
    # Ours-s deblur hard-wires the CA bottleneck to "reduction = 1" (gshift_deblur2.py:60,78);
    # the other three files use the passed reduction=4.
    forced = (not spec.plus) and (not spec.denoise)
    red = (lambda c: 1) if forced else (lambda c: 4)
    in_ch = 4 if spec.denoise else 3
    net.feat_extract = seq(ConvP(in_ch, n0, 3, True), CABP(n0, red(n0), nn.PReLU()))
    net.conv_last = ConvP(n0, 3, 3 if spec.denoise else 5, False)
    net.conv_trans = ConvP(n0, n0, 3, True)
    net.lrelu = nn.PReLU()
    net.stage1 = Stage1P(spec, red)
    for pre in ("orb", "rorb"):
        for i in range(1, 6):
            setattr(net, f"{pre}{i}", TFRUNetP(n0, spec.unet_step, red, spec.denoise))
    net.rconcat = ConvP(3 * n0, n0, 3, not spec.denoise)
