"""Static description of the four reference arch variants (product-side copy).

Citations: /root/reference/basicsr/models/archs/gshift_{deblur,denoise}{1,2}.py.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class ArchSpec:
    name: str
    denoise: bool   # 4-ch input + noise map, mid CALayer2, biased last 1x1 (gshift_denoise2.py:194,199)
    plus: bool      # Ours+ topology: 3-level stage-1, 8 pairs/block, grouped RepConv (gshift_deblur1.py)
    n0: int         # full-res width          (gshift_deblur2.py:709 / gshift_deblur1.py:738)
    c1: int         # stage-1 width           (gshift_deblur2.py:704 / gshift_deblur1.py:733)
    unet_step: int  # TFR_UNet width step     (gshift_deblur2.py:657 / gshift_deblur1.py:684)
    n_orb: int      # TFR_UNets executed per stage (gshift_deblur2.py:731-746 / gshift_deblur1.py:762-781)
    pairs: int      # (shift, CAB2, CAB1) pairs per Encoder_shift_block
    circular: bool  # temporal roll wraps (gshift_deblur2.py:504-505) vs clamped (gshift_deblur1.py:513,517)
    default_ctx: int  # ctor default for future_frames/past_frames (deblur 1, denoise 0)


ARCHS = {
    "gshift_deblur2": ArchSpec("gshift_deblur2", False, False, 14, 64, 4, 3, 4, True, 1),
    "gshift_deblur1": ArchSpec("gshift_deblur1", False, True, 24, 80, 12, 5, 8, False, 1),
    "gshift_denoise2": ArchSpec("gshift_denoise2", True, False, 14, 64, 4, 3, 4, False, 0),
    "gshift_denoise1": ArchSpec("gshift_denoise1", True, True, 24, 80, 12, 5, 8, False, 0),
}
