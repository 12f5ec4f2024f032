"""Seeded synthetic checkpoints and clips (no datasets/checkpoints exist offline).

Weights follow nn.Conv2d's default-init distribution, but every ``beta`` (zero-initialised in
the reference, gshift_deblur2.py:208,243 -- which would switch every CAB1/CAB2 branch off),
LayerNorm affine and PReLU slope is randomised, and the channel-attention logits are given a realistic
spread (non-uniform per-channel scales), so the fused kernels and every fold are actually exercised.
CPU generator => identical values in the build container and on the GPU box.
"""
import math

import torch


def randomize_(net: torch.nn.Module, seed: int = 1234) -> None:
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in net.named_parameters():
            leaf = name.rsplit(".", 1)[-1]
            if leaf == "beta":
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif ".norm." in name or name.startswith("norm."):
                base = 1.0 if leaf == "weight" else 0.0
                p.copy_(base + 0.1 * torch.randn(p.shape, generator=g))
            elif p.dim() == 1 and p.numel() == 1:                      # PReLU slope
                p.copy_(0.25 + 0.05 * torch.randn(p.shape, generator=g))
            elif p.dim() == 4:                                         # conv weight
                bound = 1.0 / math.sqrt(p.shape[1] * p.shape[2] * p.shape[3])
                if ".conv_du.2." in name:
                    # channel attention (CALayer / CALayer2, gshift_deblur2.py:54-89): with default-init weights every
                    # sigmoid sits within 0.02 of 0.5, so a scale applied to the wrong channel or the wrong side of a
                    # conv would be invisible to the parity tests; x20 spreads the per-channel scales over ~0.05..0.95
                    bound *= 20.0
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
            else:                                                      # conv bias
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.1)


def synthetic_clip(T, H, W, seed=7, noise=0.05, denoise_sigma=None):
    """(gt, x[, noise_map]) in [0,1], shapes (1,T,3,H,W).  Smooth random video + degradation."""
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(1 * T, 3, H, W, generator=g)
    gt = torch.nn.functional.avg_pool2d(gt, 5, 1, 2, count_include_pad=False).view(1, T, 3, H, W)
    if denoise_sigma is None:
        x = (gt + noise * torch.randn(gt.shape, generator=g)).clamp(0, 1)
        return gt, x
    s = denoise_sigma / 255.0
    x = gt + s * torch.randn(gt.shape, generator=g)           # unclamped (inference/test_denoise_small.py:146-147)
    return gt, x, torch.full((1, T, 1, H, W), s)
