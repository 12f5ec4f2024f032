"""The SSIM algorithm of csrc/ssim.cu (numpy restatement in tests/ssim_model.py) against the reference's scipy-based
ssim_calculate (inference/test_deblur_small.py:25-49, restated in host/infer.py): tiny, ragged and narrower-than-the-filter images."""
import numpy as np
import pytest

import golden_io as gio
from ssim_model import ssim_model


@pytest.mark.parametrize("shape", [(37, 53), (64, 96), (8, 11), (5, 4)])
def test_ssim_model_matches_scipy(shape):
    infer = gio.pkg("host.infer")
    g = np.random.default_rng(shape[0])
    H, W = shape
    gt = g.integers(0, 256, (H, W, 3), dtype=np.uint8)
    base = np.clip(gt.astype(np.float32) / 255 + 0.1 * g.standard_normal((H, W, 3)).astype(np.float32), 0, 1)
    out = (base * np.float32(255)).astype(np.float32)
    assert abs(infer.ssim_calculate(out, gt) - ssim_model(out, gt)) < 2e-7
