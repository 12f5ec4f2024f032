"""A few (shift, CAB2, CAB1) pairs at a stage-1 level-1 size (T=20, 360x640) for ncu captures of the block's kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import golden_io as gio  # noqa: E402

torch.set_grad_enabled(False)
arch = sys.argv[1] if len(sys.argv) > 1 else "gshift_deblur2"
T, H, W = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (20, 360, 640)))
sd, spec = gio.synthetic_checkpoint(arch)
eng = gio.pkg("host.engine").Engine(spec, sd, "cuda:0")
g = torch.Generator().manual_seed(5)
x = (0.5 * torch.randn(T, H, W, spec.c1, generator=g)).to("cuda:0").half()
for _ in range(2):
    y = eng.shift_block("stage1.decoder_level1", x)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
