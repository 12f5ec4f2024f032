"""One fused shift block (Ours-s, C=64) + one denoise block on a clip with more 16x16 tiles than SMs, for compute-sanitizer
(racecheck / synccheck / memcheck) runs over the tcgen05 / TMA / mbarrier kernels:

    compute-sanitizer --tool racecheck python scripts/sanitize_block.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import golden_io as gio  # noqa: E402

torch.set_grad_enabled(False)
DEV = "cuda:0"
archs = sys.argv[1:] or ["gshift_deblur2", "gshift_denoise2"]
for arch in archs:
    sd, spec = gio.synthetic_checkpoint(arch)
    eng = gio.pkg("host.engine").Engine(spec, sd, DEV)
    g = torch.Generator().manual_seed(5)
    T, H, W = 4, 96, 112                       # 6 x 7 x 4 = 168 tiles > 148 SMs: the persistent loops and cross-tile hand-offs run
    x = (0.5 * torch.randn(T, H, W, spec.c1, generator=g)).to(DEV).half()
    if spec.plus:
        # Ours+: the tcgen05 pointwise kernels (ln_pw_tc, pw_gate_tc, pass B C=80) sit in the shift block; the tcgen05 3x3 implicit
        # GEMM in the 36-channel CABs of the TFR_UNet's second level
        T, H, W = 3, 64, 80                    # 120 pixel tiles of 128 px x 3 frames > 148 SMs
        x = (0.5 * torch.randn(T, H, W, spec.c1, generator=g)).to(DEV).half()
        c2 = spec.n0 + spec.unet_step
        xc = torch.zeros(T, 2 * H, 2 * W, 40, dtype=torch.float16, device=DEV)
        xc[..., :c2] = (0.5 * torch.randn(T, 2 * H, 2 * W, c2, generator=g)).to(DEV).half()
        yc = eng.cab("orb1.encoder_level2.0", xc, c2)
        torch.cuda.synchronize()
        print(arch, "wide CAB ok", tuple(yc.shape), float(yc.float().abs().mean()))
    y = eng.shift_block("stage1.decoder_level1", x)
    torch.cuda.synchronize()
    print(arch, "shift block ok", tuple(y.shape), float(y.float().abs().mean()))
