"""One Ours+ forward at 720p (T frames) for ncu captures of the Ours+ kernels."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import golden_io as gio  # noqa: E402

torch.set_grad_enabled(False)
os.environ["GSN_CUDA_GRAPH"] = "0"
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sd, spec = gio.synthetic_checkpoint("gshift_deblur1")
net = importlib.import_module("basicsr.models.archs.gshift_deblur1").GShiftNet(future_frames=2, past_frames=2)
net.load_state_dict(sd)
net = net.half().to("cuda:0").eval()
_, x = gio.pkg("host.synth").synthetic_clip(T, 720, 1280)
y = net(x.half().to("cuda:0"))
torch.cuda.synchronize()
print("ok", tuple(y.shape))
