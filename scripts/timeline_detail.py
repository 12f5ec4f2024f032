"""Per-launch-shape CUDA-event timing of one forward (GSN_TIMELINE_DETAIL=1).  Run on the GPU box:
    python scripts/timeline_detail.py [arch] [T] [H] [W]"""
import importlib
import os
import sys

os.environ["GSN_TIMELINE_DETAIL"] = "1"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_io as gio

arch = sys.argv[1] if len(sys.argv) > 1 else "gshift_deblur2"
T, H, W = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (20, 720, 1280)
sd, spec = gio.synthetic_checkpoint(arch)
net = importlib.import_module("basicsr.models.archs." + arch).GShiftNet(future_frames=2, past_frames=2)
net.load_state_dict(sd)
net = net.half().cuda().eval()
res = gio.pkg("host.synth").synthetic_clip(T, H, W, seed=7, **({"denoise_sigma": 30} if spec.denoise else {}))
x = res[1].half().cuda()
args = (x,) if not spec.denoise else (x, res[2].half().cuda())
for _ in range(2):
    net(*args)
eng = net.engine()
eng.timeline = []
net(*args)
torch.cuda.synchronize()
tl, eng.timeline = eng.timeline, None
agg = {}
for name, pixels, a, b in tl:
    d = agg.setdefault(name, [0, 0.0, 0])
    d[0] += pixels; d[1] += a.elapsed_time(b); d[2] += 1
tot = sum(v[1] for v in agg.values())
print(f"{arch} T={T} {H}x{W}: instrumented total {tot:.1f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} n={v[2]:4d} total={v[1]:8.2f} ms  avg={1e3 * v[1] / v[2]:8.1f} us  {v[0] / v[2] / (1e3 * v[1] / v[2]) / 1e3:7.2f} Gpx/s  share={100 * v[1] / tot:5.1f}%")
