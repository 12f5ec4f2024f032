"""T-sharded single-clip mode on N GPUs (torchrun --nproc-per-node N): every rank takes its slice of ONE clip, the halo frames
travel over NCCL send/recv; rank 0 gathers the restored frames and compares them BIT-EXACTLY with its own single-GPU forward of the
whole clip.  Prints halo traffic and the time spent in the exchanges.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/tshard_check.py
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import golden_io as gio  # noqa: E402

torch.set_grad_enabled(False)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))   # a mismatched exchange must fail fast, not hang
T, H, W = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (20, 256, 448)))
ARCH = sys.argv[4] if len(sys.argv) > 4 else "gshift_deblur2"
sd, spec = gio.synthetic_checkpoint(ARCH)
net = importlib.import_module("basicsr.models.archs." + ARCH).GShiftNet(future_frames=2, past_frames=2)
net.load_state_dict(sd)
net = net.half().to(dev).eval()
nm = None
if spec.denoise:
    _, x, nm = gio.pkg("host.synth").synthetic_clip(T, H, W, denoise_sigma=30)
    nm = nm.half().to(dev)
else:
    _, x = gio.pkg("host.synth").synthetic_clip(T, H, W)
x = x.half().to(dev)
ts = gio.pkg("host.tshard").TShard(rank, world, T)
ts.time_exchanges = True
for it in range(3):
    ts.halo_bytes = ts.exchanges = 0
    ts.events = []
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out_local = net.forward_tsharded(x[:, ts.a:ts.b].contiguous(), ts, None if nm is None else nm[:, ts.a:ts.b].contiguous())
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
ms_x = sum(a.elapsed_time(b) for a, b in ts.events)
print(f"[tshard] rank {rank}/{world}: frames [{ts.a},{ts.b}) -> {out_local.shape[0]} restored, {ms:.1f} ms, {ts.exchanges} halo exchanges, "
      f"{ts.halo_bytes / 1e6:.1f} MB sent, {ms_x:.2f} ms in exchanges ({ts.halo_bytes / 1e9 / max(ms_x, 1e-9) * 1e3:.1f} GB/s per direction)", flush=True)
ok = True
if world > 1:
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([out_local.shape[0]], dtype=torch.int64, device=dev))
    m = max(int(s.item()) for s in sizes)
    pad = torch.zeros(m, 3, H, W, dtype=out_local.dtype, device=dev)
    pad[:out_local.shape[0]] = out_local
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    full = torch.cat([b[:int(s.item())] for b, s in zip(bufs, sizes)])
else:
    full = out_local
if rank == 0:
    ref = net(x, nm) if nm is not None else net(x)
    ok = tuple(full.shape) == tuple(ref.shape) and torch.equal(full, ref)
    print(f"[tshard] {ARCH} T={T} {H}x{W} on {world} rank(s): gathered {tuple(full.shape)} vs single-GPU forward: "
          f"{'BIT-EXACT' if ok else 'MISMATCH max|diff| = %.3e' % (full.float() - ref.float()).abs().max().item()}", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
