#!/usr/bin/env python
"""bench.py -- frames/sec of the Shift-Net forward hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference algorithm's CPU arm (oracle port on the host cores)

Workload (BASELINE.json configs[1]): Ours-s deblur, synthetic 720p clip, one_len=16 -> input (1,20,3,720,1280),
16 output frames per step.  A "step" is one GShiftNet forward of one clip.  Clips shard across ranks (weak scaling:
one clip per rank per step, no data-path collective; the only collective is the max-reduction of the timing).
"""
import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

ARCH = "gshift_deblur2"
ONE_LEN, CTX, H, W = 16, 2, 720, 1280
T = ONE_LEN + 2 * CTX
METRIC = "frames/sec (720p, one_len=16)"


def pkg(sub):
    return importlib.import_module("shift-net_b200." + sub)


def synthetic_net_and_sd():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_io as gio
    sd, spec = gio.synthetic_checkpoint(ARCH)
    return sd, spec


class ClockSampler(threading.Thread):
    """nvidia-smi clock / throttle-reason sampler running during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_oracle_rate(threads=None, repeats=1):
    """Times the oracle port (torch CPU fp32, all host threads) on a bounded sample of the workload and converts to the
    metric's unit.  Sample: 8 frames of the 720p clip cropped to 256x256 (BASELINE config K1 shape, about 20-30 s of CPU work).  Threads: all host
    cores up to 16 -- the first B200-box run with all 128 hardware threads took 264 s for a forward that 8 threads do in
    23 s (oversubscription of the many small convs), so more threads would only flatter the GPU ratio."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import shiftnet_oracle as O
    sd, spec = synthetic_net_and_sd()
    threads = threads or min(16, os.cpu_count())
    torch.set_num_threads(threads)
    TS, HS = 8, 256
    _, x = pkg("host.synth").synthetic_clip(TS, HS, HS)
    with torch.no_grad():
        O.gshiftnet_forward(sd, O.ARCHS[ARCH], x[:, :5, :, :64, :64])        # warm-up (allocator, oneDNN primitives)
        times = []
        for _ in range(repeats):
            t0 = time.time()
            O.gshiftnet_forward(sd, O.ARCHS[ARCH], x)
            times.append(time.time() - t0)
    sec = sorted(times)[len(times) // 2]
    px_frames_per_s = TS * HS * HS / sec                  # input frame-pixels per second on the CPU
    fps_720p = ONE_LEN / (T * H * W / px_frames_per_s)    # output frames/s the CPU would reach on the K2 clip
    sample = (f"T={TS} {HS}x{HS} crop of the clip ({sec:.1f} s per forward, median of {repeats}, {threads} threads of {os.cpu_count()}); scaled to the 720p one_len=16 clip "
              "by input frame-pixels (extrapolated)")
    return fps_720p, threads, sample, sec


def run_reference(args, rank, world):
    if rank != 0:
        return
    vals = []
    for _ in range(max(1, args.warmup > 0)):
        pass
    steps = max(1, args.steps)
    fps, threads, sample, sec = cpu_oracle_rate(repeats=min(steps, 3))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * ONE_LEN / fps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Ours-s deblur (gshift_deblur2) synthetic 720p one_len=16, T=20", "timed": "CPU oracle port of the reference algorithm"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the ~25 s CPU oracle leg (profiling runs)")
    ap.add_argument("--frames", type=int, default=T, help="clip length incl. 4 context frames (default 20)")
    ap.add_argument("--arch", default=ARCH, help="other BASELINE configs: gshift_deblur1 (Ours+), gshift_denoise2, gshift_denoise1")
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout must carry exactly ONE JSON line: NCCL writes its version banner to stdout at communicator creation
        # (whatever NCCL_DEBUG says), so fd 1 points at stderr until the first collective has run
        os.environ.pop("NCCL_DEBUG", None)
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    GShiftNet = importlib.import_module("basicsr.models.archs." + args.arch).GShiftNet
    lib = pkg("host.lib").load()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_io as gio
    sd, spec = gio.synthetic_checkpoint(args.arch)
    H, W = args.height, args.width
    net = GShiftNet(future_frames=CTX, past_frames=CTX)
    net.load_state_dict(sd)
    net = net.half().to(dev).eval()
    Tn = args.frames
    out_frames = Tn - 2 * CTX
    nm_dev = None
    if spec.denoise:
        _, x, nm = pkg("host.synth").synthetic_clip(Tn, H, W, seed=7 + rank, denoise_sigma=30)
        nm_dev = nm.half().to(dev)
        _net = net
        net = lambda inp: _net(inp, nm_dev)          # noqa: E731  (the noise map is a constant, it stays on the device)
        net.engine = _net.engine
    else:
        _net = net
        _, x = pkg("host.synth").synthetic_clip(Tn, H, W, seed=7 + rank)
    x_host = x.half().pin_memory()
    x_dev = x_host.to(dev, non_blocking=True)
    out_host = torch.empty(out_frames, 3, H, W, dtype=torch.float16).pin_memory()
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") -----------------------------------------------------
    for _ in range(args.warmup):
        net(x_dev)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _net.kernel_launches     # our kernels executed (graph replays count the kernel nodes they contain)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if os.environ.get("GSN_NCU_RANGE") == "1":      # ncu --profile-from-start off: capture exactly the timed steps
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        out = net(x_dev)
    e1.record()
    barrier()
    if os.environ.get("GSN_NCU_RANGE") == "1":
        torch.cuda.profiler.stop()
    launches = _net.kernel_launches - l0
    ms = e0.elapsed_time(e1) / args.steps

    # ---- end to end through the public API with host buffers ("e2e") --------------------------------
    for _ in range(2):
        o = net(x_host.to(dev, non_blocking=True))
        out_host.copy_(o, non_blocking=True)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        o = net(x_host.to(dev, non_blocking=True))      # H2D of the step's clip from pinned memory
        out_host.copy_(o, non_blocking=True)            # D2H of the restored frames
    e3.record()
    barrier()
    sampler.stop_flag = True
    ms_e2e = e2.elapsed_time(e3) / args.steps

    t_all = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t_all.tolist()

    # ---- roofline of the dominant kernel (fused shift + NAF pass A), CUDA events on the launching stream -------
    roof = None
    if rank == 0:
        eng = net.engine()
        # the instrumented forward launches eagerly: release the CUDA graph (and its private activation pool -- 109 GiB for the
        # 1080p one_len=96 Ours+ clip) first, the timed runs are over
        _net._graphs = {}
        out = o = None
        torch.cuda.empty_cache()
        eng.timeline = []
        net(x_dev)
        torch.cuda.synchronize()
        tl, eng.timeline = eng.timeline, None
        agg = {}
        for name, pixels, a, b in tl:
            d = agg.setdefault(name, [0, 0.0, 0])
            d[0] += pixels
            d[1] += a.elapsed_time(b)
            d[2] += 1
        total_ms = sum(v[1] for v in agg.values())
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") else (6650.0, "fallback 6.65 TB/s")
        C = spec.c1
        names = [n for n in ("cab_pass_a_shift", "cab_pass_a") if n in agg]
        if not names:                                   # Ours+ runs the split (width-generic) block: report its grouped conv
            names = [n for n in ("group_conv5",) if n in agg]
        px = sum(agg[n][0] for n in names)
        tms = sum(agg[n][1] for n in names)
        nl = sum(agg[n][2] for n in names)
        alg_bytes = px * 4 * C                           # read x (C fp16) + write z (C fp16) per pixel
        ach = alg_bytes / (tms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "cab_pass_a (fused shift + NAF block, pass A)" if "cab_pass_a" in agg else "group_conv5", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                "launches": nl, "avg_launch_ms": tms / max(nl, 1), "alg_bytes_per_launch": alg_bytes / max(nl, 1),
                "kernel_share_of_step": {k: round(v[1] / total_ms, 4) for k, v in agg.items()},
                "instrumented_step_ms": total_ms}
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            roof["traffic"] = json.load(open(prof)).get("cab_pass_a_bytes_per_launch")
        # the other kernels of the shift block, same accounting (algorithmic bytes of SURVEY.md section 8d / CUDA-event time):
        # pass B reads z + shortcut and writes out (6C B/px; +2C where it also emits the next block's LayerNorm'd operand),
        # the gather + conv1 + LayerNorm producer reads the rolled stream + the shifted half and writes the 1.5C-wide operand
        other = {}
        for nm, bpp in (("cab_pass_b", 6 * C + C), ("shift_conv1_ln", 2 * C + C + 3 * C), ("ln_planar", 5 * C)):
            if nm in agg and agg[nm][1] > 0:
                gbs = agg[nm][0] * bpp / (agg[nm][1] * 1e-3) / 1e9
                other[nm] = {"bytes_per_pixel": bpp, "achieved": gbs, "frac": gbs / peak, "launches": agg[nm][2]}
        roof["other_kernels"] = other

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    fps = world * out_frames / (ms * 1e-3)
    fps_e2e = world * out_frames / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": f"{args.arch} synthetic {H}x{W} one_len={out_frames}, input (1,{Tn},3,{H},{W}), "
                               "random-init weights with randomised beta/LN", "sharding": f"clip per rank x{world}",
                   "l2": "inputs+activations (>>126 MB) larger than L2, no flush needed",
                   "accumulate": "fp32", "storage": "fp16 NHWC"},
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": x_host.numel() * 2,
                "d2h_bytes_per_step": out_host.numel() * 2, "ms_per_step": ms_e2e},
        "gpu_launches": int(launches) * world,
        "hbm_peak_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),     # rank 0, incl. the CUDA-graph pool
        "clocks": sampler.summary(),
        "roofline": roof,
    }
    if not args.no_cpu_baseline:
        fps_cpu, threads, sample, _ = cpu_oracle_rate()
        line["cpu_baseline"] = {"value": fps_cpu, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
