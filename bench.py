#!/usr/bin/env python
"""bench.py -- frames/sec of the Shift-Net forward hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's own CPU implementation on the host cores, same config

Workload (BASELINE.json configs[1]): Ours-s deblur, synthetic 720p clip, one_len=16 -> input (1,20,3,720,1280),
16 output frames per step.  A "step" is one GShiftNet forward of one clip.  Clips shard across ranks (weak scaling:
one clip per rank per step, no data-path collective; the only collective is the max-reduction of the timing).

stdout carries exactly ONE JSON line: fd 1 points at stderr for the whole run (NCCL_DEBUG=INFO prints to stdout by default and
is left enabled so the rank/topology lines stay visible there), the line is written to the saved descriptor at the end.
"""
import argparse
import importlib
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_REAL_STDOUT = os.dup(1)
sys.stdout.flush()
os.dup2(2, 1)            # everything anybody prints (NCCL INFO lines included) goes to stderr


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


import torch  # noqa: E402

ARCH = "gshift_deblur2"
ONE_LEN, CTX, H, W = 16, 2, 720, 1280
T = ONE_LEN + 2 * CTX
METRIC = "frames/sec (720p, one_len=16)"
REF_DIR = os.path.join(ROOT, "oracle", "_ref")      # the reference's four arch files, copied by __graft_entry__.build()


def pkg(sub):
    return importlib.import_module("shift-net_b200." + sub)


def synthetic_checkpoint(arch):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_io as gio
    return gio.synthetic_checkpoint(arch)


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


# ------------------------------------------------------------------------------------------------------------------
# the reference's own implementation (CPU arm, and the same-GPU eager arm)
# ------------------------------------------------------------------------------------------------------------------
class RefImpl:
    """The reference forward on an arbitrary device/dtype.  kind "reference": the reference's own arch file loaded by path from
    oracle/_ref/ (copied there, unmodified, by __graft_entry__.build() where /root/reference exists; git-ignored, travels to the
    GPU box).  kind "port": the oracle restatement (oracle/shiftnet_oracle.py, bit-identical outputs on the goldens)."""

    def __init__(self, arch, sd, device, dtype):
        self.arch, self.device, self.dtype = arch, device, dtype
        path = os.path.join(REF_DIR, arch + ".py")
        self.kind = "port"
        if os.path.exists(path):
            try:
                spec = importlib.util.spec_from_file_location("_gsn_ref_" + arch, path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                net = mod.GShiftNet(future_frames=CTX, past_frames=CTX)
                net.load_state_dict(sd, strict=True)
                self.net = net.to(device=device, dtype=dtype).eval()
                self.kind = "reference"
            except Exception as e:          # a reference file that cannot run here: fall back to the port, say so
                print(f"bench.py: oracle/_ref/{arch}.py not usable ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)
        if self.kind == "port":
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import shiftnet_oracle as O
            self.O = O
            self.sd = {k: v.to(device=device, dtype=dtype) for k, v in sd.items()}

    @torch.no_grad()
    def __call__(self, x, nm=None):
        if self.kind == "reference":
            return self.net(x, nm) if nm is not None else self.net(x)
        return self.O.gshiftnet_forward(self.sd, self.O.ARCHS[self.arch], x, nm)


def pick_cpu_threads(ref, spec):
    """All the host threads the reference can USE: the many small convs oversubscribe a 128-thread host (round 1: 264 s with 128
    threads vs 23 s with 8 for the same forward), so the thread count is calibrated on a short clip at the workload's resolution."""
    ncpu = os.cpu_count() or 1
    cands = sorted({min(c, ncpu) for c in (16, 32, 64)})
    _, x = pkg("host.synth").synthetic_clip(5, 96, 640)
    torch.set_num_threads(cands[0])
    ref(x)                                  # warm-up (allocator, oneDNN primitive cache)
    best = (None, 1e30)
    for c in cands:
        torch.set_num_threads(c)
        t0 = time.time()
        ref(x)
        dt = time.time() - t0
        print(f"bench.py: CPU calibration {c} threads: {dt:.2f} s", file=sys.stderr)
        if dt < best[1]:
            best = (c, dt)
    torch.set_num_threads(best[0])
    return best[0]


def cpu_sample_rate(arch, out_frames, Tn, Hh, Ww):
    """cpu_baseline of the main line: the reference on a BOUNDED sample (BASELINE config K1 shape: 8 frames of the clip cropped to
    256x256, ~5-25 s of CPU work), converted to the metric by input frame-pixels (extrapolated; the --impl reference arm times
    the 720p clip itself)."""
    sd, spec = synthetic_checkpoint(arch)
    ref = RefImpl(arch, sd, "cpu", torch.float32)
    threads = min(16, os.cpu_count() or 1)
    torch.set_num_threads(threads)
    TS, HS = 8, 256
    synth = pkg("host.synth")
    if spec.denoise:
        _, x, nm = synth.synthetic_clip(TS, HS, HS, denoise_sigma=30)
    else:
        (_, x), nm = synth.synthetic_clip(TS, HS, HS), None
    ref(x[:, :5, :, :64, :64], None if nm is None else nm[:, :5, :, :64, :64])
    t0 = time.time()
    ref(x, nm)
    sec = time.time() - t0
    px_frames_per_s = TS * HS * HS / sec
    fps = out_frames / (Tn * Hh * Ww / px_frames_per_s)
    sample = (f"T={TS} {HS}x{HS} crop of the clip ({sec:.1f} s per forward, {threads} threads of {os.cpu_count()}); scaled to the "
              f"{Hh}x{Ww} one_len={out_frames} clip by input frame-pixels (extrapolated)")
    return fps, threads, sample, ref.kind


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation on THIS config (the 720p one_len=16 clip itself, fp32), all the
    host threads it can use.  One forward takes minutes on the host, so the number of timed steps is capped by a wall budget
    (GSN_REF_BUDGET_S, default 240 s; at least one full-clip step); `steps` reports what was timed."""
    if rank != 0:
        return
    sd, spec = synthetic_checkpoint(args.arch)
    ref = RefImpl(args.arch, sd, "cpu", torch.float32)
    threads = pick_cpu_threads(ref, spec)
    Tn, Hh, Ww = args.frames, args.height, args.width
    out_frames = Tn - 2 * CTX
    synth = pkg("host.synth")
    nm = None
    if spec.denoise:
        _, x, nm = synth.synthetic_clip(Tn, Hh, Ww, denoise_sigma=30)
    else:
        _, x = synth.synthetic_clip(Tn, Hh, Ww)
    budget = float(os.environ.get("GSN_REF_BUDGET_S", "240"))
    times = []
    t_start = time.time()
    while len(times) < max(1, args.steps):
        t0 = time.time()
        ref(x, nm)
        times.append(time.time() - t0)
        if time.time() - t_start + times[-1] > budget:
            break
    sec = sum(times) / len(times)
    fps = out_frames / sec
    sample = (f"the full clip (1,{Tn},3,{Hh},{Ww}) fp32, {len(times)} timed forward(s) of {sec:.1f} s, {threads} threads of "
              f"{os.cpu_count()} (calibrated), no warm-up step beyond the thread calibration")
    emit({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": len(times),
        "steps_requested": args.steps, "warmup": 0, "warmup_requested": args.warmup, "ms_per_step": 1000.0 * sec,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.arch} synthetic {Hh}x{Ww} one_len={out_frames}, input (1,{Tn},3,{Hh},{Ww}), "
                               "random-init weights with randomised beta/LN",
                   "timed": ("the reference's own arch file (oracle/_ref), torch CPU fp32" if ref.kind == "reference"
                             else "oracle port of the reference algorithm, torch CPU fp32")},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": ref.kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def gpu_eager_baseline(arch, sd, spec, x_dev, nm_dev, out_frames, steps=3):
    """SURVEY.md section 2 / BASELINE.md section 3: the path the reference actually ships -- its PyTorch-eager fp16 forward
    (inference/test_deblur_small.py:86-89,134) -- on the SAME B200, CUDA-event timed on the same clip."""
    try:
        ref = RefImpl(arch, sd, x_dev.device, torch.float16)
        for _ in range(2):
            ref(x_dev, nm_dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ref(x_dev, nm_dev)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res = {"value": out_frames / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms, "steps": steps, "kind": ref.kind,
               "dtype": "f16", "what": "reference forward, PyTorch eager + cuDNN on this GPU, same clip, CUDA events"}
        del ref
    except Exception as e:
        res = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    torch.cuda.empty_cache()
    return res


def run_strong(args, net, spec, rank, world, dev, dist):
    """--scaling strong: ONE clip, its frames split across the ranks (host/tshard.py); before every CAB2 each rank receives one
    boundary frame's half of the channels from its ring neighbour (NCCL send/recv over NVLink).  value = restored frames of the
    whole clip / max-over-ranks time."""
    Tn, Hh, Ww = args.frames, args.height, args.width
    out_frames = Tn - 2 * CTX
    _, x = pkg("host.synth").synthetic_clip(Tn, Hh, Ww, seed=7)          # the same clip on every rank
    ts = pkg("host.tshard").TShard(rank, world, Tn)
    x_host = x[:, ts.a:ts.b].half().contiguous().pin_memory()
    x_dev = x_host.to(dev, non_blocking=True)
    lo, hi = ts.local_output_range(CTX, CTX)
    out_host = torch.empty(hi - lo, 3, Hh, Ww, dtype=torch.float16).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(dev.index)        # started before the warm-up: a T-sharded step is tens of milliseconds, and the warm-up
    sampler.start()                          # runs the same load (nvidia-smi needs ~0.1-0.3 s per sample on an 8-GPU box)
    for _ in range(args.warmup):
        net.forward_tsharded(x_dev, ts)
    barrier()
    ts.halo_bytes = ts.exchanges = 0
    l0 = net.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        net.forward_tsharded(x_dev, ts)
    e1.record()
    barrier()
    launches = net.kernel_launches - l0
    ms = e0.elapsed_time(e1) / args.steps
    halo_per_step = ts.halo_bytes / args.steps
    nx = ts.exchanges // args.steps
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        o = net.forward_tsharded(x_host.to(dev, non_blocking=True), ts)
        out_host.copy_(o, non_blocking=True)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3) / args.steps
    # one more step with CUDA events around every exchange: time spent in the halo traffic
    ts.time_exchanges, ts.events = True, []
    net.forward_tsharded(x_dev, ts)
    torch.cuda.synchronize()
    ms_x = sum(a.elapsed_time(b) for a, b in ts.events)
    sampler.stop_flag = True
    t_all = torch.tensor([ms, ms_e2e, ms_x], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_x = t_all.tolist()
    if rank == 0:
        emit({
            "metric": METRIC, "value": out_frames / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": f"{args.arch} synthetic {Hh}x{Ww} one_len={out_frames}, ONE clip (1,{Tn},3,{Hh},{Ww}) T-sharded over {world} rank(s)",
                       "sharding": f"frames of one clip x{world} (rank 0 owns {ts.n_local}), halo exchange before every CAB2, " + ("CUDA graphs cut at the exchanges, NCCL between them" if (world > 1 and os.environ.get("GSN_TSHARD_GRAPH", "1") != "0" and os.environ.get("GSN_CUDA_GRAPH", "1") != "0") else "eager launches"),
                       "l2": "activations larger than L2, no flush needed", "accumulate": "fp32", "storage": "fp16 NHWC"},
            "e2e": {"value": out_frames / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": x_host.numel() * 2 * world,
                    "d2h_bytes_per_step": out_frames * 3 * Hh * Ww * 2, "ms_per_step": ms_e2e},
            "gpu_launches": int(launches) * world,
            "halo": {"exchanges_per_step": nx, "bytes_sent_per_rank_per_step": halo_per_step, "ms_in_exchanges_per_step": ms_x,
                     "achieved_gbs_per_direction": halo_per_step / 1e9 / max(ms_x * 1e-3, 1e-12),
                     "reference_gbs": 770.0, "note": "NCCL send/recv of C/2 channels of one boundary frame per CAB2 (48 per forward)"},
            "clocks": sampler.summary(),
        })
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU and GPU-eager baseline legs (profiling runs)")
    ap.add_argument("--frames", type=int, default=T, help="clip length incl. 4 context frames (default 20)")
    ap.add_argument("--arch", default=ARCH, help="other BASELINE configs: gshift_deblur1 (Ours+), gshift_denoise2, gshift_denoise1")
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: one clip per rank (default); strong: ONE clip T-sharded across the ranks with halo exchange over NVLink")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
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
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        torch.cuda.synchronize()

    GShiftNet = importlib.import_module("basicsr.models.archs." + args.arch).GShiftNet
    sd, spec = synthetic_checkpoint(args.arch)
    Hh, Ww = args.height, args.width
    net = GShiftNet(future_frames=CTX, past_frames=CTX)
    net.load_state_dict(sd)
    net = net.half().to(dev).eval()
    Tn = args.frames
    out_frames = Tn - 2 * CTX
    nm_dev = None
    _net = net
    if spec.denoise:
        _, x, nm = pkg("host.synth").synthetic_clip(Tn, Hh, Ww, seed=7 + rank, denoise_sigma=30)
        nm_dev = nm.half().to(dev)
        net = lambda inp: _net(inp, nm_dev)          # noqa: E731  (the noise map is a constant, it stays on the device)
    else:
        _, x = pkg("host.synth").synthetic_clip(Tn, Hh, Ww, seed=7 + rank)
    if args.scaling == "strong":
        run_strong(args, _net, spec, rank, world, dev, dist)
        return
    x_host = x.half().pin_memory()
    x_dev = x_host.to(dev, non_blocking=True)
    out_host = torch.empty(out_frames, 3, Hh, Ww, dtype=torch.float16).pin_memory()
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
        eng = _net.engine()
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
        Cc = spec.c1
        names = [n for n in ("cab_pass_a_shift", "cab_pass_a") if n in agg]
        if not names:                                   # Ours+ runs the split (width-generic) block: report its grouped conv
            names = [n for n in ("group_conv5",) if n in agg]
        px = sum(agg[n][0] for n in names)
        tms = sum(agg[n][1] for n in names)
        nl = sum(agg[n][2] for n in names)
        alg_bytes = px * 4 * Cc                          # pass A reads x (C fp16) + writes z (C fp16) per pixel
        ach = alg_bytes / (tms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "cab_pass_a (fused shift + NAF block, pass A)" if "cab_pass_a" in agg else "group_conv5", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                "launches": nl, "avg_launch_ms": tms / max(nl, 1), "alg_bytes_per_launch": alg_bytes / max(nl, 1),
                "kernel_share_of_step": {k: round(v[1] / total_ms, 4) for k, v in agg.items()},
                "instrumented_step_ms": total_ms}
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            tj = json.load(open(prof))
            roof["traffic"] = tj.get("cab_pass_a_bytes_per_launch")
            roof["traffic_source"] = tj.get("source")
        if "cab_pass_a" in agg:
            # the kernel's true bound: its two depthwise stages are HFMA2 work on the one FMA pipe (2.0 warp-instr/clk/SM
            # measured, profiles/r1f_ubench_pipes.txt); HFMA2 per pixel = the kernel's own instruction count (csrc comments)
            hf = getattr(eng, "pass_a_hfma2_per_pixel", None)
            sm_mhz = (sampler.summary().get("sm_mhz") or 1965)
            if hf:
                issued = px * hf / 32.0                  # warp instructions
                cap = 2.0 * 148 * (tms * 1e-3) * sm_mhz * 1e6
                roof["bound_secondary"] = {"bound": "fma", "hfma2_warp_instr": issued, "pipe_capacity_warp_instr": cap, "frac": issued / cap,
                                           "note": "HFMA2 warp-instructions issued / (2.0 per clk per SM x cycles)"}
        # the other kernels of the shift block on SURVEY.md section 8(d)'s bytes (CUDA-event time): pass B reads z + shortcut and
        # writes out = 6C B/px.  shift_conv1_ln / ln_planar have NO algorithmic bytes in 8(d) (the gather was to ride a load
        # stage): their cost shows up only in the block line below.
        other = {}
        for nm_, bpp in (("cab_pass_b", 6 * Cc),):
            if nm_ in agg and agg[nm_][1] > 0:
                gbs = agg[nm_][0] * bpp / (agg[nm_][1] * 1e-3) / 1e9
                other[nm_] = {"bytes_per_pixel": bpp, "achieved": gbs, "frac": gbs / peak, "launches": agg[nm_][2]}
        roof["other_kernels"] = other
        blk_names = [n for n in ("shift_conv1_ln", "shift_conv1", "ln_planar", "cab_pass_a_shift", "cab_pass_a", "cab_pass_a2", "cab_pass_b",
                                 "ln_pw", "group_conv5", "cab_fold") if n in agg]
        blk_ms = sum(agg[n][1] for n in blk_names)
        cab_px = sum(agg[n][0] for n in ("cab_pass_b",) if n in agg)          # one pass B per CAB: CAB-pixels of the step
        if blk_ms > 0 and cab_px:
            per_cab = (14 if spec.denoise else 10) * Cc                        # 10C B per CAB-pixel (14C denoise: three passes)
            gbs = cab_px * per_cab / (blk_ms * 1e-3) / 1e9
            roof["block"] = {"what": "whole shift block (all kernels of every (shift, CAB2, CAB1) pair)", "bytes_per_pair_pixel": 2 * per_cab,
                             "alg_bytes": cab_px * per_cab, "ms": blk_ms, "achieved": gbs, "frac": gbs / peak, "kernels": blk_names}

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
        "config": {"workload": f"{args.arch} synthetic {Hh}x{Ww} one_len={out_frames}, input (1,{Tn},3,{Hh},{Ww}), "
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
    if not args.no_cpu_baseline and world == 1:
        _net._engine = None
        torch.cuda.empty_cache()
        line["gpu_eager_baseline"] = gpu_eager_baseline(args.arch, sd, spec, x_dev, nm_dev, out_frames)
        fps_cpu, threads, sample, kind = cpu_sample_rate(args.arch, out_frames, Tn, Hh, Ww)
        line["cpu_baseline"] = {"value": fps_cpu, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
