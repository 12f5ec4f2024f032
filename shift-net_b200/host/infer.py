"""Evaluation loop behind the ``inference/test_*.py`` entry points (reference: inference/test_deblur_small.py:51-177,
inference/test_denoise_small.py:51-224), re-designed around the B200 path:

* same CLI, chunking rules, metrics and log-line formats as the reference scripts;
* work units (video, chunk) are sharded round-robin across the ranks of a ``torchrun`` launch (one process per GPU),
  weights replicated, no communication during the forward; ONE all_gather (NCCL on GPUs, gloo on CPU tests) of the
  per-frame (video, frame, psnr, ssim) records at the end so rank 0 prints the per-video / total averages;
* frames come from disk (imageio or cv2) or, when no dataset is present (``--synthetic``), from a seeded generator.
"""
from __future__ import annotations

import glob
import math
import os
import time

import numpy as np
import torch


class TraverseLogger:
    def __init__(self, result_dir, filename="inference_log.txt", enabled=True):
        self.enabled = enabled
        self.f = open(os.path.join(result_dir, filename), "a") if enabled else None

    def write_log(self, log):
        if self.enabled:
            print(log, flush=True)
            self.f.write(log + "\n")
            self.f.flush()


def psnr_255(out, gt):
    """skimage.metrics.peak_signal_noise_ratio(out, gt, data_range=255) restated (float64 MSE)."""
    mse = np.mean((np.asarray(out, dtype=np.float64) - np.asarray(gt, dtype=np.float64)) ** 2)
    return float("inf") if mse == 0 else 10.0 * math.log10(255.0 ** 2 / mse)


def ssim_calculate(img1, img2, sd=1.5, C1=0.01 ** 2, C2=0.03 ** 2):
    """Gaussian-window SSIM on /255 CHW float32 (inference/test_deblur_small.py:25-49)."""
    from scipy.ndimage import gaussian_filter
    a = (np.array(img1, dtype=np.float32) / 255).transpose((2, 0, 1))
    b = (np.array(img2, dtype=np.float32) / 255).transpose((2, 0, 1))
    mu1, mu2 = gaussian_filter(a, sd), gaussian_filter(b, sd)
    s1 = gaussian_filter(a * a, sd) - mu1 * mu1
    s2 = gaussian_filter(b * b, sd) - mu2 * mu2
    s12 = gaussian_filter(a * b, sd) - mu1 * mu2
    return float(np.mean(((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))))


def _imread(path):
    try:
        import imageio
        return np.asarray(imageio.imread(path))
    except ImportError:
        import cv2
        return cv2.imread(path, cv2.IMREAD_COLOR)[..., ::-1]


def synthetic_videos(n_videos, n_frames, h, w, seed=0):
    """{name: (blurred/noisy-free inputs uint8 list, gt uint8 list)} -- smooth random frames (no dataset offline)."""
    out = {}
    for v in range(n_videos):
        g = torch.Generator().manual_seed(seed + v)
        gt = torch.rand(n_frames, 3, h, w, generator=g)
        gt = torch.nn.functional.avg_pool2d(gt, 5, 1, 2, count_include_pad=False)
        blur = torch.nn.functional.avg_pool2d(gt, 3, 1, 1, count_include_pad=False)
        to_u8 = lambda t: [(f.permute(1, 2, 0) * 255).round().clamp(0, 255).byte().numpy() for f in t]
        out[f"syn{v:03d}"] = (to_u8(blur), to_u8(gt))
    return out


def plan_units(videos, task, one_len):
    """[(video, kk, in_frames, gt_frames)] with the reference's chunking.
    deblur (test_deblur_small.py:111-120): k_len = (n-4)//one_len, tail dropped.
    denoise (test_denoise_small.py:113-131): one_len = n-4 (halved if > 100), the last chunk takes the remainder."""
    units = []
    for v in sorted(videos):
        ins, gts = videos[v]
        n = len(ins)
        if task == "deblur":
            L = one_len
            for kk in range((n - 4) // L):
                units.append((v, kk, ins[kk * L:kk * L + L + 4], gts[kk * L + 2:kk * L + 2 + L]))
        else:
            L = n - 4
            if L > 100:
                L //= 2
            k_len, k_res = (n - 4) // L, (n - 4) % L
            for kk in range(k_len):
                add = k_res if kk == k_len - 1 else 0
                units.append((v, kk, ins[kk * L:kk * L + L + add + 4], ins[kk * L + 2:kk * L + L + add + 2]))
    return units


def frame_bases(units):
    """{(video, kk): index of the chunk's first output frame within its video} -- the reference's running `index`."""
    return {(v, kk): sum(len(u[2]) - 4 for u in units if u[0] == v and u[1] < kk) for v, kk, _, _ in units}


def load_frames(frames):
    """paths / arrays -> list of uint8 HWC arrays cropped to multiples of 4 (test_deblur_small.py:122-127)."""
    frames = [np.asarray(_imread(f) if isinstance(f, str) else f) for f in frames]
    h, w = frames[0].shape[:2]
    h, w = h - h % 4, w - w % 4
    return [f[:h, :w] for f in frames]


def to_tensor(frames):
    """uint8 HWC list -> (1,T,3,H,W) float32 in [0,1] with numpy2tensor's arithmetic: float32(u8) * float32(1/255)
    (test_deblur_small.py:191-200 -- a multiplication by the rounded reciprocal, not a division)."""
    frames = load_frames(frames)
    t = torch.from_numpy(np.ascontiguousarray(np.stack(frames))).permute(0, 3, 1, 2).float().mul_(1.0 / 255)
    return t.unsqueeze(0), frames


class DeviceIO:
    """The I/O side of the evaluation loop on the GPU: frames travel as uint8 from pinned host
    memory (1 byte per sample), become the fp16 clip on the device (gsn_u8_to_clip), and the PSNR of every restored frame is
    reduced on the device against the uint8 ground truth (gsn_psnr_sse): only 64 doubles per frame come back."""

    def __init__(self, device):
        import importlib
        self.L = importlib.import_module("shift-net_b200.host.lib")
        self.lib = self.L.load()
        self.dev = device
        self.nb = self.lib.gsn_psnr_sse_blocks()
        self.nbs = self.lib.gsn_ssim_blocks()

    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def upload_u8(self, frames):
        """list of uint8 HWC arrays -> (T,H,W,3) uint8 device tensor through pinned memory."""
        host = torch.from_numpy(np.ascontiguousarray(np.stack(frames))).pin_memory()
        return host.to(self.dev, non_blocking=True)

    def clip_from_u8(self, frames_dev, dtype=torch.float16):
        T, H, W, _ = frames_dev.shape
        clip = torch.empty(1, T, 3, H, W, dtype=dtype, device=self.dev)
        dt = self.L.DTYPE_F16 if dtype == torch.float16 else self.L.DTYPE_F32
        self.L.check(self.lib.gsn_u8_to_clip(frames_dev.data_ptr(), T, H, W, dt, clip.data_ptr(), self._stream()), "u8_to_clip")
        return clip

    def psnr(self, out, gt_dev):
        """out (T,3,H,W) fp16/fp32 as returned by the net, gt_dev (T,H,W,3) uint8 -> list of T PSNR values (float64)."""
        T, _, H, W = out.shape
        out = out.contiguous()
        part = torch.empty(T, self.nb, dtype=torch.float64, device=self.dev)
        dt = self.L.DTYPE_F16 if out.dtype == torch.float16 else self.L.DTYPE_F32
        self.L.check(self.lib.gsn_psnr_sse(out.data_ptr(), dt, gt_dev.data_ptr(), T, H, W, part.data_ptr(), self._stream()), "psnr_sse")
        res = []
        for row in part.cpu().tolist():
            sse = 0.0
            for v in row:                 # fixed order: deterministic
                sse += v
            mse = sse / (3.0 * H * W)
            res.append(float("inf") if mse == 0 else 10.0 * math.log10(255.0 ** 2 / mse))
        return res

    def ssim(self, out, gt_dev):
        """out (T,3,H,W) fp16/fp32 as returned by the net, gt_dev (T,H,W,3) uint8 -> list of T SSIM values: the reference's
        ssim_calculate (scipy 3-D Gaussian, sigma 1.5) on the device (gsn_ssim); 128 doubles per frame come back."""
        T, _, H, W = out.shape
        out = out.contiguous()
        ws = torch.empty(self.lib.gsn_ssim_workspace_bytes(T, H, W), dtype=torch.uint8, device=self.dev)
        part = torch.empty(T, self.nbs, dtype=torch.float64, device=self.dev)
        dt = self.L.DTYPE_F16 if out.dtype == torch.float16 else self.L.DTYPE_F32
        self.L.check(self.lib.gsn_ssim(out.data_ptr(), dt, gt_dev.data_ptr(), T, H, W, ws.data_ptr(), part.data_ptr(), self._stream()), "ssim")
        res = []
        for row in part.cpu().tolist():
            s = 0.0
            for v in row:                     # fixed order: deterministic
                s += v
            res.append(s / (3.0 * H * W))
        return res


def gather_records(records, device):
    """all_gather of (video_idx, frame_idx, psnr, ssim) rows across ranks -> full list on every rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return records
    world = dist.get_world_size()
    n = torch.tensor([len(records)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    m = max(int(c.item()) for c in counts)
    buf = torch.zeros(max(m, 1), 4, dtype=torch.float64, device=device)
    if records:
        buf[:len(records)] = torch.tensor(records, dtype=torch.float64, device=device)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    out = []
    for c, b in zip(counts, bufs):
        out += [tuple(r) for r in b[:int(c.item())].cpu().tolist()]
    return out


def summarize(records, names, task, log):
    per = {}
    for vi, fi, p, s in sorted(records):
        per.setdefault(int(vi), []).append((p, s))
    sp = ss = n = 0.0
    sp2 = ss2 = 0.0
    for vi in sorted(per):
        ps = [x[0] for x in per[vi]]
        sm = [x[1] for x in per[vi]]
        log("# Video:{} AVG-PSNR={:.5}, AVG-SSIM={:.4}".format(names[vi], sum(ps) / len(ps), sum(sm) / len(sm)))
        sp += sum(ps); ss += sum(sm); n += len(ps)
        sp2 += sum(ps) / len(ps); ss2 += sum(sm) / len(sm)
    if n:
        log("# Total AVG-PSNR={:.5}, AVG-SSIM={:.4}".format(sp / n, ss / n))
        if task == "denoise":                                  # test_denoise_small.py:223-224 also prints the mean of means
            log("# Total AVG-PSNR={:.5}, AVG-SSIM={:.4}".format(sp2 / len(per), ss2 / len(per)))
    return (sp / n, ss / n) if n else (float("nan"), float("nan"))


def run(net_cls, task, args):
    """Shared body of the four entry scripts.  ``net_cls`` is the arch's GShiftNet."""
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    use_cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if use_cuda else "gloo")
    os.makedirs(args.result_path, exist_ok=True)
    log = TraverseLogger(args.result_path, "inference_log_{}.txt".format(time.strftime("%Y-%m-%d %H:%M:%S")), rank == 0)
    log.write_log("Inference - {} (shiftnet_b200, {} rank(s))".format(time.strftime("%Y-%m-%d %H:%M:%S"), world))

    net = net_cls(future_frames=2, past_frames=2)
    if args.model_path and os.path.exists(args.model_path):
        net.load_state_dict(torch.load(args.model_path, map_location="cpu")["params"])
        log.write_log("Loading model from {}".format(args.model_path))
    elif args.synthetic or getattr(args, "random_weights", False):
        importlib_synth = __import__("importlib").import_module("shift-net_b200.host.synth")
        importlib_synth.randomize_(net, 1234)
        log.write_log("No checkpoint at {!r}: using the seeded synthetic checkpoint (--synthetic / --random_weights)".format(args.model_path))
    else:
        # the reference fails hard in torch.load (test_deblur_small.py:85); metrics of random weights would look plausible
        raise FileNotFoundError("checkpoint {!r} not found (pass --synthetic N or --random_weights to run without one)".format(args.model_path))
    net = net.half().to(device).eval()

    if args.synthetic:
        videos = synthetic_videos(args.synthetic, args.synthetic_frames, args.synthetic_h, args.synthetic_w)
    else:
        in_dir = os.path.join(args.data_path, "blur") if task == "deblur" else args.data_path
        gt_dir = os.path.join(args.data_path, "gt") if task == "deblur" else args.data_path
        videos = {v: (sorted(glob.glob(os.path.join(in_dir, v, "*"))), sorted(glob.glob(os.path.join(gt_dir, v, "*"))))
                  for v in sorted(os.listdir(in_dir))}
    names = sorted(videos)
    units = plan_units(videos, task, getattr(args, "one_len", 0))
    records = []
    # running frame index per video, as the reference keeps it (test_denoise_small.py:184-189): the last denoise chunk is longer
    # than the others, so kk * chunk_len would mis-number its frames
    frame_base = frame_bases(units)
    dio = DeviceIO(device) if (use_cuda and not getattr(args, "cpu_io", False)) else None
    with torch.no_grad():
        for ui in range(rank, len(units), world):             # clip sharding: unit i -> rank i % world
            v, kk, in_seq, gt_seq = units[ui]
            t0 = time.time()
            gpu_psnr = gpu_ssim = None
            if task == "deblur" and dio is not None:
                # device I/O path: uint8 H2D, /255, PSNR and SSIM reductions on the GPU -- only scalars come back over PCIe
                ins, gts = load_frames(in_seq), load_frames(gt_seq)
                x = dio.clip_from_u8(dio.upload_u8(ins))
                out = net(x)
                gt_dev = dio.upload_u8(gts)
                gpu_psnr, gpu_ssim = dio.psnr(out, gt_dev), dio.ssim(out, gt_dev)
            elif task == "deblur":
                x, _ = to_tensor(in_seq)
                _, gts = to_tensor(gt_seq)
                out = net(x.to(device).half()).float()
            else:
                sigma = args.sigma / 255.0
                gts = load_frames(gt_seq)
                if dio is not None:
                    # clean frames travel as uint8; the AWGN is drawn on the host with the per-unit seed (as --cpu_io does) or, with
                    # --device_noise, on the GPU (same distribution, different stream: no float clip crosses PCIe at all)
                    x = dio.clip_from_u8(dio.upload_u8(load_frames(in_seq)), torch.float32)
                    if getattr(args, "device_noise", False):
                        gd = torch.Generator(device=device).manual_seed(1000 * names.index(v) + kk)
                        x = (x + sigma * torch.randn(x.shape, generator=gd, device=device)).half()
                    else:
                        g = torch.Generator().manual_seed(1000 * names.index(v) + kk)
                        x = (x + (sigma * torch.randn(x.shape, generator=g)).pin_memory().to(device, non_blocking=True)).half()
                else:
                    x, _ = to_tensor(in_seq)
                    g = torch.Generator().manual_seed(1000 * names.index(v) + kk)
                    x = (x + sigma * torch.randn(x.shape, generator=g)).to(device).half()
                B, N, _, H, W = x.shape
                std = torch.full((1, 1, 1, 1, 1), sigma, device=device, dtype=torch.float16)
                ph, pw = 32 - (H // 2 % 16), 32 - (W // 2 % 16)   # 2x2 overlapped tiling, test_denoise_small.py:153-173
                out = torch.zeros(N - 4, 3, H, W, device=device)
                hh, ww = H // 2 + ph, W // 2 + pw
                nm = std.expand(B, N, 1, hh, ww)                  # the four tiles share one shape: one CUDA graph serves them all
                o = net(x[..., 0:hh, 0:ww].contiguous(), nm).float(); out[..., 0:H // 2, 0:W // 2] = o[..., 0:-ph, 0:-pw]
                o = net(x[..., 0:hh, W // 2 - pw:].contiguous(), nm).float(); out[..., 0:H // 2, W // 2:] = o[..., 0:-ph, pw:]
                o = net(x[..., H // 2 - ph:, 0:ww].contiguous(), nm).float(); out[..., H // 2:, 0:W // 2] = o[..., ph:, 0:-pw]
                o = net(x[..., H // 2 - ph:, W // 2 - pw:].contiguous(), nm).float(); out[..., H // 2:, W // 2:] = o[..., ph:, pw:]
                if dio is not None:
                    gt_dev = dio.upload_u8(gts)
                    gpu_psnr, gpu_ssim = dio.psnr(out, gt_dev), dio.ssim(out, gt_dev)
            t1 = time.time()
            imgs = None
            if gpu_psnr is None or args.save_image:       # host metrics (--cpu_io) or PNG output: the frames have to come back
                imgs = (out.float().clamp(0, 1.0) * 255).permute(0, 2, 3, 1).cpu().numpy()
            base = frame_base[(v, kk)]
            for e in range(out.shape[0]):
                p = gpu_psnr[e] if gpu_psnr is not None else psnr_255(imgs[e], gts[e])
                s = gpu_ssim[e] if gpu_ssim is not None else ssim_calculate(imgs[e], gts[e])
                records.append((names.index(v), base + e, p, s))
                if args.save_image:
                    import cv2
                    os.makedirs(os.path.join(args.result_path, v), exist_ok=True)
                    cv2.imwrite(os.path.join(args.result_path, v, "%03d.png" % (base + e)), imgs[e][..., ::-1])
            print("> [rank {}] {}-{} PSNR={:.5}, SSIM={:.4} forward_time:{:.3}s, total_time:{:.3}s".format(
                rank, v, kk, p, s, t1 - t0, time.time() - t0), flush=True)
    records = gather_records(records, device)
    res = summarize(records, names, task, log.write_log)
    if world > 1 and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
    return res, records


def add_common_args(parser):
    parser.add_argument("--save_image", action="store_true", default=False, help="save image if true")
    parser.add_argument("--border", action="store_true", help="restore border images of video if true")
    parser.add_argument("--default_data", type=str, default=".")
    parser.add_argument("--data_path", type=str, default=None)
    parser.add_argument("--model_path", type=str, default=None)
    parser.add_argument("--result_path", type=str, default=None)
    parser.add_argument("--cpu_io", action="store_true", help="reference-style host I/O: float conversion and PSNR on the CPU")
    parser.add_argument("--device_noise", action="store_true", help="denoise: draw the AWGN on the GPU instead of the host")
    parser.add_argument("--random_weights", action="store_true", help="run without a checkpoint on the seeded synthetic weights")
    parser.add_argument("--synthetic", type=int, default=0, help="number of synthetic videos (no dataset needed)")
    parser.add_argument("--synthetic_frames", type=int, default=12)
    parser.add_argument("--synthetic_h", type=int, default=64)
    parser.add_argument("--synthetic_w", type=int, default=96)
    return parser
