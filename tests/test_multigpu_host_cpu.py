"""Host-side multi-process logic on CPU (gloo, world_size 2): clip sharding + the final all_gather of PSNR records."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_io as gio


def _fake_metric(vi, fi):
    return 20.0 + 0.37 * vi + 0.011 * fi, 0.5 + 0.001 * (vi + fi)


def _records_for(units, names, rank, world):
    rec = []
    for ui in range(rank, len(units), world):
        v, kk, in_seq, gt_seq = units[ui]
        base = kk * (len(in_seq) - 4)
        for e in range(len(gt_seq)):
            p, s = _fake_metric(names.index(v), base + e)
            rec.append((names.index(v), base + e, p, s))
    return rec


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    infer = gio.pkg("host.infer")
    videos = {f"v{i}": (list(range(30 + 7 * i)), list(range(30 + 7 * i))) for i in range(3)}
    names = sorted(videos)
    units = infer.plan_units(videos, "deblur", 8)
    mine = _records_for(units, names, rank, world)
    full = infer.gather_records(mine, torch.device("cpu"))
    lines = []
    res = infer.summarize(full, names, "deblur", lines.append)
    q.put((rank, len(mine), sorted(full), res, lines))
    dist.barrier()
    dist.destroy_process_group()


def test_clip_sharding_and_all_gather_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=120) for _ in range(2)]
    [p.join(timeout=60) for p in procs]
    infer = gio.pkg("host.infer")
    videos = {f"v{i}": (list(range(30 + 7 * i)), list(range(30 + 7 * i))) for i in range(3)}
    names = sorted(videos)
    units = infer.plan_units(videos, "deblur", 8)
    single = sorted(_records_for(units, names, 0, 1))
    assert sum(g[1] for g in got) == len(single) and all(g[1] > 0 for g in got)      # both ranks did work, nothing lost
    for rank, n, full, res, lines in got:
        assert full == single                                                       # every rank sees the complete set
        assert lines[-1].startswith("# Total AVG-PSNR=")
    assert got[0][3] == got[1][3]


def test_chunking_rules_match_reference_scripts():
    infer = gio.pkg("host.infer")
    # deblur: k_len = (n-4)//one_len, tail dropped, each chunk carries 2+2 context frames
    u = infer.plan_units({"a": (list(range(50)), list(range(50)))}, "deblur", 16)
    assert [(x[1], len(x[2]), len(x[3])) for x in u] == [(0, 20, 16), (1, 20, 16)]
    assert u[1][2][0] == 16 and u[1][3][0] == 18
    # denoise: one_len = n-4, halved if > 100, last chunk takes the remainder; gt are the clean input frames
    u = infer.plan_units({"a": (list(range(131)), None)}, "denoise", 0)
    assert [(len(x[2]), len(x[3])) for x in u] == [(67, 63), (68, 64)]
    u = infer.plan_units({"a": (list(range(60)), None)}, "denoise", 0)
    assert [(len(x[2]), len(x[3])) for x in u] == [(60, 56)]


def test_metrics_formulas():
    import numpy as np
    infer = gio.pkg("host.infer")
    g = np.random.default_rng(0)
    a = g.uniform(0, 255, (24, 32, 3)); b = a + g.normal(0, 5, a.shape)
    mse = np.mean((a - b) ** 2)
    assert abs(infer.psnr_255(a, b) - 10 * np.log10(255 ** 2 / mse)) < 1e-9
    assert infer.psnr_255(a, a) == float("inf")
    assert abs(infer.ssim_calculate(a, a) - 1.0) < 1e-6 and infer.ssim_calculate(a, b) < 1.0


def test_to_tensor_matches_reference_numpy2tensor_arithmetic():
    """inference/test_deblur_small.py:191-200: float32(u8) then mul_(1/255) -- a multiplication by the float32-rounded reciprocal.
    All 256 byte values, through float32 and through the .half() the scripts apply."""
    import numpy as np
    import torch
    infer = gio.pkg("host.infer")
    img = np.arange(256, dtype=np.uint8).reshape(4, 64, 1).repeat(3, axis=2)          # HWC, H=4 W=64
    ours, frames = infer.to_tensor([img])
    ref = torch.from_numpy(np.ascontiguousarray(np.array(img).astype("float64").transpose((2, 0, 1)))).float()
    ref.mul_(1.0 / 255)
    assert torch.equal(ours[0, 0], ref) and torch.equal(ours.half()[0, 0], ref.half())
    assert frames[0].dtype == np.uint8
    # the CUDA kernel (gsn_u8_to_clip) computes float32(u8) * float32(1/255) with one rounding: the same values
    k = np.float32(1.0 / 255.0)
    assert np.array_equal((np.arange(256, dtype=np.float32) * k), ref[0].reshape(-1).numpy())


def test_denoise_chunking_numbers_frames_with_a_running_index():
    """test_denoise_small.py:113-131,184-189: 107 frames -> one_len 51, chunks of 51 and 52 (the last takes the remainder); the
    second chunk's frames start at 51, not at kk * its own length."""
    infer = gio.pkg("host.infer")
    videos = {"a": (list(range(107)), list(range(107))), "b": (list(range(20)), list(range(20)))}
    units = infer.plan_units(videos, "denoise", 0)
    assert [(u[0], u[1], len(u[2]) - 4) for u in units] == [("a", 0, 51), ("a", 1, 52), ("b", 0, 16)]
    assert infer.frame_bases(units) == {("a", 0): 0, ("a", 1): 51, ("b", 0): 0}
    units = infer.plan_units({"c": (list(range(40)), list(range(40)))}, "deblur", 8)
    assert infer.frame_bases(units) == {("c", k): 8 * k for k in range(4)}
