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


def _halo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ts_mod = gio.pkg("host.tshard")
    T, C = 11, 8
    ts = ts_mod.TShard(rank, world, T)
    n = ts.n_local
    # frame f of the global clip holds the value 100 f + channel
    glob = (100.0 * torch.arange(T).view(T, 1, 1, 1) + torch.arange(C).view(1, 1, 1, C)).expand(T, 2, 3, C).contiguous()
    full = torch.full((n + 1, 2, 3, C), -1.0)
    full[:n] = glob[ts.a:ts.b]
    ts.halo_into(full, n, reverse=False)
    fwd = full[n].clone()
    ts.halo_into(full, n, reverse=True)
    rev = full[n].clone()
    # the clamped roll: an open chain, the end ranks receive nothing for the step whose clip end they hold
    full[n] = -1.0
    ts.halo_into(full, n, reverse=False, circular=False)
    cf = full[n, 0, 0].tolist()
    full[n] = -1.0
    ts.halo_into(full, n, reverse=True, circular=False)
    cr = full[n, 0, 0].tolist()
    q.put((rank, ts.a, ts.b, fwd[0, 0].tolist(), rev[0, 0].tolist(), ts.local_output_range(2, 2), ts.halo_bytes, cf, cr,
           (ts.needs_halo(False, False), ts.needs_halo(True, False))))
    dist.barrier()
    dist.destroy_process_group()


def test_tshard_halo_ring_gloo_world3():
    """T-sharded single-clip mode (host/tshard.py): the forward halo is the HIGH half of the previous rank's last frame, the reverse
    halo the LOW half of the next rank's first frame, both around the ring (the clip's circular wrap); the final crop of the
    clip's context frames lands on the ranks that hold them."""
    world, T, C = 3, 11, 8
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    spans = [(g[1], g[2]) for g in got]
    assert spans == [(0, 4), (4, 8), (8, 11)]
    kept = 0
    for rank, a, b, fwd, rev, (lo, hi), nbytes, cf, cr, need in got:
        prev_last, next_first = (a - 1) % T, b % T
        assert fwd[C // 2:] == [100.0 * prev_last + c for c in range(C // 2, C)]          # forward: high half of frame a-1
        assert rev[:C // 2] == [100.0 * next_first + c for c in range(C // 2)]             # reverse: low half of frame b
        assert need == (rank > 0, rank < world - 1)
        assert cf[C // 2:] == ([100.0 * (a - 1) + c for c in range(C // 2, C)] if rank > 0 else [-1.0] * (C // 2))
        assert cr[:C // 2] == ([100.0 * b + c for c in range(C // 2)] if rank < world - 1 else [-1.0] * (C // 2))
        kept += hi - lo
        assert [a + i for i in range(lo, hi)] == [f for f in range(a, b) if 2 <= f < T - 2]
    assert kept == T - 4


def test_tshard_single_rank_ring_is_the_circular_wrap():
    ts = gio.pkg("host.tshard").TShard(0, 1, 5)
    full = torch.arange(6 * 1 * 1 * 4, dtype=torch.float32).view(6, 1, 1, 4)
    full[5] = -1
    ts.halo_into(full, 5, reverse=False)
    assert full[5, 0, 0].tolist() == [-1, -1, 18, 19]          # high half of its own last frame (frame 4)
    ts.halo_into(full, 5, reverse=True)
    assert full[5, 0, 0].tolist() == [0, 1, 18, 19]            # low half of its own first frame


def _frame_roll_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, gio.ROOT)
    from oracle import shiftnet_oracle as O
    T, C = 11, 8
    ts = gio.pkg("host.tshard").TShard(rank, world, T)
    g = torch.Generator().manual_seed(5)
    glob = torch.randn(T, 2, 3, C, generator=g)                       # NHWC frames of the whole clip
    ok = []
    for reverse in (False, True):
        want = O.temporal_roll(glob.permute(0, 3, 1, 2), reverse, circular=False)[0].permute(0, 2, 3, 1)[ts.a:ts.b]
        xf, own = ts.with_neighbour_frame(glob[ts.a:ts.b].contiguous(), reverse)
        have = O.temporal_roll(xf.permute(0, 3, 1, 2), reverse, circular=False)[0].permute(0, 2, 3, 1)[own]
        ok.append((bool(torch.equal(have, want)), xf.shape[0] - ts.n_local))
    q.put((rank, ok, ts.exchanges))
    dist.barrier()
    dist.destroy_process_group()


def test_tshard_whole_frame_roll_of_shift_cab_gloo_world3():
    """gshift_denoise1's Shift_CABs (gshift_denoise1.py:167-179) in T-sharded mode: TShard.with_neighbour_frame puts the neighbour
    rank's boundary frame next to the own frames so that the clamped roll over n+1 frames equals the roll of the whole clip; the end
    rank of each direction gets nothing and clamps.  Every rank calls the exchange (a rank that skipped it would deadlock NCCL)."""
    world = 3
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_frame_roll_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    for rank, ok, nex in got:
        assert nex == 2
        assert ok[0] == (True, 1 if rank > 0 else 0)                  # forward: everyone but the first rank receives a frame
        assert ok[1] == (True, 1 if rank < world - 1 else 0)          # reverse: everyone but the last rank


class _FakeGraph:
    """Stands in for torch.cuda.CUDAGraph on CPU: records the order of capture / replay calls."""
    log = []

    def capture_begin(self, pool=None, capture_error_mode="global"):
        _FakeGraph.log.append("begin")

    def capture_end(self):
        _FakeGraph.log.append("end")

    def replay(self):
        _FakeGraph.log.append("replay")


def _segmented_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ts_mod = gio.pkg("host.tshard")
    torch.cuda.CUDAGraph, torch.cuda.graph_pool_handle = _FakeGraph, (lambda: 0)      # this process only
    ts = ts_mod.TShard(rank, world, 9)
    seg = ts_mod.SegmentedGraph(ts)
    points = [(False, True), (True, True), (False, False), (True, False)]               # (reverse, circular) of four exchange points
    # capture pass: no communication, every exchange cuts the segment and hands out a static receive buffer
    ts.recorder = seg
    seg.begin()
    sends, recvs = [], []
    for k, (rev, circ) in enumerate(points):
        s = torch.full((2, 3), float(100 * rank + k))
        sends.append(s)
        recvs.append(ts.exchange(s, rev, circ))
    seg.end()
    ts.recorder = None
    captured = list(_FakeGraph.log)
    # replay twice with fresh data in the static send buffers
    out = []
    for it in range(2):
        for k, s in enumerate(sends):
            s.fill_(1000.0 * it + 100 * rank + k)
        _FakeGraph.log.clear()
        ts.exchanges = 0
        seg.replay()
        out.append([None if r is None else float(r[0, 0]) for r in recvs])
    q.put((rank, captured, list(_FakeGraph.log), ts.exchanges, out))
    dist.barrier()
    dist.destroy_process_group()


def test_segmented_graph_cuts_at_every_exchange_and_replays_the_p2p_steps_gloo_world3():
    """host/tshard.py SegmentedGraph (the T-sharded forward as CUDA graphs cut at the halo exchanges), protocol only, with stand-in
    graph objects: the capture pass communicates nothing and cuts one segment per exchange on EVERY rank (also where a clamped
    clip end neither sends nor receives); a replay runs graph, p2p, graph, p2p, ... in order and delivers the neighbour's current
    send buffer into the static receive buffer."""
    world = 3
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_segmented_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    for rank, captured, replayed, nex, out in got:
        assert captured == ["begin"] + ["end", "begin"] * 4 + ["end"]
        assert replayed == ["replay"] * 5 and nex == 4
        for it in range(2):
            base = 1000.0 * it
            prev, nxt = (rank - 1) % world, (rank + 1) % world
            want = [base + 100 * prev + 0,                                   # forward, ring: from rank-1
                    base + 100 * nxt + 1,                                    # reverse, ring: from rank+1
                    base + 100 * (rank - 1) + 2 if rank > 0 else None,       # forward, open chain
                    base + 100 * (rank + 1) + 3 if rank < world - 1 else None]
            assert out[it] == want, (rank, it, out[it], want)


def test_tshard_frame_split_properties():
    """split_frames / local_output_range over every (T, world) of interest: the ranks' ranges tile [0, T) in order, differ by at most
    one frame, the frames that survive the crop of the 2 + 2 context frames are exactly [2, T-2) with no overlap, and the open
    chain of a clamped roll has exactly world-1 receivers per direction."""
    ts_mod = gio.pkg("host.tshard")
    for world in range(1, 9):
        for T in range(max(world, 5), 70):
            shards = [ts_mod.TShard(r, world, T) for r in range(world)]
            assert shards[0].a == 0 and shards[-1].b == T
            assert all(shards[i].b == shards[i + 1].a for i in range(world - 1))
            sizes = [s.n_local for s in shards]
            assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1
            kept = []
            for s in shards:
                lo, hi = s.local_output_range(2, 2)
                assert 0 <= lo <= hi <= s.n_local
                kept += [s.a + i for i in range(lo, hi)]
            assert kept == list(range(2, T - 2))
            for rev in (False, True):
                assert sum(s.needs_halo(rev, False) for s in shards) == world - 1
                assert all(s.needs_halo(rev, True) for s in shards)
    with pytest.raises(ValueError):
        ts_mod.TShard(0, 4, 3)
    with pytest.raises(ValueError):
        ts_mod.TShard(4, 4, 20)
