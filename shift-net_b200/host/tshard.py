"""T-sharded single-clip mode: ONE clip's frames are split across the ranks of a torchrun launch,
so a long / high-resolution clip uses all GPUs of the box instead of one.

Everything in GShiftNet.forward is per-frame except the half-channel temporal roll in front of every CAB2
(channel_shift, gshift_deblur2.py:499-519): frame t reads C/2 channels of frame t-1 (forward pairs) or t+1 (reverse pairs).
Each rank therefore needs, before every CAB2, ONE boundary frame's half of the channels from its ring neighbour: the halo
exchange below (NCCL send/recv over NVLink on GPUs; gloo in the CPU tests).  48 exchanges per forward for Ours-s, 14.7 MB each
at 720p level 1.  Where the roll wraps around the clip (Ours-s deblur, gshift_deblur2.py:504-505) the ring of ranks closes the wrap;
where it clamps at the clip ends (the other nets, gshift_deblur1.py:513,517) the chain stays open and the two end ranks run their
boundary step with the clamped rule.  The four Shift_CABs of gshift_denoise1's encoder (gshift_denoise1.py:167-186) roll whole
feature frames the same way and exchange one full boundary frame each (Engine.shift_cab).

The halo frame is stored BEHIND the rank's own frames, at index Tl of a (Tl+1)-frame buffer, and the CAB2 kernels run on the Tl
own frames with circular = GSN_ROLL_HALO (include/shiftnet_b200.h): their roll wraps over Tl+1 frames, so frame 0's predecessor is
index Tl and frame Tl-1's successor is index Tl; nothing is computed for the halo frame.
"""
from __future__ import annotations

import torch


def split_frames(T: int, world: int, rank: int):
    """Contiguous, balanced split of T frames: rank r owns [a, b)."""
    base, rem = divmod(T, world)
    a = rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


class TShard:
    def __init__(self, rank: int, world: int, T_global: int, group=None):
        if not (0 <= rank < world):
            raise ValueError(f"rank {rank} outside a group of {world}")
        if T_global < world:
            raise ValueError(f"a clip of {T_global} frames cannot be T-sharded over {world} ranks (every rank needs at least one frame)")
        self.rank, self.world, self.T, self.group = rank, world, T_global, group
        self.a, self.b = split_frames(T_global, world, rank)
        self.halo_bytes = 0          # bytes this rank SENT in halo exchanges (statistics for bench.py)
        self.exchanges = 0
        self.events = []             # (start, end) CUDA event pairs around the exchanges, when timing is on
        self.time_exchanges = False
        self.recorder = None         # a SegmentedGraph during its capture pass

    @property
    def n_local(self):
        return self.b - self.a

    def needs_halo(self, reverse: bool, circular: bool) -> bool:
        """Does this rank's CAB2 step of the given direction read a neighbour's frame?  Always with the wrapping roll; with the
        clamped roll (gshift_deblur1.py:513,517) the first rank's forward step / the last rank's reverse step keep their boundary
        frame un-swapped instead."""
        if circular:
            return True
        return self.rank < self.world - 1 if reverse else self.rank > 0

    def exchange(self, send: torch.Tensor, reverse: bool, circular: bool = True):
        """Exchange of one contiguous halo tensor along the chain of ranks (a ring when the roll wraps).  forward pairs: every rank
        sends to rank+1 and receives from rank-1; reverse pairs: sends to rank-1, receives from rank+1.  Returns the received
        tensor (same shape / dtype), or None on a rank whose clip end clamps.  COLLECTIVE: every rank of the group must call it for
        every exchange point, also the end ranks that only send or only receive."""
        self.exchanges += 1
        if self.world == 1:
            self.halo_bytes += send.numel() * send.element_size() if circular else 0
            return send if circular else None            # the ring of one rank: its own boundary frame is the wrap-around
        dst = self.rank + (-1 if reverse else 1)
        src = self.rank + (1 if reverse else -1)
        if circular:
            dst, src = dst % self.world, src % self.world
        dst = dst if 0 <= dst < self.world else None
        src = src if 0 <= src < self.world else None
        if self.recorder is not None:        # capture pass of a SegmentedGraph: the launches so far become one CUDA graph
            return self.recorder.cut(send, dst, src)
        recv = torch.empty_like(send) if src is not None else None
        self._p2p(send, recv, dst, src)
        return recv

    def _p2p(self, send, recv, dst, src):
        """One point-to-point step on the current stream: send -> rank dst, recv <- rank src (either may be None)."""
        import torch.distributed as dist
        ev = None
        if self.time_exchanges and send.is_cuda:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        ops = []
        if dst is not None:
            self.halo_bytes += send.numel() * send.element_size()
            ops.append(dist.P2POp(dist.isend, send, dst, self.group))
        if src is not None:
            ops.append(dist.P2POp(dist.irecv, recv, src, self.group))
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        if ev is not None:
            ev[1].record()
            self.events.append(ev)

    def halo_into(self, full: torch.Tensor, n: int, reverse: bool, circular: bool = True) -> None:
        """full: (n+1, H, W, C) NHWC buffer holding this rank's n frames in [:n]; fills the half of frame n that the CAB2 of the
        given direction reads from the neighbour: forward -> channels [C/2, C) of the previous rank's LAST frame,
        reverse -> channels [0, C/2) of the next rank's FIRST frame."""
        h = full.shape[-1] // 2
        if reverse:
            got = self.exchange(full[0, :, :, :h].contiguous(), True, circular)
            if got is not None:
                full[n, :, :, :h] = got
        else:
            got = self.exchange(full[n - 1, :, :, h:].contiguous(), False, circular)
            if got is not None:
                full[n, :, :, h:] = got

    def with_neighbour_frame(self, x: torch.Tensor, reverse: bool, alloc=None):
        """Whole-frame clamped roll across ranks (Shift_CAB.channel_shift, gshift_denoise1.py:167-179): x (n, H, W, C) are this
        rank's frames.  COLLECTIVE.  Returns (xf, own): xf = x with the neighbour's boundary frame in FRONT (forward: the previous
        rank's last frame) or BEHIND (reverse: the next rank's first frame) and own = the slice of xf / of the rolled result that
        belongs to this rank; the clamped roll over xf's n+1 frames then gives every own frame its true neighbour.  On the rank that
        holds the clip end of this direction nothing arrives and (x, slice(0, n)) comes back: its boundary frame clamps."""
        n = x.shape[0]
        got = self.exchange(x[0 if reverse else n - 1].contiguous(), reverse, False)
        if got is None:
            return x, slice(0, n)
        xf = alloc(n + 1, *x.shape[1:]) if alloc is not None else torch.empty((n + 1,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        own = slice(0, n) if reverse else slice(1, n + 1)
        xf[own].copy_(x)
        xf[n if reverse else 0].copy_(got)
        return xf, own

    def local_output_range(self, past: int, future: int):
        """Own frames that survive the net's final crop of `past` / `future` context frames of the GLOBAL clip, as a local slice."""
        lo = min(max(past - self.a, 0), self.n_local)
        hi = self.n_local - max(self.b - (self.T - future), 0)
        return lo, max(hi, lo)


class SegmentedGraph:
    """One T-sharded forward as a chain of CUDA graphs cut at the halo exchanges: the ~3 400 kernel launches of a forward replay as
    ~50 graph launches, and the NCCL point-to-point calls run eagerly between them on static send / receive buffers.  (The eager
    T-sharded path is host-launch-bound from two ranks on; NCCL calls stay outside the graphs on purpose.)  All segments share one
    memory pool and are replayed in capture order, so tensors that live across an exchange keep their addresses."""

    def __init__(self, ts: TShard):
        self.ts = ts
        self.graphs, self.cuts = [], []
        self.pool = torch.cuda.graph_pool_handle()
        self.cur = None

    def begin(self):
        self.cur = torch.cuda.CUDAGraph()
        self.cur.capture_begin(pool=self.pool, capture_error_mode="thread_local")   # the NCCL watchdog thread keeps polling its events

    def cut(self, send, dst, src):
        """Called by TShard.exchange during the capture pass: close the running segment, reserve the receive buffer, open the next."""
        self.cur.capture_end()
        self.graphs.append(self.cur)
        recv = torch.empty_like(send) if src is not None else None
        self.cuts.append((send, recv, dst, src))
        self.begin()
        return recv

    def end(self):
        self.cur.capture_end()
        self.graphs.append(self.cur)
        self.cur = None

    def replay(self):
        ts = self.ts
        for i, g in enumerate(self.graphs):
            g.replay()
            if i < len(self.cuts):
                ts.exchanges += 1
                ts._p2p(*self.cuts[i])
