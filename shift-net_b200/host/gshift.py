"""Host-side mirror of the reference's ``GShiftNet`` nn.Module (the drop-in boundary).

Same constructor signature, same ``state_dict`` keys/shapes, same ``forward`` contract as
basicsr/models/archs/gshift_{deblur,denoise}{1,2}.py::GShiftNet -- but ``forward`` runs on the sm_100a kernels of
``csrc/`` through the C-ABI.  There is no CPU or eager fallback: a CPU tensor or a missing extension raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .archspec import ARCHS
from .params import build_param_tree


class GShiftNetB200(nn.Module):
    arch = None  # set by subclasses

    def __init__(self, n_features=48, future_frames=None, past_frames=None):
        super().__init__()
        spec = ARCHS[self.arch]
        self.spec = spec
        self.n_feats = n_features
        self.num_ff = spec.default_ctx if future_frames is None else future_frames
        self.num_fb = spec.default_ctx if past_frames is None else past_frames
        build_param_tree(self, spec)
        self._engine = None
        self._graphs = {}
        self.kernel_launches = 0     # kernels of libshiftnet_b200 executed by this module (eager launches + graph replays)
        # Replay the whole forward (~740 kernel launches) as one CUDA graph per input shape: removes the launch gaps
        # between the many short kernels.  Set GSN_CUDA_GRAPH=0 to launch eagerly (profiling per-kernel, debugging).
        import os
        self.use_cuda_graph = os.environ.get("GSN_CUDA_GRAPH", "1") != "0"

    # any change of the parameters invalidates the packed device weights
    def _apply(self, fn, *a, **k):
        self._engine, self._graphs = None, {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine, self._graphs = None, {}
        return super().load_state_dict(*a, **k)

    def engine(self):
        from .engine import Engine
        dev = next(self.parameters()).device
        if self._engine is None:
            self._engine = Engine(self.spec, self.state_dict(), dev)
        return self._engine

    @torch.no_grad()
    def forward(self, x, noise_map=None, k1=None, k2=None, k3=None):
        if not x.is_cuda:
            raise RuntimeError("shiftnet_b200.GShiftNet runs on CUDA (B200, sm_100a) only; there is no CPU fallback")
        if next(self.parameters()).device != x.device:
            raise RuntimeError("model parameters and input are on different devices; call net.to(device) first")
        eng = self.engine()
        lib = eng.lib
        if not self.use_cuda_graph or eng.timeline is not None or torch.cuda.is_current_stream_capturing():
            l0 = lib.gsn_launch_count()
            out = eng.forward(x, noise_map, past=self.num_fb, future=self.num_ff)
            self.kernel_launches += lib.gsn_launch_count() - l0
            return out
        key = (tuple(x.shape), x.dtype, None if noise_map is None else tuple(noise_map.shape), self.num_fb, self.num_ff)
        ent = self._graphs.get(key)
        if ent is None:
            # eager warm-up (packs weights, sets kernel attributes, primes the allocator), then capture
            xs = x.clone()
            ns = None if noise_map is None else noise_map.expand(noise_map.shape).clone()
            l0 = lib.gsn_launch_count()
            eng.forward(xs, ns, past=self.num_fb, future=self.num_ff)
            self.kernel_launches += lib.gsn_launch_count() - l0
            torch.cuda.synchronize(x.device)
            g = torch.cuda.CUDAGraph()
            l0 = lib.gsn_launch_count()
            with torch.cuda.graph(g):
                out_s = eng.forward(xs, ns, past=self.num_fb, future=self.num_ff)
            ent = (g, xs, ns, out_s, lib.gsn_launch_count() - l0)   # kernel nodes recorded in the graph
            self._graphs = {key: ent}            # keep one shape at a time (activations of a 720p clip are several GB)
        g, xs, ns, out_s, n_kernels = ent
        xs.copy_(x)
        if ns is not None:
            ns.copy_(noise_map)
        g.replay()
        self.kernel_launches += n_kernels
        return out_s.clone()


    @torch.no_grad()
    def forward_tsharded(self, x_local, tshard, noise_map=None):
        """One slice of a T-sharded clip (host/tshard.py): x_local (1, n_local, 3, H, W) are the frames this rank owns (noise_map
        (1, n_local, 1, H, W) for the denoise nets); returns the restored frames among them (the clip's context frames are cropped
        on the ranks that hold them).  COLLECTIVE over the ranks of `tshard`.  With more than one rank the forward is replayed as a
        chain of CUDA graphs cut at the halo exchanges (host/tshard.py SegmentedGraph; GSN_TSHARD_GRAPH=0 or GSN_CUDA_GRAPH=0
        launch eagerly); the NCCL point-to-point calls run between the graphs."""
        if not x_local.is_cuda:
            raise RuntimeError("shiftnet_b200.GShiftNet runs on CUDA (B200, sm_100a) only; there is no CPU fallback")
        eng = self.engine()
        lib = eng.lib
        import os
        graphed = (self.use_cuda_graph and tshard.world > 1 and os.environ.get("GSN_TSHARD_GRAPH", "1") != "0"
                   and eng.timeline is None and not torch.cuda.is_current_stream_capturing())
        if not graphed:
            return self._tsharded_eager(eng, x_local, tshard, noise_map)
        key = ("tshard", tuple(x_local.shape), x_local.dtype, None if noise_map is None else tuple(noise_map.shape), self.num_fb,
               self.num_ff, tshard.rank, tshard.world, tshard.T)
        ent = self._graphs.get(key)
        if ent is None:
            from .tshard import SegmentedGraph
            xs = x_local.clone()
            ns = None if noise_map is None else noise_map.expand(noise_map.shape).clone()
            self._tsharded_eager(eng, xs, tshard, ns)          # eager warm-up, halo exchanges included: packs weights, primes the allocator
            torch.cuda.synchronize(x_local.device)
            seg = SegmentedGraph(tshard)
            side = torch.cuda.Stream(x_local.device)
            side.wait_stream(torch.cuda.current_stream(x_local.device))
            eng.tshard, tshard.recorder = tshard, seg
            l0 = lib.gsn_launch_count()
            try:
                with torch.cuda.stream(side):
                    seg.begin()
                    out_s = eng.forward(xs, ns, past=self.num_fb, future=self.num_ff)
                    seg.end()
            finally:
                eng.tshard, tshard.recorder = None, None
            torch.cuda.current_stream(x_local.device).wait_stream(side)
            torch.cuda.synchronize(x_local.device)
            ent = (seg, xs, ns, out_s, lib.gsn_launch_count() - l0)
            self._graphs = {key: ent}
        seg, xs, ns, out_s, n_kernels = ent
        seg.ts = tshard              # same (rank, world, T) by the key: statistics and the process group are the caller's
        xs.copy_(x_local)
        if ns is not None:
            ns.copy_(noise_map)
        seg.replay()
        self.kernel_launches += n_kernels
        return out_s.clone()

    def _tsharded_eager(self, eng, x_local, tshard, noise_map):
        lib = eng.lib
        eng.tshard = tshard
        try:
            l0 = lib.gsn_launch_count()
            out = eng.forward(x_local, noise_map, past=self.num_fb, future=self.num_ff)
            self.kernel_launches += lib.gsn_launch_count() - l0
        finally:
            eng.tshard = None
        return out


def make_arch(arch_name):
    return type("GShiftNet", (GShiftNetB200,), {"arch": arch_name, "__doc__": GShiftNetB200.__doc__})
