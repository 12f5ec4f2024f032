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

    # any change of the parameters invalidates the packed device weights
    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
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
        return self.engine().forward(x, noise_map, past=self.num_fb, future=self.num_ff)


def make_arch(arch_name):
    return type("GShiftNet", (GShiftNetB200,), {"arch": arch_name, "__doc__": GShiftNetB200.__doc__})
