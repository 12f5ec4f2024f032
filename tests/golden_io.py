"""Seeded inputs / checkpoints shared by make_golden.py and the tests (no reference access)."""
import gzip
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ARCH_NAMES = ["gshift_deblur2", "gshift_deblur1", "gshift_denoise2", "gshift_denoise1"]
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pkg(sub):
    return importlib.import_module("shift-net_b200." + sub)


def synthetic_checkpoint(arch, seed=1234):
    """(state_dict with reference key names, ArchSpec) -- CPU fp32."""
    spec = pkg("host.archspec").ARCHS[arch]
    net = torch.nn.Module()
    pkg("host.params").build_param_tree(net, spec)
    pkg("host.synth").randomize_(net, seed)
    return {k: v.detach().clone() for k, v in net.state_dict().items()}, spec


def clip_input(spec, T=6, H=32, W=40):
    synth = pkg("host.synth")
    if spec.denoise:
        _, x, nm = synth.synthetic_clip(T, H, W, denoise_sigma=30)
        return x, nm
    _, x = synth.synthetic_clip(T, H, W)
    return x, None


_SHAPES = {  # kind -> (T, channels-selector, H, W, seed)
    "shift": (4, "c1", 20, 24, 11),
    "cab": (3, "n0", 16, 20, 12),
    "tfr": (2, "n0", 16, 24, 13),
    "stage1": (4, "n0", 32, 40, 14),
}


def module_input(kind, spec):
    T, csel, H, W, seed = _SHAPES[kind]
    g = torch.Generator().manual_seed(seed)
    return 0.5 * torch.randn(T, getattr(spec, csel), H, W, generator=g)


def load_golden(arch):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"golden_{arch}.npz")))


def load_keys(arch):
    with gzip.open(os.path.join(GOLDEN_DIR, f"keys_{arch}.json.gz"), "rt") as f:
        return json.load(f)
