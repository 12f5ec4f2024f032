"""Generate the golden fixtures in this directory from the REFERENCE classes.

Run in the build container only (needs /root/reference; the GPU box has no reference):
    python tests/golden/make_golden.py
The reference arch files are loaded *by file path* (their package import needs lmdb, see
SURVEY.md section 8c).  Weights: our parameter tree (identical keys/shapes is asserted here)
randomised by host/synth.py seed 1234, loaded strictly into the reference net.
Inputs are regenerated from seeds by the tests (tests/golden_io.py), only outputs are stored.
"""
import gzip
import hashlib
import importlib
import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_io as gio  # noqa: E402

REF = "/root/reference/basicsr/models/archs"


def load_ref(name):
    spec = importlib.util.spec_from_file_location(name, f"{REF}/{name}.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    torch.set_grad_enabled(False)
    for arch in gio.ARCH_NAMES:
        sd, spec = gio.synthetic_checkpoint(arch)
        ref = load_ref(arch).GShiftNet(future_frames=2, past_frames=2).eval()
        rk = {k: list(v.shape) for k, v in ref.state_dict().items()}
        assert rk == {k: list(v.shape) for k, v in sd.items()}, "state_dict keys/shapes differ from reference"
        ref.load_state_dict(sd, strict=True)
        with gzip.open(os.path.join(HERE, f"keys_{arch}.json.gz"), "wt") as f:
            json.dump(rk, f)
        out = {}
        # whole net
        x, nm = gio.clip_input(spec)
        out["full"] = (ref(x, nm) if spec.denoise else ref(x)).numpy()
        # shift block pieces (block = stage1.decoder_level1)
        blk = ref.stage1.decoder_level1
        xs = gio.module_input("shift", spec)
        # pure index maps: stored as sha256 of the fp32 bytes (bit-exact check) + the shape
        for nm_, rev in (("shift_fwd", False), ("shift_rev", True)):
            s = blk.channel_shift(xs, reverse=rev).contiguous().numpy().astype(np.float32)
            out[nm_ + "_sha256"] = np.frombuffer(hashlib.sha256(s.tobytes()).digest(), dtype=np.uint8)
            out[nm_ + "_shape"] = np.array(s.shape, dtype=np.int64)
        out["cab2_fwd"] = blk.encoder_level1[0](blk.channel_shift(xs)).numpy()
        out["cab2_rev"] = blk.encoder_level1_1[0](blk.channel_shift(xs, reverse=True)).numpy()
        out["cab1"] = blk.encoder_level1[1](xs).numpy()
        out["block"] = blk(xs).numpy()
        out["cab"] = ref.feat_extract[1](gio.module_input("cab", spec)).numpy()
        out["tfr"] = ref.orb1(gio.module_input("tfr", spec)).numpy()
        out["stage1"] = ref.stage1(gio.module_input("stage1", spec)).numpy()
        np.savez_compressed(os.path.join(HERE, f"golden_{arch}.npz"), **{k: (v.astype(np.float32) if v.dtype.kind == "f" else v) for k, v in out.items()})
        print(arch, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
