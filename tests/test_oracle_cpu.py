"""CPU suite: the oracle against the committed reference goldens, the host logic, and the C-ABI exports."""
import ctypes
import hashlib
import importlib
import os
import re
import sys

import numpy as np
import pytest
import torch

import golden_io as gio

sys.path.insert(0, os.path.join(gio.ROOT, "oracle"))
import shiftnet_oracle as O  # noqa: E402

torch.set_grad_enabled(False)


@pytest.mark.parametrize("arch", gio.ARCH_NAMES)
def test_state_dict_keys_match_reference(arch):
    sd, _ = gio.synthetic_checkpoint(arch)
    ref = gio.load_keys(arch)
    assert {k: list(v.shape) for k, v in sd.items()} == ref


@pytest.mark.parametrize("arch", gio.ARCH_NAMES)
def test_shift_index_map_bit_exact(arch):
    """channel_shift fwd/rev is pure data movement: sha256 of the fp32 bytes must equal the reference's."""
    _, spec = gio.synthetic_checkpoint(arch)
    g = gio.load_golden(arch)
    x = gio.module_input("shift", spec)
    for name, rev in (("shift_fwd", False), ("shift_rev", True)):
        y = O.channel_shift(x, rev, O.ARCHS[arch].circular).contiguous().numpy().astype(np.float32)
        assert list(y.shape) == list(g[name + "_shape"])
        assert hashlib.sha256(y.tobytes()).digest() == bytes(g[name + "_sha256"])


@pytest.mark.parametrize("arch", gio.ARCH_NAMES)
def test_oracle_modules_match_reference(arch):
    sd, spec = gio.synthetic_checkpoint(arch)
    osp = O.ARCHS[arch]
    g = gio.load_golden(arch)
    xs = gio.module_input("shift", spec)
    blk = "stage1.decoder_level1"
    c = spec.c1
    got = {
        "cab2_fwd": O.cab2(sd, blk + ".encoder_level1.0", O.channel_shift(xs, False, osp.circular), c, osp.denoise),
        "cab2_rev": O.cab2(sd, blk + ".encoder_level1_1.0", O.channel_shift(xs, True, osp.circular), c, osp.denoise),
        "cab1": O.cab1(sd, blk + ".encoder_level1.1", xs, osp.denoise),
        "block": O.shift_block(sd, blk, xs, osp),
        "cab": O.cab(sd, "feat_extract.1", gio.module_input("cab", spec)),
        "tfr": O.tfr_unet(sd, "orb1", gio.module_input("tfr", spec), osp),
        "stage1": O.stage1(sd, "stage1", gio.module_input("stage1", spec), osp),
    }
    for k, v in got.items():
        ref = torch.from_numpy(g[k])
        assert v.shape == ref.shape, k
        # same torch build => bit-exact here; allow a few ulp for other CPUs / oneDNN paths
        assert torch.allclose(v, ref, rtol=1e-4, atol=1e-5), (k, (v - ref).abs().max().item())


@pytest.mark.parametrize("arch", gio.ARCH_NAMES)
def test_oracle_full_forward_matches_reference(arch):
    sd, spec = gio.synthetic_checkpoint(arch)
    g = gio.load_golden(arch)
    x, nm = gio.clip_input(spec)
    out = O.gshiftnet_forward(sd, O.ARCHS[arch], x, nm)
    ref = torch.from_numpy(g["full"])
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5), (out - ref).abs().max().item()


def test_shift_offsets_table():
    for c, n_single_outer, n_inner in ((64, 1, 2), (80, 2, 1)):
        offs = O.shift_offsets(c)
        assert len(offs) == c // 2
        assert offs[0] == (8, 8) and offs[-1] == (-4, -4)
        assert all(abs(dy) in (0, 4, 8) and abs(dx) in (0, 4, 8) and (dy, dx) != (0, 0) for dy, dx in offs)
        assert len(set(offs)) == 24


def test_roll_clamped_vs_circular_edge_frames():
    x = torch.arange(3 * 4 * 2 * 2, dtype=torch.float32).view(3, 4, 2, 2)
    yc, _ = O.temporal_roll(x, False, True)
    yn, _ = O.temporal_roll(x, False, False)
    assert torch.equal(yc[0, :2], x[2, 2:]) and torch.equal(yn[0], x[0])      # wrap vs un-swapped frame 0
    assert torch.equal(yc[1:], yn[1:])
    yc, _ = O.temporal_roll(x, True, True)
    yn, _ = O.temporal_roll(x, True, False)
    assert torch.equal(yc[2, 2:], x[0, :2]) and torch.equal(yn[2], x[2])
    assert torch.equal(yc[:2], yn[:2])


def test_cabi_exports_every_declared_symbol():
    lib_mod = importlib.import_module("shift-net_b200.host.lib")
    header = open(os.path.join(gio.ROOT, "include", "shiftnet_b200.h")).read()
    declared = set(re.findall(r"\b(gsn_[a-z0-9_]+)\s*\(", header))
    assert declared == set(lib_mod.EXPORTS), declared ^ set(lib_mod.EXPORTS)
    if not os.path.exists(lib_mod.LIB_PATH):
        pytest.skip("library not built (run __graft_entry__.build())")
    handle = ctypes.CDLL(lib_mod.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.gsn_version() >= 100


def test_ctypes_structs_match_the_header(tmp_path):
    """The descriptor structs of host/lib.py must have the size and field offsets the C compiler gives include/shiftnet_b200.h
    (a silently shifted field would send garbage pointers to the kernels)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    lib_mod = importlib.import_module("shift-net_b200.host.lib")
    pairs = [("GsnConvDesc", lib_mod.ConvDesc), ("GsnCabDense", lib_mod.CabDense), ("GsnCabPassA", lib_mod.CabPassA),
             ("GsnCabPassB", lib_mod.CabPassB)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "shiftnet_b200.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-I", os.path.join(gio.ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in pairs:
        assert int(got[cname]) == ctypes.sizeof(cls), (cname, got[cname], ctypes.sizeof(cls))
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_python_constants_match_the_header():
    """Modes, dtypes, roll rules and the kernel selector of host/lib.py are the header's #defines."""
    lib_mod = importlib.import_module("shift-net_b200.host.lib")
    header = open(os.path.join(gio.ROOT, "include", "shiftnet_b200.h")).read()
    macros = {m.group(1): int(m.group(2)) for m in re.finditer(r"^#define\s+(GSN_[A-Z0-9_]+)\s+(-?\d+)\s*$", header, re.M)}
    for py, c in (("MODE_CAB1", "GSN_MODE_CAB1"), ("MODE_CAB2_FWD", "GSN_MODE_CAB2_FWD"), ("MODE_CAB2_REV", "GSN_MODE_CAB2_REV"),
                  ("ROLL_CLAMP", "GSN_ROLL_CLAMP"), ("ROLL_WRAP", "GSN_ROLL_WRAP"), ("ROLL_HALO", "GSN_ROLL_HALO"),
                  ("PASS_A_FORCE_STREAM", "GSN_PASS_A_FORCE_STREAM")):
        assert c in macros and getattr(lib_mod, py) == macros[c], (py, c)


def test_product_path_never_imports_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(gio.ROOT, "shift-net_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                if "oracle" in open(os.path.join(d, f), errors="ignore").read().replace("no oracle", ""):
                    bad.append(f)
    for d, _, files in os.walk(os.path.join(gio.ROOT, "basicsr")):
        for f in files:
            if f.endswith(".py") and "oracle" in open(os.path.join(d, f)).read():
                bad.append(f)
    assert not bad, bad


def test_cpu_input_fails_loudly():
    from basicsr.models.archs.gshift_deblur2 import GShiftNet
    net = GShiftNet(future_frames=2, past_frames=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 6, 3, 16, 16))


def test_weight_packing_layouts():
    P = gio.pkg("host.packing")
    w = torch.randn(14, 28, 3, 3)
    wp = P.pack_conv_mma(w, [14, 14], [16, 16], 16).view(9, 2, 2, 32, 4).float()
    # lane l (g=l>>2, tig=l&3) of n-tile nt, k-step ks holds W[nt*8+g][ks*16 + {2tig, 2tig+1, 2tig+8, 2tig+9}]
    for tap, ks, nt, lane in ((0, 0, 0, 0), (4, 1, 1, 13), (8, 0, 1, 31)):
        g_, tig = lane >> 2, lane & 3
        n = nt * 8 + g_
        for j, kk in enumerate((2 * tig, 2 * tig + 1, 2 * tig + 8, 2 * tig + 9)):
            kp = ks * 16 + kk           # padded input channel -> (source, real channel)
            src, c = divmod(kp, 16)
            exp = w[n, src * 14 + c, tap // 3, tap % 3].half().float() if (n < 14 and c < 14) else torch.tensor(0.0)
            assert wp[tap, ks, nt, lane, j] == exp
    pc = P.planar_chunks(torch.arange(16 * 24, dtype=torch.float32).view(16, 24))
    assert pc.shape == (3, 16, 8) and pc[2, 5, 3] == 5 * 24 + 2 * 8 + 3


@pytest.mark.parametrize("cp,cin", [(16, 14), (24, 18), (24, 22)])
def test_dense_fragment_packing(cp, cin):
    """pack_dense_frag (csrc/cab_dense.cu): rebuilding B[k][n] per tap from the per-lane mma fragment words gives the conv
    weight transposed (k = input channel, n = output channel), zero in the padding rows/columns."""
    P = gio.pkg("host.packing")
    w = torch.randn(cin, cin, 3, 3, generator=torch.Generator().manual_seed(3))
    f = P.pack_dense_frag(w, cp).float().view(9, cp // 8, -1, 32, 2)      # tap, n-tile, word, lane, element
    k16, k8 = cp // 16, (cp % 16) // 8
    assert f.shape[2] == 2 * k16 + k8
    B = torch.zeros(9, cp, cp)
    for lane in range(32):
        g_, tig = lane >> 2, lane & 3
        for nt in range(cp // 8):
            for wd in range(f.shape[2]):          # words 2s, 2s+1: k = 16s + {0, 8} + 2 tig + e ; tail word: k = 16 k16 + 2 tig + e
                k0 = 16 * (wd // 2) + 8 * (wd % 2) + 2 * tig if wd < 2 * k16 else 16 * k16 + 2 * tig
                for e in range(2):
                    B[:, k0 + e, nt * 8 + g_] = f[:, nt, wd, lane, e]
    ref = torch.zeros(9, cp, cp)
    ref[:, :cin, :cin] = w.half().float().reshape(cin, cin, 9).permute(2, 1, 0)
    assert torch.equal(B, ref)
