"""Per-stage clock64() breakdown of the tcgen05 pass A (debug_stage=9).  Run on the GPU box."""
import ctypes as C
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_io as gio

L = gio.pkg("host.lib")
sd, spec = gio.synthetic_checkpoint("gshift_deblur2")
eng = gio.pkg("host.engine").Engine(spec, {k: v.cuda() for k, v in sd.items()}, "cuda")
T, H, W = 20, 360, 640
x = (0.5 * torch.randn(T, H, W, 64, device="cuda")).half()
names = {True: ["start", "loads", "gather", "in-wait", "LN0", "LN1", "LN2", "LN3", "GEMM1", "tmem->G1", "dwA", "dwB", "GEMM2", "gate2", "store+sums"],
         False: ["start", "loads", "in-wait", "LN0", "LN1", "LN2", "LN3", "GEMM1", "tmem->G1", "dwA", "dwB", "GEMM2", "gate2", "store+sums"]}
PRE_NAMES = ["start", "GEMM1-wait", "tmem->G1", "dwA", "dwB", "GEMM2", "gate2", "store+sums"]
for p, mode, split, pre in (("stage1.encoder_level1.encoder_level1.0", L.MODE_CAB2_FWD, True, True),
                            ("stage1.encoder_level1.encoder_level1.1", L.MODE_CAB1, False, True),
                            ("stage1.encoder_level1.encoder_level1.0", L.MODE_CAB2_FWD, True, False),
                            ("stage1.encoder_level1.encoder_level1.1", L.MODE_CAB1, False, False)):
    shift = mode != L.MODE_CAB1
    blob = gio.pkg("host.packing").pack_cab_pass_a(eng.sd, p, 64, shift, 0)
    nt = eng.lib.gsn_cab_tiles(mode, H, W)
    z = torch.empty_like(x)
    part = torch.empty(T, nt, 64, device="cuda")
    dbg = torch.zeros(T * nt * 16, dtype=torch.int64, device="cuda")
    a = L.CabPassA()
    a.T, a.H, a.W, a.C, a.mode, a.circular = T, H, W, 64, mode, 1
    a.x, a.wblob, a.z, a.chan_partial = x.data_ptr(), blob.data_ptr(), z.data_ptr(), part.data_ptr()
    a.debug_stage, a.debug_out = 9, dbg.data_ptr()
    if split:
        wc1 = eng.sd[p + ".conv1.weight"].view(32, 9).t().contiguous().half()
        hw_pre = torch.empty(T, H, W, 32, dtype=torch.float16, device="cuda")
        L.check(eng.lib.gsn_shift_conv1(x.data_ptr(), T, H, W, 64, mode, 1, wc1.data_ptr(), hw_pre.data_ptr(), eng._stream()))
        a.hw_pre = hw_pre.data_ptr()
    if pre:
        ln = torch.cat((eng.sd[p + ".norm.weight"].float(), eng.sd[p + ".norm.bias"].float())).contiguous()
        a1 = torch.empty(T, 12 if shift else 8, H, W, 8, dtype=torch.float16, device="cuda")
        L.check(eng.lib.gsn_ln_planar(x.data_ptr(), a.hw_pre, T, H, W, 64, mode, 1, ln.data_ptr(), a1.data_ptr(), eng._stream()))
        a.a1_pre = a1.data_ptr()
    for _ in range(2):
        L.check(eng.lib.gsn_cab_pass_a(C.byref(a), eng._stream()))
    torch.cuda.synchronize()
    c = dbg.view(T * nt, 16).double()
    c = c[c[:, 0] > 0]          # persistent kernel: only the first tile of every CTA records its clocks
    nm = PRE_NAMES if pre else names[shift and not split]
    n = len(nm)
    d = (c[:, 1:n] - c[:, :n - 1])
    tot = (c[:, n - 1] - c[:, 0])
    print(f"mode={'shift' if shift else 'cab1'} pre={pre} split={split} tiles={T*nt} recorded={c.shape[0]} mean cycles/tile={tot.mean().item():.0f} (min {tot.min().item():.0f} max {tot.max().item():.0f})")
    for i in range(n - 1):
        print(f"   {nm[i+1]:12s} {d[:, i].mean().item():8.0f}  ({100 * d[:, i].mean().item() / tot.mean().item():4.1f}%)")
