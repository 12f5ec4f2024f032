"""ctypes binding of libshiftnet_b200.so (the C-ABI declared in include/shiftnet_b200.h).

There is NO fallback: if the CUDA library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libshiftnet_b200.so")

EXPORTS = [
    "gsn_version", "gsn_last_error", "gsn_launch_count", "gsn_conv_tiles", "gsn_conv_mma", "gsn_conv3x3_tc_tiles", "gsn_conv3x3_tc", "gsn_conv_in", "gsn_conv_in_nm",
    "gsn_conv_out", "gsn_ca_scale", "gsn_scale_residual", "gsn_upsample2x_add", "gsn_add", "gsn_cab_tiles",
    "gsn_cab_pass_a", "gsn_cab_pass_a_tiles", "gsn_cab_fold", "gsn_cab_pass_b", "gsn_cab_fold_mid", "gsn_cab_tiles_linear", "gsn_cab_pass_a2", "gsn_pw_gate_tc",
    "gsn_shift_conv1", "gsn_shift_conv1_ln", "gsn_ln_planar", "gsn_shift_ln", "gsn_ln_pw", "gsn_ln_pw_tc", "gsn_dw_gate", "gsn_gate2", "gsn_group_conv5", "gsn_roll_copy",
    "gsn_cab_dense_tiles", "gsn_cab_dense", "gsn_u8_to_clip", "gsn_psnr_sse_blocks", "gsn_psnr_sse", "gsn_ssim_blocks", "gsn_ssim_workspace_bytes", "gsn_ssim",
]

MODE_CAB1, MODE_CAB2_FWD, MODE_CAB2_REV = 0, 1, 2
DTYPE_F16, DTYPE_F32 = 0, 1
PASS_A_FORCE_STREAM = 100      # GsnCabPassA.debug_stage: GSN_PASS_A_FORCE_STREAM
ROLL_CLAMP, ROLL_WRAP, ROLL_HALO = 0, 1, 2      # include/shiftnet_b200.h GSN_ROLL_*


class ConvDesc(C.Structure):
    _fields_ = [
        ("T", C.c_int), ("Hin", C.c_int), ("Win", C.c_int), ("Hout", C.c_int), ("Wout", C.c_int),
        ("n_src", C.c_int), ("src", C.c_void_p * 3), ("src_c", C.c_int * 3),
        ("cin_p", C.c_int), ("cout_p", C.c_int), ("ks", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
        ("wpack", C.c_void_p), ("bias", C.c_void_p), ("has_prelu", C.c_int), ("prelu_slope", C.c_float),
        ("residual", C.c_void_p), ("pixel_shuffle", C.c_int), ("chan_partial", C.c_void_p), ("dst", C.c_void_p),
        ("dst_c", C.c_int),
    ]


class CabDense(C.Structure):
    _fields_ = [
        ("T", C.c_int), ("H", C.c_int), ("W", C.c_int), ("cp", C.c_int), ("x", C.c_void_p), ("w1pack", C.c_void_p),
        ("w2pack", C.c_void_p), ("bias1", C.c_void_p), ("bias2", C.c_void_p), ("has_prelu", C.c_int),
        ("prelu_slope", C.c_float), ("r", C.c_void_p), ("chan_partial", C.c_void_p),
    ]


class CabPassA(C.Structure):
    _fields_ = [
        ("T", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("mode", C.c_int), ("circular", C.c_int),
        ("x", C.c_void_p), ("wblob", C.c_void_p), ("z", C.c_void_p), ("chan_partial", C.c_void_p),
        ("debug_stage", C.c_int), ("debug_out", C.c_void_p), ("mid_ca", C.c_int),
        ("hw_pre", C.c_void_p), ("a1_pre", C.c_void_p),
    ]


class CabPassB(C.Structure):
    _fields_ = [
        ("T", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("mode", C.c_int), ("circular", C.c_int),
        ("x", C.c_void_p), ("z", C.c_void_p), ("weff", C.c_void_p), ("beff", C.c_void_p), ("out", C.c_void_p),
        ("ln_next", C.c_void_p), ("a1_next", C.c_void_p),
    ]


_lib = None


def load():
    """dlopen the library (no CUDA context is created by loading)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"shiftnet_b200: CUDA extension {LIB_PATH} is missing -- build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i, f, ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
    lib.gsn_version.restype = i
    lib.gsn_last_error.restype = C.c_char_p
    lib.gsn_launch_count.restype = C.c_ulonglong
    lib.gsn_conv_tiles.argtypes = [i, i]
    lib.gsn_conv_mma.argtypes = [C.POINTER(ConvDesc), vp]
    lib.gsn_conv3x3_tc_tiles.argtypes = [i, i]
    lib.gsn_conv3x3_tc.argtypes = [C.POINTER(ConvDesc), vp]
    lib.gsn_conv_in.argtypes = [vp, i, i, i, i, i, vp, vp, i, vp, vp]
    lib.gsn_conv_in_nm.argtypes = [vp, i, i, i, i, i, vp, ll, ll, ll, vp, vp, i, vp, vp]
    lib.gsn_conv_out.argtypes = [vp, i, i, vp, vp, i, i, i, i, i, vp, vp]
    lib.gsn_ca_scale.argtypes = [vp, i, f, vp, vp, i, i, i, i, vp, vp]
    lib.gsn_scale_residual.argtypes = [vp, vp, vp, vp, vp, i, ll, i, vp]
    lib.gsn_upsample2x_add.argtypes = [vp, vp, vp, i, i, i, i, vp]
    lib.gsn_add.argtypes = [vp, vp, vp, ll, vp]
    lib.gsn_cab_tiles.argtypes = [i, i, i]
    lib.gsn_cab_pass_a_tiles.argtypes = [i, i, i, i, i, i]
    lib.gsn_cab_pass_a.argtypes = [C.POINTER(CabPassA), vp]
    lib.gsn_cab_fold.argtypes = [vp, i, f, vp, vp, i, vp, vp, vp, i, i, vp, vp, vp]
    lib.gsn_cab_pass_b.argtypes = [C.POINTER(CabPassB), vp]
    lib.gsn_cab_fold_mid.argtypes = [vp, i, f, vp, vp, i, vp, i, i, vp, vp]
    lib.gsn_cab_tiles_linear.argtypes = [ll]
    lib.gsn_cab_pass_a2.argtypes = [vp, vp, vp, vp, i, i, i, i, i, vp]
    lib.gsn_pw_gate_tc.argtypes = [vp, vp, vp, vp, i, i, i, i, vp]
    lib.gsn_shift_conv1.argtypes = [vp, i, i, i, i, i, i, vp, vp, vp]
    lib.gsn_shift_conv1_ln.argtypes = [vp, i, i, i, i, i, i, vp, vp, vp, vp]
    lib.gsn_ln_planar.argtypes = [vp, vp, i, i, i, i, i, i, vp, vp, vp]
    lib.gsn_shift_ln.argtypes = [vp, i, i, i, i, i, i, vp, vp, vp, i, vp, vp]
    lib.gsn_ln_pw.argtypes = [vp, vp, i, i, i, i, i, i, vp, vp, vp, vp, vp]
    lib.gsn_ln_pw_tc.argtypes = [vp, vp, i, i, i, i, i, i, vp, vp, vp, vp, vp]
    lib.gsn_dw_gate.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, vp]
    lib.gsn_gate2.argtypes = [vp, vp, i, i, i, i, vp, vp, vp]
    lib.gsn_group_conv5.argtypes = [vp, i, i, i, i, vp, vp, vp, vp]
    lib.gsn_roll_copy.argtypes = [vp, vp, i, i, i, i, i, i, vp]
    lib.gsn_cab_dense_tiles.argtypes = [i, i, i]
    lib.gsn_cab_dense.argtypes = [C.POINTER(CabDense), vp]
    lib.gsn_u8_to_clip.argtypes = [vp, i, i, i, i, vp, vp]
    lib.gsn_psnr_sse_blocks.argtypes = []
    lib.gsn_psnr_sse.argtypes = [vp, i, vp, i, i, i, vp, vp]
    lib.gsn_ssim_blocks.argtypes = []
    lib.gsn_ssim_workspace_bytes.argtypes = [i, i, i]
    lib.gsn_ssim_workspace_bytes.restype = ll
    lib.gsn_ssim.argtypes = [vp, i, vp, i, i, i, vp, vp, vp]
    for n in EXPORTS:
        getattr(lib, n)       # every symbol include/shiftnet_b200.h declares must resolve (AttributeError otherwise)
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().gsn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"shiftnet_b200 {what} failed (rc={rc}): {msg}")
