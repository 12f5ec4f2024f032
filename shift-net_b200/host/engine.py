"""Forward orchestration of GShiftNet on the sm_100a kernels (csrc/), one ctypes call per kernel.

Mirrors the reference control flow (gshift_deblur2.py:587-613,682-695,731-756) but every tensor op is one of
our CUDA kernels; torch only provides device buffers and the current stream.  Activations are NHWC fp16 with
channels padded to a multiple of 16.  No CPU path: constructing an Engine requires CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import lib as L
from . import packing as P

_PAIRS = ["encoder_level1"] + [f"encoder_level1_{i}" for i in range(1, 8)]


class Engine:
    def __init__(self, spec, state_dict, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("shiftnet_b200: the engine runs on CUDA only (no CPU fallback)")
        self.spec = spec
        self.dev = torch.device(device)
        self.lib = L.load()
        # The checkpoint is packed on the HOST (host/packing.py on CPU fp32 tensors) and every packed blob reaches the device
        # by ONE memcpy: no torch kernel runs for weight handling, so the first launches of a fresh process are the
        # product's own kernels (a device-resident checkpoint comes back as one concatenated copy per dtype).
        self.sd = self._to_host(state_dict)
        self.cache = {}
        self.boff = 1 if spec.denoise else 0
        # The LayerNorm'd operand of every CAB1/CAB2 (k-chunk planar, landed by TMA in the tensor-core layout by pass A,
        # csrc/cab_pass_a_pre.cu) is emitted by its producer: CAB2's by the shift gather + conv1 kernel (gsn_shift_conv1_ln),
        # CAB1's by the epilogue of the preceding pass B (GsnCabPassB.a1_next).  GSN_LN_FUSE=0 runs the un-fused chain
        # gsn_shift_conv1 -> gsn_ln_planar instead (cross-check in the tests).
        self.ln_fuse = os.environ.get("GSN_LN_FUSE", "1") == "1"
        self.conv_tc = os.environ.get("GSN_CONV_TC", "1") == "1"       # Ours+: 40/48-channel CAB convs on tcgen05 (0: mma.sync conv_mma)
        self.ln_pw_tc = os.environ.get("GSN_LN_PW_TC", "1") == "1"     # Ours+: LayerNorm + first 1x1 on tcgen05 (0: mma.sync kernel)
        self.pass_a_stream = False     # True: route every C=64 deblur pass A to the row-streaming kernel (cab_pass_a_stream.cu)
        self.tshard = None             # host/tshard.py TShard: this engine holds only a slice of the clip's frames
        self._circ_override = None
        # HFMA2/HMUL2 thread-instructions per pixel of the 16x16-tile pass A (its two depthwise stages: 15.1 k warp-instructions per
        # 256-pixel tile incl. the halo recompute): bench.py reports the kernel against the FMA pipe with it
        self.pass_a_hfma2_per_pixel = 15.1e3 * 32 / 256
        self._a1_next = None
        # dense CAB bodies (16 / 24 stored channels): conv-PReLU-conv fused in one kernel (GSN_CAB_FUSED=0: two conv launches)
        self.cab_fused = os.environ.get("GSN_CAB_FUSED", "1") == "1"
        # optional per-kernel timing (bench.py's roofline leg): list of (name, pixels, start_event, end_event)
        self.timeline = None
        self.timeline_detail = os.environ.get("GSN_TIMELINE_DETAIL", "0") == "1"

    def _timed(self, name, pixels):
        """Context manager recording CUDA events around one launch on the launching stream (off unless profiling)."""
        eng = self

        class _T:
            def __enter__(self_inner):
                if eng.timeline is not None:
                    self_inner.a = torch.cuda.Event(enable_timing=True)
                    self_inner.b = torch.cuda.Event(enable_timing=True)
                    self_inner.a.record(torch.cuda.current_stream(eng.dev))

            def __exit__(self_inner, *exc):
                if eng.timeline is not None:
                    self_inner.b.record(torch.cuda.current_stream(eng.dev))
                    eng.timeline.append((name, pixels, self_inner.a, self_inner.b))
                return False

        return _T()

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _to_host(state_dict):
        out, groups = {}, {}
        for k, v in state_dict.items():
            v = v.detach()
            if v.is_cuda:
                groups.setdefault((v.device, v.dtype), []).append((k, v))
            else:
                out[k] = v.float()
        for items in groups.values():
            flat = torch.cat([v.reshape(-1) for _, v in items]).cpu().float()
            off = 0
            for k, v in items:
                out[k] = flat[off:off + v.numel()].view(v.shape)
                off += v.numel()
        return out

    def _up(self, t):
        """Packed host tensor (or dict of them) -> device, one memcpy each."""
        if t is None:
            return None
        if isinstance(t, dict):
            return {k: self._up(v) for k, v in t.items()}
        if isinstance(t, (tuple, list)):
            return tuple(self._up(v) for v in t)
        return t.contiguous().to(self.dev)

    def _circ(self):
        """Temporal roll wraps around the frames of the tensor at hand: the arch's rule (gshift_deblur2.py:504-505 wraps, the other three
        nets clamp), unless a T-sharded CAB2 step overrides it (a halo frame behind the rank's frames is reached by wrapping)."""
        ov = self._circ_override
        return (1 if self.spec.circular else 0) if ov is None else ov

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def _new(self, *shape, dtype=torch.float16):
        return torch.empty(*shape, dtype=dtype, device=self.dev)

    def _slope(self, key):
        v = self.cache.get(("slope", key))
        if v is None:
            v = float(self.sd[key].item())
            self.cache[("slope", key)] = v
        return v

    # ------------------------------------------------------------------ dense conv (tensor cores)
    def conv(self, key, srcs, src_real, cout, stride=1, pad=None, prelu_key=None, residual=None, pixel_shuffle=False,
             want_sums=False):
        """srcs: list of NHWC tensors (T,H,W,Cp); src_real: their real channel counts."""
        w = self.sd[key + ".weight"]
        ks = w.shape[-1]
        if pad is None:
            pad = ks // 2
        T, Hin, Win = srcs[0].shape[:3]
        Hout = (Hin + 2 * pad - ks) // stride + 1
        Wout = (Win + 2 * pad - ks) // stride + 1
        cout_p = P.pad16(cout) if pixel_shuffle else P.pad8(cout)   # shuffle: 4 conv channels per output channel
        src_pad = [s.shape[3] for s in srcs]
        ck = ("conv", key)
        if ck not in self.cache:
            wp = P.pack_conv_mma(w, src_real, src_pad, cout_p)
            b = self.sd.get(key + ".bias")
            self.cache[ck] = self._up((wp, P.pack_bias(b, cout_p) if b is not None else None))
        wp, bias = self.cache[ck]
        dst_c = 0
        if pixel_shuffle:
            dst_c = P.pad8(cout // 4)
            if dst_c == cout_p // 4:
                dst = self._new(T, 2 * Hout, 2 * Wout, dst_c)
            else:                                   # padding channels are never written by the shuffle store
                dst = torch.zeros(T, 2 * Hout, 2 * Wout, dst_c, dtype=torch.float16, device=self.dev)
        else:
            dst = self._new(T, Hout, Wout, cout_p)
        partial = None
        if want_sums:
            partial = self._new(T, self.lib.gsn_conv_tiles(Hout, Wout), cout_p, dtype=torch.float32)
        d = L.ConvDesc()
        d.T, d.Hin, d.Win, d.Hout, d.Wout = T, Hin, Win, Hout, Wout
        d.n_src = len(srcs)
        for i, s in enumerate(srcs):
            assert s.is_contiguous() and s.dtype == torch.float16
            d.src[i] = s.data_ptr()
            d.src_c[i] = s.shape[3]
        d.cin_p, d.cout_p, d.ks, d.stride, d.pad = P.pad16(sum(src_pad)), cout_p, ks, stride, pad
        d.wpack = wp.data_ptr()
        d.bias = bias.data_ptr() if bias is not None else None
        d.has_prelu = 1 if prelu_key else 0
        d.prelu_slope = self._slope(prelu_key) if prelu_key else 0.0
        d.residual = residual.data_ptr() if residual is not None else None
        d.pixel_shuffle = 1 if pixel_shuffle else 0
        d.chan_partial = partial.data_ptr() if partial is not None else None
        d.dst = dst.data_ptr()
        d.dst_c = dst_c
        with self._timed(f"conv_mma[{d.cin_p}>{cout_p} k{ks}s{stride} {Hout}x{Wout}]" if self.timeline_detail else "conv_mma", T * Hout * Wout):
            L.check(self.lib.gsn_conv_mma(C.byref(d), self._stream()), "conv_mma " + key)
        return (dst, partial) if want_sums else dst

    def conv3x3_tc(self, key, x, c, prelu_key=None, want_sums=False):
        """Same-width 3x3 / stride 1 conv of a 40- or 48-channel tensor on the tcgen05 implicit-GEMM kernel."""
        T, H, W, cp = x.shape
        ck = ("conv3tc", key)
        if ck not in self.cache:
            b = self.sd.get(key + ".bias")
            self.cache[ck] = self._up((P.pack_conv3x3_tc(self.sd[key + ".weight"]), P.pack_bias(b, 48) if b is not None else None))
        wp, bias = self.cache[ck]
        dst = self._new(T, H, W, cp)
        partial = self._new(T, self.lib.gsn_conv3x3_tc_tiles(H, W), cp, dtype=torch.float32) if want_sums else None
        d = L.ConvDesc()
        d.T, d.Hin, d.Win, d.Hout, d.Wout = T, H, W, H, W
        d.n_src = 1
        d.src[0], d.src_c[0] = x.data_ptr(), cp
        d.cin_p, d.cout_p, d.ks, d.stride, d.pad = 48, cp, 3, 1, 1
        d.wpack = wp.data_ptr()
        d.bias = bias.data_ptr() if bias is not None else None
        d.has_prelu = 1 if prelu_key else 0
        d.prelu_slope = self._slope(prelu_key) if prelu_key else 0.0
        d.chan_partial = partial.data_ptr() if partial is not None else None
        d.dst = dst.data_ptr()
        with self._timed(f"conv3x3_tc[{cp} {H}x{W}]" if self.timeline_detail else "conv3x3_tc", T * H * W):
            L.check(self.lib.gsn_conv3x3_tc(C.byref(d), self._stream()), "conv3x3_tc " + key)
        return (dst, partial) if want_sums else dst

    # ------------------------------------------------------------------ CAB (dense 3x3 + channel attention)
    def cab_body_fused(self, p, x, c):
        """conv3x3 -> PReLU -> conv3x3 of CAB.body (gshift_deblur2.py:146-150) in one kernel (csrc/cab_dense.cu)."""
        T, H, W, cp = x.shape
        ck = ("cab_dense", p)
        if ck not in self.cache:
            b1, b2 = self.sd.get(p + ".body.0.bias"), self.sd.get(p + ".body.2.bias")
            self.cache[ck] = self._up((P.pack_dense_frag(self.sd[p + ".body.0.weight"], cp), P.pack_dense_frag(self.sd[p + ".body.2.weight"], cp),
                                       P.pack_bias(b1, cp) if b1 is not None else None, P.pack_bias(b2, cp) if b2 is not None else None))
        w1, w2, b1, b2 = self.cache[ck]
        r = self._new(T, H, W, cp)
        partial = self._new(T, self.lib.gsn_cab_dense_tiles(cp, H, W), cp, dtype=torch.float32)
        d = L.CabDense()
        d.T, d.H, d.W, d.cp = T, H, W, cp
        d.x, d.w1pack, d.w2pack = x.data_ptr(), w1.data_ptr(), w2.data_ptr()
        d.bias1 = b1.data_ptr() if b1 is not None else None
        d.bias2 = b2.data_ptr() if b2 is not None else None
        d.has_prelu, d.prelu_slope = 1, self._slope(p + ".body.1.weight")
        d.r, d.chan_partial = r.data_ptr(), partial.data_ptr()
        with self._timed(f"cab_dense[{cp} {H}x{W}]" if self.timeline_detail else "cab_dense", T * H * W):
            L.check(self.lib.gsn_cab_dense(C.byref(d), self._stream()), "cab_dense " + p)
        return r, partial

    def cab(self, p, x, c, extra=None):
        """gshift_deblur2.py:143-158.  ``extra`` is an optional tensor added to the result (stage shortcuts)."""
        if self.cab_fused and x.shape[3] in (16, 24):
            r2, partial = self.cab_body_fused(p, x, c)
        elif self.conv_tc and x.shape[3] in (40, 48):      # Ours+ TFR_UNet levels 2 / 3: both 3x3 convs on tcgen05 (csrc/conv3x3_tc.cu)
            r1 = self.conv3x3_tc(p + ".body.0", x, c, prelu_key=p + ".body.1.weight")
            r2, partial = self.conv3x3_tc(p + ".body.2", r1, c, want_sums=True)
        else:
            r1 = self.conv(p + ".body.0", [x], [c], c, prelu_key=p + ".body.1.weight")
            r2, partial = self.conv(p + ".body.2", [r1], [c], c, want_sums=True)
        ck = ("ca", p)
        if ck not in self.cache:
            self.cache[ck] = self._up((self.sd[p + ".CA.conv_du.0.weight"].float().flatten(1),
                                       self.sd[p + ".CA.conv_du.2.weight"].float().flatten(1)))
        w1, w2 = self.cache[ck]
        T, H, W, cp = x.shape
        s = self._new(T, cp, dtype=torch.float32)
        L.check(self.lib.gsn_ca_scale(partial.data_ptr(), partial.shape[1], 1.0 / (H * W), w1.data_ptr(), w2.data_ptr(),
                                      c, w1.shape[0], cp, T, s.data_ptr(), self._stream()), "ca_scale")
        out = self._new(T, H, W, cp)
        with self._timed(f"scale_residual[{cp} {H}x{W}]" if self.timeline_detail else "scale_residual", T * H * W):
            L.check(self.lib.gsn_scale_residual(x.data_ptr(), r2.data_ptr(), s.data_ptr(),
                                                extra.data_ptr() if extra is not None else None, out.data_ptr(), T, H * W, cp,
                                                self._stream()), "scale_residual")
        return out

    def seq_cabs(self, p, x, c):
        i = 0
        while f"{p}.{i}.body.0.weight" in self.sd:
            x = self.cab(f"{p}.{i}", x, c)
            i += 1
        return x

    def downsample(self, p, x, cin, cout):
        if self.spec.denoise:
            return self.conv(p + ".down.0", [x], [cin], cout, stride=2, pad=1, prelu_key=p + ".down.1.weight")
        return self.conv(p + ".down", [x], [cin], cout, stride=2, pad=1)

    def skip_upsample(self, p, x, cin, skip, cout):
        """1x1 at low resolution, then bilinear x2 + skip (the two linear ops commute; gshift_deblur2.py:344-353)."""
        y = self.conv(p + ".up.1", [x], [cin], cout)
        T, h, w, cp = y.shape
        if tuple(skip.shape) != (T, 2 * h, 2 * w, cp):      # the reference's `x + y` raises here (gshift_deblur2.py:352)
            raise ValueError(f"skip_upsample {p}: skip {tuple(skip.shape)} does not match the upsampled {(T, 2 * h, 2 * w, cp)}")
        out = self._new(T, 2 * h, 2 * w, cp)
        L.check(self.lib.gsn_upsample2x_add(y.data_ptr(), skip.data_ptr(), out.data_ptr(), T, h, w, cp, self._stream()),
                "upsample2x_add")
        return out

    def add(self, a, b):
        out = torch.empty_like(a)
        L.check(self.lib.gsn_add(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), self._stream()), "add")
        return out

    def tfr_unet(self, p, x):
        """gshift_deblur2.py:682-695."""
        n0, st = self.spec.n0, self.spec.unet_step
        c1, c2, c3 = n0, n0 + st, n0 + 2 * st
        enc1 = self.seq_cabs(p + ".encoder_level1", x, c1)
        enc2 = self.seq_cabs(p + ".encoder_level2", self.downsample(p + ".down12", enc1, c1, c2), c2)
        enc3 = self.seq_cabs(p + ".encoder_level3", self.downsample(p + ".down23", enc2, c2, c3), c3)
        dec3 = self.seq_cabs(p + ".decoder_level3", enc3, c3)
        del enc3                      # intermediates are dropped as soon as their last consumer is enqueued: a 1080p
        sk = self.cab(p + ".skip_attn2", enc2, c2)   # one_len=96 clip (BASELINE config K5) has 10 GB per full-res tensor
        del enc2
        y = self.skip_upsample(p + ".up32", dec3, c3, sk, c2)
        del dec3, sk
        dec2 = self.seq_cabs(p + ".decoder_level2", y, c2)
        sk = self.cab(p + ".skip_attn1", enc1, c1)
        del enc1
        y = self.skip_upsample(p + ".up21", dec2, c2, sk, c1)
        del dec2, sk
        return self.seq_cabs(p + ".decoder_level1", y, c1)

    # ------------------------------------------------------------------ fused shift + NAF block
    def _ln_params(self, p):
        ck = ("ln", p)
        if ck not in self.cache:
            self.cache[ck] = self._up(torch.cat((self.sd[p + ".norm.weight"].float(), self.sd[p + ".norm.bias"].float())))
        return self.cache[ck]

    def _fold_and_pass_b(self, p, x, z, partial, ntiles, fw, mode, next_p=None):
        T, H, W, Cc = x.shape
        weff = self._new(T, Cc * Cc)
        beff = self._new(T, Cc, dtype=torch.float32)
        L.check(self.lib.gsn_cab_fold(partial.data_ptr(), ntiles, 1.0 / (H * W), fw["du0"].data_ptr(), fw["du2"].data_ptr(),
                                      fw["du0"].shape[0], fw["w3"].data_ptr(), fw["beta"].data_ptr(),
                                      fw["bias3"].data_ptr() if fw["bias3"] is not None else None, Cc, T, weff.data_ptr(),
                                      beff.data_ptr(), self._stream()), "cab_fold")
        if self.tshard is not None and mode == L.MODE_CAB1:
            # T-sharded clip: the next CAB2 needs a halo frame behind this rank's frames -- leave room for it (no copy later)
            full = self._new(T + 1, H, W, Cc)
            out = full[:T]
            out._gsn_full = full
        else:
            out = self._new(T, H, W, Cc)
        b = L.CabPassB()
        b.T, b.H, b.W, b.C, b.mode, b.circular = T, H, W, Cc, mode, self._circ()
        b.x, b.z, b.weff, b.beff, b.out = x.data_ptr(), z.data_ptr(), weff.data_ptr(), beff.data_ptr(), out.data_ptr()
        self._a1_next = None
        if next_p is not None:      # the LayerNorm of the block that consumes `out` rides on this kernel's epilogue
            self._a1_next = self._new(T, Cc // 8, H, W, 8)
            b.ln_next, b.a1_next = self._ln_params(next_p).data_ptr(), self._a1_next.data_ptr()
        with self._timed(f"cab_pass_b[{H}x{W}]" if self.timeline_detail else "cab_pass_b", T * H * W):
            L.check(self.lib.gsn_cab_pass_b(C.byref(b), self._stream()), "cab_pass_b " + p)
        return out

    def gated_cab_generic(self, p, x, mode):
        """Width-generic CAB1/CAB2 (Ours+, C=80, grouped RepConv): the block split at tensor boundaries
        (csrc/generic_cab.cu): shift_ln -> 1x1 (a|b) -> dw_gate -> group_conv5 -> 1x1 (a|b) -> gate2 -> fold -> pass B."""
        T, H, W, Cc = x.shape
        shift = mode != L.MODE_CAB1
        k = self.boff
        sd = self.sd
        ck = ("gcabg", p)
        if ck not in self.cache:
            ln = torch.cat((sd[p + ".norm.weight"].float(), sd[p + ".norm.bias"].float())).contiguous()
            wc1 = sd[p + ".conv1.weight"].view(Cc // 2, 9).t().contiguous().half() if shift else None
            w1 = sd[p + ".body.0.weight"].float().flatten(1)                      # (2C, cin)
            kpad = (w1.shape[1] + 15) // 16 * 16
            w1z = torch.zeros(w1.shape[0], kpad, device=w1.device)
            w1z[:, :w1.shape[1]] = w1
            w1p = P.planar_chunks(w1z).contiguous()                              # [kpad/8][2C][8], zero-padded K
            wd = sd[p + ".body.1.conv_2.weight"].view(2 * Cc, 9).t().contiguous().half()
            rp = p + f".body.{3 + k}"
            wfrag = P.pack_group_conv5(sd[rp + ".conv_1.weight"], sd[rp + ".conv_2.weight"])
            w2p = P.planar_chunks(sd[p + f".body.{4 + k}.weight"].flatten(1)).contiguous()      # [C/8][2C][8] fp16
            wfold, wvec = P.pack_ln_pw_tc(w1, sd[p + ".norm.weight"], sd[p + ".norm.bias"])
            w2h = P.planar_chunks(0.5 * sd[p + f".body.{4 + k}.weight"].float().flatten(1)).contiguous()   # half scale: exact in fp16
            self.cache[ck] = self._up((ln, wc1, wd, wfrag, P.pack_cab_fold(sd, p, k), w2p, w1p, wfold, wvec, w2h))
        ln, wc1, wd, wfrag, fw, w2p, w1p, wfold, wvec, w2h = self.cache[ck]
        hw_pre = None
        if shift:      # gather folded into conv1's load stage (TMA-staged box), written once, read by the LayerNorm kernel
            hw_pre = self._new(T, H, W, Cc // 2)
            with self._timed("shift_conv1", T * H * W):
                L.check(self.lib.gsn_shift_conv1(x.data_ptr(), T, H, W, Cc, mode, self._circ(), wc1.data_ptr(),
                                                 hw_pre.data_ptr(), self._stream()), "shift_conv1 " + p)
        # LayerNorm + first 1x1 in one kernel (the 1.5C-wide LN input never goes to HBM), a|b halves written separately
        ga, gb = self._new(T, H, W, Cc), self._new(T, H, W, Cc)
        with self._timed("ln_pw", T * H * W):
            if self.ln_pw_tc and Cc == 80:      # TMA + tcgen05 streaming kernel, LayerNorm folded around the GEMM (csrc/ln_pw_tc.cu)
                L.check(self.lib.gsn_ln_pw_tc(x.data_ptr(), hw_pre.data_ptr() if hw_pre is not None else None, T, H, W, Cc, mode,
                                              self._circ(), wfold.data_ptr(), wvec.data_ptr(), ga.data_ptr(),
                                              gb.data_ptr(), self._stream()), "ln_pw_tc " + p)
            else:
                L.check(self.lib.gsn_ln_pw(x.data_ptr(), hw_pre.data_ptr() if hw_pre is not None else None, T, H, W, Cc, mode,
                                           self._circ(), ln.data_ptr(), w1p.data_ptr(), ga.data_ptr(), gb.data_ptr(),
                                           self._stream()), "ln_pw " + p)
        del hw_pre
        ntl = self.lib.gsn_cab_tiles_linear(H * W)
        g = self._new(T, H, W, Cc)
        pg = self._new(T, ntl, Cc, dtype=torch.float32) if self.spec.denoise else None
        L.check(self.lib.gsn_dw_gate(ga.data_ptr(), gb.data_ptr(), T, H, W, Cc, wd.data_ptr(), g.data_ptr(),
                                     pg.data_ptr() if pg is not None else None, self._stream()), "dw_gate")
        del ga, gb
        s1 = None
        if self.spec.denoise:        # mid CALayer2: group_conv5 scales its staged input tile (RepConv(s*g), gshift_denoise1.py:190-191)
            s1 = self._new(T, Cc, dtype=torch.float32)
            L.check(self.lib.gsn_ca_scale(pg.data_ptr(), ntl, 1.0 / (H * W), fw["mid_du0"].data_ptr(), fw["mid_du2"].data_ptr(),
                                          Cc, fw["mid_du0"].shape[0], Cc, T, s1.data_ptr(), self._stream()), "mid ca_scale")
        u = self._new(T, H, W, Cc)
        with self._timed("group_conv5", T * H * W):
            L.check(self.lib.gsn_group_conv5(g.data_ptr(), T, H, W, Cc, wfrag.data_ptr(),
                                             s1.data_ptr() if s1 is not None else None, u.data_ptr(), self._stream()), "group_conv5")
        del g
        # second 1x1 (C -> 2C) + SimpleGate2 + per-tile sums in one kernel (u is read once, a|b never touch HBM)
        z = self._new(T, H, W, Cc)
        partial = self._new(T, ntl, Cc, dtype=torch.float32)
        with self._timed("cab_pass_a2", T * H * W):
            if self.ln_pw_tc and Cc == 80:      # TMA + tcgen05 streaming kernel (csrc/pw_gate_tc.cu)
                L.check(self.lib.gsn_pw_gate_tc(u.data_ptr(), w2h.data_ptr(), z.data_ptr(), partial.data_ptr(), T, H, W, Cc,
                                                self._stream()), "pw_gate_tc")
            else:
                L.check(self.lib.gsn_cab_pass_a2(u.data_ptr(), w2p.data_ptr(), z.data_ptr(), partial.data_ptr(), T, H, W, Cc, 0,
                                                 self._stream()), "cab_pass_a2")
        del u
        return self._fold_and_pass_b(p, x, z, partial, ntl, fw, mode)

    def shift_cab(self, p, x, c, reverse):
        """Shift_CAB of Ours+ denoise (gshift_denoise1.py:157-186): clamped temporal roll, then the CAB body."""
        T, H, W, cp = x.shape
        own = slice(0, T)
        if self.tshard is not None:     # T-sharded clip: the neighbour rank's boundary frame joins the roll (host/tshard.py)
            x, own = self.tshard.with_neighbour_frame(x, reverse, self._new)
        n = x.shape[0]
        y = self._new(n, H, W, cp)
        L.check(self.lib.gsn_roll_copy(x.data_ptr(), y.data_ptr(), n, H, W, c, cp, 1 if reverse else 0, self._stream()), "roll_copy")
        return self.cab(p, y[own], c)

    def gated_cab(self, p, x, mode, debug_stage=0, a1_pre=None, next_p=None):
        """One CAB2 (mode fwd/rev: shift folded into the load) or CAB1 step: pass A -> fold -> pass B.
        a1_pre: the block's LayerNorm output if a producer already made it (pass B epilogue of the previous block);
        next_p: name of the CAB1 that consumes the result -- its LayerNorm is then emitted by this block's pass B
        (left in ``self._a1_next``)."""
        T, H, W, Cc = x.shape
        if Cc != 64:
            return self.gated_cab_generic(p, x, mode)
        shift = mode != L.MODE_CAB1
        ck = ("gcab", p)
        if ck not in self.cache:
            self.cache[ck] = self._up((P.pack_cab_pass_a(self.sd, p, Cc, shift, self.boff), P.pack_cab_fold(self.sd, p, self.boff)))
        blob, fw = self.cache[ck]
        # pass-A kernel choice (C-ABI: GsnCabPassA.debug_stage): 0 = library default (16x16-tile kernel unless GSN_PASS_A_STREAM=1),
        # PASS_A_FORCE_STREAM = the row-streaming kernel for this call (deblur nets; tests and A/B measurements)
        stage_sel = debug_stage or (L.PASS_A_FORCE_STREAM if (self.pass_a_stream and not self.spec.denoise) else 0)
        ntiles = self.lib.gsn_cab_pass_a_tiles(T, H, W, Cc, 1 if self.spec.denoise else 0, stage_sel)
        z = self._new(T, H, W, Cc)
        partial = self._new(T, ntiles, Cc, dtype=torch.float32)
        a = L.CabPassA()
        a.T, a.H, a.W, a.C, a.mode, a.circular = T, H, W, Cc, mode, self._circ()
        a.x, a.wblob, a.z, a.chan_partial = x.data_ptr(), blob.data_ptr(), z.data_ptr(), partial.data_ptr()
        a.mid_ca = 1 if self.spec.denoise else 0
        if shift and self.ln_fuse:
            ckw = ("wc1", p)
            if ckw not in self.cache:
                self.cache[ckw] = self._up(self.sd[p + ".conv1.weight"].reshape(Cc // 2, 9).t().contiguous().half())
            a1_pre = self._new(T, 12, H, W, 8)
            with self._timed("shift_conv1_ln", T * H * W):
                L.check(self.lib.gsn_shift_conv1_ln(x.data_ptr(), T, H, W, Cc, mode, a.circular, self.cache[ckw].data_ptr(),
                                                    self._ln_params(p).data_ptr(), a1_pre.data_ptr(), self._stream()), "shift_conv1_ln " + p)
        elif shift:
            ckw = ("wc1", p)
            if ckw not in self.cache:
                self.cache[ckw] = self._up(self.sd[p + ".conv1.weight"].reshape(Cc // 2, 9).t().contiguous().half())
            hw_pre = self._new(T, H, W, Cc // 2)
            with self._timed("shift_conv1", T * H * W):
                L.check(self.lib.gsn_shift_conv1(x.data_ptr(), T, H, W, Cc, mode, a.circular, self.cache[ckw].data_ptr(),
                                                 hw_pre.data_ptr(), self._stream()), "shift_conv1 " + p)
            a.hw_pre = hw_pre.data_ptr()
        if a1_pre is None:
            a1_pre = self._new(T, 12 if shift else 8, H, W, 8)
            with self._timed("ln_planar", T * H * W):
                L.check(self.lib.gsn_ln_planar(x.data_ptr(), a.hw_pre, T, H, W, Cc, mode, a.circular, self._ln_params(p).data_ptr(),
                                               a1_pre.data_ptr(), self._stream()), "ln_planar " + p)
        a.a1_pre = a1_pre.data_ptr()
        dbg = None
        if debug_stage:
            dbg = torch.zeros(T * ntiles * 12 * 512 * 8, dtype=torch.float16, device=self.dev)
            a.debug_stage, a.debug_out = debug_stage, dbg.data_ptr()
        elif stage_sel:
            a.debug_stage, a.debug_out = stage_sel, z.data_ptr()       # no dump is written; the field only selects the kernel
        with self._timed(("cab_pass_a_shift" if shift else "cab_pass_a") + (f"[{H}x{W}]" if self.timeline_detail else ""), T * H * W):
            L.check(self.lib.gsn_cab_pass_a(C.byref(a), self._stream()), "cab_pass_a " + p)
        if self.spec.denoise:
            # z holds u = RepConv(gate); finish the block: mid CALayer2 folded into W2, then 1x1 + sigmoid gate
            u = z
            w2eff = self._new(T, 2 * Cc * Cc)
            L.check(self.lib.gsn_cab_fold_mid(partial.data_ptr(), ntiles, 1.0 / (H * W), fw["mid_du0"].data_ptr(),
                                              fw["mid_du2"].data_ptr(), fw["mid_du0"].shape[0], fw["w2"].data_ptr(), Cc, T,
                                              w2eff.data_ptr(), self._stream()), "cab_fold_mid")
            ntiles = self.lib.gsn_cab_tiles_linear(H * W)
            partial = self._new(T, ntiles, Cc, dtype=torch.float32)
            z = self._new(T, H, W, Cc)
            with self._timed("cab_pass_a2", T * H * W):
                L.check(self.lib.gsn_cab_pass_a2(u.data_ptr(), w2eff.data_ptr(), z.data_ptr(), partial.data_ptr(), T, H, W, Cc, 1,
                                                 self._stream()), "cab_pass_a2")
        out = self._fold_and_pass_b(p, x, z, partial, ntiles, fw, mode, next_p=next_p)
        if debug_stage:
            return out, z, dbg
        return out

    def shift_block(self, p, x):
        """Encoder_shift_block.forward (gshift_deblur2.py:521-530): alternating fwd/rev (shift, CAB2, CAB1) pairs."""
        ts = self.tshard
        for i in range(self.spec.pairs):
            q = f"{p}.{_PAIRS[i]}"
            fuse = self.ln_fuse and x.shape[-1] == 64
            rev = bool(i & 1)
            if ts is not None:
                # this rank's n frames + the neighbour's boundary frame behind them (host/tshard.py): the kernels' roll over n+1 frames
                # (GSN_ROLL_HALO) finds frame 0's predecessor / frame n-1's successor at index n.  For the nets whose roll CLAMPS
                # at the ends of the clip, the rank that holds that end runs the step on its n frames with the clamped rule (it still
                # serves its other neighbour in the exchange).
                n = x.shape[0]
                use_halo = ts.needs_halo(rev, self.spec.circular)
                full = getattr(x, "_gsn_full", None)
                if full is None or full.shape[0] != n + 1:
                    full = self._new(n + 1, *x.shape[1:])
                    full[:n].copy_(x)
                ts.halo_into(full, n, rev, circular=self.spec.circular)
                # the kernels compute the n own frames; with ROLL_HALO their roll reaches frame n of the same buffer
                self._circ_override = L.ROLL_HALO if use_halo else L.ROLL_CLAMP
                try:
                    y = self.gated_cab(q + ".0", full[:n], L.MODE_CAB2_REV if rev else L.MODE_CAB2_FWD, next_p=(q + ".1") if fuse else None)
                finally:
                    self._circ_override = None
                x = self.gated_cab(q + ".1", y, L.MODE_CAB1, a1_pre=self._a1_next if fuse else None)
                continue
            x = self.gated_cab(q + ".0", x, L.MODE_CAB2_REV if rev else L.MODE_CAB2_FWD, next_p=(q + ".1") if fuse else None)
            x = self.gated_cab(q + ".1", x, L.MODE_CAB1, a1_pre=self._a1_next if fuse else None)
        return x

    # ------------------------------------------------------------------ stage 1 (Encoder2, Ours-s topology)
    def stage1(self, p, x):
        """gshift_deblur2.py:587-613 / gshift_denoise2.py:583-609."""
        sp = self.spec
        if sp.plus:
            return self.stage1_plus(p, x)
        n0, c = sp.n0, sp.c1
        if isinstance(x, list):       # [tensor]: the caller hands its only reference over, the input dies after the first CAB
            x = x.pop()
        x = self.cab(p + ".concat", x, n0)
        shortcut = x
        y = self.conv(p + ".down01.0", [x], [n0], c, stride=2, pad=0, prelu_key=p + ".down01.1.weight")
        for n in ("encoder_level1", "encoder_level1_1", "encoder_level1_2"):
            y = self.shift_block(f"{p}.{n}", y)
        enc11 = y
        y = self.downsample(p + ".down12", enc11, c, c)
        for n in ("encoder_level2", "encoder_level2_1", "encoder_level2_2",
                  "decoder_level2", "decoder_level2_1", "decoder_level2_2"):
            y = self.shift_block(f"{p}.{n}", y)
        y = self.skip_upsample(p + ".up21", y, c, self.cab(p + ".skip_attn1", enc11, c), c)
        for n in ("decoder_level1", "decoder_level1_1", "decoder_level1_2"):
            y = self.shift_block(f"{p}.{n}", y)
        sk = self.cab(p + ".skip_conv", shortcut, n0)
        if sp.denoise:
            up = self.conv(p + ".upsample0.upsample_conv", [y], [c], 4 * n0, pixel_shuffle=True)
            out = self.conv(p + ".conv_hr0", [up, sk], [n0, n0], n0)                      # gshift_denoise2.py:607
        else:
            up = self.conv(p + ".upsample0.upsample_conv", [y], [c], 4 * n0, pixel_shuffle=True, prelu_key=p + ".act.weight")
            out = self.conv(p + ".conv_hr0", [up], [n0], n0, residual=sk)                 # gshift_deblur2.py:611
        return self.cab(p + ".out_conv", out, n0)

    def stage1_plus(self, p, x):
        """Encoder2 of Ours+ (gshift_deblur1.py:614-643 ; gshift_denoise1.py:640-671): plain/Shift CABs going down three
        levels, shift blocks (C=80) on the way up."""
        sp = self.spec
        n0, c = sp.n0, sp.c1
        if isinstance(x, list):
            x = x.pop()
        x = self.cab(p + ".concat", x, n0)
        shortcut = x
        if sp.denoise:
            x = self.shift_cab(p + ".encoder_level0", x, n0, False)
            x = self.shift_cab(p + ".encoder_level0_1", x, n0, True)
        y = self.conv(p + ".down01.0", [x], [n0], c, stride=2, pad=0, prelu_key=p + ".down01.1.weight")
        if sp.denoise:
            enc11 = self.shift_cab(p + ".encoder_level1_1", self.shift_cab(p + ".encoder_level1", y, c, False), c, True)
        else:
            enc11 = self.cab(p + ".encoder_level1_1", self.cab(p + ".encoder_level1", y, c), c)
        y = self.downsample(p + ".down12", enc11, c, c)
        enc22 = self.cab(p + ".encoder_level2_1", self.cab(p + ".encoder_level2", y, c), c)
        y = self.downsample(p + ".down23", enc22, c, c)
        y = self.cab(p + ".encoder_level3_1", self.cab(p + ".encoder_level3", y, c), c)
        y = self.shift_block(p + ".decoder_level3", y)
        y = self.shift_block(p + ".decoder_level3_1", y)
        sk = self.cab(p + ".skip_attn2", enc22, c)
        del enc22
        y = self.skip_upsample(p + ".up32", y, c, sk, c)
        y = self.shift_block(p + ".decoder_level2", y)
        y = self.shift_block(p + ".decoder_level2_1", y)
        sk = self.cab(p + ".skip_attn1", enc11, c)
        del enc11
        y = self.skip_upsample(p + ".up21", y, c, sk, c)
        del sk
        for n in ("decoder_level1", "decoder_level1_1", "decoder_level1_2"):
            y = self.shift_block(f"{p}.{n}", y)
        sk = self.cab(p + ".skip_conv", shortcut, n0)
        up = self.conv(p + ".upsample0.upsample_conv", [y], [c], 4 * n0, pixel_shuffle=True)
        out = self.conv(p + ".conv_hr0", [up, sk], [n0, n0], n0)                          # gshift_deblur1.py:640
        return self.cab(p + ".out_conv", out, n0)

    # ------------------------------------------------------------------ whole net
    def forward(self, x, noise_map=None, past=2, future=2):
        """GShiftNet.forward (gshift_deblur2.py:748-756, gshift_denoise2.py:744-753)."""
        sp = self.spec
        if x.dim() != 5 or x.shape[0] != 1:
            raise ValueError(f"expected input (1,T,3,H,W), got {tuple(x.shape)}")
        if x.dtype not in (torch.float16, torch.float32):
            raise TypeError(f"unsupported input dtype {x.dtype}")
        xin = x[0].contiguous()
        T, cin, H, W = xin.shape
        nm = None
        if sp.denoise:
            if noise_map is None:
                raise ValueError("denoise arch needs noise_map (1,T,1,H,W)")
            nm = noise_map[0]          # (T,1,H,W), possibly an expand()ed view: read through its strides, never copied
            if nm.dtype != xin.dtype:
                nm = nm.to(xin.dtype)
            if tuple(nm.shape) != (T, 1, H, W):
                raise ValueError(f"noise_map must be (1,{T},1,{H},{W}), got {tuple(noise_map.shape)}")
        if cin != 3:
            raise ValueError(f"expected 3 image channels, got {cin}")
        # stage 1 halves the resolution twice (Ours-s) or three times (Ours+: down01, down12, down23) with stride-2 convs whose
        # outputs come back through x2 upsampling + skip: sizes that are not multiples of 4 / 8 make the reference raise a shape
        # mismatch in SkipUpSample (gshift_deblur2.py:352 `x + y`); same error behaviour here
        m = 8 if sp.plus else 4
        if H % m or W % m:
            raise ValueError(f"H and W must be multiples of {m} for {sp.name} (got {H}x{W}); the reference scripts crop/pad to that")
        ts = self.tshard
        if ts is not None:
            if T != ts.n_local:
                raise ValueError(f"T-sharded mode: this rank owns {ts.n_local} frames, got {T}")
            lo, hi = ts.local_output_range(past, future)
        else:
            if T - past - future <= 0:
                raise ValueError("clip too short for the requested past/future context")
            lo, hi = past, T - future
        dt = L.DTYPE_F16 if xin.dtype == torch.float16 else L.DTYPE_F32
        n0 = sp.n0
        n0p = P.pad8(n0)
        if "in" not in self.cache:
            self.cache["in"] = self._up(P.pack_conv_in(self.sd["feat_extract.0.weight"], self.sd["feat_extract.0.bias"], n0p))
            self.cache["out"] = self._up(P.pack_conv_out(self.sd["conv_last.weight"], n0p))
        wi, bi = self.cache["in"]
        f0 = self._new(T, H, W, n0p)
        if nm is None:
            L.check(self.lib.gsn_conv_in(xin.data_ptr(), dt, T, cin, H, W, wi.data_ptr(), bi.data_ptr(), n0p, f0.data_ptr(),
                                         self._stream()), "conv_in")
        else:
            st = nm.stride()
            L.check(self.lib.gsn_conv_in_nm(xin.data_ptr(), dt, T, cin, H, W, nm.data_ptr(), st[0], st[2], st[3], wi.data_ptr(),
                                            bi.data_ptr(), n0p, f0.data_ptr(), self._stream()), "conv_in_nm")
        x0 = self.cab("feat_extract.1", f0, n0)
        del f0
        # stage 0 (gshift_deblur2.py:731-737)
        f = x0
        for i in range(1, sp.n_orb + 1):
            f = self.tfr_unet(f"orb{i}", f)
        if not sp.denoise:
            f = self.add(f, x0)
        sam0 = f
        sam = self.conv("conv_trans", [f], [n0], n0)
        del f
        holder = [sam]
        if sp.denoise:
            sam0 = None               # the denoise nets feed conv_trans' output to stage 2 (gshift_denoise2.py:752)
        else:
            sam = None                # deblur: stage 1 owns conv_trans' output from here on
        dec = self.stage1("stage1", holder)
        del holder
        # stage 2 on the centre frames only (gshift_deblur2.py:738-746,755)
        s = slice(lo, hi)
        if hi <= lo:                  # T-sharded: all of this rank's frames are context frames of the clip
            return torch.empty(0, 3, H, W, dtype=xin.dtype, device=self.dev)
        third = sam[s] if sp.denoise else sam0[s]
        y = self.conv("rconcat", [x0[s], third, dec[s]], [n0, n0, n0], n0,
                      prelu_key="lrelu.weight" if sp.denoise else None)
        del x0, third, dec, sam, sam0
        r = y
        for i in range(1, sp.n_orb + 1):
            r = self.tfr_unet(f"rorb{i}", r)
        if not sp.denoise:
            r = self.add(r, y)
        To = hi - lo
        out = torch.empty(To, 3, H, W, dtype=xin.dtype, device=self.dev)
        wo = self.cache["out"]
        L.check(self.lib.gsn_conv_out(r.data_ptr(), n0p, self.sd["conv_last.weight"].shape[-1], wo.data_ptr(),
                                      xin[lo:].data_ptr(), cin, dt, To, H, W, out.data_ptr(), self._stream()), "conv_out")
        return out
