"""One-time weight re-packing: reference state_dict tensors (NCHW conv weights, fp32/fp16) -> the device
layouts the kernels in csrc/ read.  Runs once per (net, device); layouts are documented next to the
kernel that consumes them.
"""
from __future__ import annotations

import torch


def pad16(c: int) -> int:
    return (c + 15) // 16 * 16


def pad8(c: int) -> int:
    """Storage width of an activation tensor: channels padded to one 16-byte vector (8 fp16)."""
    return (c + 7) // 8 * 8


def pack_conv_mma(w: torch.Tensor, src_real, src_pad, cout_p: int) -> torch.Tensor:
    """Dense conv weight (cout, sum(src_real), k, k) -> fp16 [tap][cin_p/16][cout_p/8][32 lanes][4]
    (mma.m16n8k16 B fragments; csrc/conv_mma.cu).  Each source's real channels are placed at the start of
    its padded slot, padding rows/cols are zero."""
    w = w.float()
    cout, cin, k, _ = w.shape
    assert cin == sum(src_real)
    cin_p = pad16(sum(src_pad))      # the GEMM K dimension is padded to the mma k-step in shared memory only
    assert cout_p % 8 == 0 and cout <= cout_p and all(p % 8 == 0 for p in src_pad)
    wp = torch.zeros(cout_p, cin_p, k * k, device=w.device)
    ro = po = 0
    for r, p in zip(src_real, src_pad):
        wp[:cout, po:po + r] = w[:, ro:ro + r].reshape(cout, r, k * k)
        ro += r
        po += p
    nt, ks = cout_p // 8, cin_p // 16
    # n = nt*8 + g ; kk = kstep*16 + hi*8 + tig*2 + e ; lane = g*4 + tig ; regs (b0: hi=0, b1: hi=1) x e
    v = wp.view(nt, 8, ks, 2, 4, 2, k * k)            # (nt, g, kstep, hi, tig, e, tap)
    v = v.permute(6, 2, 0, 1, 4, 3, 5).contiguous()   # (tap, kstep, nt, g, tig, hi, e)
    return v.reshape(-1).half()


def pack_dense_frag(w: torch.Tensor, cp: int) -> torch.Tensor:
    """Dense 3x3 conv weight (cout, cin, 3, 3), cout, cin <= cp -> fp16 mma B fragments for csrc/cab_dense.cu:
    [tap = ky*3+kx][cp/8 n-tiles][words][32 lanes] of 32-bit words (two fp16: k, k+1), words = 2 per 16-channel k-step
    (k = 16s + 2*tig + e ; then + 8) followed by one word for an 8-channel tail (cp % 16 == 8); lane = g*4 + tig, n = nt*8 + g."""
    w = w.float()
    cout, cin, k, _ = w.shape
    assert k == 3 and cout <= cp and cin <= cp and cp % 8 == 0
    wp = torch.zeros(cp, cp, 9, device=w.device)
    wp[:cout, :cin] = w.reshape(cout, cin, 9)
    nt, k16, k8 = cp // 8, cp // 16, (cp % 16) // 8
    v = wp.view(nt, 8, cp // 8, 4, 2, 9)              # (nt, g, kchunk, tig, e, tap): k = kchunk*8 + 2*tig + e
    v = v.permute(5, 0, 2, 1, 3, 4).contiguous()      # (tap, nt, kchunk, g, tig, e): word index = kchunk, lane = g*4 + tig
    assert v.shape[2] == 2 * k16 + k8
    return v.reshape(-1).half()


def pack_conv_in(w, b, cout_p):
    """(cout, cin, 3, 3) -> fp32 [9][cin][cout_p], bias [cout_p] (csrc/io_convs.cu conv_in)."""
    cout, cin = w.shape[:2]
    wp = torch.zeros(9, cin, cout_p, device=w.device)
    wp[:, :, :cout] = w.float().permute(2, 3, 1, 0).reshape(9, cin, cout)
    bp = torch.zeros(cout_p, device=w.device)
    if b is not None:
        bp[:cout] = b.float()
    return wp.contiguous(), bp


def pack_conv_out(w, cp):
    """(3, cin, k, k) -> fp32 [k*k][cp][3] (csrc/io_convs.cu conv_out)."""
    cout, cin, k, _ = w.shape
    wp = torch.zeros(k * k, cp, 3, device=w.device)
    wp[:, :cin] = w.float().permute(2, 3, 1, 0).reshape(k * k, cin, cout)
    return wp.contiguous()


def pack_bias(b, cout_p):
    bp = torch.zeros(cout_p, device=b.device)
    bp[:b.numel()] = b.float()
    return bp


def planar_chunks(w2d: torch.Tensor) -> torch.Tensor:
    """(N, K) -> fp16 [K/8][N][8]: the k-chunk planar operand layout of csrc/shift_cab.cu."""
    n, k = w2d.shape
    return w2d.float().view(n, k // 8, 8).permute(1, 0, 2).contiguous().half()


def pack_cab_pass_a(sd, p, C, shift, body_off=0):
    """Pass-A weight blob of one CAB1/CAB2 (byte layout = PassACfg::OFF_* in csrc/shift_cab.cu).

    body indices (gshift_deblur2.py:193-204): 0 1x1 | 1 RepConv2 | 3 RepConv | 4 1x1 ; ``body_off`` = 1 for the
    denoise variants (extra CALayer2 at index 3)."""
    k = body_off
    parts = []
    ln = torch.cat((sd[p + ".norm.weight"].float(), sd[p + ".norm.bias"].float()))
    parts.append(ln.contiguous().view(torch.uint8))
    if shift:
        c1 = sd[p + ".conv1.weight"].float().view(C // 2, 9).t().contiguous().half()      # [9][C/2]
        parts.append(c1.view(torch.uint8).reshape(-1))
    w1 = sd[p + ".body.0.weight"].float().flatten(1)                                       # (2C, CIN)
    parts.append(planar_chunks(w1).view(torch.uint8).reshape(-1))
    da = sd[p + ".body.1.conv_2.weight"].float().view(2 * C, 9).t().contiguous()          # [9][2C]
    da[4] += 1.0                  # RepConv2 = dw3x3(x) + x: the identity rides on the centre tap (one HADD2 less per output)
    da = da.half()
    parts.append(da.view(torch.uint8).reshape(-1))
    w5 = sd[p + f".body.{3 + k}.conv_1.weight"].float().clone()                            # (C,1,5,5)
    w3 = sd[p + f".body.{3 + k}.conv_2.weight"].float()
    assert w5.shape[1] == 1, "grouped RepConv (Ours+) is not supported by this kernel yet"
    w5[:, :, 1:4, 1:4] += w3
    w5[:, :, 2, 2] += 1.0         # RepConv = dw5x5 + dw3x3 + x, all merged into one 5x5 kernel per channel
    db = w5.view(C, 25).t().contiguous().half()                                            # [25][C]
    parts.append(db.view(torch.uint8).reshape(-1))
    w2 = sd[p + f".body.{4 + k}.weight"].float().flatten(1)                                # (2C, C)
    parts.append(planar_chunks(w2).view(torch.uint8).reshape(-1))
    return torch.cat([x.reshape(-1) for x in parts]).contiguous()


def pack_cab_fold(sd, p, body_off=0):
    k = body_off
    ca = p + f".body.{6 + k}.conv_du"
    last = p + f".body.{7 + k}"
    d = dict(
        du0=sd[ca + ".0.weight"].float().flatten(1).contiguous(),
        du2=sd[ca + ".2.weight"].float().flatten(1).contiguous(),
        w3=sd[last + ".weight"].float().flatten(1).contiguous(),
        beta=sd[p + ".beta"].float().reshape(-1).contiguous(),
        bias3=sd[last + ".bias"].float().contiguous() if (last + ".bias") in sd else None,
    )
    if k:   # denoise: the mid CALayer2 (body.3) is folded into the second 1x1 (body.5) per frame
        d["mid_du0"] = sd[p + ".body.3.conv_du.0.weight"].float().flatten(1).contiguous()
        d["mid_du2"] = sd[p + ".body.3.conv_du.2.weight"].float().flatten(1).contiguous()
        d["w2"] = sd[p + f".body.{4 + k}.weight"].float().flatten(1).contiguous()
    return d


def pack_group_conv5(w5, w3):
    """Grouped RepConv taps (C, 8, 5, 5) + (C, 8, 3, 3) (gshift_deblur1.py:160-161) -> merged 5x5, fp16 mma B fragments
    [C/8 groups][13 k-steps = tap pairs][32 lanes][b0 (tap 2k), b1 (tap 2k+1)] (csrc/generic_cab.cu group_conv5)."""
    w = w5.float().clone()
    w[:, :, 1:4, 1:4] += w3.float()
    C = w.shape[0]
    wm = torch.zeros(C, 8, 26, device=w.device)
    wm[:, :, :25] = w.reshape(C, 8, 25)
    v = wm.view(C // 8, 8, 4, 2, 13, 2)            # (group, n, tig, e, kstep, tsel) ; cin = 2*tig + e ; tap = 2*kstep + tsel
    v = v.permute(0, 4, 1, 2, 5, 3).contiguous()   # (group, kstep, n, tig, tsel, e) -> lane = n*4 + tig, regs (b0, b1)
    return v.reshape(-1).half()


def pack_ln_pw_tc(w1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor):
    """LayerNorm folded around the first 1x1 of a C=80 gated block for csrc/ln_pw_tc.cu:
    W . LN(x) = rstd * (W' . x - mu * rowsum(W')) + W . beta with W' = W diag(gamma).
    w1 (2C, cin) fp32 -> (W' as fp16 k-chunk planar [16 planes][2C][8], K zero-padded to 128 ; fp32 [rowsum of the ROUNDED W' | W . beta])."""
    w1 = w1.float()
    n, cin = w1.shape
    wf = (w1 * gamma.float().view(1, -1)).half()
    wz = torch.zeros(n, 128, dtype=torch.float16, device=w1.device)
    wz[:, :cin] = wf
    wsum = wf.double().sum(1).float()
    bias = (w1.double() @ beta.double()).float()
    return planar_chunks(wz.float()).contiguous(), torch.cat((wsum, bias)).contiguous()


def pack_conv3x3_tc(w: torch.Tensor) -> torch.Tensor:
    """Dense 3x3 conv weight (cout <= 48, cin <= 48, 3, 3) -> fp16 [tap = ky*3+kx][6 k-chunks][48 rows (n)][8] for csrc/conv3x3_tc.cu:
    per (tap, k-step) the no-swizzle K-major UMMA B operand (N = 48 rows, two 8-channel chunks); padding rows / channels are zero."""
    w = w.float()
    cout, cin, k, _ = w.shape
    assert k == 3 and cout <= 48 and cin <= 48
    wz = torch.zeros(48, 48, 9, device=w.device)
    wz[:cout, :cin] = w.reshape(cout, cin, 9)
    return wz.view(48, 6, 8, 9).permute(3, 1, 0, 2).contiguous().half()
