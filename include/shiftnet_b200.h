/*
 * shiftnet_b200 -- C-ABI of the B200-native (sm_100a) Shift-Net forward hot path.
 *
 * The reference (dasongli1/Shift-Net) has no FFI/operator layer of its own: its boundary is the
 * Python class ``basicsr.models.archs.gshift_*.GShiftNet``.  The entry
 * points below are what a Python/ctypes (or cgo/JNI) binding of that class's forward would bind;
 * each cites the reference code it replaces (paths relative to
 * /root/reference/basicsr/models/archs/, "d2" = gshift_deblur2.py).
 *
 * Conventions
 *   - every pointer is a raw DEVICE pointer (cudaMalloc'd / torch storage); nothing is owned or
 *     freed by the library; all calls are stream-ordered and stateless (thread-safe per stream);
 *   - activations are NHWC fp16, shape (T, H, W, Cp) with Cp = channel count padded to a multiple
 *     of 8 (one 16-byte vector; padding channels are zero and stay zero);
 *   - return value: 0 on success, negative GSN_E_* on failure; gsn_last_error() gives the message
 *     of the last failure on the calling thread;
 *   - ``stream`` is a cudaStream_t passed as void*.
 */
#ifndef SHIFTNET_B200_H
#define SHIFTNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSN_OK 0
#define GSN_E_BADARG (-1)
#define GSN_E_CUDA (-2)
#define GSN_E_UNSUPPORTED (-3)

#define GSN_DTYPE_F16 0
#define GSN_DTYPE_F32 1

int gsn_version(void);
const char *gsn_last_error(void);
/* Number of kernels this library has launched since load (bench.py's "gpu_launches" evidence). */
unsigned long long gsn_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Generic dense convolution on tensor cores (implicit GEMM), replaces every nn.Conv2d with
 * groups == 1 on the path: CAB bodies (d2:143-158), conv_trans (d2:713), DownSample (d2:333-343),
 * down01 (d2:551), SkipUpSample's 1x1 (d2:344-353, applied before the bilinear upsample -- both
 * are linear), PixelShufflePack (d2:259-281, shuffle + PReLU fused into the store), rconcat
 * (d2:727, the torch.cat of d2:739 is folded into the load of up to 3 sources), conv_hr0 (d2:572).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int T, Hin, Win, Hout, Wout;
  int n_src;            /* 1..3 sources, concatenated along channels */
  const void *src[3];   /* NHWC fp16 (T,Hin,Win,src_c[i]) */
  int src_c[3];         /* stored channel count of each source (multiple of 8) */
  int cin_p;            /* GEMM K: sum of src_c rounded up to 16 (the tail is zero-filled in shared memory only) */
  int cout_p;           /* stored output channels: 16, 24, 32, 40, 48, 64, 80 or 96 */
  int ks, stride, pad;  /* kernel size 1..3, stride 1..2, zero padding */
  const void *wpack;    /* fp16 weights in mma.m16n8k16 B-fragment order: [tap][cin_p/16][cout_p/8][32 lanes][4] */
  const float *bias;    /* cout_p floats or NULL */
  int has_prelu;        /* apply PReLU(slope) after bias */
  float prelu_slope;
  const void *residual; /* NULL or NHWC fp16 (T,Hout,Wout,cout_p) added last */
  int pixel_shuffle;    /* 1: dst is (T,2Hout,2Wout,cout_p/4), F.pixel_shuffle(.,2) fused into the store */
  float *chan_partial;  /* NULL or [T][tiles][cout_p] fp32: per-tile channel sums of the output (for CALayer) */
  void *dst;            /* NHWC fp16 */
  int dst_c;            /* pixel_shuffle only: channel pitch of dst (0 = cout_p/4); channels >= cout_p/4 are not written */
} GsnConvDesc;

/* tiles per frame the conv kernel uses for chan_partial (16x16 output tiles) */
int gsn_conv_tiles(int Hout, int Wout);
int gsn_conv_mma(const GsnConvDesc *d, void *stream);

/* The same conv for the wide CAB bodies of Ours+ (3x3, stride 1, pad 1, one source with 40 or 48 stored channels, 40 or 48 stored
 * output channels, optional bias / PReLU / chan_partial; no residual, no shuffle) as an implicit GEMM on tcgen05: the 9 taps are
 * descriptor offsets into one TMA-landed k-chunk planar tile.  wpack: fp16 [9][6][48][8] (host/packing.py pack_conv3x3_tc), bias: 48
 * floats or NULL, chan_partial: [T][gsn_conv3x3_tc_tiles(H,W)][cout_p]. */
int gsn_conv3x3_tc_tiles(int H, int W);
int gsn_conv3x3_tc(const GsnConvDesc *d, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused body of the dense channel-attention block: r = conv3x3(PReLU(conv3x3(x))) in ONE kernel (the intermediate
 * stays in shared memory) plus the per-tile channel sums of r for the CALayer pooling.  Replaces the two nn.Conv2d
 * and the nn.PReLU of CAB.body (d2:143-150,154); gsn_ca_scale + gsn_scale_residual finish the block (d2:155-158).
 * x, r: (T,H,W,cp) NHWC fp16, cp = 16 or 24 (wider CABs keep using two gsn_conv_mma launches); r must not alias x.
 * w1pack / w2pack: fp16 weights in mma B-fragment order [tap = ky*3+kx][cp/8][2*(cp/16) + (cp%16)/8 words][32 lanes]
 * (host/packing.py pack_dense_frag); bias1 / bias2: cp floats or NULL.
 * chan_partial: [T][gsn_cab_dense_tiles(cp,H,W)][cp] fp32 or NULL.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int T, H, W, cp;
  const void *x;
  const void *w1pack, *w2pack;
  const float *bias1, *bias2;
  int has_prelu;
  float prelu_slope;
  void *r;
  float *chan_partial;
} GsnCabDense;

int gsn_cab_dense_tiles(int cp, int H, int W);
int gsn_cab_dense(const GsnCabDense *d, void *stream);

/* First conv: NCHW user clip -> NHWC features. Replaces feat_extract[0] (d2:710), including the
 * x[0] un-batching of d2:750 and (denoise) the torch.cat((x, noise_map)) of gshift_denoise2.py:749.
 * x: (T,cin,H,W) fp16 or fp32; w: fp32 [9][cin][cout_p]; bias: fp32 [cout_p]; dst NHWC fp16. */
int gsn_conv_in(const void *x, int x_dtype, int T, int cin, int H, int W, const float *w, const float *bias,
                int cout_p, void *dst, void *stream);

/* The same conv for the denoise nets with the torch.cat((x, noise_map), 1) of gshift_denoise2.py:749 folded into the load:
 * x holds the `cin` image channels, conv-input channel `cin` is read from noise_map (same dtype as x) at
 * noise_map[t*nm_stride_t + y*nm_stride_y + x*nm_stride_x] (element strides; an expand()ed map has zero strides).
 * w: fp32 [9][cin+1][cout_p]. */
int gsn_conv_in_nm(const void *x, int x_dtype, int T, int cin, int H, int W, const void *noise_map, long long nm_stride_t,
                   long long nm_stride_y, long long nm_stride_x, const float *w, const float *bias, int cout_p, void *dst,
                   void *stream);

/* Last conv: NHWC features -> NCHW frames + input residual. Replaces conv_last (d2:712,745) and the
 * "+ shortcut[num_fb:frames-num_ff]" of d2:756.  src (T,H,W,cp) fp16; w fp32 [ks*ks][cp][3];
 * resid/dst: (T,3,H,W) in x_dtype, resid has cres channels per frame (3, or 4 for the cat'ed denoise input). */
int gsn_conv_out(const void *src, int cp, int ks, const float *w, const void *resid, int cres, int x_dtype, int T, int H,
                 int W, void *dst, void *stream);

/* ---------------------------------------------------------------------------------------------
 * I/O side of the evaluation loop (inference/test_deblur_small.py:134-143,191-200).
 * gsn_u8_to_clip: uint8 HWC frames (T,H,W,3), as read from disk and copied to the device unchanged, -> the (T,3,H,W) clip in
 *   [0,1] with numpy2tensor's arithmetic: float32(u8) * float32(1/255), stored as fp16 or fp32.
 * gsn_psnr_sse: partial[t][b] (b < gsn_psnr_sse_blocks()) = float64 partial sums of (clamp(out,0,1)*255 - gt)^2 over frame t,
 *   out (T,3,H,W) fp16/fp32 as returned by the net, gt (T,H,W,3) uint8; the host adds the partials of a frame in index
 *   order and finishes PSNR = 10 log10(255^2 / (sum / (3 H W))) -- skimage's peak_signal_noise_ratio(data_range=255).
 * ------------------------------------------------------------------------------------------- */
int gsn_u8_to_clip(const void *frames_u8, int T, int H, int W, int dtype, void *clip, void *stream);
int gsn_psnr_sse_blocks(void);
int gsn_psnr_sse(const void *out, int dtype, const void *gt_u8, int T, int H, int W, double *partial, void *stream);

/* SSIM of the same loop (inference/test_deblur_small.py:25-49) on the device: scipy's 3-D gaussian_filter(sigma=1.5) semantics
 * (13 taps along the channel, H and W axes, 'reflect' boundary, float64 accumulation and float32 rounding per axis pass).
 * partial[t][b] (b < gsn_ssim_blocks()) = float64 partial sums of the SSIM map of frame t; SSIM = sum / (3 H W).
 * workspace: gsn_ssim_workspace_bytes(T,H,W) bytes of device memory. */
int gsn_ssim_blocks(void);
long long gsn_ssim_workspace_bytes(int T, int H, int W);
int gsn_ssim(const void *out, int dtype, const void *gt_u8, int T, int H, int W, void *workspace, double *partial, void *stream);

/* CALayer squeeze-excite MLP (d2:54-71): s[t][c] = sigmoid(W2 relu(W1 mean_hw)), from per-tile sums.
 * partial [T][ntiles][cp]; w1 fp32 [cr][c]; w2 fp32 [c][cr]; s fp32 [T][cp]. */
int gsn_ca_scale(const float *partial, int ntiles, float inv_hw, const float *w1, const float *w2, int c, int cr, int cp,
                 int T, float *s, void *stream);

/* out = x + res * s[t][c] (+ extra)   -- the "res = CA(res); res += x" of CAB.forward (d2:154-158). */
int gsn_scale_residual(const void *x, const void *res, const float *s, const void *extra, void *out, int T, long long hw,
                       int cp, void *stream);

/* dst = bilinear_x2(src) + skip   (nn.Upsample(scale_factor=2, bilinear, align_corners=False), d2:347,350-353). */
int gsn_upsample2x_add(const void *src, const void *skip, void *dst, int T, int h, int w, int cp, void *stream);

/* out = a + b (fp16, n elements, n % 8 == 0) -- the stage-level shortcuts (d2:736,744). */
int gsn_add(const void *a, const void *b, void *out, long long n, void *stream);

/* ---------------------------------------------------------------------------------------------
 * The fused grouped spatial-temporal shift + NAF block (Encoder_shift_block, d2:443-530).
 * One (shift, CAB2) or CAB1 step is two kernels around the global average pool of CALayer2:
 *   pass A: [temporal roll + spatial shift gather -> dw3x3 (conv1) ->] LayerNorm -> 1x1 -> dw3x3+id ->
 *           gate -> dw5x5+dw3x3+id -> 1x1 -> sigmoid gate -> z (C ch) + per-tile channel sums
 *   fold  : CA MLP on the pooled z, folded with beta into a per-frame effective last-1x1 weight
 *   pass B: out = shortcut + Weff_t . z      (shortcut = the ROLLED stream for CAB2, d2:253,257)
 * ------------------------------------------------------------------------------------------- */
#define GSN_MODE_CAB1 0
#define GSN_MODE_CAB2_FWD 1
#define GSN_MODE_CAB2_REV 2
/* Values of every `circular` argument / field below: how the half-channel temporal roll treats the ends of the frame range.
 *   GSN_ROLL_CLAMP  the first (forward) / last (reverse) frame stays un-swapped (gshift_deblur1.py:513,517)
 *   GSN_ROLL_WRAP   the roll wraps around the T frames (gshift_deblur2.py:504-505)
 *   GSN_ROLL_HALO   T-sharded clip (no reference counterpart): x holds T + 1 frames, frame T being the neighbour rank's boundary
 *                   frame; the roll wraps over T + 1 frames while only the T own frames are computed and written */
#define GSN_ROLL_CLAMP 0
#define GSN_ROLL_WRAP 1
#define GSN_ROLL_HALO 2
/* GsnCabPassA.debug_stage value that routes the call to the row-streaming pass-A kernel whatever GSN_PASS_A_STREAM says */
#define GSN_PASS_A_FORCE_STREAM 100

typedef struct {
  int T, H, W, C;       /* C = 64 (Ours-s); activations (T,H,W,C) NHWC fp16 */
  int mode;             /* GSN_MODE_* */
  int circular;         /* GSN_ROLL_*: temporal roll wraps (d2:504-505), clamps (gshift_deblur1.py:513,517) or reads a halo frame */
  const void *x;        /* input of the step (un-rolled previous output) */
  const void *wblob;    /* packed pass-A weights, see host/packing.py (pack_cab_pass_a) */
  void *z;              /* out: gated features (T,H,W,C) fp16 */
  float *chan_partial;  /* out: [T][gsn_cab_tiles][C] fp32 per-tile channel sums of z */
  int debug_stage;      /* 0 = normal; >0 dumps an intermediate smem stage to debug_out (tests only) */
  void *debug_out;
  int mid_ca;           /* denoise variants (gshift_denoise2.py:194,227: a CALayer2 sits between the first gate and the
                           RepConv).  A per-channel scale commutes with the depthwise RepConv, so with mid_ca=1 pass A
                           stops after the RepConv: ``z`` receives u = RepConv(gate) (C ch) and ``chan_partial`` the
                           per-tile sums of the gated tensor; gsn_cab_fold_mid + gsn_cab_pass_a2 finish the block. */
  const void *hw_pre;   /* unused by pass A (kept for ABI stability): (T,H,W,C/2) fp16 from gsn_shift_conv1 is consumed by
                           gsn_ln_planar, which produces a1_pre */
  const void *a1_pre;   /* REQUIRED: the LayerNorm (d2:209,250) output in the k-chunk planar layout [T][CIN/8][H][W][8] fp16
                           (CIN = C for CAB1, 3C/2 for CAB2), produced by gsn_shift_conv1_ln (CAB2), by the epilogue of the
                           preceding gsn_cab_pass_b (CAB1, GsnCabPassB.a1_next) or by gsn_ln_planar: pass A lands each tile's
                           halo'd region with TMA tile loads directly in the tensor-core operand layout; x is not read. */
} GsnCabPassA;

int gsn_cab_tiles(int mode, int H, int W);
/* Number of per-frame partial-sum slots gsn_cab_pass_a writes for this problem: chan_partial is [T][slots][C]; slots depend on
 * which pass-A kernel serves the call (row-streaming pieces for C = 64 without mid_ca, 16x16 tiles otherwise). */
int gsn_cab_pass_a_tiles(int T, int H, int W, int C, int mid_ca, int debug_stage);
int gsn_cab_pass_a(const GsnCabPassA *d, void *stream);

/* LayerNorm2d of CAB1 / CAB2 (d2:19-28,44-53,209,250) as its own HBM-bound kernel, for GsnCabPassA.a1_pre:
 * a1[t][chunk][y][x][8] = LN([rolled stream | hw_pre])  (CAB1: LN(x)); x (T,H,W,C) fp16, hw_pre (T,H,W,C/2) fp16 from
 * gsn_shift_conv1 (CAB2 modes), ln = fp32 gamma[CIN] then beta[CIN]; C = 64. */
int gsn_ln_planar(const void *x, const void *hw_pre, int T, int H, int W, int C, int mode, int circular, const float *ln,
                  void *a1, void *stream);

/* Grouped spatial-temporal shift folded into the load stage of conv1 (dw3x3): out = conv1(spatial_shift2(hw)) with hw the
 * neighbour frame's half of the channels (d2:465-519, 226,254); out (T,H,W,C/2) fp16.  wc1: fp16 [9][C/2]. */
int gsn_shift_conv1(const void *x, int T, int H, int W, int C, int mode, int circular, const void *wc1, void *out, void *stream);
/* The same gather + conv1, followed in the same kernel by CAB2's LayerNorm over [rolled stream | conv1 output] (d2:250-254):
 * a1 = k-chunk planar [T][3C/16][H][W][8] fp16 for GsnCabPassA.a1_pre (the conv1 output never goes to HBM).  C = 64.
 * ln: fp32 gamma[3C/2], beta[3C/2]. */
int gsn_shift_conv1_ln(const void *x, int T, int H, int W, int C, int mode, int circular, const void *wc1, const float *ln,
                       void *a1, void *stream);

/* mid fold (denoise): w2eff[t] = W2 diag(s1_t) as fp16 [T][C/8][2C][8], s1 from the mid CALayer2 on mean(gated).
 * w_du0 [cr][C], w_du2 [C][cr], w2 [2C][C] (fp32). */
int gsn_cab_fold_mid(const float *partial, int ntiles, float inv_hw, const float *w_du0, const float *w_du2, int cr,
                     const float *w2, int C, int T, void *w2eff, void *stream);

/* pass A2 (denoise): z = a * sigmoid(b), [a|b] = w2eff_t . u  (1x1 C->2C + SimpleGate2, gshift_denoise2.py:196-197),
 * plus per-tile channel sums of z: chan_partial [T][gsn_cab_tiles_linear(H*W)][C]. */
int gsn_cab_tiles_linear(long long hw);
int gsn_cab_pass_a2(const void *u, const void *w2eff, void *z, float *chan_partial, int T, int H, int W, int C,
                    int per_frame_weights /* 1: w2eff is [T][..] from gsn_cab_fold_mid; 0: one [C/8][2C][8] weight */, void *stream);

/* The same stage for C = 80 on TMA + tcgen05 (one weight for all frames): w2half = 0.5 * W2 as fp16 k-chunk planar [C/8][2C][8]
 * (a * sigmoid(b) = (a/2) tanh(b/2) + a/2); chan_partial [T][gsn_cab_tiles_linear(H*W)][C]. */
int gsn_pw_gate_tc(const void *u, const void *w2half, void *z, float *chan_partial, int T, int H, int W, int C, void *stream);

/* fold: weff[t] = diag(beta) W3 diag(s_t) as fp16 [T][C/8][C][8] (k-chunk planar), s_t from CALayer2
 * (d2:72-89,238-239,257).  w_du0 [cr][C], w_du2 [C][cr], w3 [C][C], beta [C], bias3 [C] or NULL (fp32).
 * beff [T][C] fp32 = beta*bias3 (zeros if bias3 == NULL). */
int gsn_cab_fold(const float *partial, int ntiles, float inv_hw, const float *w_du0, const float *w_du2, int cr,
                 const float *w3, const float *beta, const float *bias3, int C, int T, void *weff, float *beff,
                 void *stream);

typedef struct {
  int T, H, W, C;
  int mode, circular;   /* shortcut = rolled x for CAB2 modes, x itself for CAB1 */
  const void *x;        /* the same tensor pass A read */
  const void *z;
  const void *weff;     /* from gsn_cab_fold */
  const float *beff;
  void *out;            /* (T,H,W,C) fp16 */
  const float *ln_next; /* optional (C = 64): fp32 gamma[C], beta[C] of the LayerNorm that consumes `out` next (CAB1.norm) */
  void *a1_next;        /* optional: that LayerNorm applied to `out`, k-chunk planar [T][C/8][H][W][8] fp16 (see
                           GsnCabPassA.a1_pre) -- saves the separate gsn_ln_planar pass over `out` */
} GsnCabPassB;

int gsn_cab_pass_b(const GsnCabPassB *d, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Width-generic pieces (C = 64 or 80) used by the Ours+ nets (gshift_deblur1.py / gshift_denoise1.py: C = 80,
 * RepConv with groups of 8, gshift_deblur1.py:157-165), whose block is split at tensor boundaries:
 *   gsn_shift_ln -> gsn_conv_mma (1x1, two halves) -> gsn_dw_gate -> gsn_group_conv5 -> gsn_conv_mma -> gsn_gate2
 *   -> gsn_cab_fold -> gsn_cab_pass_b.
 * ------------------------------------------------------------------------------------------- */
/* [roll + shift gather + conv1] + LayerNorm -> out (T,H,W,cinp) fp16, cinp = pad16(C or 3C/2), padding channels 0.
 * wc1: fp16 [9][C/2] (CAB2 modes); ln: fp32 gamma[cin] then beta[cin]. */
int gsn_shift_ln(const void *x, int T, int H, int W, int C, int mode, int circular, const void *wc1, const float *ln,
                 void *out, int cinp, const void *hw_pre /* optional (T,H,W,C/2) from gsn_shift_conv1 */, void *stream);
/* LayerNorm + first 1x1 fused: [rolled stream | hw_pre] -> LN -> W1 -> (ga | gb), each (T,H,W,C) fp16.
 * w1p: fp16 [K/8 padded to even][2C][8] (k-chunk planar, zero-padded K); ln: fp32 gamma[cin], beta[cin]. */
int gsn_ln_pw(const void *x, const void *hw_pre, int T, int H, int W, int C, int mode, int circular, const float *ln,
              const void *w1p, void *ga, void *gb, void *stream);
/* The same stage on TMA + tcgen05 (C = 80): the LayerNorm is folded around the GEMM, W . LN(x) = rstd (W' x - mu rowsum(W')) + W beta.
 * wfold: W' = W diag(gamma) as fp16 k-chunk planar [16][2C][8] (K zero-padded to 128); wvec: fp32 rowsum(W')[2C] then (W beta)[2C]
 * (host/packing.py pack_ln_pw_tc). */
int gsn_ln_pw_tc(const void *x, const void *hw_pre, int T, int H, int W, int C, int mode, int circular, const void *wfold,
                 const float *wvec, void *ga, void *gb, void *stream);
/* g = (dw3x3(a)+a) * (dw3x3(b)+b) (RepConv2 + SimpleGate); wd fp16 [9][2C]; partial (optional) [T][tiles_linear][C]. */
int gsn_dw_gate(const void *a, const void *b, int T, int H, int W, int C, const void *wd, void *out, float *partial,
                void *stream);
/* z = a * sigmoid(b) (SimpleGate2) + per-tile channel sums [T][tiles_linear][C]. */
int gsn_gate2(const void *a, const void *b, int T, int H, int W, int C, void *out, float *partial, void *stream);
/* u = (conv5x5 + conv3x3 + id)(s * g), groups of 8 channels (RepConv, gshift_deblur1.py:157-165); s = optional per-frame
 * channel scale [T][C] (fp32) of the denoise mid CALayer2 (gshift_denoise1.py:190-191,224-225), applied to the INPUT.
 * wfrag: merged 5x5 taps in mma B-fragment order [C/8][13][32 lanes][4] fp16 (host/packing.py pack_group_conv5). */
int gsn_group_conv5(const void *g, int T, int H, int W, int C, const void *wfrag, const float *scale, void *out, void *stream);
/* y = clamped half-channel temporal roll of x (Shift_CAB.channel_shift, gshift_denoise1.py:167-179); C real of cp channels. */
int gsn_roll_copy(const void *x, void *y, int T, int H, int W, int C, int cp, int reverse, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SHIFTNET_B200_H */
