// Fused grouped spatial-temporal shift + NAF block (reference: Encoder_shift_block / CAB2 / CAB1,
// gshift_deblur2.py:186-258,443-530).  Three kernels per CAB around the global average pool of CALayer2:
//
//   cab_pass_a : [temporal roll + per-channel spatial shift gather -> dw3x3 (conv1)] -> LayerNorm -> 1x1 (C'->2C)
//                -> dw3x3 + id -> gate -> dw5x5 + dw3x3 + id -> 1x1 (C->2C) -> a*sigmoid(b) -> z + per-tile sums
//   cab_fold   : CALayer2 MLP on mean(z), folded with beta into a per-frame last-1x1 weight
//   cab_pass_b : out = shortcut + Weff_t z         (shortcut = rolled stream for CAB2)
//
// Shared-memory operands use one "k-chunk planar" layout everywhere: buf[chunk][pixel] of 16-byte vectors
// (8 fp16 channels of one pixel), plane pitch (npix+1)*16 B.  It is bank-conflict free for (a) ldmatrix rows
// (8 consecutive pixels of one chunk = 128 contiguous bytes), (b) depthwise stencils (lanes = consecutive
// pixels) and (c) per-pixel all-chunk accesses, and it is exactly the no-swizzle K-major UMMA layout
// (SBO = 128 B, LBO = plane pitch) so the GEMM stages can move to tcgen05 without touching the other stages.
//
// Zero-padding semantics: every conv of the reference zero-pads its own input, so halo pixels
// outside the image are forced to 0 after the LayerNorm (A1 rows) and after the first gate (GATED).
#include <cstdlib>

#include "common.cuh"
#include "shift_common.cuh"

namespace gsn {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------------
// fold: CALayer2 MLP + beta folded into the per-frame last 1x1
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) cab_fold_kernel(const float *__restrict__ partial, int ntiles, float inv_hw,
                                                       const float *__restrict__ w_du0, const float *__restrict__ w_du2,
                                                       int cr, const float *__restrict__ w3, const float *__restrict__ beta,
                                                       const float *__restrict__ bias3, int C, __half *__restrict__ weff,
                                                       float *__restrict__ beff) {
  __shared__ float mean[128], hid[128], sc[128], part[1024];
  const int t = blockIdx.x, tid = threadIdx.x;
  {  // deterministic two-level reduction of the per-tile channel sums (1024 threads: the ~900 dependent adds per channel of a
     // 720p level-1 frame become ~60 per thread -- this latency-bound kernel runs 96 times per forward)
    const int nparts = 1024 / C, ch = tid % C, pi = tid / C;
    float a = 0.f;
    if (pi < nparts) {
      const float *p = partial + (size_t)t * ntiles * C + ch;
      for (int i = pi; i < ntiles; i += nparts) a += p[(size_t)i * C];
    }
    part[tid] = a;
    __syncthreads();
    if (tid < C) {
      float m = 0.f;
      for (int q = 0; q < nparts; ++q) m += part[q * C + tid];
      mean[tid] = m * inv_hw;
    }
  }
  __syncthreads();
  if (tid < cr) {
    float a = 0.f;
    for (int i = 0; i < C; ++i) a += w_du0[tid * C + i] * mean[i];
    hid[tid] = a > 0.f ? a : 0.f;
  }
  __syncthreads();
  if (tid < C) {
    float a = 0.f;
    for (int i = 0; i < cr; ++i) a += w_du2[tid * cr + i] * hid[i];
    sc[tid] = 1.f / (1.f + expf(-a));
    beff[(size_t)t * C + tid] = bias3 ? beta[tid] * bias3[tid] : 0.f;
  }
  __syncthreads();
  __half *wt = weff + (size_t)t * C * C;
  for (int i = tid; i < C * C; i += 1024) {
    const int co = i / C, ci = i - co * C;
    wt[((ci >> 3) * C + co) * 8 + (ci & 7)] = __float2half_rn(beta[co] * w3[co * C + ci] * sc[ci]);
  }
}

// ------------------------------------------------------------------------------------------------
// pass B: out = shortcut + Weff_t z (+ beff)
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256, (C <= 64 ? 4 : 3)) cab_pass_b_kernel(const GsnCabPassB d) {
  constexpr int MP = 128, KC = C / 8, PZ = (MP + 1) * 16, NT = C / 8;
  extern __shared__ __align__(128) unsigned char smem_b[];
  unsigned char *sz = smem_b, *ss = smem_b + KC * PZ, *sw = smem_b + 2 * KC * PZ;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = blockIdx.y;
  const long long hw = (long long)d.H * d.W, p0 = (long long)blockIdx.x * MP;
  const size_t frame = (size_t)hw * C;
  const RollSrc rs = roll_source(d.mode, d.circular, t, d.T, C);
  const __half *xg = reinterpret_cast<const __half *>(d.x);
  const __half *zg = reinterpret_cast<const __half *>(d.z) + (size_t)t * frame;
  {
    const unsigned char *wg = reinterpret_cast<const unsigned char *>(d.weff) + (size_t)t * C * C * 2;
    for (int i = tid; i < KC * C; i += 256) cp_async16(sw + i * 16, wg + i * 16, true);
    for (int i = tid; i < MP * KC; i += 256) {
      const int ch = i % KC, p = i / KC;
      const bool valid = p0 + p < hw;
      const size_t pix = valid ? (size_t)(p0 + p) * C : 0;
      cp_async16(sz + ch * PZ + p * 16, zg + pix + ch * 8, valid);
      const int hc = KC / 2;
      const __half *sp = (ch < hc) ? xg + rs.f_lo * frame + pix + rs.c_lo + ch * 8
                                   : xg + rs.f_hi * frame + pix + rs.c_hi + (ch - hc) * 8;
      cp_async16(ss + ch * PZ + p * 16, sp, valid);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  }
  const int g = lane >> 2, tig = lane & 3;
  float acc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
  const uint32_t z_s = smem_u32(sz), w_s = smem_u32(sw);
#pragma unroll
  for (int k = 0; k < KC / 2; ++k) {
    uint32_t a[4];
    ldmatrix_x4(a[0], a[1], a[2], a[3], z_s + (2 * k + (lane >> 4)) * PZ + (warp * 16 + (lane & 15)) * 16);
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      ldmatrix_x4(b[0], b[1], b[2], b[3],
                  w_s + ((2 * k + ((lane >> 3) & 1)) * C + np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * 16);
      mma16816(acc[2 * np], a, b[0], b[1]);
      mma16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
  const float *be = d.beff + (size_t)t * C;
  const bool ln_next = C == 64 && d.a1_next != nullptr;
  float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};   // LayerNorm statistics of the two pixel rows this thread holds a quarter of
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const float b0 = __ldg(be + n * 8 + tig * 2), b1 = __ldg(be + n * 8 + tig * 2 + 1);
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      uint32_t *sp = reinterpret_cast<uint32_t *>(ss + n * PZ + (warp * 16 + g + hrow * 8) * 16 + tig * 4);
      const float2 s = unpack_half2(*sp);
      const uint32_t o = pack_half2(s.x + acc[n][hrow * 2] + b0, s.y + acc[n][hrow * 2 + 1] + b1);
      *sp = o;
      if (ln_next) {   // keep the fp16-rounded values: the LayerNorm sees what the next kernel would read from HBM
        const float2 r = unpack_half2(o);
        acc[n][hrow * 2] = r.x; acc[n][hrow * 2 + 1] = r.y;
        s1[hrow] += r.x + r.y;
        s2[hrow] = fmaf(r.x, r.x, fmaf(r.y, r.y, s2[hrow]));
      }
    }
  }
  if (ln_next) {
    // a quad (tig = 0..3) holds the 64 channels of a pixel: two shuffles finish the statistics; the normalised values go to
    // the z planes (this warp's 16 rows of sz were only read by this warp's ldmatrix above: no CTA barrier needed)
    float rstd[2], nmr[2];
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      s1[hrow] += __shfl_xor_sync(0xffffffffu, s1[hrow], 1);
      s2[hrow] += __shfl_xor_sync(0xffffffffu, s2[hrow], 1);
      s1[hrow] += __shfl_xor_sync(0xffffffffu, s1[hrow], 2);
      s2[hrow] += __shfl_xor_sync(0xffffffffu, s2[hrow], 2);
      const float mu = s1[hrow] * (1.f / C);
      rstd[hrow] = rsqrtf(fmaxf(s2[hrow] * (1.f / C) - mu * mu, 0.f) + 1e-6f);
      nmr[hrow] = -mu * rstd[hrow];
    }
    __syncwarp();
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const float2 gm = __ldg(reinterpret_cast<const float2 *>(d.ln_next + n * 8 + tig * 2));
      const float2 bt = __ldg(reinterpret_cast<const float2 *>(d.ln_next + C + n * 8 + tig * 2));
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        uint32_t *sp = reinterpret_cast<uint32_t *>(sz + n * PZ + (warp * 16 + g + hrow * 8) * 16 + tig * 4);
        *sp = pack_half2(fmaf(fmaf(acc[n][hrow * 2], rstd[hrow], nmr[hrow]), gm.x, bt.x),
                         fmaf(fmaf(acc[n][hrow * 2 + 1], rstd[hrow], nmr[hrow]), gm.y, bt.y));
      }
    }
  }
  __syncthreads();
  __half *og = reinterpret_cast<__half *>(d.out) + (size_t)t * frame;
  for (int i = tid; i < MP * KC; i += 256) {
    const int ch = i % KC, p = i / KC;
    if (p0 + p < hw)
      *reinterpret_cast<uint4 *>(og + (size_t)(p0 + p) * C + ch * 8) = *reinterpret_cast<const uint4 *>(ss + ch * PZ + p * 16);
  }
  // Optional (C = 64): the LayerNorm of the NEXT block (CAB1.norm, gshift_deblur2.py:209) of the fp16 `out` values, staged in
  // the (dead) z planes above and written here in the k-chunk planar layout [T][C/8][H*W][8] that the pre-normalised pass A
  // lands by TMA (GsnCabPassA.a1_pre); lanes run along the pixels of one plane -> 512-byte contiguous stores.
  if (C == 64 && d.a1_next) {
    __half *ag = reinterpret_cast<__half *>(d.a1_next) + (size_t)t * KC * hw * 8;
    for (int i = tid; i < MP * KC; i += 256) {
      const int ch = i / MP, p = i % MP;
      if (p0 + p < hw)
        *reinterpret_cast<uint4 *>(ag + ((size_t)ch * hw + (p0 + p)) * 8) = *reinterpret_cast<const uint4 *>(sz + ch * PZ + p * 16);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// denoise variants: mid CALayer2 folded into W2, and the pointwise tail of pass A as its own kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cab_fold_mid_kernel(const float *__restrict__ partial, int ntiles, float inv_hw,
                                                           const float *__restrict__ w_du0, const float *__restrict__ w_du2,
                                                           int cr, const float *__restrict__ w2, int C,
                                                           __half *__restrict__ w2eff) {
  __shared__ float mean[128], hid[128], sc[128], part[256];
  const int t = blockIdx.x, tid = threadIdx.x;
  {
    const int nparts = 256 / C, ch = tid % C, pi = tid / C;
    float a = 0.f;
    if (pi < nparts) {
      const float *p = partial + (size_t)t * ntiles * C + ch;
      for (int i = pi; i < ntiles; i += nparts) a += p[(size_t)i * C];
    }
    part[tid] = a;
    __syncthreads();
    if (tid < C) {
      float m = 0.f;
      for (int q = 0; q < nparts; ++q) m += part[q * C + tid];
      mean[tid] = m * inv_hw;
    }
  }
  __syncthreads();
  if (tid < cr) {
    float a = 0.f;
    for (int i = 0; i < C; ++i) a += w_du0[tid * C + i] * mean[i];
    hid[tid] = a > 0.f ? a : 0.f;
  }
  __syncthreads();
  if (tid < C) {
    float a = 0.f;
    for (int i = 0; i < cr; ++i) a += w_du2[tid * cr + i] * hid[i];
    sc[tid] = 1.f / (1.f + expf(-a));
  }
  __syncthreads();
  __half *wt = w2eff + (size_t)t * 2 * C * C;
  for (int i = tid; i < 2 * C * C; i += 256) {
    const int n = i / C, k = i - n * C;
    wt[((k >> 3) * 2 * C + n) * 8 + (k & 7)] = __float2half_rn(w2[n * C + k] * sc[k]);
  }
}

template <int C>
__global__ void __launch_bounds__(256, 3) cab_pass_a2_kernel(const __half *__restrict__ u, const __half *__restrict__ w2eff,
                                                          __half *__restrict__ z, float *__restrict__ chan_partial,
                                                          long long hw, size_t w_frame_stride /* halves; 0 = one weight for all frames */) {
  constexpr int MP = 128, KC = C / 8, PZ = (MP + 1) * 16, N = 2 * C, NTH = C / 8;
  extern __shared__ __align__(128) unsigned char smem_a2[];
  unsigned char *su = smem_a2, *sw = smem_a2 + KC * PZ;
  float *red = reinterpret_cast<float *>(smem_a2 + KC * PZ + KC * N * 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * MP;
  const size_t frame = (size_t)hw * C;
  {
    const unsigned char *wg = reinterpret_cast<const unsigned char *>(w2eff + (size_t)t * w_frame_stride);
    for (int i = tid; i < KC * N; i += 256) cp_async16(sw + i * 16, wg + i * 16, true);
    const __half *ug = u + (size_t)t * frame;
    for (int i = tid; i < MP * KC; i += 256) {
      const int ch = i % KC, p = i / KC;
      const bool valid = p0 + p < hw;
      cp_async16(su + ch * PZ + p * 16, ug + (valid ? (size_t)(p0 + p) * C + ch * 8 : 0), valid);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  }
  const int g = lane >> 2, tig = lane & 3;
  float acc[2 * NTH][4];
#pragma unroll
  for (int n = 0; n < 2 * NTH; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
  const uint32_t u_s = smem_u32(su), w_s = smem_u32(sw);
#pragma unroll
  for (int k = 0; k < KC / 2; ++k) {
    uint32_t a[4];
    ldmatrix_x4(a[0], a[1], a[2], a[3], u_s + (2 * k + (lane >> 4)) * PZ + (warp * 16 + (lane & 15)) * 16);
#pragma unroll
    for (int np = 0; np < NTH; ++np) {
      uint32_t b[4];
      ldmatrix_x4(b[0], b[1], b[2], b[3], w_s + ((2 * k + ((lane >> 3) & 1)) * N + np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * 16);
      mma16816(acc[2 * np], a, b[0], b[1]);
      mma16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
  __syncwarp();
  float csum[NTH][2];
#pragma unroll
  for (int n = 0; n < NTH; ++n) {
    csum[n][0] = csum[n][1] = 0.f;
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int p = warp * 16 + g + hrow * 8;
      const float z0 = acc[n][hrow * 2] * sigmoidf_fast(acc[n + NTH][hrow * 2]);
      const float z1 = acc[n][hrow * 2 + 1] * sigmoidf_fast(acc[n + NTH][hrow * 2 + 1]);
      if (p0 + p < hw) { csum[n][0] += z0; csum[n][1] += z1; }
      // each warp only touches the rows of su it has finished reading: reuse them as the z staging tile
      *reinterpret_cast<uint32_t *>(su + n * PZ + p * 16 + tig * 4) = pack_half2(z0, z1);
    }
  }
#pragma unroll
  for (int n = 0; n < NTH; ++n)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float v = csum[n][j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (g == 0) red[warp * C + n * 8 + tig * 2 + j] = v;
    }
  __syncthreads();
  __half *zg = z + (size_t)t * frame;
  for (int i = tid; i < MP * KC; i += 256) {
    const int ch = i % KC, p = i / KC;
    if (p0 + p < hw)
      *reinterpret_cast<uint4 *>(zg + (size_t)(p0 + p) * C + ch * 8) = *reinterpret_cast<const uint4 *>(su + ch * PZ + p * 16);
  }
  if (tid < C) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w * C + tid];
    chan_partial[((size_t)t * gridDim.x + blockIdx.x) * C + tid] = s;
  }
}

int cab_pass_a_pre_dispatch(const GsnCabPassA &d, cudaStream_t st); // cab_pass_a_pre.cu
int cab_pass_a_stream_dispatch(const GsnCabPassA &d, cudaStream_t st); // cab_pass_a_stream.cu
int cab_pass_b_tc_dispatch(const GsnCabPassB &d, cudaStream_t st);  // cab_pass_b_tc.cu
int cab_pass_b80_tc_dispatch(const GsnCabPassB &d, cudaStream_t st);  // cab_pass_b80_tc.cu
bool pass_a_stream_enabled();
int pass_a_stream_tiles(int T, int H, int W);
// Two pass-A kernels serve C = 64: the 16x16-tile kernel (cab_pass_a_pre.cu, default) and the row-streaming warp-specialised
// kernel (cab_pass_a_stream.cu; deblur nets only: no mid channel attention).  They measure the same on B200,
// the tile kernel moves fewer DRAM bytes, so it stays the default; GSN_PASS_A_STREAM=1 (process-wide) or
// debug_stage == GSN_PASS_A_FORCE_STREAM (per call, tests) select the streaming kernel.
static bool use_stream(int C, int mid_ca, int debug_stage) {
  if (C != 64 || mid_ca) return false;
  if (debug_stage == GSN_PASS_A_FORCE_STREAM) return true;
  return pass_a_stream_enabled() && !debug_stage;
}

}  // namespace gsn

extern "C" int gsn_cab_tiles(int mode, int H, int W) {
  (void)mode;
  return ((H + 15) / 16) * ((W + 15) / 16);
}

extern "C" int gsn_cab_pass_a(const GsnCabPassA *dp, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(dp != nullptr, "cab_pass_a: null descriptor");
  const GsnCabPassA &d = *dp;
  GSN_REQUIRE(d.wblob && d.z && d.chan_partial, "cab_pass_a: null pointer");
  GSN_REQUIRE(d.a1_pre, "cab_pass_a: a1_pre (the LayerNorm'd planar operand of gsn_ln_planar / gsn_shift_conv1_ln / pass B) is required");
  GSN_REQUIRE(d.T > 0 && d.H > 0 && d.W > 0, "cab_pass_a: empty shape");
  GSN_REQUIRE(d.mode >= GSN_MODE_CAB1 && d.mode <= GSN_MODE_CAB2_REV, "cab_pass_a: mode=%d", d.mode);
  GSN_REQUIRE(d.debug_stage == 0 || d.debug_out, "cab_pass_a: debug_stage without debug_out");
  if (use_stream(d.C, d.mid_ca, d.debug_stage)) return cab_pass_a_stream_dispatch(d, reinterpret_cast<cudaStream_t>(stream));
  return cab_pass_a_pre_dispatch(d, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gsn_cab_pass_a_tiles(int T, int H, int W, int C, int mid_ca, int debug_stage) {
  if (gsn::use_stream(C, mid_ca, debug_stage)) return gsn::pass_a_stream_tiles(T, H, W);
  return ((H + 15) / 16) * ((W + 15) / 16);
}

extern "C" int gsn_cab_fold_mid(const float *partial, int ntiles, float inv_hw, const float *w_du0, const float *w_du2, int cr,
                                const float *w2, int C, int T, void *w2eff, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(partial && w_du0 && w_du2 && w2 && w2eff, "cab_fold_mid: null pointer");
  GSN_REQUIRE(C > 0 && C <= 128 && C % 8 == 0 && cr > 0 && cr <= 128 && ntiles > 0 && T > 0, "cab_fold_mid: bad sizes C=%d cr=%d", C, cr);
  cab_fold_mid_kernel<<<T, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(partial, ntiles, inv_hw, w_du0, w_du2, cr, w2, C,
                                                                              reinterpret_cast<__half *>(w2eff));
  count_launch();
  return check_launch("cab_fold_mid");
}

extern "C" int gsn_cab_tiles_linear(long long hw) { return (int)((hw + 127) / 128); }

extern "C" int gsn_cab_pass_a2(const void *u, const void *w2eff, void *z, float *chan_partial, int T, int H, int W, int C,
                               int per_frame_weights, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(u && w2eff && z && chan_partial, "cab_pass_a2: null pointer");
  GSN_REQUIRE(T > 0 && H > 0 && W > 0, "cab_pass_a2: empty shape");
  const long long hw = (long long)H * W;
  dim3 grid((unsigned)((hw + 127) / 128), T);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t stride = per_frame_weights ? (size_t)2 * C * C : 0;
  if (C == 64) {
    constexpr int smem = 8 * 129 * 16 + 8 * 128 * 16 + 8 * 64 * 4;
    GSN_ONCE_PER_DEVICE(cudaFuncSetAttribute(cab_pass_a2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cab_pass_a2_kernel<64><<<grid, 256, smem, st>>>(reinterpret_cast<const __half *>(u), reinterpret_cast<const __half *>(w2eff),
                                                    reinterpret_cast<__half *>(z), chan_partial, hw, stride);
  } else if (C == 80) {
    constexpr int smem = 10 * 129 * 16 + 10 * 160 * 16 + 8 * 80 * 4;
    GSN_ONCE_PER_DEVICE(cudaFuncSetAttribute(cab_pass_a2_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cab_pass_a2_kernel<80><<<grid, 256, smem, st>>>(reinterpret_cast<const __half *>(u), reinterpret_cast<const __half *>(w2eff),
                                                    reinterpret_cast<__half *>(z), chan_partial, hw, stride);
  } else {
    set_error("cab_pass_a2: C=%d unsupported (64, 80)", C);
    return GSN_E_UNSUPPORTED;
  }
  count_launch();
  return check_launch("cab_pass_a2");
}

extern "C" int gsn_cab_fold(const float *partial, int ntiles, float inv_hw, const float *w_du0, const float *w_du2, int cr,
                            const float *w3, const float *beta, const float *bias3, int C, int T, void *weff, float *beff,
                            void *stream) {
  using namespace gsn;
  GSN_REQUIRE(partial && w_du0 && w_du2 && w3 && beta && weff && beff, "cab_fold: null pointer");
  GSN_REQUIRE(C > 0 && C <= 128 && C % 8 == 0 && cr > 0 && cr <= 128 && ntiles > 0 && T > 0, "cab_fold: bad sizes C=%d cr=%d", C, cr);
  cab_fold_kernel<<<T, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(partial, ntiles, inv_hw, w_du0, w_du2, cr, w3, beta,
                                                                          bias3, C, reinterpret_cast<__half *>(weff), beff);
  count_launch();
  return check_launch("cab_fold");
}

extern "C" int gsn_cab_pass_b(const GsnCabPassB *dp, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(dp != nullptr, "cab_pass_b: null descriptor");
  const GsnCabPassB &d = *dp;
  GSN_REQUIRE(d.x && d.z && d.weff && d.beff && d.out, "cab_pass_b: null pointer");
  GSN_REQUIRE(!d.a1_next || (d.ln_next && d.C == 64), "cab_pass_b: a1_next needs ln_next and C=64");
  GSN_REQUIRE(d.T > 0 && d.H > 0 && d.W > 0, "cab_pass_b: empty shape");
  GSN_REQUIRE(d.mode >= GSN_MODE_CAB1 && d.mode <= GSN_MODE_CAB2_REV, "cab_pass_b: mode=%d", d.mode);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long hw = (long long)d.H * d.W;
  // the TMA / tcgen05 streaming kernels (cab_pass_b_tc.cu for C = 64, cab_pass_b80_tc.cu for C = 80); GSN_PASS_B_TC=0 keeps the mma.sync kernel below
  static const bool want_tc = [] { const char *e = getenv("GSN_PASS_B_TC"); return !(e && e[0] == '0'); }();
  if (d.C == 64 && want_tc) return cab_pass_b_tc_dispatch(d, st);
  if (d.C == 80 && want_tc) return cab_pass_b80_tc_dispatch(d, st);     // planar-chunk variant for the Ours+ width (cab_pass_b80_tc.cu)
  dim3 grid((unsigned)((hw + 127) / 128), d.T);
  if (d.C == 64) {
    constexpr int smem = 2 * 8 * 129 * 16 + 8 * 64 * 16;
    GSN_ONCE_PER_DEVICE(cudaFuncSetAttribute(cab_pass_b_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cab_pass_b_kernel<64><<<grid, 256, smem, st>>>(d);
  } else if (d.C == 80) {
    constexpr int smem = 2 * 10 * 129 * 16 + 10 * 80 * 16;
    GSN_ONCE_PER_DEVICE(cudaFuncSetAttribute(cab_pass_b_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cab_pass_b_kernel<80><<<grid, 256, smem, st>>>(d);
  } else {
    set_error("cab_pass_b: C=%d unsupported (64, 80)", d.C);
    return GSN_E_UNSUPPORTED;
  }
  count_launch();
  return check_launch("cab_pass_b");
}
