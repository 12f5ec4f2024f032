// The two layout-changing convolutions at the ends of GShiftNet.forward:
//   conv_in : user clip (T,cin,H,W) NCHW fp16/fp32  -> NHWC fp16 features   (feat_extract[0], d2:710,750-752)
//   conv_out: NHWC fp16 features -> (T,3,H,W) NCHW + input residual          (conv_last + shortcut, d2:712,745,756)
// Both are tiny in FLOPs (3/4 input or 3 output channels) and purely HBM-bound, so they run on CUDA
// cores with one thread per pixel; weights are broadcast from shared memory.
#include "common.cuh"

namespace gsn {

template <typename TIn, int COUT_P>
__global__ void __launch_bounds__(256) conv_in_kernel(const TIn *__restrict__ x, int T, int cin, int H, int W,
                                                      const float *__restrict__ w, const float *__restrict__ bias,
                                                      __half *__restrict__ dst, const TIn *__restrict__ nm = nullptr,
                                                      long long nm_st = 0, long long nm_sy = 0, long long nm_sx = 0) {
  // nm != null: the denoise nets' torch.cat((x, noise_map), 1) (gshift_denoise2.py:749) folded into the load -- channel `cin` of
  // the conv input is read from the (possibly expand()ed, i.e. zero-stride) noise map instead of a concatenated copy
  __shared__ __align__(16) float sw[9 * 4 * COUT_P];
  __shared__ __align__(16) float sb[COUT_P];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int cw = cin + (nm ? 1 : 0);      // channels of the conv weight
  for (int i = tid; i < 9 * cw * COUT_P; i += 256) sw[i] = w[i];
  if (tid < COUT_P) sb[tid] = bias ? bias[tid] : 0.f;
  __syncthreads();
  const int px = blockIdx.x * 32 + threadIdx.x, py = blockIdx.y * 8 + threadIdx.y, t = blockIdx.z;
  if (px >= W || py >= H) return;
  float acc[COUT_P];
#pragma unroll
  for (int c = 0; c < COUT_P; ++c) acc[c] = sb[c];
  const size_t plane = (size_t)H * W;
  const TIn *xt = x + (size_t)t * cin * plane;
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = py + ky - 1;
    if (yy < 0 || yy >= H) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = px + kx - 1;
      if (xx < 0 || xx >= W) continue;
      for (int ci = 0; ci < cin; ++ci) {
        const float v = (float)xt[ci * plane + (size_t)yy * W + xx];
        const float4 *wr = reinterpret_cast<const float4 *>(sw + ((ky * 3 + kx) * cw + ci) * COUT_P);
#pragma unroll
        for (int q = 0; q < COUT_P / 4; ++q) {
          const float4 ww = wr[q];
          acc[q * 4 + 0] += v * ww.x; acc[q * 4 + 1] += v * ww.y;
          acc[q * 4 + 2] += v * ww.z; acc[q * 4 + 3] += v * ww.w;
        }
      }
      if (nm) {
        const float v = (float)nm[t * nm_st + yy * nm_sy + xx * nm_sx];
        const float4 *wr = reinterpret_cast<const float4 *>(sw + ((ky * 3 + kx) * cw + cin) * COUT_P);
#pragma unroll
        for (int q = 0; q < COUT_P / 4; ++q) {
          const float4 ww = wr[q];
          acc[q * 4 + 0] += v * ww.x; acc[q * 4 + 1] += v * ww.y;
          acc[q * 4 + 2] += v * ww.z; acc[q * 4 + 3] += v * ww.w;
        }
      }
    }
  }
  uint4 *o = reinterpret_cast<uint4 *>(dst + (((size_t)t * H + py) * W + px) * COUT_P);
#pragma unroll
  for (int q = 0; q < COUT_P / 8; ++q) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = acc[q * 8 + j];
    o[q] = pack8(f);
  }
}

template <typename TIo, int CP, int KS>
__global__ void __launch_bounds__(256) conv_out_kernel(const __half *__restrict__ src, const float *__restrict__ w,
                                                       const TIo *__restrict__ resid, int cres, int T, int H, int W,
                                                       TIo *__restrict__ dst) {
  constexpr int R = KS / 2, TW = 32 + 2 * R, TH = 8 + 2 * R;
  __shared__ __align__(16) __half tile[TH * TW * CP];
  __shared__ __align__(16) float sw[KS * KS * CP * 3];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int t = blockIdx.z, x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
  for (int i = tid; i < KS * KS * CP * 3; i += 256) sw[i] = w[i];
  constexpr int CH = CP / 8;
  for (int i = tid; i < TH * TW * CH; i += 256) {
    const int ch = i % CH, p = i / CH, ly = p / TW, lx = p - ly * TW;
    const int gy = y0 + ly - R, gx = x0 + lx - R;
    const bool valid = gy >= 0 && gy < H && gx >= 0 && gx < W;
    const __half *sp = valid ? src + (((size_t)t * H + gy) * W + gx) * CP + ch * 8 : src;
    cp_async16(tile + (size_t)p * CP + ch * 8, sp, valid);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int px = x0 + threadIdx.x, py = y0 + threadIdx.y;
  if (px >= W || py >= H) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int ky = 0; ky < KS; ++ky)
    for (int kx = 0; kx < KS; ++kx) {
      const uint4 *tp = reinterpret_cast<const uint4 *>(tile + ((threadIdx.y + ky) * TW + threadIdx.x + kx) * CP);
      const float *wr = sw + (ky * KS + kx) * CP * 3;
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        float f[8];
        unpack8(tp[q], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float *wc = wr + (q * 8 + j) * 3;
          a0 += f[j] * wc[0]; a1 += f[j] * wc[1]; a2 += f[j] * wc[2];
        }
      }
    }
  const size_t plane = (size_t)H * W, pix = (size_t)py * W + px;
  const TIo *rt = resid + (size_t)t * cres * plane + pix;
  TIo *dt = dst + (size_t)t * 3 * plane + pix;
  dt[0] = (TIo)(a0 + (float)rt[0]);
  dt[plane] = (TIo)(a1 + (float)rt[plane]);
  dt[2 * plane] = (TIo)(a2 + (float)rt[2 * plane]);
}

}  // namespace gsn

static int conv_in_impl(const void *x, int x_dtype, int T, int cin, int H, int W, const void *nm, long long nm_st, long long nm_sy,
                        long long nm_sx, const float *w, const float *bias, int cout_p, void *dst, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(x && w && dst, "conv_in: null pointer");
  GSN_REQUIRE(cin >= 1 && cin + (nm ? 1 : 0) <= 4, "conv_in: cin=%d (expected 3, or 3 + noise map, or 4)", cin);
  GSN_REQUIRE(T > 0 && H > 0 && W > 0, "conv_in: empty shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((W + 31) / 32, (H + 7) / 8, T), block(32, 8);
  __half *d = reinterpret_cast<__half *>(dst);
#define GSN_CI(TIN, CP) conv_in_kernel<TIN, CP><<<grid, block, 0, st>>>((const TIN *)x, T, cin, H, W, w, bias, d, (const TIN *)nm, nm_st, nm_sy, nm_sx)
  if (cout_p == 16 && x_dtype == GSN_DTYPE_F16) GSN_CI(__half, 16);
  else if (cout_p == 16 && x_dtype == GSN_DTYPE_F32) GSN_CI(float, 16);
  else if (cout_p == 24 && x_dtype == GSN_DTYPE_F16) GSN_CI(__half, 24);
  else if (cout_p == 24 && x_dtype == GSN_DTYPE_F32) GSN_CI(float, 24);
  else if (cout_p == 32 && x_dtype == GSN_DTYPE_F16) GSN_CI(__half, 32);
  else if (cout_p == 32 && x_dtype == GSN_DTYPE_F32) GSN_CI(float, 32);
  else {
    set_error("conv_in: cout_p=%d dtype=%d unsupported", cout_p, x_dtype);
    return GSN_E_UNSUPPORTED;
  }
#undef GSN_CI
  count_launch();
  return check_launch("conv_in");
}

extern "C" int gsn_conv_in(const void *x, int x_dtype, int T, int cin, int H, int W, const float *w, const float *bias,
                           int cout_p, void *dst, void *stream) {
  return conv_in_impl(x, x_dtype, T, cin, H, W, nullptr, 0, 0, 0, w, bias, cout_p, dst, stream);
}

extern "C" int gsn_conv_in_nm(const void *x, int x_dtype, int T, int cin, int H, int W, const void *noise_map, long long nm_stride_t,
                              long long nm_stride_y, long long nm_stride_x, const float *w, const float *bias, int cout_p, void *dst,
                              void *stream) {
  if (!noise_map) {
    gsn::set_error("conv_in_nm: null noise map");
    return GSN_E_BADARG;
  }
  return conv_in_impl(x, x_dtype, T, cin, H, W, noise_map, nm_stride_t, nm_stride_y, nm_stride_x, w, bias, cout_p, dst, stream);
}

extern "C" int gsn_conv_out(const void *src, int cp, int ks, const float *w, const void *resid, int cres, int x_dtype,
                            int T, int H, int W, void *dst, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(src && w && resid && dst, "conv_out: null pointer");
  GSN_REQUIRE(T > 0 && H > 0 && W > 0, "conv_out: empty shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((W + 31) / 32, (H + 7) / 8, T), block(32, 8);
  const __half *s = reinterpret_cast<const __half *>(src);
#define GSN_CO(TIO, CP, KS) \
  conv_out_kernel<TIO, CP, KS><<<grid, block, 0, st>>>(s, w, (const TIO *)resid, cres, T, H, W, (TIO *)dst)
  if (x_dtype == GSN_DTYPE_F16) {
    if (cp == 16 && ks == 5) GSN_CO(__half, 16, 5);
    else if (cp == 16 && ks == 3) GSN_CO(__half, 16, 3);
    else if (cp == 32 && ks == 5) GSN_CO(__half, 32, 5);
    else if (cp == 32 && ks == 3) GSN_CO(__half, 32, 3);
    else if (cp == 24 && ks == 5) GSN_CO(__half, 24, 5);
    else if (cp == 24 && ks == 3) GSN_CO(__half, 24, 3);
    else { set_error("conv_out: cp=%d ks=%d unsupported", cp, ks); return GSN_E_UNSUPPORTED; }
  } else if (x_dtype == GSN_DTYPE_F32) {
    if (cp == 16 && ks == 5) GSN_CO(float, 16, 5);
    else if (cp == 16 && ks == 3) GSN_CO(float, 16, 3);
    else if (cp == 32 && ks == 5) GSN_CO(float, 32, 5);
    else if (cp == 32 && ks == 3) GSN_CO(float, 32, 3);
    else if (cp == 24 && ks == 5) GSN_CO(float, 24, 5);
    else if (cp == 24 && ks == 3) GSN_CO(float, 24, 3);
    else { set_error("conv_out: cp=%d ks=%d unsupported", cp, ks); return GSN_E_UNSUPPORTED; }
  } else {
    set_error("conv_out: dtype=%d unsupported", x_dtype);
    return GSN_E_UNSUPPORTED;
  }
#undef GSN_CO
  count_launch();
  return check_launch("conv_out");
}
