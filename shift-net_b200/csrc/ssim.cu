// SSIM of the evaluation loop on the device (reference: inference/test_deblur_small.py:25-49).
//
// The reference computes, per restored frame, on /255 CHW float32 images a = (clamp(out,0,1)*255)/255 and b = gt/255:
//   mu = G(a), G(b) ; sigma = G(a*a) - mu1^2, G(b*b) - mu2^2, G(a*b) - mu1*mu2 ; mean of (2 mu1 mu2 + C1)(2 s12 + C2) / ((mu1^2 + mu2^2 + C1)(s1 + s2 + C2))
// where G = scipy.ndimage.gaussian_filter(sigma = 1.5) on the 3-D (3, H, W) array: a separable 13-tap Gaussian (truncate 4.0 ->
// radius 6) along ALL THREE axes -- the channel axis included -- with the half-sample-symmetric 'reflect' boundary, axis 0 first,
// every axis pass accumulating in float64 and rounding its result to float32.  The kernels below reproduce exactly those
// rounding points (explicit _rn intrinsics: the library is built with --use_fast_math), so the result agrees with scipy to
// float32 round-off; only one float64 partial sum per (frame, block) goes back to the host.
//   pass 1 (pointwise): a, b, a*a, b*b, a*b and the channel-axis filter (a fixed 3x3 matrix: 13 taps reflected onto 3 channels)
//   pass 2: the filter along H ; pass 3: the filter along W, the SSIM map and its block sums
#include <cmath>

#include "common.cuh"

namespace gsn {

constexpr int kSsimBlocks = 128;     // partial sums per frame
constexpr int kRad = 6;              // int(4.0 * 1.5 + 0.5)

struct SsimTaps {
  double w[2 * kRad + 1];            // normalised Gaussian, sigma 1.5
  double m[3][3];                    // channel-axis pass: out[co] = sum_ci m[co][ci] in[ci]
};

__host__ __device__ inline int reflect_index(int i, int n) {   // scipy 'reflect': d c b a | a b c d | d c b a
  while (i < 0 || i >= n) {
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - 1 - i;
  }
  return i;
}

static SsimTaps make_taps() {
  SsimTaps t;
  double s = 0.0;
  for (int k = -kRad; k <= kRad; ++k) { t.w[k + kRad] = exp(-0.5 / (1.5 * 1.5) * (double)(k * k)); s += t.w[k + kRad]; }
  for (int k = 0; k <= 2 * kRad; ++k) t.w[k] /= s;
  for (int co = 0; co < 3; ++co) {
    for (int ci = 0; ci < 3; ++ci) t.m[co][ci] = 0.0;
    for (int k = -kRad; k <= kRad; ++k) t.m[co][reflect_index(co + k, 3)] += t.w[k + kRad];
  }
  return t;
}

__device__ __forceinline__ float out_as_float(const float *p, size_t i) { return p[i]; }
__device__ __forceinline__ float out_as_float(const __half *p, size_t i) { return __half2float(p[i]); }

// ws1[t][q][c][y][x] (q: a, b, aa, bb, ab), float32
template <typename T>
__global__ void __launch_bounds__(256) ssim_point_kernel(const T *__restrict__ out, const unsigned char *__restrict__ gt, long long hw,
                                                         const SsimTaps taps, float *__restrict__ ws1) {
  const int t = blockIdx.y;
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= hw) return;
  float q[5][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float o = out_as_float(out, ((size_t)t * 3 + c) * hw + p);
    const float a = __fdiv_rn(__fmul_rn(fminf(fmaxf(o, 0.f), 1.f), 255.f), 255.f);      // (clamp(out,0,1) * 255) / 255, float32
    const float b = __fdiv_rn((float)gt[((size_t)t * hw + p) * 3 + c], 255.f);
    q[0][c] = a; q[1][c] = b; q[2][c] = __fmul_rn(a, a); q[3][c] = __fmul_rn(b, b); q[4][c] = __fmul_rn(a, b);
  }
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int co = 0; co < 3; ++co) {
      const double v = taps.m[co][0] * (double)q[k][0] + taps.m[co][1] * (double)q[k][1] + taps.m[co][2] * (double)q[k][2];
      ws1[(((size_t)t * 5 + k) * 3 + co) * hw + p] = (float)v;
    }
}

// filter along H: planes = T*15
__global__ void __launch_bounds__(256) ssim_vert_kernel(const float *__restrict__ src, int H, int W, const SsimTaps taps, float *__restrict__ dst) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const float *pl = src + (size_t)blockIdx.z * H * W;
  double acc = 0.0;
#pragma unroll
  for (int k = -kRad; k <= kRad; ++k) acc += taps.w[k + kRad] * (double)pl[(size_t)reflect_index(y + k, H) * W + x];
  dst[(size_t)blockIdx.z * H * W + (size_t)y * W + x] = (float)acc;
}

// filter along W, SSIM map, block sums: partial[t][b] float64 ; every block owns a fixed set of pixels, fixed-order tree
__global__ void __launch_bounds__(256) ssim_horiz_kernel(const float *__restrict__ src, int H, int W, const SsimTaps taps, double *__restrict__ partial) {
  __shared__ double red[256];
  const int t = blockIdx.y, b = blockIdx.x, tid = threadIdx.x;
  const long long hw = (long long)H * W, n = 3 * hw;
  const float C1 = (float)(0.01 * 0.01), C2 = (float)(0.03 * 0.03);   // numpy keeps the float32 array dtype: python floats round to float32
  double acc = 0.0;
  for (long long e = (long long)b * 256 + tid; e < n; e += (long long)kSsimBlocks * 256) {
    const int c = (int)(e / hw);
    const long long p = e - (long long)c * hw;
    const int y = (int)(p / W), x = (int)(p - (long long)y * W);
    float f[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float *row = src + (((size_t)t * 5 + k) * 3 + c) * hw + (size_t)y * W;
      double s = 0.0;
#pragma unroll
      for (int j = -kRad; j <= kRad; ++j) s += taps.w[j + kRad] * (double)row[reflect_index(x + j, W)];
      f[k] = (float)s;
    }
    const float mu1 = f[0], mu2 = f[1];
    const float mu1_sq = __fmul_rn(mu1, mu1), mu2_sq = __fmul_rn(mu2, mu2), mu12 = __fmul_rn(mu1, mu2);
    const float s1 = __fsub_rn(f[2], mu1_sq), s2 = __fsub_rn(f[3], mu2_sq), s12 = __fsub_rn(f[4], mu12);
    const float num = __fmul_rn(__fadd_rn(__fmul_rn(2.f, mu12), C1), __fadd_rn(__fmul_rn(2.f, s12), C2));
    const float den = __fmul_rn(__fadd_rn(__fadd_rn(mu1_sq, mu2_sq), C1), __fadd_rn(__fadd_rn(s1, s2), C2));
    acc += (double)__fdiv_rn(num, den);
  }
  red[tid] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  if (tid == 0) partial[(size_t)t * kSsimBlocks + b] = red[0];
}

}  // namespace gsn

extern "C" int gsn_ssim_blocks(void) { return gsn::kSsimBlocks; }

extern "C" long long gsn_ssim_workspace_bytes(int T, int H, int W) { return 2LL * T * 15 * H * W * (long long)sizeof(float); }

extern "C" int gsn_ssim(const void *out, int dtype, const void *gt_u8, int T, int H, int W, void *workspace, double *partial, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(out && gt_u8 && workspace && partial, "ssim: null pointer");
  GSN_REQUIRE(T > 0 && H > 0 && W > 0, "ssim: empty shape");
  GSN_REQUIRE(dtype == GSN_DTYPE_F16 || dtype == GSN_DTYPE_F32, "ssim: dtype=%d", dtype);
  static const SsimTaps taps = make_taps();
  const long long hw = (long long)H * W;
  float *ws1 = reinterpret_cast<float *>(workspace), *ws2 = ws1 + (size_t)T * 15 * hw;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned char *g = reinterpret_cast<const unsigned char *>(gt_u8);
  dim3 g1((unsigned)((hw + 255) / 256), T);
  if (dtype == GSN_DTYPE_F16) ssim_point_kernel<__half><<<g1, 256, 0, st>>>(reinterpret_cast<const __half *>(out), g, hw, taps, ws1);
  else ssim_point_kernel<float><<<g1, 256, 0, st>>>(reinterpret_cast<const float *>(out), g, hw, taps, ws1);
  count_launch();
  dim3 g2((W + 31) / 32, (H + 7) / 8, T * 15);
  ssim_vert_kernel<<<g2, 256, 0, st>>>(ws1, H, W, taps, ws2);
  count_launch();
  dim3 g3(kSsimBlocks, T);
  ssim_horiz_kernel<<<g3, 256, 0, st>>>(ws2, H, W, taps, partial);
  count_launch();
  return check_launch("ssim");
}
