// I/O side of the evaluation loop (reference: inference/test_deblur_small.py:134-143,191-200):
//   * gsn_u8_to_clip : the uint8 HWC frames go to the device as they are (1 byte per sample instead of a float64 -> float32
//     CPU pass and 2-4 bytes over PCIe) and become the (T,3,H,W) clip in [0,1] here, with the reference's arithmetic
//     (numpy2tensor: float32(u8) * float32(1/255), then .half());
//   * gsn_psnr_sse   : per-frame sum of squared errors between clamp(out, 0, 1) * 255 (float32, not rounded: exactly what the
//     reference hands to skimage's PSNR) and the uint8 ground truth, accumulated in float64 -- only 64 doubles per frame go
//     back to the host, which finishes 10 log10(255^2 / mse) in a fixed order (deterministic).
#include "common.cuh"

namespace gsn {

constexpr int kSseBlocks = 64;   // partial sums per frame

template <typename T>
__device__ __forceinline__ T from_float(float v);
template <>
__device__ __forceinline__ float from_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }
__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__half v) { return __half2float(v); }

// frames (T,H,W,3) uint8 -> clip (T,3,H,W): one thread per pixel, coalesced plane stores
template <typename T>
__global__ void __launch_bounds__(256) u8_to_clip_kernel(const unsigned char *__restrict__ frames, long long hw, long long total_px,
                                                         T *__restrict__ clip) {
  const float k = (float)(1.0 / 255.0);       // the reference multiplies by the float32-rounded reciprocal (tensor.mul_(1/255))
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total_px; p += (long long)gridDim.x * blockDim.x) {
    const long long t = p / hw, q = p - t * hw;
    const unsigned char *s = frames + p * 3;
    T *d = clip + t * 3 * hw + q;
    d[0] = from_float<T>(__fmul_rn((float)s[0], k));
    d[hw] = from_float<T>(__fmul_rn((float)s[1], k));
    d[2 * hw] = from_float<T>(__fmul_rn((float)s[2], k));
  }
}

// out (T,3,H,W), gt (T,H,W,3) uint8 -> partial[t][kSseBlocks] float64
template <typename T>
__global__ void __launch_bounds__(256) psnr_sse_kernel(const T *__restrict__ out, const unsigned char *__restrict__ gt, long long hw,
                                                       double *__restrict__ partial) {
  __shared__ double red[256];
  const int t = blockIdx.y, b = blockIdx.x, tid = threadIdx.x;
  const T *o = out + (size_t)t * 3 * hw;
  const unsigned char *g = gt + (size_t)t * 3 * hw;
  double acc = 0.0;
  for (long long q = (long long)b * 256 + tid; q < hw; q += (long long)kSseBlocks * 256) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __fmul_rn(fminf(fmaxf(to_float(o[c * hw + q]), 0.f), 1.f), 255.f);   // clamp(0, 1) * 255 in float32
      const double dlt = (double)v - (double)g[q * 3 + c];
      acc += dlt * dlt;
    }
  }
  red[tid] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {          // fixed-order tree: deterministic
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  if (tid == 0) partial[(size_t)t * kSseBlocks + b] = red[0];
}

}  // namespace gsn

extern "C" int gsn_u8_to_clip(const void *frames_u8, int T, int H, int W, int dtype, void *clip, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(frames_u8 && clip, "u8_to_clip: null pointer");
  GSN_REQUIRE(T > 0 && H > 0 && W > 0, "u8_to_clip: empty shape");
  GSN_REQUIRE(dtype == GSN_DTYPE_F16 || dtype == GSN_DTYPE_F32, "u8_to_clip: dtype=%d", dtype);
  const long long hw = (long long)H * W, total = hw * T;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned char *f = reinterpret_cast<const unsigned char *>(frames_u8);
  if (dtype == GSN_DTYPE_F16) u8_to_clip_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>(f, hw, total, reinterpret_cast<__half *>(clip));
  else u8_to_clip_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(f, hw, total, reinterpret_cast<float *>(clip));
  count_launch();
  return check_launch("u8_to_clip");
}

extern "C" int gsn_psnr_sse_blocks(void) { return gsn::kSseBlocks; }

extern "C" int gsn_psnr_sse(const void *out, int dtype, const void *gt_u8, int T, int H, int W, double *partial, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(out && gt_u8 && partial, "psnr_sse: null pointer");
  GSN_REQUIRE(T > 0 && H > 0 && W > 0, "psnr_sse: empty shape");
  GSN_REQUIRE(dtype == GSN_DTYPE_F16 || dtype == GSN_DTYPE_F32, "psnr_sse: dtype=%d", dtype);
  const long long hw = (long long)H * W;
  dim3 grid(kSseBlocks, T);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned char *g = reinterpret_cast<const unsigned char *>(gt_u8);
  if (dtype == GSN_DTYPE_F16) psnr_sse_kernel<__half><<<grid, 256, 0, st>>>(reinterpret_cast<const __half *>(out), g, hw, partial);
  else psnr_sse_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float *>(out), g, hw, partial);
  count_launch();
  return check_launch("psnr_sse");
}
