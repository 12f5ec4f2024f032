// Memory-bound glue of the CAB / U-Net plumbing: CA squeeze-excite MLP, scale+residual, bilinear x2 + skip, add.
// All activations NHWC fp16, 16-byte vector accesses, grid-stride loops sized to a multiple of the SM count.
#include "common.cuh"

namespace gsn {

static int grid_for(long long work_items, int block) {
  long long b = (work_items + block - 1) / block;
  const long long cap = 148LL * 16;  // 16 resident 256-thread CTAs per SM would exceed the RF; 8 typical -> 2 waves
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// s[t][c] = sigmoid(W2 relu(W1 mean)), mean from deterministic per-tile partial sums (CALayer, d2:54-71)
__global__ void __launch_bounds__(256) ca_scale_kernel(const float *__restrict__ partial, int ntiles, float inv_hw,
                                                       const float *__restrict__ w1, const float *__restrict__ w2, int c,
                                                       int cr, int cp, float *__restrict__ s) {
  __shared__ float part[256];
  __shared__ float mean[128];
  __shared__ float hid[128];
  const int t = blockIdx.x, tid = threadIdx.x;
  {  // deterministic two-level reduction of the per-tile sums: 256/cp slices of the tile list, then a fixed-order sum
    const int nparts = 256 / cp, ch = tid % cp, pi = tid / cp;
    float a = 0.f;
    if (pi < nparts) {
      const float *p = partial + (size_t)t * ntiles * cp + ch;
      for (int i = pi; i < ntiles; i += nparts) a += p[(size_t)i * cp];
    }
    part[tid] = a;
    __syncthreads();
    if (tid < cp) {
      float m = 0.f;
      for (int q = 0; q < nparts; ++q) m += part[q * cp + tid];
      mean[tid] = m * inv_hw;
    }
  }
  __syncthreads();
  if (tid < cr) {
    float a = 0.f;
    for (int i = 0; i < c; ++i) a += w1[tid * c + i] * mean[i];
    hid[tid] = a > 0.f ? a : 0.f;
  }
  __syncthreads();
  if (tid < cp) {
    float v = 0.f;
    if (tid < c) {
      float a = 0.f;
      for (int i = 0; i < cr; ++i) a += w2[tid * cr + i] * hid[i];
      v = 1.f / (1.f + expf(-a));
    }
    s[(size_t)t * cp + tid] = v;
  }
}

// out = x + res * s[t][c] (+ extra)
__global__ void __launch_bounds__(256) scale_residual_kernel(const uint4 *__restrict__ x, const uint4 *__restrict__ res,
                                                             const float *__restrict__ s, const uint4 *__restrict__ extra,
                                                             uint4 *__restrict__ out, long long vec_per_frame, int chunks,
                                                             int cp, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i / vec_per_frame);
    const int ch = (int)(i % chunks);
    const float4 s0 = *reinterpret_cast<const float4 *>(s + (size_t)t * cp + ch * 8);
    const float4 s1 = *reinterpret_cast<const float4 *>(s + (size_t)t * cp + ch * 8 + 4);
    float a[8], r[8];
    unpack8(__ldg(x + i), a);
    unpack8(__ldg(res + i), r);
    a[0] += r[0] * s0.x; a[1] += r[1] * s0.y; a[2] += r[2] * s0.z; a[3] += r[3] * s0.w;
    a[4] += r[4] * s1.x; a[5] += r[5] * s1.y; a[6] += r[6] * s1.z; a[7] += r[7] * s1.w;
    if (extra) {
      float e[8];
      unpack8(__ldg(extra + i), e);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += e[j];
    }
    out[i] = pack8(a);
  }
}

// dst = bilinear_x2(src) + skip ; align_corners=False: src coord = (o + 0.5)/2 - 0.5 clamped at 0
__global__ void __launch_bounds__(256) upsample2x_add_kernel(const uint4 *__restrict__ src, const uint4 *__restrict__ skip,
                                                             uint4 *__restrict__ dst, int T, int h, int w, int chunks,
                                                             long long total) {
  const int W2 = 2 * w, H2 = 2 * h;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks);
    long long p = i / chunks;
    const int ox = (int)(p % W2); p /= W2;
    const int oy = (int)(p % H2);
    const int t = (int)(p / H2);
    float sy = (oy + 0.5f) * 0.5f - 0.5f; if (sy < 0.f) sy = 0.f;
    float sx = (ox + 0.5f) * 0.5f - 0.5f; if (sx < 0.f) sx = 0.f;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = sy - y0, lx = sx - x0;
    const uint4 *base = src + (size_t)t * h * w * chunks + ch;
    float a[8], b[8], c[8], d[8], k[8];
    unpack8(__ldg(base + ((size_t)y0 * w + x0) * chunks), a);
    unpack8(__ldg(base + ((size_t)y0 * w + x1) * chunks), b);
    unpack8(__ldg(base + ((size_t)y1 * w + x0) * chunks), c);
    unpack8(__ldg(base + ((size_t)y1 * w + x1) * chunks), d);
    unpack8(__ldg(skip + i), k);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
    for (int j = 0; j < 8; ++j) k[j] += w00 * a[j] + w01 * b[j] + w10 * c[j] + w11 * d[j];
    dst[i] = pack8(k);
  }
}

__global__ void __launch_bounds__(256) add_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b,
                                                  uint4 *__restrict__ out, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float x[8], y[8];
    unpack8(__ldg(a + i), x);
    unpack8(__ldg(b + i), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    out[i] = pack8(x);
  }
}

}  // namespace gsn

extern "C" int gsn_ca_scale(const float *partial, int ntiles, float inv_hw, const float *w1, const float *w2, int c, int cr,
                            int cp, int T, float *s, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(partial && w1 && w2 && s, "ca_scale: null pointer");
  GSN_REQUIRE(c > 0 && c <= cp && cp <= 128 && cr > 0 && cr <= 128 && ntiles > 0 && T > 0, "ca_scale: bad sizes c=%d cr=%d cp=%d", c, cr, cp);
  ca_scale_kernel<<<T, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(partial, ntiles, inv_hw, w1, w2, c, cr, cp, s);
  count_launch();
  return check_launch("ca_scale");
}

extern "C" int gsn_scale_residual(const void *x, const void *res, const float *s, const void *extra, void *out, int T,
                                  long long hw, int cp, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(x && res && s && out, "scale_residual: null pointer");
  GSN_REQUIRE(cp % 8 == 0 && T > 0 && hw > 0, "scale_residual: bad sizes");
  const int chunks = cp / 8;
  const long long vpf = hw * chunks, total = vpf * T;
  scale_residual_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const uint4 *)x, (const uint4 *)res, s, (const uint4 *)extra, (uint4 *)out, vpf, chunks, cp, total);
  count_launch();
  return check_launch("scale_residual");
}

extern "C" int gsn_upsample2x_add(const void *src, const void *skip, void *dst, int T, int h, int w, int cp, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(src && skip && dst, "upsample2x_add: null pointer");
  GSN_REQUIRE(cp % 8 == 0 && T > 0 && h > 0 && w > 0, "upsample2x_add: bad sizes");
  const int chunks = cp / 8;
  const long long total = (long long)T * 4 * h * w * chunks;
  upsample2x_add_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const uint4 *)src, (const uint4 *)skip, (uint4 *)dst, T, h, w, chunks, total);
  count_launch();
  return check_launch("upsample2x_add");
}

extern "C" int gsn_add(const void *a, const void *b, void *out, long long n, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(a && b && out && n > 0 && n % 8 == 0, "add: bad arguments");
  const long long total = n / 8;
  add_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const uint4 *)a, (const uint4 *)b,
                                                                                        (uint4 *)out, total);
  count_launch();
  return check_launch("add");
}
