// Width-generic (C = 64 or 80) building blocks of the shift + NAF block, used by the Ours+ nets (C = 80, grouped
// RepConv, gshift_deblur1.py / gshift_denoise1.py) whose 2C = 160 wide tensor does not fit the single fused tcgen05
// kernel's TMEM/shared-memory budget.  The block is split at tensor boundaries, every piece is one of our kernels:
//
//   shift_ln      : temporal roll + spatial-shift gather + conv1 (dw3x3) + LayerNorm          -> A1 (T,H,W,pad16(C'))
//   conv_mma 1x1  : first 1x1 as two half-width launches (a | b halves)                        (conv_mma.cu)
//   dw_gate       : (dw3x3 + id)(a) * (dw3x3 + id)(b)  [+ per-tile sums for the denoise mid CA] -> g (T,H,W,C)
//   group_conv5   : RepConv with groups of 8 channels (5x5 + 3x3 + id) on tensor cores (mma.sync, one n-tile per group)
//   conv_mma 1x1  : second 1x1 (two halves), then gate2: a * sigmoid(b) + per-tile sums       -> z
//   cab_fold / cab_pass_b (shift_cab.cu)
//
// Same zero-padding rules and the same deterministic per-tile partial sums as the fused path.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include <type_traits>

#include "common.cuh"
#include "shift_common.cuh"

namespace gsn {

__device__ __forceinline__ void shift_offset_rt(int C, int c, int &dy, int &dx) {
  const int number = C / 2 / 8, n2 = (number - 1) / 2, n1 = number - 2 * n2;
  if (c < 16 * n2) { const int g = c / n2; dy = kShiftOuter[g][0]; dx = kShiftOuter[g][1]; }
  else { const int g = (c - 16 * n2) / n1; dy = kShiftInner[g][0]; dx = kShiftInner[g][1]; }
}

// ---------------------------------------------------------------------------------------------------------------
// shift_ln: 16 lanes per pixel, lane = 8-channel chunk of the LayerNorm input [y_lo | y_hi | conv1(shift(hw))]
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) shift_ln_kernel(const __half *__restrict__ x, int T, int H, int W, int C, int mode,
                                                       int circular, const __half *__restrict__ wc1 /*[9][C/2]*/,
                                                       const float *__restrict__ ln /*gamma[CIN], beta[CIN]*/,
                                                       __half *__restrict__ out, int cinp,
                                                       const __half *__restrict__ hw_pre /*(T,H,W,C/2) or null*/) {
  const int lane16 = threadIdx.x & 15;
  const long long hw = (long long)H * W;
  const long long pix = (long long)blockIdx.x * 16 + (threadIdx.x >> 4);
  const int t = blockIdx.y;
  const bool shift = mode != GSN_MODE_CAB1;
  const int HC = C / 2, cin = shift ? C + HC : C;
  const bool active = pix < hw;
  const int py = active ? (int)(pix / W) : 0, px = active ? (int)(pix - (long long)py * W) : 0;
  const RollSrc rs = roll_source(mode, circular, t, T, C);
  const size_t frame = (size_t)hw * C;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  const int cb = lane16 * 8;            // first channel of this lane's chunk inside the LN input
  const bool has = active && cb < cin;
  if (has) {
    if (!shift) {
      unpack8(__ldg(reinterpret_cast<const uint4 *>(x + (size_t)t * frame + (size_t)pix * C + cb)), v);
    } else if (cb < HC) {
      unpack8(__ldg(reinterpret_cast<const uint4 *>(x + rs.f_lo * frame + (size_t)pix * C + rs.c_lo + cb)), v);
    } else if (cb < C) {
      unpack8(__ldg(reinterpret_cast<const uint4 *>(x + rs.f_hi * frame + (size_t)pix * C + rs.c_hi + cb - HC)), v);
    } else if (hw_pre) {
      unpack8(__ldg(reinterpret_cast<const uint4 *>(hw_pre + ((size_t)t * hw + pix) * HC + cb - C)), v);
    } else {
      // shifted half: neighbour frame's channels, per-channel (dy,dx), zero fill, then dw3x3 with zero padding
      const bool fwd = mode == GSN_MODE_CAB2_FWD;
      const __half *src = x + (size_t)(fwd ? rs.f_lo : rs.f_hi) * frame + (fwd ? rs.c_lo : rs.c_hi);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = cb - C + i;
        int dy, dx;
        shift_offset_rt(C, c, dy, dx);
        float a = 0.f;
#pragma unroll
        for (int ty = -1; ty <= 1; ++ty)
#pragma unroll
          for (int tx = -1; tx <= 1; ++tx) {
            const int sy = py + ty, sx = px + tx;          // position in the shifted tensor
            const int qy = sy - dy, qx = sx - dx;          // its source position
            if (sy < 0 || sy >= H || sx < 0 || sx >= W || qy < 0 || qy >= H || qx < 0 || qx >= W) continue;
            a = fmaf(__half2float(__ldg(src + ((size_t)qy * W + qx) * C + c)), __half2float(__ldg(wc1 + ((ty + 1) * 3 + tx + 1) * HC + c)), a);
          }
        v[i] = a;
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s / cin;
  float ss = 0.f;
  if (cb < cin) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float e = v[i] - mu; ss = fmaf(e, e, ss); }
  }
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / cin + 1e-6f);
  if (active && cb < cinp) {
    float o8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o8[i] = (cb < cin) ? (v[i] - mu) * rstd * __ldg(ln + cb + i) + __ldg(ln + cin + cb + i) : 0.f;
    *reinterpret_cast<uint4 *>(out + ((size_t)t * hw + pix) * cinp + cb) = pack8(o8);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// ln_pw: [rolled stream | shift_conv1 output] -> LayerNorm -> first 1x1 (C' -> 2C) in one kernel; the two halves of the
// result (a | b) are written as separate (T,H,W,C) tensors for dw_gate.  128 pixels per CTA, k-chunk planar operands,
// mma.sync m16n8k16 with fp32 accumulation; the LayerNorm input never makes a round trip through HBM.
// ---------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) ln_pw_kernel(const __half *__restrict__ x, const __half *__restrict__ hw_pre, int T,
                                                    long long hw, int mode, int circular, const float *__restrict__ ln,
                                                    const __half *__restrict__ w1p /*[KCP][2C][8]*/,
                                                    __half *__restrict__ ga, __half *__restrict__ gb) {
  constexpr int MP = 128, PZ = (MP + 1) * 16, N = 2 * C, HC = C / 2;
  extern __shared__ __align__(128) unsigned char smem[];
  const bool shift = mode != GSN_MODE_CAB1;
  const int cin = shift ? C + HC : C, kc = cin / 8, kcp = (kc + 1) / 2 * 2;   // chunks, padded to whole k-steps
  unsigned char *sa = smem, *sw = smem + 16 * PZ;                           // A planes (<= 16), W planes [kcp][N]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * MP;
  const size_t frame = (size_t)hw * C;
  const RollSrc rs = roll_source(mode, circular, t, T, C);
  for (int i = tid; i < kcp * N; i += 256) cp_async16(sw + i * 16, reinterpret_cast<const unsigned char *>(w1p) + (size_t)i * 16, true);
  for (int i = tid; i < MP * kcp; i += 256) {
    const int ch = i % kcp, p = i / kcp;
    const bool valid = (p0 + p < hw) && ch < kc;
    const size_t pix = valid ? (size_t)(p0 + p) : 0;
    const __half *sp = x;
    if (valid) {
      const int cb = ch * 8;
      if (!shift) sp = x + (size_t)t * frame + pix * C + cb;
      else if (cb < HC) sp = x + rs.f_lo * frame + pix * C + rs.c_lo + cb;
      else if (cb < C) sp = x + rs.f_hi * frame + pix * C + rs.c_hi + cb - HC;
      else sp = hw_pre + ((size_t)t * hw + pix) * HC + cb - C;
    }
    cp_async16(sa + ch * PZ + p * 16, sp, valid);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  {  // LayerNorm: two threads per pixel, each owns every other chunk; biased variance, eps 1e-6 (d2:19-28)
    const int p = tid >> 1, hsel = tid & 1;
    float s = 0.f;
    for (int ch = hsel; ch < kc; ch += 2) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4 *>(sa + ch * PZ + p * 16), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[i];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    const float mu = s / cin;
    float ss = 0.f;
    for (int ch = hsel; ch < kc; ch += 2) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4 *>(sa + ch * PZ + p * 16), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float e = v[i] - mu; ss = fmaf(e, e, ss); }
    }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    const float rstd = rsqrtf(ss / cin + 1e-6f);
    for (int ch = hsel; ch < kc; ch += 2) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4 *>(sa + ch * PZ + p * 16), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (v[i] - mu) * rstd * __ldg(ln + ch * 8 + i) + __ldg(ln + cin + ch * 8 + i);
      *reinterpret_cast<uint4 *>(sa + ch * PZ + p * 16) = pack8(v);
    }
  }
  __syncthreads();
  constexpr int NTP = N / 16;   // n-tile pairs
  float acc[2 * NTP][4];
#pragma unroll
  for (int n = 0; n < 2 * NTP; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
  const uint32_t a_s = smem_u32(sa), w_s = smem_u32(sw);
  for (int k = 0; k < kcp / 2; ++k) {
    uint32_t a[4];
    ldmatrix_x4(a[0], a[1], a[2], a[3], a_s + (2 * k + (lane >> 4)) * PZ + (warp * 16 + (lane & 15)) * 16);
#pragma unroll
    for (int np = 0; np < NTP; ++np) {
      uint32_t b[4];
      ldmatrix_x4(b[0], b[1], b[2], b[3], w_s + ((2 * k + ((lane >> 3) & 1)) * N + np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * 16);
      mma16816(acc[2 * np], a, b[0], b[1]);
      mma16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
  __syncthreads();             // every warp is done with the weight planes: reuse them as the output staging tile
  const int g = lane >> 2, tig = lane & 3;
  unsigned char *so = sw;      // [2C/8 planes][MP+1] x 16 B  (= N/8 * PZ <= kcp * N * 16 for C in {64, 80})
#pragma unroll
  for (int n = 0; n < 2 * NTP; ++n)
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow)
      *reinterpret_cast<uint32_t *>(so + n * PZ + (warp * 16 + g + hrow * 8) * 16 + tig * 4) = pack_half2(acc[n][hrow * 2], acc[n][hrow * 2 + 1]);
  __syncthreads();
  constexpr int CHN = C / 8;
  for (int i = tid; i < MP * 2 * CHN; i += 256) {
    const int ch = i % (2 * CHN), p = i / (2 * CHN);
    if (p0 + p < hw) {
      __half *dst = (ch < CHN ? ga : gb) + ((size_t)t * hw + p0 + p) * C + (ch % CHN) * 8;
      *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(so + ch * PZ + p * 16);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dw_gate / gate2: 128 pixels x C/8 chunks per CTA, thread = (pixel slot, chunk); deterministic per-CTA channel sums
// ---------------------------------------------------------------------------------------------------------------
template <bool STENCIL>
__global__ void __launch_bounds__(320) gate_kernel(const __half *__restrict__ a, const __half *__restrict__ b, int H, int W,
                                                   int C, const __half *__restrict__ wd /*[9][2C] (STENCIL)*/,
                                                   __half *__restrict__ out, float *__restrict__ partial) {
  __shared__ float red[40 * 128];
  const int nch = C / 8, slots = 320 / nch;            // C = 80: 10 chunks x 32 slots ; C = 64: 8 x 40
  const int ch = threadIdx.x % nch, slot = threadIdx.x / nch;
  const long long hw = (long long)H * W, p0 = (long long)blockIdx.x * 128;
  const int t = blockIdx.y;
  const size_t fo = (size_t)t * hw * C;
  float sum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sum[i] = 0.f;
  if (slot < slots) {
    for (int k = slot; k < 128; k += slots) {
      const long long p = p0 + k;
      if (p >= hw) break;
      float va[8], vb[8];
      const size_t ce = fo + (size_t)p * C + ch * 8;
      unpack8(__ldg(reinterpret_cast<const uint4 *>(a + ce)), va);
      unpack8(__ldg(reinterpret_cast<const uint4 *>(b + ce)), vb);
      float o[8];
      if (STENCIL) {
        const int y = (int)(p / W), xq = (int)(p - (long long)y * W);
        float ra[8], rb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { ra[i] = va[i]; rb[i] = vb[i]; }     // identity of RepConv2
#pragma unroll
        for (int ty = -1; ty <= 1; ++ty)
#pragma unroll
          for (int tx = -1; tx <= 1; ++tx) {
            const int yy = y + ty, xx = xq + tx;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const size_t ne = fo + ((size_t)yy * W + xx) * C + ch * 8;
            float na[8], nb[8], wa[8], wb[8];
            unpack8(__ldg(reinterpret_cast<const uint4 *>(a + ne)), na);
            unpack8(__ldg(reinterpret_cast<const uint4 *>(b + ne)), nb);
            const int tap = (ty + 1) * 3 + tx + 1;
            unpack8(__ldg(reinterpret_cast<const uint4 *>(wd + (size_t)tap * 2 * C + ch * 8)), wa);
            unpack8(__ldg(reinterpret_cast<const uint4 *>(wd + (size_t)tap * 2 * C + C + ch * 8)), wb);
#pragma unroll
            for (int i = 0; i < 8; ++i) { ra[i] = fmaf(na[i], wa[i], ra[i]); rb[i] = fmaf(nb[i], wb[i], rb[i]); }
          }
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = ra[i] * rb[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = va[i] * sigmoidf_fast(vb[i]);
      }
      *reinterpret_cast<uint4 *>(out + ce) = pack8(o);
#pragma unroll
      for (int i = 0; i < 8; ++i) sum[i] += o[i];
    }
  }
  if (partial) {
    if (slot < slots) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[slot * 128 + ch * 8 + i] = sum[i];
    }
    __syncthreads();
    if ((int)threadIdx.x < C) {
      float s = 0.f;
      for (int q = 0; q < slots; ++q) s += red[q * 128 + threadIdx.x];
      partial[((size_t)t * gridDim.x + blockIdx.x) * C + threadIdx.x] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// group_conv5: u = (conv5x5 + conv3x3 + id)(s * g) with groups of 8 channels; s = optional per-frame channel scale of the
// denoise mid CALayer2, applied to the staged INPUT tile (it does not commute with a conv that mixes a group's channels).  mma.sync m16n8k16: M = 16 pixels of a row, N = the 8
// outputs of one group, K = (2 taps) x (8 inputs); 13 k-steps cover the 25 taps (+1 zero tap).
// ---------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256, 2) group_conv5_kernel(const __half *__restrict__ gin, int H, int W,
                                                             const uint2 *__restrict__ wfrag /*[C/8][13][32]*/,
                                                             const float *__restrict__ scale /*[T][C] or null*/,
                                                             __half *__restrict__ out) {
  constexpr int NG = C / 8, TW = 20, PITCH = C + 8;
  extern __shared__ __align__(16) unsigned char smem[];
  __half *tile = reinterpret_cast<__half *>(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = blockIdx.z, ox0 = blockIdx.x * 16, oy0 = blockIdx.y * 16;
  const size_t fo = (size_t)t * H * W * C;
  for (int p = tid; p < TW * TW; p += 256) {
    const int ly = p / TW, lx = p - ly * TW;
    const int gy = oy0 + ly - 2, gx = ox0 + lx - 2;
    const bool valid = gy >= 0 && gy < H && gx >= 0 && gx < W;
    const __half *sp = valid ? gin + fo + ((size_t)gy * W + gx) * C : gin;
#pragma unroll
    for (int ch = 0; ch < NG; ++ch) cp_async16(tile + (size_t)p * PITCH + ch * 8, sp + (valid ? ch * 8 : 0), valid);
  }
  cp_async_commit();
  cp_async_wait<0>();
  if (scale) {
    // denoise: RepConv(s * g) (gshift_denoise1.py:190-191,224-225) -- a grouped conv mixes the 8 channels of a group, so the
    // per-channel scale must sit on the INPUT side: each thread scales the pixels it staged itself (same p loop => its own
    // cp.async data, already complete for this thread); the identity term below then reads s*g as well
    const float4 *sc4 = reinterpret_cast<const float4 *>(scale + (size_t)t * C);
    for (int p = tid; p < TW * TW; p += 256) {
#pragma unroll
      for (int ch = 0; ch < NG; ++ch) {
        uint4 *q = reinterpret_cast<uint4 *>(tile + (size_t)p * PITCH + ch * 8);
        const float4 s0 = __ldg(sc4 + ch * 2), s1 = __ldg(sc4 + ch * 2 + 1);
        float f[8];
        unpack8(*q, f);
        f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
        f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
        *q = pack8(f);
      }
    }
  }
  __syncthreads();

  float acc[2][NG][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < NG; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[m][n][j] = 0.f;
  const uint32_t tile_s = smem_u32(tile);
  const int prow = (lane & 7) + ((lane >> 3) & 1) * 8;   // pixel (x) this lane addresses
  const int tsel = lane >> 4;                            // 0: first tap of the pair, 1: second tap
#pragma unroll 1
  for (int ks = 0; ks < 13; ++ks) {
    int tap = 2 * ks + tsel;
    if (tap > 24) tap = 24;                              // 26th tap: weights are zero, any valid address will do
    const int ky = tap / 5, kx = tap - ky * 5;
#pragma unroll
    for (int n = 0; n < NG; ++n) {
      uint32_t a0[4], a1[4];
      const uint32_t base = tile_s + (uint32_t)((((2 * warp + ky) * TW + prow + kx) * PITCH + n * 8) * 2);
      ldmatrix_x4(a0[0], a0[1], a0[2], a0[3], base);
      ldmatrix_x4(a1[0], a1[1], a1[2], a1[3], base + TW * PITCH * 2);
      const uint2 b = __ldg(wfrag + ((size_t)n * 13 + ks) * 32 + lane);
      mma16816(acc[0][n], a0, b.x, b.y);
      mma16816(acc[1][n], a1, b.x, b.y);
    }
  }
  const int g = lane >> 2, tig = lane & 3;
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const int oy = oy0 + 2 * warp + m;
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int lx = g + hrow * 8, ox = ox0 + lx;
      if (oy >= H || ox >= W) continue;
      __half *dp = out + fo + ((size_t)oy * W + ox) * C + tig * 2;
      const __half *cp = tile + ((2 * warp + m + 2) * TW + lx + 2) * PITCH + tig * 2;   // centre pixel: identity term
#pragma unroll
      for (int n = 0; n < NG; ++n) {
        const float2 idv = unpack_half2(*reinterpret_cast<const uint32_t *>(cp + n * 8));
        const float v0 = acc[m][n][hrow * 2] + idv.x, v1 = acc[m][n][hrow * 2 + 1] + idv.y;
        *reinterpret_cast<uint32_t *>(dp + n * 8) = pack_half2(v0, v1);
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// shift_conv1: out = conv1(spatial_shift2(neighbour half))   -- the gather folded into the load stage of the dw3x3.
// 16x16-pixel tiles, the 34x34x(C/2) source box staged with zero-filling cp.async, per-channel sliding windows,
// results staged in smem and written with 16-byte stores.  Memory/L2 bound; 2 CTAs per SM overlap load and compute.
// ---------------------------------------------------------------------------------------------------------------
// LN (C = 64): CAB2's LayerNorm over [rolled stream | conv1 output] (gshift_deblur2.py:250-254) runs in the same kernel and the
// result leaves in the k-chunk planar layout [T][12][H][W][8] of GsnCabPassA.a1_pre; `out` then holds that tensor and the conv1
// output never goes to HBM.
template <int C, bool TMA, bool LN = false>
__global__ void __launch_bounds__(256, (C <= 64 ? 2 : 1)) shift_conv1_kernel(const __half *__restrict__ x, int T, int H, int W, int mode,
                                                             int circular, const __half *__restrict__ wc1,
                                                             __half *__restrict__ out, const __grid_constant__ CUtensorMap tmap,
                                                             const float *__restrict__ ln = nullptr) {
  constexpr int HC = C / 2, CH = HC / 8, TS = 16, BW = TS + 18;
  extern __shared__ __align__(128) unsigned char smem[];
  __half *box = reinterpret_cast<__half *>(smem);                       // [BW*BW][HC]
  __half *ot = reinterpret_cast<__half *>(smem + BW * BW * HC * 2);     // [TS*TS][HC]
  const int tid = threadIdx.x;
  const int t = blockIdx.z, x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
  const RollSrc rs = roll_source(mode, circular, t, T, C);
  const bool fwd = mode == GSN_MODE_CAB2_FWD;
  // LN: a quad of lanes owns a pixel (4 pixels per thread); the rolled stream of those pixels is requested now and consumed
  // after the conv (its latency hides under the TMA wait and the conv loop)
  uint4 rolled[LN ? 4 : 1][2];
  if (LN) {
    const int j = tid & 3;
    const size_t frame = (size_t)H * W * C;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int p = (tid >> 2) + it * 64, gy = y0 + p / TS, gx = x0 + p % TS;
      rolled[it][0] = rolled[it][1] = make_uint4(0, 0, 0, 0);
      if (gy < H && gx < W) {
        const size_t pix = ((size_t)gy * W + gx) * C;
        rolled[it][0] = __ldg(reinterpret_cast<const uint4 *>(x + rs.f_lo * frame + pix + rs.c_lo + j * 8));
        rolled[it][1] = __ldg(reinterpret_cast<const uint4 *>(x + rs.f_hi * frame + pix + rs.c_hi + j * 8));
      }
    }
  }
  if (TMA) {
    // One TMA tile load brings the whole 34x34x(C/2) box (out-of-image elements are zero-filled by the hardware).
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t bar_s = smem_u32(&bar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_s));
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      const int c0 = fwd ? rs.c_lo : rs.c_hi, f = fwd ? rs.f_lo : rs.f_hi;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_s), "r"(BW * BW * HC * 2) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
              "r"(smem_u32(box)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(x0 - 9), "r"(y0 - 9), "r"(f), "r"(bar_s)
          : "memory");
    }
    uint32_t done = 0;
    for (int spin = 0; !done; ++spin) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done) : "r"(bar_s) : "memory");
      if (spin > (1 << 24)) __trap();     // a wrong tensor map must not hang the GPU
    }
  } else {
    const size_t frame = (size_t)H * W * C;
    const __half *src = x + (size_t)(fwd ? rs.f_lo : rs.f_hi) * frame + (fwd ? rs.c_lo : rs.c_hi);
    for (int p = tid; p < BW * BW; p += 256) {
      const int by = p / BW, bx = p - by * BW;
      const int gy = y0 - 9 + by, gx = x0 - 9 + bx;
      const bool valid = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const __half *sp = valid ? src + ((size_t)gy * W + gx) * C : src;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) cp_async16(box + (size_t)p * HC + ch * 8, sp + (valid ? ch * 8 : 0), valid);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  }
  const bool interior = y0 >= 1 && y0 + TS + 1 <= H && x0 >= 1 && x0 + TS + 1 <= W;   // all destination taps in-image
  // INT = true: tile whose conv taps all land inside the image (the vast majority) -- no border predicates in the tap loop
  auto conv_items = [&](auto int_tag) {
    constexpr bool INT = decltype(int_tag)::value;
    for (int item = tid; item < HC * TS; item += 256) {
      const int c = item % HC, oy = item / HC;
      int dy, dx;
      shift_offset<C>(c, dy, dx);
      float w[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) w[i] = __half2float(__ldg(wc1 + i * HC + c));
      const __half *rp = box + ((oy - 1 - dy + 9) * BW - dx + 9) * HC + c;   // (row oy+ty-1, col) -> rp[(ty*BW + col) * HC]
      bool rowok[3];
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) { const int sy = y0 + oy + ty - 1; rowok[ty] = INT || (sy >= 0 && sy < H); }
      float v[3][3];
      auto load_col = [&](int col, float(&o)[3]) {
        const int sx = x0 + col;
        const bool cok = INT || (sx >= 0 && sx < W);
#pragma unroll
        for (int ty = 0; ty < 3; ++ty) o[ty] = (INT || (cok && rowok[ty])) ? __half2float(rp[(ty * BW + col) * HC]) : 0.f;
      };
      {
        float a[3], b[3];
        load_col(-1, a);
        load_col(0, b);
#pragma unroll
        for (int ty = 0; ty < 3; ++ty) { v[ty][1] = a[ty]; v[ty][2] = b[ty]; }
      }
#pragma unroll
      for (int ox = 0; ox < TS; ++ox) {
        float nc[3];
        load_col(ox + 1, nc);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int ty = 0; ty < 3; ++ty) { v[ty][0] = v[ty][1]; v[ty][1] = v[ty][2]; v[ty][2] = nc[ty]; }
        a0 = fmaf(v[0][0], w[0], fmaf(v[0][1], w[1], v[0][2] * w[2]));
        a1 = fmaf(v[1][0], w[3], fmaf(v[1][1], w[4], v[1][2] * w[5]));
        a2 = fmaf(v[2][0], w[6], fmaf(v[2][1], w[7], v[2][2] * w[8]));
        ot[(oy * TS + ox) * HC + c] = __float2half_rn(a0 + a1 + a2);
      }
    }
  };
  if (interior) conv_items(std::true_type{});
  else conv_items(std::false_type{});
  __syncthreads();
  if (LN) {
    constexpr int CIN = C + HC, KC = CIN / 8;
    const int j = tid & 3;
    const size_t hw = (size_t)H * W;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int p = (tid >> 2) + it * 64, gy = y0 + p / TS, gx = x0 + p % TS;
      float v[24];
      unpack8(rolled[it][0], *reinterpret_cast<float(*)[8]>(&v[0]));
      unpack8(rolled[it][1], *reinterpret_cast<float(*)[8]>(&v[8]));
      unpack8(*reinterpret_cast<const uint4 *>(ot + (size_t)p * HC + j * 8), *reinterpret_cast<float(*)[8]>(&v[16]));
      float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 24; ++i) { s4[i & 3] += v[i]; q4[i & 3] = fmaf(v[i], v[i], q4[i & 3]); }
      float s = (s4[0] + s4[1]) + (s4[2] + s4[3]), ss = (q4[0] + q4[1]) + (q4[2] + q4[3]);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      ss += __shfl_xor_sync(0xffffffffu, ss, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      ss += __shfl_xor_sync(0xffffffffu, ss, 2);
      const float mu = s * (1.f / CIN);
      const float rstd = rsqrtf(fmaxf(ss * (1.f / CIN) - mu * mu, 0.f) + 1e-6f);
      const float nmr = -mu * rstd;
      if (gy < H && gx < W) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int ch = 4 * k + j;
          const float4 g0 = __ldg(reinterpret_cast<const float4 *>(ln + ch * 8)), g1 = __ldg(reinterpret_cast<const float4 *>(ln + ch * 8 + 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4 *>(ln + CIN + ch * 8)), b1 = __ldg(reinterpret_cast<const float4 *>(ln + CIN + ch * 8 + 4));
          const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bet[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = fmaf(fmaf(v[k * 8 + i], rstd, nmr), gam[i], bet[i]);
          *reinterpret_cast<uint4 *>(out + (((size_t)t * KC + ch) * hw + (size_t)gy * W + gx) * 8) = pack8(o);
        }
      }
    }
    return;
  }
  for (int i = tid; i < TS * TS * CH; i += 256) {
    const int ch = i % CH, p = i / CH;
    const int gy = y0 + p / TS, gx = x0 + (p % TS);
    if (gy < H && gx < W)
      *reinterpret_cast<uint4 *>(out + (((size_t)t * H + gy) * W + gx) * HC + ch * 8) = *reinterpret_cast<const uint4 *>(ot + (size_t)p * HC + ch * 8);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The same gather + conv1 + LayerNorm on 16x8-pixel tiles (C = 64): the 34x26 box + the 128-pixel conv tile are 65 KB and the
// kernel fits 85 registers, so THREE CTAs are resident per SM instead of two -- the one-tile-per-CTA kernel is latency bound
// (box wait at the head of every CTA, no pipe above 31 %), and occupancy is what hides that.  One conv item (channel, row) and
// two LayerNorm pixels (as a quad of lanes) per thread; same arithmetic as shift_conv1_kernel<64, true, true>.
__global__ void __launch_bounds__(256, 3) shift_conv1_ln_h8_kernel(const __half *__restrict__ x, int T, int H, int W, int mode, int circular,
                                                                   const __half *__restrict__ wc1, __half *__restrict__ out,
                                                                   const __grid_constant__ CUtensorMap tmap, const float *__restrict__ ln) {
  constexpr int C = 64, HC = 32, TSX = 16, TSY = 8, BW = TSX + 18, BH = TSY + 18, CIN = C + HC, KC = CIN / 8;
  extern __shared__ __align__(128) unsigned char smem[];
  __half *box = reinterpret_cast<__half *>(smem);                       // [BH*BW][HC]
  __half *ot = reinterpret_cast<__half *>(smem + BH * BW * HC * 2);     // [TSY*TSX][HC]
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x, j = tid & 3;
  const int t = blockIdx.z, x0 = blockIdx.x * TSX, y0 = blockIdx.y * TSY;
  const RollSrc rs = roll_source(mode, circular, t, T, C);
  const bool fwd = mode == GSN_MODE_CAB2_FWD;
  const uint32_t bar_s = smem_u32(&bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_s));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    const int c0 = fwd ? rs.c_lo : rs.c_hi, f = fwd ? rs.f_lo : rs.f_hi;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_s), "r"(BW * BH * HC * 2) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
            "r"(smem_u32(box)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(x0 - 9), "r"(y0 - 9), "r"(f), "r"(bar_s)
        : "memory");
  }
  // rolled stream of this thread's two LayerNorm pixels: requested now, consumed after the conv
  uint4 rolled[2][2];
  {
    const size_t frame = (size_t)H * W * C;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int p = (tid >> 2) + it * 64, gy = y0 + p / TSX, gx = x0 + p % TSX;
      rolled[it][0] = rolled[it][1] = make_uint4(0, 0, 0, 0);
      if (gy < H && gx < W) {
        const size_t pix = ((size_t)gy * W + gx) * C;
        rolled[it][0] = __ldg(reinterpret_cast<const uint4 *>(x + rs.f_lo * frame + pix + rs.c_lo + j * 8));
        rolled[it][1] = __ldg(reinterpret_cast<const uint4 *>(x + rs.f_hi * frame + pix + rs.c_hi + j * 8));
      }
    }
  }
  __syncthreads();          // barrier init visible before anyone polls it
  {
    uint32_t done = 0;
    for (int spin = 0; !done; ++spin) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done) : "r"(bar_s) : "memory");
      if (spin > (1 << 24)) __trap();     // a wrong tensor map must not hang the GPU
    }
  }
  const bool interior = y0 >= 1 && y0 + TSY + 1 <= H && x0 >= 1 && x0 + TSX + 1 <= W;   // all destination taps in-image
  auto conv_item = [&](auto int_tag) {
    constexpr bool INT = decltype(int_tag)::value;
    const int c = tid % HC, oy = tid / HC;       // HC * TSY == 256 items: one per thread
    int dy, dx;
    shift_offset<C>(c, dy, dx);
    float w[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) w[i] = __half2float(__ldg(wc1 + i * HC + c));
    const __half *rp = box + ((oy - 1 - dy + 9) * BW - dx + 9) * HC + c;   // (row oy+ty-1, col) -> rp[(ty*BW + col) * HC]
    bool rowok[3];
#pragma unroll
    for (int ty = 0; ty < 3; ++ty) { const int sy = y0 + oy + ty - 1; rowok[ty] = INT || (sy >= 0 && sy < H); }
    float v[3][3];
    auto load_col = [&](int col, float(&o)[3]) {
      const int sx = x0 + col;
      const bool cok = INT || (sx >= 0 && sx < W);
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) o[ty] = (INT || (cok && rowok[ty])) ? __half2float(rp[(ty * BW + col) * HC]) : 0.f;
    };
    {
      float a[3], b[3];
      load_col(-1, a);
      load_col(0, b);
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) { v[ty][1] = a[ty]; v[ty][2] = b[ty]; }
    }
#pragma unroll
    for (int ox = 0; ox < TSX; ++ox) {
      float nc[3];
      load_col(ox + 1, nc);
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) { v[ty][0] = v[ty][1]; v[ty][1] = v[ty][2]; v[ty][2] = nc[ty]; }
      const float a0 = fmaf(v[0][0], w[0], fmaf(v[0][1], w[1], v[0][2] * w[2]));
      const float a1 = fmaf(v[1][0], w[3], fmaf(v[1][1], w[4], v[1][2] * w[5]));
      const float a2 = fmaf(v[2][0], w[6], fmaf(v[2][1], w[7], v[2][2] * w[8]));
      ot[(oy * TSX + ox) * HC + c] = __float2half_rn(a0 + a1 + a2);
    }
  };
  if (interior) conv_item(std::true_type{});
  else conv_item(std::false_type{});
  __syncthreads();
  const size_t hw = (size_t)H * W;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int p = (tid >> 2) + it * 64, gy = y0 + p / TSX, gx = x0 + p % TSX;
    float v[24];
    unpack8(rolled[it][0], *reinterpret_cast<float(*)[8]>(&v[0]));
    unpack8(rolled[it][1], *reinterpret_cast<float(*)[8]>(&v[8]));
    unpack8(*reinterpret_cast<const uint4 *>(ot + (size_t)p * HC + j * 8), *reinterpret_cast<float(*)[8]>(&v[16]));
    float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 24; ++i) { s4[i & 3] += v[i]; q4[i & 3] = fmaf(v[i], v[i], q4[i & 3]); }
    float sm = (s4[0] + s4[1]) + (s4[2] + s4[3]), ss = (q4[0] + q4[1]) + (q4[2] + q4[3]);
    sm += __shfl_xor_sync(0xffffffffu, sm, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    sm += __shfl_xor_sync(0xffffffffu, sm, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    const float mu = sm * (1.f / CIN);
    const float rstd = rsqrtf(fmaxf(ss * (1.f / CIN) - mu * mu, 0.f) + 1e-6f);
    const float nmr = -mu * rstd;
    if (gy < H && gx < W) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int ch = 4 * k + j;
        const float4 g0 = __ldg(reinterpret_cast<const float4 *>(ln + ch * 8)), g1 = __ldg(reinterpret_cast<const float4 *>(ln + ch * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(ln + CIN + ch * 8)), b1 = __ldg(reinterpret_cast<const float4 *>(ln + CIN + ch * 8 + 4));
        const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bet[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(fmaf(v[k * 8 + i], rstd, nmr), gam[i], bet[i]);
        *reinterpret_cast<uint4 *>(out + (((size_t)t * KC + ch) * hw + (size_t)gy * W + gx) * 8) = pack8(o);
      }
    }
  }
}

// y = clamped temporal roll of x (Shift_CAB.channel_shift, gshift_denoise1.py:167-179); C real channels inside cp
__global__ void __launch_bounds__(256) roll_copy_kernel(const __half *__restrict__ x, __half *__restrict__ y, int T,
                                                        long long hw, int C, int cp, int reverse) {
  const long long total = (long long)T * hw * cp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cp);
    const long long p = (i / cp) % hw;
    const int t = (int)(i / cp / hw);
    float v = 0.f;
    if (c < C) {
      const RollSrc rs = roll_source(reverse ? GSN_MODE_CAB2_REV : GSN_MODE_CAB2_FWD, 0, t, T, C);
      const int h = C / 2;
      const int f = c < h ? rs.f_lo : rs.f_hi, sc = c < h ? rs.c_lo + c : rs.c_hi + c - h;
      v = __half2float(x[((size_t)f * hw + p) * cp + sc]);
    }
    y[i] = __float2half_rn(v);
  }
}

}  // namespace gsn

extern "C" int gsn_shift_conv1(const void *x, int T, int H, int W, int C, int mode, int circular, const void *wc1, void *out,
                               void *stream) {
  using namespace gsn;
  GSN_REQUIRE(x && wc1 && out, "shift_conv1: null pointer");
  GSN_REQUIRE(T > 0 && H > 0 && W > 0, "shift_conv1: empty shape");
  GSN_REQUIRE(mode == GSN_MODE_CAB2_FWD || mode == GSN_MODE_CAB2_REV, "shift_conv1: mode=%d is not a shift mode", mode);
  if (C != 64 && C != 80) { set_error("shift_conv1: C=%d unsupported (64, 80)", C); return GSN_E_UNSUPPORTED; }
  const int smem = (34 * 34 + 16 * 16) * (C / 2) * 2;
  GSN_ONCE_PER_DEVICE(
    cudaFuncSetAttribute(shift_conv1_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (34 * 34 + 16 * 16) * 32 * 2);
    cudaFuncSetAttribute(shift_conv1_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (34 * 34 + 16 * 16) * 32 * 2);
    cudaFuncSetAttribute(shift_conv1_kernel<80, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (34 * 34 + 16 * 16) * 40 * 2);
    cudaFuncSetAttribute(shift_conv1_kernel<80, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (34 * 34 + 16 * 16) * 40 * 2));
  dim3 grid((W + 15) / 16, (H + 15) / 16, T);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // TMA path: a 4-D tensor map (C, W, H, T) over x with a (C/2, 34, 34, 1) box; GSN_SHIFT_TMA=0 selects the cp.async loader
  static const bool use_tma = [] { const char *e = getenv("GSN_SHIFT_TMA"); return !(e && e[0] == '0'); }();
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  bool tma_ok = false;
  if (use_tma) tma_ok = encode_tmap_nhwc(&tm, x, C, W, H, roll_frames(circular, T), C / 2, 34, 34);
  const __half *xh = reinterpret_cast<const __half *>(x), *wh = reinterpret_cast<const __half *>(wc1);
  __half *oh = reinterpret_cast<__half *>(out);
  if (C == 64) {
    if (tma_ok) shift_conv1_kernel<64, true><<<grid, 256, smem, st>>>(xh, T, H, W, mode, circular, wh, oh, tm);
    else shift_conv1_kernel<64, false><<<grid, 256, smem, st>>>(xh, T, H, W, mode, circular, wh, oh, tm);
  } else {
    if (tma_ok) shift_conv1_kernel<80, true><<<grid, 256, smem, st>>>(xh, T, H, W, mode, circular, wh, oh, tm);
    else shift_conv1_kernel<80, false><<<grid, 256, smem, st>>>(xh, T, H, W, mode, circular, wh, oh, tm);
  }
  count_launch();
  return check_launch("shift_conv1");
}

extern "C" int gsn_shift_conv1_ln(const void *x, int T, int H, int W, int C, int mode, int circular, const void *wc1, const float *ln,
                                  void *a1, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(x && wc1 && ln && a1, "shift_conv1_ln: null pointer");
  GSN_REQUIRE(T > 0 && H > 0 && W > 0, "shift_conv1_ln: empty shape");
  GSN_REQUIRE(mode == GSN_MODE_CAB2_FWD || mode == GSN_MODE_CAB2_REV, "shift_conv1_ln: mode=%d is not a shift mode", mode);
  if (C != 64) { set_error("shift_conv1_ln: C=%d unsupported (64)", C); return GSN_E_UNSUPPORTED; }
  constexpr int smem = (34 * 34 + 16 * 16) * 32 * 2;
  GSN_ONCE_PER_DEVICE(cudaFuncSetAttribute(shift_conv1_kernel<64, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  if (!encode_tmap_nhwc(&tm, x, C, W, H, roll_frames(circular, T), C / 2, 34, 34)) {
    set_error("shift_conv1_ln: cuTensorMapEncodeTiled failed (W=%d H=%d T=%d)", W, H, T);
    return GSN_E_CUDA;
  }
  static const bool h8 = [] { const char *e = getenv("GSN_SHIFT_H8"); return !(e && e[0] == '0'); }();
  if (h8) {       // 16x8 tiles, three CTAs per SM
    constexpr int smem8 = (34 * 26 + 16 * 8) * 32 * 2;
    GSN_ONCE_PER_DEVICE(cudaFuncSetAttribute(shift_conv1_ln_h8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem8));
    CUtensorMap tm8;
    memset(&tm8, 0, sizeof(tm8));
    if (!encode_tmap_nhwc(&tm8, x, C, W, H, roll_frames(circular, T), C / 2, 34, 26)) {
      set_error("shift_conv1_ln: cuTensorMapEncodeTiled failed (W=%d H=%d T=%d)", W, H, T);
      return GSN_E_CUDA;
    }
    dim3 grid8((W + 15) / 16, (H + 7) / 8, T);
    shift_conv1_ln_h8_kernel<<<grid8, 256, smem8, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __half *>(x), T, H, W, mode, circular, reinterpret_cast<const __half *>(wc1), reinterpret_cast<__half *>(a1), tm8, ln);
    count_launch();
    return check_launch("shift_conv1_ln");
  }
  dim3 grid((W + 15) / 16, (H + 15) / 16, T);
  shift_conv1_kernel<64, true, true><<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half *>(x), T, H, W, mode, circular, reinterpret_cast<const __half *>(wc1), reinterpret_cast<__half *>(a1), tm, ln);
  count_launch();
  return check_launch("shift_conv1_ln");
}

extern "C" int gsn_shift_ln(const void *x, int T, int H, int W, int C, int mode, int circular, const void *wc1, const float *ln,
                            void *out, int cinp, const void *hw_pre, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(x && ln && out, "shift_ln: null pointer");
  GSN_REQUIRE(C % 16 == 0 && C <= 80 && cinp % 8 == 0 && cinp <= 128 && T > 0 && H > 0 && W > 0, "shift_ln: bad sizes C=%d cinp=%d", C, cinp);
  GSN_REQUIRE(mode == GSN_MODE_CAB1 || wc1 || hw_pre, "shift_ln: conv1 weights missing");
  const long long hw = (long long)H * W;
  dim3 grid((unsigned)((hw + 15) / 16), T);
  shift_ln_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half *>(x), T, H, W, C, mode, circular, reinterpret_cast<const __half *>(wc1), ln,
      reinterpret_cast<__half *>(out), cinp, reinterpret_cast<const __half *>(hw_pre));
  count_launch();
  return check_launch("shift_ln");
}

extern "C" int gsn_ln_pw(const void *x, const void *hw_pre, int T, int H, int W, int C, int mode, int circular, const float *ln,
                         const void *w1p, void *ga, void *gb, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(x && ln && w1p && ga && gb, "ln_pw: null pointer");
  GSN_REQUIRE(mode == GSN_MODE_CAB1 || hw_pre, "ln_pw: CAB2 modes need the shift_conv1 output");
  GSN_REQUIRE(T > 0 && H > 0 && W > 0, "ln_pw: empty shape");
  if (C != 80) { set_error("ln_pw: C=%d unsupported (80)", C); return GSN_E_UNSUPPORTED; }
  const long long hw = (long long)H * W;
  const int kc = (mode == GSN_MODE_CAB1 ? C : C + C / 2) / 8, kcp = (kc + 1) / 2 * 2;
  const int smem = 16 * 129 * 16 + std::max(kcp * 2 * C * 16, 2 * C / 8 * 129 * 16);
  GSN_ONCE_PER_DEVICE(cudaFuncSetAttribute(ln_pw_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 129 * 16 + 20 * 129 * 16));
  dim3 grid((unsigned)((hw + 127) / 128), T);
  ln_pw_kernel<80><<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half *>(x), reinterpret_cast<const __half *>(hw_pre), T, hw, mode, circular, ln,
      reinterpret_cast<const __half *>(w1p), reinterpret_cast<__half *>(ga), reinterpret_cast<__half *>(gb));
  count_launch();
  return check_launch("ln_pw");
}

extern "C" int gsn_dw_gate(const void *a, const void *b, int T, int H, int W, int C, const void *wd, void *out, float *partial,
                           void *stream) {
  using namespace gsn;
  GSN_REQUIRE(a && b && wd && out, "dw_gate: null pointer");
  GSN_REQUIRE((C == 64 || C == 80) && T > 0 && H > 0 && W > 0, "dw_gate: C=%d unsupported", C);
  const long long hw = (long long)H * W;
  dim3 grid((unsigned)((hw + 127) / 128), T);
  gate_kernel<true><<<grid, 320, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half *>(a), reinterpret_cast<const __half *>(b), H, W, C, reinterpret_cast<const __half *>(wd),
      reinterpret_cast<__half *>(out), partial);
  count_launch();
  return check_launch("dw_gate");
}

extern "C" int gsn_gate2(const void *a, const void *b, int T, int H, int W, int C, void *out, float *partial, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(a && b && out && partial, "gate2: null pointer");
  GSN_REQUIRE((C == 64 || C == 80) && T > 0 && H > 0 && W > 0, "gate2: C=%d unsupported", C);
  const long long hw = (long long)H * W;
  dim3 grid((unsigned)((hw + 127) / 128), T);
  gate_kernel<false><<<grid, 320, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half *>(a), reinterpret_cast<const __half *>(b), H, W, C, nullptr, reinterpret_cast<__half *>(out),
      partial);
  count_launch();
  return check_launch("gate2");
}

extern "C" int gsn_group_conv5(const void *g, int T, int H, int W, int C, const void *wfrag, const float *scale, void *out,
                               void *stream) {
  using namespace gsn;
  GSN_REQUIRE(g && wfrag && out, "group_conv5: null pointer");
  GSN_REQUIRE(T > 0 && H > 0 && W > 0, "group_conv5: empty shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((W + 15) / 16, (H + 15) / 16, T);
  if (C == 80) {
    const size_t smem = 20 * 20 * (80 + 8) * 2;
    GSN_ONCE_PER_DEVICE(cudaFuncSetAttribute(group_conv5_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    group_conv5_kernel<80><<<grid, 256, smem, st>>>(reinterpret_cast<const __half *>(g), H, W,
                                                    reinterpret_cast<const uint2 *>(wfrag), scale, reinterpret_cast<__half *>(out));
  } else {
    set_error("group_conv5: C=%d unsupported (80)", C);
    return GSN_E_UNSUPPORTED;
  }
  count_launch();
  return check_launch("group_conv5");
}

extern "C" int gsn_roll_copy(const void *x, void *y, int T, int H, int W, int C, int cp, int reverse, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(x && y && T > 0 && H > 0 && W > 0 && C % 2 == 0 && C <= cp, "roll_copy: bad arguments");
  const long long hw = (long long)H * W, total = (long long)T * hw * cp;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  roll_copy_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half *>(x), reinterpret_cast<__half *>(y), T, hw, C, cp, reverse);
  count_launch();
  return check_launch("roll_copy");
}
