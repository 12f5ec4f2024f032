// Dense convolution as an implicit GEMM on tensor cores (mma.sync m16n8k16, fp16 in / fp32 acc).
//
// Covers every groups==1 nn.Conv2d of the reference path (see include/shiftnet_b200.h for the list of
// call sites).  NHWC fp16 activations; one CTA = 16x16 output pixels x all output channels; each of the
// 8 warps owns two output rows (two 16-pixel M tiles) so that every B fragment is used twice.
// The input tile (with halo, up to 3 channel-concatenated sources) is staged in shared memory with
// 16-byte cp.async (zero fill outside the image == the conv's zero padding); A fragments come from it
// via ldmatrix with a per-tap pixel offset, B fragments stream from a pre-packed global buffer that
// stays L1/L2 resident.
//
// HBM-bound by design for the narrow full-resolution layers (16->16 ch: 72 FLOP/B).
#include "common.cuh"

namespace gsn {

// CINP > 0: padded input width known at compile time (curated list of the layer shapes the nets use) -> the tap / k-step
// loops unroll completely and every shared-memory address is an immediate.  CINP == 0: generic runtime fallback.
template <int NT, int KS, int STRIDE, int CINP>
__global__ void __launch_bounds__(256, (NT <= 4 ? 4 : 2)) conv_mma_kernel(const GsnConvDesc d) {
  constexpr int IH = 15 * STRIDE + KS, IW = IH;   // compile-time tile geometry: no runtime integer divisions below
  const int cin_p = CINP ? CINP : d.cin_p;
  const int pitch = cin_p + 8;
  extern __shared__ __align__(16) unsigned char smem[];
  __half *tile = reinterpret_cast<__half *>(smem);
  float *red = reinterpret_cast<float *>(smem + (size_t)IH * IW * pitch * 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = blockIdx.z;
  const int ox0 = blockIdx.x * 16, oy0 = blockIdx.y * 16;
  const int ix0 = ox0 * STRIDE - d.pad, iy0 = oy0 * STRIDE - d.pad;

  // ---- stage the input tile ---------------------------------------------------------------------
  {
    const int chunks = cin_p >> 3;
    if (d.n_src == 1) {
      const __half *base = reinterpret_cast<const __half *>(d.src[0]);
      const int stored = d.src_c[0], schunks = stored >> 3;   // stored width may be 8 short of the K slot (e.g. 24 of 32)
      for (int px = tid; px < IH * IW; px += 256) {
        const int ly = px / IW, lx = px - ly * IW;
        const int gy = iy0 + ly, gx = ix0 + lx;
        const bool valid = (gy >= 0) && (gy < d.Hin) && (gx >= 0) && (gx < d.Win);
        const __half *sp = valid ? base + (((size_t)t * d.Hin + gy) * d.Win + gx) * stored : base;
        __half *dp = tile + (size_t)px * pitch;
#pragma unroll
        for (int ch = 0; ch < (CINP ? CINP / 8 : 1); ++ch) {
          if (CINP) { const bool v = valid && ch < schunks; cp_async16(dp + ch * 8, sp + (v ? ch * 8 : 0), v); }
        }
        if (!CINP)
          for (int ch = 0; ch < chunks; ++ch) { const bool v = valid && ch < schunks; cp_async16(dp + ch * 8, sp + (v ? ch * 8 : 0), v); }
      }
    } else {
      const int c1 = d.src_c[0], c2 = d.src_c[0] + d.src_c[1];
      for (int px = tid; px < IH * IW; px += 256) {
        const int ly = px / IW, lx = px - ly * IW;
        const int gy = iy0 + ly, gx = ix0 + lx;
        const bool valid = (gy >= 0) && (gy < d.Hin) && (gx >= 0) && (gx < d.Win);
        const size_t gpix = valid ? ((size_t)t * d.Hin + gy) * d.Win + gx : 0;
        __half *dp = tile + (size_t)px * pitch;
        const int stored = c2 + (d.n_src > 2 ? d.src_c[2] : 0);
        for (int ch = 0; ch < chunks; ++ch) {
          const int c = ch * 8;
          int s = 0, cb = 0;
          if (c >= c2 && d.n_src > 2) { s = 2; cb = c2; }
          else if (c >= c1) { s = 1; cb = c1; }
          const __half *base = reinterpret_cast<const __half *>(d.src[s]);
          const bool v = valid && c < stored;
          cp_async16(dp + c, v ? base + gpix * d.src_c[s] + (c - cb) : base, v);
        }
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  }

  // ---- main loop --------------------------------------------------------------------------------
  float acc[2][NT][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[m][n][j] = 0.f;

  const int arow = (lane & 7) + ((lane >> 3) & 1) * 8;  // output x within the M tile
  const int akof = (lane >> 4) * 8;                     // k half (0 / 8) this lane addresses
  const int ksteps = cin_p >> 4;
  const uint2 *wp = reinterpret_cast<const uint2 *>(d.wpack) + lane;
  // per-lane base address of the two M tiles (output rows 2*warp, 2*warp+1); taps / k-steps add constants
  const uint32_t a_base0 = smem_u32(tile) + (uint32_t)((((2 * warp) * STRIDE * IW + arow * STRIDE) * pitch + akof) * 2);
  const uint32_t a_row = (uint32_t)(STRIDE * IW * pitch * 2);
  constexpr int taps = KS * KS;

  if (CINP) {
    constexpr int KSTEPS = CINP ? CINP / 16 : 1;
    constexpr int PITCH = CINP + 8;
    // narrow layers: unroll a whole kernel row (addresses become immediates); wide layers: keep the loops rolled so the
    // compiler does not hoist dozens of B-fragment loads into registers (spills)
#pragma unroll 1
    for (int ky = 0; ky < KS; ++ky) {
      const uint32_t rowoff = (uint32_t)(ky * IW * PITCH * 2);
#pragma unroll(CINP <= 32 ? KS : 1)
      for (int kx = 0; kx < KS; ++kx) {
#pragma unroll(CINP <= 32 ? KSTEPS : 1)
        for (int k = 0; k < KSTEPS; ++k) {
          const uint32_t off = rowoff + (uint32_t)((kx * PITCH + k * 16) * 2);
          uint32_t a0[4], a1[4];
          ldmatrix_x4(a0[0], a0[1], a0[2], a0[3], a_base0 + off);
          ldmatrix_x4(a1[0], a1[1], a1[2], a1[3], a_base0 + a_row + off);
          const uint2 *wk = wp + (size_t)(((ky * KS + kx) * KSTEPS + k) * NT) * 32;
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            const uint2 b = __ldg(wk + n * 32);
            mma16816(acc[0][n], a0, b.x, b.y);
            mma16816(acc[1][n], a1, b.x, b.y);
          }
        }
      }
    }
  } else {
#pragma unroll 1
    for (int tap = 0; tap < taps; ++tap) {
      const int ky = tap / KS, kx = tap - ky * KS;
      for (int k = 0; k < ksteps; ++k) {
        const uint32_t off = (uint32_t)(((ky * IW + kx) * pitch + k * 16) * 2);
        uint32_t a0[4], a1[4];
        ldmatrix_x4(a0[0], a0[1], a0[2], a0[3], a_base0 + off);
        ldmatrix_x4(a1[0], a1[1], a1[2], a1[3], a_base0 + a_row + off);
        const uint2 *wk = wp + (size_t)((tap * ksteps + k) * NT) * 32;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const uint2 b = __ldg(wk + n * 32);
          mma16816(acc[0][n], a0, b.x, b.y);
          mma16816(acc[1][n], a1, b.x, b.y);
        }
      }
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------------
  const int g = lane >> 2, tig = lane & 3;
  float csum[NT][2], bia[NT][2];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    csum[n][0] = csum[n][1] = 0.f;
    bia[n][0] = d.bias ? __ldg(d.bias + n * 8 + tig * 2) : 0.f;
    bia[n][1] = d.bias ? __ldg(d.bias + n * 8 + tig * 2 + 1) : 0.f;
  }
  __half *dst = reinterpret_cast<__half *>(d.dst);
  const __half *res = reinterpret_cast<const __half *>(d.residual);
  const int cout_p = NT * 8;
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const int oy = oy0 + 2 * warp + m;
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int ox = ox0 + g + hrow * 8;
      if (oy >= d.Hout || ox >= d.Wout) continue;
      const size_t opix = ((size_t)t * d.Hout + oy) * d.Wout + ox;
      __half *dp = dst + opix * cout_p + tig * 2;
      const __half *rp = res ? res + opix * cout_p + tig * 2 : nullptr;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        float v0 = acc[m][n][hrow * 2 + 0] + bia[n][0], v1 = acc[m][n][hrow * 2 + 1] + bia[n][1];
        if (d.has_prelu) {
          v0 = v0 > 0.f ? v0 : v0 * d.prelu_slope;
          v1 = v1 > 0.f ? v1 : v1 * d.prelu_slope;
        }
        if (rp) {
          const float2 r = unpack_half2(*reinterpret_cast<const uint32_t *>(rp + n * 8));
          v0 += r.x; v1 += r.y;
        }
        csum[n][0] += v0; csum[n][1] += v1;
        if (!d.pixel_shuffle) {
          *reinterpret_cast<uint32_t *>(dp + n * 8) = pack_half2(v0, v1);
        } else {
          // F.pixel_shuffle(.,2): conv channel co = c*4 + i*2 + j -> dst[c, 2y+i, 2x+j]; co is even => j = 0 / 1
          const int co = n * 8 + tig * 2;
          const int cd = d.dst_c ? d.dst_c : (cout_p >> 2), c = co >> 2, i = (co >> 1) & 1;
          const size_t sp = (((size_t)t * 2 * d.Hout + 2 * oy + i) * (2 * d.Wout) + 2 * ox) * cd + c;
          dst[sp] = __float2half_rn(v0);
          dst[sp + cd] = __float2half_rn(v1);
        }
      }
    }
  }

  if (d.chan_partial) {
    // deterministic per-tile channel sums: lanes (same tig) -> warp -> CTA
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v = csum[n][j];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0) red[warp * cout_p + n * 8 + tig * 2 + j] = v;
      }
    __syncthreads();
    if (tid < cout_p) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w * cout_p + tid];
      const size_t tile_id = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
      d.chan_partial[((size_t)t * gridDim.x * gridDim.y + tile_id) * cout_p + tid] = s;
    }
  }
}

template <int NT, int KS, int STRIDE, int CINP>
static int launch_conv(const GsnConvDesc &d, cudaStream_t st) {
  constexpr int IH = 15 * STRIDE + KS, IW = IH;
  const int pitch = d.cin_p + 8;
  const size_t smem = (size_t)IH * IW * pitch * 2 + 8 * d.cout_p * sizeof(float);
  if (smem > 227 * 1024) { set_error("conv_mma: tile needs %zu B smem", smem); return GSN_E_UNSUPPORTED; }
  GSN_ONCE_PER_DEVICE(
    cudaFuncSetAttribute(conv_mma_kernel<NT, KS, STRIDE, CINP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  dim3 grid((d.Wout + 15) / 16, (d.Hout + 15) / 16, d.T);
  conv_mma_kernel<NT, KS, STRIDE, CINP><<<grid, 256, smem, st>>>(d);
  count_launch();
  return check_launch("conv_mma");
}

template <int NT>
static int launch_conv_nt(const GsnConvDesc &d, cudaStream_t st) {
  const int key = d.cin_p * 100 + d.ks * 10 + d.stride;
  // specialised (fully unrolled) instances for the layer shapes of the four nets
#define GSN_CONV_CASE(CIN, K, S) if (key == CIN * 100 + K * 10 + S) return launch_conv<NT, K, S, CIN>(d, st);
  if constexpr (NT == 2) { GSN_CONV_CASE(16, 3, 1) GSN_CONV_CASE(32, 1, 1) GSN_CONV_CASE(48, 3, 1) GSN_CONV_CASE(32, 3, 1) }
  if constexpr (NT == 3) { GSN_CONV_CASE(32, 3, 1) GSN_CONV_CASE(16, 3, 2) GSN_CONV_CASE(32, 3, 2) GSN_CONV_CASE(32, 1, 1) GSN_CONV_CASE(48, 1, 1) GSN_CONV_CASE(48, 3, 1) GSN_CONV_CASE(96, 3, 1) }
  if constexpr (NT == 5) { GSN_CONV_CASE(48, 3, 1) GSN_CONV_CASE(32, 3, 2) GSN_CONV_CASE(48, 1, 1) }
  if constexpr (NT == 4) { GSN_CONV_CASE(32, 3, 1) GSN_CONV_CASE(16, 3, 2) GSN_CONV_CASE(32, 3, 2) GSN_CONV_CASE(32, 1, 1) GSN_CONV_CASE(48, 1, 1) GSN_CONV_CASE(96, 3, 1) GSN_CONV_CASE(64, 3, 1) }
  if constexpr (NT == 6) { GSN_CONV_CASE(48, 3, 1) GSN_CONV_CASE(32, 3, 2) GSN_CONV_CASE(48, 3, 2) GSN_CONV_CASE(48, 1, 1) }
  if constexpr (NT == 8) { GSN_CONV_CASE(64, 3, 1) GSN_CONV_CASE(64, 3, 2) GSN_CONV_CASE(64, 1, 1) GSN_CONV_CASE(16, 2, 2) }
  if constexpr (NT == 12) { GSN_CONV_CASE(80, 3, 1) }
  if constexpr (NT == 10) { GSN_CONV_CASE(80, 3, 1) GSN_CONV_CASE(80, 3, 2) GSN_CONV_CASE(80, 1, 1) GSN_CONV_CASE(32, 2, 2) }
#undef GSN_CONV_CASE
  if (d.ks == 3 && d.stride == 1) return launch_conv<NT, 3, 1, 0>(d, st);
  if (d.ks == 3 && d.stride == 2) return launch_conv<NT, 3, 2, 0>(d, st);
  if (d.ks == 1 && d.stride == 1) return launch_conv<NT, 1, 1, 0>(d, st);
  if (d.ks == 2 && d.stride == 2) return launch_conv<NT, 2, 2, 0>(d, st);
  set_error("conv_mma: ks=%d stride=%d unsupported (3/1, 3/2, 1/1, 2/2)", d.ks, d.stride);
  return GSN_E_UNSUPPORTED;
}

}  // namespace gsn

extern "C" int gsn_conv_tiles(int Hout, int Wout) { return ((Hout + 15) / 16) * ((Wout + 15) / 16); }

extern "C" int gsn_conv_mma(const GsnConvDesc *dp, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(dp != nullptr, "conv_mma: null descriptor");
  const GsnConvDesc &d = *dp;
  GSN_REQUIRE(d.n_src >= 1 && d.n_src <= 3, "conv_mma: n_src=%d", d.n_src);
  int csum = 0;
  for (int i = 0; i < d.n_src; ++i) {
    GSN_REQUIRE(d.src[i] && d.src_c[i] > 0 && d.src_c[i] % 8 == 0, "conv_mma: bad source %d (c=%d)", i, d.src_c[i]);
    csum += d.src_c[i];
  }
  GSN_REQUIRE(d.cin_p == (csum + 15) / 16 * 16, "conv_mma: cin_p=%d must be the sum of the sources (%d) rounded up to 16", d.cin_p, csum);
  GSN_REQUIRE(d.ks >= 1 && d.ks <= 3 && d.stride >= 1 && d.stride <= 2, "conv_mma: ks=%d stride=%d unsupported", d.ks, d.stride);
  GSN_REQUIRE(d.T > 0 && d.Hin > 0 && d.Win > 0 && d.Hout > 0 && d.Wout > 0, "conv_mma: empty shape");
  GSN_REQUIRE((d.Hin + 2 * d.pad - d.ks) / d.stride + 1 == d.Hout && (d.Win + 2 * d.pad - d.ks) / d.stride + 1 == d.Wout,
              "conv_mma: output size %dx%d inconsistent with input %dx%d k%d s%d p%d", d.Hout, d.Wout, d.Hin, d.Win, d.ks,
              d.stride, d.pad);
  GSN_REQUIRE(d.wpack && d.dst, "conv_mma: null weights/dst");
  GSN_REQUIRE(!d.pixel_shuffle || (d.cout_p % 4 == 0 && !d.residual && !d.chan_partial), "conv_mma: bad pixel_shuffle combination");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (d.cout_p) {
    case 16: return launch_conv_nt<2>(d, st);
    case 24: return launch_conv_nt<3>(d, st);
    case 40: return launch_conv_nt<5>(d, st);
    case 32: return launch_conv_nt<4>(d, st);
    case 48: return launch_conv_nt<6>(d, st);
    case 64: return launch_conv_nt<8>(d, st);
    case 80: return launch_conv_nt<10>(d, st);
    case 96: return launch_conv_nt<12>(d, st);
    default: set_error("conv_mma: cout_p=%d unsupported (16/24/32/40/48/64/80/96)", d.cout_p); return GSN_E_UNSUPPORTED;
  }
}
