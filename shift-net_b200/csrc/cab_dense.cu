// Body of the dense channel-attention block (CAB, gshift_deblur2.py:143-158): conv3x3 -> PReLU -> conv3x3 in ONE kernel,
// plus the per-tile channel sums the CALayer pooling needs.  The intermediate tensor never leaves shared memory.
//
//   * input tile (TH+4 x 34 pixels, all CP channels) arrives by ONE TMA tile load (hardware zero fill = conv zero padding);
//   * both convs are implicit GEMMs on mma.sync (m16n8k16 + an m16n8k8 tail for CP = 24): M = 16 pixels of an image row,
//     N = 8 output channels, K = input channels of one tap.  A warp owns a 16-pixel-wide column block and R output rows;
//     an A fragment loaded for input row i and horizontal tap dx feeds the three vertical taps (output rows i, i-1, i-2), so
//     shared memory is read 3(R+2)/R times per output row instead of 9;
//   * B fragments (weights) live in shared memory in fragment order and are held in registers for one dx at a time;
//   * conv1's epilogue applies bias/PReLU, forces pixels outside the image to zero (conv2's zero padding) and writes the
//     (TH+2) x 32 intermediate tile to shared memory at a conflict-free 48-byte pixel pitch;
//   * conv2's epilogue stages the output tile in shared memory (over the dead input tile) and ONE TMA tile store writes it
//     (the hardware clips the part outside the image); channel sums are reduced lanes -> warp -> CTA in a fixed order.
//
// HBM traffic per pixel: CP*2 bytes in (+ halo re-reads from L2) and CP*2 bytes out, instead of 4x that for two conv launches.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace gsn {

template <int CP_, int R_, int TH_>
struct DenseCfg {
  static constexpr int CP = CP_, R = R_, TH = TH_;
  static constexpr int TW = 30;                 // output tile width
  static constexpr int IW = TW + 4, IH = TH + 4;   // input tile
  static constexpr int MW = 32, MH = TH + 2;    // intermediate tile computed (2 column blocks of 16)
  static constexpr int MPW = 34;                // intermediate row pitch in pixels (conv2 reads 2 columns past MW for discarded outputs)
  static constexpr int NT = CP / 8, K16 = CP / 16, K8 = (CP % 16) / 8, WORDS = 2 * K16 + K8;
  static constexpr int INP = CP * 2;            // input pixel pitch (dense TMA box)
  static constexpr int MIDP = 48;               // intermediate pixel pitch: odd multiple of 16 B -> conflict-free ldmatrix
  static_assert(CP * 2 <= MIDP, "intermediate pitch");
  static constexpr int IN_BYTES = IH * IW * INP;
  static constexpr int OUT_BYTES = TH * TW * CP * 2;
  static_assert(OUT_BYTES <= IN_BYTES, "output staging aliases the input tile");
  static constexpr int MID_OFF = (IN_BYTES + 127) / 128 * 128;
  static constexpr int MID_BYTES = (MH * MPW + 2) * MIDP;
  static constexpr int WFRAG_BYTES = 9 * NT * WORDS * 32 * 4;   // one conv
  static constexpr int W_OFF = (MID_OFF + MID_BYTES + 127) / 128 * 128;
  static constexpr int BIAS_OFF = W_OFF + 2 * WFRAG_BYTES;      // bias1[CP], bias2[CP] fp32
  static constexpr int RED_OFF = BIAS_OFF + 2 * CP * 4;         // 8 warps x CP fp32
  static constexpr int BAR_OFF = (RED_OFF + 8 * CP * 4 + 15) / 16 * 16;
  static constexpr int SMEM = BAR_OFF + 16;
  static constexpr int ROWBLK1 = (MH + R - 1) / R, ROWBLK2 = (TH + R - 1) / R;
  static_assert(ROWBLK1 * 2 <= 8 && ROWBLK2 * 2 <= 8, "one warp task per conv");
};

__device__ __forceinline__ void ldmatrix_x2(uint32_t &r0, uint32_t &r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
}
// D(16x8,f32) += A(16x8,f16,row) * B(8x8,f16,col)
__device__ __forceinline__ void mma1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}

// One warp: 3x3 conv of R output rows x 16 pixels x CP channels.  a_base = shared address of input pixel (row0, col0 + the lane's
// ldmatrix row) + the lane's k-half offset; output row o reads input rows o..o+2, output pixel m reads input pixels m..m+2.
template <int CP, int R, int PXP>
__device__ __forceinline__ void conv3x3_rows(float (&acc)[R][CP / 8][4], uint32_t a_base, uint32_t row_pitch, const uint32_t *wfrag,
                                             int lane, int rows_in) {
  constexpr int NT = CP / 8, K16 = CP / 16, K8 = (CP % 16) / 8, WORDS = 2 * K16 + K8;
#pragma unroll 1
  for (int dx = 0; dx < 3; ++dx) {
    uint32_t b[3][NT][WORDS];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int j = 0; j < WORDS; ++j) b[dy][n][j] = wfrag[(((dy * 3 + dx) * NT + n) * WORDS + j) * 32 + lane];
    const uint32_t col = a_base + dx * PXP;
#pragma unroll
    for (int i = 0; i < R + 2; ++i) {
      if (i < rows_in) {                      // warp-uniform
        uint32_t a[K16 > 0 ? K16 : 1][4], a8[2];
#pragma unroll
        for (int s = 0; s < K16; ++s) ldmatrix_x4(a[s][0], a[s][1], a[s][2], a[s][3], col + i * row_pitch + s * 32);
        if (K8) ldmatrix_x2(a8[0], a8[1], col + i * row_pitch + K16 * 32);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int o = i - dy;
          if (o >= 0 && o < R) {
#pragma unroll
            for (int n = 0; n < NT; ++n) {
#pragma unroll
              for (int s = 0; s < K16; ++s) mma16816(acc[o][n], a[s], b[dy][n][2 * s], b[dy][n][2 * s + 1]);
              if (K8) mma1688(acc[o][n], a8[0], a8[1], b[dy][n][2 * K16]);
            }
          }
        }
      }
    }
  }
}

template <int CP, int R, int TH, bool BIAS>
__global__ void __launch_bounds__(256, (DenseCfg<CP, R, TH>::SMEM <= 56 * 1024 ? 4 : (DenseCfg<CP, R, TH>::SMEM <= 113 * 1024 ? 2 : 1)))
cab_dense_kernel(const GsnCabDense d, const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out) {
  using K = DenseCfg<CP, R, TH>;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = blockIdx.z, x0 = blockIdx.x * K::TW, y0 = blockIdx.y * K::TH;
  const uint32_t bar = smem_u32(smem + K::BAR_OFF);

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(K::IN_BYTES) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
            "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(&tm_in)), "r"(0), "r"(x0 - 2), "r"(y0 - 2), "r"(t), "r"(bar)
        : "memory");
  }
  // weights (fragment order) and biases
  {
    const unsigned char *w1 = reinterpret_cast<const unsigned char *>(d.w1pack), *w2 = reinterpret_cast<const unsigned char *>(d.w2pack);
    for (int i = tid; i < K::WFRAG_BYTES / 16; i += 256) {
      cp_async16(smem + K::W_OFF + i * 16, w1 + i * 16, true);
      cp_async16(smem + K::W_OFF + K::WFRAG_BYTES + i * 16, w2 + i * 16, true);
    }
    cp_async_commit();
    float *bs = reinterpret_cast<float *>(smem + K::BIAS_OFF);
    if (tid < 2 * CP) {
      const float *src = tid < CP ? d.bias1 : d.bias2;
      bs[tid] = src ? src[tid < CP ? tid : tid - CP] : 0.f;
    }
    cp_async_wait<0>();
  }
  __syncthreads();   // weights + barrier init visible
  {
    uint32_t done = 0;
    int spin = 0;
    while (!done) {
      if (++spin > (1 << 24)) __trap();   // a lost TMA completion must fault, not hang
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"(bar)
          : "memory");
    }
  }

  const int g = lane >> 2, tig = lane & 3;
  const int arow = (lane & 7) + ((lane >> 3) & 1) * 8, akof = (lane >> 4) * 16;
  const uint32_t *wf1 = reinterpret_cast<const uint32_t *>(smem + K::W_OFF);
  const uint32_t *wf2 = reinterpret_cast<const uint32_t *>(smem + K::W_OFF + K::WFRAG_BYTES);
  const float *bias1 = reinterpret_cast<const float *>(smem + K::BIAS_OFF), *bias2 = bias1 + CP;
  const __half2 slope2 = __float2half2_rn(d.has_prelu ? d.prelu_slope : 1.f);

  // ---- conv1 + PReLU -> intermediate tile (pixels outside the image are zero) ------------------------------------------
  if (warp < K::ROWBLK1 * 2) {
    const int cb = warp & 1, rb = warp >> 1;
    const int my0 = rb * R, mx0 = cb * 16;
    float acc[R][K::NT][4];
#pragma unroll
    for (int o = 0; o < R; ++o)
#pragma unroll
      for (int n = 0; n < K::NT; ++n) acc[o][n][0] = acc[o][n][1] = acc[o][n][2] = acc[o][n][3] = 0.f;
    const uint32_t a_base = smem_u32(smem) + (uint32_t)((my0 * K::IW + mx0 + arow) * K::INP + akof);
    conv3x3_rows<CP, R, K::INP>(acc, a_base, K::IW * K::INP, wf1, lane, min(R + 2, K::IH - my0));
#pragma unroll
    for (int o = 0; o < R; ++o) {
      const int my = my0 + o;
      if (my < K::MH) {
        const int gy = y0 - 1 + my;
        const bool rowok = gy >= 0 && gy < d.H;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int mx = mx0 + g + h * 8, gx = x0 - 1 + mx;
          const bool ok = rowok && gx >= 0 && gx < d.W;
          unsigned char *dp = smem + K::MID_OFF + (my * K::MPW + mx) * K::MIDP + tig * 4;
#pragma unroll
          for (int n = 0; n < K::NT; ++n) {
            float v0 = acc[o][n][h * 2], v1 = acc[o][n][h * 2 + 1];
            if (BIAS) { v0 += bias1[n * 8 + tig * 2]; v1 += bias1[n * 8 + tig * 2 + 1]; }
            // PReLU on the packed pair: max(v,0) + slope*min(v,0) (slope = 1 when the block has no activation)
            const __half2 hv = __floats2half2_rn(v0, v1), hz = __float2half2_rn(0.f);
            const __half2 ho = __hfma2(slope2, __hmin2(hv, hz), __hmax2(hv, hz));
            *reinterpret_cast<uint32_t *>(dp + n * 16) = ok ? *reinterpret_cast<const uint32_t *>(&ho) : 0u;
          }
        }
      }
    }
  }
  __syncthreads();   // intermediate complete; the input tile is dead (its space becomes the output staging tile)

  // ---- conv2 -> output staging tile + channel sums ------------------------------------------------------------------------
  float csum[K::NT][2];
#pragma unroll
  for (int n = 0; n < K::NT; ++n) csum[n][0] = csum[n][1] = 0.f;
  if (warp < K::ROWBLK2 * 2) {
    const int cb = warp & 1, rb = warp >> 1;
    const int oy0 = rb * R, ox0 = cb * 16;
    float acc[R][K::NT][4];
#pragma unroll
    for (int o = 0; o < R; ++o)
#pragma unroll
      for (int n = 0; n < K::NT; ++n) acc[o][n][0] = acc[o][n][1] = acc[o][n][2] = acc[o][n][3] = 0.f;
    const uint32_t a_base = smem_u32(smem + K::MID_OFF) + (uint32_t)((oy0 * K::MPW + ox0 + arow) * K::MIDP + akof);
    conv3x3_rows<CP, R, K::MIDP>(acc, a_base, K::MPW * K::MIDP, wf2, lane, min(R + 2, K::MH - oy0));
#pragma unroll
    for (int o = 0; o < R; ++o) {
      const int oy = oy0 + o;
      if (oy < K::TH) {
        const bool rowok = y0 + oy < d.H;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ox = ox0 + g + h * 8;
          if (ox < K::TW) {
            const bool ok = rowok && x0 + ox < d.W;
            unsigned char *dp = smem + (oy * K::TW + ox) * (CP * 2) + tig * 4;
#pragma unroll
            for (int n = 0; n < K::NT; ++n) {
              float v0 = acc[o][n][h * 2], v1 = acc[o][n][h * 2 + 1];
              if (BIAS) { v0 += bias2[n * 8 + tig * 2]; v1 += bias2[n * 8 + tig * 2 + 1]; }
              *reinterpret_cast<uint32_t *>(dp + n * 16) = pack_half2(v0, v1);
              if (ok) {                          // fp32 sums of the in-image pixels (same as gsn_conv_mma's chan_partial)
                csum[n][0] += v0;
                csum[n][1] += v1;
              }
            }
          }
        }
      }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // staging writes -> visible to the TMA store
  // deterministic channel sums: lanes with the same tig -> warp -> CTA
  float *red = reinterpret_cast<float *>(smem + K::RED_OFF);
#pragma unroll
  for (int n = 0; n < K::NT; ++n)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float v = csum[n][j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (g == 0) red[warp * CP + n * 8 + tig * 2 + j] = v;
    }
  __syncthreads();
  if (tid == 0) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];\n" ::
                     "l"(reinterpret_cast<uint64_t>(&tm_out)), "r"(0), "r"(x0), "r"(y0), "r"(t), "r"(smem_u32(smem))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
  }
  if (tid < CP && d.chan_partial) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w * CP + tid];
    const size_t tile = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    d.chan_partial[((size_t)t * gridDim.x * gridDim.y + tile) * CP + tid] = s;
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // smem must outlive the store's reads
}

template <int CP, int R, int TH>
static int launch_cab_dense(const GsnCabDense &d, cudaStream_t st) {
  using K = DenseCfg<CP, R, TH>;
  GSN_ONCE_PER_DEVICE(
    cudaFuncSetAttribute(cab_dense_kernel<CP, R, TH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
    cudaFuncSetAttribute(cab_dense_kernel<CP, R, TH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
  CUtensorMap tm_in, tm_out;
  memset(&tm_in, 0, sizeof(tm_in));
  memset(&tm_out, 0, sizeof(tm_out));
  if (!encode_tmap_nhwc(&tm_in, d.x, CP, d.W, d.H, d.T, CP, K::IW, K::IH) ||
      !encode_tmap_nhwc(&tm_out, d.r, CP, d.W, d.H, d.T, CP, K::TW, K::TH)) {
    set_error("cab_dense: cuTensorMapEncodeTiled unavailable or misaligned tensors");
    return GSN_E_CUDA;
  }
  dim3 grid((d.W + K::TW - 1) / K::TW, (d.H + K::TH - 1) / K::TH, d.T);
  if (d.bias1 || d.bias2) cab_dense_kernel<CP, R, TH, true><<<grid, 256, K::SMEM, st>>>(d, tm_in, tm_out);
  else cab_dense_kernel<CP, R, TH, false><<<grid, 256, K::SMEM, st>>>(d, tm_in, tm_out);
  count_launch();
  return check_launch("cab_dense");
}

// A/B switch (tests, tuning): GSN_CAB_DENSE_SMALL=1 runs the 16-channel instance on 14-row tiles (4 CTAs per SM).  Read at
// every launch -- launches are host-side rare (CUDA-graph replay) -- so one process can exercise both.
static bool dense_small_tiles() {
  const char *e = getenv("GSN_CAB_DENSE_SMALL");
  return e && e[0] == '1';
}

}  // namespace gsn

extern "C" int gsn_cab_dense_tiles(int cp, int H, int W) {
  const int th = (cp == 16 && !gsn::dense_small_tiles()) ? 30 : 14;
  return ((H + th - 1) / th) * ((W + 29) / 30);
}

extern "C" int gsn_cab_dense(const GsnCabDense *dp, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(dp != nullptr, "cab_dense: null descriptor");
  const GsnCabDense &d = *dp;
  GSN_REQUIRE(d.x && d.r && d.w1pack && d.w2pack, "cab_dense: null pointer");
  GSN_REQUIRE(d.T > 0 && d.H > 0 && d.W > 0, "cab_dense: empty shape");
  GSN_REQUIRE(d.x != d.r, "cab_dense: in-place operation is not supported (tiles read their neighbours' halo)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (d.cp) {
    case 16: return dense_small_tiles() ? launch_cab_dense<16, 4, 14>(d, st) : launch_cab_dense<16, 8, 30>(d, st);
    case 24: return launch_cab_dense<24, 4, 14>(d, st);
    default: set_error("cab_dense: cp=%d unsupported (16, 24)", d.cp); return GSN_E_UNSUPPORTED;
  }
}
