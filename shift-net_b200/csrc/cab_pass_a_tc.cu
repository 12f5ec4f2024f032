// Pass A of the fused shift + NAF block, Blackwell-native version (sm_100a):
//   * both 1x1 channel-mixing convs are tcgen05.mma (M=128 pixel tiles, N=2C, K-major no-swizzle operands in the
//     k-chunk planar shared-memory layout, fp32 accumulators in TMEM, issued by one thread, completion on an mbarrier);
//   * TMEM doubles as the parking space of the 2C-wide tensor (4 x 128 columns = all 512 columns for the 22x22 halo'd
//     region), which is what lets a 16x16-pixel tile with its 3-pixel halo fit next to the gather bounding box;
//   * depthwise 3x3 / 5x5 stages run on CUDA cores with packed HFMA2 in "scatter" form: a thread walks down a column,
//     each loaded input row updates the 3 (5) output rows it contributes to, weights live in registers;
//   * the grouped spatial-temporal shift is a per-channel index-offset gather out of a staged bounding box of the
//     neighbour frame's half of the channels, fused with conv1 (dw3x3) and the LayerNorm in the load stage.
//
// Reference semantics: gshift_deblur2.py:186-258 (CAB1/CAB2), :465-519 (spatial_shift2 / channel_shift).
// Stage order and zero-padding rules are identical to the mma.sync version in shift_cab.cu (kept as a cross-check).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "shift_common.cuh"
#include "tc_common.cuh"

namespace gsn {

constexpr int kTcThreads = 512;

// ---- configuration ----------------------------------------------------------------------------------
// SHIFT: CAB2 (LayerNorm input = [rolled stream | conv1(shifted half)]).  BOX: the gather runs inside this kernel (bounding
// box staged in smem); SHIFT && !BOX: the shifted+conv1'd half was produced by shift_conv1_kernel and is read from HBM.
// TMAIN: the LayerNorm inputs of the tile (22x22 halo'd region) are staged in shared memory by TMA tile loads
// (hardware zero fill outside the image) instead of per-lane global loads.
template <int C, bool SHIFT, bool BOX = SHIFT, bool TMAIN = false>
struct TcCfg {
  static constexpr int TW = 16, TH = 16;
  static constexpr int HC = C / 2;
  static constexpr int CIN = SHIFT ? C + HC : C;
  static constexpr int KC1 = CIN / 8, KC2 = C / 8, NC = 2 * C / 8;  // k-chunks of GEMM1/GEMM2, chunks of the 2C tensor
  static constexpr int R1W = TW + 6, R1H = TH + 6, M1 = R1W * R1H;   // 22x22 = 484: region of the 2C tensor
  static constexpr int MT1 = (M1 + 127) / 128;                       // 4 UMMA M tiles
  static constexpr int R2W = TW + 4, R2H = TH + 4, M2 = R2W * R2H;   // 20x20 gated region
  static constexpr int M3 = TW * TH, MT3 = M3 / 128;                 // 256 output pixels, 2 UMMA M tiles
  static constexpr int BW = TW + 24, BH = TH + 24;                   // 40x40 gather bounding box
  static constexpr int N = 2 * C;                                    // UMMA N (128)
  static_assert(MT1 * N <= 512, "GEMM1 accumulators must fit the 512 TMEM columns");
  // weight blob offsets (identical to PassACfg in shift_cab.cu / host/packing.py)
  static constexpr int OFF_LN = 0;
  static constexpr int OFF_C1 = OFF_LN + 2 * CIN * 4;
  static constexpr int OFF_W1 = OFF_C1 + (SHIFT ? 9 * HC * 2 : 0);
  static constexpr int W1_BYTES = KC1 * N * 16;
  static constexpr int OFF_DA = OFF_W1 + W1_BYTES;
  static constexpr int DA_BYTES = 9 * 2 * C * 2;
  static constexpr int OFF_DB = OFF_DA + DA_BYTES;
  static constexpr int DB_BYTES = 25 * C * 2;
  static constexpr int OFF_W2 = OFF_DB + DB_BYTES;
  static constexpr int W2_BYTES = KC2 * N * 16;
  static constexpr int BLOB = OFF_W2 + W2_BYTES;
  static constexpr int WT2_BYTES = DA_BYTES + DB_BYTES + W2_BYTES;   // contiguous tail of the blob
  // plane pitches
  static constexpr int P1 = (M1 + 1) * 16, P2 = (M2 + 1) * 16, P3 = (M3 + 1) * 16;
  // shared memory map.  Phase 1: [X | W1 | A1 | R12]; phase 2: [X | G1 ........ | GATED | WT2], A2/Z alias G1.
  static constexpr int S_X = 0;                                      // LN params, conv1 weights, barrier, tmem ptr, sums
  static constexpr int X_LN = 0, X_C1 = 1024, X_BAR = 1728, X_TMEM = 1744, X_CNT = 1752 /* 4 x u32 arrival counters */, X_BARG = 1768 /* GEMM1 mbarrier */, X_RED = 1792, X_BYTES = 1792 + 16 * 32 * 4;   // X_RED: [16 warps][32] channel-sum partials
  static constexpr int S_W1 = (X_BYTES + 127) / 128 * 128;
  static constexpr int S_A1 = S_W1 + W1_BYTES;
  static constexpr int A1_BYTES = KC1 * P1;
  static constexpr int S_R = (S_A1 + A1_BYTES + 127) / 128 * 128;
  static constexpr int R12_BYTES = BOX ? BW * BH * HC * 2 : 0;
  static constexpr int S_G1 = S_W1;
  static constexpr int G1_BYTES = NC * P1;
  static constexpr int IN_BYTES = TMAIN ? (SHIFT ? 3 * M1 * HC * 2 : M1 * C * 2) : 0;    // TMA staging of the LN inputs
  static constexpr int END1 = S_R + (R12_BYTES > IN_BYTES ? R12_BYTES : IN_BYTES);
  // CAB1 with TMA staging has room to keep GATED and the phase-2 weights ABOVE the staging area: the next tile's inputs can
  // then be prefetched as soon as G1 is dead (after the first depthwise stage) instead of after GEMM2, and the phase-2
  // weights are loaded once per CTA instead of once per tile.  (CAB2's 1.5x wider staging does not leave that room.)
  static constexpr bool EARLY_PF = TMAIN && !SHIFT;
  static constexpr int GT_LO = EARLY_PF ? END1 : S_R;
  static constexpr int S_GT = ((S_G1 + G1_BYTES > GT_LO ? S_G1 + G1_BYTES : GT_LO) + 127) / 128 * 128;
  static constexpr int GT_BYTES = KC2 * P2;
  static constexpr int S_WT2 = (S_GT + GT_BYTES + 127) / 128 * 128;
  static constexpr int END2 = S_WT2 + WT2_BYTES;
  static constexpr bool LATE_WT2 = (BOX || TMAIN) && !EARLY_PF;   // phase-2 weights go where the box / staging lived: load them later
  static constexpr int SMEM = END1 > END2 ? END1 : END2;
  // GEMM2 operand, then the z staging tile.  With TMAIN the kernel is persistent and prefetches the NEXT tile's W1 (at
  // S_W1) and LN inputs (at S_R) while the current tile is in its tail, so A2/z must not sit on W1: it goes 40 KB into
  // the (dead) G1 area, still below S_R.
  static constexpr int S_A2 = TMAIN ? S_G1 + 40 * 1024 : S_G1;
  static_assert(!TMAIN || (S_A2 >= S_W1 + W1_BYTES && S_A2 + KC2 * P3 <= S_R), "A2 must not overlap the prefetch targets");
  static_assert(2 * CIN * 4 <= 1024 && 9 * HC * 2 <= 704, "X region layout");
  static_assert(KC2 * P3 <= G1_BYTES, "A2 aliases G1");
  static_assert(!LATE_WT2 || S_WT2 >= S_R, "WT2 must sit inside the (dead) gather box / staging, not over A1/W1");
  static_assert(LATE_WT2 || S_WT2 >= S_A1 + A1_BYTES, "WT2 must not overlap A1 while GEMM1 reads it");
  static_assert(!(BOX && TMAIN), "the in-kernel gather variant keeps its cp.async loaders");
  static_assert(S_A1 + (KC1 - 1) * P1 + (MT1 * 128) * 16 <= SMEM, "UMMA rows beyond M1 must stay inside the allocation");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <int C, bool SHIFT, bool MIDCA, bool BOX, bool TMAIN>
__global__ void __launch_bounds__(kTcThreads, 1) cab_pass_a_tc_kernel(const GsnCabPassA d, const ShiftTable tab,
                                                                      const __grid_constant__ CUtensorMap tm_x,
                                                                      const __grid_constant__ CUtensorMap tm_hw) {
  using K = TcCfg<C, SHIFT, BOX, TMAIN>;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // tiles: linear index = (frame * tiles_y + tile_y) * tiles_x + tile_x.  TMAIN variants are persistent (gridDim.x = #SMs,
  // stride gridDim.x, next tile's inputs prefetched during the current tile's tail); the others run one tile per CTA.
  const int tiles_x = (d.W + K::TW - 1) / K::TW, tiles_y = (d.H + K::TH - 1) / K::TH;
  const int total_tiles = tiles_x * tiles_y * d.T;
  int tile = blockIdx.x;
  int t, x0, y0;
  RollSrc rs;
  auto decode = [&](int ti) {
    t = ti / (tiles_x * tiles_y);
    const int r = ti - t * tiles_x * tiles_y, ty = r / tiles_x;
    y0 = ty * K::TH;
    x0 = (r - ty * tiles_x) * K::TW;
    rs = roll_source(d.mode, d.circular, t, d.T, C);
  };
  decode(tile);
  const __half *xg = reinterpret_cast<const __half *>(d.x);
  const unsigned char *wb = reinterpret_cast<const unsigned char *>(d.wblob);
  const size_t frame = (size_t)d.H * d.W * C;
  const uint32_t bar = smem_u32(smem + K::S_X + K::X_BAR);
  const uint32_t smem16 = smem_u32(smem) >> 4;   // UMMA descriptors address shared memory in 16-byte units
  // debug_stage == 9: thread 0 of every CTA records clock64() at the stage boundaries (profiling aid, tests only)
  // (the SECOND tile of a persistent CTA when it has one: steady state, inputs prefetched)
  long long *clk = nullptr;
  const bool clk_second = TMAIN && (int)blockIdx.x + (int)gridDim.x < total_tiles;
  if (d.debug_stage == 9 && tid == 0 && !clk_second) clk = reinterpret_cast<long long *>(d.debug_out) + (size_t)tile * 16;
  int clk_i = 0;
#define GSN_CLK() do { if (clk && clk_i < 16) clk[clk_i++] = clock64(); } while (0)
  GSN_CLK();
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + K::S_X + K::X_TMEM);

  // ---- P0 (once per CTA): barriers, TMEM, LN params; first tile's loads -------------------------------------------
  const uint32_t bar_in = bar + 8;
  // GEMM1 hand-off: ln_cnt[m] counts the warps that finished their share of M tile m of A1 (monotonic, 16 per tile); the warp
  // that arrives last issues that M tile's MMAs and commits them to bar_g1 (MT1 commits per phase).
  const uint32_t ln_cnt = smem_u32(smem + K::S_X + K::X_CNT), bar_g1 = smem_u32(smem + K::S_X + K::X_BARG);
  if (tid == 32) {
    mbar_init(bar, 1);
    mbar_init(bar_in, 1);
    mbar_init(bar_g1, K::MT1);
#pragma unroll
    for (int m = 0; m < K::MT1; ++m) *reinterpret_cast<volatile uint32_t *>(smem + K::S_X + K::X_CNT + 4 * m) = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  // Loads of one tile that may run ahead of its compute: the LN inputs of the 22x22 region as one (CAB1) or three (CAB2:
  // rolled low half, rolled high half, shifted half) TMA tile loads -- pixels outside the image arrive as zeros -- and W1.
  auto issue_inputs = [&](int it, int ix0, int iy0, const RollSrc &irs) {
    if (TMAIN && tid == 0) {
      fence_async_proxy();   // earlier generic-proxy accesses of the staging area stay ordered before the TMA writes
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_in), "r"(K::IN_BYTES) : "memory");
      const uint32_t dst = smem_u32(smem + K::S_R);
      auto tma4 = [&](uint32_t sdst, const CUtensorMap *tm, int c0, int f) {
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
                "r"(sdst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(ix0 - 3), "r"(iy0 - 3), "r"(f), "r"(bar_in)
            : "memory");
      };
      if (SHIFT) {
        tma4(dst, &tm_x, irs.c_lo, irs.f_lo);
        tma4(dst + K::M1 * K::HC * 2, &tm_x, irs.c_hi, irs.f_hi);
        tma4(dst + 2 * K::M1 * K::HC * 2, &tm_hw, 0, it);
      } else {
        tma4(dst, &tm_x, 0, it);
      }
    }
    for (int i = tid; i < K::W1_BYTES / 16; i += kTcThreads) cp_async16(smem + K::S_W1 + i * 16, wb + K::OFF_W1 + i * 16, true);
  };
  {
    issue_inputs(t, x0, y0, rs);
    for (int i = tid; i < (K::OFF_W1 - K::OFF_LN) / 16; i += kTcThreads) {  // LN params (+ conv1 weights)
      const int off = i * 16;
      unsigned char *dst = (off < K::OFF_C1) ? smem + K::S_X + K::X_LN + off : smem + K::S_X + K::X_C1 + (off - K::OFF_C1);
      cp_async16(dst, wb + off, true);
    }
    if (BOX) {
      const bool fwd = d.mode == GSN_MODE_CAB2_FWD;
      const __half *src = xg + (size_t)(fwd ? rs.f_lo : rs.f_hi) * frame + (fwd ? rs.c_lo : rs.c_hi);
      constexpr int CH = K::HC / 8;
      for (int p = tid; p < K::BW * K::BH; p += kTcThreads) {   // one box pixel (CH x 16 B) per thread-iteration
        const int by = p / K::BW, bx = p - by * K::BW;
        const int gy = y0 - 12 + by, gx = x0 - 12 + bx;
        const bool valid = gy >= 0 && gy < d.H && gx >= 0 && gx < d.W;
        const __half *sp = valid ? src + ((size_t)gy * d.W + gx) * C : src;
        unsigned char *dp = smem + K::S_R + (size_t)p * K::HC * 2;
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) cp_async16(dp + ch * 16, sp + (valid ? ch * 8 : 0), valid);
      }
    } else if (!K::LATE_WT2) {
      for (int i = tid; i < K::WT2_BYTES / 16; i += kTcThreads) cp_async16(smem + K::S_WT2 + i * 16, wb + K::OFF_DA + i * 16, true);
    }
    cp_async_commit();
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    cp_async_wait<0>();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tmem = *tmem_slot;
  GSN_CLK();  // 1: loads landed
  uint32_t in_parity = 0, mma_parity = 0, g1_parity = 0;   // mbarrier phases advance once per TMA batch / tcgen05.commit / tile
  // Prefetch of the next tile's inputs (persistent TMAIN variants), issued once W1 / the staging area of the current tile
  // are dead.
  auto prefetch_next = [&]() {
    if (TMAIN) {
      const int nt = tile + (int)gridDim.x;
      if (nt < total_tiles) {
        const int ntt = nt / (tiles_x * tiles_y), r = nt - ntt * tiles_x * tiles_y, ty = r / tiles_x;
        const RollSrc nrs = roll_source(d.mode, d.circular, ntt, d.T, C);
        issue_inputs(ntt, (r - ty * tiles_x) * K::TW, ty * K::TH, nrs);
        cp_async_commit();
      }
    }
  };
  for (;;) {   // ---- tile loop (a single iteration for the non-persistent variants) ----

  // ---- P1a (BOX): per-channel spatial-shift gather fused with conv1 (dw3x3, zero pad) -> raw A1 planes --------
  if (BOX) {
    const __half *wc1 = reinterpret_cast<const __half *>(smem + K::S_X + K::X_C1);
    const __half *r12 = reinterpret_cast<const __half *>(smem + K::S_R);
    constexpr int SEG = K::R1W / 2;  // 11-pixel row segments
    // destination positions of the shifted tensor touched by this tile: rows y0-4 .. y0+TH+3, cols x0-4 .. x0+TW+3
    const bool interior = y0 >= 4 && y0 + K::TH + 4 <= d.H && x0 >= 4 && x0 + K::TW + 4 <= d.W;
    for (int item = tid; item < K::HC * K::R1H * 2; item += kTcThreads) {
      const int c = item % K::HC, rest = item / K::HC;
      const int ry = rest % K::R1H, seg = rest / K::R1H;
      int dy, dx;
      shift_offset<C>(c, dy, dx);
      const int gy = y0 - 3 + ry;
      float w[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) w[i] = __half2float(wc1[i * K::HC + c]);
      const int rx0 = seg * SEG;
      // source of (shifted row ry+ty-1, shifted col) inside the box: ((ry+ty-1-dy+9)*BW + col-dx+9)
      const __half *rp = r12 + ((ry - 1 - dy + 9) * K::BW - dx + 9) * K::HC + c;
      unsigned char *dstp = smem + K::S_A1 + (C / 8 + c / 8) * K::P1 + (c & 7) * 2;
      float v[3][3];
      if (interior) {
#pragma unroll
        for (int ty = 0; ty < 3; ++ty) {
          v[ty][1] = __half2float(rp[(ty * K::BW + rx0 - 1) * K::HC]);
          v[ty][2] = __half2float(rp[(ty * K::BW + rx0) * K::HC]);
        }
#pragma unroll
        for (int i = 0; i < SEG; ++i) {
          const int rx = rx0 + i;
          float acc = 0.f;
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) {
            v[ty][0] = v[ty][1]; v[ty][1] = v[ty][2];
            v[ty][2] = __half2float(rp[(ty * K::BW + rx + 1) * K::HC]);
            acc = fmaf(v[ty][0], w[ty * 3 + 0], acc);
            acc = fmaf(v[ty][1], w[ty * 3 + 1], acc);
            acc = fmaf(v[ty][2], w[ty * 3 + 2], acc);
          }
          *reinterpret_cast<__half *>(dstp + (ry * K::R1W + rx) * 16) = __float2half_rn(acc);
        }
      } else {
        bool rowok[3];
#pragma unroll
        for (int ty = 0; ty < 3; ++ty) { const int sy = gy + ty - 1; rowok[ty] = sy >= 0 && sy < d.H; }
        auto load_col = [&](int col, float(&o)[3]) {      // col = region column of the shifted tensor (dest must be in-image)
          const int sx = x0 - 3 + col;
          const bool cok = sx >= 0 && sx < d.W;
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) o[ty] = (cok && rowok[ty]) ? __half2float(rp[(ty * K::BW + col) * K::HC]) : 0.f;
        };
        {
          float a[3], b[3];
          load_col(rx0 - 1, a);
          load_col(rx0, b);
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) { v[ty][1] = a[ty]; v[ty][2] = b[ty]; }
        }
#pragma unroll
        for (int i = 0; i < SEG; ++i) {
          const int rx = rx0 + i;
          float nc[3];
          load_col(rx + 1, nc);
          float acc = 0.f;
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) {
            v[ty][0] = v[ty][1]; v[ty][1] = v[ty][2]; v[ty][2] = nc[ty];
            acc = fmaf(v[ty][0], w[ty * 3 + 0], acc);
            acc = fmaf(v[ty][1], w[ty * 3 + 1], acc);
            acc = fmaf(v[ty][2], w[ty * 3 + 2], acc);
          }
          *reinterpret_cast<__half *>(dstp + (ry * K::R1W + rx) * 16) = __float2half_rn(acc);
        }
      }
    }
    __syncthreads();
    GSN_CLK();  // 2: gather done (SHIFT only)
    // the gather box is dead now: stream the phase-2 weights (dw taps + W2) into its tail
    for (int i = tid; i < K::WT2_BYTES / 16; i += kTcThreads) cp_async16(smem + K::S_WT2 + i * 16, wb + K::OFF_DA + i * 16, true);
    cp_async_commit();
  }

  // ---- P1b: LayerNorm over the CIN channels of every region pixel -> A1 (zero rows outside the image) ----
  {
    const float *ln_g = reinterpret_cast<const float *>(smem + K::S_X + K::X_LN);
    const float *ln_b = ln_g + K::CIN;
    constexpr int NV = SHIFT ? 24 : 16;
    constexpr int ITEMS = (K::M1 + 7) / 8 * 8 * 4;   // whole warps only (quad shuffles below)
    constexpr int NIT = (ITEMS + kTcThreads - 1) / kTcThreads;
    static_assert(NIT == K::MT1 && kTcThreads == 4 * 128, "one LayerNorm iteration = one UMMA M tile");
    const int j = tid & 3;                            // the quad lane never changes across a thread's items
    int chunk_of[NV / 8];
    if (SHIFT) { chunk_of[0] = j; chunk_of[1] = K::HC / 8 + j; chunk_of[2] = C / 8 + j; }
    else {
      // CAB1: a quad owns one pixel (128 staged bytes).  Lane j takes chunks {j, j+4}, even pixels low chunk first, odd pixels high
      // chunk first, so the 8 lanes of a quarter warp (2 pixels) hit 8 different 16-byte bank groups in each LDS.128.
      const int par = (tid >> 2) & 1;               // pixel parity: constant per thread (items advance by 128 pixels)
      chunk_of[0] = j + 4 * par; chunk_of[1] = j + 4 * (1 - par);
    }
    float gam[NV], bet[NV];
#pragma unroll
    for (int k = 0; k < NV / 8; ++k)
#pragma unroll
      for (int i = 0; i < 8; i += 4) {
        const float4 g4 = *reinterpret_cast<const float4 *>(ln_g + chunk_of[k] * 8 + i), b4 = *reinterpret_cast<const float4 *>(ln_b + chunk_of[k] * 8 + i);
        gam[k * 8 + i] = g4.x; gam[k * 8 + i + 1] = g4.y; gam[k * 8 + i + 2] = g4.z; gam[k * 8 + i + 3] = g4.w;
        bet[k * 8 + i] = b4.x; bet[k * 8 + i + 1] = b4.y; bet[k * 8 + i + 2] = b4.z; bet[k * 8 + i + 3] = b4.w;
      }
    if (TMAIN) { mbar_wait(bar_in, in_parity); in_parity ^= 1; }   // the staged LN inputs of this tile have landed
    GSN_CLK();  // input wait over
    const unsigned char *stg = smem + K::S_R;
    constexpr int PB = (SHIFT && !BOX) ? 2 : NIT;    // items whose loads are in flight together (register budget)
#pragma unroll
    for (int it0 = 0; it0 < NIT; it0 += PB) {
    uint4 raw[NIT][(SHIFT && !BOX) ? 3 : 2];
#pragma unroll
    for (int it = it0; it < it0 + PB && it < NIT; ++it) {   // issue the batch's global loads first (memory-level parallelism)
      const int q = (tid + it * kTcThreads) >> 2;
      const int ry = q / K::R1W, rx = q - ry * K::R1W;
      const int gy = y0 - 3 + ry, gx = x0 - 3 + rx;
      const bool inimg = (q < K::M1) && gy >= 0 && gy < d.H && gx >= 0 && gx < d.W;
      raw[it][0] = raw[it][1] = make_uint4(0, 0, 0, 0);
      if (SHIFT && !BOX) raw[it][(SHIFT && !BOX) ? 2 : 0] = make_uint4(0, 0, 0, 0);
      if (TMAIN) {
        if (q < K::M1) {
          if (SHIFT) {
            raw[it][0] = *reinterpret_cast<const uint4 *>(stg + (size_t)q * K::HC * 2 + j * 16);
            raw[it][1] = *reinterpret_cast<const uint4 *>(stg + K::M1 * K::HC * 2 + (size_t)q * K::HC * 2 + j * 16);
            raw[it][(SHIFT && !BOX) ? 2 : 0] = *reinterpret_cast<const uint4 *>(stg + 2 * K::M1 * K::HC * 2 + (size_t)q * K::HC * 2 + j * 16);
          } else {
            raw[it][0] = *reinterpret_cast<const uint4 *>(stg + (size_t)q * C * 2 + chunk_of[0] * 16);
            raw[it][1] = *reinterpret_cast<const uint4 *>(stg + (size_t)q * C * 2 + chunk_of[1] * 16);
          }
        }
      } else if (inimg) {
        const size_t pix = ((size_t)gy * d.W + gx) * C;
        if (SHIFT) {
          raw[it][0] = __ldg(reinterpret_cast<const uint4 *>(xg + rs.f_lo * frame + pix + rs.c_lo + j * 8));
          raw[it][1] = __ldg(reinterpret_cast<const uint4 *>(xg + rs.f_hi * frame + pix + rs.c_hi + j * 8));
          if (!BOX) raw[it][(SHIFT && !BOX) ? 2 : 0] = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const __half *>(d.hw_pre) + ((size_t)t * d.H * d.W + (size_t)gy * d.W + gx) * K::HC + j * 8));
        } else {
          raw[it][0] = __ldg(reinterpret_cast<const uint4 *>(xg + (size_t)t * frame + pix + chunk_of[0] * 8));
          raw[it][1] = __ldg(reinterpret_cast<const uint4 *>(xg + (size_t)t * frame + pix + chunk_of[1] * 8));
        }
      }
    }
#pragma unroll
    for (int it = it0; it < it0 + PB && it < NIT; ++it) {
      const int item = tid + it * kTcThreads;
      if (item < ITEMS) {                            // warp-uniform (ITEMS is a multiple of 32)
        const int q = item >> 2;
        const int ry = q / K::R1W, rx = q - ry * K::R1W;
        const int gy = y0 - 3 + ry, gx = x0 - 3 + rx;
        const bool inimg = (q < K::M1) && gy >= 0 && gy < d.H && gx >= 0 && gx < d.W;
        float v[NV];
        unpack8(raw[it][0], *reinterpret_cast<float(*)[8]>(&v[0]));
        unpack8(raw[it][1], *reinterpret_cast<float(*)[8]>(&v[8]));
        if (SHIFT && !BOX) {
          unpack8(raw[it][(SHIFT && !BOX) ? 2 : 0], *reinterpret_cast<float(*)[8]>(&v[SHIFT ? 16 : 0]));
        } else if (SHIFT) {
          if (inimg) unpack8(*reinterpret_cast<const uint4 *>(smem + K::S_A1 + (C / 8 + j) * K::P1 + q * 16), *reinterpret_cast<float(*)[8]>(&v[16]));
          else {
#pragma unroll
            for (int i = 16; i < NV; ++i) v[i] = 0.f;
          }
        }
        // one pass: s1 = sum v, s2 = sum v^2 (fp32; |mu| <~ sigma here, so E[v^2] - mu^2 loses nothing that matters in fp16)
        float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};   // 4 independent chains each (latency)
#pragma unroll
        for (int i = 0; i < NV; ++i) { s4[i & 3] += v[i]; q4[i & 3] = fmaf(v[i], v[i], q4[i & 3]); }
        float s = (s4[0] + s4[1]) + (s4[2] + s4[3]), ss = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        const float mu = s * (1.f / K::CIN);
        const float rstd = rsqrtf(fmaxf(ss * (1.f / K::CIN) - mu * mu, 0.f) + 1e-6f);
        const float nmr = -mu * rstd;
        if (q < K::M1) {
          if (inimg) {
#pragma unroll
            for (int k = 0; k < NV / 8; ++k) {
              float o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = fmaf(fmaf(v[k * 8 + i], rstd, nmr), gam[k * 8 + i], bet[k * 8 + i]);
              *reinterpret_cast<uint4 *>(smem + K::S_A1 + chunk_of[k] * K::P1 + q * 16) = pack8(o);
            }
          } else {                                       // outside the image: the 1x1 sees zero padding (P3 of SURVEY.md)
#pragma unroll
            for (int k = 0; k < NV / 8; ++k)
              *reinterpret_cast<uint4 *>(smem + K::S_A1 + chunk_of[k] * K::P1 + q * 16) = make_uint4(0, 0, 0, 0);
          }
        }
      }
      // This warp's share of M tile `it` of A1 (pixels [128 it, 128 it + 128)) is complete.  No CTA barrier: every warp bumps
      // the tile's arrival counter and runs on to the next M tile; whichever warp arrives LAST hands the tile to the tensor
      // core.  GEMM1: D[m] (128 x 2C, TMEM) = A1[m] (128 x CIN) . W1^T
      fence_async_proxy();   // generic-proxy writes of A1 (and cp.async'd W1) -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) {
        uint32_t old;
        asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;\n" : "=r"(old) : "r"(ln_cnt + 4 * it) : "memory");
        if ((old & (kTcThreads / 32 - 1)) == kTcThreads / 32 - 1) {
          tc_fence_after();
          constexpr uint32_t idesc = make_idesc_f16(128, K::N);
#pragma unroll
          for (int k = 0; k < K::KC1 / 2; ++k) {
            const uint64_t ad = smem_desc_at(smem16, K::S_A1 + 2 * k * K::P1 + it * 128 * 16, K::P1, 128);
            const uint64_t bd = smem_desc_at(smem16, K::S_W1 + 2 * k * (K::N * 16), K::N * 16, 128);
            umma_f16(tmem + it * K::N, ad, bd, idesc, k > 0);
          }
          umma_commit(bar_g1);   // tcgen05.commit tracks the issuing thread's MMAs: one commit per M tile, MT1 arrivals per phase
        }
      }
      __syncwarp();
      GSN_CLK();  // LN iteration `it` done (the last one = LN done)
    }
    }
    if (TMAIN && K::LATE_WT2) {   // once EVERY warp is done with the staging area, stream the phase-2 weights into its place
      __syncthreads();
      for (int i = tid; i < K::WT2_BYTES / 16; i += kTcThreads) cp_async16(smem + K::S_WT2 + i * 16, wb + K::OFF_DA + i * 16, true);
      cp_async_commit();
    }
  }
  if (d.debug_stage == 1) {
    __syncthreads();
    uint4 *o = reinterpret_cast<uint4 *>(d.debug_out) +
               (size_t)tile * K::KC1 * K::M1;
    for (int i = tid; i < K::KC1 * K::M1; i += kTcThreads)
      o[i] = *reinterpret_cast<uint4 *>(smem + K::S_A1 + (i / K::M1) * K::P1 + (i % K::M1) * 16);
  }

  // ---- P2: GEMM1 was issued M tile by M tile from inside the LayerNorm loop; wait for the last one -----------------
  if (K::LATE_WT2) cp_async_wait<0>();   // phase-2 weights landed (issued after the gather / the LayerNorm)
  mbar_wait(bar_g1, g1_parity);
  g1_parity ^= 1;
  tc_fence_after();
  __syncthreads();                 // A1 / W1 are dead from here on; WT2 visible to everyone
  GSN_CLK();  // GEMM1 done

  // ---- P3: TMEM -> fp16 G1 planes (all 2C channels of the region) ----------------------------------------------
  {
    const int quarter = warp & 3, sub = warp >> 2;   // a warp may only touch TMEM lanes [32*quarter, +32)
    for (int u = sub; u < K::MT1 * (K::N / 32); u += kTcThreads / 128) {
      const int m = u / (K::N / 32), cg = u % (K::N / 32);
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + m * K::N + cg * 32, v);
      const int px = m * 128 + quarter * 32 + lane;
      if (px < K::M1) {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint4 o;
          o.x = pack_half2(__uint_as_float(v[c4 * 8 + 0]), __uint_as_float(v[c4 * 8 + 1]));
          o.y = pack_half2(__uint_as_float(v[c4 * 8 + 2]), __uint_as_float(v[c4 * 8 + 3]));
          o.z = pack_half2(__uint_as_float(v[c4 * 8 + 4]), __uint_as_float(v[c4 * 8 + 5]));
          o.w = pack_half2(__uint_as_float(v[c4 * 8 + 6]), __uint_as_float(v[c4 * 8 + 7]));
          *reinterpret_cast<uint4 *>(smem + K::S_G1 + (cg * 4 + c4) * K::P1 + px * 16) = o;
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    GSN_CLK();  // TMEM -> G1 done
  }

  // ---- P4: dw3x3 + id on both halves, SimpleGate -> GATED (zero outside the image) ---------------------------------
  // The identity of RepConv2 / RepConv is folded into the centre tap by host/packing.py (w_c + 1 in fp16).
  {
    constexpr int NSTRIP = 3, SROWS = (K::R2H + NSTRIP - 1) / NSTRIP;  // 7,7,6 output rows
    const unsigned char *wda = smem + K::S_WT2;
    for (int item = tid; item < K::KC2 * K::R2W * NSTRIP; item += kTcThreads) {
      const int x = item % K::R2W, rest = item / K::R2W;
      const int p = rest % K::KC2, strip = rest / K::KC2;
      const int r0 = strip * SROWS, r1 = min(r0 + SROWS, K::R2H);
      H8 res[SROWS];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int chunk = half * K::KC2 + p;
        const unsigned char *pl = smem + K::S_G1 + chunk * K::P1;
        H8 w[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) w[i] = lds_h8(wda + (i * 2 * C + chunk * 8) * 2);
        H8 acc0, acc1;
#pragma unroll
        for (int i = 0; i < SROWS + 2; ++i) {           // input region row r0 + i feeds output rows (r0+i-2 .. r0+i)
          const int row = r0 + i;
          if (row < r1 + 2) {
            const unsigned char *rp = pl + (row * K::R1W + x) * 16;
            const H8 v0 = lds_h8(rp), v1 = lds_h8(rp + 16), v2 = lds_h8(rp + 32);
            if (i >= 2) {                               // output row r0+i-2 completes with kernel row 2
              h8_fma(acc0, v0, w[6]); h8_fma(acc0, v1, w[7]); h8_fma(acc0, v2, w[8]);
              if (half == 0) res[i - 2] = acc0;
              else {
                const int orow = row - 2;
                const int gy = y0 - 2 + orow, gx = x0 - 2 + x;
                H8 o;
                if (gy >= 0 && gy < d.H && gx >= 0 && gx < d.W) h8_mul(o, res[i - 2], acc0);
                else {
#pragma unroll
                  for (int q = 0; q < 4; ++q) o.h[q] = __float2half2_rn(0.f);
                }
                sts_h8(smem + K::S_GT + p * K::P2 + (orow * K::R2W + x) * 16, o);
              }
            }
            if (i >= 1) {                               // output row r0+i-1: kernel row 1 (centre row)
              acc0 = acc1;
              h8_fma(acc0, v0, w[3]); h8_fma(acc0, v1, w[4]); h8_fma(acc0, v2, w[5]);   // w[4] carries the "+ x" of RepConv2
            }
            h8_mul(acc1, v0, w[0]);                     // output row r0+i: kernel row 0 starts a new accumulator
            h8_fma(acc1, v1, w[1]); h8_fma(acc1, v2, w[2]);
          }
        }
      }
    }
    __syncthreads();
    GSN_CLK();  // dwA done
    if (K::EARLY_PF) prefetch_next();   // G1 is dead: W1's slot and the staging area are free for the next tile's inputs
  }
  if (d.debug_stage == 2) {
    uint4 *o = reinterpret_cast<uint4 *>(d.debug_out) +
               (size_t)tile * K::KC2 * K::M2;
    for (int i = tid; i < K::KC2 * K::M2; i += kTcThreads)
      o[i] = *reinterpret_cast<uint4 *>(smem + K::S_GT + (i / K::M2) * K::P2 + (i % K::M2) * 16);
  }

  // ---- P5: dw5x5 (+ merged dw3x3) + id on the gated tensor -> A2 (GEMM2 operand) -----------------------------------
  {
    constexpr int NSTRIP = 2, SROWS = K::TH / NSTRIP;   // 8 output rows per strip
    const unsigned char *wdb = smem + K::S_WT2 + K::DA_BYTES;
    for (int item = tid; item < 2 * K::KC2 * K::TW * NSTRIP; item += kTcThreads) {
      // lanes = (half chunk parity, x): consecutive lanes read consecutive 8-byte words -> conflict-free LDS.64
      const int e = item & 1, x = (item >> 1) % K::TW, rest = (item >> 1) / K::TW;
      const int hc = (rest % K::KC2) * 2 + e, strip = rest / K::KC2;   // hc: 4-channel half chunk
      const int r0 = strip * SROWS;
      const unsigned char *pl = smem + K::S_GT + (hc >> 1) * K::P2 + (hc & 1) * 8;
      __half2 w[25][2];
#pragma unroll
      for (int i = 0; i < 25; ++i) {
        const uint2 ww = *reinterpret_cast<const uint2 *>(wdb + (i * C + hc * 4) * 2);
        w[i][0] = *reinterpret_cast<const __half2 *>(&ww.x);
        w[i][1] = *reinterpret_cast<const __half2 *>(&ww.y);
      }
      __half2 acc[5][2];
#pragma unroll
      for (int i = 0; i < SROWS + 4; ++i) {             // gated row r0 + i feeds output rows r0+i-4 .. r0+i
        const unsigned char *rp = pl + ((r0 + i) * K::R2W + x) * 16;
        __half2 v[5][2];
#pragma unroll
        for (int tx = 0; tx < 5; ++tx) {
          const uint2 vv = *reinterpret_cast<const uint2 *>(rp + tx * 16);
          v[tx][0] = *reinterpret_cast<const __half2 *>(&vv.x);
          v[tx][1] = *reinterpret_cast<const __half2 *>(&vv.y);
        }
        // slot s = i % 5 holds the accumulator of output row (r0 + i) ; kernel row ky = i - (out row index)
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
          const int oi = i - ky;                        // output row index within the strip
          if (oi < 0 || oi >= SROWS) continue;
          const int s = oi % 5;
#pragma unroll
          for (int tx = 0; tx < 5; ++tx)
#pragma unroll
            for (int e = 0; e < 2; ++e)
              acc[s][e] = (ky == 0 && tx == 0) ? __hmul2(v[tx][e], w[ky * 5 + tx][e]) : __hfma2(v[tx][e], w[ky * 5 + tx][e], acc[s][e]);
        }
        const int od = i - 4;                           // output row completed by this input row
        if (od >= 0) {
          const int s = od % 5;
          uint2 o;
          o.x = *reinterpret_cast<uint32_t *>(&acc[s][0]);
          o.y = *reinterpret_cast<uint32_t *>(&acc[s][1]);
          *reinterpret_cast<uint2 *>(smem + K::S_A2 + (hc >> 1) * K::P3 + ((r0 + od) * K::TW + x) * 16 + (hc & 1) * 8) = o;
        }
      }
    }
    fence_async_proxy();
    __syncthreads();
    GSN_CLK();  // dwB done
  }
  if (d.debug_stage == 3) {
    uint4 *o = reinterpret_cast<uint4 *>(d.debug_out) +
               (size_t)tile * K::KC2 * K::M3;
    for (int i = tid; i < K::KC2 * K::M3; i += kTcThreads)
      o[i] = *reinterpret_cast<uint4 *>(smem + K::S_A2 + (i / K::M3) * K::P3 + (i % K::M3) * 16);
  }

  if (MIDCA) {
    // denoise variants: stop here.  u = RepConv(gate) goes to HBM, the mid CALayer2 needs the frame mean of the gated
    // tensor (sums over this tile's 16x16 centre of GATED); the scale it produces is folded into W2 by cab_fold_mid.
    __half *ug = reinterpret_cast<__half *>(d.z) + (size_t)t * frame;
    for (int i = tid; i < K::M3 * K::KC2; i += kTcThreads) {
      const int ch = i % K::KC2, p = i / K::KC2;
      const int gy = y0 + p / K::TW, gx = x0 + (p % K::TW);
      if (gy < d.H && gx < d.W)
        *reinterpret_cast<uint4 *>(ug + ((size_t)gy * d.W + gx) * C + ch * 8) = *reinterpret_cast<const uint4 *>(smem + K::S_A2 + ch * K::P3 + p * 16);
    }
    float *red = reinterpret_cast<float *>(smem + K::S_X + K::X_RED);
    for (int u = warp; u < K::KC2 * 2; u += kTcThreads / 32) {
      const int ch = u % K::KC2, hf = u / K::KC2;
      float s[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] = 0.f;
      for (int p = hf * (K::M3 / 2) + lane; p < (hf + 1) * (K::M3 / 2); p += 32) {
        const int oy = p / K::TW, ox = p % K::TW;
        if (y0 + oy < d.H && x0 + ox < d.W) {
          float f[8];
          unpack8(*reinterpret_cast<const uint4 *>(smem + K::S_GT + ch * K::P2 + ((oy + 2) * K::R2W + ox + 2) * 16), f);
#pragma unroll
          for (int i = 0; i < 8; ++i) s[i] += f[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[hf * C + ch * 8 + i] = s[i];
      }
    }
    __syncthreads();
    if (!K::EARLY_PF) prefetch_next();   // GATED / A2 reads are done: the staging area and W1 are free for the next tile
    if (tid < C) {
      d.chan_partial[(size_t)tile * C + tid] = red[tid] + red[C + tid];
    }
  } else {

  // ---- P6: GEMM2 on the tensor core: (256 x C) . W2^T -> TMEM columns [0, 2*N) -------------------------------------
  if (tid == 0) {
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_f16(128, K::N);
#pragma unroll
    for (int m = 0; m < K::MT3; ++m)
#pragma unroll
      for (int k = 0; k < K::KC2 / 2; ++k) {
        const uint64_t ad = smem_desc_at(smem16, K::S_A2 + 2 * k * K::P3 + m * 128 * 16, K::P3, 128);
        const uint64_t bd = smem_desc_at(smem16, K::S_WT2 + K::DA_BYTES + K::DB_BYTES + 2 * k * (K::N * 16), K::N * 16, 128);
        umma_f16(tmem + m * K::N, ad, bd, idesc, k > 0);
      }
    umma_commit(bar);
  }
  mbar_wait(bar, mma_parity);
  mma_parity ^= 1;
  tc_fence_after();
  __syncthreads();                 // A2 is dead: its space becomes the z staging tile
  GSN_CLK();  // GEMM2 done
  if (!K::EARLY_PF) prefetch_next();   // W1, W2 and the staging area are dead: start the next tile's loads under this tile's tail

  // ---- P7: a * sigmoid(b) (SimpleGate2) -> z tile (fp16 planes) -> coalesced global store; per-tile channel sums --------
  {
    const int quarter = warp & 3, sub = warp >> 2;
    float *red = reinterpret_cast<float *>(smem + K::S_X + K::X_RED);   // [16 warps][32 channels]
    static_assert(K::MT3 * (C / 32) == kTcThreads / 128, "one (M tile, 32-channel group) unit per warp");
    {
      const int m = sub / (C / 32), cg = sub % (C / 32);
      uint32_t a[32], b[32];
      const uint32_t base = tmem + ((uint32_t)(quarter * 32) << 16) + m * K::N + cg * 32;
      tmem_ld32(base, a);
      tmem_ld32(base + C, b);
      const int px = m * 128 + quarter * 32 + lane;
      const bool valid = (y0 + px / K::TW < d.H) && (x0 + px % K::TW < d.W);
      float z[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) z[i] = __uint_as_float(a[i]) * sigmoid_tanh(__uint_as_float(b[i]));
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4)
        *reinterpret_cast<uint4 *>(smem + K::S_A2 + (cg * 4 + c4) * K::P3 + px * 16) = pack8(*reinterpret_cast<float(*)[8]>(&z[c4 * 8]));
      // sum over the warp's 32 pixels of each of its 32 channels: butterfly that halves the live values per step;
      // lane l ends up with channel cg*32 + l (fixed order => deterministic)
      if (!valid) {
#pragma unroll
        for (int i = 0; i < 32; ++i) z[i] = 0.f;
      }
#pragma unroll
      for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
        const bool hi = lane & off;
#pragma unroll
        for (int i = 0; i < n; ++i) {
          const float send = hi ? z[i] : z[i + n], keep = hi ? z[i + n] : z[i];
          z[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      red[warp * 32 + lane] = z[0];
    }
    tc_fence_before();
    __syncthreads();
    GSN_CLK();  // gate2 -> z tile done
    __half *zg = reinterpret_cast<__half *>(d.z) + (size_t)t * frame;
    for (int i = tid; i < K::M3 * K::KC2; i += kTcThreads) {
      const int ch = i % K::KC2, p = i / K::KC2;
      const int gy = y0 + p / K::TW, gx = x0 + (p % K::TW);
      if (gy < d.H && gx < d.W)
        *reinterpret_cast<uint4 *>(zg + ((size_t)gy * d.W + gx) * C + ch * 8) = *reinterpret_cast<const uint4 *>(smem + K::S_A2 + ch * K::P3 + p * 16);
    }
    if (tid < C) {   // channel tid = group cg, lane l: the 8 warps (4 lane quarters x 2 M tiles) that own the group
      const int cg = tid >> 5, l = tid & 31;
      float s = 0.f;
#pragma unroll
      for (int m = 0; m < K::MT3; ++m)
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) s += red[(qq + 4 * (m * (C / 32) + cg)) * 32 + l];
      d.chan_partial[(size_t)tile * C + tid] = s;
    }
  }
  GSN_CLK();  // stores + sums done
  }   // !MIDCA
    clk = nullptr;                 // stage clocks are recorded for one tile of a CTA only
    if (!TMAIN) break;
    tile += (int)gridDim.x;
    if (tile >= total_tiles) break;
    decode(tile);
    cp_async_wait<0>();            // the prefetched W1 of this tile has landed
    tc_fence_before();
    __syncthreads();               // everyone is done with the previous tile (z staging, TMEM reads, channel sums)
    tc_fence_after();
    if (d.debug_stage == 9 && tid == 0 && clk_second && tile == (int)blockIdx.x + (int)gridDim.x) {
      clk = reinterpret_cast<long long *>(d.debug_out) + (size_t)tile * 16;
      GSN_CLK();   // "start"
      GSN_CLK();   // "loads" (prefetched during the previous tile)
    }
  }   // tile loop
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512));
  }
}

template <int C, bool SHIFT, bool MIDCA, bool BOX, bool TMAIN>
static int launch_pass_a_tc(const GsnCabPassA &d, cudaStream_t st, const CUtensorMap &tm_x, const CUtensorMap &tm_hw) {
  using K = TcCfg<C, SHIFT, BOX, TMAIN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(cab_pass_a_tc_kernel<C, SHIFT, MIDCA, BOX, TMAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
    attr_set = true;
  }
  static const ShiftTable tab = make_shift_table(C);
  const long long total = (long long)((d.W + K::TW - 1) / K::TW) * ((d.H + K::TH - 1) / K::TH) * d.T;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  // TMAIN variants are persistent: one CTA per SM walks the tile list and prefetches the next tile's inputs
  const unsigned grid = (unsigned)(TMAIN ? (total < num_sms ? total : num_sms) : total);
  cab_pass_a_tc_kernel<C, SHIFT, MIDCA, BOX, TMAIN><<<grid, kTcThreads, K::SMEM, st>>>(d, tab, tm_x, tm_hw);
  count_launch();
  return check_launch("cab_pass_a_tc");
}

template <int C, bool SHIFT, bool BOX, bool TMAIN>
static int launch_pass_a_mid(const GsnCabPassA &d, cudaStream_t st, const CUtensorMap &tm_x, const CUtensorMap &tm_hw) {
  return d.mid_ca ? launch_pass_a_tc<C, SHIFT, true, BOX, TMAIN>(d, st, tm_x, tm_hw)
                  : launch_pass_a_tc<C, SHIFT, false, BOX, TMAIN>(d, st, tm_x, tm_hw);
}

int cab_pass_a_tc_dispatch(const GsnCabPassA &d, cudaStream_t st) {
  if (d.C != 64) {
    set_error("cab_pass_a: C=%d unsupported by the fused kernel (64)", d.C);
    return GSN_E_UNSUPPORTED;
  }
  constexpr int C = 64;
  static const bool want_tma = [] { const char *e = getenv("GSN_PASS_A_TMA"); return !(e && e[0] == '0'); }();
  CUtensorMap tm_x, tm_hw;
  memset(&tm_x, 0, sizeof(tm_x));
  memset(&tm_hw, 0, sizeof(tm_hw));
  const bool shift = d.mode != GSN_MODE_CAB1;
  const bool split = shift && d.hw_pre != nullptr;      // shifted half precomputed by gsn_shift_conv1
  bool tma = want_tma && (!shift || split);
  if (tma) tma = encode_tmap_nhwc(&tm_x, d.x, C, d.W, d.H, d.T, shift ? C / 2 : C, 22, 22);
  if (tma && split) tma = encode_tmap_nhwc(&tm_hw, d.hw_pre, C / 2, d.W, d.H, d.T, C / 2, 22, 22);
  if (!shift) return tma ? launch_pass_a_mid<C, false, false, true>(d, st, tm_x, tm_hw) : launch_pass_a_mid<C, false, false, false>(d, st, tm_x, tm_hw);
  if (!split) return launch_pass_a_mid<C, true, true, false>(d, st, tm_x, tm_hw);
  return tma ? launch_pass_a_mid<C, true, false, true>(d, st, tm_x, tm_hw) : launch_pass_a_mid<C, true, false, false>(d, st, tm_x, tm_hw);
}

}  // namespace gsn
