// Pass A of the fused shift + NAF block, "pre-normalised" variant (sm_100a).
//
// The LayerNorm'd operand A1 of the first 1x1 (gshift_deblur2.py:209,250: ``self.norm``) is produced ONCE per pixel by an
// HBM-bound producer (ln_planar_kernel below, or the epilogue of the previous pass B / of shift_conv1) and stored in HBM in
// the k-chunk planar layout [T][KC1][H][W][8] fp16.  One TMA tile load ({22*8, 22, KC1, 1} box, zero fill outside the image
// = the zero padding the following convs see) then lands the 22x22 halo'd region of a tile DIRECTLY in the no-swizzle
// K-major UMMA operand layout: this kernel has no LayerNorm stage, no staging buffer and no generic-proxy writes of A1.
// What that buys over cab_pass_a_tc_kernel:
//   * the LayerNorm (21 % of the tile time there, run on 1.89x the pixels because of the halo) leaves the issue-bound kernel;
//   * the freed shared memory holds W1 and the phase-2 weights permanently (loaded once per CTA, also for CAB2);
//   * GEMM1 of the NEXT tile is issued while the current tile is still in its CUDA-core stages: M tiles 2,3 (TMEM columns
//     256..511, free once the accumulators were drained to shared memory) right behind GEMM2, M tiles 0,1 (columns 0..255,
//     which GEMM2 uses) as soon as the sigmoid gate has pulled GEMM2's output out of TMEM (named barrier, only the issuing
//     warp waits) -- the tensor core and the TMA run under the CUDA-core work of the previous tile;
//   * the z tile leaves through one swizzled TMA tile store.
// Stages P3..P7 (TMEM -> G1, dw3x3 + gate, dw5x5, GEMM2, sigmoid gate, store + channel sums) are those of
// cab_pass_a_tc.cu; reference semantics gshift_deblur2.py:186-258.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "shift_common.cuh"
#include "tc_common.cuh"

namespace gsn {

constexpr int kPreThreads = 512;

template <int KC1>
struct PreCfg {
  static constexpr int C = 64, TW = 16, TH = 16, HC = C / 2;
  static constexpr bool SHIFT = KC1 == 12;                            // CAB2: LN input = [rolled stream | conv1(shifted half)]
  static constexpr int CIN = KC1 * 8;
  static constexpr int KC2 = C / 8, NC = 2 * C / 8, N = 2 * C;
  static constexpr int R1W = TW + 6, R1H = TH + 6, M1 = R1W * R1H, MT1 = 4;   // 22x22 region of the 2C tensor, 4 UMMA M tiles
  static constexpr int R2W = TW + 4, R2H = TH + 4, M2 = R2W * R2H;           // 20x20 gated region
  static constexpr int M3 = TW * TH, MT3 = M3 / 128;
  // weight blob offsets (host/packing.py pack_cab_pass_a; identical to TcCfg)
  static constexpr int OFF_C1 = 2 * CIN * 4;
  static constexpr int OFF_W1 = OFF_C1 + (SHIFT ? 9 * HC * 2 : 0);
  static constexpr int W1_BYTES = KC1 * N * 16;
  static constexpr int OFF_DA = OFF_W1 + W1_BYTES;
  static constexpr int DA_BYTES = 9 * 2 * C * 2, DB_BYTES = 25 * C * 2, W2_BYTES = KC2 * N * 16;
  static constexpr int WT2_BYTES = DA_BYTES + DB_BYTES + W2_BYTES;
  // plane pitches: A1 is TMA-dense, the others carry one pad vector (bank spread of the per-plane accesses)
  static constexpr int PA1 = M1 * 16, P1 = (M1 + 1) * 16, P2 = (M2 + 1) * 16, P3 = (M3 + 1) * 16;
  // shared memory map: [X | A2 | A1 ...... ] with G1 aliasing [A2 | A1], then [GATED / z staging | WT2 | W1]
  static constexpr int X_BAR = 0, X_TMEM = 32, X_RED = 256, X_BYTES = 256 + 16 * 32 * 4;
  static constexpr int S_A2 = X_BYTES;                                // GEMM2 operand (written by the dw5x5 stage)
  static constexpr int A2_BYTES = KC2 * P3;
  static constexpr int S_G1 = S_A2;                                   // 2C-wide fp16 tensor of the region (drain .. dw3x3)
  static constexpr int G1_BYTES = NC * P1;
  static constexpr int S_A1 = S_A2 + A2_BYTES;                        // TMA destination = GEMM1 operand of the NEXT tile
  static constexpr int A1_BYTES = KC1 * PA1;
  static constexpr int LO_END = (S_G1 + G1_BYTES > S_A1 + A1_BYTES ? S_G1 + G1_BYTES : S_A1 + A1_BYTES);
  static constexpr int S_GT = (LO_END + 1023) / 1024 * 1024;
  static constexpr int GT_BYTES = KC2 * P2;
  // z staging tile (GATED is dead after the dw5x5): 256 pixels x 128 bytes, pixel-major with the 128-byte TMA swizzle
  // (16-byte chunk c of pixel p sits at chunk c ^ (p & 7): lane = pixel writes are bank-conflict free) -> ONE TMA tile store
  static constexpr int S_Z = S_GT;
  static constexpr int Z_BYTES = M3 * C * 2;
  static constexpr int S_WT2 = (S_GT + GT_BYTES + 127) / 128 * 128;
  static constexpr int S_W1 = (S_WT2 + WT2_BYTES + 127) / 128 * 128;
  static constexpr int SMEM = S_W1 + W1_BYTES;
  static_assert(S_A1 % 128 == 0, "TMA destination alignment");
  static_assert(Z_BYTES <= GT_BYTES && S_Z % 1024 == 0, "z staging fits the GATED area, swizzle-atom aligned");
  static_assert(S_A1 + (KC1 - 1) * PA1 + MT1 * 128 * 16 <= SMEM, "UMMA rows beyond M1 stay inside the allocation");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <int KC1, bool MIDCA>
__global__ void __launch_bounds__(kPreThreads, 1) cab_pass_a_pre_kernel(const GsnCabPassA d, const __grid_constant__ CUtensorMap tm_a1,
                                                                       const __grid_constant__ CUtensorMap tm_z) {
  using K = PreCfg<KC1>;
  constexpr int C = K::C;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // persistent CTAs: linear tile index = (frame * tiles_y + tile_y) * tiles_x + tile_x, stride gridDim.x
  const int tiles_x = (d.W + K::TW - 1) / K::TW, tiles_y = (d.H + K::TH - 1) / K::TH;
  const int total_tiles = tiles_x * tiles_y * d.T;
  int tile = blockIdx.x;
  int t, x0, y0;
  auto decode = [&](int ti, int &ft, int &fx0, int &fy0) {
    ft = ti / (tiles_x * tiles_y);
    const int r = ti - ft * tiles_x * tiles_y, ty = r / tiles_x;
    fy0 = ty * K::TH;
    fx0 = (r - ty * tiles_x) * K::TW;
  };
  decode(tile, t, x0, y0);
  const unsigned char *wb = reinterpret_cast<const unsigned char *>(d.wblob);
  const size_t frame = (size_t)d.H * d.W * C;
  const uint32_t bar_mma = smem_u32(smem + K::X_BAR), bar_in = bar_mma + 8, bar_g1 = bar_mma + 16;
  const uint32_t smem16 = smem_u32(smem) >> 4;   // UMMA descriptors address shared memory in 16-byte units
  // debug_stage == 9: thread 0 records clock64() at the stage boundaries of the CTA's second tile (steady state)
  long long *clk = nullptr;
  const bool clk_second = (int)blockIdx.x + (int)gridDim.x < total_tiles;
  int clk_i = 0;
#define GSN_CLK() do { if (clk && clk_i < 16) clk[clk_i++] = clock64(); } while (0)
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + K::X_TMEM);

  // A1 of one tile: ONE TMA tile load, box {22 px * 8 ch, 22 rows, KC1 planes, 1 frame}, zero fill outside the image
  auto issue_a1 = [&](int ft, int fx0, int fy0) {
    fence_async_proxy();   // earlier generic-proxy accesses of the (G1-aliased) destination stay ordered before the TMA writes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_in), "r"(K::A1_BYTES) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
            "r"(smem_u32(smem + K::S_A1)), "l"(reinterpret_cast<uint64_t>(&tm_a1)), "r"((fx0 - 3) * 8), "r"(fy0 - 3), "r"(0), "r"(ft),
        "r"(bar_in)
        : "memory");
  };
  // Single-thread roles are spread over four warps so that no warp carries all the issue latency:
  //   thread 0: GEMM2 and the A1 TMA load; kIssA: next tile's GEMM1 M tiles 2,3; kIssB: M tiles 0,1; kIssZ: the z TMA store
  constexpr int kIssA = 128, kIssB = 256, kIssZ = 384;
  uint32_t tmem = 0;
  // GEMM1 of M tiles [m0, m1): D[m] (128 x 2C fp32, TMEM columns [m*N, m*N + N)) = A1[m] (128 x CIN) . W1^T; one commit
  auto issue_gemm1 = [&](int m0, int m1) {
    constexpr uint32_t idesc = make_idesc_f16(128, K::N);
    for (int m = m0; m < m1; ++m)
#pragma unroll
      for (int k = 0; k < KC1 / 2; ++k) {
        const uint64_t ad = smem_desc_at(smem16, K::S_A1 + 2 * k * K::PA1 + m * 128 * 16, K::PA1, 128);
        const uint64_t bd = smem_desc_at(smem16, K::S_W1 + 2 * k * (K::N * 16), K::N * 16, 128);
        umma_f16(tmem + m * K::N, ad, bd, idesc, k > 0);
      }
    umma_commit(bar_g1);
  };

  // ---- P0 (once per CTA): barriers, TMEM, all weights; the first tile's A1 and GEMM1 -------------------------------
  if (tid == 0) {
    mbar_init(bar_mma, 1);
    mbar_init(bar_in, 1);
    mbar_init(bar_g1, 2);    // two commits per tile: M tiles {2,3} and {0,1}
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) issue_a1(t, x0, y0);
  for (int i = tid; i < K::W1_BYTES / 16; i += kPreThreads) cp_async16(smem + K::S_W1 + i * 16, wb + K::OFF_W1 + i * 16, true);
  for (int i = tid; i < K::WT2_BYTES / 16; i += kPreThreads) cp_async16(smem + K::S_WT2 + i * 16, wb + K::OFF_DA + i * 16, true);
  cp_async_commit();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  cp_async_wait<0>();
  fence_async_proxy();     // cp.async'd weights -> visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  tmem = *tmem_slot;
  // mbarrier phases: bar_in completes once per A1 load, bar_g1 once per tile (two commits), bar_mma once per GEMM2; every
  // thread tracks all three parities (the waiters of bar_in are the issuing threads only)
  uint32_t in_parity = 0, mma_parity = 0, g1_parity = 0;
  if (tid == 0) {
    mbar_wait(bar_in, in_parity);
    issue_gemm1(2, 4);
    issue_gemm1(0, 2);
  }
  in_parity ^= 1;

  for (;;) {   // ---- tile loop ----
  if (d.debug_stage == 9 && tid == 0 && (clk_second ? tile == (int)blockIdx.x + (int)gridDim.x : true))
    clk = reinterpret_cast<long long *>(d.debug_out) + (size_t)tile * 16;
  GSN_CLK();  // 0: tile start
  const int nt = tile + (int)gridDim.x;
  const bool has_next = nt < total_tiles;
  int nt_t = 0, nt_x0 = 0, nt_y0 = 0;
  if (has_next) {   // advance (frame, tile row, tile column) by gridDim.x tiles without integer divisions
    int tx = x0 / K::TW + (int)gridDim.x, ty = y0 / K::TH;   // TW, TH are powers of two
    nt_t = t;
    while (tx >= tiles_x) { tx -= tiles_x; ++ty; }
    while (ty >= tiles_y) { ty -= tiles_y; ++nt_t; }
    nt_x0 = tx * K::TW; nt_y0 = ty * K::TH;
  }

  // ---- P2: GEMM1 of this tile was issued during the previous tile (or in P0) ------------------------------------------
  mbar_wait(bar_g1, g1_parity);
  g1_parity ^= 1;
  tc_fence_after();
  GSN_CLK();  // 1: GEMM1 complete
  if (d.debug_stage == 1) {   // A1 as the tensor core saw it (tests)
    uint4 *o = reinterpret_cast<uint4 *>(d.debug_out) + (size_t)tile * KC1 * K::M1;
    for (int i = tid; i < KC1 * K::M1; i += kPreThreads)
      o[i] = *reinterpret_cast<uint4 *>(smem + K::S_A1 + (i / K::M1) * K::PA1 + (i % K::M1) * 16);
    __syncthreads();
  }

  // ---- P3: TMEM -> fp16 G1 planes (all 2C channels of the region) ----------------------------------------------
  {
    const int quarter = warp & 3, cg = warp >> 2;   // a warp may only touch TMEM lanes [32*quarter, +32); one 32-column group per warp
    static_assert(K::N / 32 == kPreThreads / 128 && K::MT1 % 2 == 0, "drain: one column group per warp, M tiles in pairs");
    auto put = [&](const uint32_t (&v)[32], int m) {
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
    };
#pragma unroll 1
    for (int m = 0; m < K::MT1; m += 2) {     // two TMEM loads in flight before the first conversion
      uint32_t v0[32], v1[32];
      const uint32_t ta = tmem + ((uint32_t)(quarter * 32) << 16) + m * K::N + cg * 32;
      tmem_ld32_nowait(ta, v0);
      tmem_ld32_nowait(ta + K::N, v1);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      put(v0, m);
      put(v1, m + 1);
    }
    tc_fence_before();
    __syncthreads();   // G1 complete; every thread is done with the previous tile's z staging / channel sums
    GSN_CLK();  // 2: TMEM -> G1 done
  }

  // ---- P4: dw3x3 + id on both halves, SimpleGate -> GATED (zero outside the image) ---------------------------------
  // The identity of RepConv2 / RepConv is folded into the centre tap by host/packing.py (w_c + 1 in fp16).
  {
    constexpr int NSTRIP = 3, SROWS = (K::R2H + NSTRIP - 1) / NSTRIP;  // 7,7,6 output rows
    const unsigned char *wda = smem + K::S_WT2;
    // Branch-free strips: every strip runs SROWS output rows (the 7th row of the last strip reads one region row past the
    // plane -- still inside the allocation -- and its store is predicated off); the zero padding outside the image is a
    // bit mask on the packed result.  No control flow inside the row loop, so the loads of the next rows are scheduled under
    // the HFMA2s of the current one.
    static_assert(K::S_G1 + (K::NC - 1) * K::P1 + ((NSTRIP * SROWS + 2) * K::R1W + 2) * 16 <= K::SMEM, "over-read stays inside smem");
    for (int item = tid; item < K::KC2 * K::R2W * NSTRIP; item += kPreThreads) {
      const int x = item % K::R2W, rest = item / K::R2W;
      const int p = rest % K::KC2, strip = rest / K::KC2;
      const int r0 = strip * SROWS;
      const int gx = x0 - 2 + x;
      const bool xok = gx >= 0 && gx < d.W;
      H8 res[SROWS];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int chunk = half * K::KC2 + p;
        const unsigned char *pl = smem + K::S_G1 + chunk * K::P1 + (r0 * K::R1W + x) * 16;
        H8 w[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) w[i] = lds_h8(wda + (i * 2 * C + chunk * 8) * 2);
        H8 acc0, acc1;
#pragma unroll
        for (int i = 0; i < SROWS + 2; ++i) {           // input region row r0 + i feeds output rows (r0+i-2 .. r0+i)
          const unsigned char *rp = pl + i * K::R1W * 16;
          const H8 v0 = lds_h8(rp), v1 = lds_h8(rp + 16), v2 = lds_h8(rp + 32);
          if (i >= 2) {                                 // output row r0+i-2 completes with kernel row 2
            h8_fma(acc0, v0, w[6]); h8_fma(acc0, v1, w[7]); h8_fma(acc0, v2, w[8]);
            if (half == 0) res[i - 2] = acc0;
            else {
              const int orow = r0 + i - 2;
              const int gy = y0 - 2 + orow;
              const uint32_t mask = (xok && gy >= 0 && gy < d.H) ? 0xffffffffu : 0u;
              H8 o;
              h8_mul(o, res[i - 2], acc0);
              uint4 ov;
              ov.x = *reinterpret_cast<const uint32_t *>(&o.h[0]) & mask;
              ov.y = *reinterpret_cast<const uint32_t *>(&o.h[1]) & mask;
              ov.z = *reinterpret_cast<const uint32_t *>(&o.h[2]) & mask;
              ov.w = *reinterpret_cast<const uint32_t *>(&o.h[3]) & mask;
              if (orow < K::R2H) *reinterpret_cast<uint4 *>(smem + K::S_GT + p * K::P2 + (orow * K::R2W + x) * 16) = ov;
            }
          }
          if (i >= 1 && i <= SROWS) {                   // output row r0+i-1: kernel row 1 (centre row)
            acc0 = acc1;
            h8_fma(acc0, v0, w[3]); h8_fma(acc0, v1, w[4]); h8_fma(acc0, v2, w[5]);   // w[4] carries the "+ x" of RepConv2
          }
          if (i < SROWS) {                              // output row r0+i: kernel row 0 starts a new accumulator
            h8_mul(acc1, v0, w[0]);
            h8_fma(acc1, v1, w[1]); h8_fma(acc1, v2, w[2]);
          }
        }
      }
    }
    __syncthreads();
    GSN_CLK();  // 3: dwA done
    // G1 (and with it the A1 area it aliases) is dead: bring in the next tile's A1
    if (has_next && tid == 0) issue_a1(nt_t, nt_x0, nt_y0);
  }
  if (d.debug_stage == 2) {
    uint4 *o = reinterpret_cast<uint4 *>(d.debug_out) + (size_t)tile * K::KC2 * K::M2;
    for (int i = tid; i < K::KC2 * K::M2; i += kPreThreads)
      o[i] = *reinterpret_cast<uint4 *>(smem + K::S_GT + (i / K::M2) * K::P2 + (i % K::M2) * 16);
  }

  // ---- P5: dw5x5 (+ merged dw3x3) + id on the gated tensor -> A2 (GEMM2 operand) -----------------------------------
  {
    constexpr int NSTRIP = 2, SROWS = K::TH / NSTRIP;   // 8 output rows per strip
    const unsigned char *wdb = smem + K::S_WT2 + K::DA_BYTES;
    for (int item = tid; item < 2 * K::KC2 * K::TW * NSTRIP; item += kPreThreads) {
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
            for (int ee = 0; ee < 2; ++ee)
              acc[s][ee] = (ky == 0 && tx == 0) ? __hmul2(v[tx][ee], w[ky * 5 + tx][ee]) : __hfma2(v[tx][ee], w[ky * 5 + tx][ee], acc[s][ee]);
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
    fence_async_proxy();   // generic-proxy writes of A2 -> visible to the tensor core
    __syncthreads();
    GSN_CLK();  // 4: dwB done
  }
  if (d.debug_stage == 3) {
    uint4 *o = reinterpret_cast<uint4 *>(d.debug_out) + (size_t)tile * K::KC2 * K::M3;
    for (int i = tid; i < K::KC2 * K::M3; i += kPreThreads)
      o[i] = *reinterpret_cast<uint4 *>(smem + K::S_A2 + (i / K::M3) * K::P3 + (i % K::M3) * 16);
  }

  if (MIDCA) {
    // denoise variants: stop here.  u = RepConv(gate) goes to HBM, the mid CALayer2 needs the frame mean of the gated
    // tensor (sums over this tile's 16x16 centre of GATED); the scale it produces is folded into W2 by cab_fold_mid.
    __half *ug = reinterpret_cast<__half *>(d.z) + (size_t)t * frame;
    for (int i = tid; i < K::M3 * K::KC2; i += kPreThreads) {
      const int ch = i % K::KC2, p = i / K::KC2;
      const int gy = y0 + p / K::TW, gx = x0 + (p % K::TW);
      if (gy < d.H && gx < d.W)
        *reinterpret_cast<uint4 *>(ug + ((size_t)gy * d.W + gx) * C + ch * 8) = *reinterpret_cast<const uint4 *>(smem + K::S_A2 + ch * K::P3 + p * 16);
    }
    float *red = reinterpret_cast<float *>(smem + K::X_RED);
    for (int u = warp; u < K::KC2 * 2; u += kPreThreads / 32) {
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
    tc_fence_before();
    __syncthreads();       // A2 / GATED reads are done
    if (has_next && tid == kIssA) {   // TMEM is idle since the drain: the whole GEMM1 of the next tile
      tc_fence_after();
      mbar_wait(bar_in, in_parity);
      issue_gemm1(2, 4);
      issue_gemm1(0, 2);
    }
    if (tid < C) d.chan_partial[(size_t)tile * C + tid] = red[tid] + red[C + tid];
    GSN_CLK();
  } else {

  // ---- P6: GEMM2 on the tensor core: (256 x C) . W2^T -> TMEM columns [0, 2*N); then the next tile's M tiles 2,3 ------
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
    umma_commit(bar_mma);
  } else if (has_next && tid == kIssA) {
    // TMEM columns [2N, 4N) were drained in P3: the next tile's M tiles 2,3 run right behind GEMM2 (its A1 was requested
    // before the dw5x5 stage)
    tc_fence_after();
    mbar_wait(bar_in, in_parity);
    issue_gemm1(2, 4);
  }
  mbar_wait(bar_mma, mma_parity);
  mma_parity ^= 1;
  tc_fence_after();
  GSN_CLK();  // 5: GEMM2 done

  // ---- P7: a * sigmoid(b) (SimpleGate2) -> z tile (fp16 planes) -> coalesced global store; per-tile channel sums --------
  {
    const int quarter = warp & 3, sub = warp >> 2;
    float *red = reinterpret_cast<float *>(smem + K::X_RED);   // [16 warps][32 channels]
    static_assert(K::MT3 * (C / 32) == kPreThreads / 128, "one (M tile, 32-channel group) unit per warp");
    {
      const int m = sub / (C / 32), cg = sub % (C / 32);
      uint32_t a[32], b[32];
      const uint32_t base = tmem + ((uint32_t)(quarter * 32) << 16) + m * K::N + cg * 32;
      tmem_ld32_nowait(base, a);
      tmem_ld32_nowait(base + C, b);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      // Both halves of GEMM2's output now sit in registers: TMEM columns [0, 2N) are free again.  Every warp signals that on
      // named barrier 1 without waiting; only the warp of thread kIssB waits for all 16 and then issues the rest of the next
      // tile's GEMM1 (M tiles 0,1), which runs under the sigmoid gate and the store.
      tc_fence_before();
      if (has_next && warp == kIssB / 32) {
        asm volatile("bar.sync 1, %0;\n" ::"n"(kPreThreads) : "memory");
        if (tid == kIssB) {
          tc_fence_after();
          mbar_wait(bar_in, in_parity);
          issue_gemm1(0, 2);
        }
        __syncwarp();
      } else if (has_next) {
        asm volatile("bar.arrive 1, %0;\n" ::"n"(kPreThreads) : "memory");
      }
      const int px = m * 128 + quarter * 32 + lane;
      const bool valid = (y0 + px / K::TW < d.H) && (x0 + px % K::TW < d.W);
      float z[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) z[i] = __uint_as_float(a[i]) * sigmoid_tanh(__uint_as_float(b[i]));
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4)   // pixel-major 128-byte rows, 16-byte chunks XOR-swizzled with the pixel index (TMA SWIZZLE_128B)
        *reinterpret_cast<uint4 *>(smem + K::S_Z + px * 128 + (((cg * 4 + c4) ^ (px & 7)) << 4)) = pack8(*reinterpret_cast<float(*)[8]>(&z[c4 * 8]));
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
    fence_async_proxy();   // staging writes -> visible to the TMA store
    tc_fence_before();
    __syncthreads();
    GSN_CLK();  // 6: gate2 -> z tile done
    if (tid == kIssZ) {    // z tile -> HBM: one TMA tile store (pixels beyond the image border are clipped by the hardware)
      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];\n" ::
                       "l"(reinterpret_cast<uint64_t>(&tm_z)), "r"(0), "r"(x0), "r"(y0), "r"(t), "r"(smem_u32(smem + K::S_Z))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
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
  GSN_CLK();  // 7: stores + sums done
  }   // !MIDCA
    clk = nullptr;
    if (!has_next) break;
    in_parity ^= 1;
    tile = nt;
    t = nt_t; x0 = nt_x0; y0 = nt_y0;
  }   // tile loop
#undef GSN_CLK
  if (tid == kIssZ) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512));
  }
}

// ---- producer: LayerNorm over the CIN channels of every pixel -> k-chunk planar A1 ------------------------------------
// a1[t][chunk][pixel][8] = LN([rolled stream | hw_pre])[pixel][chunk*8 .. +8]   (CAB1: LN(x)); gshift_deblur2.py:19-28,44-53.
// One quad of lanes per pixel (lane j owns chunks j, j+4 (, j+8)), fp32 statistics, same arithmetic as the in-kernel LayerNorm
// of cab_pass_a_tc_kernel.
template <bool SHIFT>
__global__ void __launch_bounds__(256) ln_planar_kernel(const __half *__restrict__ x, const __half *__restrict__ hw_pre, int T,
                                                        long long hw, int mode, int circular, const float *__restrict__ ln,
                                                        __half *__restrict__ a1) {
  constexpr int C = 64, HC = 32, CIN = SHIFT ? 96 : 64, NV = CIN / 4, KC = CIN / 8;
  const int tid = threadIdx.x, j = tid & 3;
  const int t = blockIdx.y;
  const long long q = (long long)blockIdx.x * 64 + (tid >> 2);
  const bool live = q < hw;
  const RollSrc rs = roll_source(mode, circular, t, T, C);
  int chunk_of[NV / 8];
#pragma unroll
  for (int k = 0; k < NV / 8; ++k) chunk_of[k] = 4 * k + j;
  uint4 raw[NV / 8];
#pragma unroll
  for (int k = 0; k < NV / 8; ++k) raw[k] = make_uint4(0, 0, 0, 0);
  if (live) {
    const size_t frame = (size_t)hw * C;
    if (SHIFT) {
      raw[0] = __ldg(reinterpret_cast<const uint4 *>(x + rs.f_lo * frame + (size_t)q * C + rs.c_lo + j * 8));
      raw[1] = __ldg(reinterpret_cast<const uint4 *>(x + rs.f_hi * frame + (size_t)q * C + rs.c_hi + j * 8));
      raw[NV / 8 - 1] = __ldg(reinterpret_cast<const uint4 *>(hw_pre + ((size_t)t * hw + q) * HC + j * 8));
    } else {
      raw[0] = __ldg(reinterpret_cast<const uint4 *>(x + (size_t)t * frame + (size_t)q * C + j * 8));
      raw[1] = __ldg(reinterpret_cast<const uint4 *>(x + (size_t)t * frame + (size_t)q * C + (j + 4) * 8));
    }
  }
  float v[NV];
#pragma unroll
  for (int k = 0; k < NV / 8; ++k) unpack8(raw[k], *reinterpret_cast<float(*)[8]>(&v[k * 8]));
  float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < NV; ++i) { s4[i & 3] += v[i]; q4[i & 3] = fmaf(v[i], v[i], q4[i & 3]); }
  float s = (s4[0] + s4[1]) + (s4[2] + s4[3]), ss = (q4[0] + q4[1]) + (q4[2] + q4[3]);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
  const float mu = s * (1.f / CIN);
  const float rstd = rsqrtf(fmaxf(ss * (1.f / CIN) - mu * mu, 0.f) + 1e-6f);
  const float nmr = -mu * rstd;
  if (!live) return;
#pragma unroll
  for (int k = 0; k < NV / 8; ++k) {
    const float4 g0 = __ldg(reinterpret_cast<const float4 *>(ln + chunk_of[k] * 8)), g1 = __ldg(reinterpret_cast<const float4 *>(ln + chunk_of[k] * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(ln + CIN + chunk_of[k] * 8)), b1 = __ldg(reinterpret_cast<const float4 *>(ln + CIN + chunk_of[k] * 8 + 4));
    const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bet[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(fmaf(v[k * 8 + i], rstd, nmr), gam[i], bet[i]);
    *reinterpret_cast<uint4 *>(a1 + (((size_t)t * KC + chunk_of[k]) * hw + q) * 8) = pack8(o);
  }
}

template <int KC1, bool MIDCA>
static int launch_pass_a_pre(const GsnCabPassA &d, cudaStream_t st, const CUtensorMap &tm, const CUtensorMap &tm_z) {
  using K = PreCfg<KC1>;
  GSN_ONCE_PER_DEVICE(
    cudaFuncSetAttribute(cab_pass_a_pre_kernel<KC1, MIDCA>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
  const long long total = (long long)((d.W + K::TW - 1) / K::TW) * ((d.H + K::TH - 1) / K::TH) * d.T;
  const int num_sms = sm_count();
  const unsigned grid = (unsigned)(total < num_sms ? total : num_sms);
  cab_pass_a_pre_kernel<KC1, MIDCA><<<grid, kPreThreads, K::SMEM, st>>>(d, tm, tm_z);
  count_launch();
  return check_launch("cab_pass_a_pre");
}

// d.a1_pre != NULL: the pre-normalised path
int cab_pass_a_pre_dispatch(const GsnCabPassA &d, cudaStream_t st) {
  if (d.C != 64) {
    set_error("cab_pass_a (a1_pre): C=%d unsupported (64)", d.C);
    return GSN_E_UNSUPPORTED;
  }
  const bool shift = d.mode != GSN_MODE_CAB1;
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  if (!encode_tmap_planar(&tm, d.a1_pre, d.W, d.H, shift ? 12 : 8, d.T, 22, 22)) {
    set_error("cab_pass_a (a1_pre): cuTensorMapEncodeTiled failed (W=%d H=%d T=%d)", d.W, d.H, d.T);
    return GSN_E_CUDA;
  }
  CUtensorMap tm_z;     // z (T,H,W,64) NHWC, one 16x16-pixel tile per store, 128-byte swizzle
  memset(&tm_z, 0, sizeof(tm_z));
  if (!d.mid_ca && !encode_tmap_nhwc(&tm_z, d.z, 64, d.W, d.H, d.T, 64, 16, 16, true)) {
    set_error("cab_pass_a (a1_pre): cuTensorMapEncodeTiled(z) failed (W=%d H=%d T=%d)", d.W, d.H, d.T);
    return GSN_E_CUDA;
  }
  if (shift) return d.mid_ca ? launch_pass_a_pre<12, true>(d, st, tm, tm_z) : launch_pass_a_pre<12, false>(d, st, tm, tm_z);
  return d.mid_ca ? launch_pass_a_pre<8, true>(d, st, tm, tm_z) : launch_pass_a_pre<8, false>(d, st, tm, tm_z);
}

}  // namespace gsn

extern "C" int gsn_ln_planar(const void *x, const void *hw_pre, int T, int H, int W, int C, int mode, int circular,
                             const float *ln, void *a1, void *stream) {
  using namespace gsn;
  GSN_REQUIRE(x && ln && a1, "ln_planar: null pointer");
  GSN_REQUIRE(T > 0 && H > 0 && W > 0, "ln_planar: empty shape");
  GSN_REQUIRE(mode == GSN_MODE_CAB1 || hw_pre, "ln_planar: CAB2 modes need the gsn_shift_conv1 output");
  if (C != 64) { set_error("ln_planar: C=%d unsupported (64)", C); return GSN_E_UNSUPPORTED; }
  const long long hw = (long long)H * W;
  dim3 grid((unsigned)((hw + 63) / 64), T);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __half *xh = reinterpret_cast<const __half *>(x), *hp = reinterpret_cast<const __half *>(hw_pre);
  if (mode == GSN_MODE_CAB1) ln_planar_kernel<false><<<grid, 256, 0, st>>>(xh, hp, T, hw, mode, circular, ln, reinterpret_cast<__half *>(a1));
  else ln_planar_kernel<true><<<grid, 256, 0, st>>>(xh, hp, T, hw, mode, circular, ln, reinterpret_cast<__half *>(a1));
  count_launch();
  return check_launch("ln_planar");
}
