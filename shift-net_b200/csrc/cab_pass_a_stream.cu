// Pass A of the fused shift + NAF block, ROW-STREAMING warp-specialised variant (sm_100a, C = 64, deblur nets).
//
// Same arithmetic as cab_pass_a_pre_kernel (gshift_deblur2.py:186-258, from the first 1x1 to SimpleGate2, LayerNorm'd operand
// precomputed in the k-chunk planar layout), restructured so that the FMA-bound depthwise stages never wait for anything else:
//
//   * A CTA walks DOWN a 26-pixel-wide column strip of a frame, 4 region rows (= one 128-row UMMA M tile of 4 x 32 pixels) per
//     step.  The depthwise 3x3 and 5x5 stages keep sliding accumulators in registers, so there is NO vertical halo recompute
//     (the 16x16-tile kernel ran GEMM1 on 1.89x and the dw3x3 on 1.56x the pixels; here 1.23x / 1.15x, horizontal only) and the
//     intermediates live in small row rings instead of whole-tile buffers.
//   * Warp roles, all running concurrently on different rows of the stream and handing off through mbarrier rings (every wait
//     is a hardware-suspended mbarrier.try_wait, nobody polls):
//       warps 0-7   dwA : RepConv2 (dw3x3 + id) on both halves + SimpleGate, one warp per chunk pair; lanes = (half, pixel pair),
//                         the two halves of a pixel meet through one shuffle; depthwise taps live in registers for the whole kernel
//       warps 8-15  dwB : RepConv (merged 5x5) on the gated tensor, one warp per chunk; lanes = (4-channel half chunk, pixel pair),
//                         25 taps in registers, five sliding accumulator rows; writes GEMM2's A operand; the warp that
//                         completes a block issues GEMM2
//       warps 16-19 aux : one warp per TMEM lane quarter = per row of the step: drains GEMM1's accumulators to the fp16 row ring;
//                         lane 0 of warp 19 also issues the TMA loads of the operand rows (two steps ahead) and GEMM1 (one step
//                         ahead) at the points of its loop where their inputs are known to be free
//       gate stage      : SimpleGate2 on GEMM2's accumulators a few steps later (z -> swizzled staging -> TMA store by the last
//                         task to finish, plus the deterministic per-piece channel sums of z) is split into 12 tasks per block:
//                         each aux warp takes half the channels of its row, the two dwA warps of that lane quarter a quarter each
//     GEMM1 / GEMM2 are tcgen05.mma (M=128, N=128, K=16) with fp32 accumulators in TMEM (2 + 2 slots of 128 columns).
//     20 warps = 5 per scheduler = 96 registers per thread (a 21st warp would cost every thread 16 registers).
//   * Work is cut into equal contiguous runs of 4-row blocks per CTA (a run may span several strips / frames): no tail wave.
//     Each piece of a strip costs two warm-up steps (the 3 + 2 rows of vertical context of the two stencils).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace gsn {

constexpr int kStWarps = 20, kStThreads = kStWarps * 32;

struct StGeom {
  int nstrips, nblk_col, ppc, ntiles;   // strips per frame, 4-row blocks per strip, partial-sum slots per strip, slots per frame
  long long chunk, total;               // blocks per CTA, blocks in the launch
};

template <int KC1>
struct StCfg {
  static constexpr int C = 64, N = 128, KC2 = 8, NC = 16, HC = 32;
  static constexpr int TW = 26, RW = 32, ROWS = 4, M = 128;
  static constexpr bool SHIFT = KC1 == 12;
  static constexpr int CIN = KC1 * 8;
  // weight blob offsets (host/packing.py pack_cab_pass_a; identical to PreCfg)
  static constexpr int OFF_C1 = 2 * CIN * 4;
  static constexpr int OFF_W1 = OFF_C1 + (SHIFT ? 9 * HC * 2 : 0);
  static constexpr int W1_BYTES = KC1 * N * 16;
  static constexpr int OFF_DA = OFF_W1 + W1_BYTES;
  static constexpr int DA_BYTES = 9 * 2 * C * 2, DB_BYTES = 25 * C * 2, W2_BYTES = KC2 * N * 16;
  static constexpr int OFF_DB = OFF_DA + DA_BYTES, OFF_W2 = OFF_DB + DB_BYTES;
  // rings
  static constexpr int NA1 = 2, A1_STAGE = KC1 * M * 16;     // operand rows of one step, k-chunk planar = K-major UMMA layout
  // the two row rings hold three PAIRS of rows each; producers and consumers hand over one pair per mbarrier round trip
  static constexpr int NPAIR = 3;
  static constexpr int NG1 = 2 * NPAIR, G1_ROW = NC * RW * 16;   // fp16 2C tensor, one region row: [16 planes][32 slots][16 B]
  static constexpr int NGT = 2 * NPAIR, GT_ROW = KC2 * 32 * 16;  // gated tensor, one row: [8 planes][32 slots][16 B]
  static constexpr int A2_BLK = KC2 * M * 16;                // GEMM2 operand of one step: [8 planes][4 rows x 32 px][16 B]
  static constexpr int Z_BYTES = ROWS * TW * 128;            // z staging: [4 rows][26 px][128 B], 128-byte swizzle
  static constexpr int Z_STRIDE = (Z_BYTES + 1023) / 1024 * 1024;
  // shared memory map
  static constexpr int S_BAR = 0, S_TMEM = 640, S_CNT = 656, S_RED = 704, X_BYTES = 2048;
  static constexpr int S_Z = X_BYTES;
  static constexpr int S_A1 = S_Z + 2 * Z_STRIDE;
  static constexpr int S_A2 = S_A1 + NA1 * A1_STAGE;
  static constexpr int S_G1 = S_A2 + 2 * A2_BLK;
  static constexpr int S_GT = S_G1 + NG1 * G1_ROW;
  static constexpr int S_W1 = S_GT + NGT * GT_ROW;
  static constexpr int S_W2 = S_W1 + W1_BYTES;
  static constexpr int SMEM = S_W2 + W2_BYTES;
  static_assert(S_Z % 1024 == 0 && Z_STRIDE % 1024 == 0 && S_A1 % 128 == 0, "swizzle atom / TMA destination alignment");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// mbarrier indices
enum : int { B_A1F = 0, B_A1E = 2, B_G1F = 4, B_G1E = 6, B_R1F = 8, B_R1E = 11, B_A2E = 14, B_G2F = 16, B_G2E = 18, B_ZFREE = 20,
             B_RFREE = 22, B_GTF = 23, B_GTE = 23 + 24, B_COUNT = 23 + 48 };   // GTF / GTE: [8 planes][3 row pairs]
// shared counters ("who arrived last"): A2 blocks, z staging buffers, channel-sum buffer
enum : int { N_A2 = 0, N_Z = 2, N_RED = 4 };
static_assert(B_COUNT * 8 <= 640, "barrier area");

// The contiguous run of 4-row blocks of one CTA, cut into pieces at strip boundaries.  Every role walks the same sequence.
struct StSched {
  int nblk_col;
  long long cur, end, chunk;
  int col, off, len, pidx;      // current piece: strip column (frame * nstrips + strip), first block, blocks, slot within the strip
  __device__ __forceinline__ void init(const StGeom &g, int cta) {
    nblk_col = g.nblk_col;
    chunk = g.chunk;
    cur = (long long)cta * g.chunk;
    end = cur + g.chunk;
    if (end > g.total) end = g.total;
    col = off = len = pidx = 0;
  }
  __device__ __forceinline__ bool next() {
    if (cur >= end) return false;
    col = (int)(cur / nblk_col);
    const long long c0 = (long long)col * nblk_col;
    off = (int)(cur - c0);
    const long long rem = end - cur;
    len = nblk_col - off;
    if (len > rem) len = (int)rem;
    pidx = (int)(cur / chunk - c0 / chunk);
    cur += len;
    return true;
  }
};
// a cursor over the CTA's step sequence (each piece = len + 2 steps)
struct StCur {
  StSched s;
  int k, n, t, sx, nstrips;     // step within the piece, steps of the piece, frame and strip of the piece
  bool live;
  __device__ __forceinline__ void piece() {
    live = s.next();
    k = 0;
    n = live ? s.len + 2 : 0;
    t = s.col / nstrips;
    sx = s.col - t * nstrips;
  }
  __device__ __forceinline__ void start(const StSched &s0, int nstrips_) {
    s = s0;
    nstrips = nstrips_;
    piece();
  }
  __device__ __forceinline__ void step() {
    if (++k == n) piece();
  }
};

__device__ __forceinline__ H8 ldg_h8(const unsigned char *p) {
  const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
  H8 r;
  r.h[0] = *reinterpret_cast<const __half2 *>(&v.x);
  r.h[1] = *reinterpret_cast<const __half2 *>(&v.y);
  r.h[2] = *reinterpret_cast<const __half2 *>(&v.z);
  r.h[3] = *reinterpret_cast<const __half2 *>(&v.w);
  return r;
}
__device__ __forceinline__ uint32_t h2u(const __half2 &h) { return *reinterpret_cast<const uint32_t *>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }
// hot-loop wait: hardware-suspended try_wait in a bare loop (no spin counter, no extra control flow)
__device__ __forceinline__ void mbar_wait_nt(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// "am I the last of n to arrive?" on a shared counter; acq_rel so that the last arriver sees everybody's earlier writes
__device__ __forceinline__ bool last_arrival(uint32_t cnt_addr, uint32_t n) {
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared::cta.inc.u32 %0, [%1], %2;\n" : "=r"(old) : "r"(cnt_addr), "r"(n - 1) : "memory");
  return old == n - 1;      // inc wraps to 0 at n-1: the counter is ready for its next use
}

// One warp's share of the gate stage of a block (4 output rows x 26 pixels): SimpleGate2 z = a * sigmoid(b) on GEMM2's accumulators
// of TMEM lane quarter gq (= output row gq of the block) for NCH channel chunks starting at chunk c0 -> fp16 z in the swizzled
// staging tile; the per-piece channel sums of those channels; the last of the 12 tasks of a block issues the TMA store.
template <int KC1, int NCH>
struct GateTask {
  using K = StCfg<KC1>;
  StCur cz;            // lagging cursor over the CTA's steps
  uint32_t v, np;      // valid blocks gated, pieces flushed
  int gq, c0, sum_px0, sum_off;
  float ps[8];         // lane (kq = lane & 7, cq = lane >> 3 < NCH): sums of channels 8 (c0 + cq) + 0..7 over its pixels of the piece

  __device__ __forceinline__ void init(const StSched &sc, const StGeom &geo, int gq_, int c0_, int lane) {
    cz.start(sc, geo.nstrips);
    v = np = 0;
    gq = gq_;
    c0 = c0_;
#pragma unroll
    for (int e = 0; e < 8; ++e) ps[e] = 0.f;
    // sums read the staging tile back: lane (kq, cq) reads the pixels of row gq whose staging index is == kq mod 8 (three or four
    // of the row), so its swizzled chunk position is a per-lane constant and a quarter warp hits 8 different 16-byte columns
    sum_px0 = ((lane & 7) - gq * K::TW) & 7;
    sum_off = (gq * K::TW + sum_px0) * 128 + (((c0 + (lane >> 3)) ^ (lane & 7)) << 4);
  }

  __device__ __forceinline__ void run(unsigned char *smem, uint32_t sbase, uint32_t tmem, const GsnCabPassA &d, const StGeom &geo,
                                      const CUtensorMap *tm_z, int lane) {
    constexpr int C = K::C;
    if (!cz.live) return;
    if (cz.k >= 2) {
      const uint32_t bar0 = sbase + K::S_BAR;
      const int t = cz.t, sx = cz.sx;
      const int x0 = sx * K::TW, yb = (cz.s.off + cz.k - 2) * 4;
      const uint32_t zs = v & 1, zph = (v >> 1) & 1;
      mbar_wait_nt(bar0 + 8 * (B_G2F + zs), zph);
      tc_fence_after();
      mbar_wait_nt(bar0 + 8 * (B_ZFREE + zs), zph ^ 1);     // the store of two blocks ago has read this staging buffer
      unsigned char *zt = smem + K::S_Z + zs * K::Z_STRIDE;
      const uint32_t tlane = (uint32_t)(gq * 32) << 16;
      const int pz = gq * K::TW + lane;
#pragma unroll
      for (int grp = 0; grp < NCH; ++grp) {
        uint32_t a[8], b[8];
        const uint32_t ta = tmem + tlane + 256 + zs * K::N + (c0 + grp) * 8;
        tmem_ld8_nowait(ta, a);
        tmem_ld8_nowait(ta + C, b);
        tmem_ld_wait();
        float z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float th;
          asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(__uint_as_float(b[i])));
          z[i] = fmaf(__uint_as_float(a[i]), th, __uint_as_float(a[i]));     // W2 is held at half scale: (a/2) tanh(b/2) + a/2
        }
        if (lane < K::TW) *reinterpret_cast<uint4 *>(zt + pz * 128 + (((c0 + grp) ^ (pz & 7)) << 4)) = pack8(z);
      }
      tc_fence_before();
      __syncwarp();        // also: this warp's staging writes are visible to its own lanes
      if (lane == 0) mbar_arrive(bar0 + 8 * (B_G2E + zs));   // the accumulator slot may take the GEMM2 of two blocks ahead
      if (yb + gq < d.H && (lane >> 3) < NCH) {              // channel sums over the valid pixels (fp16 values, as pass B reads them)
        int nvx = d.W - x0;
        nvx = nvx > K::TW ? K::TW : nvx;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (sum_px0 + 8 * i < nvx) {
            float f[8];
            unpack8(*reinterpret_cast<const uint4 *>(zt + sum_off + i * 1024), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) ps[e] += f[e];
          }
        }
      }
      fence_async_proxy();   // staging writes -> visible to the TMA store
      __syncwarp();
      bool last = false;
      if (lane == 0) last = last_arrival(sbase + K::S_CNT + 4 * (N_Z + zs), 12);
      if (last) {            // the twelfth task stores the tile; rows / columns beyond the image are clipped by the hardware
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];\n" ::
                         "l"(reinterpret_cast<uint64_t>(tm_z)), "r"(0), "r"(x0), "r"(yb), "r"(t), "r"(sbase + K::S_Z + zs * K::Z_STRIDE)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        mbar_arrive(bar0 + 8 * (B_ZFREE + zs));
      }
      __syncwarp();
      ++v;
      if (cz.k == cz.n - 1) {
        // last block of the piece: its channel sums (fixed order => deterministic) -> chan_partial[t][strip * ppc + pidx][C]
        float *red = reinterpret_cast<float *>(smem + K::S_RED);   // [4 rows][64 channels]
#pragma unroll
        for (int off = 1; off < 8; off <<= 1)
#pragma unroll
          for (int e = 0; e < 8; ++e) ps[e] += __shfl_xor_sync(0xffffffffu, ps[e], off);
        mbar_wait(bar0 + 8 * B_RFREE, (np & 1) ^ 1);        // the previous piece's sums have left the buffer
        if ((lane & 7) == 0 && (lane >> 3) < NCH) {
          float *rp8 = red + gq * C + (c0 + (lane >> 3)) * 8;
          *reinterpret_cast<float4 *>(rp8) = make_float4(ps[0], ps[1], ps[2], ps[3]);
          *reinterpret_cast<float4 *>(rp8 + 4) = make_float4(ps[4], ps[5], ps[6], ps[7]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) ps[e] = 0.f;
        __syncwarp();
        bool lastr = false;
        if (lane == 0) lastr = last_arrival(sbase + K::S_CNT + 4 * N_RED, 12);
        lastr = __shfl_sync(0xffffffffu, lastr ? 1 : 0, 0) != 0;
        if (lastr) {
          const size_t slot = (size_t)t * geo.ntiles + (size_t)sx * geo.ppc + cz.s.pidx;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int ch = lane + 32 * hh;
            d.chan_partial[slot * C + ch] = (red[ch] + red[C + ch]) + (red[2 * C + ch] + red[3 * C + ch]);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar0 + 8 * B_RFREE);
        }
        ++np;
      }
    }
    cz.step();
  }
};

template <int KC1>
__global__ void __launch_bounds__(kStThreads, 1) cab_pass_a_stream_kernel(const GsnCabPassA d, const __grid_constant__ CUtensorMap tm_a1,
                                                                         const __grid_constant__ CUtensorMap tm_z, const StGeom geo) {
  using K = StCfg<KC1>;
  constexpr int C = K::C;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t smem16 = sbase >> 4;
  const uint32_t bar0 = sbase + K::S_BAR;
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  auto cnt = [&](int i) { return sbase + K::S_CNT + 4u * (uint32_t)i; };
  const unsigned char *wb = reinterpret_cast<const unsigned char *>(d.wblob);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + K::S_TMEM);
  constexpr uint32_t idesc = make_idesc_f16(128, K::N);
  // The gate stage (SimpleGate2 on GEMM2's accumulators) of a block is shared: the aux warp of its row takes channel chunks 0..3,
  // the two dw3x3 warps of that TMEM lane quarter chunks 4..5 and 6..7 (measured: all of it on the dw3x3 warps 0.70 ms per launch,
  // all of it on the aux warps 0.73 ms -- either way the role that carried it was the critical path).  It runs LAG steps behind
  // the role's own position: the drain is at most 1.5 steps (the row ring) ahead of the dw3x3 warps, those at most 1.5 steps ahead
  // of the dw5x5 warps whose last row triggers GEMM2.
  constexpr uint32_t LAG_D = 3, LAG_A = 2;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_A1F + i), 1);   // TMA bytes of the stage
      mbar_init(bar(B_A1E + i), 1);   // tcgen05.commit after GEMM1
      mbar_init(bar(B_G1F + i), 1);   // tcgen05.commit after GEMM1
      mbar_init(bar(B_G1E + i), 4);   // the four drain warps
      mbar_init(bar(B_A2E + i), 1);   // tcgen05.commit after GEMM2 (or a plain arrive for warm-up blocks)
      mbar_init(bar(B_G2F + i), 1);   // tcgen05.commit after GEMM2
      mbar_init(bar(B_G2E + i), 12);  // the twelve gate tasks of a block (4 aux warps + 8 dw3x3 warps)
      mbar_init(bar(B_ZFREE + i), 1); // the thread that stored the staging tile, once the TMA has read it
    }
    for (int i = 0; i < K::NPAIR; ++i) {
      mbar_init(bar(B_R1F + i), 2);   // the two drain warps that own the rows of the pair
      mbar_init(bar(B_R1E + i), 8);   // the eight dwA warps
    }
    for (int i = 0; i < 8 * K::NPAIR; ++i) {
      mbar_init(bar(B_GTF + i), 1);   // per (chunk plane, row slot): the dwA warp that owns the plane ...
      mbar_init(bar(B_GTE + i), 1);   // ... and the dwB warp that consumes it (no all-to-all coupling between the two groups)
    }
    mbar_init(bar(B_RFREE), 1);       // the warp that summed the per-warp channel sums
    for (int i = 0; i < 8; ++i) reinterpret_cast<uint32_t *>(smem + K::S_CNT)[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int i = tid; i < K::W1_BYTES / 16; i += kStThreads) cp_async16(smem + K::S_W1 + i * 16, wb + K::OFF_W1 + i * 16, true);
  for (int i = tid; i < K::W2_BYTES / 16; i += kStThreads) cp_async16(smem + K::S_W2 + i * 16, wb + K::OFF_W2 + i * 16, true);
  cp_async_commit();
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  cp_async_wait<0>();
  // W2 is held at HALF scale (exact in fp16): GEMM2 then delivers a/2 and b/2, and SimpleGate2 a * sigmoid(b) =
  // (a/2) * tanh(b/2) + a/2 is one MUFU and one FFMA per element.  Each thread rescales the vectors it copied itself.
  for (int i = tid; i < K::W2_BYTES / 16; i += kStThreads) {
    uint4 *q = reinterpret_cast<uint4 *>(smem + K::S_W2 + i * 16);
    H8 v = lds_h8(reinterpret_cast<const unsigned char *>(q));
    const __half2 hf = __float2half2_rn(0.5f);
#pragma unroll
    for (int j = 0; j < 4; ++j) v.h[j] = __hmul2(v.h[j], hf);
    sts_h8(reinterpret_cast<unsigned char *>(q), v);
  }
  fence_async_proxy();     // weights -> visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  StSched sc;
  sc.init(geo, blockIdx.x);

  if (warp < 8) {
    // ================================================================ dwA: dw3x3 + id on both halves, SimpleGate -> GATED rows
    // lane = (half, pixel pair xp): gated pixels 2xp, 2xp+1 of the row (30 per row); region pixels 2xp .. 2xp+3 feed them.
    // Row slots of the G1 / GATED rings store pixel px at 16-byte slot (px >> 1) + 16 * (px & 1): the pair loads of consecutive
    // lanes are then contiguous (conflict-free), and this lane's own gated pixel 2xp + half lands at slot == lane.
    const int p = warp, half = lane >> 4, xp = lane & 15;
    const bool active = xp < 15;
    const int plane = half * 8 + p;
    H8 w[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) w[i] = ldg_h8(wb + K::OFF_DA + (i * 2 * C + plane * 8) * 2);   // centre tap carries RepConv2's "+ x"
    H8 accC[2], accM[2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) accC[j].h[i] = accM[j].h[i] = __float2half2_rn(0.f);
    const unsigned char *g1p = smem + K::S_G1 + plane * 512 + xp * 16;
    unsigned char *gtp = smem + K::S_GT + p * 512 + lane * 16;
    int s3 = 0;                      // row pair slot of both rings (they advance together)
    uint32_t ph3 = 0;
    const uint32_t b_r1f = bar(B_R1F), b_r1e = bar(B_R1E), b_gtf = bar(B_GTF + p * K::NPAIR), b_gte = bar(B_GTE + p * K::NPAIR);
    // this warp's share of the gate stage: TMEM lane quarter p & 3 (= output row of the block), channel chunks 4 + 2 (p >> 2) .. +1
    GateTask<KC1, 2> gate;
    gate.init(sc, geo, p & 3, 4 + 2 * (p >> 2), lane);
    uint32_t gdone = 0;
    while (sc.next()) {
      const int t = sc.col / geo.nstrips, sx = sc.col - t * geo.nstrips;
      const int x0 = sx * K::TW, ys = sc.off * 4;
      const int gx = x0 - 2 + 2 * xp + half;
      const bool xok = active && gx >= 0 && gx < d.W;
      int gy = ys - 6;
      const int nsteps = sc.len + 2;
#pragma unroll 1
      for (int k = 0; k < nsteps; ++k) {
#pragma unroll 1
        for (int hp = 0; hp < 2; ++hp) {       // two row pairs per step
          mbar_wait_nt(b_r1f + 8 * s3, ph3);
          uint4 og[2];
#pragma unroll
          for (int rr = 0; rr < 2; ++rr, ++gy) {
            const unsigned char *rp = g1p + (2 * s3 + rr) * K::G1_ROW;
            const H8 v0 = lds_h8(rp), v1 = lds_h8(rp + 256), v2 = lds_h8(rp + 16), v3 = lds_h8(rp + 272);
            // region row n: bottom tap row of gated row n-1 (completes it), centre row of gated row n, top row of gated row n+1
            H8 res0, res1;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              res0.h[i] = __hfma2(v0.h[i], w[6].h[i], accC[0].h[i]);
              res1.h[i] = __hfma2(v1.h[i], w[6].h[i], accC[1].h[i]);
            }
            h8_fma(res0, v1, w[7]); h8_fma(res0, v2, w[8]);
            h8_fma(res1, v2, w[7]); h8_fma(res1, v3, w[8]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              accC[0].h[i] = __hfma2(v0.h[i], w[3].h[i], accM[0].h[i]);
              accC[1].h[i] = __hfma2(v1.h[i], w[3].h[i], accM[1].h[i]);
            }
            h8_fma(accC[0], v1, w[4]); h8_fma(accC[0], v2, w[5]);
            h8_fma(accC[1], v2, w[4]); h8_fma(accC[1], v3, w[5]);
            h8_mul(accM[0], v0, w[0]); h8_fma(accM[0], v1, w[1]); h8_fma(accM[0], v2, w[2]);
            h8_mul(accM[1], v1, w[0]); h8_fma(accM[1], v2, w[1]); h8_fma(accM[1], v3, w[2]);
            // SimpleGate: the a-half lane finishes pixel 2xp, the b-half lane pixel 2xp+1; each sends the other its half of that pixel
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t r0 = h2u(res0.h[i]), r1 = h2u(res1.h[i]);
              const uint32_t snd = half ? r0 : r1, own = half ? r1 : r0;
              const uint32_t rcv = __shfl_xor_sync(0xffffffffu, snd, 16);
              o[i] = h2u(__hmul2(u2h(own), u2h(rcv)));
            }
            const uint32_t mask = (xok && gy >= 0 && gy < d.H) ? 0xffffffffu : 0u;   // zero padding of the following conv
            og[rr] = make_uint4(o[0] & mask, o[1] & mask, o[2] & mask, o[3] & mask);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(b_r1e + 8 * s3);
          mbar_wait_nt(b_gte + 8 * s3, ph3 ^ 1);
          if (active) {
            *reinterpret_cast<uint4 *>(gtp + (2 * s3) * K::GT_ROW) = og[0];
            *reinterpret_cast<uint4 *>(gtp + (2 * s3 + 1) * K::GT_ROW) = og[1];
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(b_gtf + 8 * s3);
          if (++s3 == K::NPAIR) { s3 = 0; ph3 ^= 1; }
        }
        if (++gdone > LAG_A) gate.run(smem, sbase, tmem, d, geo, &tm_z, lane);
      }
    }
    for (uint32_t i = 0; i < (gdone < LAG_A ? gdone : LAG_A); ++i) gate.run(smem, sbase, tmem, d, geo, &tm_z, lane);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
  } else if (warp < 16) {
    // ================================================================ dwB: merged 5x5 on the gated tensor -> A2 (GEMM2 operand)
    // lane = (e: 4-channel half of the chunk, j: output pixel pair 2j, 2j+1 of the 26); gated pixels 2j .. 2j+5 feed the pair.
    const int c = warp - 8, e = lane & 1, j = lane >> 1;
    const bool active = j < 13;
    const int hc = 2 * c + e;
    __half2 w[25][2];
#pragma unroll
    for (int i = 0; i < 25; ++i) {
      const uint2 ww = __ldg(reinterpret_cast<const uint2 *>(wb + K::OFF_DB + (i * C + hc * 4) * 2));   // centre tap carries "+ x"
      w[i][0] = u2h(ww.x);
      w[i][1] = u2h(ww.y);
    }
    // acc[a][px][ee]: after gated row m, acc[a] holds the partial sums of output row (m - a) (kernel rows 0..a done)
    __half2 acc[4][2][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[a][q >> 1][q & 1] = __float2half2_rn(0.f);
    const unsigned char *gtp = smem + K::S_GT + c * 512 + e * 8 + j * 16;
    unsigned char *a2p = smem + K::S_A2 + c * 2048 + (2 * j) * 16 + e * 8;
    int s3 = 0;
    uint32_t ph3 = 0, g = 0, v = 0;
    const uint32_t b_gtf = bar(B_GTF + c * K::NPAIR), b_gte = bar(B_GTE + c * K::NPAIR);
    while (sc.next()) {
      const int nsteps = sc.len + 2;
#pragma unroll 1
      for (int k = 0; k < nsteps; ++k, ++g) {
        const uint32_t blk = g & 1;
#pragma unroll 1
        for (int hp = 0; hp < 2; ++hp) {       // two row pairs per step
          mbar_wait_nt(b_gtf + 8 * s3, ph3);
          uint2 outp[2][2];
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const unsigned char *rp = gtp + (2 * s3 + rr) * K::GT_ROW;
            __half2 vv[6][2];
#pragma unroll
            for (int tx = 0; tx < 6; ++tx) {
              const uint2 ld = *reinterpret_cast<const uint2 *>(rp + (tx & 1) * 256 + (tx >> 1) * 16);
              vv[tx][0] = u2h(ld.x);
              vv[tx][1] = u2h(ld.y);
            }
#pragma unroll
            for (int px = 0; px < 2; ++px) {
              __half2 out[2];
#pragma unroll
              for (int ee = 0; ee < 2; ++ee) {
                // kernel row 4 completes output row m-4; rows 3..1 advance the younger accumulators; row 0 starts a new one
                __half2 o = __hfma2(vv[px][ee], w[20][ee], acc[3][px][ee]);
#pragma unroll
                for (int tx = 1; tx < 5; ++tx) o = __hfma2(vv[px + tx][ee], w[20 + tx][ee], o);
                out[ee] = o;
#pragma unroll
                for (int a = 3; a >= 1; --a) {
                  __half2 sa = __hfma2(vv[px][ee], w[5 * a][ee], acc[a - 1][px][ee]);
#pragma unroll
                  for (int tx = 1; tx < 5; ++tx) sa = __hfma2(vv[px + tx][ee], w[5 * a + tx][ee], sa);
                  acc[a][px][ee] = sa;
                }
                __half2 s0 = __hmul2(vv[px][ee], w[0][ee]);
#pragma unroll
                for (int tx = 1; tx < 5; ++tx) s0 = __hfma2(vv[px + tx][ee], w[tx][ee], s0);
                acc[0][px][ee] = s0;
              }
              outp[rr][px] = make_uint2(h2u(out[0]), h2u(out[1]));
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(b_gte + 8 * s3);
          if (hp == 0) mbar_wait_nt(bar(B_A2E + blk), ((g >> 1) & 1) ^ 1);   // GEMM2 of two steps ago has read this block
          if (active) {
            unsigned char *op = a2p + blk * K::A2_BLK + (2 * hp) * 512;
            *reinterpret_cast<uint2 *>(op) = outp[0][0];
            *reinterpret_cast<uint2 *>(op + 16) = outp[0][1];
            *reinterpret_cast<uint2 *>(op + 512) = outp[1][0];
            *reinterpret_cast<uint2 *>(op + 528) = outp[1][1];
          }
          if (++s3 == K::NPAIR) { s3 = 0; ph3 ^= 1; }
        }
        fence_async_proxy();   // generic-proxy writes of A2 -> visible to the tensor core
        __syncwarp();
        // the warp that completes the block hands it to the tensor core (or straight back, for the two warm-up blocks of a piece)
        if (lane == 0 && last_arrival(cnt(N_A2 + blk), 8)) {
          if (k < 2) {
            mbar_arrive(bar(B_A2E + blk));
          } else {
            const uint32_t zs = v & 1;
            mbar_wait(bar(B_G2E + zs), ((v >> 1) & 1) ^ 1);              // the gate stage has read the accumulator of two blocks ago
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < K::KC2 / 2; ++kk) {
              const uint64_t ad = smem_desc_at(smem16, K::S_A2 + blk * K::A2_BLK + 2 * kk * (K::M * 16), K::M * 16, 128);
              const uint64_t bd = smem_desc_at(smem16, K::S_W2 + 2 * kk * (K::N * 16), K::N * 16, 128);
              umma_f16(tmem + 256 + zs * K::N, ad, bd, idesc, kk > 0);
            }
            umma_commit(bar(B_G2F + zs));
            umma_commit(bar(B_A2E + blk));
          }
        }
        if (k >= 2) ++v;
        __syncwarp();
      }
    }
  } else {
    // ================================================================ drain: GEMM1 accumulators -> G1 rows ; TMA + GEMM1 issue
    const int q = warp - 16;                                   // TMEM lane quarter = row of the step (warp % 4 == q)
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const int slot_off = ((lane >> 1) + (lane & 1) * 16) * 16; // pixel x = lane of the row -> its 16-byte slot
    const bool issuer = q == 3 && lane == 0;                   // the row of warp 19 belongs to the later pair of the step
    uint32_t nsteps_total = 0;
    {
      StSched c0 = sc;
      while (c0.next()) nsteps_total += (uint32_t)c0.len + 2;
    }
    StCur ct;                        // cursor of the TMA loads (issuer only)
    uint32_t gt_ = 0;                // steps requested so far
    auto issue_tma = [&]() {
      const uint32_t st = gt_ & 1;
      const int t = ct.t, sx = ct.sx;
      const int x0 = sx * K::TW, y = (ct.s.off + ct.k) * 4 - 5;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar(B_A1F + st)), "r"(K::A1_STAGE) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
              "r"(sbase + K::S_A1 + st * K::A1_STAGE), "l"(reinterpret_cast<uint64_t>(&tm_a1)), "r"((x0 - 3) * 8), "r"(y), "r"(0), "r"(t),
          "r"(bar(B_A1F + st))
          : "memory");
      ct.step();
      ++gt_;
    };
    auto issue_gemm1 = [&](uint32_t g1) {     // operand stage / accumulator slot g1 & 1
      const uint32_t st = g1 & 1, u = g1 >> 1;
      mbar_wait(bar(B_G1E + st), (u & 1) ^ 1);          // the drain of two steps ago has left the accumulator slot
      mbar_wait(bar(B_A1F + st), u & 1);                // the operand rows have landed
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < KC1 / 2; ++k) {
        const uint64_t ad = smem_desc_at(smem16, K::S_A1 + st * K::A1_STAGE + 2 * k * (K::M * 16), K::M * 16, 128);
        const uint64_t bd = smem_desc_at(smem16, K::S_W1 + 2 * k * (K::N * 16), K::N * 16, 128);
        umma_f16(tmem + st * K::N, ad, bd, idesc, k > 0);
      }
      umma_commit(bar(B_G1F + st));
      umma_commit(bar(B_A1E + st));
    };
    if (issuer && nsteps_total) {
      ct.start(sc, geo.nstrips);
      issue_tma();
      if (nsteps_total > 1) issue_tma();
      issue_gemm1(0);
    }
    // ---- gate stage of the block LAG_D steps behind the drain: this warp = TMEM lane quarter q, channel chunks 0..3
    GateTask<KC1, 4> gate;
    gate.init(sc, geo, q, 0, lane);
#pragma unroll 1
    for (uint32_t g = 0; g < nsteps_total; ++g) {
      // ---- drain step g: TMEM slot g&1, lanes [32q, 32q+32) = row q of the step -> G1 ring row (4g + q) % 6
      const uint32_t gs = g & 1;
      const uint32_t pair = 2 * g + (q >> 1), u = pair / K::NPAIR, s3 = pair - u * K::NPAIR;   // row pair of the ring
      mbar_wait(bar(B_G1F + gs), (g >> 1) & 1);
      tc_fence_after();
      if (issuer) {
        // GEMM1(g) is complete: its operand stage is free for the rows of step g+2, and the tensor core for GEMM1(g+1)
        if (g + 2 < nsteps_total) issue_tma();
        if (g + 1 < nsteps_total) issue_gemm1(g + 1);
      }
      __syncwarp();
      mbar_wait(bar(B_R1E + s3), (u & 1) ^ 1);
      unsigned char *rowp = smem + K::S_G1 + (2 * s3 + (q & 1)) * K::G1_ROW + slot_off;
#pragma unroll 1
      for (int cg = 0; cg < 4; cg += 2) {
        uint32_t va[32], vb[32];
        const uint32_t ta = tmem + tlane + gs * K::N + cg * 32;
        tmem_ld32_nowait(ta, va);
        tmem_ld32_nowait(ta + 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint4 o;
          o.x = pack_half2(__uint_as_float(va[c4 * 8 + 0]), __uint_as_float(va[c4 * 8 + 1]));
          o.y = pack_half2(__uint_as_float(va[c4 * 8 + 2]), __uint_as_float(va[c4 * 8 + 3]));
          o.z = pack_half2(__uint_as_float(va[c4 * 8 + 4]), __uint_as_float(va[c4 * 8 + 5]));
          o.w = pack_half2(__uint_as_float(va[c4 * 8 + 6]), __uint_as_float(va[c4 * 8 + 7]));
          *reinterpret_cast<uint4 *>(rowp + (cg * 4 + c4) * 512) = o;
        }
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint4 o;
          o.x = pack_half2(__uint_as_float(vb[c4 * 8 + 0]), __uint_as_float(vb[c4 * 8 + 1]));
          o.y = pack_half2(__uint_as_float(vb[c4 * 8 + 2]), __uint_as_float(vb[c4 * 8 + 3]));
          o.z = pack_half2(__uint_as_float(vb[c4 * 8 + 4]), __uint_as_float(vb[c4 * 8 + 5]));
          o.w = pack_half2(__uint_as_float(vb[c4 * 8 + 6]), __uint_as_float(vb[c4 * 8 + 7]));
          *reinterpret_cast<uint4 *>(rowp + ((cg + 1) * 4 + c4) * 512) = o;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(B_G1E + gs));     // TMEM slot free for the GEMM1 of step g+2
        mbar_arrive(bar(B_R1F + s3));     // row ready for the dwA warps (the pair completes with the neighbour warp's row)
      }
      if (g >= LAG_D) gate.run(smem, sbase, tmem, d, geo, &tm_z, lane);   // block of step g - LAG_D
    }
    for (uint32_t i = 0; i < (nsteps_total < LAG_D ? nsteps_total : LAG_D); ++i) gate.run(smem, sbase, tmem, d, geo, &tm_z, lane);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512));
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------------
static StGeom stream_geometry(int T, int H, int W) {
  StGeom g;
  g.nstrips = (W + 25) / 26;
  g.nblk_col = (H + 3) / 4;
  g.total = (long long)T * g.nstrips * g.nblk_col;
  const long long sms = sm_count();
  const long long grid = g.total < sms ? g.total : sms;
  g.chunk = (g.total + grid - 1) / grid;
  g.ppc = (int)((g.nblk_col - 1) / g.chunk) + 2;
  g.ntiles = g.nstrips * g.ppc;
  return g;
}

bool pass_a_stream_enabled() {
  static const bool v = [] { const char *e = getenv("GSN_PASS_A_STREAM"); return e && e[0] == '1'; }();
  return v;
}

int pass_a_stream_tiles(int T, int H, int W) { return stream_geometry(T, H, W).ntiles; }

template <int KC1>
static int launch_stream(const GsnCabPassA &d, cudaStream_t st, const CUtensorMap &tm, const CUtensorMap &tm_z, const StGeom &geo) {
  using K = StCfg<KC1>;
  GSN_ONCE_PER_DEVICE(cudaFuncSetAttribute(cab_pass_a_stream_kernel<KC1>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
  const long long sms = sm_count();
  const unsigned grid = (unsigned)(geo.total < sms ? geo.total : sms);
  // partial-sum slots that no piece owns must read as zero in cab_fold
  cudaMemsetAsync(d.chan_partial, 0, (size_t)d.T * geo.ntiles * 64 * sizeof(float), st);
  cab_pass_a_stream_kernel<KC1><<<grid, kStThreads, K::SMEM, st>>>(d, tm, tm_z, geo);
  count_launch();
  return check_launch("cab_pass_a_stream");
}

int cab_pass_a_stream_dispatch(const GsnCabPassA &d, cudaStream_t st) {
  const bool shift = d.mode != GSN_MODE_CAB1;
  const StGeom geo = stream_geometry(d.T, d.H, d.W);
  CUtensorMap tm, tm_z;
  memset(&tm, 0, sizeof(tm));
  memset(&tm_z, 0, sizeof(tm_z));
  if (!encode_tmap_planar(&tm, d.a1_pre, d.W, d.H, shift ? 12 : 8, d.T, 32, 4)) {
    set_error("cab_pass_a (stream): cuTensorMapEncodeTiled(a1) failed (W=%d H=%d T=%d)", d.W, d.H, d.T);
    return GSN_E_CUDA;
  }
  if (!encode_tmap_nhwc(&tm_z, d.z, 64, d.W, d.H, d.T, 64, 26, 4, true)) {
    set_error("cab_pass_a (stream): cuTensorMapEncodeTiled(z) failed (W=%d H=%d T=%d)", d.W, d.H, d.T);
    return GSN_E_CUDA;
  }
  return shift ? launch_stream<12>(d, st, tm, tm_z, geo) : launch_stream<8>(d, st, tm, tm_z, geo);
}

}  // namespace gsn
