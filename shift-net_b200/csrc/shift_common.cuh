// Shared between the two pass-A implementations and pass B: shift offset table and temporal-roll source map.
#pragma once
#include "common.cuh"

namespace gsn {

// Device-side lookup of the same table (constant bank, no per-kernel parameter copy).
static __constant__ int8_t kShiftOuter[16][2] = {{8, 8}, {8, 4}, {8, 0}, {8, -4}, {8, -8}, {-8, 8}, {-8, 4}, {-8, 0}, {-8, -4},
                                                 {-8, -8}, {4, 8}, {4, -8}, {0, 8}, {0, -8}, {-4, 8}, {-4, -8}};
static __constant__ int8_t kShiftInner[8][2] = {{4, 4}, {4, 0}, {4, -4}, {0, 4}, {0, -4}, {-4, 4}, {-4, 0}, {-4, -4}};
template <int C>
__device__ __forceinline__ void shift_offset(int c, int &dy, int &dx) {
  constexpr int number = C / 2 / 8, n2 = (number - 1) / 2, n1 = number - 2 * n2;
  if (c < 16 * n2) { const int g = c / n2; dy = kShiftOuter[g][0]; dx = kShiftOuter[g][1]; }
  else { const int g = (c - 16 * n2) / n1; dy = kShiftInner[g][0]; dx = kShiftInner[g][1]; }
}

struct ShiftTable {
  int8_t dy[40];
  int8_t dx[40];
};

// Per-channel offsets of spatial_shift2 (gshift_deblur2.py:465-498): out[h,w] = in[h-dy, w-dx], zero fill.
inline ShiftTable make_shift_table(int C) {
  static const int outer[16][2] = {{8, 8}, {8, 4}, {8, 0}, {8, -4}, {8, -8}, {-8, 8}, {-8, 4}, {-8, 0}, {-8, -4}, {-8, -8},
                                   {4, 8}, {4, -8}, {0, 8}, {0, -8}, {-4, 8}, {-4, -8}};
  static const int inner[8][2] = {{4, 4}, {4, 0}, {4, -4}, {0, 4}, {0, -4}, {-4, 4}, {-4, 0}, {-4, -4}};
  ShiftTable t{};
  const int number = C / 2 / 8, n2 = (number - 1) / 2, n1 = number - 2 * n2;
  int c = 0;
  for (int g = 0; g < 16; ++g)
    for (int i = 0; i < n2; ++i, ++c) { t.dy[c] = (int8_t)outer[g][0]; t.dx[c] = (int8_t)outer[g][1]; }
  for (int g = 0; g < 8; ++g)
    for (int i = 0; i < n1; ++i, ++c) { t.dy[c] = (int8_t)inner[g][0]; t.dx[c] = (int8_t)inner[g][1]; }
  return t;
}

// Which (frame, first channel) feeds the low / high half of the rolled stream y, and which half is shifted.
// fwd: y[t,c<C/2] = x[t-1, c+C/2], y[t,c>=C/2] = x[t, c-C/2], shifted half = low   (gshift_deblur2.py:504-508)
// rev: y[t,c<C/2] = x[t, c+C/2],   y[t,c>=C/2] = x[t+1, c-C/2], shifted half = high (gshift_deblur2.py:510-511)
// clamped variants keep the boundary frame un-swapped (gshift_deblur1.py:513,517).
struct RollSrc {
  int f_lo, c_lo, f_hi, c_hi;
};
// Frames of the rolled source tensor x: T, or T + 1 with GSN_ROLL_HALO (a T-sharded clip: the neighbour rank's boundary frame is
// stored behind the T own frames, at index T, and is reached by wrapping; only the T own frames are computed).
__host__ __device__ inline int roll_frames(int circular, int T) { return circular == GSN_ROLL_HALO ? T + 1 : T; }
__host__ __device__ inline RollSrc roll_source(int mode, int circular, int t, int T, int C) {
  RollSrc r;
  const int h = C / 2, R = roll_frames(circular, T);
  if (mode == GSN_MODE_CAB2_FWD) {
    if (!circular && t == 0) { r.f_lo = 0; r.c_lo = 0; r.f_hi = 0; r.c_hi = h; }
    else { r.f_lo = (t + R - 1) % R; r.c_lo = h; r.f_hi = t; r.c_hi = 0; }
  } else if (mode == GSN_MODE_CAB2_REV) {
    if (!circular && t == T - 1) { r.f_lo = t; r.c_lo = 0; r.f_hi = t; r.c_hi = h; }
    else { r.f_lo = t; r.c_lo = h; r.f_hi = (t + 1) % R; r.c_hi = 0; }
  } else { r.f_lo = t; r.c_lo = 0; r.f_hi = t; r.c_hi = h; }
  return r;
}


}  // namespace gsn
