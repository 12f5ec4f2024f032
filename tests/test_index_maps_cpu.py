"""The product's own index maps, compiled for the HOST from csrc/shift_common.cuh with nvcc (no GPU needed), against the oracle:
the temporal-roll source map `roll_source` (wrap / clamp / GSN_ROLL_HALO) and the spatial-shift offset table `make_shift_table`."""
import os
import shutil
import subprocess
import sys

import pytest
import torch

import golden_io as gio

sys.path.insert(0, gio.ROOT)
from oracle import shiftnet_oracle as O  # noqa: E402

SRC = r'''
#include <cstdio>
#include "shift_common.cuh"
int main() {
  using namespace gsn;
  for (int mode = 0; mode <= 2; ++mode)
    for (int circ = 0; circ <= 2; ++circ)
      for (int T = 1; T <= 6; ++T)
        for (int t = 0; t < T; ++t) {
          const RollSrc r = roll_source(mode, circ, t, T, 16);
          printf("R %d %d %d %d %d %d %d %d\n", mode, circ, T, t, r.f_lo, r.c_lo, r.f_hi, r.c_hi);
        }
  for (int C : {64, 80}) {
    const ShiftTable s = make_shift_table(C);
    for (int c = 0; c < C / 2; ++c) printf("S %d %d %d %d\n", C, c, (int)s.dy[c], (int)s.dx[c]);
  }
  return 0;
}
'''


@pytest.fixture(scope="module")
def host_maps(tmp_path_factory):
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    d = tmp_path_factory.mktemp("maps")
    src = d / "maps.cu"
    src.write_text(SRC)
    exe = d / "maps"
    subprocess.run(["nvcc", "-std=c++17", "-I", os.path.join(gio.ROOT, "shift-net_b200", "csrc"), str(src), "-o", str(exe)], check=True,
                   capture_output=True)
    return subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()


def _rolled(x, rs_rows, T, C):
    """Apply a roll_source table to x (Tx, C): y[t, :C/2] = x[f_lo, c_lo:c_lo+C/2], y[t, C/2:] = x[f_hi, c_hi:c_hi+C/2]."""
    h = C // 2
    y = torch.empty(T, C)
    for t in range(T):
        f_lo, c_lo, f_hi, c_hi = rs_rows[t]
        y[t, :h] = x[f_lo, c_lo:c_lo + h]
        y[t, h:] = x[f_hi, c_hi:c_hi + h]
    return y


def test_roll_source_matches_the_oracle_roll(host_maps):
    """fwd / rev x wrap / clamp: the rolled stream built from the kernels' source map equals the oracle's temporal_roll
    (gshift_deblur2.py:499-512, gshift_deblur1.py:504-519); MODE_CAB1 is the identity."""
    C = 16
    table = {}
    for line in host_maps:
        if line.startswith("R "):
            mode, circ, T, t, *r = map(int, line.split()[1:])
            table.setdefault((mode, circ, T), {})[t] = r
    for T in range(1, 7):
        x = torch.arange(T * C, dtype=torch.float32).view(T, C)
        x4 = x.view(T, C, 1, 1)
        for circ in (0, 1):
            for mode, rev in ((1, False), (2, True)):
                want = O.temporal_roll(x4, rev, bool(circ))[0].view(T, C)
                assert torch.equal(_rolled(x, table[(mode, circ, T)], T, C), want), (mode, circ, T)
            assert torch.equal(_rolled(x, table[(0, circ, T)], T, C), x)


def test_roll_source_halo_mode_is_the_wrap_over_one_more_frame(host_maps):
    """GSN_ROLL_HALO (T-sharded clips): for the T own frames the source map equals the wrapping map over T + 1 frames -- frame 0's
    predecessor and frame T-1's successor are both the halo frame at index T -- and never points outside [0, T]."""
    table = {}
    for line in host_maps:
        if line.startswith("R "):
            mode, circ, T, t, *r = map(int, line.split()[1:])
            table.setdefault((mode, circ, T), {})[t] = r
    for T in range(1, 6):
        for mode in (1, 2):
            for t in range(T):
                assert table[(mode, 2, T)][t] == table[(mode, 1, T + 1)][t], (mode, T, t)
                f_lo, _, f_hi, _ = table[(mode, 2, T)][t]
                assert 0 <= f_lo <= T and 0 <= f_hi <= T
        assert table[(1, 2, T)][0][0] == T and table[(2, 2, T)][T - 1][2] == T
        for t in range(T):
            assert table[(0, 2, T)][t] == [t, 0, t, 8]


def test_shift_table_matches_the_oracle_offsets(host_maps):
    """Per-channel (dy, dx) of spatial_shift2 for C = 64 and C = 80 (gshift_deblur2.py:465-498, gshift_deblur1.py:470-503)."""
    got = {}
    for line in host_maps:
        if line.startswith("S "):
            C, c, dy, dx = map(int, line.split()[1:])
            got.setdefault(C, []).append((dy, dx))
    for C in (64, 80):
        assert got[C] == [tuple(o) for o in O.shift_offsets(C)]
