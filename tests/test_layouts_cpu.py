"""Host-side models of the shared-memory / HBM layouts the sm_100a kernels rely on (no GPU needed).

The CUDA kernels index these layouts with closed-form expressions; the tests restate each expression, check it is a bijection
onto the tile it addresses and check the bank-conflict claims made in DESIGN.md section 3 (a 16-byte access by 8 consecutive
lanes = one shared-memory wavefront of 128 bytes: conflict-free iff the 8 lanes hit 8 different 16-byte slots modulo 128)."""
import numpy as np


def _slots_mod128(byte_addrs):
    return sorted((a % 128) // 16 for a in byte_addrs)


def test_swizzle128_staging_tile_is_bijective_and_conflict_free():
    """cab_pass_a_pre.cu (z staging) and cab_pass_b_tc.cu (z operand / out staging): 128-byte pixel rows, 16-byte chunk c of pixel p
    at chunk c ^ (p & 7) -- the layout CU_TENSOR_MAP_SWIZZLE_128B produces and the SWIZZLE_128B UMMA descriptor expects."""
    npx = 256
    addr = lambda p, c: p * 128 + ((c ^ (p & 7)) << 4)
    seen = {addr(p, c) for p in range(npx) for c in range(8)}
    assert seen == set(range(0, npx * 128, 16))
    for c in range(8):                                   # lane = pixel, fixed logical chunk: 8 consecutive pixels per wavefront
        for p0 in range(0, npx, 8):
            assert _slots_mod128(addr(p, c) for p in range(p0, p0 + 8)) == list(range(8))
    # the hardware XOR acts on address bits [4,7) with bits [7,10): same thing written on the byte address
    for p in (0, 5, 77, 255):
        for c in range(8):
            linear = p * 128 + c * 16
            assert addr(p, c) == linear ^ (((linear >> 7) & 7) << 4)


def test_swizzle64_shortcut_halves_are_bijective_and_conflict_free():
    """cab_pass_b_tc.cu: the two rolled halves of the shortcut land as [128 px][64 B] boxes with CU_TENSOR_MAP_SWIZZLE_64B; the epilogue
    thread of pixel r reads logical chunk c at r*64 + ((c ^ ((r >> 1) & 3)) << 4)."""
    npx = 128
    addr = lambda r, c: r * 64 + ((c ^ ((r >> 1) & 3)) << 4)
    assert {addr(r, c) for r in range(npx) for c in range(4)} == set(range(0, npx * 64, 16))
    for c in range(4):
        for r0 in range(0, npx, 8):
            assert _slots_mod128(addr(r, c) for r in range(r0, r0 + 8)) == list(range(8))
    for r in (0, 3, 6, 127):
        for c in range(4):
            linear = r * 64 + c * 16
            assert addr(r, c) == linear ^ (((linear >> 7) & 3) << 4)


def test_planar_operand_layout_matches_the_tma_view():
    """The LayerNorm'd operand [T][KC][H][W][8]: producers write element (t, chunk, y, x, i) at
    (((t*KC + chunk)*H*W + y*W + x)*8 + i); pass A reads it through the 4-D tensor map (W*8, H, KC, T) with strides
    (W*16, H*W*16, KC*H*W*16) bytes and a {22*8, 22, KC, 1} box starting at ((x0-3)*8, y0-3, 0, t)."""
    T, KC, H, W = 2, 12, 20, 37
    a1 = np.arange(T * KC * H * W * 8, dtype=np.int64).reshape(T, KC, H, W, 8)
    flat = a1.reshape(-1)
    strides = (W * 16, H * W * 16, KC * H * W * 16)            # bytes, dims 1..3 of the tensor map (dim 0 is dense, 2 bytes)
    t, x0, y0 = 1, 16, 0
    box = np.zeros((KC, 22, 22 * 8), dtype=np.int64) - 1       # smem order: dim 0 fastest -> [chunk][row][x*8 + i]
    for ch in range(KC):
        for r in range(22):
            for e in range(22 * 8):
                c0, c1 = (x0 - 3) * 8 + e, y0 - 3 + r
                if 0 <= c0 < W * 8 and 0 <= c1 < H:                # out-of-bounds elements are zero-filled by the hardware
                    box[ch, r, e] = flat[(c0 * 2 + c1 * strides[0] + ch * strides[1] + t * strides[2]) // 2]
    # the box is KC planes of 484 16-byte pixel vectors: plane pitch 484*16 B = the UMMA LBO, 8 consecutive pixels = 128 B = SBO
    planes = box.reshape(KC, 22 * 22, 8)
    for ch in (0, 5, KC - 1):
        for (ry, rx) in ((3, 3), (10, 21), (21, 0)):
            gy, gx = y0 - 3 + ry, x0 - 3 + rx
            want = a1[t, ch, gy, gx] if (0 <= gy < H and 0 <= gx < W) else np.full(8, -1)
            assert np.array_equal(planes[ch, ry * 22 + rx], want)


def test_pass_a_shared_memory_map_has_no_live_overlaps():
    """PreCfg<KC1> of cab_pass_a_pre.cu restated: G1 may alias [A2 | A1] (their live ranges are disjoint), everything else is
    disjoint, z staging fits the GATED area, and the whole map fits the 227 KB a CTA can own."""
    for KC1 in (8, 12):
        C, M1, M2, M3 = 64, 22 * 22, 20 * 20, 256
        N, KC2, NC = 2 * C, C // 8, 2 * C // 8
        PA1, P1, P2, P3 = M1 * 16, (M1 + 1) * 16, (M2 + 1) * 16, (M3 + 1) * 16
        X = 256 + 16 * 32 * 4
        S_A2, A2 = X, KC2 * P3
        S_G1, G1 = S_A2, NC * P1
        S_A1, A1 = S_A2 + A2, KC1 * PA1
        lo_end = max(S_G1 + G1, S_A1 + A1)
        S_GT = (lo_end + 1023) // 1024 * 1024
        GT = KC2 * P2
        S_WT2 = (S_GT + GT + 127) // 128 * 128
        WT2 = 9 * 2 * C * 2 + 25 * C * 2 + KC2 * N * 16
        S_W1 = (S_WT2 + WT2 + 127) // 128 * 128
        SMEM = S_W1 + KC1 * N * 16
        assert SMEM <= 227 * 1024
        assert S_A1 % 128 == 0 and S_GT % 1024 == 0                      # TMA destination / swizzle atom alignment
        assert S_A2 + A2 <= S_A1 and S_A1 + A1 <= S_GT and S_G1 + G1 <= S_GT   # A2 | A1 | GATED in order; G1 inside [A2, GATED)
        assert M3 * C * 2 <= GT                                           # z staging inside the (dead) GATED area
        assert S_GT + GT <= S_WT2 and S_WT2 + WT2 <= S_W1                 # resident weights never aliased
        # UMMA rows beyond the 484 real ones (M tile 3 covers rows 384..511) read past the last A1 plane but stay inside the CTA
        assert S_A1 + (KC1 - 1) * PA1 + 4 * 128 * 16 <= SMEM
