"""Host-side model of the shared-memory fragment addressing of csrc/gemm_tma.cu (FragAddr + the lane permutations
sigma / pos0 / rho): for both operand orientations and both dtypes it checks, against the 128-byte hardware swizzle
the TMA writes, that (a) every lane reads exactly the element (row, k) the DMMA fragment layout wants, (b) the 16/8
lanes that the LSU serves in one phase hit distinct banks (conflict-free LDS.64 / LDS.128), and (c) the permutations
are bijections of the warp-tile rows, so the epilogue's renaming covers every output element once."""
import itertools

import pytest


def sigma(g): return 2 * (g & 3) + (g >> 2)
def pos0(g): return (g & 1) + 8 * ((g >> 1) & 1) + 2 * (g >> 2)
def rho(g): return (g >> 1) | ((g & 1) << 2)


def frag_index(cplx, kc, i, g):
    if cplx:
        return 8 * i + rho(g)
    return 8 * i + sigma(g) if kc else 16 * (i >> 1) + 4 * (i & 1) + pos0(g)


def frag_offset(cplx, kc, w0, i, s, g, t):
    """FragAddr::init + FragAddr::at, restated."""
    if not cplx and kc:
        sg = sigma(g)
        base = (w0 + sg) * 128 + (t & 1) * 8
        tab = (((2 * s) ^ (sg & 6)) | ((t >> 1) ^ (sg & 1))) << 4
        return base + i * 1024 + tab
    if not cplx and not kc:
        ph = 4 * ((g >> 1) & 1) + (g >> 2)
        base = (w0 // 16) * 2048 + t * 128 + (g & 1) * 8
        ih, sl = i & 1, s & 1
        tab = ((ph | (2 * ih)) ^ (4 * sl + t)) << 4
        return base + (i >> 1) * 2048 + s * 512 + tab
    rg = rho(g)
    tab = ((4 * s + t) ^ rg) << 4
    if kc:
        return (w0 + rg) * 128 + i * 1024 + tab
    return (w0 // 8) * 1024 + t * 128 + i * 1024 + s * 512 + tab


def swizzle128(linear):
    """CU_TENSOR_MAP_SWIZZLE_128B: the 16-byte chunk index (bits 4..6) is XORed with the 128-byte line index mod 8."""
    return linear ^ (((linear >> 7) & 7) << 4)


def expected_offset(cplx, kc, row, k):
    """Where the TMA put element (row, k) of the tile: KC box = [row][k] with 128-byte lines of k; RC box =
    [row_hi][k][row_lo] with 128-byte lines of rows."""
    es = 16 if cplx else 8
    per_line = 128 // es
    if kc:
        return swizzle128(row * 128 + k * es)
    sub, lo = divmod(row, per_line)
    bk = 8 if cplx else 16
    return swizzle128(sub * (bk * 128) + k * 128 + lo * es)


CONFIGS = [(False, True), (False, False), (True, True), (True, False)]


@pytest.mark.parametrize("cplx,kc", CONFIGS)
def test_every_lane_reads_its_dmma_element(cplx, kc):
    frags = (8 if not cplx else 4)                    # fragments per warp tile (WM/8 = 8 or 4; WN/8 = 4)
    ksteps = 4 if not cplx else 2
    for w0 in (0, 32, 64):
        for i, s, lane in itertools.product(range(frags), range(ksteps), range(32)):
            g, t = lane >> 2, lane & 3
            row = w0 + frag_index(cplx, kc, i, g)
            k = 4 * s + t
            assert frag_offset(cplx, kc, w0, i, s, g, t) == expected_offset(cplx, kc, row, k), (w0, i, s, lane)


@pytest.mark.parametrize("cplx,kc", CONFIGS)
def test_fragment_loads_are_bank_conflict_free(cplx, kc):
    width = 16 if cplx else 8                         # LDS.128 / LDS.64
    lanes_per_phase = 128 // width                    # 8 / 16 lanes are served together
    frags = 8 if not cplx else 4
    ksteps = 4 if not cplx else 2
    for i, s in itertools.product(range(frags), range(ksteps)):
        for phase in range(32 // lanes_per_phase):
            lanes = range(phase * lanes_per_phase, (phase + 1) * lanes_per_phase)
            slots = {(frag_offset(cplx, kc, 0, i, s, l >> 2, l & 3) % 128) // width for l in lanes}
            assert len(slots) == lanes_per_phase, (i, s, phase)


@pytest.mark.parametrize("cplx,kc", CONFIGS)
def test_row_permutation_is_a_bijection(cplx, kc):
    frags = 8 if not cplx else 4
    rows = sorted(frag_index(cplx, kc, i, g) for i in range(frags) for g in range(8))
    assert rows == list(range(8 * frags))
    # the float64 RC permutation hands each lane two ADJACENT output columns (vectorised epilogue stores)
    if not cplx and not kc:
        for j, t in itertools.product(range(4), range(4)):
            c0, c1 = frag_index(False, False, j, 2 * t), frag_index(False, False, j, 2 * t + 1)
            assert c1 == c0 + 1 and c0 % 2 == 0


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("kc", [True, False])
@pytest.mark.parametrize("rows", [128, 64])
def test_cpasync_padded_tiles_are_bank_conflict_free(cplx, kc, rows):
    """The fallback kernel (csrc/gemm.cu) pads its tiles instead of swizzling: KC [rows][BK + 4], RC [BK][rows + pad]
    (pad 4 for float64, 2 for complex128) -- same conflict-freeness per LSU phase."""
    es, bk = (16, 8) if cplx else (8, 16)
    if cplx and rows == 128 and not kc:
        pass                                           # BM = 128 for complex A tiles, BN = 64 for B tiles
    ld = (bk + 4) if kc else (rows + (2 if cplx else 4))
    lanes_per_phase = 128 // es
    for w0, i, kk in itertools.product((0, 32), range(4), range(bk // 4)):
        for phase in range(32 // lanes_per_phase):
            slots = set()
            for lane in range(phase * lanes_per_phase, (phase + 1) * lanes_per_phase):
                g, t = lane >> 2, lane & 3
                r, k = w0 + 8 * i + g, 4 * kk + t
                elem = (r * ld + k) if kc else (k * ld + r)
                slots.add((elem * es % 128) // es)
            assert len(slots) == lanes_per_phase, (w0, i, kk, phase)
