"""CPU tier: index arithmetic of the pair-search kernel restated in Python (s2tc_b200/csrc/kernels_search.cu).  The byte
parity of the kernel is a GPU test; these are the combinatorial invariants its tile shapes rely on, checked exhaustively:
every unordered pair i < j < m must be met by at least one lane of at least one tile (the reference scans all of them,
s2tc_algorithm.cpp:393-410), and the list entries must decode to the rows that were tested."""
import itertools

DIAG = 0x80000000


def tile_pairs(m):
    """(i, j) slots the kernel's tiles evaluate for m candidate rows, as (row, row) with rows possibly >= m (padding)"""
    ntile = (m + 15) // 16
    met = []
    for b in range(ntile):
        for lane in range(32):
            jj, half = lane & 15, lane >> 4
            j = 16 * b + jj
            for a in range(b):                                  # tiles below the diagonal: 8 rows x 16 columns per half
                for t in range(8):
                    met.append((16 * a + 8 * half + t, j))
            o0 = jj + 1 + 4 * half                              # diagonal tile: four rows at circular offsets
            for t in range(4):
                met.append((16 * b + ((o0 + t) & 15), j))
    return met


def test_every_pair_is_met():
    for m in (2, 5, 16, 17, 19, 33, 56, 80, 96, 129):
        want = set(itertools.combinations(range(m), 2))
        got = {(min(i, j), max(i, j)) for i, j in tile_pairs(m) if i != j and max(i, j) < m}
        assert got == want, m


def test_diagonal_tile_has_no_waste_but_the_eight_antipodes():
    met = [(min(i, j), max(i, j)) for i, j in tile_pairs(16)]
    assert len(met) == 128 and len(set(met)) == 120
    dup = [p for p in set(met) if met.count(p) == 2]
    assert sorted(dup) == [(k, k + 8) for k in range(8)]
    # VABSDIFF4 per searched block at m = 80: 10 full tiles x 32 + 5 diagonal tiles x 16 per lane; floor = pairs * 4 / 32 lanes
    ntile = 5
    assert (ntile * (ntile - 1) // 2) * 32 + ntile * 16 == 400 and 80 * 79 // 2 * 4 / 32 == 395


def test_list_entries_decode_to_the_rows_that_were_tested():
    """flush: entry = [flag | base << 16 | j]; pair t of the group is row base + t, wrapping inside the tile for diagonal groups"""
    def decode(e, t):
        base = (e >> 16) & 0x7FFF
        i = ((base & ~15) | ((base + t) & 15)) if e & DIAG else base + t
        return i, e & 0xFFFF
    for b in range(6):
        for lane in range(32):
            jj, half = lane & 15, lane >> 4
            j = 16 * b + jj
            o0 = jj + 1 + 4 * half
            e = DIAG | ((16 * b + (o0 & 15)) << 16) | j
            assert [decode(e, t) for t in range(4)] == [(16 * b + ((o0 + t) & 15), j) for t in range(4)]
            for a in range(b):
                for g in (0, 4):
                    ig = 16 * a + 8 * half + g
                    assert [decode((ig << 16) | j, t) for t in range(4)] == [(ig + t, j) for t in range(4)]
