"""Host-side model of the item enumeration of k_col_xty_slabs (insider_b200/csrc/k_stream.cu, N > 384): a block's tile range is
walked in groups of XW_GT = 4 gene tiles x row slabs of R = 128 rows; an item = (group, slab) arrives by two tensor-map box copies
whose coordinates are (slab * R, first gene of the group) for Y and (slab * R, 0) for U^T; the box is pitch4(R) = 132 rows x 64 genes.
Checks, for the shapes the configurations use and for awkward ones: every (tile, slab) of every block is covered exactly once, in an
order in which a tile's slabs are consecutive (its accumulators live in registers across slabs); partial last groups never store
tiles outside the block's range; every row the contraction reads (rows_here per slab) lies inside the box and inside the padded
matrix; boxes that reach past the matrix (rows >= ldY, genes >= P_pad) only ever do so in their unused part (zero fill)."""
import pytest

TG, XW_GT, R = 16, 4, 128


def pitch4(n):
    p = (n + 7) // 8 * 8 + 4
    return p - 8 if p - 8 >= n else p


def split_range(n, parts, idx):
    q, r = divmod(n, parts)
    b = idx * q + min(idx, r)
    return b, b + q + (1 if idx < r else 0)


@pytest.mark.parametrize("N,P,blocks", [(17382, 7025, 148), (17382, 56200, 148), (5000, 20000, 148), (600, 11850, 148), (385, 40, 3),
                                        (1000, 5000, 148), (720, 96, 148), (513, 1, 148), (4097, 16 * 149, 148)])
def test_items_cover_every_tile_and_slab_once(N, P, blocks):
    ldY = pitch4(N)
    P_pad = max(TG, (P + TG - 1) // TG * TG)
    n_tiles = P_pad // TG
    n_blocks = min(blocks, n_tiles)
    n_slabs = (ldY + R - 1) // R
    box_rows, box_genes = pitch4(R), XW_GT * TG
    seen = {}
    for b in range(n_blocks):
        t0, t1 = split_range(n_tiles, n_blocks, b)
        n_groups = (t1 - t0 + XW_GT - 1) // XW_GT
        last_slab_of_tile = {}
        for item in range(n_groups * n_slabs):
            grp, slab = divmod(item, n_slabs)
            tile0 = t0 + grp * XW_GT
            nt = min(XW_GT, t1 - tile0)
            assert 1 <= nt <= XW_GT
            r0 = slab * R
            rows_here = min(R, ldY - r0)
            assert 0 < rows_here <= R and rows_here % 4 == 0          # k-steps of 4 rows
            assert rows_here <= box_rows                              # the contraction stays inside the box
            assert r0 + rows_here <= ldY                              # ... and inside the padded matrix (zero rows beyond N)
            gene0 = tile0 * TG
            assert gene0 + nt * TG <= P_pad                           # stored / contracted tiles exist
            # box overhang only in the unused part: rows [r0 + rows_here, r0 + box_rows) and genes [gene0 + nt*TG, gene0 + box_genes)
            assert r0 < ldY and gene0 < P_pad
            for j in range(nt):
                key = (tile0 + j, slab)
                assert key not in seen, key
                seen[key] = b
                assert last_slab_of_tile.get(tile0 + j, -1) == slab - 1   # a tile's slabs are consecutive and ascending
                last_slab_of_tile[tile0 + j] = slab
        for t in range(t0, t1):
            assert last_slab_of_tile.get(t) == n_slabs - 1
    assert len(seen) == n_tiles * n_slabs


def test_box_is_the_conflict_free_pitch():
    # fragment loads read address (gene * pitch + row) with 8 genes x 4 rows per instruction: pitch % 8 == 4 spreads them over all 32 banks
    p = pitch4(R)
    assert p == 132 and p % 8 == 4 and p <= 256                       # a tensor-map box dimension is at most 256 elements
    from collections import Counter
    hits = Counter(((g * p + t) * 2) % 32 for g in range(8) for t in range(4))
    assert len(hits) == 16 and set(hits.values()) == {2}              # 32 doubles = 2 wavefronts: every bank pair exactly once per wavefront
