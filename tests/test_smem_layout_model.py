"""Executable form of the shared-memory layout argument behind `Prox3Plan` (csrc/kernels2.cuh).

Model (checked against an ncu capture: every 64-bit row access of the unmapped layout showed 4
wavefronts where 2 are ideal; after the remapping all of them show 2): a 64-bit shared-memory
access of a warp is served as two half-warp passes, and a pass takes as many wavefronts as the
largest number of distinct 8-byte words that fall on the same bank pair.  With 8 lanes per row a
half-warp holds two rows, so the two rows must sit 16 banks apart: lane groups 2i, 2i+1 take rows
i, i+8 (odd row stride P => 8 rows = 16 banks mod 32), and rows >= 8 of the Y / U tiles are
shifted by 8 elements."""

H, E, TPF, TR, P = 128, 16, 8, 16, 137


def wavefronts(addrs):
    total = 0
    for half in (addrs[:16], addrs[16:]):
        banks = {}
        for a in set(half):
            banks.setdefault(a % 16, set()).add(a)
        total += max(len(v) for v in banks.values())
    return total


def patterns(gmap, ybase):
    out = {}
    for warp in range(4):
        lanes = [(gmap((warp * 32 + l) // TPF), l % TPF) for l in range(32)]
        for p in range(E):
            out[('split a', warp, p)] = [g * P + (t + TPF * p) for g, t in lanes]
            out[('split b', warp, p)] = [g * P + (H - (t + TPF * p)) for g, t in lanes]
            out[('tile', warp, p)] = [ybase(g) + t + TPF * p for g, t in lanes]
            out[('stage-1 write', warp, p)] = [g * P + 17 * t + p for g, t in lanes]
        for i in range(2):
            for r in range(8):
                out[('stage-2 read', warp, i, r)] = [g * P + (t + 8 * i) + 17 * r for g, t in lanes]
    return out


def test_identity_mapping_is_two_way_conflicted():
    w = {k: wavefronts(a) for k, a in patterns(lambda g: g, lambda g: g * H).items()}
    assert set(w.values()) == {4}


def test_remapped_layout_is_conflict_free():
    remap = lambda g: ((g & 1) << 3) | (g >> 1)
    w = {k: wavefronts(a) for k, a in patterns(remap, lambda g: g * H + 8 * (g >> 3)).items()}
    assert set(w.values()) == {2}
    assert sorted(remap(g) for g in range(16)) == list(range(16))      # a permutation of the rows


def test_transposed_tile_accesses_keep_their_odd_stride():
    # the cp.async scatter and the post-split gather address (row, wf) with 16 rows per wf:
    # 16 lanes on 16 different rows must cover all 32 banks, which needs the odd stride P
    lanes = [r * P + 5 for r in range(16)]
    assert len({(2 * a) % 32 for a in lanes}) == 16
