"""Work decomposition of the tiled kernels (host logic, no GPU): for the launch geometries the single- and multi-GPU runs use --
whole volumes, the edge / middle split of the NCCL slab schedule and the three face-tagged ranges of peer mode at 2, 4 and 8 ranks
-- every plane of every range is covered exactly once per tile, tiles cover the slice, items are issued range by range, and the
static round-robin assignment is reasonably balanced."""
import ctypes as C

import numpy as np
import pytest

TILE = {0: (32, 16), 1: (64, 24)}          # pass A / pass B tile (solver_tiled.cu: pa::TX x pa::TY, pb::TX x pb::TY)
COST = {0: (2, 0.5), 1: (6, 0.35)}         # halo planes per item and their relative cost in make_sched


def schedule(built, pas, dims, ranges, sms=148):
    from sobfu_b200 import _capi
    X, Y, Z = dims
    n = len(ranges)
    lo = (C.c_int * 3)(*[r[0] for r in ranges] + [0] * (3 - n))
    hi = (C.c_int * 3)(*[r[1] for r in ranges] + [0] * (3 - n))
    face = (C.c_int * 3)(*[r[2] for r in ranges] + [0] * (3 - n))
    cap = 1 << 16
    items = (C.c_int * (6 * cap))()
    n_items, grid = C.c_int(), C.c_int()
    rc = _capi.lib().sobfu_b200_debug_schedule(pas, X, Y, Z, n, lo, hi, face, sms, items, cap, C.byref(n_items), C.byref(grid))
    assert rc == 0 and n_items.value <= cap
    return np.array(items[:6 * n_items.value], dtype=np.int64).reshape(-1, 6), grid.value


def peer_ranges(built, pas, dims, lo, hi, has_lo, has_hi, sms=148):
    from sobfu_b200 import _capi
    out, n = (C.c_int * 9)(), C.c_int()
    assert _capi.lib().sobfu_b200_debug_peer_ranges(pas, dims[0], dims[1], dims[2], lo, hi, int(has_lo), int(has_hi), sms, out, C.byref(n)) == 0
    return [(out[3 * r], out[3 * r + 1], out[3 * r + 2]) for r in range(n.value)]


def check(built, pas, dims, ranges, max_imbalance):
    X, Y, Z = dims
    TX, TY = TILE[pas]
    it, grid = schedule(built, pas, dims, ranges)
    ctas = 148 * (1 if pas == 1 else 4)
    assert grid == min(len(it), ctas) and np.array_equal(it[:, 0], np.arange(len(it)) % grid)
    tiles = sorted({(int(a), int(b)) for a, b in it[:, 1:3]})
    assert tiles == sorted((x, y) for x in range(0, X, TX) for y in range(0, Y, TY))
    pos = 0
    for lo, hi, face in ranges:                       # items are issued range by range, chunk by chunk, tile by tile
        if hi <= lo:
            continue
        end = pos
        while end < len(it) and it[end, 5] == face and lo <= it[end, 3] and it[end, 4] <= hi:
            end += 1
        part = it[pos:end]
        pos = end
        assert len(part) % len(tiles) == 0 and len(part) > 0, (ranges, lo, hi)
        nch = len(part) // len(tiles)
        chunks = part.reshape(nch, len(tiles), 6)
        assert (chunks[:, :, 3] == chunks[:, :1, 3]).all() and (chunks[:, :, 4] == chunks[:, :1, 4]).all()     # a chunk spans every tile
        assert sorted({(int(a), int(b)) for a, b in chunks[0, :, 1:3]}) == sorted(tiles)
        zb, ze = chunks[:, 0, 3], chunks[:, 0, 4]
        assert zb[0] == lo and ze[-1] == hi and (zb[1:] == ze[:-1]).all() and (ze > zb).all(), (lo, hi, zb, ze)  # exact cover of [lo, hi)
    assert pos == len(it)
    hp, hc = COST[pas]
    load = np.zeros(grid)
    np.add.at(load, it[:, 0], (it[:, 4] - it[:, 3]) + hp * hc + 1.0)
    assert load.max() <= max_imbalance * load.sum() / ctas, (dims, ranges, load.max(), load.sum() / ctas)
    return it


@pytest.mark.parametrize("dim", [32, 64, 96, 128, 192, 256, 512])
def test_whole_volume_schedules(built, dim):
    TILE[2], COST[2] = TILE[0], (3, 0.6)         # the default pass A: same tile, one more pipeline-fill step per item
    for pas in (0, 1, 2):
        it = check(built, pas, (dim, dim, dim), [(0, dim, 0)], max_imbalance=1.35 if dim >= 256 else 1e9)
        if dim >= 256:
            assert (it[:, 4] - it[:, 3]).min() >= 16          # big volumes keep chunks of >= 16 planes
    # partial tiles in x and y
    check(built, 0, (40, 36, 32), [(0, 32, 0)], 1e9)
    check(built, 1, (96, 40, 36), [(0, 36, 0)], 1e9)


@pytest.mark.parametrize("nranks", [2, 4, 8])
@pytest.mark.parametrize("dim", [256, 512])
def test_slab_schedules(built, nranks, dim):
    n = dim // nranks
    for rank in (0, nranks // 2, nranks - 1):
        has_lo, has_hi = rank > 0, rank < nranks - 1
        lo, hi = (-3 if has_lo else 0), (n + 3 if has_hi else n)
        dims = (dim, dim, n)
        # NCCL schedule: A_mid | A_edge (two ranges) | B_edge (two ranges) | B_mid
        check(built, 0, dims, [(1, n - 1, 0)], 1.9)
        check(built, 0, dims, [(lo, 1, 0), (n - 1, hi, 0)], 1e9)
        check(built, 1, dims, [(0, 4, 0), (n - 4, n, 0)], 1e9)
        check(built, 1, dims, [(4, n - 4, 0)], 1.9)
        # peer schedule: | lower face chunk | upper face chunk | middle | per launch, the face chunks first and ONE z chunk each
        ra, rb = peer_ranges(built, 0, dims, lo, hi, has_lo, has_hi), peer_ranges(built, 1, dims, 0, n, has_lo, has_hi)
        for r, (l, h) in ((ra, (lo, hi)), (rb, (0, n))):
            faces = [x for x in r if x[2]]
            assert [x[2] for x in faces] == ([1] if has_lo else []) + ([2] if has_hi else [])
            assert r[:len(faces)] == faces                                   # issued first
            assert all(x[1] - x[0] >= 4 for x in faces)                      # the planes a neighbour needs / that read halo planes
            cover = sorted((x[0], x[1]) for x in r)
            assert cover[0][0] == l and cover[-1][1] == h and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
        a = check(built, 0, dims, ra, 2.0)
        b = check(built, 1, dims, rb, 2.0)
        # what the neighbours count: tiles_x * tiles_y items per face, on every rank
        for f, on in ((1, has_lo), (2, has_hi)):
            assert (a[:, 5] == f).sum() == ((dim // 32) * (dim // 16) if on else 0)
            assert (b[:, 5] == f).sum() == (-(-dim // 64) * -(-dim // 24) if on else 0)


def test_bad_arguments(built):
    from sobfu_b200 import _capi
    z = (C.c_int * 3)()
    n, g = C.c_int(), C.c_int()
    assert _capi.lib().sobfu_b200_debug_schedule(3, 64, 64, 64, 1, z, z, z, 148, None, 0, C.byref(n), C.byref(g)) != 0
    assert _capi.lib().sobfu_b200_debug_schedule(0, 64, 64, 64, 4, z, z, z, 148, None, 0, C.byref(n), C.byref(g)) != 0
