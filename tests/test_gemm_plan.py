"""Work list of the persistent dense-layer kernel (whole-tile rounds + stream-K remainder), checked on the host.

The same `Sched` code runs inside `gemm_tc2p_kernel` (csrc/gemm_tcgen05_persist.cuh); `mfm_debug_gemm_plan` evaluates
it on the CPU.  Properties: every (tile, k-block) is computed exactly once; every tile has exactly one pair that
runs its epilogue and that pair lists exactly the pairs holding the tile's other k-ranges; a pair contributes at
most once (one workspace slot per pair), as its FIRST item (no dependency chains); the part that finishes a tile
comes second and the whole tiles last (the fix-up epilogue hides under their main loops).
"""
import ctypes

import numpy as np
import pytest

BK, TILE_M, TILE_N = 32, 256, 256
FULL, CONTRIB, FINISH = 0, 1, 2


def plan(lib, M, N, K, pairs, streamk=1):
    cap = 1 << 16
    buf = (ctypes.c_int * (7 * cap))()
    n = lib.mfm_debug_gemm_plan(M, N, K, pairs, streamk, buf, cap)
    assert 0 <= n <= cap
    return np.frombuffer(buf, dtype=np.int32, count=7 * n).reshape(n, 7).copy()


SHAPES = [(8192, 1024, 1024), (8192, 1600, 1024), (8192, 1024, 1600), (8192, 1024, 256), (16384, 1024, 1024),
          (65536, 1024, 1024), (65536, 1600, 1024), (1024, 1024, 1600), (4096 + 77, 1088, 520), (256, 256, 4096),
          (300, 200, 100), (256, 64, 32), (5000, 1600, 1600), (12345, 1024, 288), (74 * 256, 256, 1024)]


@pytest.mark.parametrize("pairs", [74, 66, 2])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_plan_covers_every_k_block_once(lib, M, N, K, pairs):
    rows = plan(lib, M, N, K, pairs)
    mt, nt, KT = -(-M // TILE_M), -(-N // TILE_N), -(-K // BK)
    total = mt * nt
    cover = np.zeros((total, KT), dtype=np.int32)
    finisher = {}
    slots = {}
    per_pair = {}
    for pair, tile, kb0, kb1, kind, c_first, c_count in rows:
        assert 0 <= tile < total and 0 <= kb0 < kb1 <= KT
        cover[tile, kb0:kb1] += 1
        per_pair.setdefault(pair, []).append((tile, kb0, kb1, kind, c_first, c_count))
        if kind == CONTRIB:
            assert kb1 < KT
            assert pair not in slots, "a pair has one workspace slot"
            slots[pair] = (tile, kb0, kb1)
        else:
            assert kb1 == KT
            assert tile not in finisher
            finisher[tile] = (pair, kb0, c_first, c_count)
            assert (kind == FINISH) == (kb0 > 0)
    assert (cover == 1).all()
    assert sorted(finisher) == list(range(total))
    # the finishing pair names exactly the pairs that hold the rest of its tile
    for tile, (pair, kb0, c_first, c_count) in finisher.items():
        holders = sorted(q for q, (t, _, _) in slots.items() if t == tile)
        assert holders == list(range(c_first, c_first + c_count)) if kb0 > 0 else holders == []
        assert all(q < pair for q in holders)
        assert c_count <= 5
    # contributions come before the finishing item of the same pair; the remainder's items before the whole tiles
    for pair, items in per_pair.items():
        kinds = [k for (_, _, _, k, _, _) in items]
        if CONTRIB in kinds and FINISH in kinds:
            assert kinds.index(CONTRIB) < kinds.index(FINISH)
        base = (total // pairs) * pairs
        partial = [i for i, (t, a, b, k, _, _) in enumerate(items) if k != FULL or (a, b) != (0, KT)]
        assert len(partial) <= 2
        if partial:                       # stream-K plan: [contribution] [finishing part] whole tiles ...
            rem_items = [i for i, (t, _, _, _, _, _) in enumerate(items) if t >= base]
            assert rem_items == list(range(len(rem_items))) and len(rem_items) <= 2


def test_plan_small_remainders_are_cut_four_ways(lib):
    # 8 192 chains x 1 600 columns: 224 tiles = 3 whole rounds + 2 tiles; 8 pairs share them, a quarter tile each
    rows = plan(lib, 8192, 1600, 1024, 74)
    rem = [r for r in rows if r[1] >= 3 * 74]
    assert len({r[0] for r in rem}) == 8
    assert max(r[3] - r[2] for r in rem) == 8
    work = np.zeros(74, dtype=np.int64)
    for pair, tile, kb0, kb1, *_ in rows:
        work[pair] += kb1 - kb0
    assert work.max() == 3 * 32 + 8 and work.min() == 3 * 32
    # a handful of active chains (end of an ODE solve): 4 tiles on 16 pairs
    rows = plan(lib, 256, 1024, 1024, 74)
    assert len({r[0] for r in rows}) == 16 and sum(r[3] - r[2] for r in rows) == 4 * 32
    # K = 1 600: 50 k-blocks do not divide by 4 -> ranges of 12 and 13 k-blocks, still one item per pair (4 ranges per tile)
    rows = plan(lib, 1024, 1024, 1600, 74)
    assert len(rows) == 64 and len({r[0] for r in rows}) == 64
    assert sorted({r[3] - r[2] for r in rows}) == [12, 13]


def test_plan_large_remainders_stay_whole_tiles(lib):
    # measured slower than whole tiles (epilogue-bound round), see gemm_tcgen05_persist.cuh
    for shape in [(8192, 1024, 1024), (16384, 1024, 1024), (65536, 1024, 1024)]:
        rows = plan(lib, *shape, 74)
        assert all(r[4] == FULL for r in rows)
    # too little K per range: whole tiles only
    rows = plan(lib, 8192, 1600, 256, 74)
    assert all(r[4] == FULL for r in rows)
    # a full last round: nothing to balance
    rows = plan(lib, 74 * 256, 256, 1024, 74)
    assert all(r[4] == FULL for r in rows) and len(rows) == 74
    # switched off
    rows = plan(lib, 8192, 1600, 1024, 74, streamk=0)
    assert all(r[4] == FULL for r in rows)
