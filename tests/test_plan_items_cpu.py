"""CPU suite: invariants of the work-item cutter (rspmm_plan.cu) through the debug C entry point, which runs the HOST
MIRROR of the device cutter - the same block-walk / visitor functions the device kernels execute, block after block
(tests/test_plan_device_gpu.py checks on the GPU that the device kernels produce exactly this).  No GPU involved."""
import ctypes

import numpy as np
import pytest

MERGED = 1 << 30


def cut(seg_ptr, ITEM_EDGES=64):
    from ultra_b200 import build, _capi
    build.build()
    lib = _capi.lib()
    assert ITEM_EDGES == 64  # the finer-item tunable of round 1 was measured (no gain at any size) and removed
    seg_ptr = np.ascontiguousarray(seg_ptr, dtype=np.int32)
    S = len(seg_ptr) - 1
    E = int(seg_ptr[-1])
    cap = 5 * ((E + ITEM_EDGES - 1) // ITEM_EDGES) + S // 64 + 16
    items = np.zeros((cap, 4), np.int32)
    fix = np.zeros((max(S, 1), 4), np.int32)
    nonempty = np.zeros(S + 1, np.int32)
    n = [ctypes.c_int(0) for _ in range(5)]
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.ultra_rspmm_debug_cut_items(p(seg_ptr), S, p(items), cap, ctypes.byref(n[0]), p(fix), max(S, 1),
                                         ctypes.byref(n[1]), ctypes.byref(n[2]), ctypes.byref(n[3]), p(nonempty), S + 1,
                                         ctypes.byref(n[4]))
    assert rc == 0, lib.ultra_rspmm_last_error()
    return (items[:n[0].value], fix[:n[1].value], n[2].value, n[3].value, nonempty[:n[4].value])


def degrees_cases():
    rng = np.random.default_rng(0)
    yield "empty graph", np.zeros(5, np.int64)
    yield "single edge", np.array([0, 1, 0])
    yield "all short", rng.integers(0, 9, 500)
    yield "exactly 64 / 65 / 256 / 257", np.array([64, 65, 0, 256, 257, 1, 0, 63])
    yield "one hub", np.concatenate([rng.integers(0, 5, 100), [36008], rng.integers(0, 5, 100)])
    yield "zipf", (20000 / np.arange(1, 400)).astype(np.int64)
    yield "uniform 37", rng.poisson(37, 3000)
    yield "relation-graph like", np.full(474, 1896)
    yield "many cutter blocks, short segments", rng.integers(0, 3, 5000)
    yield "block boundaries at 64-segment multiples", np.tile(np.array([10] * 63 + [70]), 9)


@pytest.mark.parametrize("ITEM_EDGES", [64], ids=["items64"])
@pytest.mark.parametrize("name,deg", list(degrees_cases()), ids=[n for n, _ in degrees_cases()])
def test_cut_items_invariants(name, deg, ITEM_EDGES):
    check_invariants(deg, ITEM_EDGES)


@pytest.mark.parametrize("ITEM_EDGES", [64], ids=["items64"])
def test_cut_items_invariants_randomized(ITEM_EDGES):
    """Seeded random degree sequences: mixtures of empty, short, exactly-full, long and very long segments."""
    rng = np.random.default_rng(100 + ITEM_EDGES)
    for _ in range(25):
        S = int(rng.integers(1, 400))
        kind = rng.integers(0, 5, S)
        deg = np.where(kind == 0, 0, np.where(kind == 1, rng.integers(1, ITEM_EDGES, S),
                       np.where(kind == 2, ITEM_EDGES * rng.integers(1, 6, S),
                                np.where(kind == 3, rng.integers(ITEM_EDGES, 6 * ITEM_EDGES, S),
                                         rng.integers(1, 3, S)))))
        if rng.random() < 0.3:
            deg[int(rng.integers(0, S))] = int(rng.integers(1000, 20000))
        check_invariants(deg.astype(np.int64), ITEM_EDGES)


def check_invariants(deg, ITEM_EDGES):
    seg_ptr = np.concatenate([[0], np.cumsum(deg)])
    S, E = len(deg), int(seg_ptr[-1])
    items, fix, nlong, nslots, nonempty = cut(seg_ptr, ITEM_EDGES)
    assert len(items) % 4 == 0
    # every edge is covered exactly once
    cover = np.zeros(E + 1, np.int64)
    for e0, e1, z, slot in items:
        assert 0 <= e0 <= e1 <= E and e1 - e0 <= ITEM_EDGES
        cover[e0] += 1
        cover[e1] -= 1
    assert (np.cumsum(cover)[:E] == 1).all()
    # non-empty list
    ne_ids = np.nonzero(deg)[0]
    assert list(nonempty[:-1]) == list(ne_ids) and nonempty[-1] == 0
    seg_of_edge = np.repeat(np.arange(S), deg)
    slots_seen = {}
    merged_seen = True
    for k, (e0, e1, z, slot) in enumerate(items):
        j, merged = z & (MERGED - 1), bool(z & MERGED)
        if e1 == e0:
            continue
        segs = seg_of_edge[e0:e1]
        if slot == -1:   # whole segments: starts and ends on segment boundaries, first segment is nonempty[j]
            assert not merged
            assert e0 == seg_ptr[segs[0]] and e1 == seg_ptr[segs[-1] + 1]
            assert nonempty[j] == segs[0]
            assert all(deg[s] <= ITEM_EDGES for s in np.unique(segs))
        else:            # chunk of ONE long segment
            assert (segs == segs[0]).all() and deg[segs[0]] > ITEM_EDGES and nonempty[j] == segs[0]
            if merged:
                assert merged_seen, "merged groups must come first"
                assert e1 - e0 == ITEM_EDGES
                lead = k - k % 4
                assert all(items[lead + t][2] & MERGED for t in range(4))
                assert items[lead][3] >= 0 and all(items[lead + t][3] == -2 for t in range(1, 4))
                assert all(items[lead + t][0] == items[lead][0] + t * ITEM_EDGES for t in range(4))
                if k == lead:
                    slots_seen.setdefault(int(segs[0]), []).append((e0, int(slot)))
            else:
                assert slot >= 0
                slots_seen.setdefault(int(segs[0]), []).append((e0, int(slot)))
        if not merged:
            merged_seen = False
    # fix list: long segments first (their slots are consecutive and in edge order), then empty segments
    long_ids = [s for s in range(S) if deg[s] > ITEM_EDGES]
    assert nlong == len(long_ids) and list(fix[:nlong, 0]) == long_ids
    assert sorted(fix[nlong:, 0]) == [s for s in range(S) if deg[s] == 0]
    total = 0
    for s, slot0, cnt, _ in fix[:nlong]:
        got = [sl for _, sl in sorted(slots_seen[int(s)])]
        assert got == list(range(slot0, slot0 + cnt))
        total += cnt
    assert total == nslots
    # whole-segment items after the merged groups are sorted longest first (lock-step balance)
    rest = [(-(e1 - e0), e0) for e0, e1, z, slot in items if not (z & MERGED) and e1 > e0]
    assert rest == sorted(rest)
