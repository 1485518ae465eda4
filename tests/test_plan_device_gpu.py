"""The device-side work-item cutter (-m gpu): plan_build cuts items with CUDA kernels (count -> scan -> write -> sort ->
finalize, no host round trip); here its output - items, fix list, non-empty ids - is compared EXACTLY with the host
mirror of the same functions (ultra_rspmm_debug_cut_items), for all three edge orders of several graphs."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def host_cut(seg_ptr):
    from ultra_b200 import _capi
    lib = _capi.lib()
    seg_ptr = np.ascontiguousarray(seg_ptr, dtype=np.int32)
    S, E = len(seg_ptr) - 1, int(seg_ptr[-1])
    cap = 5 * ((E + 63) // 64) + S // 64 + 16
    items, fix, nonempty = np.zeros((cap, 4), np.int32), np.zeros((max(S, 1), 4), np.int32), np.zeros(S + 1, np.int32)
    n = [ctypes.c_int(0) for _ in range(5)]
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    rc = lib.ultra_rspmm_debug_cut_items(p(seg_ptr), S, p(items), cap, ctypes.byref(n[0]), p(fix), max(S, 1),
                                         ctypes.byref(n[1]), ctypes.byref(n[2]), ctypes.byref(n[3]), p(nonempty), S + 1,
                                         ctypes.byref(n[4]))
    assert rc == 0, lib.ultra_rspmm_last_error()
    return items[:n[0].value], fix[:n[1].value], n[2].value, n[3].value, nonempty[:n[4].value]


def device_read(plan, order, what, count, width):
    from ultra_b200 import _capi
    out = np.zeros((max(count, 1), width), np.int32)
    rc = _capi.lib().ultra_rspmm_plan_debug_read(plan.handle, order, what, out.ctypes.data_as(ctypes.c_void_p), out.nbytes)
    assert rc == 0, _capi.lib().ultra_rspmm_last_error()
    return out[:count]


def graphs():
    from ultra_b200.synth import synth_kg, synth_scale, relation_graph_of
    g = torch.Generator().manual_seed(3)
    yield "tiny", torch.randint(0, 9, (2, 40), generator=g), torch.randint(0, 3, (40,), generator=g), 12, 3
    ei = torch.randint(0, 300, (2, 9000), generator=g)
    ei[0, :2500] = 7  # a hub row: merged groups + single chunks
    yield "hub", ei, torch.randint(0, 5, (9000,), generator=g), 305, 5
    yield "no edges", torch.zeros(2, 0, dtype=torch.long), torch.zeros(0, dtype=torch.long), 5, 2
    ei, et, R = synth_kg(5000, 20, 60000, alpha=1.0, seed=5)
    yield "zipf 5k", ei, et, 5000, R
    ei, et, R, N = synth_scale("fb15k237", alpha=1.0, seed=1236)
    yield "fb15k237 scale", ei, et, N, R
    rei, ret, rR, rN = relation_graph_of(ei, et, N, R)
    yield "relation graph", rei, ret, rN, rR


@pytest.mark.parametrize("name", ["tiny", "hub", "no edges", "zipf 5k", "fb15k237 scale", "relation graph"])
def test_device_cutter_equals_host_mirror(cuda_device, name):
    from ultra_b200.plan import GraphPlan
    ei, et, N, R = next(g[1:] for g in graphs() if g[0] == name)
    plan = GraphPlan(ei.to(cuda_device), et.to(cuda_device), N, R)
    for order, S in ((0, N), (1, N), (2, R)):
        seg_ptr = device_read(plan, order, 3, S + 1, 1)[:, 0]
        # the CSR pointer itself: against numpy
        key = (ei[0], ei[1], et)[order].numpy()
        assert np.array_equal(seg_ptr, np.concatenate([[0], np.cumsum(np.bincount(key, minlength=S))]))
        items, fix, nlong, nslots, nonempty = host_cut(seg_ptr)
        n_items = plan.info(3) if order == 0 else plan.info(7 + order)
        assert n_items == len(items), (name, order)
        assert np.array_equal(device_read(plan, order, 0, n_items, 4), items), (name, order, "items")
        assert plan.info(10 + 3 * order) == len(fix) and plan.info(11 + 3 * order) == nlong
        assert np.array_equal(device_read(plan, order, 1, len(fix), 4), fix), (name, order, "fix")
        assert plan.info(4 + order) == nslots
        got_ne = device_read(plan, order, 2, S + 1, 1)[:len(nonempty), 0]
        assert np.array_equal(got_ne, nonempty), (name, order, "nonempty")


@pytest.mark.parametrize("name", ["tiny", "hub", "zipf 5k", "relation graph"])
def test_segment_edge_order(cuda_device, name):
    """Inside a segment the edges are sorted by the gathered node id (sweep locality of the row gathers); graphs small
    enough for shared-memory-resident tiles ((N + R) * 128 B <= 78 KB) are sorted by relation type first, the order the
    relation-run reuse of the small-graph kernel needs (include/ultra_rspmm.h, tunable plan_relation_runs).  Either way
    the sequence is a permutation of the caller's edges, grouped by the primary key."""
    from ultra_b200.plan import GraphPlan
    ei, et, N, R = next(g[1:] for g in graphs() if g[0] == name)
    plan = GraphPlan(ei.to(cuda_device), et.to(cuda_device), N, R)
    E = ei.shape[1]
    small = (N + R) * 128 <= 78 * 1024
    row, col, typ = ei[0].numpy(), ei[1].numpy(), et.numpy()
    for order in (0, 1, 2):
        eid = device_read(plan, order, 5, E, 1)[:, 0]
        assert np.array_equal(np.sort(eid), np.arange(E)), (name, order)
        if order == 2:
            keys = (row[eid], typ[eid])                      # np.lexsort: last key is the primary one
        else:
            pri, sec = (row, col) if order == 0 else (col, row)
            keys = (sec[eid], typ[eid], pri[eid]) if small else (sec[eid], pri[eid])
        k = np.stack(keys[::-1], axis=1).astype(np.int64)    # columns: primary first
        flat = np.zeros(E, np.int64)
        for c in range(k.shape[1]):
            flat = flat * (max(N, R) + 1) + k[:, c]
        assert np.all(np.diff(flat) >= 0), (name, order, "edges not sorted by %d keys" % k.shape[1])
