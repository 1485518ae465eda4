"""Edge masks on a built plan (-m gpu): ultra_rspmm_plan_apply_mask re-cuts the plan for a sub-graph on the device, with
no sort and no host synchronisation - the incremental path for the training loop's per-step edge removal
(remove_easy_edges, /root/reference/ultra/base_nbfnet.py:54-77).  Because the compaction preserves every order and the same
cutter runs on the same segment pointers, the masked plan must compute BIT-identically to a plan built from the filtered
edge list - which is what these tests assert (and, through that plan, parity with the oracle)."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import random_case, oracle_all, assert_close_ref, assert_equal_values

pytestmark = pytest.mark.gpu


def _run(ei, et, w, rel, x, og, sum, mul, need_wgrad=False):
    from ultra_b200.rspmm import generalized_rspmm
    relg, xg = rel.clone().requires_grad_(), x.clone().requires_grad_()
    wg = w.clone().requires_grad_() if need_wgrad else w
    out = generalized_rspmm(ei, et, wg, relg, xg, sum=sum, mul=mul)
    out.backward(og)
    return out.detach(), relg.grad, xg.grad, (wg.grad if need_wgrad else None)


def _mask(E, seed, frac, device, ei=None):
    g = torch.Generator().manual_seed(seed)
    keep = torch.rand(E, generator=g) >= frac
    if ei is not None:  # empty one whole row and thin out the hub (breaks merged chunk groups into single chunks)
        keep &= ei[0] != 5
        hub = (ei[0] == 1).nonzero().flatten()
        keep[hub[torch.rand(len(hub), generator=g) < 0.4]] = False
    return keep.to(device)


@pytest.mark.parametrize("sum,mul", [("add", "mul"), ("max", "add"), ("min", "mul")])
def test_masked_plan_is_bit_identical_to_a_plan_of_the_filtered_edges(cuda_device, sum, mul):
    from ultra_b200.plan import plan_cache
    d = cuda_device
    ei, et, w, rel, x, og = random_case(81, 400, 30000, 6, 192, hub=4000, unit_weight=True)
    eid, etd, wd, reld, xd, ogd = (t.to(d) for t in (ei, et, w, rel, x, og))
    base = _run(eid, etd, wd, reld, xd, ogd, sum, mul)
    plan = plan_cache.get(eid, etd, 400, 6)
    for seed, frac in ((1, 0.05), (2, 0.5)):
        keep = _mask(ei.shape[1], seed, frac, d, ei)
        plan.apply_edge_mask(keep)
        try:
            got = _run(eid, etd, wd, reld, xd, ogd, sum, mul)
        finally:
            plan.apply_edge_mask(None)
        fei, fet = eid[:, keep].contiguous(), etd[keep].contiguous()
        exp = _run(fei, fet, torch.ones(fei.shape[1], device=d), reld, xd, ogd, sum, mul)
        for name, a, b in zip(("output", "relation_grad", "input_grad"), got, exp):
            assert torch.equal(a, b), "%s differs under mask (seed %d)" % (name, seed)
        # and against the oracle on the filtered list
        k = keep.cpu()
        e_out, _, e_rg, e_ig = oracle_all(sum, mul, ei[:, k], et[k], torch.ones(int(k.sum())), rel, x, og)
        (assert_close_ref if sum == "add" else assert_equal_values)(got[0], e_out, "masked forward vs oracle")
        assert_close_ref(got[2], e_ig, "masked input_grad vs oracle")
    # mask cleared: the base graph again, bit for bit
    again = _run(eid, etd, wd, reld, xd, ogd, sum, mul)
    for a, b in zip(again[:3], base[:3]):
        assert torch.equal(a, b)


def test_masked_plan_arrays_equal_the_host_mirror(cuda_device):
    """Segment pointers of the masked orders against numpy, the re-cut items / fix list against the host mirror."""
    from ultra_b200 import _capi
    from ultra_b200.plan import GraphPlan
    from test_plan_device_gpu import host_cut
    d = cuda_device
    ei, et, _, _, _, _ = random_case(82, 300, 20000, 5, 32, hub=3000, unit_weight=True)
    plan = GraphPlan(ei.to(d), et.to(d), 300, 5)
    keep = _mask(ei.shape[1], 3, 0.3, d, ei)
    plan.apply_edge_mask(keep)
    torch.cuda.synchronize()
    k = keep.cpu().numpy()
    lib = _capi.lib()

    def read(order, what, count, width):
        out = np.zeros((max(count, 1), width), np.int32)
        rc = lib.ultra_rspmm_plan_debug_read(plan.handle, order, what, out.ctypes.data_as(ctypes.c_void_p), out.nbytes)
        assert rc == 0, lib.ultra_rspmm_last_error()
        return out[:count]
    for order, S in ((0, 300), (1, 300), (2, 5)):
        key = (ei[0], ei[1], et)[order].numpy()[k]
        seg_ptr = read(order, 3, S + 1, 1)[:, 0]
        assert np.array_equal(seg_ptr, np.concatenate([[0], np.cumsum(np.bincount(key, minlength=S))]))
        counts = read(order, 4, 8, 1)[:, 0]
        items, fix, nlong, nslots, nonempty = host_cut(seg_ptr)
        assert counts[0] == len(items) and counts[1] == len(fix) and counts[2] == nlong and counts[3] == nslots
        assert counts[7] == int(k.sum())
        assert np.array_equal(read(order, 0, len(items), 4), items)
        assert np.array_equal(read(order, 1, len(fix), 4), fix)


def test_weight_grad_under_mask_keeps_base_order(cuda_device):
    from ultra_b200.plan import plan_cache
    d = cuda_device
    ei, et, w, rel, x, og = random_case(83, 120, 3000, 4, 64, hub=500)
    eid, etd, wd, reld, xd, ogd = (t.to(d) for t in (ei, et, w, rel, x, og))
    keep = _mask(ei.shape[1], 4, 0.3, d)
    plan = plan_cache.get(eid, etd, 120, 4)
    plan.apply_edge_mask(keep)
    try:
        out, rg, ig, wgrad = _run(eid, etd, wd, reld, xd, ogd, "add", "mul", need_wgrad=True)
    finally:
        plan.apply_edge_mask(None)
    k = keep.cpu()
    e_out, e_wg, e_rg, e_ig = oracle_all("add", "mul", ei[:, k], et[k], w[k], rel, x, og)
    assert_close_ref(out, e_out, "masked weighted forward")
    assert wgrad.shape == w.shape and bool((wgrad[~keep] == 0).all())
    assert_close_ref(wgrad[keep], e_wg, "weight_grad of the kept edges, base order")


def test_apply_mask_forward_backward_capture_as_one_cuda_graph(cuda_device):
    """No host synchronisation anywhere: mask application + forward + backward captured once, replayed with new masks
    written into the static mask tensor."""
    from ultra_b200.plan import plan_cache
    from ultra_b200.rspmm import generalized_rspmm
    d = cuda_device
    ei, et, w, rel, x, og = random_case(84, 200, 8000, 5, 128, hub=1500, unit_weight=True)
    eid, etd, wd, ogd = ei.to(d), et.to(d), w.to(d), og.to(d)
    s_rel, s_x = rel.to(d).requires_grad_(), x.to(d).requires_grad_()
    s_keep = torch.ones(ei.shape[1], dtype=torch.bool, device=d)
    plan = plan_cache.get(eid, etd, 200, 5)

    def step():
        plan.apply_edge_mask(s_keep)
        out = generalized_rspmm(eid, etd, wd, s_rel, s_x, sum="add", mul="mul")
        out.backward(ogd)
        return out
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
            s_rel.grad = s_x.grad = None
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_out = step()
    g_rg, g_ig = s_rel.grad, s_x.grad
    try:
        for seed in (5, 6):
            keep = _mask(ei.shape[1], seed, 0.2, d, ei)
            s_keep.copy_(keep)
            graph.replay()
            k = keep.cpu()
            e_out, _, e_rg, e_ig = oracle_all("add", "mul", ei[:, k], et[k], torch.ones(int(k.sum())), rel, x, og)
            assert_close_ref(g_out.detach(), e_out, "replayed masked forward")
            assert_close_ref(g_ig, e_ig, "replayed masked input_grad")
            assert_close_ref(g_rg, e_rg, "replayed masked relation_grad")
    finally:
        plan.apply_edge_mask(None)


@pytest.mark.parametrize("dtype,d", [(torch.float32, 64), (torch.float64, 16)])
def test_masked_plan_head_major_and_fp64(cuda_device, dtype, d):
    """The head-major entry points (the layer's native (B, N, d) layout, boundary folded) and fp64 under an edge mask:
    bit-identical to the same call on the filtered edge list."""
    from ultra_b200.fused import rspmm_heads
    from ultra_b200.plan import plan_cache
    dev = cuda_device
    g = torch.Generator().manual_seed(91)
    B, N, R, E = 4, 150, 6, 6000
    ei = torch.randint(0, N - 2, (2, E), generator=g)
    ei[0, :1500] = 3
    et = torch.randint(0, R, (E,), generator=g)
    rel = torch.randn(B, R, d, generator=g, dtype=dtype).to(dev)
    x = torch.randn(B, N, d, generator=g, dtype=dtype).to(dev)
    bd = torch.randn(B, N, d, generator=g, dtype=dtype).to(dev)
    og = torch.randn(B, N, d, generator=g, dtype=dtype).to(dev)
    eid, etd = ei.to(dev), et.to(dev)
    w = torch.ones(E, dtype=dtype, device=dev)
    keep = _mask(E, 7, 0.25, dev, ei)

    def run(e_i, e_t, ww, sum):
        r, xx = rel.clone().requires_grad_(), x.clone().requires_grad_()
        out = rspmm_heads(e_i, e_t, ww, r, xx, bd, sum=sum, mul="mul")
        out.backward(og)
        return out.detach(), r.grad, xx.grad
    plan = plan_cache.get(eid, etd, N, R)
    for sum in ("add", "max"):
        plan.apply_edge_mask(keep)
        try:
            got = run(eid, etd, w, sum)
        finally:
            plan.apply_edge_mask(None)
        exp = run(eid[:, keep].contiguous(), etd[keep].contiguous(), torch.ones(int(keep.sum()), dtype=dtype, device=dev), sum)
        for a, b in zip(got, exp):
            assert torch.equal(a, b)


def test_plan_cache_is_bounded_by_bytes(cuda_device):
    """a caller that hands a fresh edge tensor to every call (the reference's training loop) must not pin an
    unbounded number of stale plan workspaces: the LRU evicts by device bytes as well as by entries."""
    from ultra_b200.plan import PlanCache
    dev = cuda_device
    cache = PlanCache(capacity=64, max_bytes=40 << 20)
    g = torch.Generator().manual_seed(92)
    for _ in range(12):
        ei = torch.randint(0, 2000, (2, 60000), generator=g).to(dev)   # a new tensor every time: ~9 MB of workspace each
        et = torch.randint(0, 10, (60000,), generator=g).to(dev)
        cache.get(ei, et, 2000, 10)
    total = sum(e[0].nbytes() for e in cache._entries.values())
    assert cache.misses == 12 and 1 <= len(cache._entries) < 12 and total <= (40 << 20) + (12 << 20)
