"""GPU parity suite (-m gpu): the sm_100a kernels, called through the reference-facing Python surface (which goes
through the C ABI), against the golden vectors of the unmodified reference, the C oracle, and oracle/_ref.

Tolerance (north star): 1e-5 fp32 relative, atol = 1e-5 * max|expected| (see helpers.assert_close_ref);
min/max forward must be bit-exact by value; fp64 1e-12."""
import os

import numpy as np
import pytest
import torch

from helpers import (OPS, GOLDEN_CASES, load_golden, random_case, oracle_all, assert_close_ref,
                     assert_equal_values)

pytestmark = pytest.mark.gpu


def run_ours(sum, mul, ei, et, w, rel, x, og=None, device="cuda", weight_grad=True):
    from ultra_b200.rspmm import generalized_rspmm
    d = torch.device(device)
    ei, et = ei.to(d), et.to(d)
    # detach(): fresh leaves that share storage (keeps pinned-ness for the host-tensor test)
    w = w.to(d).detach().requires_grad_(og is not None and weight_grad)
    rel = rel.to(d).detach().requires_grad_(og is not None)
    x = x.to(d).detach().requires_grad_(og is not None)
    out = generalized_rspmm(ei, et, w, rel, x, sum=sum, mul=mul)
    if og is None:
        return out.detach().cpu()
    out.backward(og.to(d))
    wg = w.grad.cpu() if weight_grad else None
    return out.detach().cpu(), wg, rel.grad.cpu(), x.grad.cpu()


def check_against(sum, mul, ours, expected, what=""):
    out, wg, rg, ig = ours
    e_out, e_wg, e_rg, e_ig = expected
    if sum == "add":
        assert_close_ref(out, e_out, what + "forward")
    else:
        assert_equal_values(out, e_out, what + "forward")
    assert_close_ref(rg, e_rg, what + "relation_grad")
    assert_close_ref(ig, e_ig, what + "input_grad")
    if wg is not None:
        assert_close_ref(wg, e_wg, what + "weight_grad")


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("sum,mul", OPS)
def test_cuda_matches_reference_golden(cuda_device, case, sum, mul):
    g = load_golden()
    t = lambda k: torch.from_numpy(g[case + "/" + k])
    ours = run_ours(sum, mul, t("edge_index"), t("edge_type"), t("edge_weight"), t("relation"), t("input"),
                    t("output_grad"))
    k = "%s_%s/" % (sum, mul)
    check_against(sum, mul, ours, (t(k + "output"), t(k + "weight_grad"), t(k + "relation_grad"),
                                   t(k + "input_grad")), case + " ")


def test_hand_case_survey_8c(cuda_device):
    FLT_MAX = float(np.finfo(np.float32).max)
    ei = torch.tensor([[0, 0, 1], [1, 1, 2]])
    et = torch.tensor([0, 0, 1])
    w = torch.ones(3)
    rel = torch.tensor([[2., 3.], [1., 1.]])
    x = torch.tensor([[1., 1.], [5., 7.], [4., 4.]])
    exp = {"add": [[20, 42], [4, 4], [0, 0]], "min": [[10, 21], [4, 4], [FLT_MAX, FLT_MAX]],
           "max": [[10, 21], [4, 4], [-FLT_MAX, -FLT_MAX]]}
    for s in ("add", "min", "max"):
        out, wg, rg, ig = run_ours(s, "mul", ei, et, w, rel, x, torch.ones(3, 2))
        assert_equal_values(out, torch.tensor(exp[s], dtype=torch.float32), s)
        assert_equal_values(rg, torch.tensor([[10., 14.], [4., 4.]]), s)
        assert_equal_values(ig, torch.tensor([[0., 0.], [4., 6.], [1., 1.]]), s)
        assert_equal_values(wg, torch.tensor([31., 31., 8.]), s)


SHAPES = [
    # seed, N, E, R, D, dtype, hub, unit_weight
    (31, 64, 900, 5, 32, torch.float32, 0, False),       # one tile
    (32, 200, 6000, 11, 96, torch.float32, 1500, False),  # 3 tiles, hub row split in many chunks
    (33, 97, 1000, 7, 100, torch.float32, 0, True),      # D % 32 != 0 -> partial last tile; unit weights
    (34, 33, 400, 3, 7, torch.float32, 0, False),        # D % 4 != 0 -> generic kernel
    (35, 150, 5000, 900, 64, torch.float32, 300, False),  # R too large for the smem relation tile
    (36, 80, 2000, 6, 48, torch.float64, 400, False),    # fp64 vector path
    (37, 80, 500, 6, 5, torch.float64, 0, False),        # fp64 generic path
    (38, 5, 0, 2, 32, torch.float32, 0, False),          # no edges at all: every row is the identity
    (39, 3000, 4000, 4, 64, torch.float32, 0, False),    # mostly empty / degree-1 rows
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "N%d_E%d_R%d_D%d_%s" % (s[1], s[2], s[3], s[4], str(s[5])[-2:]))
@pytest.mark.parametrize("sum,mul", OPS)
def test_cuda_matches_oracle(cuda_device, shape, sum, mul):
    seed, N, E, R, D, dtype, hub, unit = shape
    ei, et, w, rel, x, og = random_case(seed, N, E, R, D, dtype=dtype, hub=hub, unit_weight=unit)
    ours = run_ours(sum, mul, ei, et, w, rel, x, og)
    expected = oracle_all(sum, mul, ei, et, w, rel, x, og)
    check_against(sum, mul, ours, expected)


@pytest.mark.parametrize("sum,mul", OPS)
def test_function_apply_presorted_and_assert(cuda_device, sum, mul):
    """RSPMM*Function.apply is the reference's inner surface: pre-sorted edges, AssertionError otherwise."""
    import oracle
    from ultra_b200 import rspmm as mod
    ei, et, w, rel, x, og = random_case(41, 50, 600, 4, 32)
    Function = getattr(mod, "RSPMM%s%sFunction" % (sum.capitalize(), mul.capitalize()))
    d = cuda_device
    with pytest.raises(AssertionError, match="Expect sorted"):
        Function.apply(ei.to(d), et.to(d), w.to(d), rel.to(d), x.to(d))
    sei, stt, sw, _ = oracle.sort_edges(ei.numpy(), et.numpy(), w.numpy())
    out = Function.apply(torch.from_numpy(sei).to(d), torch.from_numpy(stt).to(d), torch.from_numpy(sw).to(d),
                         rel.to(d), x.to(d))
    exp = oracle_all(sum, mul, ei, et, w, rel, x)
    (assert_close_ref if sum == "add" else assert_equal_values)(out, exp, "forward")


def test_out_of_range_index_raises(cuda_device):
    from ultra_b200.rspmm import generalized_rspmm
    d = cuda_device
    ei = torch.tensor([[0, 5], [1, 1]], device=d)
    with pytest.raises(IndexError):
        generalized_rspmm(ei, torch.zeros(2, dtype=torch.long, device=d), torch.ones(2, device=d),
                          torch.ones(1, 4, device=d), torch.ones(3, 4, device=d))


def test_weight_grad_none_when_not_required(cuda_device):
    ei, et, w, rel, x, og = random_case(43, 40, 300, 4, 32)
    out, wg, rg, ig = run_ours("add", "mul", ei, et, w, rel, x, og, weight_grad=False)
    exp = oracle_all("add", "mul", ei, et, w, rel, x, og)
    assert wg is None
    assert_close_ref(rg, exp[2], "relation_grad")
    assert_close_ref(ig, exp[3], "input_grad")


def test_noncontiguous_inputs(cuda_device):
    """The reference makes private contiguous copies (rspmm.cu:227-231); strided views must work."""
    from ultra_b200.rspmm import generalized_rspmm
    ei, et, w, rel, x, og = random_case(44, 40, 300, 4, 32)
    d = cuda_device
    x2 = torch.randn(40, 64).index_copy_(1, torch.arange(0, 64, 2), x)  # x lives in the even columns
    rel_t = rel.t().contiguous().t()  # column-major storage
    out = generalized_rspmm(ei.to(d), et.to(d), w.to(d), rel_t.to(d), x2.to(d)[:, ::2], sum="add", mul="mul")
    assert_close_ref(out, oracle_all("add", "mul", ei, et, w, rel, x), "forward")


def test_plan_cache_reused_across_layers(cuda_device):
    from ultra_b200.plan import plan_cache
    from ultra_b200.rspmm import generalized_rspmm
    ei, et, w, rel, x, og = random_case(45, 40, 300, 4, 32)
    d = cuda_device
    ei, et, w = ei.to(d), et.to(d), w.to(d)
    plan_cache.clear()
    m0 = plan_cache.misses
    h = x.to(d)
    for _ in range(6):  # six layers over the same graph object -> one plan build
        h = generalized_rspmm(ei, et, w, rel.to(d), h)
    assert plan_cache.misses == m0 + 1
    ei[0, 0] = (ei[0, 0] + 1) % 38  # in-place edit bumps the version -> new plan
    generalized_rspmm(ei, et, w, rel.to(d), x.to(d))
    assert plan_cache.misses == m0 + 2


@pytest.mark.parametrize("D", [1100, 2852], ids=["3_slabs", "6_slabs_ragged"])
@pytest.mark.parametrize("sum,mul", [("add", "mul"), ("max", "add")])
def test_host_tensors_are_staged_through_gpu(cuda_device, sum, mul, D):
    """CPU tensors in -> CPU tensors out (reference dispatches on input.device.type, rspmm.py:19-22); the work
    runs on the GPU in pipelined column slabs (ultra_rspmm_forward_host).  1100 columns -> 3 slabs of 512; 2852
    columns -> 6 slabs, the last one ragged (both staging buffers are reused twice)."""
    ei, et, w, rel, x, og = random_case(46, 300, 5000, 9, D, hub=900)
    ours = run_ours(sum, mul, ei, et, w, rel.pin_memory(), x.pin_memory(), og, device="cpu")
    assert ours[0].device.type == "cpu"
    check_against(sum, mul, ours, oracle_all(sum, mul, ei, et, w, rel, x, og))


def test_c_abi_rejects_bad_arguments(cuda_device):
    import ctypes
    from ultra_b200 import _capi
    from ultra_b200.plan import GraphPlan
    ei, et, w, rel, x, og = random_case(47, 40, 300, 4, 32)
    plan = GraphPlan(ei.to(cuda_device), et.to(cuda_device), 40, 4)
    lib = _capi.lib()
    xd = x.to(cuda_device)
    rc = lib.ultra_rspmm_forward(plan.handle, 7, 0, 0, xd.data_ptr(), xd.data_ptr(), xd.data_ptr(), 32, None, 0, None)
    assert rc != 0 and b"sum_op" in lib.ultra_rspmm_last_error()
    rc = lib.ultra_rspmm_forward(None, 0, 0, 0, xd.data_ptr(), xd.data_ptr(), xd.data_ptr(), 32, None, 0, None)
    assert rc != 0 and b"null plan" in lib.ultra_rspmm_last_error()


def test_forward_backward_are_cuda_graph_capturable(cuda_device):
    """The reference syncs the host on every call (sortedness assert, rspmm.py:15-17); with a cached plan the whole
    forward+backward is stream-ordered, so it can be captured once and replayed (relation-graph regime)."""
    from ultra_b200.rspmm import generalized_rspmm
    d = cuda_device
    ei, et, w, rel, x, og = random_case(51, 120, 3000, 4, 64, hub=500, unit_weight=True)
    ei, et, w = ei.to(d), et.to(d), w.to(d)
    s_rel, s_x, s_og = rel.to(d).requires_grad_(), x.to(d).requires_grad_(), og.to(d)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # warm-up outside capture: plan build, weight bind, scratch allocation
        for _ in range(2):
            out = generalized_rspmm(ei, et, w, s_rel, s_x, sum="add", mul="mul")
            out.backward(s_og)
            s_rel.grad = s_x.grad = None
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_out = generalized_rspmm(ei, et, w, s_rel, s_x, sum="add", mul="mul")
        g_out.backward(s_og)
    g_xgrad, g_rgrad = s_x.grad, s_rel.grad
    for seed in (1, 2):
        gen = torch.Generator().manual_seed(seed)
        nx, nrel = torch.randn(x.shape, generator=gen), torch.randn(rel.shape, generator=gen)
        with torch.no_grad():
            s_x.copy_(nx)
            s_rel.copy_(nrel)
        graph.replay()
        e_out, _, e_rg, e_ig = oracle_all("add", "mul", ei.cpu(), et.cpu(), w.cpu(), nrel, nx, og)
        assert_close_ref(g_out.detach(), e_out, "replayed forward")
        assert_close_ref(g_xgrad, e_ig, "replayed input_grad")
        assert_close_ref(g_rgrad, e_rg, "replayed relation_grad")


@pytest.mark.parametrize("mul", ["mul", "add"])
@pytest.mark.parametrize("dtype,D", [(torch.float32, 96), (torch.float64, 40)])
def test_pna_one_pass_equals_four_calls(cuda_device, mul, dtype, D):
    """generalized_rspmm_pna == the reference layer's four rspmm calls (layers.py:211-215), forward and backward."""
    from ultra_b200.rspmm import generalized_rspmm, generalized_rspmm_pna
    d = cuda_device
    ei, et, w, rel, x, og = random_case(61, 150, 5000, 7, D, dtype=dtype, hub=900)
    ei, et, w = ei.to(d), et.to(d), w.to(d)
    g = torch.Generator().manual_seed(62)
    ogs = [torch.randn(x.shape, generator=g, dtype=dtype).to(d) for _ in range(4)]
    r1, x1 = rel.to(d).requires_grad_(), x.to(d).requires_grad_()
    outs = generalized_rspmm_pna(ei, et, w, r1, x1, mul=mul)
    sum((o * g_).sum() for o, g_ in zip(outs, ogs)).backward()
    r2, x2 = rel.to(d).requires_grad_(), x.to(d).requires_grad_()
    refs = [generalized_rspmm(ei, et, w, r2, x2, sum="add", mul=mul),
            generalized_rspmm(ei, et, w, r2 ** 2, x2 ** 2, sum="add", mul=mul),
            generalized_rspmm(ei, et, w, r2, x2, sum="max", mul=mul),
            generalized_rspmm(ei, et, w, r2, x2, sum="min", mul=mul)]
    sum((o * g_).sum() for o, g_ in zip(refs, ogs)).backward()
    for name, a, b in zip(("sum", "sq_sum", "max", "min"), outs, refs):
        if name in ("max", "min"):
            assert_equal_values(a.detach(), b.detach().cpu(), name)
        else:
            assert_close_ref(a.detach(), b.detach().cpu(), name)
    assert_close_ref(x1.grad, x2.grad.cpu(), "input_grad")
    assert_close_ref(r1.grad, r2.grad.cpu(), "relation_grad")
    # and against the oracle directly
    for name, s_, a in (("sum", "add", outs[0]), ("max", "max", outs[2]), ("min", "min", outs[3])):
        exp = oracle_all(s_, mul, ei.cpu(), et.cpu(), w.cpu(), rel, x)
        (assert_close_ref if s_ == "add" else assert_equal_values)(a.detach(), exp, name + " vs oracle")


@pytest.mark.parametrize("mul", ["mul", "add"])
def test_zero_one_weights_remove_edges_for_sum(cuda_device, mul):
    """The incremental alternative to re-planning a masked graph (GraphPlan.edge_mask_weight): for the sum aggregate a
    0/1 edge weight on the BASE graph's plan equals rspmm on the physically filtered edge list (forward and grads)."""
    from ultra_b200.rspmm import generalized_rspmm
    from ultra_b200.plan import plan_cache
    ei, et, w, rel, x, og = random_case(61, 500, 20000, 7, 128, hub=3000, unit_weight=True)
    g = torch.Generator().manual_seed(62)
    keep = torch.rand(ei.shape[1], generator=g) > 0.1
    d = cuda_device
    eid, etd = ei.to(d), et.to(d)
    plan = plan_cache.get(eid, etd, 500, 7)
    relg, xg = rel.to(d).requires_grad_(), x.to(d).requires_grad_()
    out = generalized_rspmm(eid, etd, plan.edge_mask_weight(keep.to(d)), relg, xg, sum="add", mul=mul)
    out.backward(og.to(d))
    e_out, _, e_rg, e_ig = oracle_all("add", mul, ei[:, keep], et[keep], torch.ones(int(keep.sum())), rel, x, og)
    assert_close_ref(out.detach(), e_out, "masked forward")
    assert_close_ref(relg.grad, e_rg, "masked relation_grad")
    assert_close_ref(xg.grad, e_ig, "masked input_grad")
