"""GPU parity at the remaining BASELINE.json scales (-m gpu): C3 (relation graph + entity graph at batch 256), C4 (CoDEx-L
scale) and C5 (synthetic sweep corner points up to 1M nodes).  Round-1 review: these scales were timed but never checked.

What each scale exercises that the small-shape tests do not:
  C3 relation graph  474 nodes, 4 relation rows, ~0.9 M edges: the small-graph kernel variant (gathered tile staged in
                     shared memory by TMA, SM=2) at full size, ~1 900-edge rows (every row is split into chunk groups)
  C3 entity, B=256   16 384 columns, 953 MB matrices: 32-bit row-byte offsets near their limit, 512 column tiles
  C4                 77 951 nodes: the per-tile gather footprint (10 MB) with few heads per GPU (8: D = 512)
  C5                 1M nodes x 20 edges: int32/unsigned address arithmetic (row * row_bytes > 4 GiB), gather footprint
                     beyond L2, hub rows of ~10^5 edges (Zipf) -> thousands of partial slots

Element-wise comparison is against the reference's own CPU kernel (oracle/_ref, OpenMP) when shipped, else the C oracle,
on crops / sampled heads small enough for the CPU to finish in seconds; the full widths are tied to those crops through
column independence (bit-exact).  Tolerances: helpers.assert_close_ref (1e-5 relative to max|ref|), min/max forward
bit-exact; rows longer than a few thousand edges use assert_close_with_truth (the fp32 reference's own rounding error
against fp64 exceeds 1e-5 there)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import ref
from helpers import OPS, assert_close_ref, assert_close_with_truth, assert_equal_values

pytestmark = pytest.mark.gpu


def _sorted_cpu(ei, et, w):
    sei, stt, sw, order = oracle.sort_edges(ei.numpy(), et.numpy(), w.numpy())
    return sei, stt, sw, order


def _cpu_forward(sum, mul, sorted_edges, rel, x):
    """Reference CPU kernel (OpenMP) when shipped, else the scalar C restatement; inputs pre-sorted numpy + CPU tensors."""
    sei, stt, sw, _ = sorted_edges
    if ref.available() and x.dtype == torch.float32:
        return ref.ref_forward(sum, mul, torch.from_numpy(sei), torch.from_numpy(stt), torch.from_numpy(sw), rel, x)
    return torch.from_numpy(oracle.c_forward(sum, mul, sei, stt, sw, rel.numpy(), x.numpy()))


def _cpu_backward(sum, mul, sorted_edges, rel, x, out, og):
    """C oracle backward (the reference's CPU backward takes a mutex per element: minutes at these sizes)."""
    sei, stt, sw, _ = sorted_edges
    _, rg, ig = oracle.c_backward(sum, mul, sei, stt, sw, rel.numpy(), x.numpy(), out.numpy(), og.numpy())
    return torch.from_numpy(rg), torch.from_numpy(ig)


def _ours(ei, et, w, rel, x, sum, mul):
    from ultra_b200.rspmm import generalized_rspmm
    return generalized_rspmm(ei, et, w, rel, x, sum=sum, mul=mul)


def _cmp_fwd(sum, got, exp, what, truth=None):
    if sum != "add":
        assert_equal_values(got, exp, what)
    elif truth is not None:
        assert_close_with_truth(got, exp, truth, what)
    else:
        assert_close_ref(got, exp, what)


# ---------------------------------------------------------------------------------------------------------------------
# C3: relation graph of the FB15k-237-scale KG
@pytest.fixture(scope="module")
def relation_graph():
    from ultra_b200.synth import synth_scale, relation_graph_of
    ei, et, R, N = synth_scale("fb15k237", alpha=1.0, seed=1234 + 3)
    rei, ret, rR, rN = relation_graph_of(ei, et, N, R)
    assert rN == 474 and rR == 4 and rei.shape[1] > 300_000  # dense: a large share of the 4 * 474^2 possible edges
    w = torch.ones(rei.shape[1])
    return rei, ret, w, rR, rN, _sorted_cpu(rei, ret, w)


@pytest.mark.parametrize("sum,mul", OPS)
def test_c3_relation_graph_fwd_and_grads(cuda_device, relation_graph, sum, mul):
    """All six op pairs, forward + relation_grad + input_grad, B = 16 heads (D = 1024): the SM=2 kernel at full size."""
    rei, ret, w, R, N, se = relation_graph
    d = cuda_device
    B, dim = 16, 64
    D = B * dim
    g = torch.Generator().manual_seed(31)
    # the relation stack broadcasts ONE (4, 64) embedding over the heads (layers.py:76); keep that structure
    rel = torch.randn(R, dim, generator=g).repeat(1, B).contiguous()
    x = torch.randn(N, D, generator=g)
    og = torch.randn(N, D, generator=g)
    relg, xg = rel.to(d).requires_grad_(), x.to(d).requires_grad_()
    out = _ours(rei.to(d), ret.to(d), w.to(d), relg, xg, sum, mul)
    out.backward(og.to(d))
    exp = _cpu_forward(sum, mul, se, rel, x)
    truth = None
    if sum == "add":  # ~1 900 edges per row: compare through the fp64 truth
        truth = torch.from_numpy(oracle.c_forward(sum, mul, se[0], se[1], se[2].astype(np.float64),
                                                  rel.double().numpy(), x.double().numpy()))
    _cmp_fwd(sum, out.detach(), exp, "C3 relation graph forward %s-%s" % (sum, mul), truth)
    # grads: the backward of min/max must see the SAME output the forward produced (ties by equality): feed ours
    out_cpu = out.detach().cpu()
    e_rg, e_ig = _cpu_backward(sum, mul, se, rel, x, out_cpu, og)
    t_rg, t_ig = (torch.from_numpy(a) for a in oracle.c_backward(
        sum, mul, se[0], se[1], se[2].astype(np.float64), rel.double().numpy(), x.double().numpy(),
        out_cpu.double().numpy(), og.double().numpy())[1:])
    # relation_grad rows fold ~E/4 = 10^5 terms, input_grad rows ~1 900
    assert_close_with_truth(relg.grad, e_rg, t_rg, "C3 relation graph relation_grad %s-%s" % (sum, mul))
    assert_close_with_truth(xg.grad, e_ig, t_ig, "C3 relation graph input_grad %s-%s" % (sum, mul))


def test_c3_relation_graph_wide_batch_column_independence(cuda_device, relation_graph):
    """B = 256 heads on the relation graph: heads computed alone are bit-identical to the same heads in the batch."""
    rei, ret, w, R, N, se = relation_graph
    d = cuda_device
    B, dim = 256, 64
    g = torch.Generator(device=d).manual_seed(32)
    rel = torch.randn(R, B * dim, generator=g, device=d)
    x = torch.randn(N, B * dim, generator=g, device=d)
    eid, etd, wd = rei.to(d), ret.to(d), w.to(d)
    full = _ours(eid, etd, wd, rel, x, "add", "mul")
    for b in (0, 131, 255):
        sl = slice(b * dim, (b + 1) * dim)
        alone = _ours(eid, etd, wd, rel[:, sl].contiguous(), x[:, sl].contiguous(), "add", "mul")
        assert_equal_values(alone, full[:, sl].cpu(), "relation graph head %d alone vs in batch of 256" % b)


# ---------------------------------------------------------------------------------------------------------------------
# C3: entity graph at batch 256 (D = 16 384)
@pytest.mark.parametrize("sum,mul", [("add", "mul"), ("max", "mul")])
def test_c3_entity_batch256_sampled_heads(cuda_device, sum, mul):
    from ultra_b200.synth import synth_scale
    ei, et, R, N = synth_scale("fb15k237", alpha=1.0, seed=1234 + 3)
    d = cuda_device
    B, dim = 256, 64
    D = B * dim
    g = torch.Generator(device=d).manual_seed(33)
    rel = torch.randn(R, D, generator=g, device=d)
    x = torch.randn(N, D, generator=g, device=d)
    og = torch.randn(N, D, generator=g, device=d)
    w = torch.ones(ei.shape[1])
    se = _sorted_cpu(ei, et, w)
    relg, xg = rel.clone().requires_grad_(), x.clone().requires_grad_()
    out = _ours(ei.to(d), et.to(d), w.to(d), relg, xg, sum, mul)
    out.backward(og)
    for b in (0, 97, 255):
        sl = slice(b * dim, (b + 1) * dim)
        r_c, x_c, og_c = rel[:, sl].cpu().contiguous(), x[:, sl].cpu().contiguous(), og[:, sl].cpu().contiguous()
        exp = _cpu_forward(sum, mul, se, r_c, x_c)
        _cmp_fwd(sum, out.detach()[:, sl], exp, "C3 entity B=256 head %d forward" % b)
        e_rg, e_ig = _cpu_backward(sum, mul, se, r_c, x_c, out.detach()[:, sl].cpu().contiguous(), og_c)
        assert_close_ref(xg.grad[:, sl], e_ig, "C3 entity B=256 head %d input_grad" % b)
        assert_close_ref(relg.grad[:, sl], e_rg, "C3 entity B=256 head %d relation_grad" % b, rtol=2e-5)


# ---------------------------------------------------------------------------------------------------------------------
# C4: CoDEx-L scale
@pytest.mark.parametrize("alpha", [1.0, 0.0])
def test_c4_codex_l_crop_with_grads_and_full_width(cuda_device, alpha):
    from ultra_b200.synth import synth_scale
    ei, et, R, N = synth_scale("codex_l", alpha=alpha, seed=1234 + 4)
    assert N == 77951 and R == 138
    d = cuda_device
    dim = 64
    w = torch.ones(ei.shape[1])
    se = _sorted_cpu(ei, et, w)
    eid, etd, wd = ei.to(d), et.to(d), w.to(d)
    # (1) what one GPU of the 8-GPU run owns: 8 heads, D = 512, forward + grads, all of it against the CPU
    g = torch.Generator().manual_seed(41)
    D = 8 * dim
    rel, x, og = torch.randn(R, D, generator=g), torch.randn(N, D, generator=g), torch.randn(N, D, generator=g)
    for sum, mul in (("add", "mul"), ("min", "add")):
        relg, xg = rel.to(d).requires_grad_(), x.to(d).requires_grad_()
        out = _ours(eid, etd, wd, relg, xg, sum, mul)
        out.backward(og.to(d))
        exp = _cpu_forward(sum, mul, se, rel, x)
        truth = None
        if sum == "add" and alpha > 0:  # Zipf hub rows fold tens of thousands of edges
            truth = torch.from_numpy(oracle.c_forward(sum, mul, se[0], se[1], se[2].astype(np.float64),
                                                      rel.double().numpy(), x.double().numpy()))
        _cmp_fwd(sum, out.detach(), exp, "C4 (alpha %.0f) 8-head forward %s-%s" % (alpha, sum, mul), truth)
        out_cpu = out.detach().cpu()
        e_rg, e_ig = _cpu_backward(sum, mul, se, rel, x, out_cpu, og)
        t_rg, t_ig = (torch.from_numpy(a) for a in oracle.c_backward(
            sum, mul, se[0], se[1], se[2].astype(np.float64), rel.double().numpy(), x.double().numpy(),
            out_cpu.double().numpy(), og.double().numpy())[1:])
        assert_close_with_truth(relg.grad, e_rg, t_rg, "C4 relation_grad %s-%s" % (sum, mul))
        assert_close_with_truth(xg.grad, e_ig, t_ig, "C4 input_grad %s-%s" % (sum, mul))
    # (2) the whole batch of 64 heads on one GPU: sampled heads against the CPU, plus column independence
    B = 64
    gd = torch.Generator(device=d).manual_seed(42)
    relf = torch.randn(R, B * dim, generator=gd, device=d)
    xf = torch.randn(N, B * dim, generator=gd, device=d)
    full = _ours(eid, etd, wd, relf, xf, "add", "mul")
    for b in (5, 63):
        sl = slice(b * dim, (b + 1) * dim)
        alone = _ours(eid, etd, wd, relf[:, sl].contiguous(), xf[:, sl].contiguous(), "add", "mul")
        assert_equal_values(alone, full[:, sl].cpu(), "C4 head %d alone vs in batch" % b)
    sl = slice(17 * dim, 18 * dim)
    r_c, x_c = relf[:, sl].cpu().contiguous(), xf[:, sl].cpu().contiguous()
    exp = _cpu_forward("add", "mul", se, r_c, x_c)
    truth = torch.from_numpy(oracle.c_forward("add", "mul", se[0], se[1], se[2].astype(np.float64),
                                              r_c.double().numpy(), x_c.double().numpy()))
    assert_close_with_truth(full[:, sl], exp, truth, "C4 head 17 of 64 vs CPU")


# ---------------------------------------------------------------------------------------------------------------------
# C5: synthetic sweep corner points at 1M nodes
@pytest.mark.parametrize("alpha", [0.0, 1.0])
@pytest.mark.parametrize("deg", [5, 20])
def test_c5_million_nodes_sampled_columns(cuda_device, deg, alpha):
    from ultra_b200.synth import synth_kg_torch
    d = cuda_device
    N, R_base, dim, B = 1_000_000, 128, 64, 32
    D = B * dim
    ei, et, R = synth_kg_torch(N, R_base, N * deg // 2, alpha=alpha, seed=1234 + 5, device=d)
    E = ei.shape[1]
    assert R == 256 and E == N * deg
    g = torch.Generator(device=d).manual_seed(51)
    rel = torch.randn(R, D, generator=g, device=d)
    x = torch.randn(N, D, generator=g, device=d)       # 8.2 GB: row offsets exceed 32 bits
    w = torch.ones(E, device=d)
    outs = {}
    for sum, mul in (("add", "mul"), ("max", "add")):
        outs[sum] = _ours(ei, et, w, rel, x, sum, mul)
    # CPU check on the first and the last head (the columns whose byte offsets are smallest / largest)
    se = _sorted_cpu(ei.cpu(), et.cpu(), w.cpu())
    for b in (0, B - 1):
        sl = slice(b * dim, (b + 1) * dim)
        r_c, x_c = rel[:, sl].cpu().contiguous(), x[:, sl].cpu().contiguous()
        for sum, mul in (("add", "mul"), ("max", "add")):
            exp = _cpu_forward(sum, mul, se, r_c, x_c)
            truth = None
            if sum == "add":
                truth = torch.from_numpy(oracle.c_forward(sum, mul, se[0], se[1], se[2].astype(np.float64),
                                                          r_c.double().numpy(), x_c.double().numpy()))
            _cmp_fwd(sum, outs[sum][:, sl], exp, "C5 N=1M deg=%d alpha=%.0f head %d %s" % (deg, alpha, b, sum), truth)
    # backward at this scale: input_grad / relation_grad of one more head against the C oracle
    sl = slice(7 * dim, 8 * dim)
    relg, xg = rel[:, sl].clone().requires_grad_(), x[:, sl].clone().requires_grad_()
    og = torch.randn(N, dim, generator=g, device=d)
    out = _ours(ei, et, w, relg, xg, "add", "mul")
    out.backward(og)
    r_c, x_c, og_c = relg.detach().cpu(), xg.detach().cpu(), og.cpu()
    out_c = out.detach().cpu()
    e_rg, e_ig = _cpu_backward("add", "mul", se, r_c, x_c, out_c, og_c)
    t_rg, t_ig = (torch.from_numpy(a) for a in oracle.c_backward(
        "add", "mul", se[0], se[1], se[2].astype(np.float64), r_c.double().numpy(), x_c.double().numpy(),
        out_c.double().numpy(), og_c.double().numpy())[1:])
    assert_close_with_truth(xg.grad, e_ig, t_ig, "C5 input_grad")
    assert_close_with_truth(relg.grad, e_rg, t_rg, "C5 relation_grad")
