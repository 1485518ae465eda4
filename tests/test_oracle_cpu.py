"""CPU suite: the C oracle against (1) golden vectors produced by the unmodified reference CPU kernels,
(2) the hand-computed case of SURVEY.md section 8c, (3) a dense PyTorch formula, (4) oracle/_ref when present."""
import numpy as np
import pytest
import torch

import oracle
from oracle import ref
from helpers import OPS, GOLDEN_CASES, load_golden, random_case, oracle_all, assert_close_ref, assert_equal_values

FLT_MAX = float(np.finfo(np.float32).max)


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("sum,mul", OPS)
def test_oracle_matches_reference_golden(case, sum, mul):
    g = load_golden()
    t = lambda k: torch.from_numpy(g[case + "/" + k])
    out, wg, rg, ig = oracle_all(sum, mul, t("edge_index"), t("edge_type"), t("edge_weight"), t("relation"),
                                 t("input"), t("output_grad"))
    k = "%s_%s/" % (sum, mul)
    assert_equal_values(out, t(k + "output"), "forward")  # same sequential order as rspmm.cpp:62-71 -> bit-exact
    assert_close_ref(wg, t(k + "weight_grad"), "weight_grad")
    assert_close_ref(rg, t(k + "relation_grad"), "relation_grad")
    assert_close_ref(ig, t(k + "input_grad"), "input_grad")


def test_hand_case_survey_8c():
    ei = torch.tensor([[0, 0, 1], [1, 1, 2]])
    et = torch.tensor([0, 0, 1])
    w = torch.ones(3)
    rel = torch.tensor([[2., 3.], [1., 1.]])
    x = torch.tensor([[1., 1.], [5., 7.], [4., 4.]])
    og = torch.ones(3, 2)
    exp = {"add": [[20, 42], [4, 4], [0, 0]], "min": [[10, 21], [4, 4], [FLT_MAX, FLT_MAX]],
           "max": [[10, 21], [4, 4], [-FLT_MAX, -FLT_MAX]]}
    for s in ("add", "min", "max"):
        out, wg, rg, ig = oracle_all(s, "mul", ei, et, w, rel, x, og)
        assert_equal_values(out, torch.tensor(exp[s], dtype=torch.float32), s)
        assert_equal_values(rg, torch.tensor([[10., 14.], [4., 4.]]), s + " relation_grad")
        assert_equal_values(ig, torch.tensor([[0., 0.], [4., 6.], [1., 1.]]), s + " input_grad")
        assert_equal_values(wg, torch.tensor([31., 31., 8.]), s + " weight_grad")  # ties: both get gradient


@pytest.mark.parametrize("mul", ["mul", "add"])
def test_oracle_vs_dense_formula(mul):
    ei, et, w, rel, x, og = random_case(11, N=40, E=500, R=6, D=24, dtype=torch.float64)
    msg = (rel[et] * x[ei[1]] if mul == "mul" else rel[et] + x[ei[1]]) * w[:, None]
    dense = torch.zeros(40, 24, dtype=torch.float64).index_add_(0, ei[0], msg)
    assert_close_ref(oracle_all("add", mul, ei, et, w, rel, x), dense, "add")
    for s, red in (("min", "amin"), ("max", "amax")):
        ident = torch.finfo(torch.float64).max * (1 if s == "min" else -1)
        d = torch.full((40, 24), ident, dtype=torch.float64).scatter_reduce(
            0, ei[0][:, None].expand(-1, 24), msg, reduce=red, include_self=True)
        assert_equal_values(oracle_all(s, mul, ei, et, w, rel, x), d, s)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("sum,mul", OPS)
def test_oracle_vs_live_reference(sum, mul):
    ei, et, w, rel, x, og = random_case(21, N=120, E=3000, R=9, D=40, hub=700)
    sei, stt, sw, order = oracle.sort_edges(ei.numpy(), et.numpy(), w.numpy())
    tei, ttt, tw = torch.from_numpy(sei), torch.from_numpy(stt), torch.from_numpy(sw)
    r_out = ref.ref_forward(sum, mul, tei, ttt, tw, rel, x)
    out, wg, rg, ig = oracle_all(sum, mul, ei, et, w, rel, x, og)
    assert_equal_values(out, r_out, "forward")
    r_wg, r_rg, r_ig = ref.ref_backward(sum, mul, tei, ttt, tw, rel, x, r_out, og)
    assert_close_ref(rg, r_rg, "relation_grad")
    assert_close_ref(ig, r_ig, "input_grad")
    assert_close_ref(wg[torch.from_numpy(order)], r_wg, "weight_grad")


def test_oracle_rejects_unsorted():
    ei = np.array([[1, 0], [0, 0]])
    with pytest.raises(AssertionError, match="Expect sorted"):
        oracle.c_forward("add", "mul", ei, np.zeros(2, np.int64), np.ones(2, np.float32),
                         np.ones((1, 4), np.float32), np.ones((2, 4), np.float32))
