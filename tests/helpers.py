"""Shared helpers for the parity tests (test infrastructure; may use oracle/)."""
import os

import numpy as np
import torch

import oracle

OPS = [(s, m) for s in ("add", "min", "max") for m in ("mul", "add")]
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rspmm_golden.npz")
GOLDEN_CASES = ["rand_f32", "rand_f64", "hub_f32", "ties_f32", "oddD_f32", "unitw_f32"]


def load_golden():
    return np.load(GOLDEN)


def random_case(seed, N, E, R, D, dtype=torch.float32, hub=0, unit_weight=False, isolated=2):
    g = torch.Generator().manual_seed(seed)
    hi = max(N - isolated, 1)
    ei = torch.randint(0, hi, (2, E), generator=g)
    if hub:
        ei[0, :hub] = min(1, hi - 1)
    et = torch.randint(0, R, (E,), generator=g)
    rel = torch.randn(R, D, generator=g, dtype=dtype)
    x = torch.randn(N, D, generator=g, dtype=dtype)
    w = torch.ones(E, dtype=dtype) if unit_weight else torch.rand(E, generator=g, dtype=dtype) + 0.5
    og = torch.randn(N, D, generator=g, dtype=dtype)
    return ei, et, w, rel, x, og


def oracle_all(sum, mul, ei, et, w, rel, x, og=None):
    """C oracle on unsorted CPU tensors; grads returned in the caller's edge order."""
    sei, stt, sw, order = oracle.sort_edges(ei.numpy(), et.numpy(), w.numpy())
    out = oracle.c_forward(sum, mul, sei, stt, sw, rel.numpy(), x.numpy())
    if og is None:
        return torch.from_numpy(out)
    wg, rg, ig = oracle.c_backward(sum, mul, sei, stt, sw, rel.numpy(), x.numpy(), out, og.numpy())
    wg_caller = np.empty_like(wg)
    wg_caller[order] = wg
    return tuple(torch.from_numpy(a) for a in (out, wg_caller, rg, ig))


def assert_close_ref(actual, expected, what, rtol=1e-5):
    """north-star tolerance: 1e-5 fp32 relative, with atol = 1e-5 * max|expected| (SURVEY.md section 8d:
    tree-order vs sequential fp32 sums differ by ~1e-7 * sum|terms|, so a pure relative test is ill-posed under
    cancellation).  fp64 uses 1e-12."""
    expected = torch.as_tensor(expected)
    actual = torch.as_tensor(actual).cpu()
    if expected.dtype == torch.float64:
        rtol = 1e-12
    scale = float(expected.abs().max()) if expected.numel() else 0.0
    if not np.isfinite(scale) or scale > 1e30:  # +-FLT_MAX identities of empty rows: compare finite part by value
        fin = expected.abs() < 1e30
        assert torch.equal(actual[~fin], expected[~fin]), what + ": identity (empty-row) entries differ"
        if fin.any():
            scale = float(expected[fin].abs().max())
            torch.testing.assert_close(actual[fin], expected[fin], rtol=rtol, atol=rtol * scale, msg=lambda m: what + ": " + m)
        return
    torch.testing.assert_close(actual, expected, rtol=rtol, atol=rtol * scale, msg=lambda m: what + ": " + m)


def assert_equal_values(actual, expected, what):
    """bit-exact by value (min/max forward involves no re-association; -0.0 == +0.0 compare equal)."""
    actual = torch.as_tensor(actual).cpu()
    expected = torch.as_tensor(expected)
    assert actual.shape == expected.shape, what
    assert bool((actual == expected).all()), "%s: %d mismatching elements" % (what, int((actual != expected).sum()))


def assert_close_with_truth(actual, expected32, truth64, what, rtol=1e-5):
    """fp32 comparison that stays well-posed for very long rows (SURVEY.md section 8d "sum order vs the oracle"):
    the reference folds a row sequentially in fp32, our kernels fold it chunk-wise; for a row of tens of thousands
    of edges the reference's OWN rounding error exceeds 1e-5 * max|out|.  Criterion: we may differ from the fp32
    reference by the north-star tolerance plus (a small multiple of) the reference's own distance to the fp64 truth,
    and we must be within that same budget of the truth."""
    actual = torch.as_tensor(actual).cpu().double()
    expected32 = torch.as_tensor(expected32).double()
    truth64 = torch.as_tensor(truth64).double()
    fin = truth64.abs() < 1e30
    assert torch.equal(actual[~fin], expected32[~fin]), what + ": identity (empty-row) entries differ"
    if not fin.any():
        return
    scale = float(truth64[fin].abs().max())
    ref_err = (expected32 - truth64).abs()[fin]
    budget = rtol * scale + rtol * truth64.abs()[fin] + 4 * ref_err
    for name, other in (("reference", expected32), ("fp64 truth", truth64)):
        err = (actual - other).abs()[fin]
        bad = err > budget
        assert not bool(bad.any()), "%s vs %s: %d elements beyond budget, worst %.3e (budget there %.3e, ref err %.3e)" % (
            what, name, int(bad.sum()), float(err[bad].max()), float(budget[bad][err[bad].argmax()]),
            float(ref_err[bad][err[bad].argmax()]))


# ---- layer-level fixture (tests/golden/layer_golden.npz, made by tests/golden/make_layer_golden.py) ------------------
LAYER_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layer_golden.npz")


def load_layer_golden():
    return np.load(LAYER_GOLDEN)


def ultra_modules_from_state_dict(sd, device="cpu"):
    """The product-side layer classes (ultra_b200.layer) populated from a reference checkpoint's state-dict:
    (relation_layers, entity_layers), configuration of the shipped ultra_* checkpoints (distmult, sum, layer_norm,
    relu; /root/reference/config/transductive/inference.yaml).  Keys are loaded with strict=True, i.e. the module
    parameter names are the reference's."""
    from ultra_b200.layer import GeneralizedRelationalConv

    def stack(prefix, **kw):
        n = 1 + max(int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix))
        layers = []
        for i in range(n):
            layer = GeneralizedRelationalConv(64, 64, kw["num_relation"], 64, "distmult", "sum", layer_norm=True,
                                              activation="relu", dependent=False,
                                              project_relations=kw["project_relations"])
            own = {k[len(prefix) + len(str(i)) + 1:]: v for k, v in sd.items() if k.startswith("%s%d." % (prefix, i))}
            layer.load_state_dict(own, strict=True)
            layers.append(layer.to(device).eval())
        return layers

    relation_layers = stack("relation_model.layers.", num_relation=4, project_relations=False)
    entity_layers = stack("entity_model.layers.", num_relation=1, project_relations=True)
    return relation_layers, entity_layers


def ultra_forward(relation_layers, entity_layers, sd, fx, device="cpu"):
    """Ultra.forward (/root/reference/ultra/models.py:18-27) through ultra_b200.layer: relation-graph stack ->
    relation representations -> entity-graph stack -> tail scores.  Returns (rel_repr, h6, score)."""
    from ultra_b200.layer import bellmanford_entity

    t = lambda k: torch.from_numpy(np.asarray(fx[k])).to(device)  # noqa: E731
    num_node, num_rel = int(fx["num_node"]), int(fx["num_relation"])
    h_index, r_index, t_index = t("h_index"), t("r_index"), t("t_index")
    # RelNBFNet.bellmanford (models.py:55-95): all-ones query at the query relation, output = last hidden
    ones = torch.ones(len(r_index), 64, device=device)
    rel_repr = bellmanford_entity(relation_layers, ones, r_index, t("rel_edge_index"), t("rel_edge_type"), num_rel)
    # EntityNBFNet.forward / bellmanford (models.py:131-186)
    for layer in entity_layers:
        layer.relation = rel_repr
    query = rel_repr[torch.arange(len(r_index), device=device), r_index]
    h6 = bellmanford_entity(entity_layers, query, h_index, t("edge_index"), t("edge_type"), num_node)
    feature = torch.cat([h6, query.unsqueeze(1).expand(-1, num_node, -1)], dim=-1)
    f = feature.gather(1, t_index.unsqueeze(-1).expand(-1, -1, feature.shape[-1]))
    w = lambda k: sd[k].to(device)  # noqa: E731
    f = torch.relu(torch.nn.functional.linear(f, w("entity_model.mlp.0.weight"), w("entity_model.mlp.0.bias")))
    score = torch.nn.functional.linear(f, w("entity_model.mlp.2.weight"), w("entity_model.mlp.2.bias")).squeeze(-1)
    return rel_repr, h6, score
