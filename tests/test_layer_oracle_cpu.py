"""Layer-level oracle (last row) on CPU: the fixture tests/golden/layer_golden.npz holds the full
ULTRA forward (relation-graph stack -> entity-graph stack -> scores) with the REAL ultra_3g checkpoint weights and every
rspmm computed by the unmodified reference CPU extension.  Here, without /root/reference:

* the functional restatement (oracle/layer_ref.py) on top of the plain-C rspmm oracle reproduces it;
* the product-side layer classes (ultra_b200/layer.py) load the checkpoint's state-dict keys unchanged and, with their
  rspmm call replaced by the oracle (there is no CPU rspmm in the product), reproduce it too - so the GPU test of the
  same fixture (tests/test_layer_gpu.py) exercises only the CUDA kernels, not new glue.
"""
import numpy as np
import pytest
import torch

import oracle
from oracle import layer_ref
from helpers import load_layer_golden, ultra_forward, ultra_modules_from_state_dict

TOL = 2e-5  # relative to max|expected|: six layers of fp32 Linear/LayerNorm on top of 1e-5-level rspmm differences


def close(actual, expected, what, tol=TOL):
    expected = torch.as_tensor(np.asarray(expected))
    scale = float(expected.abs().max())
    torch.testing.assert_close(torch.as_tensor(actual).cpu(), expected, rtol=tol, atol=tol * scale,
                               msg=lambda m: what + ": " + m)


@pytest.fixture(scope="module")
def fx():
    return load_layer_golden()


def test_fixture_is_the_shipped_checkpoint(fx):
    sd = layer_ref.state_dict_from_npz(fx)
    assert sum(v.numel() for v in sd.values()) == 168705  # parameter count of ultra_3g.pth
    assert sd["relation_model.layers.0.relation.weight"].shape == (4, 64)
    assert sd["entity_model.layers.5.relation_projection.2.weight"].shape == (64, 64)
    assert int(fx["edge_index"].shape[1]) == 2 * 260 and int(fx["num_relation"]) == 20


def test_c_oracle_layer_stack_reproduces_reference_fixture(fx):
    sd = layer_ref.state_dict_from_npz(fx)
    t = lambda k: torch.from_numpy(np.asarray(fx[k]))  # noqa: E731
    with torch.no_grad():
        rel = layer_ref.relation_stack(sd, t("rel_edge_index"), t("rel_edge_type"), int(fx["num_relation"]),
                                       t("r_index"), rspmm="c")
        close(rel, fx["relation_representations"], "relation representations")
        # entity stack from the FIXTURE's relation representations (isolates the two stacks)
        feature, hiddens = layer_ref.entity_stack(sd, t("edge_index"), t("edge_type"), int(fx["num_node"]),
                                                  t("relation_representations"), t("h_index"), t("r_index"),
                                                  rspmm="c", all_hiddens=True)
        close(hiddens[0], fx["entity_hidden_1"], "entity layer 1")
        close(hiddens[-1], fx["entity_hidden_6"], "entity layer 6")
        close(layer_ref.score(sd, feature, t("t_index")), fx["score"], "scores")


def test_product_layer_classes_load_checkpoint_and_match_fixture(fx, monkeypatch):
    import ultra_b200.layer as product_layer

    def oracle_rspmm(edge_index, edge_type, edge_weight, relation, input, sum="add", mul="mul"):
        sei, stt, sw, _ = oracle.sort_edges(edge_index.numpy(), edge_type.numpy(), edge_weight.numpy())
        return torch.from_numpy(oracle.c_forward(sum, mul, sei, stt, sw, relation.contiguous().numpy(),
                                                 input.contiguous().numpy()))

    monkeypatch.setattr(product_layer, "generalized_rspmm", oracle_rspmm)
    monkeypatch.setattr(product_layer.GeneralizedRelationalConv, "fuse", False)
    sd = layer_ref.state_dict_from_npz(fx)
    relation_layers, entity_layers = ultra_modules_from_state_dict(sd)
    assert len(relation_layers) == 6 and len(entity_layers) == 6
    with torch.no_grad():
        rel, h6, score = ultra_forward(relation_layers, entity_layers, sd, fx)
    close(rel, fx["relation_representations"], "relation representations")
    close(h6, fx["entity_hidden_6"], "entity layer 6", tol=5e-5)  # 12 layers end to end
    close(score, fx["score"], "scores", tol=5e-5)


def test_relation_graph_restatements_agree(fx):
    """ultra_b200.synth.relation_graph_of (dense products) == the fixture's sparse restatement of
    build_relation_graph (/root/reference/ultra/tasks.py:144-199): same edges, same order."""
    from ultra_b200.synth import relation_graph_of

    ei, et, num_rel_rel, num_rel_node = relation_graph_of(torch.from_numpy(fx["edge_index"]),
                                                          torch.from_numpy(fx["edge_type"]), int(fx["num_node"]),
                                                          int(fx["num_relation"]))
    assert (num_rel_rel, num_rel_node) == (4, int(fx["num_relation"]))
    assert torch.equal(ei, torch.from_numpy(fx["rel_edge_index"]))
    assert torch.equal(et, torch.from_numpy(fx["rel_edge_type"]))


def test_ultra_nbfnet_runner_matches_fixture(fx, monkeypatch):
    """ultra_b200.nbfnet.UltraNBFNet (the full-stack runner the C3 bench / head-shard scaling use) loads the checkpoint
    with strict=True under the reference's key names and, with rspmm replaced by the oracle, reproduces the fixture -
    including score_all's split-weight scoring (no (B, N, 128) concatenation) against the gathered-tail scores."""
    import ultra_b200.layer as product_layer
    from ultra_b200.nbfnet import UltraNBFNet

    def oracle_rspmm(edge_index, edge_type, edge_weight, relation, input, sum="add", mul="mul"):
        sei, stt, sw, _ = oracle.sort_edges(edge_index.numpy(), edge_type.numpy(), edge_weight.numpy())
        return torch.from_numpy(oracle.c_forward(sum, mul, sei, stt, sw, relation.contiguous().numpy(),
                                                 input.contiguous().numpy()))

    monkeypatch.setattr(product_layer, "generalized_rspmm", oracle_rspmm)
    monkeypatch.setattr(product_layer.GeneralizedRelationalConv, "fuse", False)
    model = UltraNBFNet().eval()
    model.load_state_dict(layer_ref.state_dict_from_npz(fx), strict=True)
    t = lambda k: torch.from_numpy(np.asarray(fx[k]))  # noqa: E731
    rel_graph = (t("rel_edge_index"), t("rel_edge_type"), int(fx["num_relation"]))
    graph = (t("edge_index"), t("edge_type"), int(fx["num_node"]))
    with torch.no_grad():
        rel = model.relation_representations(*rel_graph, t("r_index"))
        close(rel, fx["relation_representations"], "relation representations")
        score = model(rel_graph, graph, t("h_index"), t("r_index"), t("t_index"))
        close(score, fx["score"], "scores", tol=5e-5)
        all_scores = model(rel_graph, graph, t("h_index"), t("r_index"))  # every entity as the tail
        assert all_scores.shape == (4, int(fx["num_node"]))
        close(all_scores.gather(1, t("t_index")), fx["score"], "score_all at the fixture's tails", tol=5e-5)


def test_merge_parallel_edges_preserves_the_sum_aggregate():
    """merge_parallel_edges: (sum_t rel[t]) * x[j] over merged edges == sum over the original parallel edges (oracle on
    both graphs), for a random multigraph where node pairs carry arbitrary subsets of the 4 types; a graph with a
    duplicated (pair, type) edge is refused (every copy counts in the reference)."""
    from ultra_b200.nbfnet import merge_parallel_edges
    g = torch.Generator().manual_seed(9)
    n, R, d = 30, 4, 16
    pairs = torch.randint(0, n, (2, 300), generator=g)
    pairs = torch.unique(pairs[0] * n + pairs[1])
    rows, cols, types = [], [], []
    for key in pairs.tolist():
        subset = torch.nonzero(torch.rand(R, generator=g) < 0.5).flatten().tolist() or [int(torch.randint(0, R, (1,), generator=g))]
        for t in subset:
            rows.append(key // n); cols.append(key % n); types.append(t)
    ei, et = torch.tensor([rows, cols]), torch.tensor(types)
    perm = torch.randperm(ei.shape[1], generator=g)
    ei, et = ei[:, perm], et[perm]
    mei, met, combos = merge_parallel_edges(ei, et, R)
    assert mei.shape[1] == pairs.numel() and combos.shape[1] == R and int(met.max()) + 1 == combos.shape[0]
    rel, x = torch.randn(R, d, generator=g), torch.randn(n, d, generator=g)

    def run(e_i, e_t, relation):
        w = torch.ones(e_i.shape[1])
        sei, stt, sw, _ = oracle.sort_edges(e_i.numpy(), e_t.numpy(), w.numpy())
        return torch.from_numpy(oracle.c_forward("add", "mul", sei, stt, sw, relation.numpy(), x.numpy()))
    exp, got = run(ei, et, rel), run(mei, met, combos @ rel)
    torch.testing.assert_close(got, exp, rtol=1e-5, atol=1e-5 * float(exp.abs().max()))
    dup_ei, dup_et = torch.cat([ei, ei[:, :1]], 1), torch.cat([et, et[:1]])
    assert merge_parallel_edges(dup_ei, dup_et, R) is None


def test_easy_edge_mask_matches_brute_force():
    """UltraNBFNet.easy_edge_mask == remove_easy_edges' mask (/root/reference/ultra/base_nbfnet.py:54-77), by brute force."""
    from ultra_b200.nbfnet import UltraNBFNet
    g = torch.Generator().manual_seed(13)
    N, R, E = 40, 6, 500  # R counts inverse relations: r + R/2 is the inverse of r
    ei = torch.randint(0, N, (2, E), generator=g)
    et = torch.randint(0, R, (E,), generator=g)
    pick = torch.randint(0, E, (8,), generator=g)
    pick = pick[et[pick] < R // 2]
    h, t, r = ei[0, pick], ei[1, pick], et[pick]
    for one_hop in (False, True):
        mask = UltraNBFNet.easy_edge_mask(ei, et, N, R, h, t, r, remove_one_hop=one_hop)
        batch = set()
        for a, b, c in zip(h.tolist(), t.tolist(), r.tolist()):
            batch.add((a, b, c)); batch.add((b, a, c + R // 2))
        pairs = {(a, b) for a, b, _ in batch}
        exp = [((u, v) not in pairs) if one_hop else ((u, v, w) not in batch)
               for u, v, w in zip(ei[0].tolist(), ei[1].tolist(), et.tolist())]
        assert mask.tolist() == exp
        assert not mask[pick].any()
