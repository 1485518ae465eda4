"""Generates tests/golden/layer_golden.npz: the layer-level oracle of SURVEY.md section 8c with REAL weights.

Run in the build container only (needs /root/reference):   python tests/golden/make_layer_golden.py

What it does: loads the reference's shipped checkpoint /root/reference/ckpts/ultra_3g.pth (168,705 parameters), builds a
small synthetic KG with inverse edges and its relation graph (sparse restatement of build_relation_graph,
/root/reference/ultra/tasks.py:144-199), and runs the full ULTRA forward - relation-graph NBFNet (6 layers) -> relation
representations -> entity-graph NBFNet (6 layers) -> tail scores - through ``oracle/layer_ref.py`` with every rspmm
computed by the UNMODIFIED reference CPU extension (oracle/_ref).  The reference's own ``ultra/layers.py`` /
``ultra/models.py`` cannot be imported (torch_geometric / torch_scatter are not installed), so the module code around the
kernel is the functional restatement in oracle/layer_ref.py; weights and rspmm are the reference's.

The fixture stores the checkpoint tensors (so tests need neither /root/reference nor the .pth), the graph, the batch and
the outputs of both stacks.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import layer_ref  # noqa: E402
from ultra_b200.synth import synth_kg  # noqa: E402  (graph generator only; no product compute)

CKPT = "/root/reference/ckpts/ultra_3g.pth"


def build_relation_graph(edge_index, edge_type, num_node, num_rel):
    """tasks.py:144-199 without torch_scatter / PyG: the four sparse products, coalesced; only their index pattern is
    used downstream (edge weights of the relation graph are all ones, models.py:68)."""
    def incidence(node):  # unique (node, relation) pairs, tasks.py:153,166
        return torch.vstack([node, edge_type]).T.unique(dim=0)

    def mats(pairs):
        deg = torch.bincount(pairs[:, 0], minlength=num_node).to(torch.float32)  # tasks.py:154,168
        transposed = torch.sparse_coo_tensor(torch.flip(pairs, dims=[1]).T, torch.ones(pairs.shape[0]) / deg[pairs[:, 0]],
                                             (num_rel, num_node))
        plain = torch.sparse_coo_tensor(pairs.T, torch.ones(pairs.shape[0]), (num_node, num_rel))
        return transposed, plain

    EhT, Eh = mats(incidence(edge_index[0]))
    EtT, Et = mats(incidence(edge_index[1]))
    prods = [torch.sparse.mm(EhT, Eh), torch.sparse.mm(EtT, Et), torch.sparse.mm(EhT, Et), torch.sparse.mm(EtT, Eh)]
    eis, ets = [], []
    for k, a in enumerate(prods):  # 0 h2h, 1 t2t, 2 h2t, 3 t2h (tasks.py:186-189)
        idx = a.coalesce().indices()
        eis.append(idx)
        ets.append(torch.full((idx.shape[1],), k, dtype=torch.long))
    return torch.cat(eis, dim=1), torch.cat(ets)


def main():
    torch.set_num_threads(1)
    sd = torch.load(CKPT, map_location="cpu", weights_only=False)["model"]
    sd = {k: v.float().contiguous() for k, v in sd.items()}
    assert sum(v.numel() for v in sd.values()) == 168705
    num_node, num_rel_base = 120, 10
    edge_index, edge_type, num_rel = synth_kg(num_node, num_rel_base, 260, alpha=1.0, seed=11)
    rel_ei, rel_et = build_relation_graph(edge_index, edge_type, num_node, num_rel)
    g = torch.Generator().manual_seed(5)
    B, T = 4, 6
    h_index = torch.randint(0, num_node, (B,), generator=g)
    r_index = torch.randint(0, num_rel, (B,), generator=g)
    t_index = torch.randint(0, num_node, (B, T), generator=g)
    with torch.no_grad():
        rel_repr = layer_ref.relation_stack(sd, rel_ei, rel_et, num_rel, r_index, rspmm="ref")
        feature, hiddens = layer_ref.entity_stack(sd, edge_index, edge_type, num_node, rel_repr, h_index, r_index,
                                                  rspmm="ref", all_hiddens=True)
        score = layer_ref.score(sd, feature, t_index)
    data = {"weights/" + k: v.numpy() for k, v in sd.items()}
    data.update({
        "edge_index": edge_index.numpy(), "edge_type": edge_type.numpy(), "num_node": np.int64(num_node),
        "num_relation": np.int64(num_rel), "rel_edge_index": rel_ei.numpy(), "rel_edge_type": rel_et.numpy(),
        "h_index": h_index.numpy(), "r_index": r_index.numpy(), "t_index": t_index.numpy(),
        "relation_representations": rel_repr.numpy(), "entity_hidden_1": hiddens[0].numpy(),
        "entity_hidden_6": hiddens[-1].numpy(), "score": score.numpy(),
    })
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "layer_golden.npz")
    np.savez_compressed(out, **data)
    print("wrote %s: %d arrays, %d bytes; |rel_repr| max %.3f, |h6| max %.3f, score range [%.3f, %.3f], rel-graph edges %d"
          % (out, len(data), os.path.getsize(out), rel_repr.abs().max(), hiddens[-1].abs().max(), score.min(), score.max(),
             rel_ei.shape[1]))


if __name__ == "__main__":
    main()
