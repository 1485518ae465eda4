"""TEST INFRASTRUCTURE ONLY - layer-level oracle (last row).

A PyG-free, CPU, functional restatement of the caller side of the hot path, written independently of
``ultra_b200/layer.py`` (which is the product-side mirror) and operating directly on a reference checkpoint's
state-dict (``/root/reference/ckpts/ultra_3g.pth``: keys ``relation_model.layers.N.*``, ``entity_model.layers.N.*``,
``entity_model.mlp.*``):

* ``conv``            - GeneralizedRelationalConv.forward on the fused route: relation features
                        (/root/reference/ultra/layers.py:67-83), message_and_aggregate with aggregate_func == "sum"
                        (layers.py:183-201,230) and update (layers.py:233-240);
* ``relation_stack``  - RelNBFNet.bellmanford (/root/reference/ultra/models.py:55-95), concat_hidden = False;
* ``entity_stack``    - EntityNBFNet.bellmanford (models.py:131-175) after forward() set query / layer.relation
                        (models.py:177-186);
* ``score``           - tail gather + MLP (models.py:202-208).

The rspmm inside is pluggable: ``rspmm="ref"`` calls the UNMODIFIED reference CPU extension (oracle/_ref) through the
restated per-call-sort wrapper, ``rspmm="c"`` the plain-C restatement.  Configuration as shipped in the reference's
config/*/*.yaml for ultra_3g: 6 + 6 layers, 64-d, distmult, sum, layer_norm, relu, short_cut.
"""
import numpy as np
import torch
from torch.nn import functional as F

from . import c_forward, sort_edges
from . import ref as _ref


def _rspmm(kind, edge_index, edge_type, edge_weight, relation, input):
    """generalized_rspmm(..., sum="add", mul="mul") (rspmm.py:167-178) on CPU tensors."""
    if kind == "ref":
        return _ref.ref_generalized_rspmm(edge_index, edge_type, edge_weight, relation.contiguous(),
                                          input.contiguous(), sum="add", mul="mul")
    sei, stt, sw, _ = sort_edges(edge_index.numpy(), edge_type.numpy(), edge_weight.numpy())
    return torch.from_numpy(c_forward("add", "mul", sei, stt, sw, relation.contiguous().numpy(),
                                      input.contiguous().numpy()))


def conv(sd, prefix, input, boundary, edge_index, edge_type, relation_repr=None, rspmm="ref"):
    """One GeneralizedRelationalConv layer.  input, boundary: (B, N, 64).  relation_repr: (B, R, 64) for the entity
    model (project_relations = True), None for the relation model (per-layer nn.Embedding(4, 64))."""
    batch_size, num_node = input.shape[:2]
    if relation_repr is None:  # layers.py:76
        relation = sd[prefix + "relation.weight"].expand(batch_size, -1, -1)
    else:  # layers.py:80: Linear -> ReLU -> Linear
        hidden = F.relu(F.linear(relation_repr, sd[prefix + "relation_projection.0.weight"],
                                 sd[prefix + "relation_projection.0.bias"]))
        relation = F.linear(hidden, sd[prefix + "relation_projection.2.weight"], sd[prefix + "relation_projection.2.bias"])
    edge_weight = torch.ones(edge_type.shape[0])  # models.py:68,144
    # layers.py:190-192: (B, N, d) -> (N, B*d)
    x = input.transpose(0, 1).flatten(1)
    rel = relation.transpose(0, 1).flatten(1)
    bnd = boundary.transpose(0, 1).flatten(1)
    update = _rspmm(rspmm, edge_index, edge_type, edge_weight, rel, x) + bnd  # layers.py:199-201
    update = update.view(num_node, batch_size, -1).transpose(0, 1)  # layers.py:230
    out = F.linear(torch.cat([input, update], dim=-1), sd[prefix + "linear.weight"], sd[prefix + "linear.bias"])
    out = F.layer_norm(out, (out.shape[-1],), sd[prefix + "layer_norm.weight"], sd[prefix + "layer_norm.bias"])
    return F.relu(out)  # layers.py:233-240


def _stack(sd, model, num_layers, boundary, edge_index, edge_type, relation_repr, rspmm):
    layer_input = boundary
    hiddens = []
    for i in range(num_layers):
        hidden = conv(sd, "%s.layers.%d." % (model, i), layer_input, boundary, edge_index, edge_type, relation_repr, rspmm)
        if hidden.shape == layer_input.shape:  # short_cut (models.py:76-78,158-160)
            hidden = hidden + layer_input
        hiddens.append(hidden)
        layer_input = hidden
    return hiddens


def relation_stack(sd, rel_edge_index, rel_edge_type, num_rel_node, query_rels, num_layers=6, rspmm="ref"):
    """RelNBFNet.forward: (B, num_rel_node, 64) relation representations conditioned on each query relation."""
    query = torch.ones(query_rels.shape[0], 64)  # models.py:59
    boundary = torch.zeros(query_rels.shape[0], num_rel_node, 64)
    boundary.scatter_add_(1, query_rels.view(-1, 1, 1).expand(-1, 1, 64), query.unsqueeze(1))  # models.py:60-66
    return _stack(sd, "relation_model", num_layers, boundary, rel_edge_index, rel_edge_type, None, rspmm)[-1]


def entity_stack(sd, edge_index, edge_type, num_node, relation_repr, h_index, r_index, num_layers=6, rspmm="ref",
                 all_hiddens=False):
    """EntityNBFNet.bellmanford -> node_feature (B, N, 128) = [h_6 | query] (concat_hidden = False)."""
    batch_size = h_index.shape[0]
    query = relation_repr[torch.arange(batch_size), r_index]  # models.py:134
    boundary = torch.zeros(batch_size, num_node, 64)
    boundary.scatter_add_(1, h_index.view(-1, 1, 1).expand(-1, 1, 64), query.unsqueeze(1))  # models.py:135-141
    hiddens = _stack(sd, "entity_model", num_layers, boundary, edge_index, edge_type, relation_repr, rspmm)
    node_query = query.unsqueeze(1).expand(-1, num_node, -1)
    feature = torch.cat([hiddens[-1], node_query], dim=-1)  # models.py:166-170
    return (feature, hiddens) if all_hiddens else feature


def score(sd, feature, t_index):
    """models.py:202-208: gather the tails' features, MLP 128 -> 128 -> 1."""
    index = t_index.unsqueeze(-1).expand(-1, -1, feature.shape[-1])
    f = feature.gather(1, index)
    f = F.relu(F.linear(f, sd["entity_model.mlp.0.weight"], sd["entity_model.mlp.0.bias"]))
    return F.linear(f, sd["entity_model.mlp.2.weight"], sd["entity_model.mlp.2.bias"]).squeeze(-1)


def state_dict_from_npz(npz, prefix="weights/"):
    """The checkpoint tensors stored in a golden fixture -> {key: torch tensor}."""
    return {k[len(prefix):]: torch.from_numpy(np.asarray(npz[k])) for k in npz.files if k.startswith(prefix)}
