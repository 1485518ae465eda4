"""GPU suite: the PyG-free GeneralizedRelationalConv restatement (ultra_b200.layer) on top of the CUDA rspmm against
a dense PyTorch formula of /root/reference/ultra/layers.py:183-240 evaluated on the CPU (fp32, 1e-4 tolerance for the
Linear+LayerNorm epilogue)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def dense_layer_reference(layer, input, query, boundary, edge_index, edge_type):
    """Plain torch restatement: messages materialised per edge, index_add / scatter_reduce aggregation."""
    B, N, d = input.shape
    if layer.dependent:
        relation = layer.relation_linear(query).view(B, layer.num_relation, d)
    elif not layer.project_relations:
        relation = layer.relation.weight.expand(B, -1, -1)
    else:
        relation = layer.relation_projection(layer.relation)
    x = input.transpose(0, 1).flatten(1)
    rel = relation.transpose(0, 1).flatten(1)
    bd = boundary.transpose(0, 1).flatten(1)
    row, col = edge_index
    msg = rel[edge_type] * x[col] if layer.message_func == "distmult" else rel[edge_type] + x[col]
    D = x.shape[1]
    idx = row[:, None].expand(-1, D)
    deg = torch.bincount(edge_index[1], minlength=N).float().unsqueeze(-1) + 1
    big = torch.finfo(torch.float32).max
    s = torch.zeros(N, D).index_add_(0, row, msg)
    if layer.aggregate_func == "sum":
        upd = s + bd
    elif layer.aggregate_func == "mean":
        upd = (s + bd) / deg
    elif layer.aggregate_func == "max":
        upd = torch.max(torch.full((N, D), -big).scatter_reduce(0, idx, msg, "amax"), bd)
    else:
        mx = torch.max(torch.full((N, D), -big).scatter_reduce(0, idx, msg, "amax"), bd)
        mn = torch.min(torch.full((N, D), big).scatter_reduce(0, idx, msg, "amin"), bd)
        rel2, x2 = rel ** 2, x ** 2
        msg2 = rel2[edge_type] * x2[col] if layer.message_func == "distmult" else rel2[edge_type] + x2[col]
        sq = torch.zeros(N, D).index_add_(0, row, msg2)
        mean = (s + bd) / deg
        sq_mean = (sq + bd ** 2) / deg
        std = (sq_mean - mean ** 2).clamp(min=layer.eps).sqrt()
        feats = torch.cat([mean.unsqueeze(-1), mx.unsqueeze(-1), mn.unsqueeze(-1), std.unsqueeze(-1)], -1).flatten(-2)
        scale = deg.log()
        scale = scale / scale.mean()
        scales = torch.cat([torch.ones_like(scale), scale, 1 / scale.clamp(min=1e-2)], dim=-1)
        upd = (feats.unsqueeze(-1) * scales.unsqueeze(-2)).flatten(-2)
    upd = upd.view(N, B, -1).transpose(0, 1)
    return layer.update(upd, input)


@pytest.mark.parametrize("message,aggregate", [("distmult", "sum"), ("transe", "sum"), ("distmult", "mean"),
                                               ("distmult", "max"), ("transe", "pna"), ("distmult", "pna")])
@pytest.mark.parametrize("mode", ["embedding", "project", "dependent"])
def test_layer_matches_dense_formula(cuda_device, message, aggregate, mode):
    from ultra_b200.layer import GeneralizedRelationalConv
    torch.manual_seed(0)
    B, N, R, d, E = 3, 60, 6, 64, 700
    layer = GeneralizedRelationalConv(d, d, R, d, message_func=message, aggregate_func=aggregate, layer_norm=True,
                                      dependent=(mode == "dependent"), project_relations=(mode == "project"))
    edge_index = torch.randint(0, N - 2, (2, E))
    edge_type = torch.randint(0, R, (E,))
    x, bd, q = torch.randn(B, N, d), torch.randn(B, N, d), torch.randn(B, d)
    if mode == "project":
        layer.relation = torch.randn(B, R, d)
    with torch.no_grad():
        ref = dense_layer_reference(layer, x, q, bd, edge_index, edge_type)
    dev = cuda_device
    glayer = GeneralizedRelationalConv(d, d, R, d, message_func=message, aggregate_func=aggregate, layer_norm=True,
                                       dependent=(mode == "dependent"), project_relations=(mode == "project"))
    glayer.load_state_dict(layer.state_dict())
    glayer = glayer.to(dev)
    if mode == "project":
        glayer.relation = layer.relation.to(dev)
    xg = x.to(dev).requires_grad_()
    out = glayer(xg, q.to(dev), bd.to(dev), edge_index.to(dev), edge_type.to(dev), (N, N))
    torch.testing.assert_close(out.detach().cpu(), ref, rtol=1e-4, atol=1e-4)
    out.sum().backward()  # autograd flows through the custom Function
    assert xg.grad is not None and bool(torch.isfinite(xg.grad).all())
    assert glayer.linear.weight.grad is not None


def test_bellmanford_stack_head_sharding_is_exact(cuda_device):
    """Running the 6-layer entity stack on a slice of heads gives bit-identical states for those heads."""
    from ultra_b200.layer import GeneralizedRelationalConv, bellmanford_entity
    from ultra_b200.shard import shard_heads
    torch.manual_seed(1)
    dev = cuda_device
    B, N, R, d, E = 8, 200, 10, 64, 3000
    layers = [GeneralizedRelationalConv(d, d, R, d, "distmult", "sum", layer_norm=True, project_relations=True).to(dev)
              for _ in range(3)]
    rel_repr = torch.randn(B, R, d, device=dev)
    query = torch.randn(B, d, device=dev)
    h_index = torch.randint(0, N, (B,), device=dev)
    ei = torch.randint(0, N, (2, E), device=dev)
    et = torch.randint(0, R, (E,), device=dev)
    with torch.no_grad():
        for layer in layers:
            layer.relation = rel_repr
        full = bellmanford_entity(layers, query, h_index, ei, et, N)
        for rank in range(2):
            for layer in layers:
                layer.relation = shard_heads(rel_repr, rank, 2)
            part = bellmanford_entity(layers, shard_heads(query, rank, 2), shard_heads(h_index, rank, 2), ei, et, N)
            torch.testing.assert_close(part, shard_heads(full, rank, 2), rtol=1e-5, atol=1e-5)
