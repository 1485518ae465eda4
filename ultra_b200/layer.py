"""PyG-free restatement of the ONE caller of the hot path: GeneralizedRelationalConv's fused route
(/root/reference/ultra/layers.py:67-88 forward, :183-231 message_and_aggregate, :233-240 update).

The reference class is a torch_geometric MessagePassing subclass; PyG is not part of this environment and the layer
is outside the rspmm contract, so this module exists for two things only: measuring "NBFNet-layer ms"
(BASELINE.json metric) and layer-level parity tests.  Parameter names match the reference so its checkpoints'
state-dict keys (`linear`, `layer_norm`, `relation`, `relation_linear`, `relation_projection.{0,2}`) load unchanged.
Only the fused path is restated: `rotate` messages and the PyG unfused fallback (layers.py:91-94,135-181) are not.
"""
import torch
from torch import nn
from torch.nn import functional as F

from .rspmm import generalized_rspmm, generalized_rspmm_pna
from . import fused


class GeneralizedRelationalConv(nn.Module):
    eps = 1e-6
    message2mul = {"transe": "add", "distmult": "mul"}  # layers.py:16-19

    def __init__(self, input_dim, output_dim, num_relation, query_input_dim, message_func="distmult",
                 aggregate_func="pna", layer_norm=False, activation="relu", dependent=False,
                 project_relations=False):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.num_relation, self.query_input_dim = num_relation, query_input_dim
        self.message_func, self.aggregate_func = message_func, aggregate_func
        self.dependent, self.project_relations = dependent, project_relations
        self.layer_norm = nn.LayerNorm(output_dim) if layer_norm else None
        self.activation = getattr(F, activation) if isinstance(activation, str) else activation
        self.linear = nn.Linear(input_dim * (13 if aggregate_func == "pna" else 2), output_dim)  # layers.py:44-47
        if dependent:
            self.relation_linear = nn.Linear(query_input_dim, num_relation * input_dim)
        elif not project_relations:
            self.relation = nn.Embedding(num_relation, input_dim)
        else:
            self.relation = None  # set by the caller after the relation-graph pass (models.py:184-185)
            self.relation_projection = nn.Sequential(nn.Linear(input_dim, input_dim), nn.ReLU(),
                                                     nn.Linear(input_dim, input_dim))

    fuse = True  # class-wide switch for the head-major / fused-epilogue fast path (set False to force the restatement)

    def forward(self, input, query, boundary, edge_index, edge_type, size, edge_weight=None, shortcut=False,
                relation=None):
        """input, boundary: (B, N, d); query: (B, d_q); returns (B, N, output_dim).  layers.py:67-88.
        `shortcut=True` (not in the reference signature) also adds `input` to the result, i.e. the residual of the
        Bellman-Ford loop (models.py:158-160), so that it can be folded into the fused epilogue.
        `relation` (not in the reference signature): this layer's already-projected (B, R, d) relation, when the caller
        projected all layers at once (ultra_b200.nbfnet); None = compute it here like the reference."""
        batch_size = len(query)
        shared = False  # relation is ONE (R, d) block for every head (embedding mode, layers.py:76)
        if relation is not None:
            shared = relation.dim() == 2  # one (R', d) block for every head (e.g. the merged relation-graph table)
        elif self.dependent:
            relation = self.relation_linear(query).view(batch_size, self.num_relation, self.input_dim)
        elif not self.project_relations:
            relation, shared = self.relation.weight, True
        else:
            relation = self.relation_projection(self.relation)
        if edge_weight is None:
            edge_weight = torch.ones(len(edge_type), device=input.device, dtype=input.dtype)
        if self.message_func not in self.message2mul:
            raise ValueError("Unknown message function `%s`" % self.message_func)
        fast = (self.fuse and self.aggregate_func in ("sum", "mean", "max") and fused.heads_layout_ok(input)
                and not edge_weight.requires_grad and relation.dtype == input.dtype)
        if fast:
            update = self.message_and_aggregate_heads(edge_index, input, relation, boundary, edge_type, edge_weight,
                                                      size[0])
            if fused.nbfnet_update_ok(input, update, self.linear, self.activation, self.layer_norm):
                from .plan import plan_cache
                plan = plan_cache.get(edge_index, edge_type, input.shape[1], relation.shape[1])  # cache hit
                return fused.nbfnet_update(input, update, self.linear, self.layer_norm,
                                           relu=self.activation is not None, shortcut=shortcut, sm_limit=plan.sm_limit)
        else:
            if shared:
                relation = relation.expand(batch_size, -1, -1)
            update = self.message_and_aggregate(edge_index, input, relation, boundary, edge_type, edge_weight, size[0])
        output = self.update(update, input)
        return output + input if shortcut else output

    def message_and_aggregate_heads(self, edge_index, input, relation, boundary, edge_type, edge_weight, dim_size):
        """Same result as message_and_aggregate for sum / mean / max, but the (B, N, d) state is consumed and produced
        in place by the head-major kernels: no transpose copies, boundary folded into the store."""
        mul = self.message2mul[self.message_func]
        # relation is (B, R, d), or the (R, d) embedding itself in the relation-graph stack: rspmm_heads reads that one
        # block for every head (head stride 0), the reference's expand (layers.py:76) is never materialised
        if self.aggregate_func == "max":
            return fused.rspmm_heads(edge_index, edge_type, edge_weight, relation, input, boundary, sum="max", mul=mul)
        update = fused.rspmm_heads(edge_index, edge_type, edge_weight, relation, input, boundary, sum="add", mul=mul)
        if self.aggregate_func == "mean":
            update = update / self._degree_out(edge_index, dim_size, input).unsqueeze(0)
        return update

    def message_and_aggregate(self, edge_index, input, relation, boundary, edge_type, edge_weight, dim_size):
        """layers.py:183-231: (B,N,d) -> (N,B*d) copies, rspmm, boundary as the self-loop message."""
        batch_size, num_node = input.shape[:2]
        input = input.transpose(0, 1).flatten(1)
        relation = relation.transpose(0, 1).flatten(1)
        boundary = boundary.transpose(0, 1).flatten(1)
        mul = self.message2mul[self.message_func]
        rspmm = lambda rel, x, s: generalized_rspmm(edge_index, edge_type, edge_weight, rel, x, sum=s, mul=mul)
        if self.aggregate_func == "sum":
            update = rspmm(relation, input, "add") + boundary
        elif self.aggregate_func == "mean":
            update = (rspmm(relation, input, "add") + boundary) / self._degree_out(edge_index, dim_size, input)
        elif self.aggregate_func == "max":
            update = torch.max(rspmm(relation, input, "max"), boundary)
        elif self.aggregate_func == "pna":  # layers.py:208-226
            degree_out = self._degree_out(edge_index, dim_size, input)
            if self.fuse and input.is_cuda and input.shape[1] % 4 == 0:  # one gather pass for all four aggregates
                sum_, sq_sum, max_, min_ = generalized_rspmm_pna(edge_index, edge_type, edge_weight, relation, input,
                                                                 mul=mul)
            else:  # the reference's four calls (layers.py:211-215)
                sum_, sq_sum = rspmm(relation, input, "add"), rspmm(relation ** 2, input ** 2, "add")
                max_, min_ = rspmm(relation, input, "max"), rspmm(relation, input, "min")
            max_ = torch.max(max_, boundary)
            min_ = torch.min(min_, boundary)
            mean = (sum_ + boundary) / degree_out
            sq_mean = (sq_sum + boundary ** 2) / degree_out
            std = (sq_mean - mean ** 2).clamp(min=self.eps).sqrt()
            features = torch.cat([mean.unsqueeze(-1), max_.unsqueeze(-1), min_.unsqueeze(-1), std.unsqueeze(-1)],
                                 dim=-1).flatten(-2)
            scale = degree_out.log()
            scale = scale / scale.mean()
            scales = torch.cat([torch.ones_like(scale), scale, 1 / scale.clamp(min=1e-2)], dim=-1)
            update = (features.unsqueeze(-1) * scales.unsqueeze(-2)).flatten(-2)
        else:
            raise ValueError("Unknown aggregation function `%s`" % self.aggregate_func)
        return update.view(num_node, batch_size, -1).transpose(0, 1)

    @staticmethod
    def _degree_out(edge_index, dim_size, like):
        # PyG `degree(index, dim_size)` with index = edge_index[1] (flow source_to_target), +1 for the self loop
        deg = torch.bincount(edge_index[1], minlength=dim_size).to(like.dtype)
        return deg.unsqueeze(-1) + 1

    def update(self, update, input):  # layers.py:233-240
        output = self.linear(torch.cat([input, update], dim=-1))
        if self.layer_norm is not None:
            output = self.layer_norm(output)
        if self.activation:
            output = self.activation(output)
        return output


def bellmanford_entity(layers, query, h_index, edge_index, edge_type, num_node, short_cut=True, hook=None):
    """Entity-graph Bellman-Ford loop of EntityNBFNet.bellmanford (/root/reference/ultra/models.py:131-175) over a
    list of GeneralizedRelationalConv layers whose `.relation` is already set.  query: (B, d), h_index: (B,).
    `hook(layer_idx, hidden)` is called after every layer (used by the head-sharded runner to launch all-gathers)."""
    batch_size = len(h_index)
    boundary = torch.zeros(batch_size, num_node, query.shape[-1], device=query.device, dtype=query.dtype)
    index = h_index.unsqueeze(-1).expand_as(query)
    boundary.scatter_add_(1, index.unsqueeze(1), query.unsqueeze(1))
    edge_weight = torch.ones(edge_index.shape[1], device=query.device, dtype=query.dtype)
    size = (num_node, num_node)
    layer_input = boundary
    for li, layer in enumerate(layers):
        residual = short_cut and layer.output_dim == layer_input.shape[-1]
        hidden = layer(layer_input, query, boundary, edge_index, edge_type, size, edge_weight, shortcut=residual)
        if hook is not None:
            hook(li, hidden)
        layer_input = hidden
    return layer_input


class GraphedBellmanFord:
    """The Bellman-Ford layer loop captured once as a CUDA graph and replayed.

    The reference cannot do this: every rspmm call synchronises the host (sortedness assert, rspmm.py:15-17).  With the
    cached plan the whole loop is stream-ordered, so for the launch-bound regime (relation graph, small batches: a layer
    is a few hundred microseconds of GPU work behind ~10 launches and their Python overhead) one graph launch replaces
    ~60.  Inference only; shapes, graph (edge tensors) and layer weights are frozen at capture, the per-call inputs
    (query, h_index, and the layers' `relation` representations when given) are copied into static buffers.
    """

    def __init__(self, layers, query, h_index, edge_index, edge_type, num_node, relation=None, short_cut=True):
        self.layers = list(layers)
        self.static_query = query.clone()
        self.static_h_index = h_index.clone()
        self.static_relation = None if relation is None else relation.clone()
        self.edge_index, self.edge_type, self.num_node, self.short_cut = edge_index, edge_type, num_node, short_cut
        if self.static_relation is not None:
            for layer in self.layers:
                layer.relation = self.static_relation
        side = torch.cuda.Stream(device=query.device)
        side.wait_stream(torch.cuda.current_stream(query.device))
        with torch.cuda.stream(side), torch.no_grad():  # warm-up: plan build, weight bind, scratch allocation
            for _ in range(2):
                self._run()
        torch.cuda.current_stream(query.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = self._run()

    def _run(self):
        return bellmanford_entity(self.layers, self.static_query, self.static_h_index, self.edge_index, self.edge_type,
                                  self.num_node, short_cut=self.short_cut)

    @torch.no_grad()
    def __call__(self, query, h_index, relation=None):
        self.static_query.copy_(query)
        self.static_h_index.copy_(h_index)
        if relation is not None:
            self.static_relation.copy_(relation)
        self.graph.replay()
        # the captured weight bind rewrote the plan's device weights: an eager call that bound another weight tensor
        # before must re-bind (the Python-side key would otherwise still name that tensor)
        from .plan import plan_cache
        for entry in list(plan_cache._entries.values()):
            if entry[1] is self.edge_index:
                entry[0].invalidate_weight()
        return self.static_out
