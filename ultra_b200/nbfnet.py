"""PyG-free restatement of the two NBFNet stacks around the hot path - what calls `rspmm` 12 times per forward:

    Ultra.forward                 /root/reference/ultra/models.py:18-26
      RelNBFNet.bellmanford       models.py:55-95     6 layers over the relation graph (4 relation rows, one embedding
                                                      shared by every head, boundary = ones at the query relation)
      EntityNBFNet.forward        models.py:177-209   6 layers over the entity graph (relations = per-layer projections of
                                                      the relation representations), tail features -> MLP -> scores

It exists for measurement ("NBFNet-layer ms", the C3 full-stack configuration of BASELINE.json, head-shard scaling) and
layer-level parity (the shipped checkpoints load with `load_state_dict(ckpt["model"])`, strict: module and parameter
names are the reference's).  Datasets, samplers, losses, the training loop and the query executor stay out of scope.

What is done differently from the reference, all bit-compatible restructurings of the same arithmetic:
  * both stacks run on the (B, N, d) layout through the head-major kernels + the fused tcgen05 epilogue (ultra_b200.fused)
  * the relation stack's shared embedding is read with head stride 0 (never expanded to (B, 4, d));
  * the six relation_projection MLPs of the entity stack are ONE launch (ultra_relation_project_f32);
  * scoring every entity as a tail ("all negatives", /root/reference/ultra/tasks.py:79-91 + models.py:166-170,202-208)
    never builds the (B, N, 128) concatenation: mlp[0](cat[h, q]) = h . Wa^T + (q . Wb^T + b);
  * the whole forward is stream-ordered (no host syncs), so `CudaGraphRunner` can capture it once and replay it.
One restructuring is NOT bit-compatible (it changes roundings at the 1e-7 level, inside the 1e-5 bar; switch it off with
`model.merge_relation_graph = False`): for distmult + sum the parallel edges of the relation graph are merged, see
`merge_parallel_edges` - 4x fewer edges on the dense relation graph of the FB15k-237-scale KG.
"""
import torch
from torch import nn

from . import fused
from .layer import GeneralizedRelationalConv


def merge_parallel_edges(edge_index, edge_type, num_relation):
    """Merge the parallel edges of a graph with few relation types: all edges (i <- j, t) between the same node pair
    become ONE edge whose type names the SET of original types (bit mask - 1), for distmult + sum aggregation:
        sum_t rel[t] * x[j]  ==  (sum_t rel[t]) * x[j]        (one rounding of the combined relation row instead of one per
                                                              product: ~1e-7 relative, inside the 1e-5 parity bar)
    ULTRA's relation graph has 4 edge types (h2h, t2t, h2t, t2h, /root/reference/ultra/tasks.py:144-199) and is dense -
    relations that share a head usually share tails too - so most node pairs carry several of them: on the relation
    graph of the FB15k-237-scale KG (474 nodes, 898 704 edges) this leaves 224 676 edges, 4x fewer gathers per layer.
    Returns (edge_index', edge_type', combos) with combos (C, num_relation) 0/1 fp32: merged relation table = combos @ rel,
    which also routes the gradient back to the original rows through autograd."""
    if num_relation > 6:
        raise ValueError("merge_parallel_edges is meant for graphs with a handful of relation types")
    n = int(edge_index.max()) + 1 if edge_index.numel() else 1
    key = edge_index[0] * n + edge_index[1]
    uniq, inv = torch.unique(key, return_inverse=True)
    mask = torch.zeros(uniq.numel(), dtype=torch.long, device=edge_index.device)
    # duplicates of the same (pair, type) must each count (the reference sums every edge): only merge when there are none
    full = key * num_relation + edge_type
    if torch.unique(full).numel() != full.numel():
        return None
    mask.scatter_add_(0, inv, torch.ones_like(inv) << edge_type)
    used, new_type = torch.unique(mask, return_inverse=True)
    combos = ((used.unsqueeze(1) >> torch.arange(num_relation, device=used.device)) & 1).to(torch.float32)
    merged_index = torch.stack([uniq // n, uniq % n])
    return merged_index.contiguous(), new_type.contiguous(), combos


class _Stack(nn.Module):
    """Holds `layers` (+ optionally `mlp`) under the reference's attribute names (state-dict compatibility)."""

    def __init__(self, layers, mlp=None):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        if mlp is not None:
            self.mlp = mlp


class UltraNBFNet(nn.Module):
    """relation_model (RelNBFNet) + entity_model (EntityNBFNet) of the shipped ultra_* checkpoints: 6 + 6 layers, dim 64,
    distmult messages, sum aggregation, layer norm, ReLU, shortcut (/root/reference/config/transductive/inference.yaml)."""

    def __init__(self, dim=64, num_layers=6, message_func="distmult", aggregate_func="sum", layer_norm=True,
                 short_cut=True, num_mlp_layers=2):
        super().__init__()
        self.dim, self.short_cut = dim, short_cut
        conv = lambda **kw: GeneralizedRelationalConv(dim, dim, kw.pop("num_relation"), dim, message_func,  # noqa: E731
                                                      aggregate_func, layer_norm, "relu", dependent=False, **kw)
        self.relation_model = _Stack([conv(num_relation=4) for _ in range(num_layers)])
        mlp = []
        for _ in range(num_mlp_layers - 1):  # models.py:122-128
            mlp += [nn.Linear(2 * dim, 2 * dim), nn.ReLU()]
        mlp.append(nn.Linear(2 * dim, 1))
        self.entity_model = _Stack([conv(num_relation=1, project_relations=True) for _ in range(num_layers)],
                                   nn.Sequential(*mlp))
        self._ones = {}
        self._merged = {}
        self.merge_relation_graph = message_func == "distmult" and aggregate_func == "sum"  # (mean divides by the edge count)

    # ---- pieces ------------------------------------------------------------------------------------------------------
    def _unit_weight(self, num_edge, like):
        key = (num_edge, like.device, like.dtype)
        if key not in self._ones:  # one tensor per graph size: the plan's weight binding is keyed on tensor identity
            self._ones[key] = torch.ones(num_edge, device=like.device, dtype=like.dtype)
        return self._ones[key]

    @staticmethod
    def _boundary(index, value, num_node):
        """zeros (B, num_node, d) with `value[b]` at node `index[b]` (models.py:59-66,135-141)."""
        boundary = torch.zeros(len(index), num_node, value.shape[-1], device=value.device, dtype=value.dtype)
        boundary.scatter_add_(1, index.view(-1, 1, 1).expand(-1, 1, value.shape[-1]), value.unsqueeze(1))
        return boundary

    @staticmethod
    def easy_edge_mask(edge_index, edge_type, num_node, num_relation, h_index, t_index, r_index, remove_one_hop=False):
        """keep-mask of remove_easy_edges (/root/reference/ultra/base_nbfnet.py:54-77): False for every edge
        (h, t, r) / (t, h, r + R/2) of the training batch (or every edge between the pairs when remove_one_hop).  The
        reference's edge_match (tasks.py:7-39) is a hashing + bucketize search; an integer-key `isin` gives the same set.
        Returns a bool (E,) tensor: feed it to the forward as `edge_mask` - the base graph's plan is re-cut for it on the
        device (GraphPlan.apply_edge_mask), nothing is re-sorted and the host is never synchronised."""
        h = torch.cat([h_index.flatten(), t_index.flatten()])
        t = torch.cat([t_index.flatten(), h_index.flatten()])
        key = edge_index[0] * num_node + edge_index[1]
        easy = h * num_node + t
        if not remove_one_hop:
            r = torch.cat([r_index.flatten(), r_index.flatten() + num_relation // 2])
            key = key * num_relation + edge_type
            easy = easy * num_relation + r
        return ~torch.isin(key, easy)

    def _bellmanford(self, layers, boundary, query, edge_index, edge_type, num_node, relations=None, hook=None,
                     edge_weight=None):
        w = self._unit_weight(edge_index.shape[1], boundary) if edge_weight is None else edge_weight
        x = boundary
        for i, layer in enumerate(layers):
            residual = self.short_cut and layer.output_dim == x.shape[-1]
            x = layer(x, query, boundary, edge_index, edge_type, (num_node, num_node), w, shortcut=residual,
                      relation=None if relations is None else relations[i])
            if hook is not None:
                hook(i, x)
        return x

    def _merged_relation_graph(self, rel_edge_index, rel_edge_type):
        key = (rel_edge_index.data_ptr(), rel_edge_type.data_ptr(), tuple(rel_edge_index.shape))
        if key not in self._merged:  # once per graph; the tensors are kept so the addresses cannot be recycled
            num_rel = self.relation_model.layers[0].num_relation
            self._merged[key] = (merge_parallel_edges(rel_edge_index, rel_edge_type, num_rel), rel_edge_index,
                                 rel_edge_type)
        return self._merged[key][0]

    def relation_representations(self, rel_edge_index, rel_edge_type, num_relation, r_index):
        """RelNBFNet.forward (models.py:97-102): (B, num_relation, d) representations conditioned on the query relation."""
        ones = torch.ones(len(r_index), self.dim, device=r_index.device)
        boundary = self._boundary(r_index, ones, num_relation)
        layers = self.relation_model.layers
        merged = self._merged_relation_graph(rel_edge_index, rel_edge_type) if self.merge_relation_graph else None
        if merged is None:
            return self._bellmanford(layers, boundary, ones, rel_edge_index, rel_edge_type, num_relation)
        mei, met, combos = merged
        # per layer: the merged relation table (C, d) = combos @ embedding (4, d); one shared block for every head
        relations = [combos @ layer.relation.weight for layer in layers]
        return self._bellmanford(layers, boundary, ones, mei, met, num_relation, relations)

    def entity_states(self, edge_index, edge_type, num_node, rel_repr, h_index, r_index, hook=None, edge_mask=None):
        """EntityNBFNet.bellmanford (models.py:131-175) -> (h (B, N, d) last hidden, query (B, d)).
        edge_mask (bool (E,), optional): edges to KEEP this step (training: easy_edge_mask)."""
        layers = self.entity_model.layers
        query = rel_repr[torch.arange(len(r_index), device=r_index.device), r_index]
        projections = [layer.relation_projection for layer in layers]
        if fused.relation_project_ok(rel_repr, projections):
            relations = fused.relation_project(rel_repr, projections)  # (L, B, R, d), one launch
        else:
            relations = [p(rel_repr) for p in projections]
        boundary = self._boundary(h_index, query, num_node)
        weight = None
        if edge_index.is_cuda:
            # the incremental plan (ultra_rspmm_plan_apply_mask): re-cut on the device for this step's edge set, no sort,
            # no host sync; all six layers and their backward passes then run on the masked plan.  None restores the graph.
            from .plan import plan_cache
            plan = plan_cache.get(edge_index, edge_type, num_node, rel_repr.shape[1])
            if edge_mask is not None or plan.masked:
                plan.apply_edge_mask(edge_mask)
        elif edge_mask is not None:  # CPU tensors (tests with the oracle): a 0/1 weight is an edge removal for `sum`
            if layers[0].aggregate_func != "sum":
                raise ValueError("edge_mask as a 0/1 edge weight is only an edge removal for `sum`")
            weight = edge_mask.to(boundary.dtype)
        return self._bellmanford(layers, boundary, query, edge_index, edge_type, num_node, relations, hook, weight), query

    def score(self, h, query, t_index):
        """models.py:166-170,202-208 for explicit tails t_index (B, K): gather cat[h, query] rows, MLP -> (B, K)."""
        feature = torch.cat([h, query.unsqueeze(1).expand(-1, h.shape[1], -1)], dim=-1)
        feature = feature.gather(1, t_index.unsqueeze(-1).expand(-1, -1, feature.shape[-1]))
        return self.entity_model.mlp(feature).squeeze(-1)

    def score_all(self, h, query, chunk=32):
        """Scores of EVERY entity as the tail (the evaluation protocol, tasks.py:79-91: t_index = arange(N)) -> (B, N).
        mlp[0] on cat[h, q] is split into its h and q halves, so the (B, N, 2d) concatenation is never built."""
        mlp = self.entity_model.mlp
        if fused.score_head_ok(h, mlp):
            return fused.score_head(h, query, mlp)   # one pass over h, the hidden layer never reaches HBM
        if len(mlp) != 3:
            t_index = torch.arange(h.shape[1], device=h.device).expand(h.shape[0], -1)
            return self.score(h, query, t_index)
        d = h.shape[-1]
        wa, wb = mlp[0].weight[:, :d], mlp[0].weight[:, d:]
        qb = torch.addmm(mlp[0].bias, query, wb.t())                       # (B, 2d)
        out = h.new_empty(h.shape[0], h.shape[1])
        for lo in range(0, h.shape[0], chunk):                             # bounds the (chunk, N, 2d) intermediate
            hid = torch.matmul(h[lo:lo + chunk], wa.t())
            hid.add_(qb[lo:lo + chunk].unsqueeze(1)).relu_()
            out[lo:lo + chunk] = torch.matmul(hid, mlp[2].weight.t()).squeeze(-1) + mlp[2].bias
        return out

    # ---- Ultra.forward -----------------------------------------------------------------------------------------------
    def forward(self, rel_graph, graph, h_index, r_index, t_index=None, edge_mask=None):
        """rel_graph / graph: (edge_index, edge_type, num_node) of the relation graph / the entity graph (already in tail-
        prediction form: one (h, r) query per head).  Returns scores (B, K) for t_index (B, K), or (B, N) for all tails.
        edge_mask: bool (E,) keep-mask over the entity graph's edges for this step (training, see easy_edge_mask)."""
        rel_repr = self.relation_representations(rel_graph[0], rel_graph[1], rel_graph[2], r_index)
        h, query = self.entity_states(graph[0], graph[1], graph[2], rel_repr, h_index, r_index, edge_mask=edge_mask)
        return self.score_all(h, query) if t_index is None else self.score(h, query, t_index)


class CudaGraphRunner:
    """Capture `fn(*static_inputs)` once as a CUDA graph and replay it.

    The reference cannot do this: every rspmm call synchronises the host (sortedness assert, rspmm.py:15-17).  With cached
    plans the whole Ultra forward is stream-ordered.  Inference only; shapes, graphs (edge tensors) and weights are frozen
    at capture; per-call inputs are copied into the static buffers."""

    def __init__(self, fn, *inputs, warmup=2):
        self.fn = fn
        self.static_inputs = [t.clone() for t in inputs]
        dev = inputs[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():  # plan builds, weight binds, scratch allocation happen here
            for _ in range(warmup):
                self.fn(*self.static_inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = self.fn(*self.static_inputs)

    @torch.no_grad()
    def __call__(self, *inputs):
        for s, t in zip(self.static_inputs, inputs):
            s.copy_(t)
        self.graph.replay()
        return self.static_out
