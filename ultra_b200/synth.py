"""Synthetic knowledge graphs at the scale of BASELINE.json's configs (no datasets / network in this environment).

Follows SURVEY.md section 8d: heads/tails drawn from a Zipf(alpha)-distributed, randomly permuted node popularity,
relation ids from Zipf(alpha) over the base relations, triples deduplicated, inverse edges (t, h, r + R_base)
appended exactly like the reference datasets do (/root/reference/ultra/datasets.py:112-115,215-216,318-319), so
E = 2 * E_base and R = 2 * R_base.  alpha = 0 gives the uniform-random variant.
"""
import numpy as np
import torch

# (num_node, base relations, base triples) of the public datasets named in BASELINE.json configs
SCALES = {
    "wn18rr": (40943, 11, 86835),        # C1
    "fb15k237": (14541, 237, 272115),    # C2 / C3
    "codex_l": (77951, 69, 551193),      # C4
}


def _zipf_sampler(n, alpha, rng):
    if alpha == 0.0:
        return lambda size: rng.integers(0, n, size=size, dtype=np.int64)
    p = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), alpha)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    perm = rng.permutation(n).astype(np.int64)  # hubs are not the low ids
    return lambda size: perm[np.minimum(np.searchsorted(cdf, rng.random(size), side="right"), n - 1)]


def synth_kg(num_node, num_rel_base, num_triples, alpha=1.0, seed=0, inverse=True):
    """Returns (edge_index (2,E) int64, edge_type (E) int64, num_relation) as CPU torch tensors, unsorted."""
    rng = np.random.default_rng(seed)
    node = _zipf_sampler(num_node, alpha, rng)
    rel = _zipf_sampler(num_rel_base, alpha, rng)
    keys = np.empty(0, dtype=np.int64)
    cap = num_node * num_node * num_rel_base
    target = min(num_triples, cap)
    while keys.size < target:
        m = int((target - keys.size) * 1.5) + 1024
        h, t, r = node(m), node(m), rel(m)
        k = (h * num_rel_base + r) * num_node + t
        keys = np.unique(np.concatenate([keys, k]))
        if keys.size > target:  # drop a random surplus, not the largest keys
            keys = keys[np.sort(rng.choice(keys.size, size=target, replace=False))]
    keys = keys[rng.permutation(keys.size)]
    t = keys % num_node
    hr = keys // num_node
    r = hr % num_rel_base
    h = hr // num_rel_base
    if inverse:
        src = np.concatenate([h, t])
        dst = np.concatenate([t, h])
        typ = np.concatenate([r, r + num_rel_base])
        num_relation = 2 * num_rel_base
    else:
        src, dst, typ, num_relation = h, t, r, num_rel_base
    edge_index = torch.from_numpy(np.stack([src, dst]).astype(np.int64))
    edge_type = torch.from_numpy(typ.astype(np.int64))
    return edge_index, edge_type, num_relation


def synth_scale(name, alpha=1.0, seed=0):
    n, r, e = SCALES[name]
    return synth_kg(n, r, e, alpha=alpha, seed=seed) + (n,)


def algorithmic_bytes(num_node, num_edge, num_relation, dim, elem=4):
    """Edge-gather model of SURVEY.md section 8d / BASELINE.md: every edge moves one D-wide input row slice into an SM;
    relation, CSR and output move once."""
    return elem * dim * (num_edge + num_node + num_relation) + 12 * num_edge + 4 * (num_node + 1)


def compulsory_bytes(num_node, num_edge, num_relation, dim, elem=4):
    """Compulsory-traffic floor: input read once, output written once."""
    return elem * dim * (2 * num_node + num_relation) + 12 * num_edge + 4 * (num_node + 1)


def relation_graph_of(edge_index, edge_type, num_node, num_relation):
    """Torch-only restatement of build_relation_graph (/root/reference/ultra/tasks.py:144-199): relations become
    nodes; 4 edge types (0 h2h, 1 t2t, 2 h2t, 3 t2h) connect relations that share a head/tail entity.  Only the
    sparsity pattern of the four products is used by the reference, so dense boolean products suffice (R <= ~1.5k).
    Returns (edge_index (2,E'), edge_type (E'), num_relation=4, num_node=R)."""
    dev = edge_index.device
    Eh = torch.zeros(num_node, num_relation, device=dev)
    Et = torch.zeros(num_node, num_relation, device=dev)
    Eh[edge_index[0], edge_type] = 1.0
    Et[edge_index[1], edge_type] = 1.0
    prods = [Eh.t() @ Eh, Et.t() @ Et, Eh.t() @ Et, Et.t() @ Eh]  # hh, tt, ht, th  (tasks.py:181-184)
    eis, ets = [], []
    for k, a in enumerate(prods):
        idx = (a > 0).nonzero().t()
        eis.append(idx)
        ets.append(torch.full((idx.shape[1],), k, dtype=torch.long, device=dev))
    return torch.cat(eis, dim=1), torch.cat(ets), 4, num_relation


def synth_kg_torch(num_node, num_rel_base, num_triples, alpha=1.0, seed=0, device="cuda"):
    """GPU generator for the large sweep graphs (same recipe as synth_kg, torch ops, not bit-identical to it)."""
    g = torch.Generator(device=device).manual_seed(seed)

    def sampler(n):
        if alpha == 0.0:
            return lambda m: torch.randint(0, n, (m,), generator=g, device=device)
        p = 1.0 / torch.arange(1, n + 1, dtype=torch.float64, device=device).pow(alpha)
        cdf = torch.cumsum(p, 0)
        cdf = cdf / cdf[-1]
        perm = torch.randperm(n, generator=g, device=device)
        return lambda m: perm[torch.searchsorted(cdf, torch.rand(m, generator=g, device=device, dtype=torch.float64)
                                                 ).clamp_(max=n - 1)]
    node, rel = sampler(num_node), sampler(num_rel_base)
    keys = torch.empty(0, dtype=torch.long, device=device)
    target = min(num_triples, num_node * num_node * num_rel_base)
    while keys.numel() < target:
        m = int((target - keys.numel()) * 1.5) + 1024
        k = (node(m) * num_rel_base + rel(m)) * num_node + node(m)
        keys = torch.unique(torch.cat([keys, k]))
        if keys.numel() > target:
            keys = keys[torch.randperm(keys.numel(), generator=g, device=device)[:target]]
    keys = keys[torch.randperm(keys.numel(), generator=g, device=device)]
    t = keys % num_node
    hr = keys // num_node
    r = hr % num_rel_base
    h = hr // num_rel_base
    edge_index = torch.stack([torch.cat([h, t]), torch.cat([t, h])])
    edge_type = torch.cat([r, r + num_rel_base])
    return edge_index, edge_type, 2 * num_rel_base
