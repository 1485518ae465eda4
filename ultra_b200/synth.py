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
