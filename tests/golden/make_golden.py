"""Generates tests/golden/rspmm_golden.npz by running the UNMODIFIED reference CPU kernels (oracle/_ref, compiled
from /root/reference/ultra/rspmm/source/rspmm.cpp by oracle/build_ref.py) on small seeded inputs.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The reference ships no golden vectors of its own; these fixtures pin the oracle and the CUDA path
to the reference's outputs.  Cases: a random multigraph with duplicate edges (min/max ties), isolated nodes
(empty rows), a hub row longer than one work item, non-unit weights, fp32 and fp64, plus the hand-computed case
of SURVEY.md section 8c.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref, sort_edges  # noqa: E402

OPS = [(s, m) for s in ("add", "min", "max") for m in ("mul", "add")]


def make_case(name, seed, N, E, R, D, dtype, hub=0, unit_weight=False, quantize=False):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, N, (2, E), generator=g)
    if hub:  # node 1 aggregates `hub` edges; nodes N-2, N-1 stay isolated (empty rows)
        ei[0, :hub] = 1
    ei = ei.clamp(max=N - 3)
    et = torch.randint(0, R, (E,), generator=g)
    if quantize:  # small integer values: many exact ties for min/max, exact sums in any order
        rel = torch.randint(-2, 3, (R, D), generator=g).to(dtype)
        x = torch.randint(-2, 3, (N, D), generator=g).to(dtype)
        w = torch.randint(1, 3, (E,), generator=g).to(dtype)
    else:
        rel = torch.randn(R, D, generator=g, dtype=dtype)
        x = torch.randn(N, D, generator=g, dtype=dtype)
        w = torch.rand(E, generator=g, dtype=dtype) + 0.5
    if unit_weight:
        w = torch.ones(E, dtype=dtype)
    og = torch.randn(N, D, generator=g, dtype=dtype)
    sei, stt, sw, order = sort_edges(ei.numpy(), et.numpy(), w.numpy())
    out = {name + "/edge_index": ei.numpy(), name + "/edge_type": et.numpy(), name + "/edge_weight": w.numpy(),
           name + "/relation": rel.numpy(), name + "/input": x.numpy(), name + "/output_grad": og.numpy()}
    tei, ttt, tw = torch.from_numpy(sei), torch.from_numpy(stt), torch.from_numpy(sw)
    for s, m in OPS:
        o = ref.ref_forward(s, m, tei, ttt, tw, rel, x)
        wg, rg, ig = ref.ref_backward(s, m, tei, ttt, tw, rel, x, o, og)
        wg_caller = torch.empty_like(wg)
        wg_caller[torch.from_numpy(order)] = wg  # back to the caller's (unsorted) edge order
        k = "%s/%s_%s/" % (name, s, m)
        out[k + "output"] = o.numpy()
        out[k + "weight_grad"] = wg_caller.numpy()
        out[k + "relation_grad"] = rg.numpy()
        out[k + "input_grad"] = ig.numpy()
    return out


def main():
    torch.set_num_threads(1)  # deterministic accumulation order in the reference's mutex-guarded backward
    data = {}
    data.update(make_case("rand_f32", 1, N=37, E=300, R=5, D=32, dtype=torch.float32))
    data.update(make_case("rand_f64", 2, N=37, E=300, R=5, D=16, dtype=torch.float64))
    data.update(make_case("hub_f32", 3, N=50, E=600, R=7, D=64, dtype=torch.float32, hub=200))
    data.update(make_case("ties_f32", 4, N=20, E=400, R=3, D=32, dtype=torch.float32, quantize=True))
    data.update(make_case("oddD_f32", 5, N=23, E=150, R=4, D=7, dtype=torch.float32))
    data.update(make_case("unitw_f32", 6, N=64, E=500, R=6, D=96, dtype=torch.float32, unit_weight=True))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rspmm_golden.npz"), **data)
    print("wrote", len(data), "arrays")


if __name__ == "__main__":
    main()
