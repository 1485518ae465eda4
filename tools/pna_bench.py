"""PNA aggregation at C2: four rspmm calls (reference layer) vs the one-pass kernel."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.kbench import time_fn
from ultra_b200.rspmm import generalized_rspmm, generalized_rspmm_pna
from ultra_b200.synth import synth_scale
d = torch.device("cuda:0")
ei, et, R, N = synth_scale("fb15k237", alpha=1.0, seed=1236)
ei, et = ei.to(d), et.to(d)
for heads in (32, 128):
    D = heads * 64
    rel, x = torch.randn(R, D, device=d), torch.randn(N, D, device=d)
    w = torch.ones(ei.shape[1], device=d)
    with torch.no_grad():
        four = lambda: (generalized_rspmm(ei, et, w, rel, x, sum="add"), generalized_rspmm(ei, et, w, rel ** 2, x ** 2, sum="add"),
                        generalized_rspmm(ei, et, w, rel, x, sum="max"), generalized_rspmm(ei, et, w, rel, x, sum="min"))
        one = lambda: generalized_rspmm_pna(ei, et, w, rel, x)
        t4, _ = time_fn(four, iters=5)
        t1, _ = time_fn(one, iters=5)
    print(json.dumps({"heads": heads, "four_calls_ms": round(t4, 3), "one_pass_ms": round(t1, 3), "speedup": round(t4 / t1, 2)}))
