"""Launch-bound regime: 6-layer relation-graph stack, eager vs one CUDA graph."""
import os, sys, json, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultra_b200.layer import GeneralizedRelationalConv, bellmanford_entity, GraphedBellmanFord
from ultra_b200.synth import SCALES, synth_kg_torch, relation_graph_of
d = torch.device("cuda:0")
n, r, e = SCALES["fb15k237"]
ei, et, R = synth_kg_torch(n, r, e, 1.0, 1236, d)
rei, ret, RR, RN = relation_graph_of(ei, et, n, R)
torch.manual_seed(0)
layers = [GeneralizedRelationalConv(64, 64, RR, 64, "distmult", "sum", layer_norm=True).to(d) for _ in range(6)]
for B in (4, 16, 64):
    q, h = torch.ones(B, 64, device=d), torch.randint(0, RN, (B,), device=d)
    graphed = GraphedBellmanFord(layers, q, h, rei, ret, RN)
    def timeit(fn, reps=20):
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / reps
    with torch.no_grad():
        t_eager = timeit(lambda: bellmanford_entity(layers, q, h, rei, ret, RN))
    t_graph = timeit(lambda: graphed(q, h))
    print(json.dumps({"stack": "relation graph, 6 layers", "B": B, "nodes": RN, "edges": rei.shape[1], "eager_ms": round(t_eager, 3),
                      "graph_ms": round(t_graph, 3), "speedup": round(t_eager / t_graph, 2)}))
