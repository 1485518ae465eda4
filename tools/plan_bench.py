"""Plan build time (one-time per graph): python tools/plan_bench.py"""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultra_b200.plan import GraphPlan
from ultra_b200.synth import synth_scale
d = torch.device("cuda:0")
for name in ("wn18rr", "fb15k237", "codex_l"):
    ei, et, R, N = synth_scale(name, alpha=1.0, seed=1)
    ei, et = ei.to(d), et.to(d)
    GraphPlan(ei, et, N, R)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        p = GraphPlan(ei, et, N, R)
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
        del p
    # the reference's per-call work on the same graph: argsort + 3 gathers + sortedness check (rspmm.py:15-17,174-178)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        key = ei[0] * (ei[1].max() + 1) + ei[1]
        order = key.argsort()
        a, b = ei[:, order], et[order]
        k2 = a[0] * (a[1].max() + 1) + a[1]
        assert (k2.diff() >= 0).all()
    torch.cuda.synchronize()
    ref_ms = 1e3 * (time.perf_counter() - t0) / 5
    # incremental re-cut for a masked edge set (training's remove_easy_edges): device time (CUDA events: nothing syncs)
    # and host wall time of the enqueue
    p = GraphPlan(ei, et, N, R)
    keep = torch.rand(ei.shape[1], device=d) > 0.002
    for _ in range(3):
        p.apply_edge_mask(keep)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(10):
        p.apply_edge_mask(keep)
    b.record()
    enqueue_ms = 1e3 * (time.perf_counter() - t0) / 10
    torch.cuda.synchronize()
    mask_ms = a.elapsed_time(b) / 10
    print(json.dumps({"graph": name, "E": ei.shape[1], "plan_build_ms_min": round(min(ts), 3), "plan_build_ms_med": round(sorted(ts)[2], 3),
                      "apply_mask_device_ms": round(mask_ms, 3), "apply_mask_enqueue_ms": round(enqueue_ms, 3),
                      "reference_per_call_sort_ms": round(ref_ms, 3)}))
