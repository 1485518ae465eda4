"""e2e (host buffers) timing for a few slab widths: python tools/e2e_bench.py"""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ultra_b200.rspmm import rspmm as R
wl = bench.make_workload(bench.HEADS, "cpu")
rel_h, x_h = wl["rel"].pin_memory(), wl["x"].pin_memory()
for slab in (256, 512, 1024, 2048, 4096):
    R._HOST_SLAB_DIM = slab
    for _ in range(2):
        R.generalized_rspmm(wl["ei"], wl["et"], wl["w"], rel_h, x_h)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        out = R.generalized_rspmm(wl["ei"], wl["et"], wl["w"], rel_h, x_h)
        del out
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(json.dumps({"slab": slab, "ms": round(dt * 1e3, 3), "Medges_s": round(wl["E"] / dt / 1e6, 1)}))
