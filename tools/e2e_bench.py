"""e2e (host buffers) timing for a few slab widths: python tools/e2e_bench.py"""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ultra_b200.rspmm import rspmm as R
wl = bench.make_workload(bench.HEADS, "cpu")
rel_h, x_h = wl["rel"].pin_memory(), wl["x"].pin_memory()
from ultra_b200 import _capi
for slab, split in ((512, 1), (512, 2), (512, 4), (1024, 2), (1024, 4), (2048, 4)):
    R._HOST_SLAB_DIM = slab
    _capi.set_tunable("host_copy_split", split)
    for _ in range(2):
        R.generalized_rspmm(wl["ei"], wl["et"], wl["w"], rel_h, x_h)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        out = R.generalized_rspmm(wl["ei"], wl["et"], wl["w"], rel_h, x_h)
        del out
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(json.dumps({"slab": slab, "split": split, "ms": round(dt * 1e3, 3), "Medges_s": round(wl["E"] / dt / 1e6, 1)}))
