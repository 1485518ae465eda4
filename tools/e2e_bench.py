"""e2e (host buffers) study: the PCIe ceilings on this box next to the public-API call, for a few slab widths.
    python tools/e2e_bench.py"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ultra_b200 import _capi  # noqa: E402
from ultra_b200.rspmm import rspmm as R  # noqa: E402

d = torch.device("cuda:0")
wl = bench.make_workload(bench.HEADS)
rel_h, x_h = wl["rel"].pin_memory(), wl["x"].pin_memory()
out_h = torch.empty_like(x_h).pin_memory()
x_d, o_d = torch.empty_like(x_h, device=d), torch.randn(x_h.shape, device=d)
nbytes = x_h.numel() * 4


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        x_d.copy_(x_h, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        out_h.copy_(o_d, non_blocking=True)


def both():
    h2d()
    d2h()


def slabs(width):
    """both directions concurrently, column slabs of `width` columns (the strided pattern of forward_host)"""
    D = x_h.shape[1]
    for c0 in range(0, D, width):
        with torch.cuda.stream(s1):
            x_d[:, c0:c0 + width].copy_(x_h[:, c0:c0 + width], non_blocking=True)
        with torch.cuda.stream(s2):
            out_h[:, c0:c0 + width].copy_(o_d[:, c0:c0 + width], non_blocking=True)


for name, fn, total in (("H2D contiguous", h2d, nbytes), ("D2H contiguous", d2h, nbytes), ("H2D + D2H concurrently", both, 2 * nbytes),
                        ("both, 512-column slabs", lambda: slabs(512), 2 * nbytes),
                        ("both, 1024-column slabs", lambda: slabs(1024), 2 * nbytes),
                        ("both, 2048-column slabs", lambda: slabs(2048), 2 * nbytes)):
    dt = timed(fn)
    print(json.dumps({"pcie": name, "ms": round(dt * 1e3, 3), "GBs": round(total / dt / 1e9, 1)}), flush=True)
del x_d, o_d
for slab, split in ((256, 1), (512, 1), (1024, 1), (1024, 2), (2048, 1)):
    R._HOST_SLAB_DIM = slab
    _capi.set_tunable("host_copy_split", split)
    dt = timed(lambda: R.generalized_rspmm(wl["ei"], wl["et"], wl["w"], rel_h, x_h))
    print(json.dumps({"e2e_call": True, "slab": slab, "split": split, "ms": round(dt * 1e3, 3),
                      "Medges_s": round(wl["E"] / dt / 1e6, 1)}), flush=True)
