"""One fused NBFNet layer at C2 (for ncu captures of the head-major rspmm + update kernel)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = bench.make_workload(bench.HEADS)
d = torch.device("cuda:0")
print(bench.bench_layer(wl, d, 1, 0, bench.make_timer(d, 1, None), bench.load_counters()))
