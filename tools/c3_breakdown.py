"""Kernel-level breakdown of one full Ultra forward (bench.py's strong-scaling configs, one GPU, eager launches).

Run under ``ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file X
python tools/c3_breakdown.py C3`` and summarise X with ``python tools/c3_breakdown.py --summarise X`` (per-kernel launch
count and summed duration of the second forward, the NVTX range "timed"; the first one builds the plans).  ncu serialises launches and flushes caches between them:
shares, not absolute times (profiles/README.md).
"""
import csv
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(name):
    import torch
    import bench
    from ultra_b200.nbfnet import UltraNBFNet
    from ultra_b200.synth import synth_scale, relation_graph_of
    device = torch.device("cuda:0")
    cfg = bench.STRONG[name]
    B = cfg["batch"]
    ei, et, R, N = synth_scale(cfg["scale"], alpha=bench.ALPHA, seed=cfg["seed"])
    ei, et = ei.to(device), et.to(device)
    rei, ret, rR, rN = relation_graph_of(ei, et, N, R)
    torch.manual_seed(11)
    model = UltraNBFNet().to(device).eval()
    g = torch.Generator().manual_seed(cfg["seed"])
    h = torch.randint(0, N, (B,), generator=g).to(device)
    r = torch.randint(0, R, (B,), generator=g).to(device)
    with torch.no_grad():
        for it in range(2):
            torch.cuda.synchronize()
            if it == 1:
                torch.cuda.nvtx.range_push("timed")
            model((rei.contiguous(), ret.contiguous(), rN), (ei, et, N), h, r)
            torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()


def summarise(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    ix = {k: i for i, k in enumerate(rows[0])}
    recs = []
    for r in rows[1:]:
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3}.get(u, 1e-6)
        recs.append((r[ix["Kernel Name"]], r[ix["Grid Size"]], v))
    tot, cnt = collections.Counter(), collections.Counter()
    for n, gsz, v in recs:
        key = re.sub(r"\(.*", "", n)[:90]
        tot[key] += v
        cnt[key] += 1
    total = sum(tot.values())
    print("launches %d   summed duration %.3f ms (serialised, cold caches)" % (len(recs), total))
    for k, v in tot.most_common(30):
        print("%-92s %5d %9.3f ms %5.1f %%" % (k, cnt[k], v, 100 * v / total))


if __name__ == "__main__":
    if sys.argv[1] == "--summarise":
        summarise(sys.argv[2])
    else:
        run(sys.argv[1])
