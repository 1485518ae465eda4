"""Measure rspmm forward/backward at every BASELINE.json config (C1..C5), single GPU, plan pre-built.
Writes one JSON line per point (to stdout and --out).  Developer / profiles tool, not the bench contract."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultra_b200 import _capi  # noqa: E402
from ultra_b200.plan import GraphPlan, _stream_ptr  # noqa: E402
from ultra_b200.synth import (SCALES, synth_kg_torch, relation_graph_of, algorithmic_bytes,  # noqa: E402
                              compulsory_bytes)


def time_fn(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def parity_first_head(ei, et, rel, x, out, cols=64, chunk=4_000_000):
    """Same-run correctness of the timed forward: the first `cols` columns against the dense PyTorch formula
    out[i] = sum_{e: row_e = i} rel[type_e] * x[col_e] accumulated in fp64 on the GPU,
    north-star tolerance 1e-5 relative to max|ref| (+ the fp32 reference's own rounding on very long rows is not an
    issue here because the reference side is fp64)."""
    N = x.shape[0]
    ref = torch.zeros(N, cols, dtype=torch.float64, device=x.device)
    r64, x64 = rel[:, :cols].double(), x[:, :cols].double()
    for lo in range(0, ei.shape[1], chunk):
        sl = slice(lo, lo + chunk)
        ref.index_add_(0, ei[0, sl], r64[et[sl]] * x64[ei[1, sl]])
    scale = float(ref.abs().max())
    err = float((out[:, :cols].double() - ref).abs().max())
    return {"cols": cols, "max_abs_err": err, "max_abs_ref": scale, "rel_to_max": err / scale if scale else 0.0,
            "ok": bool(err <= 1e-5 * scale)}


def measure(name, ei, et, N, R, D, peak, bwd=True, extra=None):
    d = ei.device
    lib = _capi.lib()
    E = ei.shape[1]
    g = torch.Generator(device=d).manual_seed(1)
    rel = torch.randn(R, D, device=d, generator=g)
    x = torch.randn(N, D, device=d, generator=g)
    out = torch.empty_like(x)
    plan = GraphPlan(ei, et, N, R)
    scratch = plan.scratch(D, torch.float32)

    def fwd():
        rc = lib.ultra_rspmm_forward(plan.handle, 0, 0, 0, rel.data_ptr(), x.data_ptr(), out.data_ptr(), D,
                                     scratch.data_ptr(), scratch.numel(), _stream_ptr(d))
        assert rc == 0, lib.ultra_rspmm_last_error()
    ms = time_fn(fwd)
    parity = parity_first_head(ei, et, rel, x, out)
    ba, bm = algorithmic_bytes(N, E, R, D), compulsory_bytes(N, E, R, D)
    rec = {"config": name, "N": N, "E": E, "R": R, "D": D, "fwd_ms": round(ms, 4),
           "Medges_s": round(E / ms / 1e3, 1), "alg_GBs": round(ba / ms / 1e6, 1),
           "frac_alg": round(ba / ms / 1e6 / peak, 3), "frac_compulsory": round(bm / ms / 1e6 / peak, 3),
           "max_deg": int(torch.bincount(ei[0], minlength=N).max()), "parity": parity}
    if bwd:
        og = torch.randn_like(x)
        ig, rg = torch.empty_like(x), torch.empty_like(rel)

        def b():
            rc = lib.ultra_rspmm_backward(plan.handle, 0, 0, 0, rel.data_ptr(), x.data_ptr(), out.data_ptr(),
                                          og.data_ptr(), None, rg.data_ptr(), ig.data_ptr(), D, scratch.data_ptr(),
                                          scratch.numel(), _stream_ptr(d))
            assert rc == 0, lib.ultra_rspmm_last_error()
        rec["bwd_ms"] = round(time_fn(b, iters=3), 4)
    if extra:
        rec.update(extra)
    print(json.dumps(rec), flush=True)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/r02_sweep_out.jsonl")
    ap.add_argument("--peak", type=float, default=6583.2)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    d = torch.device("cuda:0")
    recs = []
    for alpha in (1.0, 0.0):
        tag = "zipf1" if alpha else "uniform"
        n, r, e = SCALES["wn18rr"]
        ei, et, R = synth_kg_torch(n, r, e, alpha, 1235, d)
        recs.append(measure("C1 wn18rr B=4 " + tag, ei, et, n, R, 4 * 64, args.peak))
        n, r, e = SCALES["fb15k237"]
        ei, et, R = synth_kg_torch(n, r, e, alpha, 1236, d)
        recs.append(measure("C2 fb15k237 B=128 " + tag, ei, et, n, R, 128 * 64, args.peak))
        recs.append(measure("C3 fb15k237 entity B=256 " + tag, ei, et, n, R, 256 * 64, args.peak, bwd=False))
        rei, ret, RR, RN = relation_graph_of(ei, et, n, R)
        recs.append(measure("C3 relation-graph B=256 " + tag, rei, ret, RN, RR, 256 * 64, args.peak))
        n, r, e = SCALES["codex_l"]
        ei, et, R = synth_kg_torch(n, r, e, alpha, 1238, d)
        recs.append(measure("C4 codex_l per-GPU B=8 " + tag, ei, et, n, R, 8 * 64, args.peak))
        recs.append(measure("C4 codex_l 1-GPU B=64 " + tag, ei, et, n, R, 64 * 64, args.peak))
    sizes = [(10_000, 5), (10_000, 50), (100_000, 20), (1_000_000, 5)] if args.quick else \
        [(n, k) for n in (10_000, 100_000, 1_000_000) for k in (5, 20, 50)]
    for alpha in (1.0, 0.0):
        tag = "zipf1" if alpha else "uniform"
        for n, k in sizes:
            ei, et, R = synth_kg_torch(n, 128, n * k // 2, alpha, 1239, d)
            recs.append(measure("C5 N=%d E/N=%d B=32 %s" % (n, k, tag), ei, et, n, R, 32 * 64, args.peak,
                                bwd=(n * k <= 5_000_000)))
            del ei, et
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        for r in recs:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
