"""Developer micro-benchmark: rspmm forward/backward kernel time at a named scale (CUDA events, plan pre-built)."""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultra_b200 import _capi  # noqa: E402
from ultra_b200.plan import GraphPlan, _DTYPE_ID, _stream_ptr  # noqa: E402
from ultra_b200.synth import synth_scale, synth_kg, algorithmic_bytes, compulsory_bytes  # noqa: E402


def time_fn(fn, warmup=3, iters=10):
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
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", default="fb15k237")
    ap.add_argument("--heads", type=int, default=128)
    ap.add_argument("--alphas", default="1.0,0.0")
    ap.add_argument("--ops", default="add_mul")
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--peak", type=float, default=6583.2)
    ap.add_argument("--l2-percent", type=int, default=0)
    ap.add_argument("--sm-limit", type=int, default=0)
    ap.add_argument("--tunable", action="append", default=[], help="name=value for ultra_rspmm_set_tunable (repeatable)")
    ap.add_argument("--work-queue", type=int, default=-1, help="experiment: 1 ticket work queue (default), 0 static round-robin")
    ap.add_argument("--smem-tiles", type=int, default=-1, help="experiment: highest smem residency mode (2 / 1 / 0)")
    ap.add_argument("--mask-frac", type=float, default=0.0,
                    help="bind a 0/1 edge weight that removes this fraction of the edges (the weighted kernel path)")
    args = ap.parse_args()
    d = torch.device("cuda:0")
    lib = _capi.lib()
    for kv in args.tunable:
        k, v = kv.split("=")
        _capi.set_tunable(k, int(v))
    if args.work_queue >= 0:
        _capi.set_tunable("work_queue", args.work_queue)
    if args.smem_tiles >= 0:
        _capi.set_tunable("smem_tiles", args.smem_tiles)
    if args.l2_percent:
        _capi.set_tunable("l2_percent", args.l2_percent)
    for alpha in [float(a) for a in args.alphas.split(",")]:
        ei, et, R, N = synth_scale(args.scale, alpha=alpha, seed=1234)
        E = ei.shape[1]
        D = args.heads * 64
        g = torch.Generator(device=d).manual_seed(1)
        rel = torch.randn(R, D, device=d, generator=g)
        x = torch.randn(N, D, device=d, generator=g)
        out = torch.empty_like(x)
        plan = GraphPlan(ei.to(d), et.to(d), N, R)
        if args.sm_limit:
            plan.set_sm_limit(args.sm_limit)
        if args.mask_frac > 0:
            keep = torch.rand(E, device=d, generator=g) >= args.mask_frac
            wmask = plan.edge_mask_weight(keep)
            med, best = time_fn(lambda: (plan.invalidate_weight(), plan.bind_weight(wmask)))
            print(json.dumps({"weight_bind_ms_med": round(med, 4), "removed_edges": int((~keep).sum())}))
        scratch = plan.scratch(D, torch.float32)
        deg = torch.bincount(ei[0], minlength=N)
        print(json.dumps({"alpha": alpha, "N": N, "E": E, "R": R, "D": D, "max_deg": int(deg.max()),
                          "items": plan.info(3), "slots": [plan.info(4), plan.info(5), plan.info(6)],
                          "scratch_MB": scratch.numel() / 1e6}))
        for op in args.ops.split(","):
            s, m = op.split("_")
            sid, mid = _capi.SUM_IDS[s], _capi.MUL_IDS[m]

            def fwd():
                rc = lib.ultra_rspmm_forward(plan.handle, sid, mid, 0, rel.data_ptr(), x.data_ptr(), out.data_ptr(), D,
                                             scratch.data_ptr(), scratch.numel(), _stream_ptr(d))
                assert rc == 0, lib.ultra_rspmm_last_error()
            med, best = time_fn(fwd)
            ba = algorithmic_bytes(N, E, R, D)
            print(json.dumps({"op": op, "pass": "fwd", "alpha": alpha, "ms_med": round(med, 4), "ms_min": round(best, 4),
                              "Medges_s": round(E / med / 1e3, 1), "alg_GBs": round(ba / med / 1e6, 1),
                              "frac_of_peak": round(ba / med / 1e6 / args.peak, 3),
                              "min_GBs": round(compulsory_bytes(N, E, R, D) / med / 1e6, 1)}))
            if args.bwd:
                og = torch.randn_like(x)
                ig, rg = torch.empty_like(x), torch.empty_like(rel)

                def bwd():
                    rc = lib.ultra_rspmm_backward(plan.handle, sid, mid, 0, rel.data_ptr(), x.data_ptr(), out.data_ptr(),
                                                  og.data_ptr(), None, rg.data_ptr(), ig.data_ptr(), D,
                                                  scratch.data_ptr(), scratch.numel(), _stream_ptr(d))
                    assert rc == 0, lib.ultra_rspmm_last_error()
                for fused in (0, 1):
                    _capi.set_tunable("bwd_fused_relgrad", fused)
                    med, best = time_fn(bwd, iters=5)
                    print(json.dumps({"op": op, "pass": "bwd", "fused_relgrad": fused, "alpha": alpha,
                                      "ms_med": round(med, 4), "ms_min": round(best, 4)}))
                _capi.set_tunable("bwd_fused_relgrad", 0)
        del plan


if __name__ == "__main__":
    main()
