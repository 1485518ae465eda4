"""A/B of compile-time kernel variants (tools/build_variants.py) against the default build, on the GPU.

Parent: ``python tools/ab_bench.py [--out FILE]`` runs one child per library (``ULTRA_RSPMM_LIB``), prints one JSON line
per (library, case) and a final ``{"winner": ...}`` line: the library with the lowest summed forward time over the cases
whose outputs are bit-identical to the default build's (same arithmetic order, so they must be).
Child: ``--child`` times C2 (FB15k-237 scale, 128 heads) forward on the Zipf and the uniform graph, one backward, and a
max x add forward at 32 heads, and the fused layer epilogue at C2 size, and prints an int64 checksum of every result.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def checksum(t):
    import torch
    v = t.contiguous().view(torch.int32).to(torch.int64)
    w = torch.arange(v.numel(), device=v.device, dtype=torch.int64).view_as(v) % 1000003 + 1
    return int((v * w).sum().item())  # int64 wrap-around is fine: order-independent, position-sensitive


def child():
    import torch
    from tools.kbench import time_fn
    from ultra_b200 import _capi
    from ultra_b200.plan import GraphPlan, _stream_ptr
    from ultra_b200.synth import synth_scale, relation_graph_of

    d = torch.device("cuda:0")
    lib = _capi.lib()
    for kv in filter(None, os.environ.get("ULTRA_AB_TUNABLES", "").split(",")):  # e.g. plan_relation_runs=1
        k, v = kv.split("=")
        assert lib.ultra_rspmm_set_tunable(k.encode(), int(v)) == 0, lib.ultra_rspmm_last_error()
    # C2 entity graph (Zipf, uniform), a min/max-type op at 32 heads, and the relation graph of the C2 KG (474 nodes,
    # 4 relation rows, ~0.9 M edges: the small-graph kernel variant with the gathered tile in shared memory)
    for alpha, heads, op, bwd, relgraph in ((1.0, 128, "add_mul", True, False), (0.0, 128, "add_mul", False, False),
                                            (1.0, 32, "max_add", True, False), (1.0, 64, "add_mul", False, True)):
        ei, et, R, N = synth_scale("fb15k237", alpha=alpha, seed=1234)
        if relgraph:
            ei, et, R, N = relation_graph_of(ei, et, N, R)
        E, D = ei.shape[1], heads * 64
        g = torch.Generator(device=d).manual_seed(1)
        rel = torch.randn(R, D, device=d, generator=g)
        x = torch.randn(N, D, device=d, generator=g)
        og = torch.randn(N, D, device=d, generator=g)
        out = torch.empty_like(x)
        plan = GraphPlan(ei.to(d), et.to(d), N, R)
        scratch = plan.scratch(D, torch.float32)
        s, m = op.split("_")
        sid, mid = _capi.SUM_IDS[s], _capi.MUL_IDS[m]

        def fwd():
            rc = lib.ultra_rspmm_forward(plan.handle, sid, mid, 0, rel.data_ptr(), x.data_ptr(), out.data_ptr(), D,
                                         scratch.data_ptr(), scratch.numel(), _stream_ptr(d))
            assert rc == 0, lib.ultra_rspmm_last_error()
        med, best = time_fn(fwd, warmup=3, iters=15)
        rec = {"alpha": alpha, "heads": heads, "op": op, "relation_graph": relgraph, "N": N, "E": E,
               "fwd_ms": round(med, 4), "fwd_ms_min": round(best, 4),
               "fwd_sum": checksum(out), "fwd_l1": float(out.double().abs().sum().item())}
        if bwd:
            ig, rg = torch.empty_like(x), torch.empty_like(rel)

            def bw():
                rc = lib.ultra_rspmm_backward(plan.handle, sid, mid, 0, rel.data_ptr(), x.data_ptr(), out.data_ptr(),
                                              og.data_ptr(), None, rg.data_ptr(), ig.data_ptr(), D,
                                              scratch.data_ptr(), scratch.numel(), _stream_ptr(d))
                assert rc == 0, lib.ultra_rspmm_last_error()
            med, best = time_fn(bw, warmup=2, iters=7)
            rec.update({"bwd_ms": round(med, 4), "ig_sum": checksum(ig), "rg_sum": checksum(rg),
                        "ig_l1": float(ig.double().abs().sum().item()), "rg_l1": float(rg.double().abs().sum().item())})
        print(json.dumps(rec), flush=True)
        del plan
    # fused layer epilogue at C2 size (library default implementation of ultra_nbfnet_update_f32)
    from ultra_b200.fused import nbfnet_update
    torch.manual_seed(3)
    lin = torch.nn.Linear(128, 64).to(d)
    ln = torch.nn.LayerNorm(64).to(d)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.uniform_(-0.5, 0.5)
        M = 128 * 14541
        x = torch.randn(M, 64, device=d)
        u = torch.randn(M, 64, device=d) * 3
        got = nbfnet_update(x, u, lin, ln, relu=True, shortcut=True)
        med, best = time_fn(lambda: nbfnet_update(x, u, lin, ln, relu=True, shortcut=True), warmup=3, iters=15)
    print(json.dumps({"case": "update", "upd_ms": round(med, 4), "upd_ms_min": round(best, 4),
                      "upd_sum": checksum(got)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--timeout", type=int, default=150)
    ap.add_argument("--tunable-sets", default="", help="run the DEFAULT library again under these tunables, e.g. "
                    "'pf0:l2_prefetch_tiles=0 pf8:l2_prefetch_tiles=8' (name:k=v[,k=v])")
    args = ap.parse_args()
    if args.child:
        return child()
    vdir = os.path.join(ROOT, "ultra_b200", "lib", "variants")
    libs = {"default": os.path.join(ROOT, "ultra_b200", "lib", "libultra_rspmm.so")}
    if os.path.isdir(vdir):
        for f in sorted(os.listdir(vdir)):
            if f.endswith(".so"):
                libs[f[len("libultra_rspmm_"):-3]] = os.path.join(vdir, f)
    # the same libraries again with plan-time tunables (results are then NOT bit-identical: another summation order)
    tunables = {name: "" for name in libs}
    for name in (("default", "relreuse") if os.environ.get("ULTRA_AB_RELRUNS") else ()):
        if name in libs:
            libs[name + "@relruns"] = libs[name]
            tunables[name + "@relruns"] = "plan_relation_runs=1"
    for item in args.tunable_sets.split():
        name, kv = item.split(":")
        libs["default@" + name] = libs["default"]
        tunables["default@" + name] = kv
    results, lines = {}, []
    for name, path in libs.items():
        env = dict(os.environ, ULTRA_RSPMM_LIB=path, ULTRA_AB_TUNABLES=tunables[name])
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True,
                               text=True, timeout=args.timeout)
            recs = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
            err = None if r.returncode == 0 else r.stderr[-400:]
        except subprocess.TimeoutExpired:
            recs, err = [], "timeout"
        results[name] = recs
        for rec in recs:
            lines.append(dict(rec, lib=name))
        if err:
            lines.append({"lib": name, "error": err})
        print("\n".join(json.dumps(l) for l in lines[-(len(recs) + (1 if err else 0)):]), flush=True)
    base = results.get("default") or []
    sums = lambda recs: [[r.get(k) for k in ("fwd_sum", "ig_sum", "rg_sum", "upd_sum")] for r in recs]  # noqa: E731
    score = {}
    for name, recs in results.items():
        if len(recs) == len(base) and base and sums(recs) == sums(base):
            score[name] = round(sum(r.get("fwd_ms", 0.0) + 0.25 * r.get("bwd_ms", 0.0) + r.get("upd_ms", 0.0)
                                    for r in recs), 4)
    winner = min(score, key=score.get) if score else "default"
    # a variant must win by more than run-to-run noise to displace the default
    if winner != "default" and "default" in score and score[winner] > 0.985 * score["default"]:
        winner = "default"
    final = {"winner": winner, "winner_lib": libs[winner], "score_ms": score,
             "bit_identical": sorted(score), "libs": sorted(libs)}
    lines.append(final)
    print(json.dumps(final), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(json.dumps(l) for l in lines) + "\n")


if __name__ == "__main__":
    main()
