#!/usr/bin/env python
"""bench.py -- rspmm edges/s (BASELINE.json metric) on B200, with roofline, e2e and the reference CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] -- FB15k-237 scale entity graph (14,541 nodes, 474 relation
rows, 544,230 edges incl. inverse edges), dim=64, 128 heads per GPU -> D = 8192 fp32 columns, distmult + sum.
Synthetic Zipf(1.0) graph (ultra_b200.synth; no datasets / network here).  One STEP = one rspmm forward over the
whole batch (the call GeneralizedRelationalConv.message_and_aggregate makes, /root/reference/ultra/layers.py:200).

  value     : edges/s, inputs resident in HBM, plan pre-built (the reference re-sorts per call), CUDA events
  e2e       : edges/s through ultra_b200.rspmm.generalized_rspmm with pinned HOST tensors: H2D of relation+input
              and D2H of the output inside the timed region (pipelined column slabs in the C ABI); each rank is pinned
              to its GPU's NUMA node before it allocates pinned memory
  roofline  : algorithmic bytes (edge-gather model) / measured per-call device time vs the measured
              HBM copy peak in MEASURED_PEAKS.json -- PLUS the binding-resource fractions of the same kernel build from
              the committed ncu capture (profiles/roofline_traffic.json): dram_frac, l1tex_lsu_frac, l2_frac.  The
              edge-gather `frac` charges L1/L2-served gathers to the HBM peak and can exceed 1; dram_frac is the honest
              HBM fraction, l1tex_lsu_frac the resource the kernel actually leans on
  backward  : (extra, SURVEY 8d iii) ms per ultra_rspmm_backward call on the same workload, with its own roofline
  nbfnet_layer : (extra) one GeneralizedRelationalConv forward and the 6-layer entity stack at 128 heads per GPU
  strong_scaling : (extra; BASELINE configs[2] and [3]) the FULL Ultra forward - relation-graph stack, entity-graph stack,
              all-tail scoring - at a FIXED total batch (C3: 256 queries on the FB15k-237-scale KG, C4: 64 on the
              CoDEx-L-scale KG) sharded by heads over the N ranks, with the one collective the caller needs (all-gather
              of the (B/N, num_node) score rows) INSIDE the timed region, next to the same forward at the full batch on
              one GPU (every rank times it) -> in-run strong-scaling efficiency t1 / (N * tN)
  cpu_baseline / --impl reference : the reference's own CPU kernel (oracle/_ref, built from /root/reference by
              oracle/build_ref.py) on this box's host cores; thread count = the fastest of {affinity, 64, 32}; kernel-only
              and wrapper (per-call argsort, rspmm.py:174-178) timed separately, min and median, identically in both legs

N > 1: one process per GPU, heads sharded (each GPU owns 128 heads; rspmm needs no collective because columns are
independent) -> "scaling": "weak"; value = N * E / max-over-ranks device time.  The strong-scaling
records above are where a collective is on the timed path.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rspmm_edges_per_sec"
UNIT = "edges/s"
SCALE = "fb15k237"
HEADS = 128
DIM = 64
ALPHA = 1.0
SEED = 1234 + 2


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_counters():
    """Per-kernel ncu counters of the committed captures (profiles/roofline_traffic.json): {role: {...}} or {}."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            data = json.load(f)
        return data if "forward" in data else {"forward": data}  # round-1 file held the forward kernel only
    return {}


L2_GATHER_PEAK_GBS = 7164.5  # best LDG.128 row-gather rate of tools/gather_microbench.cu (profiles/r01_gather_microbench.jsonl)


def l2_gather(counters, ms):
    if not counters or not counters.get("l2_to_l1_bytes"):
        return None
    achieved = counters["l2_to_l1_bytes"] / (ms * 1e-3) / 1e9
    return {"achieved": achieved, "peak": L2_GATHER_PEAK_GBS, "unit": "GB/s", "frac": achieved / L2_GATHER_PEAK_GBS,
            "bytes_per_launch": counters["l2_to_l1_bytes"], "l1_hit_rate": counters.get("l1_hit_rate"),
            "peak_note": "rate of the 4-loads-in-flight gather microbench, NOT a hard ceiling: the two-gather backward pass "
                         "of the same template sustains 8.5 TB/s (profiles/roofline_traffic.json backward_rel); the forward "
                         "pass waits on the L2 round trip under load: neither more loads in flight nor fewer L1 wavefronts "
                         "moved it (DESIGN.md 3.1)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_workload(heads):
    from ultra_b200.synth import synth_scale
    ei, et, R, N = synth_scale(SCALE, alpha=ALPHA, seed=SEED)
    D = heads * DIM
    g = torch.Generator().manual_seed(SEED)
    # entity stack: relation rows are per-head distinct (projection of (B,R,64), layers.py:80)
    rel = torch.randn(R, D, generator=g)
    x = torch.randn(N, D, generator=g)
    w = torch.ones(ei.shape[1])
    return dict(ei=ei, et=et, w=w, rel=rel, x=x, N=N, R=R, E=ei.shape[1], D=D)


def workload_config(wl, n_gpus):
    return {"workload": "BASELINE configs[1]: FB15k-237-scale entity graph, rspmm add-mul forward (distmult+sum), "
                        "dim=64, %d heads per GPU" % HEADS,
            "num_node": wl["N"], "num_relation": wl["R"], "num_edge": wl["E"], "dim": wl["D"], "heads_per_gpu": HEADS,
            "graph": "synthetic Zipf(alpha=%.1f) + inverse edges, seed %d" % (ALPHA, SEED),
            "l2": "no flush: input+output are 2x476 MB per call, larger than the 126 MB L2",
            "parallelism": "head-shard x%d (no data-path collective in rspmm; the strong_scaling records carry the "
                           "score all-gather)" % n_gpus}


# ---- the reference's CPU path, timed the same way in both legs ---------------------------------------------------------
def _stats_ms(ts):
    return {"min": 1e3 * min(ts), "median": 1e3 * statistics.median(ts), "mean": 1e3 * sum(ts) / len(ts), "n": len(ts)}


def reference_timing(wl, steps, warmup=1):
    """oracle/_ref (the unmodified reference CPU kernel) on this box's host cores.

    Thread count: the reference uses every core torch gives it (at::parallel_for over rows, rspmm.cpp:55); on a shared
    box `os.cpu_count()` threads oversubscribe (round 1: 1902 ms vs 272 ms for the same call on two boxes), so the kernel
    is probed at {CPUs this process may run on, 64, 32} threads and the FASTEST setting is kept - the reference gets its
    best configuration.  Then `steps` iterations time the kernel alone (pre-sorted edges) and the wrapper (per-call
    argsort + gathers + kernel, what the reference's public generalized_rspmm does), each reported as min / median / mean."""
    import oracle
    from oracle import ref
    ei, et, w, rel, x = wl["ei"], wl["et"], wl["w"], wl["rel"], wl["x"]
    sei, stt, sw, _ = oracle.sort_edges(ei.numpy(), et.numpy(), w.numpy())
    tei, ttt, tw = torch.from_numpy(sei), torch.from_numpy(stt), torch.from_numpy(sw)
    avail = len(os.sched_getaffinity(0))
    tried = {}
    out = None
    for th in sorted({avail, min(avail, 64), min(avail, 32)}, reverse=True):
        torch.set_num_threads(th)
        out = ref.ref_forward("add", "mul", tei, ttt, tw, rel, x)  # warm-up at this setting
        t0 = time.perf_counter()
        ref.ref_forward("add", "mul", tei, ttt, tw, rel, x)
        tried[th] = time.perf_counter() - t0
    best = min(tried, key=tried.get)
    torch.set_num_threads(best)
    for _ in range(warmup):
        ref.ref_generalized_rspmm(ei, et, w, rel, x)
    tk, tw_ = [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        ref.ref_forward("add", "mul", tei, ttt, tw, rel, x)
        tk.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        ref.ref_generalized_rspmm(ei, et, w, rel, x)
        tw_.append(time.perf_counter() - t0)
    cpu_model = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    cpu_model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"threads": best, "threads_tried_kernel_ms": {str(k): 1e3 * v for k, v in tried.items()},
            "cpus_allowed": avail, "os_cpu_count": os.cpu_count(), "cpu": cpu_model,
            "kernel_ms": _stats_ms(tk), "wrapper_ms": _stats_ms(tw_), "wrapper_seconds": tw_, "output": out}


def run_reference(args):
    """--impl reference: the reference's CPU rspmm (oracle/_ref) on the host cores, same workload and metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/rspmm_ref.so not built (needs /root/reference)"}))
        return 0
    wl = make_workload(HEADS)
    rt = reference_timing(wl, args.steps, warmup=max(args.warmup, 1))
    total = sum(rt["wrapper_seconds"])  # EXACTLY K timed steps through the reference's public wrapper
    value = wl["E"] * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": rt["threads"], "kind": "reference",
                         "sample": "full C2 batch (128 heads, D=8192) per step through the reference wrapper "
                                   "(per-call argsort + rspmm_add_mul_forward_cpu), %d steps" % args.steps,
                         "kernel_ms": rt["kernel_ms"], "wrapper_ms": rt["wrapper_ms"],
                         "kernel_value_median": wl["E"] / (rt["kernel_ms"]["median"] * 1e-3),
                         "threads_tried_kernel_ms": rt["threads_tried_kernel_ms"], "cpus_allowed": rt["cpus_allowed"],
                         "os_cpu_count": rt["os_cpu_count"], "cpu": rt["cpu"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def parity_against(ref_out, gpu_out):
    """SURVEY 8d "parity check in the same run": the GPU forward of the timed workload against the reference CPU kernel's
    output on the same tensors, north-star tolerance 1e-5 relative (to max|ref|)."""
    try:
        if gpu_out is None:
            return {"checked": False, "why": "no GPU output"}
        scale = float(ref_out.abs().max())
        err = float((gpu_out - ref_out).abs().max())
        return {"checked": True, "max_abs_err": err, "max_abs_ref": scale, "rel_to_max": err / scale if scale else 0.0,
                "tolerance": 1e-5, "ok": bool(err <= 1e-5 * scale),
                "what": "ultra_rspmm_forward output of the timed workload (all %d x %d elements) vs "
                        "rspmm_add_mul_forward_cpu of the reference" % tuple(ref_out.shape)}
    except Exception as exc:  # noqa: BLE001
        return {"checked": False, "why": repr(exc)}


def cpu_baseline(wl, gpu_out=None):
    """Reference CPU kernel (oracle/_ref) on this box's host cores, bounded sample: full C2 batch, 3 timed steps."""
    from oracle import ref
    import oracle
    if not ref.available():
        # C restatement (scalar, 1 thread) on a 4-head crop, normalised to the full width
        crop = 4 * DIM
        sei, stt, sw, _ = oracle.sort_edges(wl["ei"].numpy(), wl["et"].numpy(), wl["w"].numpy())
        rel, x = wl["rel"][:, :crop].contiguous().numpy(), wl["x"][:, :crop].contiguous().numpy()
        t0 = time.perf_counter()
        oracle.c_forward("add", "mul", sei, stt, sw, rel, x)
        dt = (time.perf_counter() - t0) * (wl["D"] / crop)
        return {"value": wl["E"] / dt, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": "C oracle, 4 of 128 heads, time scaled x32 to D=8192"}
    rt = reference_timing(wl, 3, warmup=1)
    parity = parity_against(rt["output"], gpu_out)
    return {"value": wl["E"] / (rt["kernel_ms"]["min"] * 1e-3), "unit": UNIT, "cores": rt["threads"], "kind": "reference",
            "sample": "full C2 batch (128 heads, D=8192), kernel call on pre-sorted edges, min of 3 after warm-up; "
                      "threads = fastest of the settings tried",
            "wrapper_value": wl["E"] / (rt["wrapper_ms"]["min"] * 1e-3), "kernel_ms": rt["kernel_ms"],
            "wrapper_ms": rt["wrapper_ms"], "threads_tried_kernel_ms": rt["threads_tried_kernel_ms"],
            "cpus_allowed": rt["cpus_allowed"], "cpu": rt["cpu"], "os_cpu_count": rt["os_cpu_count"], "parity": parity}


# ---- device timing helper ------------------------------------------------------------------------------------------------
def make_timer(device, world, dist):
    def timeit(fn, reps=5, warm=2, collective=True):
        """mean ms per call over `reps` back-to-back calls, CUDA events on the current stream; max over ranks."""
        for _ in range(warm):
            fn()
        if world > 1 and collective:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=device, dtype=torch.float64)
        if world > 1 and collective:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return timeit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="e2e steps (default min(steps, 10))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-layer", action="store_true", help="skip the NBFNet-layer / 6-layer stack timing")
    ap.add_argument("--no-strong", action="store_true", help="skip the C3 / C4 strong-scaling records")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the rspmm product path has no CPU implementation)")
    # pin the rank (and with it every pinned allocation below: first touch) to the GPU's NUMA node
    from ultra_b200.numa import pin_to_gpu_numa
    numa = pin_to_gpu_numa(local_rank) if world > 1 else {"node": None, "why": "single rank: not pinned"}
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    from ultra_b200 import _capi
    from ultra_b200.plan import GraphPlan, _stream_ptr
    from ultra_b200.rspmm import generalized_rspmm
    from ultra_b200.synth import algorithmic_bytes, compulsory_bytes
    lib = _capi.lib()

    wl = make_workload(HEADS)
    N, R, E, D = wl["N"], wl["R"], wl["E"], wl["D"]
    ei_d, et_d, w_d = wl["ei"].to(device), wl["et"].to(device), wl["w"].to(device)
    rel_d, x_d = wl["rel"].to(device), wl["x"].to(device)
    out_d = torch.empty_like(x_d)
    torch.cuda.synchronize()
    t_plan = time.perf_counter()
    plan = GraphPlan(ei_d, et_d, N, R)  # one-time per graph: 3 device sorts + item cutting (reference: a sort per CALL)
    torch.cuda.synchronize()
    first_plan_ms = 1e3 * (time.perf_counter() - t_plan)
    t_plan = time.perf_counter()
    plan2 = GraphPlan(ei_d, et_d, N, R)
    torch.cuda.synchronize()
    plan_build_ms = 1e3 * (time.perf_counter() - t_plan)
    del plan2
    # incremental re-cut for a masked edge set (training's per-step edge removal): device time, nothing synchronises
    keep_d = torch.rand(E, device=device) > 0.002
    for _ in range(2):
        plan.apply_edge_mask(keep_d)
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record()
    for _ in range(5):
        plan.apply_edge_mask(keep_d)
    m1.record()
    plan.apply_edge_mask(None)
    torch.cuda.synchronize()
    apply_mask_ms = m0.elapsed_time(m1) / 5
    plan.bind_weight(w_d)
    scratch = plan.scratch(D, torch.float32)
    has_empty = bool((torch.bincount(wl["ei"][0], minlength=N) == 0).any())
    launches_per_step = 1 + (1 if plan.info(4) > 0 else 0) + (1 if has_empty else 0)  # tiled kernel, fold, identity fill

    def step():
        rc = lib.ultra_rspmm_forward(plan.handle, 0, 0, 0, rel_d.data_ptr(), x_d.data_ptr(), out_d.data_ptr(), D,
                                     scratch.data_ptr(), scratch.numel(), _stream_ptr(device))
        if rc != 0:
            raise RuntimeError(lib.ultra_rspmm_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * E * args.steps / (total_ms * 1e-3)

    # ---- e2e: public API, pinned host tensors, H2D + D2H inside the timed region -------------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    rel_h, x_h = wl["rel"].pin_memory(), wl["x"].pin_memory()
    ei_h, et_h, w_h = wl["ei"], wl["et"], wl["w"]
    for _ in range(2):
        out_h = generalized_rspmm(ei_h, et_h, w_h, rel_h, x_h, sum="add", mul="mul")
        del out_h
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for _ in range(e2e_steps):
        out_h = generalized_rspmm(ei_h, et_h, w_h, rel_h, x_h, sum="add", mul="mul")  # returns a complete CPU tensor
        checksum += float(out_h[0, 0])
        del out_h
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * E * e2e_steps / float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    del rel_h, x_h

    timeit = make_timer(device, world, dist)
    peak, peak_src = load_peaks()
    counters = load_counters()
    # ---- rspmm backward (relation_grad + input_grad, deterministic default) - guarded extra
    backward_info = None
    try:
        og_d = torch.randn_like(x_d)
        ig_d, rg_d = torch.empty_like(x_d), torch.empty_like(rel_d)

        def bwd():
            rc = lib.ultra_rspmm_backward(plan.handle, 0, 0, 0, rel_d.data_ptr(), x_d.data_ptr(), out_d.data_ptr(),
                                          og_d.data_ptr(), None, rg_d.data_ptr(), ig_d.data_ptr(), D,
                                          scratch.data_ptr(), scratch.numel(), _stream_ptr(device))
            if rc != 0:
                raise RuntimeError(lib.ultra_rspmm_last_error().decode())
        bwd_ms = timeit(bwd, reps=5, warm=2, collective=False)
        # edge-gather model of the backward: per edge one output_grad row slice for input_grad and (two-pass default) one
        # output_grad + one input row slice for relation_grad; input_grad and relation_grad written once; relation read once
        bwd_bytes = 4 * D * (3 * E + N + 2 * R) + 2 * 12 * E + 4 * (N + R + 2)
        _capi.set_tunable("bwd_fused_relgrad", 1)
        try:
            fused_ms = timeit(bwd, reps=5, warm=2, collective=False)
        finally:
            _capi.set_tunable("bwd_fused_relgrad", 0)
        cb = counters.get("backward_in"), counters.get("backward_rel")
        l2b = None
        if cb[0] and cb[1]:
            l2b = l2_gather({"l2_to_l1_bytes": cb[0]["l2_to_l1_bytes"] + cb[1]["l2_to_l1_bytes"]}, bwd_ms)
        backward_info = {"ms_per_call": bwd_ms, "calls": 5, "ratio_to_forward": bwd_ms / (sum(per_step) / len(per_step)),
                         "fused_relgrad_ms_per_call": fused_ms, "l2_gather": l2b,
                         "fused_note": "opt-in one-pass backward (tunable bwd_fused_relgrad): relation_grad by vector "
                                       "reductions at L2, order-dependent in its last bits like the reference's atomicAdd "
                                       "backward; wins on uniform relation frequencies, loses on Zipf",
                         "what": "ultra_rspmm_backward add-mul, same workload: input_grad + relation_grad, no atomics",
                         "roofline": {"bound": "hbm", "achieved": bwd_bytes / (bwd_ms * 1e-3) / 1e9, "peak": peak,
                                      "unit": "GB/s", "frac": bwd_bytes / (bwd_ms * 1e-3) / 1e9 / peak,
                                      "model": "edge-gather: 4*D*(3E+N+2R)+24E+4(N+R+2) = %d bytes per call" % bwd_bytes,
                                      "counters": counters.get("backward")}}
        del og_d, ig_d, rg_d
    except Exception as exc:  # noqa: BLE001 - extras only
        backward_info = {"error": repr(exc)}

    layer_info = None
    if not args.no_layer:
        try:
            layer_info = bench_layer(wl, device, world, rank, timeit, counters)
        except Exception as exc:  # noqa: BLE001 - extras only
            layer_info = {"error": repr(exc)}
    torch.cuda.empty_cache()
    strong = None
    if not args.no_strong:
        strong = []
        for cfg in ("C3", "C3-b16", "C4"):
            try:
                strong.append(bench_strong(cfg, device, world, rank, dist, timeit))
            except Exception as exc:  # noqa: BLE001 - extras only (every rank takes the same path: shapes are static)
                strong.append({"config": cfg, "error": repr(exc)})
            torch.cuda.empty_cache()

    if rank == 0:
        kernel_ms = sum(per_step) / len(per_step)
        ba = algorithmic_bytes(N, E, R, D)
        achieved = ba / (kernel_ms * 1e-3) / 1e9
        fwd_c = counters.get("forward") or {}
        dram = fwd_c.get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(wl, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": dram, "peak_source": peak_src,
                         "model": "edge-gather: 4*D*(E+N+R)+12*E+4*(N+1) = %d bytes per call; duration = whole forward "
                                  "call (tiled kernel + split-row fix-up), mean of %d CUDA-event samples on the "
                                  "launching stream.  This model charges every L1/L2-served gather to HBM, so frac can "
                                  "exceed 1; see dram_frac / l1tex_lsu_frac for the binding resources" % (ba, len(per_step)),
                         "compulsory_bytes": compulsory_bytes(N, E, R, D),
                         "achieved_compulsory_GBs": compulsory_bytes(N, E, R, D) / (kernel_ms * 1e-3) / 1e9,
                         # binding-resource fractions: DRAM bytes of the ncu capture over THIS run's duration vs the HBM
                         # peak; pipe utilisations as captured (percent of peak sustained, same kernel build)
                         "dram_frac": (dram / (kernel_ms * 1e-3) / 1e9 / peak) if dram else None,
                         "l1tex_lsu_frac": fwd_c.get("l1tex_lsu_frac"), "l2_frac": fwd_c.get("l2_frac"),
                         # what actually bounds the kernel: bytes L1 pulls from L2 (ncu lts__t_sectors_srcunit_tex_op_read of
                         # the capture) over THIS run's duration, against the measured ceiling for random 128-byte row
                         # gathers out of an L2-resident tile (tools/gather_microbench.cu, profiles/r01_gather_microbench.jsonl)
                         "l2_gather": l2_gather(fwd_c, kernel_ms),
                         "counters_source": fwd_c.get("source")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * D * (N + R),
                    "d2h_bytes_per_step": 4 * D * N, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "api": "ultra_b200.rspmm.generalized_rspmm(pinned CPU tensors) -> ultra_rspmm_forward_host",
                    "checksum": checksum, "numa": numa},
            "gpu_launches": launches_per_step * args.steps,
            "backward": backward_info,
            "plan_build_ms": {"warm": plan_build_ms, "first_call_incl_context": first_plan_ms,
                              "forward_incl_plan_build_ms": kernel_ms + plan_build_ms,
                              "apply_mask_device_ms": apply_mask_ms,
                              "note": "once per graph, outside the timed region; the reference wrapper re-sorts the "
                                      "edges on every call (rspmm.py:174-178).  apply_mask = re-cut of the built plan for "
                                      "a masked edge set (remove_easy_edges), on the device, no sort, no host sync"},
            "clocks": clocks,
            "nbfnet_layer": layer_info,
            "strong_scaling": strong,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:  # the timed step's own output, brought to the host for the parity check against the reference kernel
                step()
                torch.cuda.synchronize()
                gpu_out = out_d.cpu()
            except Exception:  # noqa: BLE001 - the parity extra must never fail the run
                gpu_out = None
            line["cpu_baseline"] = cpu_baseline(wl, gpu_out)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def bench_layer(wl, device, world, rank, timeit, counters):
    """NBFNet-layer ms (second half of BASELINE.json's metric): one GeneralizedRelationalConv forward of the entity
    stack (rspmm with folded boundary, Linear 128->64, LayerNorm, ReLU, shortcut; /root/reference/ultra/layers.py
    :67-88,183-240 restated PyG-free in ultra_b200.layer) at HEADS heads per GPU, and the 6-layer Bellman-Ford stack."""
    from ultra_b200.layer import GeneralizedRelationalConv, bellmanford_entity
    from ultra_b200.fused import nbfnet_update
    N, R = wl["N"], wl["R"]
    B, d = HEADS, DIM
    torch.manual_seed(7)
    ei, et = wl["ei"].to(device), wl["et"].to(device)
    layers = [GeneralizedRelationalConv(d, d, R, d, "distmult", "sum", layer_norm=True, project_relations=True).to(device)
              for _ in range(6)]
    rel_repr = torch.randn(B, R, d, device=device)
    query = torch.randn(B, d, device=device)
    h_index = torch.randint(0, N, (B,), device=device)
    for layer in layers:
        layer.relation = rel_repr
    x = torch.randn(B, N, d, device=device)
    bd = torch.zeros(B, N, d, device=device)
    w = torch.ones(ei.shape[1], device=device)
    with torch.no_grad():
        layer_ms = timeit(lambda: layers[0](x, query, bd, ei, et, (N, N), w, shortcut=True))
        stack_ms = timeit(lambda: bellmanford_entity(layers, query, h_index, ei, et, N), reps=3, warm=1)
        u = torch.randn_like(x)
        upd_ms = timeit(lambda: nbfnet_update(x, u, layers[0].linear, layers[0].layer_norm, relu=True, shortcut=True),
                        reps=10, warm=3, collective=False)
        upd_bytes = 3 * x.numel() * 4
        peak, _ = load_peaks()
        info = {"layer_ms": layer_ms, "stack6_ms": stack_ms, "heads_per_gpu": B,
                "what": "GeneralizedRelationalConv forward (distmult+sum, layer_norm, relu, shortcut), fp32: "
                        "relation_projection MLP + head-major rspmm with folded boundary + fused tcgen05 "
                        "Linear/LayerNorm/ReLU/residual kernel (SURVEY 8f row f2); stack6 = 6-layer Bellman-Ford loop",
                "update_kernel": {"ms": upd_ms, "roofline": {"bound": "hbm", "achieved": upd_bytes / (upd_ms * 1e-3) / 1e9,
                                                             "peak": peak, "unit": "GB/s",
                                                             "frac": upd_bytes / (upd_ms * 1e-3) / 1e9 / peak,
                                                             "model": "compulsory: read x, read u, write out = %d bytes"
                                                                      % upd_bytes,
                                                             "counters": counters.get("update")}}}
    # one TRAINING step of the layer (second half): forward + backward through rspmm and the epilogue,
    # gradients to the node states, the relation representations and every layer parameter
    try:
        layer = layers[0]
        x_t = x.clone().requires_grad_()
        layer.relation = rel_repr.clone().requires_grad_()
        gout = torch.randn_like(x)

        def train_step():
            layer.zero_grad(set_to_none=True)
            x_t.grad = None
            layer.relation.grad = None
            out = layer(x_t, query, bd, ei, et, (N, N), w, shortcut=True)
            out.backward(gout)
        info["training_layer_ms"] = timeit(train_step, reps=3, warm=2)
        info["training_layer_what"] = ("forward + backward of one entity layer at %d heads (grads: input, relation "
                                       "representations, relation_projection, linear, layer_norm)" % B)
        del x_t, gout
    except Exception as exc:  # noqa: BLE001 - extras only
        info["training_layer_error"] = repr(exc)
    del layers, x, bd, u
    torch.cuda.empty_cache()
    # a whole TRAINING step of the 6-layer entity stack (forward + backward, this step's easy edges removed through the
    # incremental plan: ultra_rspmm_plan_apply_mask, no sort, no host sync), 32 heads
    try:
        from ultra_b200.nbfnet import UltraNBFNet
        torch.manual_seed(9)
        model = UltraNBFNet().to(device).train()
        Bt = 32
        rel_t = torch.randn(Bt, R, d, device=device, requires_grad=True)
        h_t = torch.randint(0, N, (Bt,), device=device)
        r_t = torch.randint(0, R // 2, (Bt,), device=device)
        t_t = torch.randint(0, N, (Bt,), device=device)
        keep = UltraNBFNet.easy_edge_mask(ei, et, N, R, h_t, t_t, r_t)

        def train_stack():
            model.zero_grad(set_to_none=True)
            rel_t.grad = None
            hidden, _ = model.entity_states(ei, et, N, rel_t, h_t, r_t, edge_mask=keep)
            hidden.square().mean().backward()
        info["training_stack6_ms"] = timeit(train_stack, reps=3, warm=2)
        info["training_stack6_what"] = ("6-layer entity stack, forward + backward, %d heads, per-step edge mask applied through the "
                                        "incremental plan (%d of %d edges removed)" % (Bt, int((~keep).sum()), ei.shape[1]))
        from ultra_b200.plan import plan_cache
        plan_cache.get(ei, et, N, R).apply_edge_mask(None)
        del model, rel_t
    except Exception as exc:  # noqa: BLE001 - extras only
        info["training_stack6_error"] = repr(exc)
    torch.cuda.empty_cache()
    return info


STRONG = {
    # BASELINE configs[2]: FB15k-237 inductive-scale KG, full relation-graph + entity-graph stack, batch 256
    "C3": dict(scale="fb15k237", batch=256, seed=1234 + 3,
               what="BASELINE configs[2]: FB15k-237-scale KG + its relation graph, full Ultra forward, batch 256"),
    # the launch-bound end of the same configuration: 16 queries in total
    "C3-b16": dict(scale="fb15k237", batch=16, seed=1234 + 3,
                   what="BASELINE configs[2] at batch 16: FB15k-237-scale KG + its relation graph, full Ultra forward"),
    # BASELINE configs[3]: CoDEx-L scale (77,951 nodes, 1.1 M edges), 6 layers, batch 64 (8 per GPU at 8 GPUs)
    "C4": dict(scale="codex_l", batch=64, seed=1234 + 4,
               what="BASELINE configs[3]: CoDEx-L-scale KG + its relation graph, full Ultra forward, batch 64"),
}


def bench_strong(name, device, world, rank, dist, timeit):
    """Strong head-shard scaling of the full Ultra forward at a FIXED total batch (see the module docstring)."""
    from ultra_b200 import _capi
    from ultra_b200.nbfnet import UltraNBFNet, CudaGraphRunner
    from ultra_b200.plan import plan_cache, _stream_ptr
    from ultra_b200.shard import head_range, all_gather_heads
    from ultra_b200.synth import synth_scale, relation_graph_of
    cfg = STRONG[name]
    B = cfg["batch"]
    ei, et, R, N = synth_scale(cfg["scale"], alpha=ALPHA, seed=cfg["seed"])
    ei, et = ei.to(device), et.to(device)
    rei, ret, rR, rN = relation_graph_of(ei, et, N, R)
    rei, ret = rei.contiguous(), ret.contiguous()
    graph, rel_graph = (ei, et, N), (rei, ret, rN)
    torch.manual_seed(11)  # identical random-init weights on every rank (no checkpoints / network here)
    model = UltraNBFNet().to(device).eval()
    g = torch.Generator().manual_seed(cfg["seed"])
    h_index = torch.randint(0, N, (B,), generator=g).to(device)
    r_index = torch.randint(0, R, (B,), generator=g).to(device)
    lo, hi = head_range(B, rank, world)
    fwd = lambda h, r: model(rel_graph, graph, h, r)  # noqa: E731  -> (heads, N) scores of every entity as the tail

    info = {"config": name, "what": cfg["what"], "total_batch": B, "heads_per_gpu": hi - lo, "n_gpus": world,
            "num_node": N, "num_edge": int(ei.shape[1]), "num_relation": R,
            "relation_graph": {"num_node": rN, "num_edge": int(rei.shape[1])},
            "collective": "one all_gather_into_tensor of the (B/N, num_node) fp32 score rows (what Ultra.forward returns, "
                          "models.py:208), inside the timed region",
            "weights": "random init (seeded, identical on every rank)"}
    with torch.no_grad():
        # (a) per-GPU rspmm forward at the local head count (entity graph, flat layout, the headline kernel)
        Dl = (hi - lo) * DIM
        plan = plan_cache.get(ei, et, N, R)
        plan.bind_weight(model._unit_weight(ei.shape[1], torch.empty(0, device=device)))
        rel_l, x_l = torch.randn(R, Dl, device=device), torch.randn(N, Dl, device=device)
        out_l = torch.empty_like(x_l)
        scr = plan.scratch(Dl, torch.float32)
        lib = _capi.lib()

        def rspmm_local():
            rc = lib.ultra_rspmm_forward(plan.handle, 0, 0, 0, rel_l.data_ptr(), x_l.data_ptr(), out_l.data_ptr(), Dl,
                                         scr.data_ptr(), scr.numel(), _stream_ptr(device))
            if rc != 0:
                raise RuntimeError(lib.ultra_rspmm_last_error().decode())
        info["rspmm_ms_per_gpu"] = timeit(rspmm_local, reps=10, warm=3)
        # work unit that is comparable across N: one edge x one head (an rspmm call at B/N heads walks the same E edges
        # over 1/N of the columns); aggregate over the ranks = E * B / (max-over-ranks time)
        info["rspmm_edge_heads_per_sec_aggregate"] = int(ei.shape[1]) * B / (info["rspmm_ms_per_gpu"] * 1e-3)
        del rel_l, x_l, out_l
        # (a2) the all-tails scoring MLP at the local head count (ultra_score_head_f32; models.py:166-170,202-208)
        h_s, q_s = torch.randn(hi - lo, N, DIM, device=device), torch.randn(hi - lo, DIM, device=device)
        info["score_head_ms_per_gpu"] = timeit(lambda: model.score_all(h_s, q_s), reps=10, warm=3)
        del h_s, q_s
        # (b) the sharded forward, captured as ONE CUDA graph per rank, + the score all-gather on the timed path
        local = CudaGraphRunner(fwd, h_index[lo:hi], r_index[lo:hi])
        scores = torch.empty(B, N, device=device)
        h_l, r_l = h_index[lo:hi].clone(), r_index[lo:hi].clone()
        info["local_ms"] = timeit(lambda: local(h_l, r_l), reps=5, warm=2)

        def sharded():
            s = local(h_l, r_l)
            if world > 1:
                all_gather_heads(s, B, out=scores)
            return s
        info["stack_ms"] = timeit(sharded, reps=5, warm=2)
        info["exposed_collective_ms"] = info["stack_ms"] - info["local_ms"]
        info["collective_bytes_received_per_gpu"] = 4 * N * (B - (hi - lo)) if world > 1 else 0
        if world > 1 and info["exposed_collective_ms"] > 0:
            info["collective_GBs"] = info["collective_bytes_received_per_gpu"] / (info["exposed_collective_ms"] * 1e-3) / 1e9
        # parity of the sharding itself, same run: the gathered score rows of this rank's heads == its local scores
        if world > 1:
            info["gathered_equals_local"] = bool(torch.equal(scores[lo:hi], local.static_out))
        del local
        torch.cuda.empty_cache()
        # (c) the same forward at the FULL batch on one GPU (every rank does the identical work; max over ranks)
        if world > 1:
            full = CudaGraphRunner(fwd, h_index, r_index)
            info["single_gpu_full_batch_ms"] = timeit(lambda: full(h_index, r_index), reps=3, warm=1)
            if info.get("gathered_equals_local") is not None:
                ref_rows = full(h_index, r_index)[lo:hi]
                err = float((ref_rows - scores[lo:hi]).abs().max())
                info["sharded_vs_full_batch_max_abs_diff"] = err  # heads are independent: bit-identical expected
            del full
        else:
            info["single_gpu_full_batch_ms"] = info["stack_ms"]
        info["strong_scaling_efficiency"] = info["single_gpu_full_batch_ms"] / (world * info["stack_ms"])
        info["queries_per_sec"] = B / (info["stack_ms"] * 1e-3)
    torch.cuda.empty_cache()
    return info


if __name__ == "__main__":
    sys.exit(main())
