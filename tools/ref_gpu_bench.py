"""The reference's OWN CUDA kernels (ultra/rspmm/source/rspmm.cu, unmodified, compiled for sm_100a into oracle/_ref by
`python oracle/build_ref.py --cuda`) timed on the same B200 beside ours, at BASELINE configs[1].  Test/measurement
infrastructure: nothing in the product imports this.  Prints JSON lines; also checks our output against theirs."""
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.kbench import time_fn  # noqa: E402
from ultra_b200.rspmm import generalized_rspmm  # noqa: E402
from ultra_b200.synth import synth_scale  # noqa: E402

so = os.path.join(ROOT, "oracle", "_ref", "rspmm_ref_cuda.so")
spec = importlib.util.spec_from_file_location("rspmm_ref_cuda", so)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

d = torch.device("cuda:0")
for alpha in (1.0, 0.0):
    ei, et, R, N = synth_scale("fb15k237", alpha=alpha, seed=1236)
    E = ei.shape[1]
    ei, et = ei.to(d), et.to(d)
    w = torch.ones(E, device=d)
    for heads in (32, 128):
        D = heads * 64
        g = torch.Generator(device=d).manual_seed(1)
        rel, x = torch.randn(R, D, device=d, generator=g), torch.randn(N, D, device=d, generator=g)
        og = torch.randn(N, D, device=d, generator=g)
        # reference wrapper semantics (rspmm.py:167-178 + :15-25): per-call argsort, gathers, sortedness assert, kernel
        def ref_wrapper():
            key = ei[0] * (ei[1].max() + 1) + ei[1]
            order = key.argsort()
            sei, stt, sw = ei[:, order], et[order], w[order]
            k2 = sei[0] * (sei[1].max() + 1) + sei[1]
            assert (k2.diff() >= 0).all()
            return ref.rspmm_add_mul_forward_cuda(sei, stt, sw, rel, x)
        key = ei[0] * (ei[1].max() + 1) + ei[1]
        order = key.argsort()
        sei, stt, sw = ei[:, order].contiguous(), et[order].contiguous(), w[order].contiguous()
        ref_kernel = lambda: ref.rspmm_add_mul_forward_cuda(sei, stt, sw, rel, x)   # launcher incl. ind2ptr + allocs
        ours = lambda: generalized_rspmm(ei, et, w, rel, x)
        out_ref = ref_kernel()
        out_ours = ours()
        err = float((out_ours - out_ref).abs().max() / out_ref.abs().max())
        t_rk, _ = time_fn(ref_kernel, iters=5)
        t_rw, _ = time_fn(ref_wrapper, iters=5)
        t_o, _ = time_fn(ours, iters=5)
        ref_bwd = lambda: ref.rspmm_add_mul_backward_cuda(sei, stt, sw, rel, x, out_ref, og)
        relg, xg = rel.clone().requires_grad_(), x.clone().requires_grad_()
        def ours_fb():
            o = generalized_rspmm(ei, et, w, relg, xg)
            o.backward(og)
            relg.grad = xg.grad = None
        t_rb, _ = time_fn(ref_bwd, iters=3)
        t_ofb, _ = time_fn(ours_fb, iters=3)
        print(json.dumps({"graph": "fb15k237-scale zipf" if alpha else "fb15k237-scale uniform", "heads": heads, "D": D,
                          "reference_cuda_fwd_ms": round(t_rk, 3), "reference_wrapper_fwd_ms": round(t_rw, 3),
                          "ours_fwd_ms_python_api": round(t_o, 3), "speedup_vs_reference_cuda_kernel": round(t_rk / t_o, 2),
                          "reference_cuda_bwd_ms": round(t_rb, 3), "ours_fwd_plus_bwd_ms": round(t_ofb, 3),
                          "max_rel_diff_vs_reference_cuda": err}), flush=True)
