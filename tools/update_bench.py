"""Fused layer epilogue (ultra_nbfnet_update_f32) at C2 size, timed with CUDA events and checked against an fp64
PyTorch evaluation.  Optional argument: comma-separated sm_limit values (0 = every SM)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.kbench import time_fn  # noqa: E402
from ultra_b200 import _capi  # noqa: E402
from ultra_b200.fused import nbfnet_update  # noqa: E402

d = torch.device("cuda:0")
limits = [int(a) for a in (sys.argv[1] if len(sys.argv) > 1 else "0").split(",")]
torch.manual_seed(3)
lin = torch.nn.Linear(128, 64).to(d)
ln = torch.nn.LayerNorm(64).to(d)
with torch.no_grad():
    ln.weight.uniform_(0.5, 1.5)
    ln.bias.uniform_(-0.5, 0.5)
M = 128 * 14541
x = torch.randn(M, 64, device=d)
u = torch.randn(M, 64, device=d) * 3
idx = torch.randint(0, M, (200000,), device=d)
idx[:300] = torch.arange(M - 300, M, device=d)
with torch.no_grad():
    ref = torch.nn.functional.linear(torch.cat([x[idx], u[idx]], -1).double(), lin.weight.double(), lin.bias.double())
    ref = torch.relu(torch.nn.functional.layer_norm(ref, (64,), ln.weight.double(), ln.bias.double(), ln.eps)) + x[idx].double()
for lim in limits:
    with torch.no_grad():
        got = nbfnet_update(x, u, lin, ln, relu=True, shortcut=True, sm_limit=lim)
        torch.cuda.synchronize()
        err = float((got[idx].double() - ref).abs().max())
        med, best = time_fn(lambda: nbfnet_update(x, u, lin, ln, relu=True, shortcut=True, sm_limit=lim), warmup=3, iters=10)
    gb = M * (128 + 64) * 4 / 1e9
    print(json.dumps({"sm_limit": lim, "rows": M, "ms_med": round(med, 4), "ms_min": round(best, 4),
                      "max_abs_err_vs_fp64": err, "compulsory_GBs": round(gb / med * 1e3, 1)}), flush=True)
