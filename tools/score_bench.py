"""Fused scoring MLP (ultra_score_head_f32) against the eager chunked formulation it replaces, at the sizes of the
strong-scaling configs: C3 (256 queries x 14 541 entities) and C4 (64 x 77 951).  CUDA events; the states (0.24 - 1.3 GB) are larger than L2."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.kbench import time_fn  # noqa: E402
from ultra_b200 import fused  # noqa: E402


def eager_score_all(h, query, mlp, chunk=32):
    d = h.shape[-1]
    wa, wb = mlp[0].weight[:, :d], mlp[0].weight[:, d:]
    qb = torch.addmm(mlp[0].bias, query, wb.t())
    out = h.new_empty(h.shape[0], h.shape[1])
    for lo in range(0, h.shape[0], chunk):
        hid = torch.matmul(h[lo:lo + chunk], wa.t())
        hid.add_(qb[lo:lo + chunk].unsqueeze(1)).relu_()
        out[lo:lo + chunk] = torch.matmul(hid, mlp[2].weight.t()).squeeze(-1) + mlp[2].bias
    return out


def main():
    d = torch.device("cuda:0")
    torch.manual_seed(0)
    mlp = torch.nn.Sequential(torch.nn.Linear(128, 128), torch.nn.ReLU(), torch.nn.Linear(128, 1)).to(d)
    for name, B, N in (("C3", 256, 14541), ("C4", 64, 77951), ("C3 / 8 GPUs", 32, 14541)):
        h = torch.randn(B, N, 64, device=d)
        q = torch.randn(B, 64, device=d)
        with torch.no_grad():
            ref = eager_score_all(h, q, mlp)
            got = fused.score_head(h, q, mlp)
            err = float((got - ref).abs().max() / ref.abs().max())
            e_med, e_min = time_fn(lambda: eager_score_all(h, q, mlp), warmup=2, iters=7)
            f_med, f_min = time_fn(lambda: fused.score_head(h, q, mlp), warmup=3, iters=15)
        flop = 2.0 * B * N * 64 * 128
        print(json.dumps({"case": name, "rows": B * N, "eager_ms": round(e_med, 4), "fused_ms": round(f_med, 4),
                          "fused_ms_min": round(f_min, 4), "fused_TFLOPs": round(flop / f_med / 1e9, 2),
                          "h_read_GBs": round(B * N * 256 / f_med / 1e6, 1), "rel_err_vs_eager": err}), flush=True)


if __name__ == "__main__":
    main()
