"""PNA aggregation in one pass.

The reference layer issues FOUR rspmm calls over the same edges for `aggregate_func == "pna"`
(/root/reference/ultra/layers.py:211-215): sum, sum over (relation**2, input**2), max, min.
`generalized_rspmm_pna` gathers every input row slice once and returns all four (ultra_rspmm_forward_pna).
Gradients compose from the four existing backward kernels (same cost as the reference's four backward calls).
"""
import torch
from torch import autograd

from .. import _capi
from ..plan import plan_cache, _DTYPE_ID, _stream_ptr
from .rspmm import _forward_check, _run_backward


class RSPMMPnaFunction(autograd.Function):

    @staticmethod
    def forward(ctx, edge_index, edge_type, edge_weight, relation, input, mul):
        _forward_check(edge_index, edge_type, edge_weight, relation, input)
        if input.device.type != "cuda":
            raise RuntimeError("generalized_rspmm_pna needs CUDA tensors")
        lib = _capi.lib()
        mul_id = _capi.MUL_IDS[mul]
        plan = plan_cache.get(edge_index, edge_type, input.size(0), relation.size(0))
        plan.bind_weight(edge_weight)
        rel_c, in_c = relation.detach().contiguous(), input.detach().contiguous()
        dim = in_c.size(1)
        out4 = torch.empty((4,) + tuple(in_c.shape), dtype=in_c.dtype, device=in_c.device)
        need = 4 * lib.ultra_rspmm_scratch_bytes(plan.handle, dim, _DTYPE_ID[in_c.dtype])
        scratch = getattr(plan, "_scratch_pna", None)
        if scratch is None or scratch.numel() < need:
            scratch = torch.empty(max(need, 256), dtype=torch.uint8, device=in_c.device)
            plan._scratch_pna = scratch
        with torch.cuda.device(in_c.device):
            rc = lib.ultra_rspmm_forward_pna(plan.handle, mul_id, _DTYPE_ID[in_c.dtype], rel_c.data_ptr(),
                                             in_c.data_ptr(), out4.data_ptr(), dim, scratch.data_ptr(),
                                             scratch.numel(), _stream_ptr(in_c.device))
        _capi.check(rc, "ultra_rspmm_forward_pna")
        ctx.plan, ctx.mul_id, ctx.num_edge = plan, mul_id, edge_index.size(1)
        ctx.save_for_backward(edge_weight, rel_c, in_c, out4)
        return out4[0], out4[1], out4[2], out4[3]

    @staticmethod
    def backward(ctx, g_sum, g_sq, g_max, g_min):
        edge_weight, rel, x, out4 = ctx.saved_tensors
        plan = ctx.plan
        plan.bind_weight(edge_weight)
        need_wg = ctx.needs_input_grad[2]
        z = lambda g: torch.zeros_like(x) if g is None else g.contiguous()
        ADD, MIN, MAX = _capi.SUM_IDS["add"], _capi.SUM_IDS["min"], _capi.SUM_IDS["max"]
        run = lambda s, r, i, o, g: _run_backward(plan, s, ctx.mul_id, r, i, o, g, need_wg, ctx.num_edge)
        wg1, rg1, ig1 = run(ADD, rel, x, out4[0], z(g_sum))
        rel2, x2 = rel * rel, x * x
        wg2, rg2, ig2 = run(ADD, rel2, x2, out4[1], z(g_sq))            # d/d(rel**2), d/d(in**2): chain rule below
        wg3, rg3, ig3 = run(MAX, rel, x, out4[2], z(g_max))
        wg4, rg4, ig4 = run(MIN, rel, x, out4[3], z(g_min))
        relation_grad = rg1 + 2 * rel * rg2 + rg3 + rg4
        input_grad = ig1 + 2 * x * ig2 + ig3 + ig4
        weight_grad = (wg1 + wg2 + wg3 + wg4) if need_wg else None
        return None, None, weight_grad, relation_grad, input_grad, None


def generalized_rspmm_pna(edge_index, edge_type, edge_weight, relation, input, mul="mul"):
    """-> (sum, sq_sum, max, min), each (N, D): the four aggregates of layers.py:211-215 in one gather pass."""
    if mul not in _capi.MUL_IDS:
        raise ValueError("No generalized rspmm implementation found for summation `pna` and multiplication `%s`" % mul)
    return RSPMMPnaFunction.apply(edge_index, edge_type, edge_weight, relation, input, mul)
