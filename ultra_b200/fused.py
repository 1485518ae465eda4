"""Head-major rspmm and the fused layer epilogue - the callers' side of the hot path.

    rspmm_heads(edge_index, edge_type, edge_weight, relation (B,R,d), input (B,N,d), boundary (B,N,d)|None, sum, mul)
        == generalized_rspmm on the transposed-and-flattened operands of /root/reference/ultra/layers.py:190-192,
           folded with the boundary term (layers.py:199-207) and transposed back (layers.py:230) - without any of
           those copies: the kernels address the (B, rows, d) layout directly.
    nbfnet_update(x, u, linear, layer_norm, relu, shortcut)
        == relu(layer_norm(linear(cat[x, u]))) (+ x)    layers.py:233-240 + models.py:158-160, one kernel, inference.
"""
import torch
from torch import autograd

from . import _capi
from .plan import plan_cache, _DTYPE_ID, _stream_ptr


def heads_layout_ok(input):
    """The head-major kernels need head_dim * itemsize to be a multiple of the 128-byte column tile."""
    return (input.is_cuda and input.dtype in _DTYPE_ID and input.dim() == 3
            and (input.shape[-1] * input.element_size()) % 128 == 0)


class _RSPMMHeadsFunction(autograd.Function):

    @staticmethod
    def forward(ctx, edge_index, edge_type, edge_weight, relation, input, boundary, sum, mul):
        lib = _capi.lib()
        sum_id, mul_id = _capi.SUM_IDS[sum], _capi.MUL_IDS[mul]
        B, N, d = input.shape
        # relation (B, R, d), or (R, d): ONE block shared by every head (the relation-graph stack, layers.py:76) that
        # the kernels read with head stride 0 - the (B, R, d) expand is never materialised
        broadcast = relation.dim() == 2
        R = relation.shape[-2]
        if relation.shape[-1] != d or (not broadcast and (relation.dim() != 3 or relation.shape[0] != B)):
            raise RuntimeError("rspmm_heads: relation must be (B, R, d) or (R, d) matching input (B, N, d)")
        if boundary is not None and boundary.shape != input.shape:
            raise RuntimeError("rspmm_heads: boundary must have the shape of input")
        if relation.dtype != input.dtype or edge_weight.dtype != input.dtype or \
                (boundary is not None and boundary.dtype != input.dtype):
            raise RuntimeError("Expected tensors 'edge_weight', 'relation', 'input' to have the same type; got %s, %s, %s"
                               % (edge_weight.dtype, relation.dtype, input.dtype))
        devs = {t.device for t in (edge_index, edge_type, edge_weight, relation, input)}
        if len(devs) != 1:
            raise RuntimeError("Expected all tensors to be on the same device, got %s" % sorted(map(str, devs)))
        # index dtype / shapes are validated by GraphPlan, the weight length by bind_weight
        plan = plan_cache.get(edge_index, edge_type, N, R)
        plan.bind_weight(edge_weight)
        rel_c, in_c = relation.detach().contiguous(), input.detach().contiguous()
        bnd_c = None if boundary is None else boundary.detach().contiguous()
        output = torch.empty_like(in_c)
        scratch = plan.scratch(B * d, input.dtype)
        with torch.cuda.device(input.device):
            rc = lib.ultra_rspmm_forward_heads(plan.handle, sum_id, mul_id, _DTYPE_ID[input.dtype], rel_c.data_ptr(),
                                               in_c.data_ptr(), None if bnd_c is None else bnd_c.data_ptr(),
                                               output.data_ptr(), B, d, 1 if broadcast else 0, scratch.data_ptr(),
                                               scratch.numel(), _stream_ptr(input.device))
        _capi.check(rc, "ultra_rspmm_forward_heads")
        ctx.plan, ctx.ids, ctx.sum, ctx.broadcast = plan, (sum_id, mul_id), sum, broadcast
        ctx.has_boundary = boundary is not None
        ctx.save_for_backward(edge_weight, rel_c, in_c, output, bnd_c if bnd_c is not None else torch.empty(0))
        return output

    @staticmethod
    def backward(ctx, output_grad):
        if ctx.needs_input_grad[2]:
            raise RuntimeError("rspmm_heads: edge_weight gradients need the flat path (generalized_rspmm)")
        lib = _capi.lib()
        edge_weight, rel_c, in_c, output, bnd_c = ctx.saved_tensors
        plan = ctx.plan
        plan.bind_weight(edge_weight)
        B, N, d = in_c.shape
        og = output_grad.contiguous()
        relation_grad = in_c.new_empty((B,) + tuple(rel_c.shape[-2:]))
        input_grad = torch.empty_like(in_c)
        scratch = plan.scratch(B * d, in_c.dtype)
        with torch.cuda.device(in_c.device):
            rc = lib.ultra_rspmm_backward_heads(plan.handle, ctx.ids[0], ctx.ids[1], _DTYPE_ID[in_c.dtype],
                                                rel_c.data_ptr(), in_c.data_ptr(), output.data_ptr(), og.data_ptr(),
                                                relation_grad.data_ptr(), input_grad.data_ptr(), B, d,
                                                1 if ctx.broadcast else 0, scratch.data_ptr(), scratch.numel(),
                                                _stream_ptr(in_c.device))
        _capi.check(rc, "ultra_rspmm_backward_heads")
        if ctx.broadcast:  # d/d(shared embedding) = sum over the heads that read it (what autograd does for expand)
            relation_grad = relation_grad.sum(0)
        boundary_grad = None
        if ctx.has_boundary and ctx.needs_input_grad[5]:
            # out = Nary(agg, boundary): d out / d boundary = 1 (sum) or [out == boundary] (min / max, operator.cuh:62,75)
            boundary_grad = og if ctx.sum == "add" else og * (output == bnd_c).to(og.dtype)
        return None, None, None, relation_grad, input_grad, boundary_grad, None, None


def rspmm_heads(edge_index, edge_type, edge_weight, relation, input, boundary=None, sum="add", mul="mul"):
    if sum not in _capi.SUM_IDS or mul not in _capi.MUL_IDS:
        raise ValueError("No generalized rspmm implementation found for summation `%s` and multiplication `%s`"
                         % (sum, mul))
    if not heads_layout_ok(input):
        raise RuntimeError("rspmm_heads needs CUDA fp32/fp64 (B, N, d) input with d * itemsize % 128 == 0")
    return _RSPMMHeadsFunction.apply(edge_index, edge_type, edge_weight, relation, input, boundary, sum, mul)


def relation_project_ok(rel_repr, projections):
    """All `projections` are Sequential(Linear(64,64), ReLU, Linear(64,64)) on fp32 CUDA (B, R, 64); inference only."""
    if not (rel_repr.is_cuda and rel_repr.dtype == torch.float32 and rel_repr.shape[-1] == 64
            and not torch.is_grad_enabled()):
        return False
    for p in projections:
        if not (len(p) == 3 and isinstance(p[0], torch.nn.Linear) and isinstance(p[1], torch.nn.ReLU)
                and isinstance(p[2], torch.nn.Linear) and p[0].weight.shape == (64, 64) and p[2].weight.shape == (64, 64)
                and p[0].bias is not None and p[2].bias is not None and p[0].weight.dtype == torch.float32):
            return False
    return True


def relation_project(rel_repr, projections):
    """Every layer's relation_projection MLP of the same (B, R, 64) relation representations in ONE launch
    (layers.py:60-64,80).  Returns (L, B, R, 64)."""
    lib = _capi.lib()
    x = rel_repr.contiguous()
    L = len(projections)
    with torch.no_grad():
        w1 = torch.stack([p[0].weight.t() for p in projections]).contiguous()   # (L, 64, 64) [k][n]
        b1 = torch.stack([p[0].bias for p in projections]).contiguous()
        w2 = torch.stack([p[2].weight.t() for p in projections]).contiguous()
        b2 = torch.stack([p[2].bias for p in projections]).contiguous()
    out = x.new_empty((L,) + tuple(x.shape))
    rows = x.numel() // 64
    with torch.cuda.device(x.device):
        rc = lib.ultra_relation_project_f32(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                            out.data_ptr(), rows, 64, L, _stream_ptr(x.device))
    _capi.check(rc, "ultra_relation_project_f32")
    return out


def score_head_ok(h, mlp):
    """`mlp` is Sequential(Linear(128,128), ReLU, Linear(128,1)) (models.py:114-118 with feature_dim 128) on fp32 CUDA
    states (B, N, 64); inference only (the kernel has no backward: under autograd the caller keeps its eager MLP)."""
    return (h.is_cuda and h.dtype == torch.float32 and h.dim() == 3 and h.shape[-1] == 64 and not torch.is_grad_enabled()
            and len(mlp) == 3 and isinstance(mlp[0], torch.nn.Linear) and isinstance(mlp[1], torch.nn.ReLU)
            and isinstance(mlp[2], torch.nn.Linear) and mlp[0].weight.shape == (128, 128) and mlp[2].weight.shape == (1, 128)
            and mlp[0].bias is not None and mlp[2].bias is not None and mlp[0].weight.dtype == torch.float32)


def score_head(h, query, mlp):
    """Scores of every entity as the tail of every query, (B, N), in one pass over the final states h (B, N, 64):
    mlp(cat[h, query]) of models.py:166-170,202-208 with the first Linear split into its h and query halves."""
    lib = _capi.lib()
    B, N, d = h.shape
    h = h.contiguous()
    with torch.no_grad():
        w1 = mlp[0].weight
        qb = torch.addmm(mlp[0].bias, query, w1[:, d:].t()).contiguous()     # (B, 128): query half + bias
        w1h_t = w1[:, :d].t().contiguous()                                    # (64, 128) [k][n]
        w2 = mlp[2].weight.reshape(-1).contiguous()
        b2 = mlp[2].bias.contiguous()
    out = h.new_empty(B, N)
    with torch.cuda.device(h.device):
        rc = lib.ultra_score_head_f32(h.data_ptr(), w1h_t.data_ptr(), qb.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                      out.data_ptr(), B * N, N, d, w1.shape[0], _stream_ptr(h.device))
    _capi.check(rc, "ultra_score_head_f32")
    return out


def nbfnet_update_ok(x, u, linear, activation, layer_norm=None):
    """Shapes / dtypes the fused epilogue covers.  With autograd recording, the training variant (which needs LayerNorm
    and the Linear bias) is used; otherwise the inference kernel."""
    import torch.nn.functional as F
    ok = (x.is_cuda and x.dtype == torch.float32 and u.dtype == torch.float32 and x.shape == u.shape
          and x.shape[-1] == 64 and linear.in_features == 128 and linear.out_features == 64
          and linear.weight.dtype == torch.float32 and (activation is None or activation is F.relu))
    if ok and torch.is_grad_enabled():
        ok = layer_norm is not None and layer_norm.elementwise_affine and linear.bias is not None
    return ok


class _NBFNetUpdateFunction(autograd.Function):
    """out = (shortcut ? x : 0) + relu?( LayerNorm( [x ; u] . W^T + b ) )  with a hand-written forward AND backward
    (layers.py:233-240 + models.py:158-160).  Forward: the tcgen05 epilogue kernel, which also stores z (the Linear
    output) and the LayerNorm row statistics.  Backward: ONE kernel for ReLU + LayerNorm (-> dz, d gamma, d beta, d bias),
    then the GEMM-shaped gradients as library GEMMs: d[x;u] = dz . W (two GEMMs, the residual folded into the first) and
    dW = dz^T . [x;u] as a head-batched GEMM summed over heads."""

    @staticmethod
    def forward(ctx, x, u, weight, bias, gamma, beta, eps, relu, shortcut, sm_limit):
        lib = _capi.lib()
        xc, uc = x.detach().contiguous(), u.detach().contiguous()
        out = torch.empty_like(xc)
        z = torch.empty_like(xc)
        rows = xc.numel() // 64
        stats = xc.new_empty(rows, 2)
        wt = weight.detach().t().contiguous()
        with torch.cuda.device(x.device):
            rc = lib.ultra_nbfnet_update_train_f32(xc.data_ptr(), uc.data_ptr(), wt.data_ptr(), bias.detach().data_ptr(),
                                                   gamma.detach().data_ptr(), beta.detach().data_ptr(), out.data_ptr(),
                                                   z.data_ptr(), stats.data_ptr(), rows, 128, 64, float(eps),
                                                   1 if relu else 0, 1 if shortcut else 0, int(sm_limit),
                                                   _stream_ptr(x.device))
        _capi.check(rc, "ultra_nbfnet_update_train_f32")
        ctx.save_for_backward(xc, uc, weight, gamma, beta, z, stats)
        ctx.relu, ctx.shortcut = relu, shortcut
        return out

    @staticmethod
    def backward(ctx, out_grad):
        lib = _capi.lib()
        xc, uc, weight, gamma, beta, z, stats = ctx.saved_tensors
        og = out_grad.contiguous()
        rows = og.numel() // 64
        dz = torch.empty_like(og)
        grid = lib.ultra_nbfnet_update_bwd_grid(rows)
        partials = og.new_empty(grid, 3, 64)
        with torch.cuda.device(og.device):
            rc = lib.ultra_nbfnet_update_bwd_f32(og.data_ptr(), z.data_ptr(), stats.data_ptr(), gamma.detach().data_ptr(),
                                                 beta.detach().data_ptr(), dz.data_ptr(), partials.data_ptr(), rows, 64,
                                                 1 if ctx.relu else 0, _stream_ptr(og.device))
        _capi.check(rc, "ultra_nbfnet_update_bwd_f32")
        sums = partials.sum(0)  # fixed order over the CTAs: deterministic
        dgamma, dbeta, dbias = sums[0], sums[1], sums[2]
        dz2, og2 = dz.view(rows, 64), og.view(rows, 64)
        w = weight.detach()
        dx = du = dw = None
        if ctx.needs_input_grad[0]:
            dx = (torch.addmm(og2, dz2, w[:, :64]) if ctx.shortcut else torch.mm(dz2, w[:, :64])).view_as(xc)
        if ctx.needs_input_grad[1]:
            du = torch.mm(dz2, w[:, 64:]).view_as(uc)
        if ctx.needs_input_grad[2]:
            # dW = dz^T . [x ; u]: reduce over rows head by head (batched GEMM over the leading dim), then over the heads
            lead = xc.shape[0] if xc.dim() == 3 else 1
            dz3 = dz.view(lead, -1, 64).transpose(1, 2)
            dw = torch.cat([torch.bmm(dz3, xc.view(lead, -1, 64)).sum(0), torch.bmm(dz3, uc.view(lead, -1, 64)).sum(0)], 1)
        return dx, du, dw, dbias, dgamma, dbeta, None, None, None, None


def nbfnet_update(x, u, linear, layer_norm=None, relu=True, shortcut=False, sm_limit=0):
    """One-pass layer epilogue.  x, u: (..., 64) fp32 CUDA; returns a tensor of x's shape.  Differentiable when autograd
    is recording (training variant, needs layer_norm).
    sm_limit > 0 launches at most that many persistent CTAs (SMs left to a concurrent collective)."""
    if torch.is_grad_enabled() and (x.requires_grad or u.requires_grad or linear.weight.requires_grad):
        return _NBFNetUpdateFunction.apply(x, u, linear.weight, linear.bias, layer_norm.weight, layer_norm.bias,
                                           layer_norm.eps, relu, shortcut, sm_limit)
    return _nbfnet_update_inference(x, u, linear, layer_norm, relu, shortcut, sm_limit)


def _nbfnet_update_inference(x, u, linear, layer_norm=None, relu=True, shortcut=False, sm_limit=0):
    lib = _capi.lib()
    xc, uc = x.contiguous(), u.contiguous()
    out = torch.empty_like(xc)
    wt = linear.weight.detach().t().contiguous()  # (128, 64); 32 KB
    bias = linear.bias
    gamma = layer_norm.weight if layer_norm is not None else None
    beta = layer_norm.bias if layer_norm is not None else None
    ptr = lambda t: None if t is None else t.detach().data_ptr()
    rows = xc.numel() // 64
    with torch.cuda.device(x.device):
        rc = lib.ultra_nbfnet_update_f32(xc.data_ptr(), uc.data_ptr(), wt.data_ptr(), ptr(bias), ptr(gamma), ptr(beta),
                                         out.data_ptr(), rows, 128, 64,
                                         float(layer_norm.eps) if layer_norm is not None else 0.0,
                                         1 if layer_norm is not None else 0, 1 if relu else 0, 1 if shortcut else 0,
                                         int(sm_limit), _stream_ptr(x.device))
    _capi.check(rc, "ultra_nbfnet_update_f32")
    return out
