"""GPU parity of the fused scoring MLP (ultra_score_head_f32) against the reference's
formulation  mlp(cat[h, query])  of /root/reference/ultra/models.py:166-170,202-208 evaluated by PyTorch in fp32.
Tolerance: 1e-5 of the largest |score| (fp32 accumulation in another order), the bar of the path."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def make_mlp(device, seed):
    torch.manual_seed(seed)
    mlp = torch.nn.Sequential(torch.nn.Linear(128, 128), torch.nn.ReLU(), torch.nn.Linear(128, 1)).to(device)
    return mlp


def reference_scores(h, query, mlp):
    feature = torch.cat([h, query.unsqueeze(1).expand(-1, h.shape[1], -1)], dim=-1)   # models.py:166-170
    return mlp(feature).squeeze(-1)                                                     # models.py:202-208


@pytest.mark.parametrize("B,N", [(1, 1), (3, 7), (2, 128), (5, 129), (4, 1000), (16, 14541), (7, 333)])
def test_score_head_matches_eager_mlp(cuda_device, B, N):
    from ultra_b200 import fused
    mlp = make_mlp(cuda_device, 100 + B)
    g = torch.Generator(device=cuda_device).manual_seed(B * 1000 + N)
    h = torch.randn(B, N, 64, device=cuda_device, generator=g) * 2.0
    query = torch.randn(B, 64, device=cuda_device, generator=g)
    with torch.no_grad():
        assert fused.score_head_ok(h, mlp)
        got = fused.score_head(h, query, mlp)
        ref = reference_scores(h, query, mlp)
    assert got.shape == (B, N)
    tol = 1e-5 * float(ref.abs().max()) + 1e-6
    assert float((got - ref).abs().max()) <= tol


def test_score_head_noncontiguous_states_and_graph_capture(cuda_device):
    """h arrives as a view (the layer stack's (B, N, d) state may be a slice); the call is stream-ordered without a host
    read (b2 stays on the device), so it can be captured in a CUDA graph like the rest of the forward."""
    from ultra_b200 import fused
    mlp = make_mlp(cuda_device, 7)
    g = torch.Generator(device=cuda_device).manual_seed(5)
    big = torch.randn(6, 300, 96, device=cuda_device, generator=g)
    h = big[:, :, 16:80]
    query = torch.randn(6, 64, device=cuda_device, generator=g)
    with torch.no_grad():
        ref = reference_scores(h, query, mlp)
        got = fused.score_head(h, query, mlp)
        assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-6
        hc = h.contiguous()
        side = torch.cuda.Stream(device=cuda_device)
        side.wait_stream(torch.cuda.current_stream(cuda_device))
        with torch.cuda.stream(side):
            fused.score_head(hc, query, mlp)
        torch.cuda.current_stream(cuda_device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = fused.score_head(hc, query, mlp)
        hc.mul_(0.5)
        graph.replay()
        torch.cuda.synchronize(cuda_device)
        assert float((out - reference_scores(hc, query, mlp)).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-6


def test_score_head_used_by_the_model_only_without_autograd(cuda_device):
    from ultra_b200 import fused
    mlp = make_mlp(cuda_device, 9)
    h = torch.randn(2, 50, 64, device=cuda_device)
    assert not fused.score_head_ok(h, mlp)                      # grad mode on: the eager MLP keeps autograd
    with torch.no_grad():
        assert fused.score_head_ok(h, mlp)
        assert not fused.score_head_ok(h.double(), mlp)
        assert not fused.score_head_ok(torch.randn(2, 50, 32, device=cuda_device), mlp)
        small = torch.nn.Sequential(torch.nn.Linear(128, 64), torch.nn.ReLU(), torch.nn.Linear(64, 1)).to(cuda_device)
        assert not fused.score_head_ok(h, small)


def test_score_head_abi_errors(cuda_device):
    from ultra_b200 import _capi
    lib = _capi.lib()
    x = torch.zeros(4, 64, device=cuda_device)
    w = torch.zeros(64, 128, device=cuda_device)
    q = torch.zeros(1, 128, device=cuda_device)
    v = torch.zeros(128, device=cuda_device)
    b = torch.zeros(1, device=cuda_device)
    o = torch.zeros(4, device=cuda_device)
    args = lambda dim, hid, rph, out: (x.data_ptr(), w.data_ptr(), q.data_ptr(), v.data_ptr(), b.data_ptr(), out, 4, rph, dim, hid, None)  # noqa: E731
    assert lib.ultra_score_head_f32(*args(32, 128, 4, o.data_ptr())) == 1 and b"dim == 64" in lib.ultra_rspmm_last_error()
    assert lib.ultra_score_head_f32(*args(64, 64, 4, o.data_ptr())) == 1
    assert lib.ultra_score_head_f32(*args(64, 128, 0, o.data_ptr())) == 1
    assert lib.ultra_score_head_f32(*args(64, 128, 4, None)) == 1
    assert lib.ultra_score_head_f32(x.data_ptr() + 4, w.data_ptr(), q.data_ptr(), v.data_ptr(), b.data_ptr(), o.data_ptr(),
                                    3, 3, 64, 128, None) == 1 and b"aligned" in lib.ultra_rspmm_last_error()
    assert lib.ultra_score_head_f32(*args(64, 128, 4, o.data_ptr())) == 0
    torch.cuda.synchronize(cuda_device)
