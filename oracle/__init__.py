"""TEST INFRASTRUCTURE ONLY - parity oracle for the rspmm hot path.

Two checkers live here, neither is ever on the product path (ultra_b200/ never imports this package):

* ``oracle.c_oracle``  - plain-C scalar restatement (``rspmm_oracle.c``) of
  /root/reference/ultra/rspmm/source/rspmm.cpp:50-119 + operator.cuh:13-80, bound with ctypes.
* ``oracle.ref``       - the UNMODIFIED reference CPU extension, compiled from /root/reference by
  ``oracle/build_ref.py`` into ``oracle/_ref/rspmm_ref.so`` (git-ignored, travels to the GPU box).

Allowed callers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline leg and --impl reference).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_C_SRC = os.path.join(_HERE, "rspmm_oracle.c")
_C_SO = os.path.join(_HERE, "_build", "librspmm_oracle.so")

SUM_IDS = {"add": 0, "min": 1, "max": 2}
MUL_IDS = {"mul": 0, "add": 1}


def build_c(force: bool = False) -> str:
    """gcc -O2 (no fast-math: evaluation order must stay as written) -> oracle/_build/librspmm_oracle.so"""
    if (not force and os.path.exists(_C_SO) and os.path.getmtime(_C_SO) >= os.path.getmtime(_C_SRC)):
        return _C_SO
    os.makedirs(os.path.dirname(_C_SO), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-o", _C_SO, _C_SRC])
    return _C_SO


_lib = None


def _get_lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c())
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(edge_index, edge_type, edge_weight, relation, input):
    edge_index = np.ascontiguousarray(edge_index, dtype=np.int64)
    edge_type = np.ascontiguousarray(edge_type, dtype=np.int64)
    dt = np.asarray(input).dtype
    assert dt in (np.float32, np.float64), dt
    edge_weight = np.ascontiguousarray(edge_weight, dtype=dt)
    relation = np.ascontiguousarray(relation, dtype=dt)
    input = np.ascontiguousarray(input, dtype=dt)
    assert edge_index.ndim == 2 and edge_index.shape[0] == 2
    # the reference asserts sortedness of the row-major key before calling the kernel (rspmm.py:15-17)
    key = edge_index[0] * (edge_index[1].max(initial=0) + 1) + edge_index[1]
    assert (np.diff(key) >= 0).all(), "Expect sorted `edge_index`"
    return edge_index, edge_type, edge_weight, relation, input, ("f32" if dt == np.float32 else "f64")


def sort_edges(edge_index, edge_type, edge_weight):
    """numpy restatement of the per-call sort in generalized_rspmm (rspmm.py:174-178); stable."""
    edge_index = np.asarray(edge_index)
    key = edge_index[0] * (edge_index[1].max(initial=0) + 1) + edge_index[1]
    order = np.argsort(key, kind="stable")
    return edge_index[:, order], np.asarray(edge_type)[order], np.asarray(edge_weight)[order], order


def c_forward(sum, mul, edge_index, edge_type, edge_weight, relation, input):
    """Oracle forward on pre-sorted numpy inputs -> (N, D) numpy array."""
    ei, et, ew, rel, inp, sfx = _prep(edge_index, edge_type, edge_weight, relation, input)
    n, d = inp.shape
    out = np.empty((n, d), dtype=inp.dtype)
    fn = getattr(_get_lib(), "rspmm_oracle_forward_" + sfx)
    rc = fn(ctypes.c_int(SUM_IDS[sum]), ctypes.c_int(MUL_IDS[mul]), _ptr(ei), _ptr(et), _ptr(ew), _ptr(rel),
            _ptr(inp), _ptr(out), ctypes.c_int64(n), ctypes.c_int64(ei.shape[1]), ctypes.c_int64(d))
    assert rc == 0
    return out


def c_backward(sum, mul, edge_index, edge_type, edge_weight, relation, input, output, output_grad):
    """Oracle backward on pre-sorted numpy inputs -> (weight_grad, relation_grad, input_grad)."""
    ei, et, ew, rel, inp, sfx = _prep(edge_index, edge_type, edge_weight, relation, input)
    n, d = inp.shape
    output = np.ascontiguousarray(output, dtype=inp.dtype)
    output_grad = np.ascontiguousarray(output_grad, dtype=inp.dtype)
    wg = np.zeros_like(ew)
    rg = np.zeros_like(rel)
    ig = np.zeros_like(inp)
    fn = getattr(_get_lib(), "rspmm_oracle_backward_" + sfx)
    rc = fn(ctypes.c_int(SUM_IDS[sum]), ctypes.c_int(MUL_IDS[mul]), _ptr(ei), _ptr(et), _ptr(ew), _ptr(rel),
            _ptr(inp), _ptr(output), _ptr(output_grad), _ptr(wg), _ptr(rg), _ptr(ig), ctypes.c_int64(n),
            ctypes.c_int64(ei.shape[1]), ctypes.c_int64(d))
    assert rc == 0
    return wg, rg, ig
