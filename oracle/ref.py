"""TEST INFRASTRUCTURE ONLY - loader for the unmodified reference CPU extension (oracle/_ref/rspmm_ref.so).

The module exposes the reference's 12 pybind functions ``rspmm_{add,min,max}_{mul,add}_{forward,backward}_cpu``
(/root/reference/ultra/rspmm/source/rspmm.cpp:256-268) exactly as the reference's own
``ultra/rspmm/rspmm.py`` sees them.  It is compiled in the build container by ``oracle/build_ref.py``; on the
GPU box (where /root/reference does not exist) the prebuilt ``.so`` is imported by path.
"""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "rspmm_ref.so")
_mod = None


def available() -> bool:
    return os.path.exists(_SO)


def load_ref():
    """Import oracle/_ref/rspmm_ref.so (needs torch's libs loaded first)."""
    global _mod
    if _mod is None:
        import torch  # noqa: F401  (libtorch symbols for the pybind module)

        if not available():
            from . import build_ref

            if build_ref.build() is None:
                raise RuntimeError("oracle/_ref/rspmm_ref.so missing and /root/reference not present to build it")
        spec = importlib.util.spec_from_file_location("rspmm_ref", _SO)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def ref_forward(sum, mul, edge_index, edge_type, edge_weight, relation, input):
    """Reference CPU kernel call on pre-sorted CPU torch tensors (what RSPMM*Function.forward dispatches to)."""
    return getattr(load_ref(), "rspmm_%s_%s_forward_cpu" % (sum, mul))(edge_index, edge_type, edge_weight,
                                                                      relation, input)


def ref_backward(sum, mul, edge_index, edge_type, edge_weight, relation, input, output, output_grad):
    return getattr(load_ref(), "rspmm_%s_%s_backward_cpu" % (sum, mul))(edge_index, edge_type, edge_weight,
                                                                       relation, input, output, output_grad)


def ref_generalized_rspmm(edge_index, edge_type, edge_weight, relation, input, sum="add", mul="mul"):
    """Restatement of the reference wrapper (rspmm.py:167-178) on top of the reference kernels: per-call sort."""
    node_in, node_out = edge_index
    key = node_in * (node_out.max() + 1) + node_out
    order = key.argsort()
    return ref_forward(sum, mul, edge_index[:, order], edge_type[order], edge_weight[order], relation, input)
