"""Graph plan cache: one plan per (edge_index, edge_type) graph, reused by every layer and by forward/backward.

Replaces the work the reference redoes on EVERY rspmm call (12-24x per Ultra.forward over the same graph):
``key.argsort()`` + gathers (/root/reference/ultra/rspmm/rspmm.py:174-178), the sortedness assert with its
device->host sync (rspmm.py:15-17) and ``ind2ptr`` (/root/reference/ultra/rspmm/source/rspmm.cpp:40-48).
"""
import collections
import ctypes
import threading

import torch

from . import _capi

_DTYPE_ID = {torch.float32: _capi.F32, torch.float64: _capi.F64}


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class GraphPlan:
    """Device-resident sorted CSR views + work items of one relational graph (see include/ultra_rspmm.h)."""

    def __init__(self, edge_index, edge_type, num_node, num_relation):
        if edge_index.device.type != "cuda":
            raise RuntimeError("GraphPlan needs CUDA tensors (ultra_b200 has no CPU implementation of rspmm)")
        lib = _capi.lib()
        self.device = edge_index.device
        self.num_node, self.num_relation = int(num_node), int(num_relation)
        self.num_edge = int(edge_index.shape[1])
        ei = edge_index.contiguous()
        et = edge_type.contiguous()
        with torch.cuda.device(self.device):
            nbytes = lib.ultra_rspmm_plan_workspace_bytes(self.num_node, self.num_edge, self.num_relation)
            self.workspace = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            handle = ctypes.c_void_p()
            rc = lib.ultra_rspmm_plan_build(ctypes.byref(handle), ei.data_ptr(), et.data_ptr(), self.num_node,
                                            self.num_edge, self.num_relation, self.workspace.data_ptr(),
                                            self.workspace.numel(), _stream_ptr(self.device))
        if rc != 0:
            msg = lib.ultra_rspmm_last_error().decode()
            if rc == 3:
                raise IndexError(msg)
            raise RuntimeError("ultra_rspmm_plan_build failed (code %d): %s" % (rc, msg))
        self._handle = handle
        self.input_sorted = bool(lib.ultra_rspmm_plan_input_sorted(handle))
        self._scratch = None
        self._weight_key = None
        self._lock = threading.Lock()

    @property
    def handle(self):
        return self._handle

    def info(self, what):
        return int(_capi.lib().ultra_rspmm_plan_info(self._handle, what))

    def scratch(self, dim, dtype):
        """Split-row partial buffer, grown on demand and reused across calls (stream-ordered)."""
        need = _capi.lib().ultra_rspmm_scratch_bytes(self._handle, int(dim), _DTYPE_ID[dtype])
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
        return self._scratch

    def bind_weight(self, edge_weight):
        """Permute edge weights into the plan orders (no host sync); skipped if the same tensor is re-bound."""
        key = (edge_weight.data_ptr(), edge_weight._version, edge_weight.dtype, tuple(edge_weight.shape),
               edge_weight.stride())
        if key == self._weight_key:
            return
        ew = edge_weight.detach().contiguous()
        rc = _capi.lib().ultra_rspmm_set_edge_weight(self._handle, ew.data_ptr(), _DTYPE_ID[ew.dtype],
                                                     _stream_ptr(self.device))
        _capi.check(rc, "ultra_rspmm_set_edge_weight")
        self._weight_key = key
        self._weight_ref = edge_weight  # keep the storage alive so the key cannot be recycled

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _capi.lib().ultra_rspmm_plan_destroy(h)
            except Exception:
                pass
            self._handle = None


class PlanCache:
    """LRU of GraphPlans keyed on tensor identity + version (holds the tensors so addresses are not recycled)."""

    def __init__(self, capacity=16):
        self.capacity = capacity
        self._entries = collections.OrderedDict()
        self._lock = threading.Lock()
        self.hits = 0
        self.misses = 0

    @staticmethod
    def _key(edge_index, edge_type, num_node, num_relation):
        return (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape), edge_index.stride(),
                edge_type.data_ptr(), edge_type._version, tuple(edge_type.shape), edge_type.stride(),
                str(edge_index.device), int(num_node), int(num_relation))

    def get(self, edge_index, edge_type, num_node, num_relation, device=None):
        """`device` is only needed for host-resident edge tensors (their plan lives on that GPU)."""
        key = self._key(edge_index, edge_type, num_node, num_relation) + (str(device),)
        with self._lock:
            hit = self._entries.get(key)
            if hit is not None:
                self._entries.move_to_end(key)
                self.hits += 1
                return hit[0]
        if edge_index.device.type != "cuda":
            plan = GraphPlan(edge_index.to(device), edge_type.to(device), num_node, num_relation)
        else:
            plan = GraphPlan(edge_index, edge_type, num_node, num_relation)
        with self._lock:
            self.misses += 1
            self._entries[key] = (plan, edge_index, edge_type)
            while len(self._entries) > self.capacity:
                self._entries.popitem(last=False)
        return plan

    def clear(self):
        with self._lock:
            self._entries.clear()


plan_cache = PlanCache()
