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


def _version(t):
    """Version counter of a tensor for cache keys.  Inference tensors do not track one (reading `_version` raises), and
    they cannot be mutated in place outside inference mode either, so a constant is a sound key for them.  Writes
    through `.data` never bump the counter: callers that mutate edge tensors that way must call `plan_cache.clear()`
    (or `GraphPlan.invalidate_weight()`)."""
    return 0 if t.is_inference() else t._version


class GraphPlan:
    """Device-resident sorted CSR views + work items of one relational graph (see include/ultra_rspmm.h)."""

    def __init__(self, edge_index, edge_type, num_node, num_relation):
        if edge_index.device.type != "cuda":
            raise RuntimeError("GraphPlan needs CUDA tensors (ultra_b200 has no CPU implementation of rspmm)")
        if edge_index.dtype != torch.int64 or edge_type.dtype != torch.int64:
            raise RuntimeError("expected scalar type Long for edge_index / edge_type, got %s / %s"
                               % (edge_index.dtype, edge_type.dtype))
        if edge_index.dim() != 2 or edge_index.shape[0] != 2 or edge_type.dim() != 1 or \
                edge_type.shape[0] != edge_index.shape[1] or edge_type.device != edge_index.device:
            raise RuntimeError("GraphPlan: edge_index must be (2, E) and edge_type (E,) on the same device, got %s / %s"
                               % (tuple(edge_index.shape), tuple(edge_type.shape)))
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
        self.sm_limit = 0
        self.masked = False

    def nbytes(self):
        """Device bytes this plan pins (workspace + scratch + host-path staging), for the cache's byte budget."""
        n = self.workspace.numel()
        for name in ("_scratch", "_host_staging"):
            t = getattr(self, name, None)
            if t is not None:
                n += t.numel()
        return n

    def set_sm_limit(self, sm_limit):
        """Size this plan's persistent grids for `sm_limit` SMs (0 = all): leaves SMs to a concurrent collective."""
        _capi.check(_capi.lib().ultra_rspmm_plan_set_sm_limit(self._handle, int(sm_limit)), "ultra_rspmm_plan_set_sm_limit")
        self.sm_limit = int(sm_limit)

    def invalidate_weight(self):
        """Forget which weight tensor is bound (after something else rewrote the plan's device weights, e.g. a CUDA-graph
        replay that captured a bind, or after the caller mutated the weight tensor through `.data`)."""
        self._weight_key = None

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
        if edge_weight.dim() != 1 or edge_weight.shape[0] != self.num_edge or edge_weight.dtype not in _DTYPE_ID:
            raise RuntimeError("bind_weight: edge_weight must be a float32/float64 vector of length num_edge (%d), got %s %s"
                               % (self.num_edge, tuple(edge_weight.shape), edge_weight.dtype))
        if edge_weight.device != self.device:
            raise RuntimeError("bind_weight: edge_weight is on %s, the plan on %s" % (edge_weight.device, self.device))
        key = (edge_weight.data_ptr(), _version(edge_weight), edge_weight.dtype, tuple(edge_weight.shape),
               edge_weight.stride())
        if key == self._weight_key:
            return
        ew = edge_weight.detach().contiguous()
        with torch.cuda.device(self.device):  # the permute kernel must launch on the plan's device, not the current one
            rc = _capi.lib().ultra_rspmm_set_edge_weight(self._handle, ew.data_ptr(), _DTYPE_ID[ew.dtype],
                                                         _stream_ptr(self.device))
        _capi.check(rc, "ultra_rspmm_set_edge_weight")
        self._weight_key = key
        self._weight_ref = edge_weight  # keep the storage alive so the key cannot be recycled

    def apply_edge_mask(self, keep_mask):
        """Re-cut this plan for the sub-graph that keeps edge e iff keep_mask[e] (bool (E,), caller's edge order) - on the
        device, no sort, no host synchronisation, CUDA-graph capturable (ultra_rspmm_plan_apply_mask).  The incremental
        plan for remove_easy_edges (/root/reference/ultra/base_nbfnet.py:54-77): afterwards every rspmm call on this graph's
        (unchanged) edge tensors computes on the masked edge set, exactly like a plan built from `edge_index[:, mask]`.
        `None` restores the base graph.  The mask is plan state: forward and the matching backward must see the same one."""
        if keep_mask is not None and (keep_mask.dtype != torch.bool or keep_mask.shape != (self.num_edge,)
                                      or keep_mask.device != self.device):
            raise RuntimeError("apply_edge_mask: keep_mask must be a bool vector of length num_edge on the plan's device")
        km = None if keep_mask is None else keep_mask.contiguous()
        with torch.cuda.device(self.device):
            rc = _capi.lib().ultra_rspmm_plan_apply_mask(self._handle, None if km is None else km.data_ptr(),
                                                         _stream_ptr(self.device))
        _capi.check(rc, "ultra_rspmm_plan_apply_mask")
        self._mask_ref = km  # the kernels enqueued above read it asynchronously
        self.masked = km is not None

    def edge_mask_weight(self, keep_mask, dtype=torch.float32):
        """A 0/1 edge-weight vector that REMOVES the edges where `keep_mask` is False - the incremental alternative to a
        plan rebuild for the training loop's per-step edge removal (remove_easy_edges,
        /root/reference/ultra/base_nbfnet.py:54-77: `data.edge_index[:, mask]` + a fresh sort inside every rspmm call).
        For the sum aggregate a weight of 0 is exactly an absent edge (y = 0 * (rel (x) in) = 0, acc + 0 == acc), so
        the base graph's plan is reused as is: passing the returned tensor as `edge_weight` costs one permute kernel
        (~10 us, no host sync, CUDA-graph capturable) instead of 3 sorts + a synchronisation per step.  NOT valid for
        min / max aggregation (0 is not their identity) - those need a plan of the masked graph."""
        if keep_mask.dtype != torch.bool or keep_mask.shape != (self.num_edge,) or keep_mask.device != self.device:
            raise RuntimeError("edge_mask_weight: keep_mask must be a bool vector of length num_edge on the plan's device")
        return keep_mask.to(dtype)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _capi.lib().ultra_rspmm_plan_destroy(h)
            except Exception:
                pass
            self._handle = None


class PlanCache:
    """LRU of GraphPlans keyed on tensor identity + version (holds the tensors so addresses are not recycled).

    Bounded by entries AND by device bytes: a caller that hands a fresh edge tensor to every call (the reference's
    training loop after remove_easy_edges, /root/reference/ultra/base_nbfnet.py:54-77) misses every time, and each
    stale plan pins its workspace; the byte budget (default 1 GiB, ULTRA_PLAN_CACHE_MB) evicts them early.  Cache hits
    need stable edge tensors - keep `edge_index` / `edge_type` objects alive across layers and steps."""

    def __init__(self, capacity=8, max_bytes=None):
        import os
        self.capacity = capacity
        self.max_bytes = int(os.environ.get("ULTRA_PLAN_CACHE_MB", "1024")) << 20 if max_bytes is None else max_bytes
        self._entries = collections.OrderedDict()
        self._lock = threading.Lock()
        self.hits = 0
        self.misses = 0

    @staticmethod
    def _key(edge_index, edge_type, num_node, num_relation):
        return (edge_index.data_ptr(), _version(edge_index), tuple(edge_index.shape), edge_index.stride(),
                edge_type.data_ptr(), _version(edge_type), tuple(edge_type.shape), edge_type.stride(),
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
            while len(self._entries) > 1 and (len(self._entries) > self.capacity or
                                              sum(e[0].nbytes() for e in self._entries.values()) > self.max_bytes):
                self._entries.popitem(last=False)
        return plan

    def clear(self):
        with self._lock:
            self._entries.clear()


plan_cache = PlanCache()
