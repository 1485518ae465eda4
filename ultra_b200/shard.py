"""Head-sharded execution of the hot path across the GPUs of one box.

Columns of the (N, B*d) state are independent inside rspmm, and heads (blocks of d=64 columns) stay independent
through the whole layer and the whole 6-layer stack (Linear / LayerNorm act on the last dim,
/root/reference/ultra/layers.py:233-240).  So rank g of G owns heads [g*B/G, (g+1)*B/G): rspmm, the layer and the
stack need NO data-path collective.  The one collective the north star prescribes is an all-gather of the per-layer
node states of the entity-graph pass (NCCL over NVLink); it is issued on a side stream right after each layer and
overlaps the next layer's kernels, which only touch local heads.

The reference scales the same axis with batch-sharded DDP (/root/reference/script/run.py:33,45); its only
collectives are DDP's gradient all-reduce and metric all-reduces.
"""
import torch
import torch.distributed as dist


def head_range(batch_size, rank, world_size):
    """Contiguous, balanced split of `batch_size` heads: the first (batch_size % world_size) ranks get one extra."""
    base, extra = divmod(batch_size, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_heads(tensor, rank, world_size, dim=0):
    """Slice the head (batch) dimension of a (B, ...) tensor for this rank."""
    lo, hi = head_range(tensor.shape[dim], rank, world_size)
    return tensor.narrow(dim, lo, hi - lo)


def all_gather_heads(local, batch_size, group=None, out=None, async_op=False):
    """Gather (B_local, N, d) node states of every rank into (B, N, d), heads in rank order.

    Equal shards use one all_gather_into_tensor straight into the output (the (B, N, d) layout is head-major, so
    rank blocks are contiguous).  Ragged shards (B % G != 0) are padded to ceil(B/G) heads for the collective
    (NCCL/gloo all-gathers need equal sizes) and compacted afterwards (synchronous)."""
    world = dist.get_world_size(group)
    if out is None:
        out = local.new_empty((batch_size,) + tuple(local.shape[1:]))
    local = local.contiguous()
    if batch_size % world == 0:
        work = dist.all_gather_into_tensor(out, local, group=group, async_op=async_op)
        return out, work
    per = -(-batch_size // world)
    padded = local.new_zeros((per,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    buf = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(buf, padded, group=group)
    for r in range(world):
        lo, hi = head_range(batch_size, r, world)
        out[lo:hi] = buf[r * per:r * per + (hi - lo)]
    return out, None


class OverlappedStateGather:
    """Per-layer all-gather of entity-graph node states on a side stream (CUDA) so that it overlaps the next layer.

    Usage (entity-graph pass only):
        g = OverlappedStateGather(batch_size)
        hidden = bellmanford_entity(layers, ..., hook=g.after_layer)     # launches one all-gather per layer
        states = g.wait()                                                # list of (B, N, d) tensors, one per layer
    """

    def __init__(self, batch_size, group=None, layers=None, reserve_sms=8, plan=None):
        """reserve_sms: SMs left free for NCCL while gathers are in flight (the rspmm / update kernels are persistent
        and otherwise fill every SM, which would serialise the collective behind them).  The reservation is a field of
        the GraphPlan the layers run on (`plan`, per plan, no process-global state); restored by wait()."""
        self.batch_size = batch_size
        self.reserve_sms = reserve_sms
        self.plan = plan
        self._limited = False
        self.group = group
        self.layers = None if layers is None else set(layers)  # which layer indices to gather (None = every layer)
        self.results = []
        self._pending = []
        self._stream = None

    def after_layer(self, layer_idx, hidden):
        if self.layers is not None and layer_idx not in self.layers:
            return
        if hidden.is_cuda:
            if self.reserve_sms and self.plan is not None and not self._limited:
                sms = self.plan.info(7)
                self.plan.set_sm_limit(max(sms - self.reserve_sms, 1))
                self._limited = True
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=hidden.device)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(hidden.device))
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(ready)
                out, _ = all_gather_heads(hidden, self.batch_size, group=self.group)
                hidden.record_stream(self._stream)
                done = torch.cuda.Event()
                done.record(self._stream)
            self._pending.append((out, done))
        else:  # gloo / CPU tensors: synchronous
            out, _ = all_gather_heads(hidden, self.batch_size, group=self.group)
            self._pending.append((out, None))

    def wait(self):
        for out, done in self._pending:
            if done is not None:
                cur = torch.cuda.current_stream(out.device)
                cur.wait_event(done)
                out.record_stream(cur)  # allocated on the side stream, consumed on this one
            self.results.append(out)
        self._pending = []
        if self._limited:
            self.plan.set_sm_limit(0)
            self._limited = False
        return self.results


class ShardedUltra:
    """Head-sharded Ultra forward: rank g of G owns the queries (heads) [g*B/G, (g+1)*B/G) of the batch.

    Heads ARE the batch samples (one (h, r) query each), so this is the reference's own scaling axis - batch-sharded DDP
    (/root/reference/script/run.py:33,45) - inside one process group: both NBFNet stacks and the scoring MLP run on local
    heads with no communication, and ONE all-gather of the (B_local, N) score rows gives every rank what the caller of
    `Ultra.forward` consumes (models.py:208; at C3: 256 x 14 541 x 4 B = 14.9 MB in total).  Nothing downstream in the
    reference reads replicated per-layer node states (models.py:157-170 uses the last hidden + query, then gathers tails,
    :202-208), which is why the per-layer full-state gather of round 1 (953 MB per rank per layer at C3 - more bytes than
    the layer itself moves) is no longer on the default path; `OverlappedStateGather` remains for callers that want
    replicated states (path visualisation, base_nbfnet.py:156-172)."""

    def __init__(self, model, group=None):
        self.model = model
        self.group = group

    def local_scores(self, rel_graph, graph, h_index, r_index):
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        lo, hi = head_range(len(h_index), rank, world)
        return self.model(rel_graph, graph, h_index[lo:hi], r_index[lo:hi])

    def gather_scores(self, local, batch_size, out=None):
        out, _ = all_gather_heads(local, batch_size, group=self.group, out=out)
        return out

    def __call__(self, rel_graph, graph, h_index, r_index):
        """h_index, r_index: the FULL batch (B,) on every rank.  Returns the replicated (B, N) score matrix."""
        return self.gather_scores(self.local_scores(rel_graph, graph, h_index, r_index), len(h_index))
