/*
 * ultra_rspmm.h -- C ABI of the Blackwell-native (sm_100a) generalized relational SpMM ("rspmm").
 *
 * This is the drop-in boundary for the ONE hot path of DeepGraphLearning/ULTRA:
 *     out[i,:] = AGG_{e : edge_index[0,e]==i}  w_e * ( relation[type_e,:] (*|+) input[edge_index[1,e],:] )
 * It replaces the 24 pybind11 entry points  rspmm_{add,min,max}_{mul,add}_{forward,backward}_{cpu,cuda}
 * of the reference (/root/reference/ultra/rspmm/source/rspmm.cpp:256-283, declared in rspmm.h:22-105) with
 * a plan-based interface: the per-call argsort + sortedness assert + ind2ptr of the reference
 * (/root/reference/ultra/rspmm/rspmm.py:15-17,174-178; rspmm.cpp:40-48) become ONE plan build per graph.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; ultra_rspmm_last_error() gives the message
 *     (thread-local).  No C++ exceptions cross the ABI.
 *   - the library owns NO device memory: the caller allocates every device buffer (workspace, scratch,
 *     outputs).  Plans are small host objects holding pointers into the caller's workspace.
 *   - all device work is enqueued on the `stream` argument (a cudaStream_t passed as void*); forward/backward
 *     never synchronise the host and are CUDA-graph capturable.  plan_build synchronises `stream` exactly once, at its
 *     end (sorts, CSR construction and the work-item cutter all run on the device).
 *   - a plan is a single-stream object: calls that use the same plan must be stream-ordered with respect to each other
 *     (the plan's workspace holds the bound weights, the work-queue tickets of the persistent kernels and, under an edge
 *     mask, the active edge arrays).  Different plans are independent; there is no process-global launch state.
 *   - dtype: ULTRA_RSPMM_F32 or ULTRA_RSPMM_F64 (the reference dispatches AT_DISPATCH_FLOATING_TYPES,
 *     rspmm.cu:251,313,332).  Indices are int64 at the boundary (reference convention, rspmm.cpp:22),
 *     int32 inside the plan.
 *   - there is NO CPU implementation behind this ABI (north star: "no CPU fallback").
 */
#ifndef ULTRA_RSPMM_H_
#define ULTRA_RSPMM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* NaryOp of the reference (operator.cuh:42-80) */
enum { ULTRA_RSPMM_SUM_ADD = 0, ULTRA_RSPMM_SUM_MIN = 1, ULTRA_RSPMM_SUM_MAX = 2 };
/* BinaryOp of the reference (operator.cuh:13-40): MUL = distmult, ADD = transe */
enum { ULTRA_RSPMM_MUL_MUL = 0, ULTRA_RSPMM_MUL_ADD = 1 };
enum { ULTRA_RSPMM_F32 = 0, ULTRA_RSPMM_F64 = 1 };

typedef struct ultra_rspmm_plan ultra_rspmm_plan; /* opaque host object */

/* Library / build identification ("ultra_rspmm sm_100a <version>"). */
const char *ultra_rspmm_version(void);

/*
 * Process-wide launch tunables (experiment knobs; nothing on the data path depends on them being set).
 * "l2_percent": share of L2 (default 25) the concurrently gathered input tiles may occupy (tile-major schedule).
 * "smem_tiles": highest shared-memory residency mode the segment kernel may choose (2 = relation tile + gathered tile for
 * small graphs, 1 = relation tile only, 0 = none; default 2) - an experiment knob (0 costs +13 % at FB15k-237 scale).
 * "host_copy_split": row blocks per slab copy in forward_host.
 * "work_queue": 1 (default) = the persistent CTAs draw their blocks of work from a global ticket counter, 0 = static
 * round-robin (A/B knob; same results bit for bit).
 * "plan_relation_runs": read by plan_build.  -1 (default) = graphs whose gathered tile and relation tile are both
 * shared-memory resident ((num_node + num_relation) * 128 B <= 78 KB: ULTRA's relation graphs) get their segments'
 * edges sorted by relation type first, and the kernel re-reads a relation slice only when the type changes;
 * 0 = never, 1 = every graph (measured slower on large graphs: the gathers lose their sweep locality).  The order only
 * changes the floating-point summation order inside a segment.
 * "bwd_fused_relgrad": 0 (default) = backward is two deterministic gather passes (input_grad by column, relation_grad
 * by relation); 1 = ONE pass over the by-column order that also adds each edge's relation_grad contribution with a
 * 16-byte vector reduction at L2 (red.global.add.v4.f32), like the reference's atomicAdd backward (rspmm.cu:198-209):
 * relation_grad then depends on the reduction order in its last bits, input_grad / weight_grad stay deterministic.
 * Measured at FB15k-237 scale, 128 heads: add x mul 6.7 -> 3.7 ms (uniform degrees) but 4.9 -> 5.6 ms (Zipf: the
 * reductions collide on hot relations); min/max x mul 14.0 -> 5.5 / 8.6 -> 6.1 ms; add x add is faster two-pass.
 */
int ultra_rspmm_set_tunable(const char *name, int value);

/*
 * Host-only (no GPU needed): run the HOST MIRROR of the plan's device-side work-item cutter on a CSR segment pointer (num_seg + 1 ints) and return
 * items (int4 {e_begin, e_end, j0 | merged_bit(1<<30), slot}), the fix list (int4 {segment, slot_begin, slot_count, 0};
 * the first num_fix_long entries are long segments, the rest empty ones) and the non-empty segment ids.  Exists so
 * that the CPU test-suite can check the partition invariants; not used by the product path.
 */
int ultra_rspmm_debug_cut_items(const int *seg_ptr, int num_seg, int *items_out, int64_t items_cap, int *num_items,
                                int *fix_out, int64_t fix_cap, int *num_fix, int *num_fix_long, int *num_slots,
                                int *nonempty_out, int64_t nonempty_cap, int *num_nonempty);

/*
 * Test hook: copy one array of a built plan to the host (synchronous).  order: 0 by row, 1 by col, 2 by relation;
 * what: 0 items (int4 x plan_info items), 1 fix list (int4), 2 non-empty segment ids (num_seg + 1 ints, entries past
 * the count are unspecified), 3 segment pointer (num_seg + 1 ints), 4 the device-side counts (8 ints: items, fix entries,
 * long-segment fix entries, slots, non-empty entries, merged groups, other items, edges), 5 the caller's edge ids in
 * this order's sequence (num_edge ints; base graph).  The GPU test-suite compares these with the host
 * mirror (ultra_rspmm_debug_cut_items); not used by the product path.
 */
int ultra_rspmm_plan_debug_read(const ultra_rspmm_plan *plan, int order, int what, void *host_out, int64_t cap_bytes);

/* Message of the last failing call on this thread ("" if none). */
const char *ultra_rspmm_last_error(void);

/* Bytes of device workspace a plan for (num_node, num_edge, num_relation) needs (upper bound). */
size_t ultra_rspmm_plan_workspace_bytes(int64_t num_node, int64_t num_edge, int64_t num_relation);

/*
 * Build the graph plan: replaces generalized_rspmm's per-call `key.argsort()` + gathers (rspmm.py:174-178),
 * the `assert (key.diff() >= 0).all()` host sync (rspmm.py:15-17) and ind2ptr (rspmm.cpp:40-48).
 *   edge_index : device int64 (2, num_edge) row-major; [0]=aggregating node (row), [1]=gathered node (col)
 *   edge_type  : device int64 (num_edge)
 *   workspace  : device buffer of >= ultra_rspmm_plan_workspace_bytes() bytes, 256-byte aligned; must stay
 *                alive and untouched for the life of the plan
 * Produces int32 CSR by row (forward), by col (input_grad), by relation (relation_grad), each with its
 * edge-balanced work-item list.  Fails if any index is out of range (UB in the reference, rspmm.cu:234-236).
 */
int ultra_rspmm_plan_build(ultra_rspmm_plan **plan_out, const int64_t *edge_index, const int64_t *edge_type,
                           int64_t num_node, int64_t num_edge, int64_t num_relation, void *workspace,
                           size_t workspace_bytes, void *stream);

/*
 * Edge masks - the incremental plan for the training loop's per-step edge removal (remove_easy_edges,
 * /root/reference/ultra/base_nbfnet.py:54-77; the reference indexes the edge tensors and then re-sorts the result inside
 * every rspmm call, /root/reference/ultra/rspmm/rspmm.py:174-178).  keep_mask: device bytes, one per edge of the BASE
 * graph in the caller's edge order (a torch.bool tensor), non-zero = keep.  The plan is re-cut for that sub-graph entirely
 * on the device - removing edges keeps every order sorted, so: scan, compact, gather the new segment pointers, work-item
 * cutter; no sort - and WITHOUT a host synchronisation (CUDA-graph capturable): the host keeps upper bounds for grid and
 * scratch sizing, the kernels read the exact counts from device memory.  Forward / backward afterwards compute exactly
 * what a plan built from the filtered edge list would (same cutter on the same segment pointers: bit-identical).
 * Edge weights and weight_grad keep the base graph's length and order (entries of removed edges are ignored / zero).
 * ultra_rspmm_scratch_bytes grows while a mask is applied (split-row partials are bounded, not counted).
 * keep_mask == NULL restores the base graph.
 */
int ultra_rspmm_plan_apply_mask(ultra_rspmm_plan *plan, const unsigned char *keep_mask, void *stream);

void ultra_rspmm_plan_destroy(ultra_rspmm_plan *plan);

/*
 * Per-plan SM reservation: size this plan's persistent grids for `sm_limit` SMs (0 = all) so that a concurrent NCCL
 * collective (the overlapped all-gather of node states, ultra_b200/shard.py) finds free SMs.  Host-side field of the
 * plan; takes effect for launches issued afterwards.  (Round 1 had a process-global tunable here; a plan is the unit
 * a sharded runner owns, so two runners / streams can no longer disturb each other.)
 */
int ultra_rspmm_plan_set_sm_limit(ultra_rspmm_plan *plan, int sm_limit);

/* 1 if the caller's edge order was already sorted by the reference key (rspmm.py:15-17), else 0. */
int ultra_rspmm_plan_input_sorted(const ultra_rspmm_plan *plan);

/* Plan facts: what=0 num_node, 1 num_edge, 2 num_relation, 3 fwd work items, 4 fwd partial slots,
 * 5 input_grad partial slots, 6 relation_grad partial slots, 7 SM count used for grid sizing,
 * 8 / 9 input_grad / relation_grad work items, 10 + 3 * order + {0, 1, 2}: fix entries, long-segment fix entries, segments
 * of that order. */
int64_t ultra_rspmm_plan_info(const ultra_rspmm_plan *plan, int what);

/* Bytes of device scratch (split-row partials) forward/backward need for feature width `dim`. */
size_t ultra_rspmm_scratch_bytes(const ultra_rspmm_plan *plan, int64_t dim, int dtype);

/*
 * Bind per-call edge weights: permutes `edge_weight` (device, caller edge order, length num_edge, dtype)
 * into the plan's three edge orders inside the plan workspace.  Must precede forward/backward whenever the
 * weight VALUES changed.  edge_weight == NULL selects the all-ones specialisation (ULTRA's only case,
 * models.py:68,144,225), which is bit-identical because 1*x == x.
 */
int ultra_rspmm_set_edge_weight(ultra_rspmm_plan *plan, const void *edge_weight, int dtype, void *stream);

/*
 * Forward: replaces rspmm_{add,min,max}_{mul,add}_forward_cuda (rspmm.cu:216-267, kernel :21-83).
 *   relation (num_relation, dim), input (num_node, dim), output (num_node, dim): device, row-major, dtype
 *   scratch : device buffer of >= ultra_rspmm_scratch_bytes(plan, dim, dtype) (may be NULL if that is 0)
 * Every element of `output` is written (empty rows get the Nary identity 0 / +MAX / -MAX, operator.cuh:53,66,79).
 */
int ultra_rspmm_forward(const ultra_rspmm_plan *plan, int sum_op, int mul_op, int dtype, const void *relation,
                        const void *input, void *output, int64_t dim, void *scratch, size_t scratch_bytes,
                        void *stream);

/*
 * Backward: replaces rspmm_{...}_backward_cuda (rspmm.cu:269-351, kernels :85-214) WITHOUT atomics:
 *   input_grad    = gather-reduce over the by-col CSR, relation_grad = gather-reduce over the by-relation CSR,
 *   weight_grad   = per-edge dot product, written in the CALLER's edge order; pass NULL to skip it (the
 *                   reference leaves zeros unless edge_weight.requires_grad, rspmm.cu:295,312-331).
 * Every element of relation_grad / input_grad (/ weight_grad) is written; no pre-zeroing needed.
 */
int ultra_rspmm_backward(const ultra_rspmm_plan *plan, int sum_op, int mul_op, int dtype, const void *relation,
                         const void *input, const void *output, const void *output_grad, void *weight_grad,
                         void *relation_grad, void *input_grad, int64_t dim, void *scratch,
                         size_t scratch_bytes, void *stream);

/*
 * PNA aggregation in ONE pass: the reference issues four rspmm calls over the same edges
 * (/root/reference/ultra/layers.py:211-215: sum, sum over (relation**2, input**2), max, min); this entry point
 * gathers every row slice once and produces all four.  output4 is (4, num_node, dim): [0] sum, [1] sum of squares,
 * [2] max, [3] min, each exactly what the corresponding ultra_rspmm_forward call would write.  scratch must hold
 * 4 * ultra_rspmm_scratch_bytes(plan, dim, dtype).  Forward only: gradients compose from the four backward calls.
 */
int ultra_rspmm_forward_pna(const ultra_rspmm_plan *plan, int mul_op, int dtype, const void *relation,
                            const void *input, void *output4, int64_t dim, void *scratch, size_t scratch_bytes,
                            void *stream);

/*
 * Head-major variants: the NBFNet layer keeps its state as (num_heads, rows, head_dim) = (B, N, d); the reference
 * transposes it into (N, B*d) copies before every rspmm call and back afterwards
 * (/root/reference/ultra/layers.py:190-192,230).  These entry points read and write the (B, rows, d) layout
 * directly - relation (B, R, d), input/output/output_grad/input_grad (B, N, d), relation_grad (B, R, d) - and can
 * fold the layer's boundary term (the self-loop message, layers.py:199-207) into the store:
 *     output = Nary(rspmm(...), boundary)    i.e. + for sum, max / min for max / min;   boundary may be NULL.
 * relation_broadcast != 0: `relation` is ONE (R, d) block shared by every head - the relation-graph stack expands a
 * single nn.Embedding over the batch (/root/reference/ultra/layers.py:76); the kernels read it with head stride 0 instead
 * of a materialised (B, R, d) copy.  relation_grad is still written per head, (B, R, d): the caller sums it over B.
 * Requirements: head_dim * sizeof(dtype) is a multiple of 128 bytes, pointers 16-byte aligned.  For min/max the
 * `output` handed to backward_heads is the (boundary-folded) forward output.  weight_grad is not offered here.
 */
int ultra_rspmm_forward_heads(const ultra_rspmm_plan *plan, int sum_op, int mul_op, int dtype, const void *relation,
                              const void *input, const void *boundary, void *output, int64_t num_heads,
                              int64_t head_dim, int relation_broadcast, void *scratch, size_t scratch_bytes,
                              void *stream);
int ultra_rspmm_backward_heads(const ultra_rspmm_plan *plan, int sum_op, int mul_op, int dtype, const void *relation,
                               const void *input, const void *output, const void *output_grad, void *relation_grad,
                               void *input_grad, int64_t num_heads, int64_t head_dim, int relation_broadcast,
                               void *scratch, size_t scratch_bytes, void *stream);

/*
 * End-to-end forward from HOST buffers (pinned for full speed): relation/input/output live in host memory;
 * the call pipelines column slabs  H2D copy -> kernel -> D2H copy  on three internal streams forked from and
 * joined back to `stream`.  `staging` is a caller-provided device buffer; ultra_rspmm_host_staging_bytes()
 * gives the size for a given slab width.  Mirrors what the reference does when handed CPU tensors
 * (dispatch on input.device.type, rspmm.py:19-22) -- except the arithmetic still runs on the GPU.
 */
size_t ultra_rspmm_host_staging_bytes(const ultra_rspmm_plan *plan, int64_t dim, int64_t slab_dim, int dtype);
int ultra_rspmm_forward_host(const ultra_rspmm_plan *plan, int sum_op, int mul_op, int dtype,
                             const void *h_relation, const void *h_input, void *h_output, int64_t dim,
                             int64_t slab_dim, void *staging, size_t staging_bytes, void *stream);

/*
 * Fused NBFNet layer epilogue, fp32:
 *     out[m,:] = (shortcut ? x[m,:] : 0) + act( LN( weight . [x[m,:] ; u[m,:]] + bias ) )
 * replaces GeneralizedRelationalConv.update (/root/reference/ultra/layers.py:233-240: cat, Linear(128->64), LayerNorm,
 * ReLU) and the residual add of the Bellman-Ford loop (/root/reference/ultra/models.py:158-160).
 *   x, u, out : device (num_rows, 64) row-major, num_rows = heads * nodes (the (B, N, 64) layer state as is)
 *   weight_t  : device (128, 64) = nn.Linear.weight transposed;  bias / ln_gamma / ln_beta : device (64) or NULL
 * Only in_dim == 128, out_dim == 64 is implemented (sum / mean / max aggregation at ULTRA's dim 64); anything else
 * returns an error and the caller keeps its own epilogue.  One implementation: pipelined tcgen05 kernel with a BF16x3
 * operand split (fp32-level accuracy: 3.2e-6 max abs error against fp64 on outputs of magnitude 7).
 * sm_limit: persistent CTAs to launch at most (0 = one per SM), see ultra_rspmm_plan_set_sm_limit.
 */
int ultra_nbfnet_update_f32(const float *x, const float *u, const float *weight_t, const float *bias,
                            const float *ln_gamma, const float *ln_beta, float *out, int64_t num_rows, int64_t in_dim,
                            int64_t out_dim, float ln_eps, int layer_norm, int relu, int shortcut, int sm_limit,
                            void *stream);

/*
 * Training-mode epilogue (second half).  Forward: the same tcgen05 kernel as
 * ultra_nbfnet_update_f32 (LayerNorm always on), which additionally keeps what the backward needs:
 *   z_out (num_rows, 64)  the Linear output W.[x;u] + bias BEFORE LayerNorm;  stats_out (num_rows, 2) = (mean, rstd) per row.
 * Backward of  y = relu(LayerNorm(z) * gamma + beta)  in ONE pass (read out_grad and z, write z_grad):
 *   z_grad (num_rows, 64);  partials (ultra_nbfnet_update_bwd_grid(num_rows), 3, 64): per-CTA partial sums of
 *   [0] d gamma, [1] d beta, [2] d bias (= column sums of z_grad) - the caller sums over the first axis (fixed order:
 *   deterministic).  The three GEMM-shaped gradients that remain - d[x;u] = z_grad . W, dW = z_grad^T . [x;u] - are plain
 *   library GEMMs on the caller's side (ultra_b200.fused), as is the residual's pass-through of out_grad.
 */
int ultra_nbfnet_update_train_f32(const float *x, const float *u, const float *weight_t, const float *bias,
                                  const float *ln_gamma, const float *ln_beta, float *out, float *z_out,
                                  float *stats_out, int64_t num_rows, int64_t in_dim, int64_t out_dim, float ln_eps,
                                  int relu, int shortcut, int sm_limit, void *stream);
int ultra_nbfnet_update_bwd_grid(int64_t num_rows);
int ultra_nbfnet_update_bwd_f32(const float *out_grad, const float *z, const float *stats, const float *ln_gamma,
                                const float *ln_beta, float *z_grad, float *partials, int64_t num_rows, int64_t dim,
                                int relu, void *stream);

/*
 * All per-layer relation projections of the entity stack in ONE launch, fp32:
 *     out[l, m, :] = W2[l] . relu( W1[l] . x[m, :] + b1[l] ) + b2[l]          l < num_layers, m < num_rows
 * replaces the nn.Sequential(Linear(64,64), ReLU, Linear(64,64)) every entity layer applies to the SAME (B, R, 64)
 * relation representations (/root/reference/ultra/layers.py:60-64,80): 3 launches x 6 layers in the reference.
 *   x (num_rows, 64);  w1_t, w2_t (num_layers, 64, 64) = the nn.Linear weights TRANSPOSED ([k][n]);  b1, b2 (num_layers, 64)
 *   out (num_layers, num_rows, 64).  Only dim == 64; other sizes return an error and the caller keeps its own MLP.
 */
int ultra_relation_project_f32(const float *x, const float *w1_t, const float *b1, const float *w2_t, const float *b2,
                               float *out, int64_t num_rows, int64_t dim, int64_t num_layers, void *stream);

/*
 * The entity model's scoring MLP over every (query, entity) pair in ONE pass, fp32:
 *     out[m] = w2 . relu( W1h . h[m, :] + qb[m / rows_per_head, :] ) + b2                    m < num_rows
 * replaces  mlp(cat[h, query])  of /root/reference/ultra/models.py:166-170,202-208 (Linear(128,128) -> ReLU -> Linear(128,1))
 * for the all-tails evaluation protocol (tasks.py:79-91): the first Linear is split into its h half W1h (128, 64) and its
 * query half, whose product with the query (+ bias) the caller passes as qb (num_heads, 128).
 *   h (num_rows, 64) = the (B, N, 64) final states, rows_per_head = N;  w1h_t (64, 128) = W1h TRANSPOSED ([k][n]);
 *   w2 (128);  b2 (1, on the device: no host read, CUDA-graph capturable);  out (num_rows).  Only dim == 64, hidden == 128; other sizes return an error and the caller keeps its MLP.
 */
int ultra_score_head_f32(const float *h, const float *w1h_t, const float *qb, const float *w2, const float *b2,
                         float *out, int64_t num_rows, int64_t rows_per_head, int64_t dim, int64_t hidden, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ULTRA_RSPMM_H_ */
