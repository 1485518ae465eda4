// rspmm_kernels.cuh -- sm_100a device code for the generalized relational SpMM.
//
// One templated "segment gather-reduce" kernel serves the three passes of the hot path:
//   FWD     : segment = aggregating node (row)  -> output          replaces rspmm.cu:21-83   (reference)
//   BWD_IN  : segment = gathered node  (col)    -> input_grad      replaces rspmm.cu:153-214 (atomics)
//   BWD_REL : segment = relation type           -> relation_grad   replaces rspmm.cu:153-214 (atomics)
// Functor semantics follow /root/reference/ultra/rspmm/source/operator.cuh:13-80 exactly (see ops below).
//
// Data layout (all built once per graph by rspmm_plan.cu):
//   meta[e]   int2  (a, b | last_of_segment<<31)  edges sorted by (segment, secondary key)
//   items[k]  int4  {e_begin, e_end, j0, slot}    <= kItemEdges edges; whole segments (slot=-1), a single chunk of a
//                                                 long segment (slot>=0 -> partial[slot]), or a member of a merged
//                                                 group of 4 full chunks (j0 |= kMergedBit; leader slot>=0, followers -2)
//   nonempty[j]     ids of non-empty segments, in order (j0 indexes it)
//   fix[f]    int4  {segment, slot_begin, slot_count, 0}  empty segments (count 0) and long segments
//
// Work decomposition: a column TILE is 128 bytes of a feature row (32 fp32 / 16 fp64).  8 lanes own one edge's
// tile slice (one 16-byte vector load per lane, a full 128-byte L2 line per edge), so a warp walks 4 items
// at once with no cross-lane traffic.  CTAs are persistent and walk (tile, item-chunk) units TILE-MAJOR so
// that the concurrently active tiles of `input` (N x 128 B each) stay L2-resident: HBM sees ~compulsory
// traffic and the gathers are served by L2/L1.  Shared memory holds, per CTA: the relation tile (R x 128 B, staged
// once per (CTA, tile) visit by TMA bulk copies - cp.async.bulk, one 128-byte row slice each, mbarrier completion),
// for small graphs also the gathered matrix' tile (N x 128 B), and each warp's item metadata (coalesced loads at item
// start, broadcast LDS in the loop).  The per-edge gathers themselves stay LDG.128 -> registers:
// tools/gather_microbench.cu measures 128-byte TMA bulk gathers at <= the LDG rate (both ~7 TB/s, the L2 limit) while
// costing an smem round trip and the L1 hits.
#pragma once
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>
#include <type_traits>

namespace ultra {

#ifndef ULTRA_FWD_UNROLL
#define ULTRA_FWD_UNROLL 4  // edges per lane in flight in the single-gather passes (register budget: 64/thread)
#endif
#ifndef ULTRA_FWD_UNROLL_W
#define ULTRA_FWD_UNROLL_W ULTRA_FWD_UNROLL  // same, for the (rare) non-unit-weight path, which also holds U weights per lane
#endif
#ifndef ULTRA_FWD_PIPE
#define ULTRA_FWD_PIPE 0  // experiment: software-pipelined gathers in the forward pass (see walk_items_impl)
#endif
#ifndef ULTRA_META_PAIR
#define ULTRA_META_PAIR 1  // item metadata read two edges at a time (explicit LDS.128); 0 = one LDS.64 per edge
#endif
#ifndef ULTRA_BOTH_UNROLL
#define ULTRA_BOTH_UNROLL 4  // same for the fused backward pass (BWD_BOTH)
#endif
constexpr int kLanesPerEdge = 8;
constexpr int kSubPerWarp = 32 / kLanesPerEdge;  // 4 items walked concurrently by one warp
constexpr int kTileBytes = 128;
constexpr int kItemEdges = 64;   // max edges of a whole-segment item and chunk length of a long segment
#ifndef ULTRA_META_PAD
#define ULTRA_META_PAD 16  // bytes of padding between the 4 sub-groups' metadata regions (16: bank-conflict-free; 0: none)
#endif
// per warp: the (a, b) metadata of its 4 items, staged in shared memory.  The 4 sub-groups read their entry k in the same
// instruction; with 512-byte regions those 4 addresses fall on the same banks (ncu: 52 M conflict wavefronts per forward
// call, 12 % of the data-pipe traffic), with 528-byte regions they are 16 bytes apart.
constexpr int kMetaSubBytes = kItemEdges * 8 + ULTRA_META_PAD;
constexpr int kMetaStageBytes = kSubPerWarp * kMetaSubBytes;
constexpr int kMergedBit = 1 << 30;  // items[k].z flag: member of a merged group of 4 full chunks
#ifndef ULTRA_CTA_THREADS
#define ULTRA_CTA_THREADS 512  // threads per persistent CTA ...
#endif
#ifndef ULTRA_CTAS_PER_SM
#define ULTRA_CTAS_PER_SM 2    // ... and resident CTAs per SM (512 x 2 and 1024 x 1 both give 32 warps/SM at 64 registers; the
#endif                         // latter keeps ONE relation tile per SM in shared memory and so leaves L1 ~3x the capacity)
constexpr int kCtaThreads = ULTRA_CTA_THREADS;
constexpr int kCtasPerSm = ULTRA_CTAS_PER_SM;
constexpr int kCtaWarps = kCtaThreads / 32;

enum { FWD = 0, BWD_IN = 1, BWD_REL = 2, BWD_BOTH = 3 };
// BWD_BOTH (opt-in, tunable "bwd_fused_relgrad"): the BWD_IN pass that ALSO scatters each edge's relation_grad
// contribution with one 16-byte vector reduction per lane (red.global.add.v4.f32 -> REDG.E.ADD.F32x4) into a pre-zeroed
// relation_grad, so the two-gather BWD_REL pass disappears.  input_grad stays a deterministic gather-reduce;
// relation_grad becomes order-dependent in the last bits, like the reference's atomicAdd backward (rspmm.cu:198-209).
__host__ __device__ constexpr bool by_col(int mode) { return mode == BWD_IN || mode == BWD_BOTH; }
__host__ __device__ constexpr int order_of(int mode) { return mode == BWD_BOTH ? BWD_IN : mode; }
enum { S_ADD = 0, S_MIN = 1, S_MAX = 2, S_PNA = 3 };  // S_PNA: forward only, 4 aggregates in one pass (sum, sum of
                                                       // squares, max, min - the four rspmm calls of layers.py:211-215)
enum { M_MUL = 0, M_ADD = 1 };

// ---------------------------------------------------------------------------------------------------------
// 16-byte vectors
template <typename T> struct alignas(16) Vec { T v[16 / sizeof(T)]; };
template <typename T> struct VecN { static constexpr int value = 16 / sizeof(T); };

template <typename T> __device__ __forceinline__ Vec<T> ld_gather(const T *p) {  // read-only path, L1-allocating
    Vec<T> r;
    if constexpr (sizeof(T) == 4) {
        float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
        double2 t = __ldg(reinterpret_cast<const double2 *>(p));
        r.v[0] = t.x; r.v[1] = t.y;
    }
    return r;
}
template <typename T> __device__ __forceinline__ void st_stream(T *p, const Vec<T> &x) {  // evict-first store
    if constexpr (sizeof(T) == 4) {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(x.v[0], x.v[1], x.v[2], x.v[3]));
    } else {
        __stcs(reinterpret_cast<double2 *>(p), make_double2(x.v[0], x.v[1]));
    }
}
template <typename T> __device__ __forceinline__ void red_add_vec(T *p, const Vec<T> &x) {  // fire-and-forget += at L2
    if constexpr (sizeof(T) == 4) {
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(x.v[0]), "f"(x.v[1]), "f"(x.v[2]), "f"(x.v[3])
                     : "memory");
    } else {
        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(x.v[0]) : "memory");
        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p + 1), "d"(x.v[1]) : "memory");
    }
}
template <typename T> __device__ __forceinline__ void st_plain(T *p, const Vec<T> &x) {
    *reinterpret_cast<Vec<T> *>(p) = x;
}

// ---------------------------------------------------------------------------------------------------------
// functors (operator.cuh:13-80)
template <typename T> struct Lim;
template <> struct Lim<float> { static __device__ __forceinline__ float max() { return FLT_MAX; } };
template <> struct Lim<double> { static __device__ __forceinline__ double max() { return DBL_MAX; } };

template <int MUL, typename T> __device__ __forceinline__ T bin_fwd(T rel, T in) {  // operator.cuh:15,30
    return MUL == M_MUL ? rel * in : rel + in;
}
template <int SUM, typename T> __device__ __forceinline__ T nary_zero() {  // operator.cuh:53,66,79
    return SUM == S_ADD ? T(0) : (SUM == S_MIN ? Lim<T>::max() : -Lim<T>::max());
}
template <int SUM, typename T> __device__ __forceinline__ T nary_fwd(T r, T x) {  // operator.cuh:45,58,71
    if (SUM == S_ADD) return r + x;
    if (SUM == S_MIN) return r < x ? r : x;
    return r > x ? r : x;
}
// accumulator a of a (possibly multi-aggregate) reduction: which Nary op it folds with
__host__ __device__ constexpr int acc_count(int sum) { return sum == S_PNA ? 4 : 1; }
__host__ __device__ constexpr int acc_op(int sum, int a) {
    return sum != S_PNA ? sum : (a < 2 ? S_ADD : (a == 2 ? S_MAX : S_MIN));
}
template <typename T> __device__ __forceinline__ T nary_zero_dyn(int op) {
    return op == S_ADD ? T(0) : (op == S_MIN ? Lim<T>::max() : -Lim<T>::max());
}
template <typename T> __device__ __forceinline__ T nary_fwd_dyn(int op, T r, T x) {
    if (op == S_ADD) return r + x;
    if (op == S_MIN) return r < x ? r : x;
    return r > x ? r : x;
}
template <int SUM, typename T> __device__ __forceinline__ T nary_bwd(T r, T x) {  // operator.cuh:49,62,75
    if (SUM == S_ADD) return T(1);
    return r == x ? T(1) : T(0);
}

// ---------------------------------------------------------------------------------------------------------
// Feature-matrix layout.  A 128-byte column tile t of row r lives at
//     base + (t / tpg) * grp_bytes(rows) + r * row_bytes + (t % tpg) * 128
//   flat  (rows, D)      : tpg = #tiles, row_bytes = D * sizeof(T)                (the reference's rspmm layout)
//   heads (B, rows, d)   : tpg = d * sizeof(T) / 128, row_bytes = d * sizeof(T), grp_bytes = rows * row_bytes
//                          (the NBFNet layer's native layout: no transpose copies, ultra/layers.py:190-192)
struct Layout {
    int tpg;                     // tiles per group (head)
    int group_cols;              // columns per group (D flat, d heads) - ragged last tile check
    unsigned row_bytes;
    long long grp_node_bytes;    // group stride of the N-row matrices (input, output, outgrad, boundary, input_grad)
    long long grp_rel_bytes;     // group stride of the relation matrix; 0 = ONE (R, d) block broadcast over the heads (the
                                 // relation-graph stack expands a single embedding, ultra/layers.py:76 - never materialised)
    long long grp_rgrad_bytes;   // group stride of relation_grad (always per head; the caller sums a broadcast one)
};

template <typename T> struct SegParams {
    const int2 *meta;
    const T *w;  // per-edge weight in this order, or nullptr = all ones
    const int *w_nonunit;  // device flag written by set_edge_weight: 0 -> every weight == 1 (skip the loads)
    int *sched;            // {next block ticket, CTAs finished}: dynamic tile-major work queue, reset by the last CTA
    const int *nonempty;
    const int4 *items;
    const int *counts;     // device-side result block of the plan's cutter: [0] num_items, [1] num_fix, [2] num_fix_long,
                           // [3] num_slots, [7] num_edge.  The launch geometry is derived from it IN the kernel, so a plan
                           // whose edge set was re-cut on the device (ultra_rspmm_plan_apply_mask) needs no host round trip
    const T *relation;  // (R, D)
    const T *input;     // (N, D)
    const T *output;    // (N, D)   backward only
    const T *outgrad;   // (N, D)   backward only
    T *dst;             // output / input_grad / relation_grad
    T *rgrad;           // BWD_BOTH only: relation_grad (pre-zeroed), receives vector reductions
    T *partial;         // (num_slots, D), always flat
    long long dst_plane;      // S_PNA: element stride between the 4 output matrices (sum, sq_sum, max, min)
    long long partial_plane;  // S_PNA: element stride between the 4 partial buffers
    const T *boundary;  // forward only, optional: out = Nary(agg, boundary) (the NBFNet self-loop message); dst layout
    Layout lay;         // where tile t / row r of the feature matrices live (flat (rows, D) or head-major (B, rows, d))
    int D;
    int R;
    int N;
    int num_tiles;
    int items_per_warp;
    int run;
};

// which tensors a (MODE, SUM, MUL) combination touches per edge
template <int MODE, int SUM, int MUL> struct Needs {
    static constexpr bool out = (MODE != FWD) && (SUM != S_ADD);                       // output[a] for dout/dy
    static constexpr bool inb = (MODE == BWD_REL) && (MUL == M_MUL || SUM != S_ADD);   // input[b]
    static constexpr bool rel = (MODE == FWD) || (by_col(MODE) && (MUL == M_MUL || SUM != S_ADD));
    // the segment's own input / relation row (BWD_BOTH x mul: d x / d rel = input[col], the segment's row)
    static constexpr bool self = ((MODE != FWD) && (SUM != S_ADD)) || (MODE == BWD_BOTH && MUL == M_MUL);
    static constexpr int red = (MODE == FWD) ? SUM : S_ADD;   // S_PNA: see acc_op()
    static constexpr int nacc = acc_count(red);
    static constexpr int gathers = 1 + (out ? 1 : 0) + (inb ? 1 : 0);  // row slices fetched per edge
    static constexpr int unroll = (MODE == BWD_BOTH) ? (gathers == 1 ? ULTRA_BOTH_UNROLL : 2) : (SUM != S_PNA && gathers == 1) ? ULTRA_FWD_UNROLL : ((SUM != S_PNA && gathers == 2 && SUM == S_ADD) ? 4 : 2);  // 64 regs
};

// shared-memory vector load through a 32-bit shared-window address (address math stays in 32 bits)
template <typename T> __device__ __forceinline__ Vec<T> lds_vec(uint32_t addr) {
    Vec<T> r;
    if constexpr (sizeof(T) == 4) {
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "r"(addr));
    } else {
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "r"(addr));
    }
    return r;
}
// row gather: base + row * row_bytes is ONE IMAD.WIDE.U32 (32x32+64) instead of a 64x64 multiply chain
template <typename T> __device__ __forceinline__ Vec<T> ld_row(const char *base, unsigned row, unsigned row_bytes) {
    return ld_gather(reinterpret_cast<const T *>(base + (unsigned long long)row * row_bytes));
}

// One warp walks 4 items (one per 8-lane sub-group) of one column tile.
//   UNITW : every edge weight is 1 (device flag) -> no weight loads, y = Bin(rel, in) (1 * x == x exactly)
// The loop has a predicate-free main part (all 4 sub-groups still have edges) and a predicated tail.
template <typename T, int MODE, int SUM, int MUL, int SM, bool UNITW>
__device__ __forceinline__ void walk_items_impl(const SegParams<T> &p, int wi, int tile, uint32_t rel_s_addr,
                                                uint32_t g1_s_addr, uint32_t meta_s_addr) {
    constexpr bool REL_SMEM = SM >= 1;
    constexpr bool G1_SMEM = SM == 2;
    using N = Needs<MODE, SUM, MUL>;
    constexpr int V = VecN<T>::value;
    constexpr int TILE = kLanesPerEdge * V;
    constexpr bool kSingleGather = MODE != BWD_BOTH && SUM != S_PNA && N::gathers == 1;
    constexpr int U = (kSingleGather && !UNITW) ? ULTRA_FWD_UNROLL_W : N::unroll;
    const int lane = threadIdx.x & 31;
    const int q = lane & (kLanesPerEdge - 1);
    const int sg = lane >> 3;
    const int grp = tile / p.lay.tpg, tin = tile - grp * p.lay.tpg;
    int cin = tin * TILE + q * V;  // column inside the group
    const bool st_ok = cin < p.lay.group_cols;  // lanes past a ragged last tile compute on column 0, never store
    if (!st_ok) cin = 0;
    const int c0 = grp * p.lay.group_cols + cin;  // global column (flat partial buffer)
    const unsigned row_bytes = p.lay.row_bytes;
    const size_t off_n = (size_t)grp * p.lay.grp_node_bytes + (size_t)cin * sizeof(T);
    const size_t off_r = (size_t)grp * p.lay.grp_rel_bytes + (size_t)cin * sizeof(T);
    const size_t off_rg = (size_t)grp * p.lay.grp_rgrad_bytes + (size_t)cin * sizeof(T);

    const int4 it = __ldg(p.items + (size_t)wi * kSubPerWarp + sg);
    const int e0 = it.x, slot = it.w;
    const bool merged = (it.z & kMergedBit) != 0;  // warp-uniform: 4 full chunks of ONE long segment (cut_items)
    int j = it.z & (kMergedBit - 1);
    const int len = it.y - it.x;
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    const int minlen = __reduce_min_sync(0xffffffffu, len);
    const int len_m1 = len > 0 ? len - 1 : 0;
    unsigned seg = (unsigned)__ldg(p.nonempty + j);

    const int2 *mp = p.meta + e0;
    const T *wp = UNITW ? nullptr : p.w + e0;
    const char *g1_c = reinterpret_cast<const char *>(MODE == FWD ? p.input : p.outgrad) + off_n;
    const char *out_c = reinterpret_cast<const char *>(p.output) + off_n;
    const char *in_c = reinterpret_cast<const char *>(p.input) + off_n;
    const char *rel_c = reinterpret_cast<const char *>(p.relation) + off_r;
    const char *self_c = by_col(MODE) ? in_c : rel_c;
    char *dst_c = reinterpret_cast<char *>(p.dst) + (MODE == BWD_REL ? off_rg : off_n);
    char *rg_c = reinterpret_cast<char *>(p.rgrad) + off_rg;
    const char *bnd_c = reinterpret_cast<const char *>(p.boundary) + off_n;
    const bool has_bnd = (MODE == FWD) && p.boundary != nullptr;
    uint32_t rel_q = rel_s_addr + (uint32_t)q * 16u;
    uint32_t g1_q = g1_s_addr + (uint32_t)q * 16u;
    // Pin the per-lane bases in registers: otherwise ptxas keeps "uniform base + lane offset" apart and re-adds the
    // uniform half for every gather (IMAD.WIDE + IADD3 + IADD3.X instead of one IMAD.WIDE), and rematerialises the
    // shared-window base (S2UR/ULEA) every iteration.
    asm volatile("" : "+l"(g1_c), "+l"(out_c), "+l"(in_c), "+l"(rel_c), "+r"(rel_q), "+r"(g1_q));

    constexpr int NA = N::nacc;
    Vec<T> acc[NA], self;
#pragma unroll
    for (int a = 0; a < NA; a++)
#pragma unroll
        for (int k = 0; k < V; k++) acc[a].v[k] = nary_zero_dyn<T>(acc_op(N::red, a));
#pragma unroll
    for (int k = 0; k < V; k++) self.v[k] = T(0);
    if (N::self) self = ld_row<T>(self_c, seg, row_bytes);

    // Relation-run reuse, small graphs only (SM == 2: gathered tile AND relation tile in shared memory, so the pass is
    // bound by shared-memory wavefronts, two LDS.128 per edge-tile).  plan_build sorts such a graph's segments by relation
    // type first (rspmm_plan.cu), the lane keeps the current relation slice in registers and re-reads it only when the type
    // changes: ULTRA's relation graph has 4 edge types, so nearly every relation LDS disappears (measured on the
    // relation graph of the FB15k-237-scale KG, 64 heads: forward 1.20 -> 0.91 ms).  Same values, same order of operations.
    // NOT applied to the large-graph kernels: there the same trick (and its twin, reusing the gathered row of parallel
    // edges, 35 % of the Zipf benchmark graph's edges) removed up to a third of the L1 wavefronts and bought nothing
    // (profiles/r02_run_reuse_ab.txt) - that pass is bound by L2 latency under load, not by the L1 data pipe.
    constexpr bool kReuseRel = N::rel && SM == 2;
    unsigned prev_rid = 0xffffffffu;
    Vec<T> rcur;
#pragma unroll
    for (int k = 0; k < V; k++) rcur.v[k] = T(0);

    // Stage this item's metadata (<= 64 x 8 bytes) in shared memory with coalesced loads: the hot loop then reads it
    // with a ~30-cycle broadcast LDS instead of a dependent global load, so each group of U edges exposes ONE long
    // latency (the row gather) instead of two (metadata -> gather).  The stall profile was 72 % long_scoreboard.
    // Entries past the item's end replicate its last edge, so the predicated tail needs no index clamping.
    const uint32_t my_meta = meta_s_addr + (uint32_t)sg * kMetaSubBytes;
    __syncwarp();  // the previous item's metadata is no longer being read by the other lanes
#pragma unroll
    for (int t = 0; t < kItemEdges / kLanesPerEdge; t++) {
        const int k = t * kLanesPerEdge + q;
        const int2 mm = __ldg(mp + (k < len ? k : len_m1));
        asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(my_meta + (uint32_t)k * 8u), "r"(mm.x), "r"(mm.y) : "memory");
    }
    __syncwarp();

    // a segment ended inside a whole-segment item: emit the finished row, start the next one
    auto end_segment = [&]() {
        if (st_ok) {
            if (has_bnd && NA == 1) {
                const Vec<T> bv = ld_row<T>(bnd_c, seg, row_bytes);
#pragma unroll
                for (int k = 0; k < V; k++) acc[0].v[k] = nary_fwd_dyn<T>(N::red, acc[0].v[k], bv.v[k]);
            }
#pragma unroll
            for (int a = 0; a < NA; a++)
                st_stream(reinterpret_cast<T *>(dst_c + (unsigned long long)seg * row_bytes) + (size_t)a * p.dst_plane,
                          acc[a]);
        }
#pragma unroll
        for (int a = 0; a < NA; a++)
#pragma unroll
            for (int k = 0; k < V; k++) acc[a].v[k] = nary_zero_dyn<T>(acc_op(N::red, a));
        j++;
        seg = (unsigned)__ldg(p.nonempty + j);  // padded by one entry (0): always a valid row
        if (N::self) self = ld_row<T>(self_c, seg, row_bytes);
    };

    auto body = [&](auto pred_tag, int i) {
        constexpr bool PRED = decltype(pred_tag)::value;
        int2 m[U];
        T wv[U];
        bool ok[U];
#if ULTRA_META_PAIR
        // two edges' metadata per LDS.128: i and U are even (16-byte aligned), and entries past the item's end replicate
        // its last edge, so no clamping is needed.  Each LDS is one wavefront on the L1/shared data pipe, the pipe this
        // kernel saturates (gather + relation-row LDS + metadata = ~2.25 wavefronts per edge-tile + L1 fills)
        static_assert(U % 2 == 0, "pair loads need an even unroll");
        static_assert(kItemEdges % U == 0, "pair loads are not clamped: the last group must end inside the 64 staged entries");
#pragma unroll
        for (int u = 0; u < U; u += 2)
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(m[u].x), "=r"(m[u].y), "=r"(m[u + 1].x), "=r"(m[u + 1].y)
                         : "r"(my_meta + (uint32_t)(i + u) * 8u));
#endif
#pragma unroll
        for (int u = 0; u < U; u++) {
            int k = i + u;
            ok[u] = true;
            if (PRED) { ok[u] = k < len; k = k < len ? k : len_m1; }
#if !ULTRA_META_PAIR
            asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(m[u].x), "=r"(m[u].y) : "r"(my_meta + (uint32_t)k * 8u));
#endif
            wv[u] = T(1);
            if (!UNITW) wv[u] = __ldg(wp + k);
        }
        Vec<T> g1[U], g2[U], g3[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (G1_SMEM) g1[u] = lds_vec<T>(g1_q + ((uint32_t)m[u].x << 7));
            else g1[u] = ld_row<T>(g1_c, (unsigned)m[u].x, row_bytes);
            if (N::out) g2[u] = ld_row<T>(out_c, (unsigned)m[u].x, row_bytes);
            if (N::inb) g3[u] = ld_row<T>(in_c, (unsigned)m[u].y & 0x7fffffffu, row_bytes);
        }
        Vec<T> rr[U];
        if (N::rel && !REL_SMEM) {  // relation rows from L1/L2: issue all loads up front (long latency)
#pragma unroll
            for (int u = 0; u < U; u++) rr[u] = ld_row<T>(rel_c, (unsigned)m[u].y & 0x7fffffffu, row_bytes);
        }
        auto consume = [&](int u) {
            // relation tile in shared memory: fetch the vector just in time (~30 cycles, other warps cover it) instead
            // of holding U of them in registers - the registers are better spent on row gathers in flight.
            // (b << 7): 128 bytes per relation row of the smem tile; the shift drops the flag bit 31
            if (N::rel && REL_SMEM) {
                if (kReuseRel) {
                    const unsigned rid = (unsigned)m[u].y << 7;
                    if (rid != prev_rid) rcur = lds_vec<T>(rel_q + rid);
                    prev_rid = rid;
                    rr[u] = rcur;
                } else rr[u] = lds_vec<T>(rel_q + ((uint32_t)m[u].y << 7));
            }
            Vec<T> nxt[NA], rg;
#pragma unroll
            for (int k = 0; k < V; k++) {
                if (MODE == FWD) {
                    // rspmm.cu:68-70   x = Bin(rel, in); y = w * x; out = Nary(out, y)
                    T y = bin_fwd<MUL>(rr[u].v[k], g1[u].v[k]);
                    if (!UNITW) y = wv[u] * y;
                    if (SUM == S_PNA) {
                        // layers.py:211-215: sum, sum over (relation**2, input**2), max, min of the same messages
                        T y2 = bin_fwd<MUL>(rr[u].v[k] * rr[u].v[k], g1[u].v[k] * g1[u].v[k]);
                        if (!UNITW) y2 = wv[u] * y2;
                        nxt[0].v[k] = acc[0].v[k] + y;
                        nxt[NA > 1 ? 1 : 0].v[k] = acc[NA > 1 ? 1 : 0].v[k] + y2;
                        nxt[NA > 2 ? 2 : 0].v[k] = acc[NA > 2 ? 2 : 0].v[k] > y ? acc[NA > 2 ? 2 : 0].v[k] : y;
                        nxt[NA > 3 ? 3 : 0].v[k] = acc[NA > 3 ? 3 : 0].v[k] < y ? acc[NA > 3 ? 3 : 0].v[k] : y;
                    } else {
                        nxt[0].v[k] = nary_fwd_dyn<T>(N::red, acc[0].v[k], y);
                    }
                } else if (by_col(MODE)) {
                    // rspmm.cu:198-209  input_grad[col] += og * dout_dy * w * dx_din ; segment = col
                    T g = g1[u].v[k];
                    if (N::out) {
                        T y = bin_fwd<MUL>(rr[u].v[k], self.v[k]);
                        if (!UNITW) y = wv[u] * y;
                        g = g * nary_bwd<SUM>(g2[u].v[k], y);
                    }
                    if (!UNITW) g = g * wv[u];
                    if (MODE == BWD_BOTH) rg.v[k] = MUL == M_MUL ? g * self.v[k] : g;  // ... * dx_drel
                    if (MUL == M_MUL) g = g * rr[u].v[k];
                    nxt[0].v[k] = acc[0].v[k] + g;
                } else {
                    // rspmm.cu:198-208  relation_grad[type] += og * dout_dy * w * dx_drel ; segment = type
                    T g = g1[u].v[k];
                    if (N::out) {
                        T y = bin_fwd<MUL>(self.v[k], g3[u].v[k]);
                        if (!UNITW) y = wv[u] * y;
                        g = g * nary_bwd<SUM>(g2[u].v[k], y);
                    }
                    if (!UNITW) g = g * wv[u];
                    if (MUL == M_MUL) g = g * g3[u].v[k];
                    nxt[0].v[k] = acc[0].v[k] + g;
                }
            }
#pragma unroll
            for (int a = 0; a < NA; a++)
#pragma unroll
                for (int k = 0; k < V; k++) acc[a].v[k] = ok[u] ? nxt[a].v[k] : acc[a].v[k];
            if (MODE == BWD_BOTH) {  // relation_grad[type] += ... (rspmm.cu:207-208), 16 bytes per lane
                if (ok[u] && st_ok)
                    red_add_vec(reinterpret_cast<T *>(rg_c + (unsigned long long)((unsigned)m[u].y & 0x7fffffffu) * row_bytes), rg);
            }
        };
        // one row-end test for the whole group of U edges (row ends are rare: the common path is pure FMA)
        int any = 0;
#pragma unroll
        for (int u = 0; u < U; u++) any |= m[u].y;
        if (!(any < 0 && slot == -1)) {
#pragma unroll
            for (int u = 0; u < U; u++) consume(u);
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) {
                consume(u);
                if (m[u].y < 0 && ok[u]) end_segment();
            }
        }
    };

    int i = 0;
#if ULTRA_FWD_PIPE
    // Experiment (compile-time, off by default): two register stages in the predicate-free part of the forward pass - the
    // gathers of group g + 1 are issued BEFORE group g is consumed, so a lane keeps 2 * U row gathers in flight instead of
    // U with a full stop in between.  Needs ~16 more registers than the 64 the 512 x 2 geometry allows (it spills there);
    // meant for the 768 x 1 geometry (85 registers).  Same arithmetic in the same order: bit-identical results.
    if constexpr (MODE == FWD && SUM != S_PNA && UNITW && SM == 1 && ULTRA_META_PAIR) {
        int2 mA[U], mB[U];
        Vec<T> gA[U], gB[U];
        auto issue = [&](int2 (&m)[U], Vec<T> (&g)[U], int i0) {
#pragma unroll
            for (int u = 0; u < U; u += 2)
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(m[u].x), "=r"(m[u].y), "=r"(m[u + 1].x), "=r"(m[u + 1].y)
                             : "r"(my_meta + (uint32_t)(i0 + u) * 8u));
#pragma unroll
            for (int u = 0; u < U; u++) g[u] = ld_row<T>(g1_c, (unsigned)m[u].x, row_bytes);
        };
        auto retire = [&](int2 (&m)[U], Vec<T> (&g)[U]) {
            int any = 0;
#pragma unroll
            for (int u = 0; u < U; u++) any |= m[u].y;
            const bool ends = any < 0 && slot == -1;
#pragma unroll
            for (int u = 0; u < U; u++) {
                const Vec<T> rr = lds_vec<T>(rel_q + ((uint32_t)m[u].y << 7));
#pragma unroll
                for (int k = 0; k < V; k++)
                    acc[0].v[k] = nary_fwd_dyn<T>(N::red, acc[0].v[k], bin_fwd<MUL>(rr.v[k], g[u].v[k]));
                if (ends && m[u].y < 0) end_segment();
            }
        };
        const int n_full = minlen / U;
        int g = 0;
        if (n_full > 0) {
            issue(mA, gA, 0);
            while (g + 2 < n_full) {
                issue(mB, gB, (g + 1) * U);
                retire(mA, gA);
                issue(mA, gA, (g + 2) * U);
                retire(mB, gB);
                g += 2;
            }
            if (g + 1 < n_full) {
                issue(mB, gB, (g + 1) * U);
                retire(mA, gA);
                retire(mB, gB);
                g += 2;
            } else {
                retire(mA, gA);
                g += 1;
            }
        }
        i = g * U;
    }
#endif
    for (; i + U <= minlen; i += U) body(std::false_type{}, i);
    for (; i < maxlen; i += U) body(std::true_type{}, i);
    if (merged) {  // warp-uniform: fold the 4 chunk partials (fixed tree order), the leader stores one slot
#pragma unroll
        for (int off = kLanesPerEdge; off < 32; off <<= 1) {
#pragma unroll
            for (int a = 0; a < NA; a++)
#pragma unroll
                for (int k = 0; k < V; k++)
                    acc[a].v[k] = nary_fwd_dyn<T>(acc_op(N::red, a), acc[a].v[k],
                                                  __shfl_xor_sync(0xffffffffu, acc[a].v[k], off));
        }
        if (sg == 0 && st_ok) {
#pragma unroll
            for (int a = 0; a < NA; a++)
                st_plain(p.partial + (size_t)a * p.partial_plane + (size_t)slot * p.D + c0, acc[a]);
        }
    } else if (slot >= 0 && st_ok) {  // single chunk of a long segment
#pragma unroll
        for (int a = 0; a < NA; a++)
            st_plain(p.partial + (size_t)a * p.partial_plane + (size_t)slot * p.D + c0, acc[a]);
    }
}

template <typename T, int MODE, int SUM, int MUL, int SM>
__device__ __forceinline__ void walk_items(const SegParams<T> &p, int wi, int tile, uint32_t rel_s_addr,
                                           uint32_t g1_s_addr, uint32_t meta_s_addr, bool use_w) {
    if (use_w) walk_items_impl<T, MODE, SUM, MUL, SM, false>(p, wi, tile, rel_s_addr, g1_s_addr, meta_s_addr);
    else walk_items_impl<T, MODE, SUM, MUL, SM, true>(p, wi, tile, rel_s_addr, g1_s_addr, meta_s_addr);
}

// Stage `nrows` 128-byte row slices (row r at src + r * row_bytes) into shared memory at smem_addr + r * 128.
//   full tile   : one TMA bulk copy (cp.async.bulk -> UBLKCP) per row; completion is signalled on the CTA's mbarrier
//                 (the caller has already posted expect_tx for the bytes and waits on it afterwards)
//   ragged tile : 128-bit vector loads with zero fill (caller __syncthreads afterwards)
template <typename T>
__device__ __forceinline__ void stage_rows(uint32_t smem_addr, const char *src, int nrows, unsigned row_bytes,
                                           bool full, int valid_cols, uint32_t bar_addr) {
    constexpr int V = VecN<T>::value;
    if (full) {
        for (int r = threadIdx.x; r < nrows; r += kCtaThreads)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_addr + (uint32_t)r * kTileBytes),
                           "l"(src + (unsigned long long)(unsigned)r * row_bytes), "r"(kTileBytes), "r"(bar_addr)
                         : "memory");
    } else {
        for (int idx = threadIdx.x; idx < nrows * kLanesPerEdge; idx += kCtaThreads) {
            const int r = idx >> 3, qq = idx & 7;
            Vec<T> x;
#pragma unroll
            for (int k = 0; k < V; k++) x.v[k] = T(0);
            if (qq * V < valid_cols) x = ld_gather(reinterpret_cast<const T *>(src + (size_t)r * row_bytes + qq * 16));
            if constexpr (sizeof(T) == 4)
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(smem_addr + (uint32_t)idx * 16u),
                             "f"(x.v[0]), "f"(x.v[1]), "f"(x.v[2]), "f"(x.v[3]) : "memory");
            else
                asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(smem_addr + (uint32_t)idx * 16u), "d"(x.v[0]),
                             "d"(x.v[1]) : "memory");
        }
    }
}

// SM: what lives in shared memory for the current column tile
//   0  nothing (relation rows come through L1/L2)
//   1  the relation tile (R x 128 B)
//   2  the relation tile (if the op needs it) AND the gathered matrix' tile (N x 128 B: input for FWD, output_grad
//      for the backward passes) - small graphs such as ULTRA's relation graph (N = #relations <= ~800): every gather
//      becomes an LDS.128 (half the L1 data-pipe cost of an LDG.128) and L2 sees only the staging traffic
template <typename T, int MODE, int SUM, int MUL, int SM>
__global__ void __launch_bounds__(kCtaThreads, kCtasPerSm) seg_reduce_kernel(const SegParams<T> p) {
    using N = Needs<MODE, SUM, MUL>;
    constexpr int V = VecN<T>::value;
    constexpr int TILE = kLanesPerEdge * V;
    constexpr bool kRelS = N::rel && SM >= 1;
    constexpr bool kG1S = SM == 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    const uint32_t rel_s_addr = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t g1_s_addr = rel_s_addr + (kRelS ? (uint32_t)p.R * kTileBytes : 0u);
    // per-warp metadata staging area sits behind the tiles: [relation tile][gathered tile][16 x 2 KB metadata]
    const uint32_t meta_s_addr = g1_s_addr + (kG1S ? (uint32_t)p.N * kTileBytes : 0u) + (uint32_t)warp * kMetaStageBytes;
    __shared__ __align__(8) unsigned long long slab_bar;  // mbarrier of the TMA-staged tiles
    const uint32_t slab_bar_addr = (uint32_t)__cvta_generic_to_shared(&slab_bar);
    unsigned slab_phase = 0;
    if (kRelS || kG1S) {
        if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(slab_bar_addr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    const bool use_w = p.w != nullptr && __ldg(p.w_nonunit) != 0;  // uniform over the grid
    int cur_tile = -1;

    // Work queue: blocks of `run` consecutive units are handed out in tile-major order by a global ticket counter, so a
    // CTA that drew short items simply takes the next block (static round-robin left a ~1/waves tail: 15 % at 32 heads
    // per GPU, where a CTA only sees ~5 blocks).  The order in which tiles are visited - what keeps the gathered tiles
    // L2-resident - is unchanged.  Results do not depend on which CTA computes a unit.
    // The next ticket is drawn at the START of a block and published at its end, so the atomic's round trip (~1 us,
    // 19 blocks per CTA at C2 = 1 % of the call when it was exposed) overlaps the block's work.
    const int num_warp_items = __ldg(p.counts) / kSubPerWarp;
    const int units_per_tile = (num_warp_items + kCtaWarps * p.items_per_warp - 1) / (kCtaWarps * p.items_per_warp);
    const long long total_units = (long long)units_per_tile * p.num_tiles;
    // p.sched == nullptr selects a static round-robin (blocks blockIdx.x, + gridDim.x, ...; tunable "work_queue" = 0), kept
    // for A/B runs only: without the queue the CTAs drift apart, the window of active tiles outgrows its L2 budget and the
    // forward pass re-reads 3x the compulsory bytes from DRAM (2.35 vs 1.96 ms at C2, profiles/r02_work_queue_ab.txt)
    const bool dynamic = p.sched != nullptr;
    __shared__ int s_ticket;
    if (dynamic) {
        if (threadIdx.x == 0) s_ticket = atomicAdd(p.sched, 1);
        __syncthreads();
    }
    long long blk = dynamic ? (long long)s_ticket : (long long)blockIdx.x;
    for (;;) {
        if (blk * p.run >= total_units) break;
        int next_ticket = 0;
        if (dynamic && threadIdx.x == 0) next_ticket = atomicAdd(p.sched, 1);  // consumed only after the block (no stall here)
        for (int uu = 0; uu < p.run; uu++) {
            const long long unit = blk * p.run + uu;
            if (unit >= total_units) break;
            const int tile = (int)(unit / units_per_tile);
            const int chunk = (int)(unit - (long long)tile * units_per_tile);
            if ((kRelS || kG1S) && tile != cur_tile) {
                __syncthreads();  // everyone is done reading the previous tile's rows
                const int grp = tile / p.lay.tpg, tin = tile - grp * p.lay.tpg;
                const bool full = (tin + 1) * TILE <= p.lay.group_cols;
                const int valid_cols = p.lay.group_cols - tin * TILE;
                const char *rel_src = reinterpret_cast<const char *>(p.relation) +
                                      (size_t)grp * p.lay.grp_rel_bytes + (size_t)tin * kTileBytes;
                const char *g1_src = reinterpret_cast<const char *>(MODE == FWD ? p.input : p.outgrad) +
                                     (size_t)grp * p.lay.grp_node_bytes + (size_t)tin * kTileBytes;
                if (full) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads -> async writes
                    if (threadIdx.x == 0)
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(slab_bar_addr),
                                     "r"(((kRelS ? p.R : 0) + (kG1S ? p.N : 0)) * kTileBytes) : "memory");
                }
                if (kRelS) stage_rows<T>(rel_s_addr, rel_src, p.R, p.lay.row_bytes, full, valid_cols, slab_bar_addr);
                if (kG1S) stage_rows<T>(g1_s_addr, g1_src, p.N, p.lay.row_bytes, full, valid_cols, slab_bar_addr);
                if (full) {
                    unsigned done = 0;
                    while (!done)
                        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; "
                                     "selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(slab_bar_addr), "r"(slab_phase)
                                     : "memory");
                    slab_phase ^= 1u;
                } else {
                    __syncthreads();
                }
                cur_tile = tile;
            }
#pragma unroll 1
            for (int k = 0; k < p.items_per_warp; k++) {
                const int wi = (chunk * p.items_per_warp + k) * kCtaWarps + warp;
                if (wi < num_warp_items) walk_items<T, MODE, SUM, MUL, SM>(p, wi, tile, rel_s_addr, g1_s_addr, meta_s_addr, use_w);
            }
        }
        if (dynamic) {
            __syncthreads();  // every thread has read this block's ticket
            if (threadIdx.x == 0) s_ticket = next_ticket;
            __syncthreads();
            blk = s_ticket;
        } else {
            blk += gridDim.x;
        }
    }
    // the last CTA to leave resets the queue for the next launch on this plan (launches of one plan are stream-ordered)
    if (dynamic && threadIdx.x == 0) {
        if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
            p.sched[0] = 0;
            p.sched[1] = 0;
            __threadfence();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// byte offset of (row `seg`, global column c0) in a matrix with layout `lay` and group stride grp_bytes
template <typename T>
__device__ __forceinline__ size_t dst_offset(const Layout &lay, long long grp_bytes, int c0, int seg) {
    const int grp = c0 / lay.group_cols, cin = c0 - grp * lay.group_cols;
    return (size_t)grp * grp_bytes + (size_t)seg * lay.row_bytes + (size_t)cin * sizeof(T);
}

// fix-up, part 1: long segments fold their chunk-group partials.  One 256-thread CTA per (segment, 128-column
// chunk): warp w folds partials w, w+8, ... in order, then warp 0 folds the 8 warp results in order -> deterministic.
template <typename T, int RED>
__global__ void __launch_bounds__(256) fold_partials_kernel(const int4 *__restrict__ fix, const int *__restrict__ counts,
                                                            const T *__restrict__ partial, T *__restrict__ dst,
                                                            const T *__restrict__ boundary, int D, Layout lay,
                                                            long long grp_bytes, int ncc) {
    constexpr int V = VecN<T>::value;
    __shared__ Vec<T> part[8][32];
    const int num_fix_long = __ldg(counts + 2);  // the grid is sized for an upper bound (block-uniform exit)
    if (num_fix_long == 0 || blockIdx.x >= (unsigned)num_fix_long * (unsigned)ncc) return;
    const int f = blockIdx.x % num_fix_long;
    const int cc = blockIdx.x / num_fix_long;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = cc * 32 * V + lane * V;
    const bool ok = c0 < D;
    const int4 fx = __ldg(fix + f);
    Vec<T> acc;
#pragma unroll
    for (int k = 0; k < V; k++) acc.v[k] = nary_zero<RED, T>();
    if (ok) {
        const T *src = partial + (size_t)fx.y * D + c0;
#pragma unroll 4
        for (int s = warp; s < fx.z; s += 8) {
            const Vec<T> x = *reinterpret_cast<const Vec<T> *>(src + (size_t)s * D);
#pragma unroll
            for (int k = 0; k < V; k++) acc.v[k] = nary_fwd<RED>(acc.v[k], x.v[k]);
        }
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && ok) {
#pragma unroll
        for (int w = 1; w < 8; w++) {
            const Vec<T> x = part[w][lane];
#pragma unroll
            for (int k = 0; k < V; k++) acc.v[k] = nary_fwd<RED>(acc.v[k], x.v[k]);
        }
        const size_t off = dst_offset<T>(lay, grp_bytes, c0, fx.x);
        if (boundary != nullptr) {
            const Vec<T> bv = *reinterpret_cast<const Vec<T> *>(reinterpret_cast<const char *>(boundary) + off);
#pragma unroll
            for (int k = 0; k < V; k++) acc.v[k] = nary_fwd<RED>(acc.v[k], bv.v[k]);
        }
        st_stream(reinterpret_cast<T *>(reinterpret_cast<char *>(dst) + off), acc);
    }
}

// fix-up, part 2: empty segments get the Nary identity (operator.cuh:53,66,79).  One warp per (segment, 128 columns).
template <typename T, int RED>
__global__ void __launch_bounds__(256) fill_identity_kernel(const int4 *__restrict__ fix_all,
                                                            const int *__restrict__ counts, T *__restrict__ dst,
                                                            const T *__restrict__ boundary, int D, Layout lay,
                                                            long long grp_bytes) {
    constexpr int V = VecN<T>::value;
    const int ncc = (D + 32 * V - 1) / (32 * V);
    const int num_long = __ldg(counts + 2), num_fix = __ldg(counts + 1) - num_long;  // empty segments follow the long ones
    const int4 *fix = fix_all + num_long;
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)num_fix * ncc, stride = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; gw < warps; gw += stride) {
        const int f = (int)(gw % num_fix);
        const int cc = (int)(gw / num_fix);
        const int c0 = cc * 32 * V + lane * V;
        if (c0 >= D) continue;
        Vec<T> acc;
#pragma unroll
        for (int k = 0; k < V; k++) acc.v[k] = nary_zero<RED, T>();
        const size_t off = dst_offset<T>(lay, grp_bytes, c0, __ldg(fix + f).x);
        if (boundary != nullptr) {  // Nary(identity, boundary) == boundary
            acc = *reinterpret_cast<const Vec<T> *>(reinterpret_cast<const char *>(boundary) + off);
        }
        st_stream(reinterpret_cast<T *>(reinterpret_cast<char *>(dst) + off), acc);
    }
}

// ---------------------------------------------------------------------------------------------------------
// weight_grad[e] = sum_d og * dout_dy * x   (rspmm.cu:137-147): one warp per edge of the by-row order.
template <typename T, int SUM, int MUL>
__global__ void __launch_bounds__(256) wgrad_kernel(const int2 *__restrict__ meta, const int *__restrict__ eid,
                                                    const int *__restrict__ row_of_edge,
                                                    const T *__restrict__ w, const T *__restrict__ relation,
                                                    const T *__restrict__ input, const T *__restrict__ output,
                                                    const T *__restrict__ outgrad, T *__restrict__ wgrad,
                                                    const int *__restrict__ counts, int D) {
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= __ldg(counts + 7)) return;  // launched for the base graph's edge count
    const int e = (int)gw;
    const int2 m = __ldg(meta + e);
    const int col = m.x, type = m.y & 0x7fffffff, row = __ldg(row_of_edge + e);
    const T wv = w ? __ldg(w + e) : T(1);
    T s = T(0);
    for (int d = lane; d < D; d += 32) {
        const T rel = __ldg(relation + (size_t)type * D + d);
        const T in = __ldg(input + (size_t)col * D + d);
        const T x = bin_fwd<MUL>(rel, in);
        T g = __ldg(outgrad + (size_t)row * D + d);
        if (SUM != S_ADD) g = g * nary_bwd<SUM>(__ldg(output + (size_t)row * D + d), wv * x);
        s += g * x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) wgrad[__ldg(eid + e)] = s;
}

// ---------------------------------------------------------------------------------------------------------
// generic path (any D / alignment): one thread per (segment, column), edges folded sequentially in plan order.
template <typename T, int MODE, int SUM, int MUL>
__global__ void __launch_bounds__(256) seg_reduce_generic_kernel(const int2 *__restrict__ meta,
                                                                 const T *__restrict__ w,
                                                                 const int *__restrict__ seg_ptr, int num_seg,
                                                                 const T *__restrict__ relation,
                                                                 const T *__restrict__ input,
                                                                 const T *__restrict__ output,
                                                                 const T *__restrict__ outgrad, T *__restrict__ dst,
                                                                 int D) {
    using N = Needs<MODE, SUM, MUL>;
    const long long total = (long long)num_seg * D;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int seg = (int)(idx / D);
        const int d = (int)(idx - (long long)seg * D);
        T acc = nary_zero<N::red, T>();
        const int e_begin = seg_ptr[seg], e_end = seg_ptr[seg + 1];
        for (int e = e_begin; e < e_end; e++) {
            const int2 m = meta[e];
            const int a = m.x, b = m.y & 0x7fffffff;
            const T wv = w ? w[e] : T(1);
            if (MODE == FWD) {
                const T y = wv * bin_fwd<MUL>(relation[(size_t)b * D + d], input[(size_t)a * D + d]);
                acc = nary_fwd<SUM>(acc, y);
            } else if (MODE == BWD_IN) {
                T g = outgrad[(size_t)a * D + d];
                const T r = N::rel ? relation[(size_t)b * D + d] : T(0);
                if (N::out) {
                    const T y = wv * bin_fwd<MUL>(r, input[(size_t)seg * D + d]);
                    g = g * nary_bwd<SUM>(output[(size_t)a * D + d], y);
                }
                g = g * wv;
                if (MUL == M_MUL) g = g * r;
                acc += g;
            } else {
                T g = outgrad[(size_t)a * D + d];
                const T in = N::inb ? input[(size_t)b * D + d] : T(0);
                if (N::out) {
                    const T y = wv * bin_fwd<MUL>(relation[(size_t)seg * D + d], in);
                    g = g * nary_bwd<SUM>(output[(size_t)a * D + d], y);
                }
                g = g * wv;
                if (MUL == M_MUL) g = g * in;
                acc += g;
            }
        }
        dst[idx] = acc;
    }
}

template <typename T>
__global__ void permute_weight_kernel(const T *__restrict__ w, const int *__restrict__ eid0,
                                      const int *__restrict__ eid1, const int *__restrict__ eid2,
                                      T *__restrict__ w0, T *__restrict__ w1, T *__restrict__ w2,
                                      const int *__restrict__ counts, int *__restrict__ nonunit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= __ldg(counts + 7)) return;  // launched for the base graph's edge count
    if (w[i] != T(1)) *nonunit = 1;  // benign race: every writer stores the same value
    w0[i] = w[eid0[i]];
    w1[i] = w[eid1[i]];
    w2[i] = w[eid2[i]];
}

}  // namespace ultra
