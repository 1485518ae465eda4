// rspmm_plan.cu -- graph plan builder: the one-time replacement for the reference's per-call
//   key.argsort() + 3 gathers      (/root/reference/ultra/rspmm/rspmm.py:174-178)
//   sortedness assert + host sync  (/root/reference/ultra/rspmm/rspmm.py:15-17)
//   ind2ptr scatter_add + cumsum   (/root/reference/ultra/rspmm/source/rspmm.cpp:40-48)
// Three stable radix sorts (CUB, device) give the by-row / by-col / by-relation edge orders as int32 CSR;
// the edge-balanced work-item lists are cut on the host from the three segment-pointer arrays.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <climits>

#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

#include "rspmm_kernels.cuh"
#include "rspmm_plan.h"

namespace ultra {

namespace {

constexpr size_t kAlign = 256;
#ifndef ULTRA_CUT_BLOCK
#define ULTRA_CUT_BLOCK 64
#endif
constexpr int kCutBlock = ULTRA_CUT_BLOCK;    // segments per cutter thread (see cut_walk_block)
constexpr int kPlanCounts = 8;   // ints per order in the device-side result block (cut_finalize_kernel)
struct CutCounts { int whole, groups, chunks, slots, nlong, nempty, nonempty; };  // per cutter block / running totals
struct LongRec { int seg, b, e, j, slot0, group0, chunk0, pad; };  // a long segment + where its items / slots go
inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }
inline int bits_for(int64_t n) {  // bits needed to represent values in [0, n)
    int b = 1;
    while ((int64_t(1) << b) < n) b++;
    return b;
}
// item capacity: whole items <= 2E/64 + 1 (two consecutive overflow-closed items hold > 64 edges) + one forced close per
// cutter block and per long segment; chunk items <= E/64 + #long; group items <= E/64; #long <= E/65
inline int64_t max_items(int64_t E, int64_t S) { return 5 * ((E + kItemEdges - 1) / kItemEdges) + S / kCutBlock + 16; }

struct WsLayout {
    size_t meta[3], eid[3], seg_ptr[3], nonempty[3], items[3], fix[3], w[3];
    size_t row_of_edge, keysA, keysB, valsA, valsB, flags, cub_temp, cub_temp_bytes, total;
    // device cutter / mask temporaries, one set PER ORDER (the three orders are processed on three streams), + result counts
    size_t cut_cnt[3], cut_incl[3], rkeys[3][2], ritems[3][2], cut_temp[3], cut_temp_bytes, counts, longs[3];
    size_t meta_m[3], eid_m[3], seg_ptr_m[3], row_of_edge_m, mask_pos[3];  // masked views (plan_apply_mask)
    int64_t max_blocks, rest_cap;
};

WsLayout make_layout(int64_t N, int64_t E, int64_t R) {
    WsLayout L;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(std::max<size_t>(bytes, 16)); return o; };
    const int64_t S[3] = {N, N, R};
    for (int o = 0; o < 3; o++) {
        L.meta[o] = take((size_t)E * 8);
        L.eid[o] = take((size_t)E * 4);
        L.seg_ptr[o] = take((size_t)(S[o] + 1) * 4);
        L.nonempty[o] = take((size_t)(S[o] + 1) * 4);
        L.items[o] = take((size_t)max_items(E, S[o]) * 16);
        L.fix[o] = take((size_t)std::max<int64_t>(S[o], 1) * 16);
        L.w[o] = take((size_t)E * 8);
    }
    L.row_of_edge = take((size_t)E * 4);
    L.flags = take(64);  // [0] range error, [1] unsorted, [2] w_nonunit, [4..9] work-queue tickets of the three orders
    L.cub_temp_bytes = (size_t)(1 << 20) + (size_t)E;
    L.cub_temp = take(L.cub_temp_bytes);
    const int64_t Smax = std::max(N, R);
    L.max_blocks = (Smax + kCutBlock - 1) / kCutBlock + 1;
    L.rest_cap = max_items(E, Smax);
    L.cut_temp_bytes = (size_t)(2 << 20) + (size_t)E / 32;
    for (int o = 0; o < 3; o++) {
        L.cut_cnt[o] = take((size_t)L.max_blocks * sizeof(CutCounts));
        L.cut_incl[o] = take((size_t)L.max_blocks * sizeof(CutCounts));
        for (int k = 0; k < 2; k++) {
            L.rkeys[o][k] = take((size_t)L.rest_cap * 8);
            L.ritems[o][k] = take((size_t)L.rest_cap * 16);
        }
        L.cut_temp[o] = take(L.cut_temp_bytes);
        L.longs[o] = take((size_t)std::min<int64_t>(S[o], E / (kItemEdges + 1) + 1) * sizeof(LongRec));
        L.mask_pos[o] = take((size_t)E * 4);
    }
    L.counts = take(3 * kPlanCounts * sizeof(int));
    for (int o = 0; o < 3; o++) {
        L.meta_m[o] = take((size_t)E * 8);
        L.eid_m[o] = take((size_t)E * 4);
        L.seg_ptr_m[o] = take((size_t)(S[o] + 1) * 4);
    }
    L.row_of_edge_m = take((size_t)E * 4);
    // the edge sorts' key / value double buffers are only alive during plan_build, the masked views only after it:
    // they share memory (the sort scratch used to stay reserved for the plan's whole life)
    L.keysA = L.meta_m[0]; L.keysB = L.meta_m[1]; L.valsA = L.eid_m[0]; L.valsB = L.eid_m[1];
    L.total = off;
    return L;
}

__global__ void validate_kernel(const int64_t *__restrict__ ei, const int64_t *__restrict__ et, int64_t N,
                                int64_t E, int64_t R, int *flags) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t row = ei[e], col = ei[E + e], t = et[e];
    if (row < 0 || row >= N || col < 0 || col >= N || t < 0 || t >= R) atomicOr(flags + 0, 1);
    if (e > 0) {
        const int64_t prow = ei[e - 1], pcol = ei[E + e - 1];
        if (prow > row || (prow == row && pcol > col)) atomicOr(flags + 1, 1);  // rspmm.py:15-17 key order
    }
}

// order 0: (row, col)   order 1: (col, row)   order 2: (type, row).
// The secondary key matters for speed, not correctness: with every segment's edges sorted by the gathered node id,
// all warps sweep the gathered matrix from low to high rows roughly in step, which is what gives the gathers their
// L1/L2 temporal locality (measured: grouping by relation type instead costs +40% forward time on a uniform graph,
// +20 % on the Zipf one).  Graphs small enough for the gathered tile to sit in shared memory have no such locality to
// lose: their orders 0 and 1 are (row, type, col) / (col, type, row).
__global__ void make_keys_kernel(int order, const int64_t *__restrict__ ei, const int64_t *__restrict__ et,
                                 int64_t E, int64_t N, int64_t R, int sec_bits, int node_bits, unsigned long long *keys,
                                 int *vals) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    // out-of-range values (flagged by validate_kernel; the plan is rejected at the end of the build) are clamped here so
    // that the kernels enqueued behind this one stay inside their buffers without an early host synchronisation
    auto clampi = [](int64_t v, int64_t hi) { return (unsigned long long)(v < 0 ? 0 : (v >= hi ? hi - 1 : v)); };
    const unsigned long long row = clampi(ei[e], N), col = clampi(ei[E + e], N);
    unsigned long long pri, sec;
    if (order == 0) { pri = row; sec = col; }
    else if (order == 1) { pri = col; sec = row; }
    else { pri = clampi(et[e], R); sec = row; }
    // small graphs (tunable "plan_relation_runs"): relation type between the primary and the node key, so that a segment's
    // edges form runs of one relation type and the kernel re-reads the relation slice only when the type changes
    // (sec_bits > node_bits only then; see walk_items_impl)
    if (order != 2 && sec_bits > node_bits) sec |= clampi(et[e], R) << node_bits;
    keys[e] = (pri << sec_bits) | sec;
    vals[e] = (int)e;
}

__global__ void build_meta_kernel(int order, const unsigned long long *__restrict__ keys,
                                  const int *__restrict__ vals, const int64_t *__restrict__ ei,
                                  const int64_t *__restrict__ et, int64_t E, int S, int sec_bits, int2 *meta,
                                  int *eid, int *seg_ptr, int *row_of_edge) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    const int id = vals[i];
    const int pri = (int)(keys[i] >> sec_bits);
    const int prev = i > 0 ? (int)(keys[i - 1] >> sec_bits) : -1;
    const int next = i + 1 < E ? (int)(keys[i + 1] >> sec_bits) : S;
    const int row = (int)ei[id], col = (int)ei[E + id], t = (int)et[id];
    int a, b;
    if (order == 0) { a = col; b = t; }        // gather input[col], relation[type]
    else if (order == 1) { a = row; b = t; }   // gather outgrad[row], relation[type]
    else { a = row; b = col; }                 // gather outgrad[row], input[col]
    if (next != pri) b |= 0x80000000;          // last edge of its segment
    meta[i] = make_int2(a, b);
    eid[i] = id;
    if (order == 0) row_of_edge[i] = row;
    for (int s = prev + 1; s <= pri; s++) seg_ptr[s] = (int)i;
    if (i == E - 1)
        for (int s = pri + 1; s <= S; s++) seg_ptr[s] = (int)E;
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t err__ = (expr);                                                            \
        if (err__ != cudaSuccess) {                                                            \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(err__));                  \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------------
// Work-item cutter.  One rule, two executors: the device kernels below (the product path: plan_build never cuts on the
// host any more) and a host mirror of the same functions (ultra_rspmm_debug_cut_items, CPU test-suite).
//   * the segment array is processed in blocks of kCutBlock (64) segments, one thread each; a block's whole-segment
//     items never reach into the next block (costs < 1 extra item per 64 segments and makes the greedy packing parallel)
//   * whole segments are packed greedily up to kItemEdges edges per item                       {e0, e1, j0,  -1}
//   * a segment longer than kItemEdges is cut into 64-edge chunks.  Every run of 4 FULL chunks becomes a merged group
//     walked by one warp and combined with shuffles: leader {e0, e1, j|kMergedBit, slot}, followers {.., .., j|kMergedBit, -2};
//     the remaining (<= 3 full + 1 partial) chunks are single chunk items                     {e0, e1, j,  slot}
//     Partial slots of one segment are consecutive and in edge order (groups, then single chunks).
// Item order: merged groups first (a multiple of 4, so a warp never mixes a group with other items), then all other
// items by (length descending, first edge ascending), so the 4 items a warp walks in lock-step have (nearly) equal
// lengths; padded with empty items to a multiple of 4.  fix list: long segments (in order), then empty segments.
template <typename V>
__host__ __device__ inline void cut_walk_block(const int *ptr, int s0, int s1, V &v) {
    int cur_begin = -1, cur_end = -1, cur_j = 0, nj = 0;  // nj: non-empty segments of this block seen so far
    for (int s = s0; s < s1; s++) {
        const int b = ptr[s], e = ptr[s + 1], deg = e - b;
        if (deg == 0) { v.empty(s); continue; }
        const int j = nj++;
        v.nonempty(s);
        if (deg > kItemEdges) {
            if (cur_begin >= 0) { v.whole(cur_begin, cur_end, cur_j); cur_begin = -1; }
            v.long_seg(s, b, e, j);
        } else {
            if (cur_begin >= 0 && e - cur_begin > kItemEdges) { v.whole(cur_begin, cur_end, cur_j); cur_begin = -1; }
            if (cur_begin < 0) { cur_begin = b; cur_j = j; }
            cur_end = e;
        }
    }
    if (cur_begin >= 0) v.whole(cur_begin, cur_end, cur_j);
}

__host__ __device__ inline CutCounts cut_add(const CutCounts &a, const CutCounts &b) {
    CutCounts r;
    r.whole = a.whole + b.whole; r.groups = a.groups + b.groups; r.chunks = a.chunks + b.chunks;
    r.slots = a.slots + b.slots; r.nlong = a.nlong + b.nlong; r.nempty = a.nempty + b.nempty;
    r.nonempty = a.nonempty + b.nonempty;
    return r;
}
struct CutSum {
    __host__ __device__ CutCounts operator()(const CutCounts &a, const CutCounts &b) const { return cut_add(a, b); }
};

struct CountVisitor {
    CutCounts c{0, 0, 0, 0, 0, 0, 0};
    __host__ __device__ void empty(int) { c.nempty++; }
    __host__ __device__ void nonempty(int) { c.nonempty++; }
    __host__ __device__ void whole(int, int, int) { c.whole++; }
    __host__ __device__ void long_seg(int, int b, int e, int) {
        const int deg = e - b, groups = deg / (kItemEdges * kSubPerWarp);
        const int rest = (deg - groups * kItemEdges * kSubPerWarp + kItemEdges - 1) / kItemEdges;
        c.groups += groups; c.chunks += rest; c.slots += groups + rest; c.nlong++;
    }
};

__host__ __device__ inline unsigned long long rest_key(int e0, int e1) {  // (length descending, first edge ascending)
    return ((unsigned long long)(kItemEdges - (e1 - e0)) << 32) | (unsigned)e0;
}

struct WriteVisitor {
    CutCounts off;      // exclusive offsets of this block
    CutCounts tot;      // grand totals
    int4 *items;        // merged-group items go straight to their final place
    unsigned long long *rkeys;
    int4 *ritems;       // whole-segment items [0, tot.whole), single chunks [tot.whole, tot.whole + tot.chunks)
    int4 *fix;
    int *nonempty_out;
    LongRec *longs;
    int w = 0, g = 0, c = 0, sl = 0, l = 0, em = 0, ne = 0;
    __host__ __device__ void empty(int s) { fix[tot.nlong + off.nempty + em++] = make_int4(s, 0, 0, 0); }
    __host__ __device__ void nonempty(int s) { nonempty_out[off.nonempty + ne++] = s; }
    __host__ __device__ void whole(int e0, int e1, int j) {
        const int i = off.whole + w++;
        ritems[i] = make_int4(e0, e1, off.nonempty + j, -1);
        rkeys[i] = rest_key(e0, e1);
    }
    __host__ __device__ void long_seg(int s, int b, int e, int j) {
        // only the bookkeeping here: the segment's chunk / group items are written by emit_long, one CTA per long
        // segment (a hub row of 10^5 edges is thousands of items - far too many for the one thread that walks this block)
        const int deg = e - b, groups = deg / (kItemEdges * kSubPerWarp);
        const int rest = (deg - groups * kItemEdges * kSubPerWarp + kItemEdges - 1) / kItemEdges;
        const int slot0 = off.slots + sl;
        longs[off.nlong + l] = LongRec{s, b, e, off.nonempty + j, slot0, off.groups + g, off.chunks + c, 0};
        fix[off.nlong + l] = make_int4(s, slot0, groups + rest, 0);
        l++; sl += groups + rest; g += groups; c += rest;
    }
};

// items of one long segment; thread t of nt (device: a CTA per segment; host mirror: t = 0, nt = 1)
__host__ __device__ inline void emit_long(const LongRec &r, const CutCounts &tot, int4 *items, unsigned long long *rkeys,
                                          int4 *ritems, int t, int nt) {
    const int kGroup = kItemEdges * kSubPerWarp;
    const int groups = (r.e - r.b) / kGroup;
    for (int k = t; k < groups * kSubPerWarp; k += nt) {
        const int gi = k / kSubPerWarp, tt = k % kSubPerWarp, cpos = r.b + gi * kGroup;
        items[(size_t)(r.group0 + gi) * kSubPerWarp + tt] =
            make_int4(cpos + tt * kItemEdges, cpos + (tt + 1) * kItemEdges, r.j | kMergedBit, tt ? -2 : r.slot0 + gi);
    }
    const int cstart = r.b + groups * kGroup, rest = (r.e - cstart + kItemEdges - 1) / kItemEdges;
    for (int k = t; k < rest; k += nt) {
        const int cpos = cstart + k * kItemEdges, e1 = cpos + kItemEdges < r.e ? cpos + kItemEdges : r.e;
        const int i = tot.whole + r.chunk0 + k;
        ritems[i] = make_int4(cpos, e1, r.j, r.slot0 + groups + k);
        rkeys[i] = rest_key(cpos, e1);
    }
}

__global__ void cut_long_kernel(const LongRec *__restrict__ longs, const CutCounts *__restrict__ incl, int nb, int4 *items,
                                unsigned long long *rkeys, int4 *ritems) {
    const CutCounts tot = incl[nb - 1];
    for (int li = blockIdx.x; li < tot.nlong; li += gridDim.x) emit_long(longs[li], tot, items, rkeys, ritems, threadIdx.x, blockDim.x);
}

__global__ void cut_count_kernel(const int *__restrict__ ptr, int S, CutCounts *__restrict__ cnt, int nb) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    CountVisitor v;
    const int s0 = b * kCutBlock, s1 = min(S, s0 + kCutBlock);
    if (s0 < s1) cut_walk_block(ptr, s0, s1, v);
    cnt[b] = v.c;
}

__global__ void cut_write_kernel(const int *__restrict__ ptr, int S, const CutCounts *__restrict__ incl, int nb,
                                 int4 *items, unsigned long long *rkeys, int4 *ritems, int4 *fix, int *nonempty,
                                 LongRec *longs) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    WriteVisitor v;
    v.longs = longs;
    v.off = b ? incl[b - 1] : CutCounts{0, 0, 0, 0, 0, 0, 0};
    v.tot = incl[nb - 1];
    v.items = items; v.rkeys = rkeys; v.ritems = ritems; v.fix = fix; v.nonempty_out = nonempty;
    const int s0 = b * kCutBlock, s1 = min(S, s0 + kCutBlock);
    if (s0 < s1) cut_walk_block(ptr, s0, s1, v);
}

// sorted rest items -> behind the merged groups; pad to a multiple of 4; result counts for the launchers
__global__ void cut_finalize_kernel(const CutCounts *__restrict__ incl, int nb, const int4 *__restrict__ sorted, int4 *items,
                                    int *nonempty, int *counts, int num_edge, const int *num_edge_dev) {
    const CutCounts tot = incl[nb - 1];
    const int G4 = tot.groups * kSubPerWarp, rest = tot.whole + tot.chunks;
    const int padded = (rest + kSubPerWarp - 1) / kSubPerWarp * kSubPerWarp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += gridDim.x * blockDim.x)
        items[G4 + i] = i < rest ? sorted[i] : make_int4(0, 0, 0, -1);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        nonempty[tot.nonempty] = 0;  // pad: the walker prefetches nonempty[j + 1]
        counts[0] = G4 + padded;            // num_items
        counts[1] = tot.nlong + tot.nempty;  // num_fix
        counts[2] = tot.nlong;               // num_fix_long
        counts[3] = tot.slots;               // num_slots
        counts[4] = tot.nonempty + 1;        // entries of nonempty (incl. the pad)
        counts[5] = tot.groups;
        counts[6] = rest;
        counts[7] = num_edge_dev ? *num_edge_dev : num_edge;  // edges of the (possibly masked) graph
    }
}

// Fork the caller's stream into three lanes (lane 0 IS the caller's stream, lanes 1-2 are plan-owned side streams created
// on first use) and join them back; works under CUDA-graph capture (the lanes become parallel branches of the graph).
struct ForkJoin {
    ultra_rspmm_plan *plan;
    cudaStream_t main;
    int rc = 0;
    ForkJoin(ultra_rspmm_plan *p, cudaStream_t s) : plan(p), main(s) {
        for (int i = 0; i < 2 && rc == 0; i++) {
            if (plan->side[i] == nullptr && cudaStreamCreateWithFlags(&plan->side[i], cudaStreamNonBlocking) != cudaSuccess) rc = 2;
            if (rc == 0 && plan->side_ev[i] == nullptr &&
                cudaEventCreateWithFlags(&plan->side_ev[i], cudaEventDisableTiming) != cudaSuccess) rc = 2;
        }
        if (rc == 0 && plan->fork_ev == nullptr && cudaEventCreateWithFlags(&plan->fork_ev, cudaEventDisableTiming) != cudaSuccess) rc = 2;
        if (rc == 0 && cudaEventRecord(plan->fork_ev, main) != cudaSuccess) rc = 2;
        for (int i = 0; i < 2 && rc == 0; i++)
            if (cudaStreamWaitEvent(plan->side[i], plan->fork_ev, 0) != cudaSuccess) rc = 2;
        if (rc) set_error("plan: could not fork the side streams");
    }
    cudaStream_t lane(int o) const { return o == 0 ? main : plan->side[o - 1]; }
    int join() {
        for (int i = 0; i < 2; i++) {
            if (cudaEventRecord(plan->side_ev[i], plan->side[i]) != cudaSuccess ||
                cudaStreamWaitEvent(main, plan->side_ev[i], 0) != cudaSuccess) {
                set_error("plan: could not join the side streams");
                return 2;
            }
        }
        return 0;
    }
};

// Device cutter for one segment order: count -> scan -> write -> sort the non-group items -> finalize.  No host sync.
int cut_items_device(int o, const int *seg_ptr, int S, int64_t E, const WsLayout &L, char *ws, SegOrder &so, int *counts,
                     cudaStream_t stream, const int *num_edge_dev = nullptr) {
    const int nb = std::max(1, (S + kCutBlock - 1) / kCutBlock);
    auto *cnt = reinterpret_cast<CutCounts *>(ws + L.cut_cnt[o]);
    auto *incl = reinterpret_cast<CutCounts *>(ws + L.cut_incl[o]);
    const int cap = (int)std::min<int64_t>(max_items(E, S), L.rest_cap);
    cub::DoubleBuffer<unsigned long long> dk(reinterpret_cast<unsigned long long *>(ws + L.rkeys[o][0]),
                                             reinterpret_cast<unsigned long long *>(ws + L.rkeys[o][1]));
    cub::DoubleBuffer<int4> dv(reinterpret_cast<int4 *>(ws + L.ritems[o][0]), reinterpret_cast<int4 *>(ws + L.ritems[o][1]));
    const int threads = 128, blocks = (nb + threads - 1) / threads;
    cut_count_kernel<<<blocks, threads, 0, stream>>>(seg_ptr, S, cnt, nb);
    CUDA_TRY(cudaGetLastError());
    size_t temp_bytes = 0;
    CUDA_TRY(cub::DeviceScan::InclusiveScan(nullptr, temp_bytes, cnt, incl, CutSum(), nb, stream));
    if (temp_bytes > L.cut_temp_bytes) { set_error("plan_build: internal scan scratch under-reserved"); return 4; }
    temp_bytes = L.cut_temp_bytes;
    CUDA_TRY(cub::DeviceScan::InclusiveScan(ws + L.cut_temp[o], temp_bytes, cnt, incl, CutSum(), nb, stream));
    CUDA_TRY(cudaMemsetAsync(dk.Current(), 0xff, (size_t)cap * 8, stream));  // unused entries sort to the end
    auto *longs = reinterpret_cast<LongRec *>(ws + L.longs[o]);
    cut_write_kernel<<<blocks, threads, 0, stream>>>(seg_ptr, S, incl, nb, so.items, dk.Current(), dv.Current(), so.fix,
                                                     so.nonempty, longs);
    CUDA_TRY(cudaGetLastError());
    const int long_ub = (int)std::min<int64_t>(S, E / (kItemEdges + 1) + 1);  // the exact count lives in incl[nb - 1]
    cut_long_kernel<<<std::max(1, std::min(long_ub, 592)), 128, 0, stream>>>(longs, incl, nb, so.items, dk.Current(),
                                                                             dv.Current());
    CUDA_TRY(cudaGetLastError());
    temp_bytes = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, dk, dv, cap, 0, 39, stream));
    if (temp_bytes > L.cut_temp_bytes) { set_error("plan_build: internal sort scratch under-reserved"); return 4; }
    temp_bytes = L.cut_temp_bytes;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(ws + L.cut_temp[o], temp_bytes, dk, dv, cap, 0, 39, stream));  // 7 + 32 key bits
    cut_finalize_kernel<<<std::max(1, std::min(cap / 256, 64)), 256, 0, stream>>>(incl, nb, dv.Current(), so.items,
                                                                                  so.nonempty, counts, (int)E, num_edge_dev);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace

size_t plan_workspace_bytes(int64_t N, int64_t E, int64_t R) { return make_layout(N, E, R).total; }

// Host mirror of the device cutter (same functions, run block after block): the CPU test-suite checks its invariants,
// the GPU test-suite that the device kernels produce exactly this.
int cut_items_host(const int *seg_ptr, int num_seg, int *items_out, int64_t items_cap, int *num_items, int *fix_out,
                   int64_t fix_cap, int *num_fix, int *num_fix_long, int *num_slots, int *nonempty_out,
                   int64_t nonempty_cap, int *num_nonempty) {
    const int S = num_seg, nb = std::max(1, (S + kCutBlock - 1) / kCutBlock);
    std::vector<CutCounts> incl(nb);
    CutCounts run{0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < nb; b++) {
        CountVisitor v;
        const int s0 = b * kCutBlock, s1 = std::min(S, s0 + kCutBlock);
        if (s0 < s1) cut_walk_block(seg_ptr, s0, s1, v);
        run = cut_add(run, v.c);
        incl[b] = run;
    }
    const CutCounts tot = incl[nb - 1];
    const int rest = tot.whole + tot.chunks, G4 = tot.groups * kSubPerWarp;
    const int padded = (rest + kSubPerWarp - 1) / kSubPerWarp * kSubPerWarp;
    if (G4 + padded > items_cap || tot.nlong + tot.nempty > fix_cap || tot.nonempty + 1 > nonempty_cap) {
        set_error("cut_items_host: output capacity too small");
        return 1;
    }
    std::vector<int4> ritems(std::max(rest, 1));
    std::vector<unsigned long long> rkeys(std::max(rest, 1));
    std::vector<LongRec> longs(std::max(tot.nlong, 1));
    for (int b = 0; b < nb; b++) {
        WriteVisitor v;
        v.longs = longs.data();
        v.off = b ? incl[b - 1] : CutCounts{0, 0, 0, 0, 0, 0, 0};
        v.tot = tot;
        v.items = reinterpret_cast<int4 *>(items_out); v.rkeys = rkeys.data(); v.ritems = ritems.data();
        v.fix = reinterpret_cast<int4 *>(fix_out); v.nonempty_out = nonempty_out;
        const int s0 = b * kCutBlock, s1 = std::min(S, s0 + kCutBlock);
        if (s0 < s1) cut_walk_block(seg_ptr, s0, s1, v);
    }
    for (int li = 0; li < tot.nlong; li++)
        emit_long(longs[li], tot, reinterpret_cast<int4 *>(items_out), rkeys.data(), ritems.data(), 0, 1);
    std::vector<int> order(rest);
    for (int i = 0; i < rest; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return rkeys[a] < rkeys[b]; });  // keys are distinct
    int4 *items = reinterpret_cast<int4 *>(items_out);
    for (int i = 0; i < padded; i++) items[G4 + i] = i < rest ? ritems[order[i]] : make_int4(0, 0, 0, -1);
    nonempty_out[tot.nonempty] = 0;
    *num_items = G4 + padded; *num_fix = tot.nlong + tot.nempty; *num_fix_long = tot.nlong; *num_slots = tot.slots;
    *num_nonempty = tot.nonempty + 1;
    return 0;
}

int plan_build(ultra_rspmm_plan **out, const int64_t *edge_index, const int64_t *edge_type, int64_t N, int64_t E,
               int64_t R, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (N < 0 || E < 0 || R < 0 || N >= (int64_t(1) << 30) || E >= (int64_t(1) << 31) - 64 ||
        R >= (int64_t(1) << 24)) {
        set_error("plan_build: num_node / num_edge / num_relation out of int32 range");
        return 1;
    }
    const WsLayout L = make_layout(N, E, R);
    if (workspace_bytes < L.total || (workspace == nullptr && L.total > 0)) {
        set_error("plan_build: workspace too small");
        return 1;
    }
    if ((reinterpret_cast<uintptr_t>(workspace) & (kAlign - 1)) != 0) {
        set_error("plan_build: workspace must be 256-byte aligned");
        return 1;
    }
    char *ws = static_cast<char *>(workspace);
    std::unique_ptr<ultra_rspmm_plan> plan(new ultra_rspmm_plan());
    plan->N = N; plan->E = E; plan->R = R;
    CUDA_TRY(cudaGetDevice(&plan->device));
    {   // cudaGetDeviceProperties costs tens of ms per call; three attribute queries are microseconds
        int sms = 0, l2 = 0, smem = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, plan->device));
        CUDA_TRY(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, plan->device));
        CUDA_TRY(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, plan->device));
        plan->sm_count = sms;
        plan->l2_bytes = (size_t)l2;
        plan->smem_optin = (size_t)smem;
    }

    const int64_t S[3] = {N, N, R};
    for (int o = 0; o < 3; o++) {
        SegOrder &so = plan->ord[o];
        so.meta = reinterpret_cast<int2 *>(ws + L.meta[o]);
        so.eid = reinterpret_cast<int *>(ws + L.eid[o]);
        so.seg_ptr = reinterpret_cast<int *>(ws + L.seg_ptr[o]);
        so.nonempty = reinterpret_cast<int *>(ws + L.nonempty[o]);
        so.items = reinterpret_cast<int4 *>(ws + L.items[o]);
        so.fix = reinterpret_cast<int4 *>(ws + L.fix[o]);
        so.w = ws + L.w[o];
        so.num_seg = (int)S[o];
        so.meta_base = so.meta; so.eid_base = so.eid; so.seg_ptr_base = so.seg_ptr;
        so.meta_mask = reinterpret_cast<int2 *>(ws + L.meta_m[o]);
        so.eid_mask = reinterpret_cast<int *>(ws + L.eid_m[o]);
        so.seg_ptr_mask = reinterpret_cast<int *>(ws + L.seg_ptr_m[o]);
        so.counts = reinterpret_cast<const int *>(ws + L.counts) + o * kPlanCounts;
    }
    plan->ws = ws;
    plan->row_of_edge = reinterpret_cast<int *>(ws + L.row_of_edge);
    plan->row_of_edge_base = plan->row_of_edge;
    plan->row_of_edge_mask = reinterpret_cast<int *>(ws + L.row_of_edge_m);
    auto *keysA = reinterpret_cast<unsigned long long *>(ws + L.keysA);
    auto *keysB = reinterpret_cast<unsigned long long *>(ws + L.keysB);
    auto *valsA = reinterpret_cast<int *>(ws + L.valsA);
    auto *valsB = reinterpret_cast<int *>(ws + L.valsB);
    int *flags = reinterpret_cast<int *>(ws + L.flags);
    int *counts = reinterpret_cast<int *>(ws + L.counts);
    plan->w_nonunit = flags + 2;
    plan->sched = flags + 4;
    plan->counts = counts;

    const int threads = 256;
    const int blocks = (int)((E + threads - 1) / threads);
    CUDA_TRY(cudaMemsetAsync(flags, 0, 64, stream));
    if (E > 0) {
        validate_kernel<<<blocks, threads, 0, stream>>>(edge_index, edge_type, N, E, R, flags);
        CUDA_TRY(cudaGetLastError());
    }
    // Everything below is enqueued without looking at the flags: if an index is out of range the keys are garbage but
    // every kernel stays inside its buffers (indices are only used as sort keys / stored values until forward runs),
    // and the plan is rejected after the ONE synchronisation at the end.
    for (int o = 0; o < 3; o++) {
        SegOrder &so = plan->ord[o];
        const int node_bits = bits_for(std::max<int64_t>(N, 2));
        const int pri_bits = bits_for(std::max<int64_t>(S[o], 2));
        int sec_bits = node_bits;
        if (o != 2 && plan_relation_runs(N, R) && pri_bits + node_bits + bits_for(std::max<int64_t>(R, 2)) <= 64)
            sec_bits += bits_for(std::max<int64_t>(R, 2));
        if (E > 0) {
            make_keys_kernel<<<blocks, threads, 0, stream>>>(o, edge_index, edge_type, E, N, R, sec_bits, node_bits, keysA,
                                                             valsA);
            CUDA_TRY(cudaGetLastError());
            cub::DoubleBuffer<unsigned long long> dk(keysA, keysB);
            cub::DoubleBuffer<int> dv(valsA, valsB);
            size_t temp_bytes = 0;
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, dk, dv, (int)E, 0, sec_bits + pri_bits,
                                                     stream));
            if (temp_bytes > L.cub_temp_bytes) { set_error("plan_build: internal sort scratch under-reserved"); return 4; }
            temp_bytes = L.cub_temp_bytes;
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(ws + L.cub_temp, temp_bytes, dk, dv, (int)E, 0,
                                                     sec_bits + pri_bits, stream));
            build_meta_kernel<<<blocks, threads, 0, stream>>>(o, dk.Current(), dv.Current(), edge_index, edge_type,
                                                              E, (int)S[o], sec_bits, so.meta, so.eid, so.seg_ptr,
                                                              plan->row_of_edge);
            CUDA_TRY(cudaGetLastError());
        } else {
            CUDA_TRY(cudaMemsetAsync(so.seg_ptr, 0, (size_t)(S[o] + 1) * 4, stream));
        }
    }
    {   // the three cutters are chains of tiny kernels: run them side by side
        ForkJoin fj(plan.get(), stream);
        if (fj.rc) return fj.rc;
        for (int o = 0; o < 3; o++)
            if (int rc = cut_items_device(o, plan->ord[o].seg_ptr, (int)S[o], E, L, ws, plan->ord[o], counts + o * kPlanCounts,
                                          fj.lane(o)))
                return rc;
        if (int rc = fj.join()) return rc;
    }
    // the one synchronisation of a plan build: validation flags + the three orders' item / fix / slot counts
    int h[4 + 3 * kPlanCounts];
    CUDA_TRY(cudaMemcpyAsync(h, flags, 16, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(h + 4, counts, 3 * kPlanCounts * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (h[0]) {
        set_error("plan_build: edge_index / edge_type value out of range [0, num_node) / [0, num_relation)");
        return 3;
    }
    plan->input_sorted = h[1] ? 0 : 1;
    for (int o = 0; o < 3; o++) {
        SegOrder &so = plan->ord[o];
        const int *c = h + 4 + o * kPlanCounts;
        so.num_items = c[0]; so.num_fix = c[1]; so.num_fix_long = c[2]; so.num_slots = c[3];
        so.base_items = c[0]; so.base_fix = c[1]; so.base_fix_long = c[2]; so.base_slots = c[3];
        if ((int64_t)so.num_items > max_items(E, S[o])) { set_error("plan_build: item bound exceeded"); return 4; }
    }
    *out = plan.release();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Edge masks: the sub-graph that keeps a subset of the base graph's edges, re-cut WITHOUT sorting and WITHOUT a host
// synchronisation (the reference re-sorts the masked edge list inside every rspmm call of a training step,
// /root/reference/ultra/base_nbfnet.py:54-77 + /root/reference/ultra/rspmm/rspmm.py:174-178).  Removing edges keeps every
// order sorted, so per order: scan the keep flags in plan order -> compact metadata / edge ids -> gather the new segment
// pointer from the scan -> re-flag the last edge of every segment -> run the device cutter on the new pointer.  The host
// never learns the new counts: it keeps upper bounds for grid and scratch sizing, the kernels read the exact ones.
namespace {

struct KeepInOrder {  // keep flag of the i-th edge of a plan order
    const unsigned char *keep;
    const int *eid;
    __host__ __device__ int operator()(int i) const { return keep[eid[i]] ? 1 : 0; }
};

__global__ void mask_compact_kernel(const unsigned char *__restrict__ keep, const int *__restrict__ pos,
                                    const int2 *__restrict__ meta_b,
                                    const int *__restrict__ eid_b, const int *__restrict__ row_b, int2 *__restrict__ meta_m,
                                    int *__restrict__ eid_m, int *__restrict__ row_m, int E) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    const int e = eid_b[i];
    if (!keep[e]) return;
    const int j = pos[i] - 1;  // inclusive scan
    const int2 m = meta_b[i];
    meta_m[j] = make_int2(m.x, m.y & 0x7fffffff);  // last-of-segment flags are set again below
    eid_m[j] = e;  // still the caller's BASE edge id: edge weights / weight_grad keep the base graph's length and order
    if (row_m != nullptr) row_m[j] = row_b[i];
}

__global__ void mask_segptr_kernel(const int *__restrict__ seg_ptr_b, const int *__restrict__ pos, int *__restrict__ seg_ptr_m,
                                   int S) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > S) return;
    const int b = seg_ptr_b[s];
    seg_ptr_m[s] = b > 0 ? pos[b - 1] : 0;
}

__global__ void mask_lastflag_kernel(const int *__restrict__ seg_ptr_m, int2 *__restrict__ meta_m, int S, int *num_edge_out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0 && num_edge_out != nullptr) *num_edge_out = seg_ptr_m[S];
    if (s >= S) return;
    const int e = seg_ptr_m[s + 1];
    if (e > seg_ptr_m[s]) meta_m[e - 1].y |= 0x80000000;
}

}  // namespace

int plan_apply_mask(ultra_rspmm_plan *plan, const unsigned char *keep, cudaStream_t stream) {
    const int64_t N = plan->N, E = plan->E, R = plan->R;
    const WsLayout L = make_layout(N, E, R);
    char *ws = static_cast<char *>(plan->ws);
    int *counts = reinterpret_cast<int *>(ws + L.counts);
    const int64_t S[3] = {N, N, R};
    ForkJoin fj(plan, stream);  // the three orders are independent chains of small kernels: three streams
    if (fj.rc) return fj.rc;
    if (keep == nullptr) {  // back to the base graph: base arrays, exact counts, items re-cut from the base pointer
        if (!plan->masked) return fj.join();
        for (int o = 0; o < 3; o++) {
            SegOrder &so = plan->ord[o];
            so.meta = so.meta_base; so.eid = so.eid_base; so.seg_ptr = so.seg_ptr_base;
            so.num_items = so.base_items; so.num_fix = so.base_fix; so.num_fix_long = so.base_fix_long;
            so.num_slots = so.base_slots;
            if (int rc = cut_items_device(o, so.seg_ptr, (int)S[o], E, L, ws, so, counts + o * kPlanCounts, fj.lane(o)))
                return rc;
        }
        plan->row_of_edge = plan->row_of_edge_base;
        plan->masked = false;
        return fj.join();
    }
    if (E == 0) return fj.join();
    const int threads = 256, blocks = (int)((E + threads - 1) / threads);
    for (int o = 0; o < 3; o++) {
        SegOrder &so = plan->ord[o];
        cudaStream_t st = fj.lane(o);
        int *pos = reinterpret_cast<int *>(ws + L.mask_pos[o]);
        cub::TransformInputIterator<int, KeepInOrder, cub::CountingInputIterator<int>> in_o(
            cub::CountingInputIterator<int>(0), KeepInOrder{keep, so.eid_base});
        size_t temp_bytes = L.cut_temp_bytes;
        CUDA_TRY(cub::DeviceScan::InclusiveSum(ws + L.cut_temp[o], temp_bytes, in_o, pos, (int)E, st));
        mask_compact_kernel<<<blocks, threads, 0, st>>>(keep, pos, so.meta_base, so.eid_base,
                                                        o == 0 ? plan->row_of_edge_base : nullptr, so.meta_mask,
                                                        so.eid_mask, o == 0 ? plan->row_of_edge_mask : nullptr, (int)E);
        CUDA_TRY(cudaGetLastError());
        const int sblocks = (int)((S[o] + 1 + threads - 1) / threads);
        mask_segptr_kernel<<<sblocks, threads, 0, st>>>(so.seg_ptr_base, pos, so.seg_ptr_mask, (int)S[o]);
        CUDA_TRY(cudaGetLastError());
        int *num_edge_dev = counts + o * kPlanCounts + 7;  // finalize reads it back from here (same slot)
        mask_lastflag_kernel<<<sblocks, threads, 0, st>>>(so.seg_ptr_mask, so.meta_mask, (int)S[o], num_edge_dev);
        CUDA_TRY(cudaGetLastError());
        so.meta = so.meta_mask; so.eid = so.eid_mask; so.seg_ptr = so.seg_ptr_mask;
        if (int rc = cut_items_device(o, so.seg_ptr, (int)S[o], E, L, ws, so, counts + o * kPlanCounts, st, num_edge_dev))
            return rc;
        // host-side upper bounds for grids and scratch (a merged group that loses edges can break into 4 single chunks,
        // so slots are bounded by the chunk count, not by the base graph's slot count)
        so.num_items = so.base_items;                   // a HINT for grid / run sizing only: the kernel derives its units from
                                                        // the device-side count and the persistent CTAs draw tickets until done
        so.num_fix_long = so.base_fix_long;             // a sub-graph's long segments are a subset of the base graph's
        so.num_fix = so.base_fix_long + (int)S[o];      // ... but any segment may have become empty
        so.num_slots = (int)((E + kItemEdges - 1) / kItemEdges) + so.base_fix_long;
    }
    if (int rc = fj.join()) return rc;
    plan->row_of_edge = plan->row_of_edge_mask;
    plan->masked = true;
    return 0;
}

}  // namespace ultra
