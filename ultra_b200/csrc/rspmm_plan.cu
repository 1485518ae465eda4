// rspmm_plan.cu -- graph plan builder: the one-time replacement for the reference's per-call
//   key.argsort() + 3 gathers      (/root/reference/ultra/rspmm/rspmm.py:174-178)
//   sortedness assert + host sync  (/root/reference/ultra/rspmm/rspmm.py:15-17)
//   ind2ptr scatter_add + cumsum   (/root/reference/ultra/rspmm/source/rspmm.cpp:40-48)
// Three stable radix sorts (CUB, device) give the by-row / by-col / by-relation edge orders as int32 CSR;
// the edge-balanced work-item lists are cut on the host from the three segment-pointer arrays.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <vector>

#include "rspmm_kernels.cuh"
#include "rspmm_plan.h"

namespace ultra {

namespace {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }
inline int bits_for(int64_t n) {  // bits needed to represent values in [0, n)
    int b = 1;
    while ((int64_t(1) << b) < n) b++;
    return b;
}
// Edges per work item (tunable "item_edges": 64 = kItemEdges, the kernels' staging capacity and the default; 32 / 16
// cut finer items for calls so small that 64-edge items leave a second, nearly empty wave of warps - WN18RR x 4 heads
// is 344 work units on 296 CTAs).  Read when the workspace is sized AND when the plan is built; a plan built with a
// workspace sized for a coarser setting fails the size check instead of overflowing.
std::atomic<int> g_item_edges{kItemEdges};
// whole items <= 2E/L+1; chunks <= E/L + #long with #long <= E/L
inline int64_t max_items(int64_t E, int L) { return 4 * ((E + L - 1) / L) + 8; }

struct WsLayout {
    size_t meta[3], eid[3], seg_ptr[3], nonempty[3], items[3], fix[3], w[3];
    size_t row_of_edge, keysA, keysB, valsA, valsB, flags, cub_temp, cub_temp_bytes, total;
};

WsLayout make_layout(int64_t N, int64_t E, int64_t R, int item_edges) {
    WsLayout L;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(std::max<size_t>(bytes, 16)); return o; };
    const int64_t S[3] = {N, N, R};
    for (int o = 0; o < 3; o++) {
        L.meta[o] = take((size_t)E * 8);
        L.eid[o] = take((size_t)E * 4);
        L.seg_ptr[o] = take((size_t)(S[o] + 1) * 4);
        L.nonempty[o] = take((size_t)(S[o] + 1) * 4);
        L.items[o] = take((size_t)max_items(E, item_edges) * 16);
        L.fix[o] = take((size_t)std::max<int64_t>(S[o], 1) * 16);
        L.w[o] = take((size_t)E * 8);
    }
    L.row_of_edge = take((size_t)E * 4);
    L.keysA = take((size_t)E * 8);
    L.keysB = take((size_t)E * 8);
    L.valsA = take((size_t)E * 4);
    L.valsB = take((size_t)E * 4);
    L.flags = take(64);  // [0] range error, [1] unsorted, [2] w_nonunit, [4..9] work-queue tickets of the three orders
    L.cub_temp_bytes = (size_t)(1 << 20) + (size_t)E;
    L.cub_temp = take(L.cub_temp_bytes);
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
// L1/L2 temporal locality (measured: grouping by relation type instead costs +40% forward time on a uniform graph).
__global__ void make_keys_kernel(int order, const int64_t *__restrict__ ei, const int64_t *__restrict__ et,
                                 int64_t E, int sec_bits, unsigned long long *keys, int *vals) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const unsigned long long row = (unsigned long long)ei[e], col = (unsigned long long)ei[E + e];
    unsigned long long pri, sec;
    if (order == 0) { pri = row; sec = col; }
    else if (order == 1) { pri = col; sec = row; }
    else { pri = (unsigned long long)et[e]; sec = row; }
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

// Cut one segment-pointer array into work items (host).
//   * whole segments are packed greedily up to kItemEdges edges per item              {e0, e1, j0,  -1}
//   * a segment longer than kItemEdges is cut into 64-edge chunks.  Every run of 4 FULL chunks becomes a merged
//     group walked by one warp and combined with shuffles: leader {e0, e1, j|kMergedBit, slot}, followers {.., .., j|kMergedBit, -2};
//     the remaining (<= 3 full + 1 partial) chunks are single chunk items            {e0, e1, j,  slot}
//   Partial slots of one segment are consecutive and in edge order (groups, then single chunks).
// Item order: merged groups first (a multiple of 4, so a warp never mixes a group with other items), then all other
// items longest first, so the 4 items a warp walks in lock-step have (nearly) equal lengths.
// fix list: long segments first, then empty segments.
void cut_items(const std::vector<int> &ptr, std::vector<int4> &items, std::vector<int> &nonempty,
               std::vector<int4> &fix, int &num_slots, int &num_fix_long, const int kItemEdges) {  // shadows the constant
    const int S = (int)ptr.size() - 1;
    const int kGroupEdges = kItemEdges * kSubPerWarp;
    std::vector<int4> groups, rest, fix_empty;
    items.clear(); nonempty.clear(); fix.clear();
    num_slots = 0;
    int cur_begin = -1, cur_end = -1, cur_j0 = 0;
    auto close = [&]() {
        if (cur_begin >= 0) rest.push_back(make_int4(cur_begin, cur_end, cur_j0, -1));
        cur_begin = -1;
    };
    for (int s = 0; s < S; s++) {
        const int b = ptr[s], e = ptr[s + 1], deg = e - b;
        if (deg == 0) { fix_empty.push_back(make_int4(s, 0, 0, 0)); continue; }
        const int j = (int)nonempty.size();
        nonempty.push_back(s);
        if (deg > kItemEdges) {
            close();
            const int slot0 = num_slots;
            int c = b;
            for (; c + kGroupEdges <= e; c += kGroupEdges) {
                for (int k = 0; k < kSubPerWarp; k++)
                    groups.push_back(make_int4(c + k * kItemEdges, c + (k + 1) * kItemEdges, j | kMergedBit, k ? -2 : num_slots));
                num_slots++;
            }
            for (; c < e; c += kItemEdges) rest.push_back(make_int4(c, std::min(c + kItemEdges, e), j, num_slots++));
            fix.push_back(make_int4(s, slot0, num_slots - slot0, 0));
        } else {
            if (cur_begin >= 0 && (e - cur_begin) > kItemEdges) close();
            if (cur_begin < 0) { cur_begin = b; cur_j0 = j; }
            cur_end = e;
        }
    }
    close();
    std::stable_sort(rest.begin(), rest.end(),
                     [](const int4 &a, const int4 &b) { return (a.y - a.x) > (b.y - b.x); });
    items = groups;
    items.insert(items.end(), rest.begin(), rest.end());
    while (items.size() % kSubPerWarp) items.push_back(make_int4(0, 0, 0, -1));
    num_fix_long = (int)fix.size();
    fix.insert(fix.end(), fix_empty.begin(), fix_empty.end());
    nonempty.push_back(0);  // pad: the walker prefetches nonempty[j+1]
}

}  // namespace

size_t plan_workspace_bytes(int64_t N, int64_t E, int64_t R) { return make_layout(N, E, R, g_item_edges.load()).total; }

int set_item_edges(int edges) {
    if (edges != 16 && edges != 32 && edges != kItemEdges) return 1;
    g_item_edges.store(edges);
    return 0;
}

// Host-only access to the work-item cutter (pure CPU; used by the CPU test-suite to check its invariants).
int cut_items_host(const int *seg_ptr, int num_seg, int *items_out, int64_t items_cap, int *num_items, int *fix_out,
                   int64_t fix_cap, int *num_fix, int *num_fix_long, int *num_slots, int *nonempty_out,
                   int64_t nonempty_cap, int *num_nonempty) {
    std::vector<int> ptr(seg_ptr, seg_ptr + num_seg + 1);
    std::vector<int4> items, fix;
    std::vector<int> nonempty;
    int slots = 0, nlong = 0;
    cut_items(ptr, items, nonempty, fix, slots, nlong, g_item_edges.load());
    if ((int64_t)items.size() > items_cap || (int64_t)fix.size() > fix_cap || (int64_t)nonempty.size() > nonempty_cap) {
        set_error("cut_items_host: output capacity too small");
        return 1;
    }
    if (!items.empty()) std::memcpy(items_out, items.data(), items.size() * sizeof(int4));
    if (!fix.empty()) std::memcpy(fix_out, fix.data(), fix.size() * sizeof(int4));
    if (!nonempty.empty()) std::memcpy(nonempty_out, nonempty.data(), nonempty.size() * sizeof(int));
    *num_items = (int)items.size(); *num_fix = (int)fix.size(); *num_fix_long = nlong; *num_slots = slots;
    *num_nonempty = (int)nonempty.size();
    return 0;
}

int plan_build(ultra_rspmm_plan **out, const int64_t *edge_index, const int64_t *edge_type, int64_t N, int64_t E,
               int64_t R, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (N < 0 || E < 0 || R < 0 || N >= (int64_t(1) << 30) || E >= (int64_t(1) << 31) - 64 ||
        R >= (int64_t(1) << 24)) {
        set_error("plan_build: num_node / num_edge / num_relation out of int32 range");
        return 1;
    }
    const int item_edges = g_item_edges.load();  // one value for the size check and the cutter
    const WsLayout L = make_layout(N, E, R, item_edges);
    if (workspace_bytes < L.total || (workspace == nullptr && L.total > 0)) {
        set_error("plan_build: workspace too small");
        return 1;
    }
    if ((reinterpret_cast<uintptr_t>(workspace) & (kAlign - 1)) != 0) {
        set_error("plan_build: workspace must be 256-byte aligned");
        return 1;
    }
    char *ws = static_cast<char *>(workspace);
    auto *plan = new ultra_rspmm_plan();
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
    }
    plan->row_of_edge = reinterpret_cast<int *>(ws + L.row_of_edge);
    auto *keysA = reinterpret_cast<unsigned long long *>(ws + L.keysA);
    auto *keysB = reinterpret_cast<unsigned long long *>(ws + L.keysB);
    auto *valsA = reinterpret_cast<int *>(ws + L.valsA);
    auto *valsB = reinterpret_cast<int *>(ws + L.valsB);
    int *flags = reinterpret_cast<int *>(ws + L.flags);
    plan->w_nonunit = flags + 2;
    plan->sched = flags + 4;

    const int threads = 256;
    const int blocks = (int)((E + threads - 1) / threads);
    int h_flags[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemsetAsync(flags, 0, 64, stream));
    if (E > 0) {
        validate_kernel<<<blocks, threads, 0, stream>>>(edge_index, edge_type, N, E, R, flags);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaMemcpyAsync(h_flags, flags, 16, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (h_flags[0]) {
        delete plan;
        set_error("plan_build: edge_index / edge_type value out of range [0, num_node) / [0, num_relation)");
        return 3;
    }
    plan->input_sorted = h_flags[1] ? 0 : 1;

    std::vector<int> h_ptr;
    std::vector<int4> items, fix;
    std::vector<int> nonempty;
    for (int o = 0; o < 3; o++) {
        SegOrder &so = plan->ord[o];
        const int sec_bits = bits_for(std::max<int64_t>(N, 2));
        const int pri_bits = bits_for(std::max<int64_t>(S[o], 2));
        if (E > 0) {
            make_keys_kernel<<<blocks, threads, 0, stream>>>(o, edge_index, edge_type, E, sec_bits, keysA, valsA);
            CUDA_TRY(cudaGetLastError());
            cub::DoubleBuffer<unsigned long long> dk(keysA, keysB);
            cub::DoubleBuffer<int> dv(valsA, valsB);
            size_t temp_bytes = 0;
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, dk, dv, (int)E, 0, sec_bits + pri_bits,
                                                     stream));
            if (temp_bytes > L.cub_temp_bytes) {
                delete plan;
                set_error("plan_build: internal sort scratch under-reserved");
                return 4;
            }
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
        h_ptr.resize((size_t)S[o] + 1);
        CUDA_TRY(cudaMemcpyAsync(h_ptr.data(), so.seg_ptr, (size_t)(S[o] + 1) * 4, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        cut_items(h_ptr, items, nonempty, fix, so.num_slots, so.num_fix_long, item_edges);
        if ((int64_t)items.size() > max_items(E, item_edges)) {
            delete plan;
            set_error("plan_build: item bound exceeded");
            return 4;
        }
        so.num_items = (int)items.size();
        so.num_fix = (int)fix.size();
        if (!items.empty())
            CUDA_TRY(cudaMemcpyAsync(so.items, items.data(), items.size() * 16, cudaMemcpyHostToDevice, stream));
        if (!fix.empty())
            CUDA_TRY(cudaMemcpyAsync(so.fix, fix.data(), fix.size() * 16, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(so.nonempty, nonempty.data(), nonempty.size() * 4, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));  // host vectors are reused by the next order
    }
    *out = plan;
    return 0;
}

}  // namespace ultra
