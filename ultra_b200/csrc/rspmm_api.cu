// rspmm_api.cu -- extern "C" surface declared in include/ultra_rspmm.h: argument checks, kernel selection and
// launch geometry.  Replaces the host launchers of the reference (rspmm.cu:216-267 forward, :269-351 backward).
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>

#include "../../include/ultra_rspmm.h"
#include "rspmm_kernels.cuh"
#include "rspmm_plan.h"

namespace ultra {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }

namespace {

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t err__ = (expr);                                                            \
        if (err__ != cudaSuccess) {                                                            \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(err__));                  \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)

constexpr size_t kMetaSmem = (size_t)kCtaWarps * kMetaStageBytes;  // per-warp metadata staging per CTA (33 KB at 16 warps)
// tiles in smem while kCtasPerSm CTAs (tiles + metadata + 1 KB reserved each + static smem) still fit in 227 KB: 78 KB at 512 x 2
constexpr size_t kSlabSmemLimit = ((size_t)227 * 1024 / kCtasPerSm - kMetaSmem - 2560) / 1024 * 1024;
static_assert(kCtasPerSm != 2 || kCtaThreads != 512 || kSlabSmemLimit == 78 * 1024, "default geometry: 78 KB of tiles");

// process-wide launch tunables (ultra_rspmm_set_tunable); the SM reservation is per plan (ultra_rspmm_plan_set_sm_limit)
std::atomic<int> g_l2_percent{25};        // share of L2 the concurrently gathered tiles may occupy
std::atomic<int> g_host_copy_split{1};   // forward_host: row blocks (streams) per slab copy
std::atomic<int> g_bwd_fused_relgrad{0};
std::atomic<int> g_plan_relation_runs{-1};  // read at plan build: sort a segment's edges by relation type first
                                            // (-1 = only graphs whose tiles are shared-memory resident, 0 never, 1 always)
std::atomic<int> g_work_queue{1};       // experiment knob: 1 = ticket work queue (default), 0 = static round-robin
std::atomic<int> g_smem_tiles{2};       // experiment knob: highest SM mode allowed (2 all, 1 relation tile only, 0 none) // 1 = one backward pass, relation_grad by vector reductions (order-dependent)

template <typename T> bool vec_ok(int64_t D, std::initializer_list<const void *> ptrs) {
    if (D % VecN<T>::value != 0) return false;
    for (const void *p : ptrs)
        if (p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15) != 0) return false;
    return true;
}
}  // namespace
bool plan_relation_runs(int64_t N, int64_t R) {
    const int v = g_plan_relation_runs.load();
    return v < 0 ? (size_t)(N + R) * kTileBytes <= kSlabSmemLimit : v != 0;  // the sm_mode == 2 condition of launch_pass
}
namespace {

// ---- one segment pass (tiled kernel + fix-up, or the generic kernel) -------------------------------------
// heads == 0: flat (rows, D) matrices.  heads > 0: head-major (heads, rows, D / heads) matrices.
template <typename T, int MODE, int SUM, int MUL>
int launch_pass(const ultra_rspmm_plan *plan, const T *relation, const T *input, const T *output,
                const T *outgrad, T *dst, T *scratch, int64_t D, cudaStream_t stream, int64_t heads = 0,
                const T *boundary = nullptr, T *rgrad = nullptr, bool rel_broadcast = false) {
    using Nd = Needs<MODE, SUM, MUL>;
    constexpr int ORD = order_of(MODE);  // BWD_BOTH walks the by-column order
    const SegOrder &so = plan->ord[ORD];
    const T *w = plan->unit_weight ? nullptr : static_cast<const T *>(so.w);
    if (so.num_seg == 0 || D == 0) return 0;

    constexpr int V = VecN<T>::value;
    constexpr int TILE = kLanesPerEdge * V;
    Layout lay;
    if (heads > 0) {
        const int64_t d = D / heads;
        if (d * heads != D || (d * (int64_t)sizeof(T)) % kTileBytes != 0 ||
            !vec_ok<T>(D, {relation, input, output, outgrad, dst, scratch, boundary})) {
            set_error("head-major layout needs head_dim * sizeof(T) to be a multiple of 128 bytes and 16-byte "
                      "aligned pointers");
            return 1;
        }
        lay.tpg = (int)(d * sizeof(T) / kTileBytes);
        lay.group_cols = (int)d;
        lay.row_bytes = (unsigned)(d * sizeof(T));
        lay.grp_node_bytes = (long long)plan->N * lay.row_bytes;
        lay.grp_rgrad_bytes = (long long)plan->R * lay.row_bytes;
        lay.grp_rel_bytes = rel_broadcast ? 0 : lay.grp_rgrad_bytes;
    } else {
        lay.tpg = (int)((D + TILE - 1) / TILE);
        lay.group_cols = (int)D;
        lay.row_bytes = (unsigned)(D * sizeof(T));
        lay.grp_node_bytes = 0;
        lay.grp_rel_bytes = 0;
        lay.grp_rgrad_bytes = 0;
    }
    const long long dst_grp_bytes = MODE == BWD_REL ? lay.grp_rgrad_bytes : lay.grp_node_bytes;

    if (heads == 0 && !vec_ok<T>(D, {relation, input, output, outgrad, dst, scratch, boundary, rgrad})) {
        if constexpr (MODE == BWD_BOTH) { set_error("fused backward needs the vector path"); return 1; } else {
        if (boundary != nullptr) { set_error("boundary fusion needs the vector path (D % 4 == 0, aligned)"); return 1; }
        if constexpr (SUM == S_PNA) { set_error("forward_pna needs the vector path (D % 4 == 0, aligned)"); return 1; } else {
        const long long total = (long long)so.num_seg * D;
        const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)plan->sm_count * 16);
        seg_reduce_generic_kernel<T, MODE, SUM, MUL><<<blocks, 256, 0, stream>>>(
            so.meta, w, so.seg_ptr, so.num_seg, relation, input, output, outgrad, dst, (int)D);
        CUDA_TRY(cudaGetLastError());
        return 0;
        }
        }
    }

    SegParams<T> p;
    p.meta = so.meta; p.w = w; p.w_nonunit = plan->w_nonunit; p.sched = g_work_queue.load() ? plan->sched + 2 * ORD : nullptr; p.nonempty = so.nonempty; p.items = so.items;
    p.counts = so.counts;
    const int ub_warp_items = so.num_items / kSubPerWarp;  // exact for the base graph, an upper bound under an edge mask
    p.relation = relation; p.input = input; p.output = output; p.outgrad = outgrad;
    p.dst = dst; p.rgrad = rgrad; p.partial = scratch; p.boundary = boundary; p.lay = lay;
    p.dst_plane = (long long)plan->ord[ORD].num_seg * D;          // S_PNA: (4, N, D) outputs
    p.partial_plane = (long long)plan->ord[ORD].num_slots * D;    // S_PNA: (4, slots, D) partials
    p.D = (int)D; p.R = (int)plan->R; p.N = (int)plan->N;
    p.num_tiles = (int)((D + TILE - 1) / TILE);

    if (ub_warp_items > 0) {
        const int sm_limit = plan->sm_limit;
        const int ctas = ((sm_limit > 0 && sm_limit < plan->sm_count) ? sm_limit : plan->sm_count) * kCtasPerSm;
        // items per warp per unit: 2 when that still leaves every CTA several units
        int ipw = 2;
        auto units_per_tile = [&](int k) { return (ub_warp_items + kCtaWarps * k - 1) / (kCtaWarps * k); };
        if ((long long)units_per_tile(ipw) * p.num_tiles < 4LL * ctas) ipw = 1;
        p.items_per_warp = ipw;
        const int ub_units_per_tile = units_per_tile(ipw);  // the kernel recomputes both from the device-side count
        const long long ub_total_units = (long long)ub_units_per_tile * p.num_tiles;
        // tile-major walk: each CTA takes `run` consecutive units, so ~ctas*run/units_per_tile tiles are in
        // flight; keep their gathered footprint (segments' source rows x 128 B) within a fraction of L2.
        const size_t l2_budget = (plan->l2_bytes ? plan->l2_bytes : (size_t)96 << 20) / 100 * (size_t)g_l2_percent.load();
        const size_t rows_gathered = (size_t)std::max<int64_t>(plan->N, 1) * Nd::gathers;  // matrices gathered from
        const double tiles_in_flight = std::max(1.0, (double)l2_budget / ((double)rows_gathered * kTileBytes));
        long long run = (long long)(tiles_in_flight * ub_units_per_tile / ctas);
        run = std::max<long long>(1, std::min<long long>(run, ub_units_per_tile));
        // ... but never so long that CTAs go idle: keep >= 4 blocks of `run` units per CTA when the work allows it
        // (small graphs have huge tiles_in_flight; without this cap each CTA would own a whole tile and a
        // 32-tile problem would occupy 32 of 296 CTAs)
        run = std::max<long long>(1, std::min<long long>(run, ub_total_units / (4LL * ctas)));
        p.run = (int)run;
        const long long nblk = (ub_total_units + p.run - 1) / p.run;
        const int grid = (int)std::min<long long>(nblk, ctas);

        // shared-memory residency: relation tile (R x 128 B) and, for small graphs, the gathered matrix' tile too
        const size_t rel_bytes = Nd::rel ? (size_t)plan->R * kTileBytes : 0;
        const size_t g1_bytes = (size_t)plan->N * kTileBytes;
        int sm_mode = 0;
        if (rel_bytes + g1_bytes <= kSlabSmemLimit) sm_mode = 2;
        else if (Nd::rel && rel_bytes <= kSlabSmemLimit) sm_mode = 1;
        sm_mode = std::min(sm_mode, g_smem_tiles.load());
        auto launch = [&](auto kern, size_t smem) -> int {
            if (smem > 48 * 1024) {  // opt-in above 48 KB; cudaFuncSetAttribute is cheap and idempotent
                CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)(kSlabSmemLimit + kMetaSmem)));
            }
            kern<<<grid, kCtaThreads, smem, stream>>>(p);
            return 0;
        };
        int lrc = 0;
        if (sm_mode == 2) lrc = launch(seg_reduce_kernel<T, MODE, SUM, MUL, 2>, rel_bytes + g1_bytes + kMetaSmem);
        else if (sm_mode == 1) lrc = launch(seg_reduce_kernel<T, MODE, SUM, MUL, 1>, rel_bytes + kMetaSmem);
        else lrc = launch(seg_reduce_kernel<T, MODE, SUM, MUL, 0>, kMetaSmem);
        if (lrc) return lrc;
        CUDA_TRY(cudaGetLastError());
    }
    const int ncc = (int)((D + 32 * V - 1) / (32 * V));
    const int num_empty = so.num_fix - so.num_fix_long;  // upper bounds under an edge mask
    // one fold + one identity fill per accumulator plane (1, or 4 for S_PNA: sum, sum of squares, max, min)
    auto fix_plane = [&](auto red_tag, int a) -> int {
        constexpr int RED = decltype(red_tag)::value;
        T *dst_a = dst + (size_t)a * p.dst_plane;
        const T *part_a = scratch + (size_t)a * p.partial_plane;
        if (so.num_fix_long > 0) {
            fold_partials_kernel<T, RED><<<so.num_fix_long * ncc, 256, 0, stream>>>(
                so.fix, so.counts, part_a, dst_a, boundary, (int)D, lay, dst_grp_bytes, ncc);
            CUDA_TRY(cudaGetLastError());
        }
        if (num_empty > 0) {
            const long long warps = (long long)num_empty * ncc;
            const int blocks = (int)std::min<long long>((warps * 32 + 255) / 256, (long long)plan->sm_count * 32);
            fill_identity_kernel<T, RED><<<blocks, 256, 0, stream>>>(so.fix, so.counts, dst_a, boundary, (int)D, lay,
                                                                     dst_grp_bytes);
            CUDA_TRY(cudaGetLastError());
        }
        return 0;
    };
    if constexpr (Nd::red == S_PNA) {
        if (int rc = fix_plane(std::integral_constant<int, S_ADD>{}, 0)) return rc;
        if (int rc = fix_plane(std::integral_constant<int, S_ADD>{}, 1)) return rc;
        if (int rc = fix_plane(std::integral_constant<int, S_MAX>{}, 2)) return rc;
        if (int rc = fix_plane(std::integral_constant<int, S_MIN>{}, 3)) return rc;
    } else {
        if (int rc = fix_plane(std::integral_constant<int, Nd::red>{}, 0)) return rc;
    }
    return 0;
}

template <typename T, int SUM, int MUL>
int forward_t(const ultra_rspmm_plan *plan, const void *relation, const void *input, void *output, int64_t D,
              void *scratch, cudaStream_t stream, int64_t heads = 0, const void *boundary = nullptr,
              bool rel_broadcast = false) {
    return launch_pass<T, FWD, SUM, MUL>(plan, (const T *)relation, (const T *)input, nullptr, nullptr, (T *)output,
                                         (T *)scratch, D, stream, heads, (const T *)boundary, nullptr, rel_broadcast);
}

template <typename T, int SUM, int MUL>
int backward_t(const ultra_rspmm_plan *plan, const void *relation, const void *input, const void *output,
               const void *outgrad, void *wgrad, void *rgrad, void *igrad, int64_t D, void *scratch,
               cudaStream_t stream, int64_t heads = 0, bool rel_broadcast = false) {
    int rc;
    const bool fused = g_bwd_fused_relgrad.load() != 0 && plan->R > 0 &&
                       (heads > 0 || vec_ok<T>(D, {relation, input, output, outgrad, igrad, rgrad, scratch}));
    if (fused) {
        // one pass over the by-column order: input_grad by gather-reduce, relation_grad by REDG.F32x4 into zeros
        CUDA_TRY(cudaMemsetAsync(rgrad, 0, (size_t)plan->R * (size_t)D * sizeof(T), stream));
        rc = launch_pass<T, BWD_BOTH, SUM, MUL>(plan, (const T *)relation, (const T *)input, (const T *)output,
                                                (const T *)outgrad, (T *)igrad, (T *)scratch, D, stream, heads,
                                                nullptr, (T *)rgrad, rel_broadcast);
        if (rc) return rc;
    } else {
        rc = launch_pass<T, BWD_IN, SUM, MUL>(plan, (const T *)relation, (const T *)input, (const T *)output,
                                              (const T *)outgrad, (T *)igrad, (T *)scratch, D, stream, heads, nullptr,
                                              nullptr, rel_broadcast);
        if (rc) return rc;
        rc = launch_pass<T, BWD_REL, SUM, MUL>(plan, (const T *)relation, (const T *)input, (const T *)output,
                                               (const T *)outgrad, (T *)rgrad, (T *)scratch, D, stream, heads, nullptr,
                                               nullptr, rel_broadcast);
        if (rc) return rc;
    }
    if (wgrad != nullptr && heads > 0) { set_error("weight_grad is not available in the head-major layout"); return 1; }
    if (wgrad != nullptr && plan->E > 0) {
        const SegOrder &so = plan->ord[FWD];
        const T *w = plan->unit_weight ? nullptr : static_cast<const T *>(so.w);
        if (plan->masked)  // removed edges receive no gradient: their (base-order) entries are zero
            CUDA_TRY(cudaMemsetAsync(wgrad, 0, (size_t)plan->E * sizeof(T), stream));
        const long long threads = (long long)plan->E * 32;
        wgrad_kernel<T, SUM, MUL><<<(int)((threads + 255) / 256), 256, 0, stream>>>(
            so.meta, so.eid, plan->row_of_edge, w, (const T *)relation, (const T *)input, (const T *)output,
            (const T *)outgrad, (T *)wgrad, so.counts, (int)D);
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

#define DISPATCH_OPS(FN, T, ...)                                                          \
    switch (sum_op * 2 + mul_op) {                                                        \
        case 0: return FN<T, S_ADD, M_MUL>(__VA_ARGS__);                                  \
        case 1: return FN<T, S_ADD, M_ADD>(__VA_ARGS__);                                  \
        case 2: return FN<T, S_MIN, M_MUL>(__VA_ARGS__);                                  \
        case 3: return FN<T, S_MIN, M_ADD>(__VA_ARGS__);                                  \
        case 4: return FN<T, S_MAX, M_MUL>(__VA_ARGS__);                                  \
        case 5: return FN<T, S_MAX, M_ADD>(__VA_ARGS__);                                  \
    }

int check_common(const ultra_rspmm_plan *plan, int sum_op, int mul_op, int dtype, int64_t dim) {
    if (plan == nullptr) { set_error("null plan"); return 1; }
    if (sum_op < 0 || sum_op > 2 || mul_op < 0 || mul_op > 1) { set_error("unknown sum_op / mul_op"); return 1; }
    if (dtype != ULTRA_RSPMM_F32 && dtype != ULTRA_RSPMM_F64) { set_error("unknown dtype"); return 1; }
    if (dim < 0 || dim >= (int64_t(1) << 31) || dim * (dtype == ULTRA_RSPMM_F32 ? 4 : 8) >= (int64_t(1) << 32)) {
        set_error("dim out of range (a feature row must be shorter than 4 GiB)");
        return 1;
    }
    if (!plan->unit_weight && plan->w_dtype != dtype) {
        set_error("edge weights were bound with a different dtype than relation/input (checkAllSameType)");
        return 1;
    }
    return 0;
}

size_t scratch_bytes(const ultra_rspmm_plan *plan, int64_t dim, int dtype) {
    int slots = 0;
    for (int o = 0; o < 3; o++) slots = std::max(slots, plan->ord[o].num_slots);
    return (size_t)slots * (size_t)dim * (dtype == ULTRA_RSPMM_F32 ? 4 : 8);
}

}  // namespace
}  // namespace ultra

// Streams / events of the host-buffer pipeline (ultra_rspmm_forward_host), owned by the plan and created lazily.
struct ultra::HostLanes {
    static constexpr int kMax = 4;
    cudaStream_t in[kMax] = {}, out[kMax] = {}, k = nullptr;
    cudaEvent_t start = nullptr, done = nullptr, ein[2][kMax] = {}, ek[2] = {}, eout[2][kMax] = {};
    int ready = 0;  // lanes [0, ready) exist
    int ensure(int split) {
        if (k == nullptr) {
            CUDA_TRY(cudaStreamCreateWithFlags(&k, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
            for (int b = 0; b < 2; b++) CUDA_TRY(cudaEventCreateWithFlags(&ek[b], cudaEventDisableTiming));
        }
        for (; ready < split; ready++) {
            CUDA_TRY(cudaStreamCreateWithFlags(&in[ready], cudaStreamNonBlocking));
            CUDA_TRY(cudaStreamCreateWithFlags(&out[ready], cudaStreamNonBlocking));
            for (int b = 0; b < 2; b++) {
                CUDA_TRY(cudaEventCreateWithFlags(&ein[b][ready], cudaEventDisableTiming));
                CUDA_TRY(cudaEventCreateWithFlags(&eout[b][ready], cudaEventDisableTiming));
            }
        }
        return 0;
    }
    ~HostLanes() {  // pending work completes first (stream destruction is deferred by the runtime)
        for (int i = 0; i < kMax; i++) {
            if (in[i]) cudaStreamDestroy(in[i]);
            if (out[i]) cudaStreamDestroy(out[i]);
            for (int b = 0; b < 2; b++) {
                if (ein[b][i]) cudaEventDestroy(ein[b][i]);
                if (eout[b][i]) cudaEventDestroy(eout[b][i]);
            }
        }
        if (k) cudaStreamDestroy(k);
        if (start) cudaEventDestroy(start);
        if (done) cudaEventDestroy(done);
        for (int b = 0; b < 2; b++) if (ek[b]) cudaEventDestroy(ek[b]);
    }
};

using namespace ultra;

ultra_rspmm_plan::~ultra_rspmm_plan() {
    delete host_lanes;
    for (int i = 0; i < 2; i++) {
        if (side[i]) cudaStreamDestroy(side[i]);
        if (side_ev[i]) cudaEventDestroy(side_ev[i]);
    }
    if (fork_ev) cudaEventDestroy(fork_ev);
}

extern "C" {

const char *ultra_rspmm_version(void) { return "ultra_rspmm sm_100a 0.1.0"; }

int ultra_rspmm_set_tunable(const char *name, int value) {
    if (name == nullptr) { set_error("set_tunable: null name"); return 1; }
    const std::string n(name);
    if (n == "host_copy_split") { g_host_copy_split.store(value); return 0; }
    if (n == "bwd_fused_relgrad") { g_bwd_fused_relgrad.store(value != 0); return 0; }
    if (n == "work_queue") { g_work_queue.store(value != 0); return 0; }
    if (n == "plan_relation_runs") { g_plan_relation_runs.store(value < 0 ? -1 : (value != 0)); return 0; }
    if (n == "smem_tiles") { g_smem_tiles.store(value < 0 ? 0 : (value > 2 ? 2 : value)); return 0; }
    if (n == "l2_percent") { g_l2_percent.store(value < 1 ? 1 : (value > 100 ? 100 : value)); return 0; }
    set_error("set_tunable: unknown tunable '" + n + "' (l2_percent, host_copy_split, bwd_fused_relgrad, smem_tiles, work_queue, plan_relation_runs)");
    return 1;
}

int ultra_rspmm_plan_set_sm_limit(ultra_rspmm_plan *plan, int sm_limit) {
    if (plan == nullptr) { set_error("null plan"); return 1; }
    plan->sm_limit = sm_limit < 0 ? 0 : sm_limit;
    return 0;
}

int ultra_rspmm_debug_cut_items(const int *seg_ptr, int num_seg, int *items_out, int64_t items_cap, int *num_items,
                                int *fix_out, int64_t fix_cap, int *num_fix, int *num_fix_long, int *num_slots,
                                int *nonempty_out, int64_t nonempty_cap, int *num_nonempty) {
    if (!seg_ptr || num_seg < 0 || !num_items || !num_fix || !num_fix_long || !num_slots || !num_nonempty) {
        set_error("debug_cut_items: bad arguments");
        return 1;
    }
    return cut_items_host(seg_ptr, num_seg, items_out, items_cap, num_items, fix_out, fix_cap, num_fix, num_fix_long,
                          num_slots, nonempty_out, nonempty_cap, num_nonempty);
}

int ultra_rspmm_plan_debug_read(const ultra_rspmm_plan *plan, int order, int what, void *host_out, int64_t cap_bytes) {
    if (plan == nullptr || order < 0 || order > 2 || host_out == nullptr) { set_error("plan_debug_read: bad arguments"); return 1; }
    const SegOrder &so = plan->ord[order];
    const void *src = nullptr;
    size_t bytes = 0;
    switch (what) {
        case 0: src = so.items; bytes = (size_t)so.num_items * 16; break;
        case 1: src = so.fix; bytes = (size_t)so.num_fix * 16; break;
        case 2: src = so.nonempty; bytes = (size_t)(so.num_seg + 1) * 4; break;
        case 3: src = so.seg_ptr; bytes = (size_t)(so.num_seg + 1) * 4; break;
        case 4: src = so.counts; bytes = 8 * 4; break;
        case 5: src = so.eid; bytes = (size_t)plan->E * 4; break;
        default: set_error("plan_debug_read: what must be 0 (items), 1 (fix), 2 (nonempty), 3 (seg_ptr), 4 (counts), 5 (edge ids)"); return 1;
    }
    if (plan->masked && (what == 0 || what == 1)) bytes = (size_t)cap_bytes;  // exact sizes live on the device (what = 4)
    if ((int64_t)bytes > cap_bytes) { set_error("plan_debug_read: host buffer too small"); return 1; }
    if (bytes > 0) CUDA_TRY(cudaMemcpy(host_out, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

const char *ultra_rspmm_last_error(void) { return g_last_error.c_str(); }

size_t ultra_rspmm_plan_workspace_bytes(int64_t num_node, int64_t num_edge, int64_t num_relation) {
    return plan_workspace_bytes(num_node, num_edge, num_relation);
}

int ultra_rspmm_plan_build(ultra_rspmm_plan **plan_out, const int64_t *edge_index, const int64_t *edge_type,
                           int64_t num_node, int64_t num_edge, int64_t num_relation, void *workspace,
                           size_t workspace_bytes, void *stream) {
    if (plan_out == nullptr) { set_error("plan_out is null"); return 1; }
    *plan_out = nullptr;
    return plan_build(plan_out, edge_index, edge_type, num_node, num_edge, num_relation, workspace,
                      workspace_bytes, static_cast<cudaStream_t>(stream));
}

int ultra_rspmm_plan_apply_mask(ultra_rspmm_plan *plan, const unsigned char *keep_mask, void *stream) {
    if (plan == nullptr) { set_error("null plan"); return 1; }
    return plan_apply_mask(plan, keep_mask, static_cast<cudaStream_t>(stream));
}

void ultra_rspmm_plan_destroy(ultra_rspmm_plan *plan) {
    delete plan;
}

int ultra_rspmm_plan_input_sorted(const ultra_rspmm_plan *plan) { return plan ? plan->input_sorted : 0; }

int64_t ultra_rspmm_plan_info(const ultra_rspmm_plan *plan, int what) {
    if (plan == nullptr) return -1;
    switch (what) {
        case 0: return plan->N;
        case 1: return plan->E;
        case 2: return plan->R;
        case 3: return plan->ord[0].num_items;
        case 4: return plan->ord[0].num_slots;
        case 5: return plan->ord[1].num_slots;
        case 6: return plan->ord[2].num_slots;
        case 7: return plan->sm_count;
        case 8: return plan->ord[1].num_items;
        case 9: return plan->ord[2].num_items;
    }
    if (what >= 10 && what < 19) {
        const SegOrder &so = plan->ord[(what - 10) / 3];
        const int k = (what - 10) % 3;
        return k == 0 ? so.num_fix : (k == 1 ? so.num_fix_long : so.num_seg);
    }
    return -1;
}

size_t ultra_rspmm_scratch_bytes(const ultra_rspmm_plan *plan, int64_t dim, int dtype) {
    return plan ? scratch_bytes(plan, dim, dtype) : 0;
}

int ultra_rspmm_set_edge_weight(ultra_rspmm_plan *plan, const void *edge_weight, int dtype, void *stream_) {
    if (plan == nullptr) { set_error("null plan"); return 1; }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (edge_weight == nullptr) { plan->unit_weight = true; plan->w_dtype = -1; return 0; }
    if (dtype != ULTRA_RSPMM_F32 && dtype != ULTRA_RSPMM_F64) { set_error("unknown dtype"); return 1; }
    plan->unit_weight = false;
    plan->w_dtype = dtype;
    if (plan->E == 0) return 0;
    const int E = (int)plan->E, blocks = (E + 255) / 256;
    CUDA_TRY(cudaMemsetAsync(plan->w_nonunit, 0, sizeof(int), stream));
    if (dtype == ULTRA_RSPMM_F32)
        permute_weight_kernel<float><<<blocks, 256, 0, stream>>>(
            (const float *)edge_weight, plan->ord[0].eid, plan->ord[1].eid, plan->ord[2].eid,
            (float *)plan->ord[0].w, (float *)plan->ord[1].w, (float *)plan->ord[2].w, plan->ord[0].counts,
            plan->w_nonunit);
    else
        permute_weight_kernel<double><<<blocks, 256, 0, stream>>>(
            (const double *)edge_weight, plan->ord[0].eid, plan->ord[1].eid, plan->ord[2].eid,
            (double *)plan->ord[0].w, (double *)plan->ord[1].w, (double *)plan->ord[2].w, plan->ord[0].counts,
            plan->w_nonunit);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int ultra_rspmm_forward(const ultra_rspmm_plan *plan, int sum_op, int mul_op, int dtype, const void *relation,
                        const void *input, void *output, int64_t dim, void *scratch, size_t scratch_bytes_,
                        void *stream_) {
    if (int rc = check_common(plan, sum_op, mul_op, dtype, dim)) return rc;
    if (scratch_bytes_ < scratch_bytes(plan, dim, dtype)) { set_error("forward: scratch too small"); return 1; }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (dtype == ULTRA_RSPMM_F32) {
        DISPATCH_OPS(forward_t, float, plan, relation, input, output, dim, scratch, stream)
    } else {
        DISPATCH_OPS(forward_t, double, plan, relation, input, output, dim, scratch, stream)
    }
    return 1;
}

int ultra_rspmm_backward(const ultra_rspmm_plan *plan, int sum_op, int mul_op, int dtype, const void *relation,
                         const void *input, const void *output, const void *output_grad, void *weight_grad,
                         void *relation_grad, void *input_grad, int64_t dim, void *scratch, size_t scratch_bytes_,
                         void *stream_) {
    if (int rc = check_common(plan, sum_op, mul_op, dtype, dim)) return rc;
    if (scratch_bytes_ < scratch_bytes(plan, dim, dtype)) { set_error("backward: scratch too small"); return 1; }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (dtype == ULTRA_RSPMM_F32) {
        DISPATCH_OPS(backward_t, float, plan, relation, input, output, output_grad, weight_grad, relation_grad,
                     input_grad, dim, scratch, stream)
    } else {
        DISPATCH_OPS(backward_t, double, plan, relation, input, output, output_grad, weight_grad, relation_grad,
                     input_grad, dim, scratch, stream)
    }
    return 1;
}

int ultra_rspmm_forward_pna(const ultra_rspmm_plan *plan, int mul_op, int dtype, const void *relation,
                            const void *input, void *output4, int64_t dim, void *scratch, size_t scratch_bytes_,
                            void *stream_) {
    if (int rc = check_common(plan, 0, mul_op, dtype, dim)) return rc;
    if (scratch_bytes_ < 4 * scratch_bytes(plan, dim, dtype)) {
        set_error("forward_pna: scratch too small (needs 4 x ultra_rspmm_scratch_bytes)");
        return 1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (dtype == ULTRA_RSPMM_F32) {
        if (mul_op == 0) return forward_t<float, S_PNA, M_MUL>(plan, relation, input, output4, dim, scratch, stream);
        return forward_t<float, S_PNA, M_ADD>(plan, relation, input, output4, dim, scratch, stream);
    }
    if (mul_op == 0) return forward_t<double, S_PNA, M_MUL>(plan, relation, input, output4, dim, scratch, stream);
    return forward_t<double, S_PNA, M_ADD>(plan, relation, input, output4, dim, scratch, stream);
}

int ultra_rspmm_forward_heads(const ultra_rspmm_plan *plan, int sum_op, int mul_op, int dtype, const void *relation,
                              const void *input, const void *boundary, void *output, int64_t num_heads,
                              int64_t head_dim, int relation_broadcast, void *scratch, size_t scratch_bytes_,
                              void *stream_) {
    if (num_heads <= 0 || head_dim <= 0) { set_error("forward_heads: num_heads / head_dim must be positive"); return 1; }
    const int64_t dim = num_heads * head_dim;
    if (int rc = check_common(plan, sum_op, mul_op, dtype, dim)) return rc;
    if (scratch_bytes_ < scratch_bytes(plan, dim, dtype)) { set_error("forward_heads: scratch too small"); return 1; }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (dtype == ULTRA_RSPMM_F32) {
        DISPATCH_OPS(forward_t, float, plan, relation, input, output, dim, scratch, stream, num_heads, boundary,
                     relation_broadcast != 0)
    } else {
        DISPATCH_OPS(forward_t, double, plan, relation, input, output, dim, scratch, stream, num_heads, boundary,
                     relation_broadcast != 0)
    }
    return 1;
}

int ultra_rspmm_backward_heads(const ultra_rspmm_plan *plan, int sum_op, int mul_op, int dtype, const void *relation,
                               const void *input, const void *output, const void *output_grad, void *relation_grad,
                               void *input_grad, int64_t num_heads, int64_t head_dim, int relation_broadcast,
                               void *scratch, size_t scratch_bytes_, void *stream_) {
    if (num_heads <= 0 || head_dim <= 0) { set_error("backward_heads: num_heads / head_dim must be positive"); return 1; }
    const int64_t dim = num_heads * head_dim;
    if (int rc = check_common(plan, sum_op, mul_op, dtype, dim)) return rc;
    if (scratch_bytes_ < scratch_bytes(plan, dim, dtype)) { set_error("backward_heads: scratch too small"); return 1; }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (dtype == ULTRA_RSPMM_F32) {
        DISPATCH_OPS(backward_t, float, plan, relation, input, output, output_grad, nullptr, relation_grad, input_grad,
                     dim, scratch, stream, num_heads, relation_broadcast != 0)
    } else {
        DISPATCH_OPS(backward_t, double, plan, relation, input, output, output_grad, nullptr, relation_grad,
                     input_grad, dim, scratch, stream, num_heads, relation_broadcast != 0)
    }
    return 1;
}

// ---- end-to-end from host buffers: column slabs pipelined H2D -> kernel -> D2H ------------------------------
size_t ultra_rspmm_host_staging_bytes(const ultra_rspmm_plan *plan, int64_t dim, int64_t slab_dim, int dtype) {
    if (plan == nullptr || slab_dim <= 0) return 0;
    const size_t es = dtype == ULTRA_RSPMM_F32 ? 4 : 8;
    slab_dim = std::min(slab_dim, std::max<int64_t>(dim, 1));
    // double-buffered {relation slab, input slab, output slab, scratch}
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t per = up((size_t)plan->R * slab_dim * es) + 2 * up((size_t)plan->N * slab_dim * es) +
                       up(scratch_bytes(plan, slab_dim, dtype));
    return 2 * per;
}

int ultra_rspmm_forward_host(const ultra_rspmm_plan *plan, int sum_op, int mul_op, int dtype,
                             const void *h_relation, const void *h_input, void *h_output, int64_t dim,
                             int64_t slab_dim, void *staging, size_t staging_bytes, void *stream_) {
    if (int rc = check_common(plan, sum_op, mul_op, dtype, dim)) return rc;
    if (slab_dim <= 0) { set_error("forward_host: slab_dim must be positive"); return 1; }
    slab_dim = std::min(slab_dim, std::max<int64_t>(dim, 1));
    if (staging_bytes < ultra_rspmm_host_staging_bytes(plan, dim, slab_dim, dtype)) {
        set_error("forward_host: staging too small");
        return 1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t es = dtype == ULTRA_RSPMM_F32 ? 4 : 8;
    const size_t N = (size_t)plan->N, R = (size_t)plan->R;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t sb = scratch_bytes(plan, slab_dim, dtype);
    const size_t per = up(R * slab_dim * es) + 2 * up(N * slab_dim * es) + up(sb);
    char *base = static_cast<char *>(staging);
    struct Buf { char *rel, *in, *out, *scr; } buf[2];
    for (int b = 0; b < 2; b++) {
        char *p = base + (size_t)b * per;
        buf[b].rel = p; p += up(R * slab_dim * es);
        buf[b].in = p; p += up(N * slab_dim * es);
        buf[b].out = p; p += up(N * slab_dim * es);
        buf[b].scr = p;
    }
    // lanes: copy-in (x split), compute, copy-out (x split); events chain them per slab and protect buffer reuse.
    // `split` > 1 issues each slab copy as `split` row blocks on separate streams so that more than one copy engine
    // works on the strided (2 KB per row) transfers.  The lanes (9 streams + 13 events at split 1) are created on the
    // first host call of a plan and reused by every later one (round 1 created and destroyed them per call).
    const int split = std::max(1, std::min(g_host_copy_split.load(), HostLanes::kMax));
    ultra_rspmm_plan *mplan = const_cast<ultra_rspmm_plan *>(plan);
    if (mplan->host_lanes == nullptr) mplan->host_lanes = new HostLanes();
    HostLanes &L = *mplan->host_lanes;
    if (int rc0 = L.ensure(split)) return rc0;
    CUDA_TRY(cudaEventRecord(L.start, stream));
    CUDA_TRY(cudaStreamWaitEvent(L.k, L.start, 0));
    for (int i = 0; i < split; i++) {
        CUDA_TRY(cudaStreamWaitEvent(L.in[i], L.start, 0));
        CUDA_TRY(cudaStreamWaitEvent(L.out[i], L.start, 0));
    }
    auto rows_of = [&](size_t total, int i, size_t &r0, size_t &cnt) {
        const size_t per = (total + split - 1) / split;
        r0 = std::min(total, per * i);
        cnt = std::min(total - r0, per);
    };
    int rc = 0;
    int it = 0;
    for (int64_t c0 = 0; c0 < dim && rc == 0; c0 += slab_dim, it++) {
        const int b = it & 1;
        const int64_t wd = std::min(slab_dim, dim - c0);
        const size_t wbytes = (size_t)wd * es, pitch_h = (size_t)dim * es;
        for (int i = 0; i < split; i++) {
            // buffer b: its previous kernel must be done before H2D overwrites rel / in
            if (it >= 2) CUDA_TRY(cudaStreamWaitEvent(L.in[i], L.ek[b], 0));
            size_t r0, cnt;
            rows_of(R, i, r0, cnt);
            if (cnt > 0)
                CUDA_TRY(cudaMemcpy2DAsync(buf[b].rel + r0 * wbytes, wbytes,
                                           (const char *)h_relation + r0 * pitch_h + c0 * es, pitch_h, wbytes, cnt,
                                           cudaMemcpyHostToDevice, L.in[i]));
            rows_of(N, i, r0, cnt);
            if (cnt > 0)
                CUDA_TRY(cudaMemcpy2DAsync(buf[b].in + r0 * wbytes, wbytes,
                                           (const char *)h_input + r0 * pitch_h + c0 * es, pitch_h, wbytes, cnt,
                                           cudaMemcpyHostToDevice, L.in[i]));
            CUDA_TRY(cudaEventRecord(L.ein[b][i], L.in[i]));
            CUDA_TRY(cudaStreamWaitEvent(L.k, L.ein[b][i], 0));
            // ... and its previous D2H before the kernel rewrites out
            if (it >= 2) CUDA_TRY(cudaStreamWaitEvent(L.k, L.eout[b][i], 0));
        }
        rc = ultra_rspmm_forward(plan, sum_op, mul_op, dtype, buf[b].rel, buf[b].in, buf[b].out, wd, buf[b].scr,
                                 sb, L.k);
        if (rc) break;
        CUDA_TRY(cudaEventRecord(L.ek[b], L.k));
        for (int i = 0; i < split; i++) {
            CUDA_TRY(cudaStreamWaitEvent(L.out[i], L.ek[b], 0));
            size_t r0, cnt;
            rows_of(N, i, r0, cnt);
            if (cnt > 0)
                CUDA_TRY(cudaMemcpy2DAsync((char *)h_output + r0 * pitch_h + c0 * es, pitch_h,
                                           buf[b].out + r0 * wbytes, wbytes, wbytes, cnt, cudaMemcpyDeviceToHost,
                                           L.out[i]));
            CUDA_TRY(cudaEventRecord(L.eout[b][i], L.out[i]));
        }
    }
    // join every lane back into the caller's stream
    CUDA_TRY(cudaEventRecord(L.done, L.k));
    CUDA_TRY(cudaStreamWaitEvent(stream, L.done, 0));
    for (int i = 0; i < split; i++) {
        CUDA_TRY(cudaEventRecord(L.done, L.in[i]));
        CUDA_TRY(cudaStreamWaitEvent(stream, L.done, 0));
        CUDA_TRY(cudaEventRecord(L.done, L.out[i]));
        CUDA_TRY(cudaStreamWaitEvent(stream, L.done, 0));
    }
    return rc;
}

}  // extern "C"
