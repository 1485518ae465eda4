// rspmm_plan.h -- host-side plan object shared by rspmm_plan.cu (builder) and rspmm_api.cu (launchers).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>

namespace ultra {

// One sorted view of the edge list (see rspmm_kernels.cuh for the layout contract).
struct SegOrder {
    int2 *meta = nullptr;      // (E)
    int *eid = nullptr;        // (E)  plan position -> caller edge id
    int *seg_ptr = nullptr;    // (S+1)
    int *nonempty = nullptr;   // (#non-empty + 1)
    int4 *items = nullptr;     // (num_items), multiple of 4
    int4 *fix = nullptr;       // (num_fix)
    void *w = nullptr;         // (E) weights in this order, dtype of the last set_edge_weight
    int num_seg = 0;
    // Host copies of the cutter's counts.  For the base graph they are exact; while an edge mask is applied
    // (ultra_rspmm_plan_apply_mask) they are UPPER BOUNDS used only to size grids and scratch - the kernels read the
    // exact values from the device-side block `counts`.
    int num_items = 0;
    int num_fix = 0;
    int num_fix_long = 0;      // fix[0 .. num_fix_long) are long segments, the rest are empty segments
    int num_slots = 0;
    const int *counts = nullptr;  // device: {num_items, num_fix, num_fix_long, num_slots, nonempty entries, groups, rest, num_edge}
    // base-graph arrays / exact counts (the active pointers above switch between these and the masked copies)
    int2 *meta_base = nullptr, *meta_mask = nullptr;
    int *eid_base = nullptr, *eid_mask = nullptr;
    int *seg_ptr_base = nullptr, *seg_ptr_mask = nullptr;
    int base_items = 0, base_fix = 0, base_fix_long = 0, base_slots = 0;
};

struct HostLanes;  // rspmm_api.cu: streams / events of the host-buffer pipeline, created on first use

}  // namespace ultra

struct ultra_rspmm_plan {
    ~ultra_rspmm_plan();  // rspmm_api.cu: releases the plan-owned streams / events (never device memory: the caller owns it)
    int64_t N = 0, E = 0, R = 0;
    ultra::SegOrder ord[3];  // FWD (by row), BWD_IN (by col), BWD_REL (by relation)
    int *row_of_edge = nullptr;  // (E) row of each edge in FWD order (weight_grad kernel)
    int *row_of_edge_base = nullptr, *row_of_edge_mask = nullptr;
    cudaStream_t side[2] = {nullptr, nullptr};   // plan-owned side streams / events (plan build + apply_mask run the three
    cudaEvent_t side_ev[2] = {nullptr, nullptr}, fork_ev = nullptr;  // orders side by side); freed by plan_destroy
    bool masked = false;
    void *ws = nullptr;          // the caller's workspace (cutter temporaries live there)
    int *w_nonunit = nullptr;    // device flag: some bound weight != 1
    int *counts = nullptr;       // (3 orders x 8) device copy of {num_items, num_fix, num_fix_long, num_slots, ...}
    int *sched = nullptr;        // (3 orders x 2) device work-queue tickets {next block, CTAs done}; self-resetting
    bool unit_weight = true;
    int w_dtype = -1;
    int input_sorted = 0;
    int device = 0;
    int sm_count = 148;
    ultra::HostLanes *host_lanes = nullptr;  // owned; freed by ultra_rspmm_plan_destroy
    int sm_limit = 0;            // 0 = size persistent grids for every SM; >0 = leave the rest to a concurrent collective
    size_t l2_bytes = 0;
    size_t smem_optin = 0;
};

namespace ultra {
void set_error(const std::string &msg);
bool plan_relation_runs(int64_t N, int64_t R);  // rspmm_api.cu: relation-first secondary sort key for this graph?
size_t plan_workspace_bytes(int64_t N, int64_t E, int64_t R);
int cut_items_host(const int *seg_ptr, int num_seg, int *items_out, int64_t items_cap, int *num_items, int *fix_out,
                   int64_t fix_cap, int *num_fix, int *num_fix_long, int *num_slots, int *nonempty_out,
                   int64_t nonempty_cap, int *num_nonempty);
// Re-cut the plan for the sub-graph that keeps edge e iff keep[e] != 0 (device bytes in the CALLER's edge order), entirely
// on the device and without a host synchronisation; keep == nullptr restores the base graph.
int plan_apply_mask(ultra_rspmm_plan *plan, const unsigned char *keep, cudaStream_t stream);
int plan_build(ultra_rspmm_plan **out, const int64_t *edge_index, const int64_t *edge_type, int64_t N, int64_t E,
               int64_t R, void *workspace, size_t workspace_bytes, cudaStream_t stream);
}  // namespace ultra
