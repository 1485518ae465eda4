/* GPU-side plain-C consumer of the drop-in boundary (include/ultra_rspmm.h): what a non-Python binder (cgo / JNI /
 * N-API) would do - CUDA runtime C API for memory, then plan_build -> forward -> backward through the C ABI, checked
 * against the hand case of SURVEY.md section 8c (values produced by the reference CPU kernel):
 *   edge_index = [[0,0,1],[1,1,2]], edge_type = [0,0,1], w = 1, relation = [[2,3],[1,1]], input = [[1,1],[5,7],[4,4]]
 *   add.mul -> [[20,42],[4,4],[0,0]];  max.mul -> [[10,21],[4,4],[-FLT_MAX,-FLT_MAX]]
 *   backward with ones: relation_grad = [[10,14],[4,4]], input_grad = [[0,0],[4,6],[1,1]], weight_grad = [31,31,8]
 * Built and run by tests/test_abi_gpu.py (gcc -std=c99 -pedantic, -lcudart). */
#include <cuda_runtime_api.h>
#include <float.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "ultra_rspmm.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 10; } } while (0)
#define RC(x) do { int r_ = (x); if (r_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #x, r_, ultra_rspmm_last_error()); return 20; } } while (0)

static int same(const float *a, const float *b, int n) {
    int i;
    for (i = 0; i < n; i++) if (a[i] != b[i]) return 0;
    return 1;
}

int main(void) {
    const int64_t h_ei[6] = {0, 0, 1, 1, 1, 2};
    const int64_t h_et[3] = {0, 0, 1};
    const float h_w[3] = {1.f, 1.f, 1.f};
    const float h_rel[4] = {2.f, 3.f, 1.f, 1.f};
    const float h_in[6] = {1.f, 1.f, 5.f, 7.f, 4.f, 4.f};
    const float h_og[6] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    const float exp_add[6] = {20.f, 42.f, 4.f, 4.f, 0.f, 0.f};
    const float exp_max[6] = {10.f, 21.f, 4.f, 4.f, -FLT_MAX, -FLT_MAX};
    const float exp_rg[4] = {10.f, 14.f, 4.f, 4.f};
    const float exp_ig[6] = {0.f, 0.f, 4.f, 6.f, 1.f, 1.f};
    const float exp_wg[3] = {31.f, 31.f, 8.f};
    const int64_t N = 3, E = 3, R = 2, D = 2;
    int64_t *d_ei, *d_et;
    float *d_w, *d_rel, *d_in, *d_out, *d_og, *d_rg, *d_ig, *d_wg;
    void *d_ws, *d_scratch;
    float got[6];
    ultra_rspmm_plan *plan = NULL;
    cudaStream_t stream;
    size_t ws_bytes, sc_bytes;
    int sum;

    CK(cudaSetDevice(0));
    CK(cudaStreamCreate(&stream));
    CK(cudaMalloc((void **)&d_ei, sizeof h_ei)); CK(cudaMalloc((void **)&d_et, sizeof h_et));
    CK(cudaMalloc((void **)&d_w, sizeof h_w)); CK(cudaMalloc((void **)&d_rel, sizeof h_rel));
    CK(cudaMalloc((void **)&d_in, sizeof h_in)); CK(cudaMalloc((void **)&d_out, sizeof h_in));
    CK(cudaMalloc((void **)&d_og, sizeof h_og)); CK(cudaMalloc((void **)&d_rg, sizeof h_rel));
    CK(cudaMalloc((void **)&d_ig, sizeof h_in)); CK(cudaMalloc((void **)&d_wg, sizeof h_w));
    CK(cudaMemcpy(d_ei, h_ei, sizeof h_ei, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_et, h_et, sizeof h_et, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w, h_w, sizeof h_w, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_rel, h_rel, sizeof h_rel, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_in, h_in, sizeof h_in, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_og, h_og, sizeof h_og, cudaMemcpyHostToDevice));

    ws_bytes = ultra_rspmm_plan_workspace_bytes(N, E, R);
    CK(cudaMalloc(&d_ws, ws_bytes));
    RC(ultra_rspmm_plan_build(&plan, d_ei, d_et, N, E, R, d_ws, ws_bytes, stream));
    if (!ultra_rspmm_plan_input_sorted(plan)) { fprintf(stderr, "hand case is sorted\n"); return 30; }
    if (ultra_rspmm_plan_info(plan, 0) != N || ultra_rspmm_plan_info(plan, 1) != E) return 31;
    RC(ultra_rspmm_set_edge_weight(plan, d_w, ULTRA_RSPMM_F32, stream));
    sc_bytes = ultra_rspmm_scratch_bytes(plan, D, ULTRA_RSPMM_F32);
    CK(cudaMalloc(&d_scratch, sc_bytes > 0 ? sc_bytes : 256));

    for (sum = 0; sum < 2; sum++) {
        const int sum_op = sum == 0 ? ULTRA_RSPMM_SUM_ADD : ULTRA_RSPMM_SUM_MAX;
        RC(ultra_rspmm_forward(plan, sum_op, ULTRA_RSPMM_MUL_MUL, ULTRA_RSPMM_F32, d_rel, d_in, d_out, D, d_scratch,
                               sc_bytes, stream));
        RC(ultra_rspmm_backward(plan, sum_op, ULTRA_RSPMM_MUL_MUL, ULTRA_RSPMM_F32, d_rel, d_in, d_out, d_og, d_wg, d_rg,
                                d_ig, D, d_scratch, sc_bytes, stream));
        CK(cudaStreamSynchronize(stream));
        CK(cudaMemcpy(got, d_out, sizeof h_in, cudaMemcpyDeviceToHost));
        if (!same(got, sum == 0 ? exp_add : exp_max, 6)) { fprintf(stderr, "forward mismatch (sum %d)\n", sum); return 40 + sum; }
        CK(cudaMemcpy(got, d_rg, sizeof h_rel, cudaMemcpyDeviceToHost));
        if (!same(got, exp_rg, 4)) { fprintf(stderr, "relation_grad mismatch (sum %d)\n", sum); return 50 + sum; }
        CK(cudaMemcpy(got, d_ig, sizeof h_in, cudaMemcpyDeviceToHost));
        if (!same(got, exp_ig, 6)) { fprintf(stderr, "input_grad mismatch (sum %d)\n", sum); return 60 + sum; }
        CK(cudaMemcpy(got, d_wg, sizeof h_w, cudaMemcpyDeviceToHost));
        if (!same(got, exp_wg, 3)) { fprintf(stderr, "weight_grad mismatch (sum %d)\n", sum); return 70 + sum; }
    }
    /* error convention: an out-of-range index is a code + message, never UB / an exception */
    {
        const int64_t bad_ei[6] = {0, 0, 7, 1, 1, 2};
        ultra_rspmm_plan *bad = NULL;
        CK(cudaMemcpy(d_ei, bad_ei, sizeof bad_ei, cudaMemcpyHostToDevice));
        if (ultra_rspmm_plan_build(&bad, d_ei, d_et, N, E, R, d_ws, ws_bytes, stream) == 0 || bad != NULL) return 80;
        if (strstr(ultra_rspmm_last_error(), "out of range") == NULL) return 81;
    }
    ultra_rspmm_plan_destroy(plan);
    printf("abi gpu consumer ok\n");
    return 0;
}
