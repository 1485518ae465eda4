// relation_project.cu -- the per-layer relation projections of the entity stack, all layers in ONE launch (SURVEY.md
// section 8f, row f4).  The reference runs, inside every one of the 6 entity layers,
//     relation = relation_projection(self.relation)        nn.Sequential(Linear(64,64), ReLU, Linear(64,64))
// on the SAME (B, R, 64) relation representations (/root/reference/ultra/layers.py:60-64,80): 18 small library
// launches per forward.  Here:  Y[l, m, :] = W2[l] . relu(W1[l] . X[m, :] + b1[l]) + b2[l]   for l < L, m < M = B * R.
//
// fp32 FFMA (K = 64 accumulated in order, like the reference's fp32 Linear; the whole op is ~6 GFLOP at C2 size, far
// below anything worth a tensor-core pipeline).  CTA = 256 threads, 64-row tile of X resident in shared memory for all
// L layers; per layer W1^T then W2^T (16 KB each) pass through one smem buffer; thread (ty, tx) owns rows ty*4..+4,
// columns tx*4..+4 -> per k: 4 broadcast LDS + 1 LDS.128 feed 16 FFMA.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/ultra_rspmm.h"

namespace ultra {
void set_error(const std::string &msg);

namespace {

constexpr int kDim = 64, kRows = 64, kThreads = 256;
constexpr int kLd = kDim + 4;  // padded row (floats): rows ty*4+i of one warp's two ty values fall on different banks
constexpr size_t kSmemBytes = (size_t)(2 * kRows * kLd + kDim * kDim) * sizeof(float);

__device__ __forceinline__ void load_w(float *Ws, const float *__restrict__ Wt, int tid) {  // Wt: [k][n] row-major
    for (int idx = tid; idx < kDim * kDim / 4; idx += kThreads)
        reinterpret_cast<float4 *>(Ws)[idx] = __ldg(reinterpret_cast<const float4 *>(Wt) + idx);
}

// acc[i][j] = sum_k A[row0 + i][k] * Ws[k][col0 + j]
__device__ __forceinline__ void tile_gemm(const float *As, const float *Ws, int row0, int col0, float (&acc)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
#pragma unroll 8
    for (int k = 0; k < kDim; k++) {
        const float4 b = *reinterpret_cast<const float4 *>(Ws + k * kDim + col0);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float a = As[(row0 + i) * kLd + k];
            acc[i][0] += a * b.x; acc[i][1] += a * b.y; acc[i][2] += a * b.z; acc[i][3] += a * b.w;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
relation_project_kernel(const float *__restrict__ X, const float *__restrict__ W1t, const float *__restrict__ b1,
                        const float *__restrict__ W2t, const float *__restrict__ b2, float *__restrict__ Y,
                        long long M, int L) {
    extern __shared__ __align__(16) float smem_f[];  // 51 200 B: above the 48 KB static limit -> dynamic, opted in
    float *Xs = smem_f, *Hs = smem_f + kRows * kLd, *Ws = smem_f + 2 * kRows * kLd;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = ty * 4, col0 = tx * 4;
    for (long long tile = blockIdx.x; tile * kRows < M; tile += gridDim.x) {
        const long long m0 = tile * kRows;
        __syncthreads();  // previous tile's readers of Xs / Ws are done
        for (int idx = tid; idx < kRows * kDim / 4; idx += kThreads) {
            const int r = idx >> 4, c4 = idx & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < M) v = __ldg(reinterpret_cast<const float4 *>(X + (m0 + r) * kDim) + c4);
            *reinterpret_cast<float4 *>(Xs + r * kLd + c4 * 4) = v;
        }
        for (int l = 0; l < L; l++) {
            load_w(Ws, W1t + (size_t)l * kDim * kDim, tid);
            __syncthreads();
            float acc[4][4];
            tile_gemm(Xs, Ws, row0, col0, acc);
            const float4 bb = __ldg(reinterpret_cast<const float4 *>(b1 + (size_t)l * kDim + col0));
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float4 h = make_float4(fmaxf(acc[i][0] + bb.x, 0.f), fmaxf(acc[i][1] + bb.y, 0.f),
                                       fmaxf(acc[i][2] + bb.z, 0.f), fmaxf(acc[i][3] + bb.w, 0.f));
                *reinterpret_cast<float4 *>(Hs + (row0 + i) * kLd + col0) = h;
            }
            __syncthreads();  // Hs complete, Ws (W1) no longer read
            load_w(Ws, W2t + (size_t)l * kDim * kDim, tid);
            __syncthreads();
            tile_gemm(Hs, Ws, row0, col0, acc);
            const float4 b2v = __ldg(reinterpret_cast<const float4 *>(b2 + (size_t)l * kDim + col0));
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (m0 + row0 + i < M)
                    *reinterpret_cast<float4 *>(Y + ((size_t)l * M + m0 + row0 + i) * kDim + col0) =
                        make_float4(acc[i][0] + b2v.x, acc[i][1] + b2v.y, acc[i][2] + b2v.z, acc[i][3] + b2v.w);
            }
            __syncthreads();  // Ws (W2) and Hs free for the next layer
        }
    }
}

}  // namespace
}  // namespace ultra

extern "C" int ultra_relation_project_f32(const float *x, const float *w1_t, const float *b1, const float *w2_t,
                                          const float *b2, float *out, int64_t num_rows, int64_t dim,
                                          int64_t num_layers, void *stream_) {
    using namespace ultra;
    if (dim != kDim) { set_error("relation_project: only dim == 64 is implemented"); return 1; }
    if (num_rows < 0 || num_layers < 0) { set_error("relation_project: negative size"); return 1; }
    if (num_rows == 0 || num_layers == 0) return 0;
    for (const void *p : {(const void *)x, (const void *)w1_t, (const void *)b1, (const void *)w2_t, (const void *)b2,
                          (const void *)out})
        if (p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) != 0) {
            set_error("relation_project: pointers must be non-null and 16-byte aligned");
            return 1;
        }
    int dev = 0, sms = 148;
    static thread_local int attr_dev = -1;
    cudaError_t err = cudaGetDevice(&dev);
    if (err == cudaSuccess && attr_dev != dev) {
        err = cudaFuncSetAttribute(relation_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        attr_dev = dev;
    }
    if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) { set_error(std::string("relation_project: ") + cudaGetErrorString(err)); return 2; }
    const long long tiles = (num_rows + kRows - 1) / kRows;
    const int grid = (int)(tiles < 4LL * sms ? tiles : 4LL * sms);
    relation_project_kernel<<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream_)>>>(
        x, w1_t, b1, w2_t, b2, out, (long long)num_rows, (int)num_layers);
    err = cudaGetLastError();
    if (err != cudaSuccess) { set_error(std::string("relation_project launch: ") + cudaGetErrorString(err)); return 2; }
    return 0;
}
