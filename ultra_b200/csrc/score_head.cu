// score_head.cu -- the scoring MLP of the entity model over EVERY (query, entity) pair in one pass (SURVEY.md section 8f,
// row f4).  The reference scores a tail t of query b with
//     feature = cat[ h[b, t, :], query[b, :] ]                 (/root/reference/ultra/models.py:166-170: final states + query)
//     score   = Linear(128 -> 1)( relu( Linear(128 -> 128)(feature) ) )          (models.py:114-118 mlp, 202-208)
// and the evaluation protocol asks for all N tails of every query (tasks.py:79-91).  The first Linear splits into its
// h half and its query half; the query half (+ bias) is one (B, 128) GEMM on the caller's side, so per row m = b * N + t
//     score[m] = w2 . relu( W1h . h[m, :] + qb[b, :] ) + b2           W1h (128, 64), qb = W1q . query + b1
// In eager PyTorch that is, per chunk of heads, a (rows, 64) x (64, 128) GEMM that writes a (rows, 128) fp32
// intermediate (2 x the size of h), a broadcast add, a clamp, a GEMV and a bias add over it: 4.2 ms of the 40 ms full
// forward at FB15k-237 scale with 256 queries, 5.6 of 30 ms at CoDEx-L scale (ncu launch list, profiles/r02_forward_breakdown.txt).
// Here h is read once (the HBM floor of the op: rows x 256 B) and (rows) floats are written; the hidden layer lives in
// registers.
//
// fp32 FFMA, K accumulated in order like an fp32 GEMM (the parity bar of the path is 1e-5; a single TF32/BF16 MMA is
// not an option, and at K = 64 the op is ~16 KFLOP per 256 loaded bytes: FFMA-bound at ~1 ms, not worth a second
// BF16x3 tensor-core pipeline next to nbfnet_update.cu).  CTA = 256 threads, 2 per SM; tile = 128 rows of h, double
// buffered in shared memory with cp.async (LDGSTS), W1h^T (64 x 128, 32 KB) resident for the CTA's life.
// Thread (ty, tx) owns 8 rows (every other row of a 16-row band) and columns tx*4..+4 and 64+tx*4..+4: per 4 k-steps 8 LDS.128 of A (two distinct
// addresses per warp) and 8 LDS.128 of B (256 contiguous bytes per half-warp) feed 256 FFMA.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/ultra_rspmm.h"

namespace ultra {
void set_error(const std::string &msg);

namespace {

constexpr int kK = 64, kH = 128, kRows = 128, kThreads = 256;
constexpr int kLdx = kK + 4;  // padded row of the h tile (floats): keeps rows 16-byte aligned for LDS.128 / cp.async
constexpr size_t kXBytes = (size_t)kRows * kLdx * sizeof(float);
constexpr size_t kSmemBytes = 2 * kXBytes + (size_t)kK * kH * sizeof(float);  // 2 x 34 KB + 32 KB = 100 KB

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool valid) {
    // src-size 0 zero-fills the 16 destination bytes (rows past the end of the matrix)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}

__device__ __forceinline__ void stage_tile(uint32_t xs_addr, const float *__restrict__ X, long long m0, long long M, int tid) {
    // 128 rows x 16 chunks of 16 bytes; consecutive lanes take consecutive chunks of a row (256 contiguous bytes per 16 lanes)
#pragma unroll
    for (int it = 0; it < kRows * (kK / 4) / kThreads; it++) {
        const int idx = it * kThreads + tid;
        const int r = idx >> 4, c4 = idx & 15;
        const bool ok = m0 + r < M;
        cp_async16(xs_addr + (uint32_t)(r * kLdx + c4 * 4) * 4u, X + (ok ? (m0 + r) * kK + c4 * 4 : 0), ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(kThreads, 2)
score_head_kernel(const float *__restrict__ X, const float *__restrict__ W1t, const float *__restrict__ qb,
                  const float *__restrict__ w2, const float *__restrict__ b2p, float *__restrict__ out, long long M,
                  long long rows_per_head) {
    extern __shared__ __align__(16) float smem_f[];
    float *Xs0 = smem_f, *Ws = smem_f + 2 * kRows * kLdx;
    const uint32_t xs_addr0 = (uint32_t)__cvta_generic_to_shared(smem_f);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    // rows of this thread: rbase + 2 * i.  The two half-warps of a warp (ty even / odd) then read ADJACENT rows, 272 bytes
    // apart = 4 banks apart, so an A-operand LDS.128 is one wavefront (8 rows apart they would share their banks)
    const int rbase = (ty >> 1) * 16 + (ty & 1), colA = tx * 4, colB = 64 + tx * 4;
    const long long num_tiles = (M + kRows - 1) / kRows;
    long long tile = blockIdx.x;
    if (tile >= num_tiles) return;

    stage_tile(xs_addr0, X, tile * kRows, M, tid);  // first tile in flight while W is copied
    for (int idx = tid; idx < kK * kH / 4; idx += kThreads)
        reinterpret_cast<float4 *>(Ws)[idx] = __ldg(reinterpret_cast<const float4 *>(W1t) + idx);
    const float4 w2a = __ldg(reinterpret_cast<const float4 *>(w2 + colA));
    const float4 w2b = __ldg(reinterpret_cast<const float4 *>(w2 + colB));
    const float b2 = __ldg(b2p);

    int buf = 0;
    for (; tile < num_tiles; tile += gridDim.x) {
        const long long m0 = tile * kRows;
        const long long next = tile + gridDim.x;
        // the other buffer was last read in the previous iteration, which ended with a CTA barrier
        if (next < num_tiles) stage_tile(xs_addr0 + (uint32_t)((buf ^ 1) * kXBytes), X, next * kRows, M, tid);
        if (next < num_tiles) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // this tile's rows (and, the first time, W) are visible to every thread

        const float *Xs = Xs0 + (size_t)buf * kRows * kLdx;
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
#pragma unroll 2
        for (int k = 0; k < kK; k += 4) {
            float4 a[8];
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = *reinterpret_cast<const float4 *>(Xs + (rbase + 2 * i) * kLdx + k);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const float4 b0 = *reinterpret_cast<const float4 *>(Ws + (k + kk) * kH + colA);
                const float4 b1 = *reinterpret_cast<const float4 *>(Ws + (k + kk) * kH + colB);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
                    acc[i][0] += av * b0.x; acc[i][1] += av * b0.y; acc[i][2] += av * b0.z; acc[i][3] += av * b0.w;
                    acc[i][4] += av * b1.x; acc[i][5] += av * b1.y; acc[i][6] += av * b1.z; acc[i][7] += av * b1.w;
                }
            }
        }
        // hidden = relu(acc + qb[head]); score = w2 . hidden + b2: this thread's 8 columns, then the 16 lanes of the row
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const long long m = m0 + rbase + 2 * i;
            const long long head = (m < M ? m : M - 1) / rows_per_head;
            const float4 qa = __ldg(reinterpret_cast<const float4 *>(qb + head * kH + colA));
            const float4 qc = __ldg(reinterpret_cast<const float4 *>(qb + head * kH + colB));
            float s = fmaxf(acc[i][0] + qa.x, 0.f) * w2a.x;
            s += fmaxf(acc[i][1] + qa.y, 0.f) * w2a.y;
            s += fmaxf(acc[i][2] + qa.z, 0.f) * w2a.z;
            s += fmaxf(acc[i][3] + qa.w, 0.f) * w2a.w;
            s += fmaxf(acc[i][4] + qc.x, 0.f) * w2b.x;
            s += fmaxf(acc[i][5] + qc.y, 0.f) * w2b.y;
            s += fmaxf(acc[i][6] + qc.z, 0.f) * w2b.z;
            s += fmaxf(acc[i][7] + qc.w, 0.f) * w2b.w;
#pragma unroll
            for (int off = 8; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);  // stays inside the half-warp
            if (tx == 0 && m < M) out[m] = s + b2;
        }
        __syncthreads();  // everyone is done with this buffer before the next iteration's prefetch overwrites it
        buf ^= 1;
    }
}

}  // namespace
}  // namespace ultra

extern "C" int ultra_score_head_f32(const float *h, const float *w1h_t, const float *qb, const float *w2,
                                    const float *b2, float *out, int64_t num_rows, int64_t rows_per_head, int64_t dim,
                                    int64_t hidden, void *stream_) {
    using namespace ultra;
    if (dim != kK || hidden != kH) { set_error("score_head: only dim == 64, hidden == 128 is implemented"); return 1; }
    if (num_rows < 0 || rows_per_head <= 0) { set_error("score_head: bad sizes"); return 1; }
    if (num_rows == 0) return 0;
    for (const void *p : {(const void *)h, (const void *)w1h_t, (const void *)qb, (const void *)w2})
        if (p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) != 0) {
            set_error("score_head: pointers must be non-null and 16-byte aligned");
            return 1;
        }
    if (out == nullptr || b2 == nullptr) { set_error("score_head: out / b2 is null"); return 1; }
    int dev = 0, sms = 148;
    static thread_local int attr_dev = -1;
    cudaError_t err = cudaGetDevice(&dev);
    if (err == cudaSuccess && attr_dev != dev) {
        err = cudaFuncSetAttribute(score_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        attr_dev = dev;
    }
    if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) { set_error(std::string("score_head: ") + cudaGetErrorString(err)); return 2; }
    const long long tiles = (num_rows + kRows - 1) / kRows;
    const int grid = (int)(tiles < 2LL * sms ? tiles : 2LL * sms);
    score_head_kernel<<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream_)>>>(
        h, w1h_t, qb, w2, b2, out, (long long)num_rows, (long long)rows_per_head);
    err = cudaGetLastError();
    if (err != cudaSuccess) { set_error(std::string("score_head launch: ") + cudaGetErrorString(err)); return 2; }
    return 0;
}
