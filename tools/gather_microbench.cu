// gather_microbench.cu -- design evidence for DESIGN.md section 3.1: how should 128-byte row slices be gathered?
//   A) LDG.128 per lane (8 lanes per row) straight into registers           [what seg_reduce_kernel does]
//   B) cp.async.bulk (TMA, UBLKCP) one 128-byte copy per row into shared memory, mbarrier completion, then LDS.128
// Both sum E randomly chosen rows of an (N, D) fp32 matrix restricted to one 128-byte column tile (row stride D*4,
// the tile is L2-resident like in the real kernel) and report gathered GB/s.  Build: nvcc -arch=sm_100a (tools/bin/).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kWarps = 8;

__global__ void __launch_bounds__(kWarps * 32) gather_ldg(const float *__restrict__ mat, const int *__restrict__ idx,
                                                          int E, int D, int tiles, float *out) {
    const int lane = threadIdx.x & 31, q = lane & 7, sg = lane >> 3;
    const long long warp = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kWarps;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int t = 0; t < tiles; t++) {
        const char *base = reinterpret_cast<const char *>(mat) + (size_t)t * 128 + q * 16;
        for (long long e0 = warp * 16; e0 + 16 <= E; e0 += nwarps * 16) {
            int r[4];
#pragma unroll
            for (int u = 0; u < 4; u++) r[u] = __ldg(idx + e0 + u * 4 + sg);
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                v[u] = __ldg(reinterpret_cast<const float4 *>(base + (unsigned long long)(unsigned)r[u] * (unsigned)(D * 4)));
#pragma unroll
            for (int u = 0; u < 4; u++) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
    }
    if (acc.x == 12345.f) out[0] = acc.x + acc.y + acc.z + acc.w;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// B: each warp owns 2 stages x 32 rows x 128 B.  Every lane issues one bulk copy per stage fill.
__global__ void __launch_bounds__(kWarps * 32) gather_tma(const float *__restrict__ mat, const int *__restrict__ idx,
                                                          int E, int D, int tiles, float *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[kWarps][2];
    const int lane = threadIdx.x & 31, q = lane & 7, sg = lane >> 3, w = threadIdx.x >> 5;
    unsigned char *buf = smem + (size_t)w * 2 * 4096;
    if (lane == 0) {
        for (int s = 0; s < 2; s++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[w][s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const long long warp = (long long)blockIdx.x * kWarps + w;
    const long long nwarps = (long long)gridDim.x * kWarps;
    float4 acc = make_float4(0, 0, 0, 0);
    unsigned phase[2] = {0, 0};
    for (int t = 0; t < tiles; t++) {
        const char *base = reinterpret_cast<const char *>(mat) + (size_t)t * 128;
        auto fill = [&](int s, long long e0) {
            const int r = __ldg(idx + e0 + lane);
            const uint32_t bar = smem_u32(&bars[w][s]);
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4096) : "memory");
            __syncwarp();
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(buf + s * 4096 + lane * 128)),
                           "l"(base + (unsigned long long)(unsigned)r * (unsigned)(D * 4)), "r"(128), "r"(bar)
                         : "memory");
        };
        long long e0 = warp * 32;
        int s = 0;
        if (e0 + 32 <= E) fill(0, e0);
        for (; e0 + 32 <= E; e0 += nwarps * 32, s ^= 1) {
            const long long nxt = e0 + nwarps * 32;
            if (nxt + 32 <= E) fill(s ^ 1, nxt);
            const uint32_t bar = smem_u32(&bars[w][s]);
            unsigned ok = 0;
            while (!ok) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(bar), "r"(phase[s]) : "memory");
            }
            phase[s] ^= 1;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const float4 v = *reinterpret_cast<const float4 *>(buf + s * 4096 + (u * 4 + sg) * 128 + q * 16);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            __syncwarp();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before the next async writes
        }
    }
    if (acc.x == 12345.f) out[0] = acc.x + acc.y + acc.z + acc.w;
}

int main(int argc, char **argv) {
    const int N = 14541, D = 8192, E = 544230 / 32 * 32, tiles = 32;
    float *mat, *out;
    int *idx;
    CK(cudaMalloc(&mat, (size_t)N * D * 4));
    CK(cudaMemset(mat, 0, (size_t)N * D * 4));
    CK(cudaMalloc(&out, 4));
    std::vector<int> h(E);
    srand(1);
    for (int &x : h) x = rand() % N;
    CK(cudaMalloc(&idx, E * 4));
    CK(cudaMemcpy(idx, h.data(), E * 4, cudaMemcpyHostToDevice));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const size_t smem_tma = (size_t)kWarps * 2 * 4096;
    CK(cudaFuncSetAttribute(gather_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    const double bytes = (double)E * 128 * tiles;
    for (int ctas_per_sm : {2, 3, 4, 6, 8}) {
        const int grid = prop.multiProcessorCount * ctas_per_sm;
        for (int which = 0; which < 2; which++) {
            float best = 1e9f;
            for (int rep = 0; rep < 5; rep++) {
                CK(cudaEventRecord(a));
                if (which == 0) gather_ldg<<<grid, kWarps * 32>>>(mat, idx, E, D, tiles, out);
                else gather_tma<<<grid, kWarps * 32, smem_tma>>>(mat, idx, E, D, tiles, out);
                CK(cudaEventRecord(b));
                CK(cudaEventSynchronize(b));
                CK(cudaGetLastError());
                float ms;
                CK(cudaEventElapsedTime(&ms, a, b));
                if (ms < best) best = ms;
            }
            printf("{\"variant\": \"%s\", \"ctas_per_sm\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"gathered_GBs\": %.1f}\n",
                   which == 0 ? "LDG.128 -> registers" : "cp.async.bulk 128B -> smem -> LDS.128", ctas_per_sm,
                   ctas_per_sm * kWarps, best, bytes / best / 1e6);
        }
    }
    return 0;
}
