// l1_microbench.cu -- L1 data-pipe cost of row gathers that HIT in L1 (hub rows): 128-bit vs 256-bit loads, vs LDS.128
// A 256-row x 128-byte table (32 KB, L1 resident) is gathered with random row ids kept in registers.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE>  // 0: LDG.128 8 lanes/row, 1: LDG.256 4 lanes/row, 2: LDS.128 8 lanes/row (table copied to smem)
__global__ void __launch_bounds__(512) k(const float *__restrict__ tab, int iters, float *out, unsigned seed) {
    __shared__ __align__(128) float stab[256 * 32];
    if (MODE == 2) {
        for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) stab[i] = tab[i];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    unsigned s = seed + blockIdx.x * 977 + threadIdx.x / (MODE == 1 ? 4 : 8) * 31;
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            s = s * 1664525u + 1013904223u;
            const unsigned row = (s >> 16) & 255u;
            if (MODE == 0) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(tab + row * 32) + (lane & 7));
                acc += v.x + v.y + v.z + v.w;
            } else if (MODE == 1) {
                float a, b, c, d, e, f, g, h;
                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=f"(a), "=f"(b), "=f"(c), "=f"(d), "=f"(e), "=f"(f), "=f"(g), "=f"(h)
                             : "l"(tab + row * 32 + (lane & 3) * 8));
                acc += a + b + c + d + e + f + g + h;
            } else {
                const float4 v = *(reinterpret_cast<const float4 *>(stab + row * 32) + (lane & 7));
                acc += v.x + v.y + v.z + v.w;
            }
        }
    }
    if (acc == 1234.5f) out[0] = acc;
}

int main() {
    float *tab, *out;
    CK(cudaMalloc(&tab, 256 * 128));
    CK(cudaMemset(tab, 0, 256 * 128));
    CK(cudaMalloc(&out, 4));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int iters = 4096, grid = 148 * 2;
    const char *names[3] = {"LDG.128 (8 lanes/row, L1 hit)", "LDG.256 (4 lanes/row, L1 hit)", "LDS.128 (8 lanes/row)"};
    for (int mode = 0; mode < 3; mode++) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(a));
            if (mode == 0) k<0><<<grid, 512>>>(tab, iters, out, 1u);
            if (mode == 1) k<1><<<grid, 512>>>(tab, iters, out, 1u);
            if (mode == 2) k<2><<<grid, 512>>>(tab, iters, out, 1u);
            CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, a, b));
            if (ms < best) best = ms;
        }
        const double rows = (double)grid * 512 / (mode == 1 ? 4 : 8) * 4.0 * iters;
        printf("{\"variant\": \"%s\", \"ms\": %.4f, \"rows_per_s_G\": %.2f, \"GBs\": %.0f, \"bytes_per_clk_per_sm\": %.1f}\n",
               names[mode], best, rows / best / 1e6, rows * 128 / best / 1e6, rows * 128 / (best * 1e-3) / 148 / 1.95e9);
    }
    return 0;
}
