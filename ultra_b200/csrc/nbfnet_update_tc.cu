// nbfnet_update_tc.cu -- tensor-core (tcgen05) variant of the fused NBFNet layer epilogue:
//     out[m,:] = (shortcut ? x[m,:] : 0) + relu( LayerNorm( W . [x[m,:] ; u[m,:]] + bias ) )
// Same contract as nbfnet_update.cu (the FFMA version, kept as the reference implementation and fallback); this one
// runs the Linear(128 -> 64) on the 5th-generation tensor cores WITHOUT giving up fp32 accuracy:
//   3xTF32: every fp32 operand is split into hi = tf32(a) (mantissa truncated to 10 bits) and lo = a - hi (exact);
//           C = Ah.Wh + Ah.Wl + Al.Wh accumulated in fp32 in TMEM reproduces the fp32 product to ~2^-21 relative
//           (measured: 4.9e-6 max abs error on outputs of magnitude 7, tools/update_tc_test.cu).
// Per 128-row tile: all 512 threads split the prefetched [x | u] rows into the hi / lo operand tiles in shared memory
// (K-major, no-swizzle canonical layout of 8 x 16-byte core matrices), ONE thread issues 48 tcgen05.mma.kind::tf32
// (16 k-steps x 3 products, M = 128, N = 64, K = 8) into a 128-lane x 64-column TMEM accumulator and commits to an
// mbarrier; all 16 warps read the accumulator back with tcgen05.ld (warp w: lanes 32 (w & 3).., columns 16 (w >> 2)..),
// apply bias / LayerNorm (row statistics combined through shared memory in a fixed order) / ReLU / residual (x == hi + lo
// is still in the operand tiles), stage the rows in shared memory and store them coalesced.  The next tile's rows are
// fetched into registers while the tensor core and the epilogue work on the current one.
// SASS: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), UTCATOMSWS (TMEM alloc).
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/ultra_rspmm.h"

#ifndef ULTRA_POLL_NS
#define ULTRA_POLL_NS 64  // sleep between failed mbarrier polls
#endif

namespace ultra {
void set_error(const std::string &msg);

namespace tc {

constexpr int kOut = 64, kK = 128, kBM = 128, kThreads = 512;
constexpr int kChunksK = kK / 4;                   // 16-byte chunks along K (4 fp32 each): 32
constexpr int kSboBytes = kChunksK * 128;          // core-matrix row-block stride: 4096 B
constexpr int kATileBytes = kBM * kK * 4;          // 64 KB
constexpr int kWTileBytes = kOut * kK * 4;         // 32 KB
constexpr int kLoads = kBM * kChunksK / kThreads;   // 16-byte loads per thread per tile: 8
constexpr size_t kSmemBytes = 2 * kATileBytes + 2 * kWTileBytes + 1024 + 3 * 64 * 4 + 2 * 4 * kBM * 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r, k) inside a K-major no-swizzle operand tile: 8 x 16 B core matrices, the 32 core
// matrices of one 8-row block are consecutive along K (LBO = 128 B), row blocks are kSboBytes apart (SBO)
__device__ __forceinline__ uint32_t op_offset(int r, int k) {
    return (uint32_t)((r >> 3) * kSboBytes + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);              // start address
    d |= (uint64_t)((128 >> 4) & 0x3fff) << 16;          // leading byte offset: next core matrix along K
    d |= (uint64_t)((kSboBytes >> 4) & 0x3fff) << 32;    // stride byte offset: next 8-row block
    d |= (uint64_t)1 << 46;                              // descriptor version 1 (sm_100)
    return d;                                            // layout_type = SWIZZLE_NONE (0), base_offset 0
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 64, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kOut >> 3) << 17) | ((kBM >> 4) << 24);

__device__ __forceinline__ void split_tf32(float a, float &hi, float &lo) {
    hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
    lo = a - hi;
}

__global__ void __launch_bounds__(kThreads, 1)
update_tc_kernel(const float *__restrict__ x, const float *__restrict__ u, const float *__restrict__ W,
                 const float *__restrict__ bias, const float *__restrict__ gamma, const float *__restrict__ beta,
                 float *__restrict__ out, long long M, float eps, int layer_norm, int relu, int shortcut) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *a_hi = smem, *a_lo = smem + kATileBytes;
    unsigned char *w_hi = smem + 2 * kATileBytes, *w_lo = w_hi + kWTileBytes;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(w_lo + kWTileBytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);
    unsigned char *stage = a_hi;  // aliases the hi operand tile once its last reader (MMA, residual) is done:
                                  // [128 rows][16 chunks ^ (row & 15)] x 16 B
    float *prm = reinterpret_cast<float *>(w_lo + kWTileBytes + 1024);  // bias | gamma | beta
    float *red = prm + 192;                                             // [2][4 column groups][128 rows] LN partials
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time: W hi / lo operand tiles, mbarrier, TMEM allocation -----------------------------------------
    // W arrives transposed (weight_t[k][n], the layout nbfnet_update.cu uses): operand element (n, k) = weight_t[k*64+n]
    for (int idx = tid; idx < kOut * kChunksK; idx += kThreads) {
        const int n = idx & 7 | ((idx >> 8) << 3), c = (idx >> 3) & 31;  // lanes: n%8 fastest, then chunk
        float4 v;
        v.x = __ldg(W + (size_t)(c * 4 + 0) * kOut + n);
        v.y = __ldg(W + (size_t)(c * 4 + 1) * kOut + n);
        v.z = __ldg(W + (size_t)(c * 4 + 2) * kOut + n);
        v.w = __ldg(W + (size_t)(c * 4 + 3) * kOut + n);
        float4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        *reinterpret_cast<float4 *>(w_hi + op_offset(n, c * 4)) = h;
        *reinterpret_cast<float4 *>(w_lo + op_offset(n, c * 4)) = l;
    }
    if (tid < 64) {
        prm[tid] = bias ? bias[tid] : 0.f;
        prm[64 + tid] = gamma ? gamma[tid] : 1.f;
        prm[128 + tid] = beta ? beta[tid] : 0.f;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    unsigned phase = 0;

    const long long num_tiles = (M + kBM - 1) / kBM;
    // software pipeline: the NEXT tile's [x | u] rows are fetched into registers while the tensor core and the
    // epilogue work on the current tile (8 independent 16-byte loads per thread = 64 KB in flight per SM)
    float4 pre[kLoads];
    auto prefetch = [&](long long t_) {
        const long long mm = t_ * kBM;
#pragma unroll
        for (int t = 0; t < kLoads; t++) {
            const int idx = tid + t * kThreads;
            const int rr = idx & 7, cc = (idx >> 3) & 3, cg = (idx >> 5) & 7, rb = idx >> 8;
            const int r = rb * 8 + rr, c = cg * 4 + cc;   // chunk c of the 128-wide [x|u] row
            pre[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t_ < num_tiles && mm + r < M) {
                const float *src = (c < 16) ? (x + (mm + r) * kOut + c * 4) : (u + (mm + r) * kOut + (c - 16) * 4);
                pre[t] = __ldg(reinterpret_cast<const float4 *>(src));
            }
        }
    };
    prefetch(blockIdx.x);
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long long m0 = tile * kBM;
        // ---- split the prefetched [x | u] rows and store the operand tiles (lanes: row%8 fastest, then 4 chunks) ----
#pragma unroll
        for (int t = 0; t < kLoads; t++) {
            const int idx = tid + t * kThreads;
            const int rr = idx & 7, cc = (idx >> 3) & 3, cg = (idx >> 5) & 7, rb = idx >> 8;
            const int r = rb * 8 + rr, c = cg * 4 + cc;
            const float4 v = pre[t];
            float4 h, l;
            split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
            *reinterpret_cast<float4 *>(a_hi + op_offset(r, c * 4)) = h;
            *reinterpret_cast<float4 *>(a_lo + op_offset(r, c * 4)) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> tensor-core reads
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // ---- MMA: one thread, 3 products x 16 k-steps into the same TMEM accumulator -------------------------------
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // descriptors differ only in the start-address field (16-byte units): +16 per k-step of 8 tf32 (256 B)
            const uint64_t dah = make_desc(smem_u32(a_hi)), dal = make_desc(smem_u32(a_lo));
            const uint64_t dwh = make_desc(smem_u32(w_hi)), dwl = make_desc(smem_u32(w_lo));
#pragma unroll
            for (int prod = 0; prod < 3; prod++) {
                const uint64_t da0 = prod == 2 ? dal : dah, db0 = prod == 1 ? dwl : dwh;
#pragma unroll
                for (int ks = 0; ks < kK / 8; ks++) {
                    const uint32_t acc = (prod | ks) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                                 ::"r"(tmem_base), "l"(da0 + (uint64_t)(ks * 16)), "l"(db0 + (uint64_t)(ks * 16)),
                                   "r"(kIdesc), "r"(acc) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         ::"r"(smem_u32(mbar)) : "memory");
        }
        prefetch(tile + gridDim.x);  // global loads of the next tile overlap the MMAs and the epilogue
        // ---- epilogue part 1: all 16 warps.  Warp w reads TMEM lanes 32*(w&3).. (rows) x columns 16*(w>>2).. -------
        {
            // back-off between polls: failed try_waits are wavefronts on the L1/shared data pipe that also feeds the
            // tensor core's operand reads (ncu: 6 650 poll wavefronts per tile vs 2 304 operand wavefronts)
            unsigned done = 0;
            while (true) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(smem_u32(mbar)), "r"(phase) : "memory");
                if (done) break;
                __nanosleep(ULTRA_POLL_NS);
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int qd = warp & 3, grp = warp >> 2, r = qd * 32 + lane, cb = grp * 16;
            float v[16];
            uint32_t *vr = reinterpret_cast<uint32_t *>(v);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                         "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(vr[0]), "=r"(vr[1]), "=r"(vr[2]), "=r"(vr[3]), "=r"(vr[4]), "=r"(vr[5]), "=r"(vr[6]),
                           "=r"(vr[7]), "=r"(vr[8]), "=r"(vr[9]), "=r"(vr[10]), "=r"(vr[11]), "=r"(vr[12]), "=r"(vr[13]),
                           "=r"(vr[14]), "=r"(vr[15])
                         : "r"(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)cb));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            float xres[16];  // residual: x == hi + lo exactly, still in the operand tiles (chunks 0..15 of the row)
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
                float4 xh = make_float4(0.f, 0.f, 0.f, 0.f), xl = xh;
                if (shortcut) {
                    xh = *reinterpret_cast<const float4 *>(a_hi + op_offset(r, cb + c));
                    xl = *reinterpret_cast<const float4 *>(a_lo + op_offset(r, cb + c));
                }
                xres[c] = xh.x + xl.x; xres[c + 1] = xh.y + xl.y; xres[c + 2] = xh.z + xl.z; xres[c + 3] = xh.w + xl.w;
            }
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 16; c++) { v[c] += prm[cb + c]; s += v[c]; }
            if (!layer_norm) __syncthreads();  // every residual read of a_hi done before it is reused as staging
            if (layer_norm) {  // row statistics across the 4 column groups through shared memory (fixed order)
                red[grp * kBM + r] = s;
                __syncthreads();
                const float mean = (red[r] + red[kBM + r] + red[2 * kBM + r] + red[3 * kBM + r]) * (1.f / kOut);
                float q = 0.f;
#pragma unroll
                for (int c = 0; c < 16; c++) { v[c] -= mean; q += v[c] * v[c]; }
                red[4 * kBM + grp * kBM + r] = q;
                __syncthreads();
                const float var = (red[4 * kBM + r] + red[5 * kBM + r] + red[6 * kBM + r] + red[7 * kBM + r]) * (1.f / kOut);
                const float rstd = 1.f / sqrtf(var + eps);
#pragma unroll
                for (int c = 0; c < 16; c++) v[c] = v[c] * rstd * prm[64 + cb + c] + prm[128 + cb + c];
            }
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
                float4 o = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                o.x += xres[c]; o.y += xres[c + 1]; o.z += xres[c + 2]; o.w += xres[c + 3];
                *reinterpret_cast<float4 *>(stage + r * 256 + ((((cb + c) >> 2) ^ (r & 15)) << 4)) = o;
            }
        }
        phase ^= 1u;
        __syncthreads();  // accumulator drained, staged rows visible
        // ---- epilogue part 2: all threads, coalesced residual add + store ----------------------------------------
        {
            constexpr int kIt = kBM * 16 / kThreads;  // float4 per thread: 4
            float4 o[kIt];
#pragma unroll
            for (int t = 0; t < kIt; t++) {
                const int idx = tid + t * kThreads;
                // a warp covers 8 rows x 4 chunks: conflict-free reads of the core-matrix operand layout and the staging
                const int r = (idx >> 7) * 8 + (idx & 7), c4 = ((idx >> 5) & 3) * 4 + ((idx >> 3) & 3);
                o[t] = *reinterpret_cast<const float4 *>(stage + r * 256 + ((c4 ^ (r & 15)) << 4));
            }
#pragma unroll
            for (int t = 0; t < kIt; t++) {
                const int idx = tid + t * kThreads;
                const int r = (idx >> 7) * 8 + (idx & 7), c4 = ((idx >> 5) & 3) * 4 + ((idx >> 3) & 3);
                if (m0 + r < M) __stcs(reinterpret_cast<float4 *>(out + (m0 + r) * kOut) + c4, o[t]);
            }
        }
        __syncthreads();  // operand tiles (residual source) and staging free again before the next tile's stores
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base));
}


}  // namespace tc

int launch_update_tc(const float *x, const float *u, const float *W, const float *bias, const float *gamma,
                     const float *beta, float *out, int64_t num_rows, float eps, int layer_norm, int relu, int shortcut,
                     int sms, cudaStream_t stream) {
    using namespace tc;
    static thread_local int attr_dev = -1;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err == cudaSuccess && attr_dev != dev) {
        err = cudaFuncSetAttribute(update_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        attr_dev = dev;
    }
    if (err != cudaSuccess) { set_error(std::string("nbfnet_update_tc: ") + cudaGetErrorString(err)); return 2; }
    const long long tiles = (num_rows + kBM - 1) / kBM;
    const int grid = (int)(tiles < sms ? tiles : sms);
    update_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(x, u, W, bias, gamma, beta, out, (long long)num_rows, eps,
                                                             layer_norm, relu, shortcut);
    err = cudaGetLastError();
    if (err != cudaSuccess) { set_error(std::string("nbfnet_update_tc launch: ") + cudaGetErrorString(err)); return 2; }
    return 0;
}

}  // namespace ultra
