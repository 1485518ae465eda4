// nbfnet_update_tc2.cu -- pipelined tcgen05 variant of the fused NBFNet layer epilogue:
//     out[m,:] = (shortcut ? x[m,:] : 0) + relu( LayerNorm( W . [x[m,:] ; u[m,:]] + bias ) )
// Same contract and the same 3xTF32 arithmetic as nbfnet_update_tc.cu (hi = tf32(a), lo = a - hi,
// C = Ah.Wh + Ah.Wl + Al.Wh in an fp32 TMEM accumulator), but the three stages of a 128-row tile no longer run one after
// the other.  ncu on the first version (profiles/r01_ncu_v5_layer_summary.md): the tensor pipe is busy ~50 % of the
// kernel (the K = 8 tf32 MMAs are operand-fetch bound, ~125 clk each) and sits idle while the 16 warps split operands
// or normalise rows, and vice versa.  Here:
//   * the [x | u] operand tile is a ring of 4 K-chunks (32 fp32 columns each, hi + lo = 32 KB per chunk); as soon as a
//     chunk is split and stored, one thread issues its 12 MMAs and commits them to that chunk's "empty" mbarrier, so the
//     NEXT tile's chunk can be written while the tensor core is still working on the later chunks of this tile;
//   * two TMEM accumulators (2 x 128 columns): the tensor core fills one while all 16 warps run the epilogue
//     (tcgen05.ld, bias, LayerNorm, ReLU, residual, store) of the previous tile out of the other;
//   * the epilogue stores its 64-byte row pieces straight to global memory (full 32-byte sectors; L2 merges the four
//     column groups of a row) and re-reads the residual x from L2 (it was fetched one tile earlier), so no shared-memory
//     staging, no second pass and no operand-tile lifetime to respect.
// Steady state per tile: thread work (split + epilogue) and tensor work (48 MMAs) overlap completely.
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

namespace tc2 {

constexpr int kOut = 64, kK = 128, kBM = 128, kThreads = 512;
constexpr int kStages = 4;                          // K-chunks per tile
constexpr int kStageK = kK / kStages;               // 32 fp32 columns
constexpr int kStageChunks = kStageK / 4;           // 8 sixteen-byte chunks along K
// A operand stages use a PADDED core-matrix pitch along K (leading byte offset 144 B instead of 128 B): 8 lanes that hold
// the 8 sixteen-byte chunks of ONE row then store to 8 different bank groups (c * 144 mod 128 = 16 c), so the global
// loads can be row-contiguous (a warp = 4 rows x 128 B = 4 data-pipe wavefronts) and the shared stores stay
// conflict-free.  With the unpadded pitch a conflict-free store needs 8 different ROWS per 8 lanes, and ncu counted
// ~32 L1 data-pipe wavefronts per such LDG.128 (6 650 per tile, the pipe 100 % busy together with the tensor core's
// operand reads) - that, not the tensor core, bounded the first kernel.
constexpr int kALbo = 144;                          // A: core-matrix pitch along K
constexpr int kASbo = kStageChunks * kALbo;         // A: 8-row block stride inside a stage: 1152 B
constexpr int kAStageBytes = (kBM / 8) * kASbo;     // 18 KB (hi or lo of one stage)
constexpr int kWSbo = (kK / 4) * 128;               // W: 8-row block stride: 4096 B
constexpr int kWTileBytes = kOut * kK * 4;          // 32 KB
constexpr int kItemsPerStage = kBM * kStageChunks / kThreads;  // 16-byte items per thread per stage: 2
constexpr int kSlots = 3;                           // operand ring: K-chunk g = 4 * tile + c lives in slot g % 3
constexpr int kOutStageBytes = kBM * kOut * 4;      // 32 KB: epilogue rows staged for coalesced stores
constexpr int kBarBytes = 128;                      // empty[3], acc_full[2], TMEM slot
constexpr size_t kSmemBytes = 2 * kSlots * kAStageBytes + 2 * kWTileBytes + kOutStageBytes + kBarBytes + 3 * 64 * 4 +
                              2 * 4 * kBM * 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no-swizzle canonical layout: 8 rows x 16 B core matrices (128 contiguous bytes each); the core matrices of
// one 8-row block follow each other along K every `lbo` bytes; 8-row blocks are `sbo` bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
    return d;
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n
// The two products that share the A operand, Ah.Wh and Ah.Wl, are issued as ONE N = 128 MMA against the stacked B
// operand [Wh ; Wl] (w_lo follows w_hi in shared memory = 8 more 8-row blocks): the accumulator is 128 columns
// (Ah.Wh + Al.Wh | Ah.Wl), the epilogue adds the halves.  32 MMAs per tile instead of 48; the tensor-active cycles do
// not change (ncu: hmma_cycles_active is proportional to N) but Ah is fetched from shared memory once instead of twice.
constexpr uint32_t idesc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((kBM >> 4) << 24); }
constexpr int kAccCols = 2 * kOut;                  // TMEM columns per accumulator

__device__ __forceinline__ void split_tf32(float a, float &hi, float &lo) {
    hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
    lo = a - hi;
}

// mbarrier wait with back-off.  Every failed mbarrier.try_wait is a wavefront on the SM's L1/shared data pipe - the
// pipe the tensor core fetches its shared-memory operands through.  ncu on the first tcgen05 kernel (16 warps spinning
// while the MMAs ran) showed 6 650 such wavefronts per 128-row tile against 2 304 operand wavefronts, the data pipe
// 100 % busy and the MMAs starved; so: one immediate try, then sleep between polls.
__device__ __forceinline__ bool mbar_try(uint32_t bar, unsigned parity) {
    unsigned done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity) {
    if (mbar_try(bar, parity)) return;
    do { __nanosleep(ULTRA_POLL_NS); } while (!mbar_try(bar, parity));
}

__global__ void __launch_bounds__(kThreads, 1)
update_tc2_kernel(const float *__restrict__ x, const float *__restrict__ u, const float *__restrict__ W,
                  const float *__restrict__ bias, const float *__restrict__ gamma, const float *__restrict__ beta,
                  float *__restrict__ out, long long M, float eps, int layer_norm, int relu, int shortcut) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *a_hi = smem, *a_lo = smem + kSlots * kAStageBytes;           // [slot][18 KB] each
    unsigned char *w_hi = smem + 2 * kSlots * kAStageBytes, *w_lo = w_hi + kWTileBytes;
    unsigned char *ostage = w_lo + kWTileBytes;  // [128 rows][16 chunks ^ (row & 15)] x 16 B
    uint64_t *bars = reinterpret_cast<uint64_t *>(ostage + kOutStageBytes);     // empty[0..2], acc_full[0..1]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 6);
    float *prm = reinterpret_cast<float *>(ostage + kOutStageBytes + kBarBytes);  // bias | gamma | beta
    float *red = prm + 192;                                                     // [2][4 column groups][128 rows]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = smem_u32(bars);

    // ---- one-time: W hi / lo operand tiles, mbarriers, TMEM allocation ----------------------------------------
    // W arrives transposed (weight_t[k][n]): operand element (n, k) = weight_t[k * 64 + n]
    for (int idx = tid; idx < kOut * (kK / 4); idx += kThreads) {
        const int n = (idx & 7) | ((idx >> 8) << 3), c = (idx >> 3) & 31;
        float4 v;
        v.x = __ldg(W + (size_t)(c * 4 + 0) * kOut + n);
        v.y = __ldg(W + (size_t)(c * 4 + 1) * kOut + n);
        v.z = __ldg(W + (size_t)(c * 4 + 2) * kOut + n);
        v.w = __ldg(W + (size_t)(c * 4 + 3) * kOut + n);
        float4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        const uint32_t off = (uint32_t)((n >> 3) * kWSbo + c * 128 + (n & 7) * 16);
        *reinterpret_cast<float4 *>(w_hi + off) = h;
        *reinterpret_cast<float4 *>(w_lo + off) = l;
    }
    if (tid < 64) {
        prm[tid] = bias ? bias[tid] : 0.f;
        prm[64 + tid] = gamma ? gamma[tid] : 1.f;
        prm[128 + tid] = beta ? beta[tid] : 0.f;
    }
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < 5; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // W tiles: generic-proxy writes -> tensor-core reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const long long num_tiles = (M + kBM - 1) / kBM;
    const long long first = blockIdx.x, step = gridDim.x;
    const int n_my = first < num_tiles ? (int)((num_tiles - first + step - 1) / step) : 0;

    // this thread's two 16-byte items of a stage: chunk = item % 8 (fastest), row = item / 8
    float4 pre[kStages * kItemsPerStage];
    auto fetch_chunk = [&](long long tile, int c, bool live) {
#pragma unroll
        for (int j = 0; j < kItemsPerStage; j++) {
            const int item = tid + j * kThreads;
            const int r = item >> 3, gc = c * kStageChunks + (item & 7);  // 8 lanes = one row's 128 bytes
            const long long row = tile * kBM + r;
            pre[c * kItemsPerStage + j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && row < M) {
                const float *src = gc < 16 ? x + row * kOut + gc * 4 : u + row * kOut + (gc - 16) * 4;
                pre[c * kItemsPerStage + j] = __ldg(reinterpret_cast<const float4 *>(src));
            }
        }
    };
#pragma unroll
    for (int c = 0; c < kStages; c++) fetch_chunk(first, c, n_my > 0);

    const uint64_t dah = make_desc(smem_u32(a_hi), kALbo, kASbo), dal = make_desc(smem_u32(a_lo), kALbo, kASbo);
    const uint64_t dwh = make_desc(smem_u32(w_hi), 128, kWSbo), dwl = make_desc(smem_u32(w_lo), 128, kWSbo);

    for (int i = 0; i <= n_my; i++) {
        const long long tile = first + (long long)i * step;
        // ================= producer + MMA issue for tile i =========================================================
        if (i < n_my) {
#pragma unroll
            for (int c = 0; c < kStages; c++) {
                // ring slot of this K-chunk; its previous user (chunk g - 3) must have been consumed by the tensor core
                const int g = i * kStages + c, slot = g % kSlots, use = g / kSlots;
                if (use >= 1) mbar_wait(bar0 + 8u * slot, (unsigned)(use - 1) & 1u);
#pragma unroll
                for (int j = 0; j < kItemsPerStage; j++) {
                    const int item = tid + j * kThreads;
                    const int r = item >> 3;
                    const uint32_t off = (uint32_t)(slot * kAStageBytes + (r >> 3) * kASbo + (item & 7) * kALbo + (r & 7) * 16);
                    const float4 v = pre[c * kItemsPerStage + j];
                    float4 h, l;
                    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z);
                    split_tf32(v.w, h.w, l.w);
#ifdef ULTRA_TF32_RAW_HI  // experiment: does kind::tf32 ignore the 13 low mantissa bits (hi == raw operand)?
                    *reinterpret_cast<float4 *>(a_hi + off) = v;
#else
                    *reinterpret_cast<float4 *>(a_hi + off) = h;
#endif
                    *reinterpret_cast<float4 *>(a_lo + off) = l;
                }
                fetch_chunk(tile + step, c, i + 1 < n_my);  // next tile's rows of this chunk: in flight for a whole tile
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();
                if (tid == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d_tmem = tmem_base + (uint32_t)(i & 1) * kAccCols;
                    const uint64_t dah0 = dah + (uint64_t)(slot * (kAStageBytes >> 4));
                    const uint64_t dal0 = dal + (uint64_t)(slot * (kAStageBytes >> 4));
                    const uint64_t db0 = dwh + (uint64_t)(c * (kStageK / 8) * 16);
#pragma unroll
                    for (int ks = 0; ks < kStageK / 8; ks++) {  // one MMA = 8 tf32 of K = 2 chunks
                        const uint64_t ka = (uint64_t)(ks * (2 * kALbo >> 4)), kb = (uint64_t)(ks * 16);
                        const uint32_t acc = (c | ks) ? 1u : 0u;
                        // Ah . [Wh ; Wl]^T -> columns 0..127
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                                     ::"r"(d_tmem), "l"(dah0 + ka), "l"(db0 + kb), "r"(idesc(2 * kOut)), "r"(acc) : "memory");
                        // Al . Wh^T -> += columns 0..63
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                                     ::"r"(d_tmem), "l"(dal0 + ka), "l"(db0 + kb), "r"(idesc(kOut)), "r"(1u) : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 ::"r"(bar0 + 8u * slot) : "memory");
                    if (c == kStages - 1)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                     ::"r"(bar0 + 8u * (kSlots + (i & 1))) : "memory");
                }
            }
        }
        // ================= epilogue of tile i - 1 (overlaps the MMAs of tile i) =====================================
        if (i >= 1) {
            const int e = i - 1;
            const long long m0 = (tile - step) * kBM;
            const int qd = warp & 3, grp = warp >> 2, r = qd * 32 + lane, cb = grp * 16;
            // residual rows for part B (this SM fetched them a tile ago: L2 hits), issued before anything can stall
            constexpr int kIt = kBM * 16 / kThreads;  // float4 per thread in part B: 4
            float4 xr[kIt];
#pragma unroll
            for (int t = 0; t < kIt; t++) {
                const int idx = tid + t * kThreads, rr = idx >> 4, c4 = idx & 15;
                xr[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (shortcut && m0 + rr < M) xr[t] = __ldg(reinterpret_cast<const float4 *>(x + (m0 + rr) * kOut) + c4);
            }
            // part A: thread = (row r, 16 columns): accumulator -> bias -> LayerNorm -> ReLU -> staging (XOR-swizzled)
            mbar_wait(bar0 + 8u * (kSlots + (e & 1)), (unsigned)(e >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float v[16], v2[16];
            uint32_t *vr = reinterpret_cast<uint32_t *>(v), *vr2 = reinterpret_cast<uint32_t *>(v2);
            const uint32_t t_acc = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)((e & 1) * kAccCols + cb);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                         "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(vr[0]), "=r"(vr[1]), "=r"(vr[2]), "=r"(vr[3]), "=r"(vr[4]), "=r"(vr[5]), "=r"(vr[6]),
                           "=r"(vr[7]), "=r"(vr[8]), "=r"(vr[9]), "=r"(vr[10]), "=r"(vr[11]), "=r"(vr[12]), "=r"(vr[13]),
                           "=r"(vr[14]), "=r"(vr[15])
                         : "r"(t_acc));
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                         "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(vr2[0]), "=r"(vr2[1]), "=r"(vr2[2]), "=r"(vr2[3]), "=r"(vr2[4]), "=r"(vr2[5]), "=r"(vr2[6]),
                           "=r"(vr2[7]), "=r"(vr2[8]), "=r"(vr2[9]), "=r"(vr2[10]), "=r"(vr2[11]), "=r"(vr2[12]), "=r"(vr2[13]),
                           "=r"(vr2[14]), "=r"(vr2[15])
                         : "r"(t_acc + (uint32_t)kOut));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // ordered before the next barrier
#pragma unroll
            for (int c = 0; c < 16; c++) v[c] += v2[c];  // (Ah.Wh + Al.Wh) + Ah.Wl
            const float4 *prm4 = reinterpret_cast<const float4 *>(prm);
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
                const float4 b = prm4[(cb + c) >> 2];
                v[c] += b.x; v[c + 1] += b.y; v[c + 2] += b.z; v[c + 3] += b.w;
                s += (v[c] + v[c + 1]) + (v[c + 2] + v[c + 3]);
            }
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
                for (int c = 0; c < 16; c += 4) {
                    const float4 ga = prm4[16 + ((cb + c) >> 2)], be = prm4[32 + ((cb + c) >> 2)];
                    v[c] = v[c] * rstd * ga.x + be.x; v[c + 1] = v[c + 1] * rstd * ga.y + be.y;
                    v[c + 2] = v[c + 2] * rstd * ga.z + be.z; v[c + 3] = v[c + 3] * rstd * ga.w + be.w;
                }
            }
            if (!layer_norm) __syncthreads();  // the previous tile's part-B reads of the staging rows are done
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
                float4 o = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                *reinterpret_cast<float4 *>(ostage + r * 256 + ((((cb + c) >> 2) ^ (r & 15)) << 4)) = o;
            }
            __syncthreads();
            // part B: 16 lanes = one 256-byte row: residual add + store, every access a full 128-byte line per 8 lanes
            {
#pragma unroll
                for (int t = 0; t < kIt; t++) {
                    const int idx = tid + t * kThreads, rr = idx >> 4, c4 = idx & 15;
                    float4 o = *reinterpret_cast<const float4 *>(ostage + rr * 256 + ((c4 ^ (rr & 15)) << 4));
                    o.x += xr[t].x; o.y += xr[t].y; o.z += xr[t].z; o.w += xr[t].w;
                    if (m0 + rr < M) __stcs(reinterpret_cast<float4 *>(out + (m0 + rr) * kOut) + c4, o);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base));
}

}  // namespace tc2

int launch_update_tc2(const float *x, const float *u, const float *W, const float *bias, const float *gamma,
                      const float *beta, float *out, int64_t num_rows, float eps, int layer_norm, int relu, int shortcut,
                      int sms, cudaStream_t stream) {
    using namespace tc2;
    static thread_local int attr_dev = -1;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err == cudaSuccess && attr_dev != dev) {
        err = cudaFuncSetAttribute(update_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        attr_dev = dev;
    }
    if (err != cudaSuccess) { set_error(std::string("nbfnet_update_tc2: ") + cudaGetErrorString(err)); return 2; }
    const long long tiles = (num_rows + kBM - 1) / kBM;
    const int grid = (int)(tiles < sms ? tiles : sms);
    update_tc2_kernel<<<grid, kThreads, kSmemBytes, stream>>>(x, u, W, bias, gamma, beta, out, (long long)num_rows, eps,
                                                              layer_norm, relu, shortcut);
    err = cudaGetLastError();
    if (err != cudaSuccess) { set_error(std::string("nbfnet_update_tc2 launch: ") + cudaGetErrorString(err)); return 2; }
    return 0;
}

}  // namespace ultra
