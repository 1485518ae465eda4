// nbfnet_update.cu -- fused NBFNet layer epilogue on the 5th-gen tensor cores:
//     out[m,:] = (shortcut ? x[m,:] : 0) + relu( LayerNorm( W . [x[m,:] ; u[m,:]] + bias ) )
// replaces  GeneralizedRelationalConv.update  (/root/reference/ultra/layers.py:233-240: cat -> Linear(128->64) ->
// LayerNorm -> ReLU) plus the residual add of the Bellman-Ford loop (/root/reference/ultra/models.py:158-160):
// five PyTorch kernels and ~8 passes over the (B, N, 64) state become one pass (read x, read u, write out).
// x, u, out are (M, 64) fp32 row-major with M = heads * nodes -- exactly the (B, N, 64) layer state, no transposes.
//
// Pipeline: one persistent 512-thread CTA per SM, 128-row tiles; the [x|u] tile is cut into four K-chunks that go
// through a ring of 3 operand slots (K-major, no-swizzle core matrices, hand-built smem descriptors); one thread issues
// tcgen05.mma (SASS UTCHMMA) per chunk and commits to the slot's mbarrier; the accumulator is double-buffered in TMEM
// so all 16 warps run the epilogue of tile i-1 (tcgen05.ld, bias, LayerNorm, ReLU, staged coalesced stores) under the
// MMAs of tile i.  (Round 1 also shipped an FFMA kernel and two earlier tcgen05 generations - 0.90 / 0.64 / 0.40 ms at
// C2 size against this kernel's 0.386 ms; they were removed once this one had been validated, see profiles/README.md.)
//
// Arithmetic: the reference's Linear is an fp32 GEMM and the parity bar is 1e-5, so one TF32/BF16 MMA is not an option.
// BF16 keeps fp32's exponent range, so an fp32 value splits EXACTLY into three
// BF16 terms (8 significand bits each, by truncation):  a = a0 + a1 + a2.  The fp32 product is recovered to ~2^-21 by
// the six term products with i + j <= 2 (dropped: a1.w2 + a2.w1 ~ 2^-21, a2.w2 ~ 2^-28):
//     a.w ~= a0.w0 + a0.w1 + a1.w0 + a0.w2 + a1.w1 + a2.w0        (fp32 accumulation in TMEM)
// Products that share the A term are ONE MMA against the stacked B operand [W0 ; W1 ; W2] (192 rows):
//     A0 x [W0;W1;W2] (N = 192)   A1 x [W0;W1] (N = 128)   A2 x [W0] (N = 64)      K = 16 per MMA
// = 3 MMAs per k-step, 24 per 128-row tile (3xTF32: 32), into a 192-column accumulator
//     cols 0..63: a0w0 + a1w0 + a2w0 | cols 64..127: a0w1 + a1w1 | cols 128..191: a0w2 ;  the epilogue adds the thirds.
// Operand tiles are half the bytes of the TF32 ones (3 x 2 B instead of 2 x 4 B per element): 96 KB -> 768 store
// wavefronts per tile instead of 1 024, and 1 536 tensor-core operand wavefronts instead of 1 792.
// Core-matrix pitch along K is padded to 160 B: an 8-lane store phase is 2 rows x 4 chunks, (q * 160 + r * 16) mod 128
// hits 8 different bank groups, while 4 lanes read 128 contiguous bytes of one row from global memory.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/ultra_rspmm.h"

#ifndef ULTRA_TC3_PREFETCH
// Where the NEXT tile's global loads are issued inside a K-chunk step.  0 (default, the measured build): right after the
// operand stores, i.e. BEFORE fence.proxy.async (SASS: MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC.S) and the CTA barrier.  The ncu
// source page of that build (profiles/r01_ncu_v9_update_bf16x3_summary.md) puts the long_scoreboard stalls exactly on
// that FENCE and BAR.SYNC and the barrier stalls on the branch behind them: the membar waits for the just-issued loads,
// which puts a DRAM round trip on the critical path of every chunk step.  1: issue them after the barrier of the same
// step; 2: issue all four chunks after the last step's barrier (the epilogue covers the latency).  To be A/B-measured.
#define ULTRA_TC3_PREFETCH 0
#endif
#ifndef ULTRA_TC3_L2PF
// L2 prefetch distance in tiles (0 = off): at the start of tile i one thread asks L2 for the x- and u-rows of tile
// i + ULTRA_TC3_L2PF with two bulk prefetches (cp.async.bulk.prefetch.L2, 32 KB each, no registers, no shared memory).  The
// register prefetch of the next tile then hits L2 instead of DRAM, and the DRAM requests of 1 + PF tiles are in flight
// per SM instead of one - the kernel's limiter is bytes in flight.
#define ULTRA_TC3_L2PF 0
#endif
#ifndef ULTRA_POLL_NS
#define ULTRA_POLL_NS 64  // sleep between failed mbarrier polls
#endif

namespace ultra {
void set_error(const std::string &msg);

namespace tc3 {

constexpr int kOut = 64, kK = 128, kBM = 128, kThreads = 512;
constexpr int kTerms = 3;
constexpr int kStages = 4;                          // K-chunks per tile
constexpr int kStageK = kK / kStages;               // 32 columns
constexpr int kStageChunks = kStageK / 8;           // 4 sixteen-byte chunks (8 bf16) along K per row and term
constexpr int kALbo = 160;                          // A: padded core-matrix pitch along K
constexpr int kASbo = kStageChunks * kALbo;         // A: 8-row block stride inside a stage: 640 B
constexpr int kATermBytes = (kBM / 8) * kASbo;      // 10 KB: one term of one stage
constexpr int kAStageBytes = kTerms * kATermBytes;  // 30 KB
constexpr int kWSbo = (kK / 8) * 128;               // W: 8-row block stride: 2048 B
constexpr int kWTermBytes = (kOut / 8) * kWSbo;     // 16 KB; [W0 ; W1 ; W2] are consecutive = one 192-row B operand
constexpr int kSlots = 3;                           // operand ring: K-chunk g = 4 * tile + c lives in slot g % 3
constexpr int kOutStageBytes = kBM * kOut * 4;      // 32 KB: epilogue rows staged for coalesced stores
constexpr int kBarBytes = 128;                      // empty[3], acc_full[2], TMEM slot
constexpr int kAccStride = 256;                     // TMEM columns between the two accumulators (192 used)
constexpr size_t kSmemBytes = kSlots * kAStageBytes + kTerms * kWTermBytes + kOutStageBytes + kBarBytes + 3 * 64 * 4 +
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

// instruction descriptor (kind::f16): D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), K-major, M = 128, N = n
constexpr uint32_t idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((kBM >> 4) << 24); }

// a = t0 + t1 + t2 exactly, each term a BF16 (the top 16 bits of an fp32): returns the three bit patterns (low half 0)
__device__ __forceinline__ void split_bf16x3(float a, uint32_t &t0, uint32_t &t1, uint32_t &t2) {
    t0 = __float_as_uint(a) & 0xffff0000u;
    const float r1 = a - __uint_as_float(t0);
    t1 = __float_as_uint(r1) & 0xffff0000u;
    const float r2 = r1 - __uint_as_float(t1);
    t2 = __float_as_uint(r2) & 0xffff0000u;
}
// eight consecutive-k fp32 -> three 16-byte chunks of 8 bf16 (element j at byte 2 j)
__device__ __forceinline__ void split_chunk(const float4 &lo4, const float4 &hi4, uint4 &c0, uint4 &c1, uint4 &c2) {
    const float a[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
    uint32_t t0[8], t1[8], t2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) split_bf16x3(a[j], t0[j], t1[j], t2[j]);
    // pack (even k -> low half, odd k -> high half): bytes {x.b2, x.b3, y.b2, y.b3}
    c0 = make_uint4(__byte_perm(t0[0], t0[1], 0x7632), __byte_perm(t0[2], t0[3], 0x7632),
                    __byte_perm(t0[4], t0[5], 0x7632), __byte_perm(t0[6], t0[7], 0x7632));
    c1 = make_uint4(__byte_perm(t1[0], t1[1], 0x7632), __byte_perm(t1[2], t1[3], 0x7632),
                    __byte_perm(t1[4], t1[5], 0x7632), __byte_perm(t1[6], t1[7], 0x7632));
    c2 = make_uint4(__byte_perm(t2[0], t2[1], 0x7632), __byte_perm(t2[2], t2[3], 0x7632),
                    __byte_perm(t2[4], t2[5], 0x7632), __byte_perm(t2[6], t2[7], 0x7632));
}

// mbarrier wait with back-off
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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t *vr = reinterpret_cast<uint32_t *>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(vr[0]), "=r"(vr[1]), "=r"(vr[2]), "=r"(vr[3]), "=r"(vr[4]), "=r"(vr[5]), "=r"(vr[6]), "=r"(vr[7]),
                   "=r"(vr[8]), "=r"(vr[9]), "=r"(vr[10]), "=r"(vr[11]), "=r"(vr[12]), "=r"(vr[13]), "=r"(vr[14]),
                   "=r"(vr[15])
                 : "r"(taddr));
}

__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
update_tc3_kernel(const float *__restrict__ x, const float *__restrict__ u, const float *__restrict__ W,
                  const float *__restrict__ bias, const float *__restrict__ gamma, const float *__restrict__ beta,
                  float *__restrict__ out, long long M, float eps, int layer_norm, int relu, int shortcut,
                  float *__restrict__ z_out, float *__restrict__ stats_out) {
    // z_out / stats_out (training forward, both or neither): the pre-LayerNorm linear output z (M, 64) and the row
    // statistics (mean, rstd) (M, 2) that ultra_nbfnet_update_bwd_f32 consumes; nullptr for inference
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *a_st = smem;                                       // [slot][term][10 KB]
    unsigned char *w_st = smem + kSlots * kAStageBytes;               // [term][16 KB] = 192-row B operand
    unsigned char *ostage = w_st + kTerms * kWTermBytes;              // [128 rows][16 chunks ^ (row & 15)] x 16 B
    uint64_t *bars = reinterpret_cast<uint64_t *>(ostage + kOutStageBytes);  // empty[0..2], acc_full[0..1]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 6);
    float *prm = reinterpret_cast<float *>(ostage + kOutStageBytes + kBarBytes);  // bias | gamma | beta
    float *red = prm + 192;                                           // [2][4 column groups][128 rows]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = smem_u32(bars);

    // ---- one-time: the three BF16 terms of W, mbarriers, TMEM allocation ---------------------------------------
    // W arrives transposed (weight_t[k][n]): operand element (n, k) = weight_t[k * 64 + n]
    for (int idx = tid; idx < kOut * (kK / 8); idx += kThreads) {
        const int n = idx & 63, ch = idx >> 6;  // chunk ch = k / 8
        float4 lo4, hi4;
        lo4.x = __ldg(W + (size_t)(ch * 8 + 0) * kOut + n); lo4.y = __ldg(W + (size_t)(ch * 8 + 1) * kOut + n);
        lo4.z = __ldg(W + (size_t)(ch * 8 + 2) * kOut + n); lo4.w = __ldg(W + (size_t)(ch * 8 + 3) * kOut + n);
        hi4.x = __ldg(W + (size_t)(ch * 8 + 4) * kOut + n); hi4.y = __ldg(W + (size_t)(ch * 8 + 5) * kOut + n);
        hi4.z = __ldg(W + (size_t)(ch * 8 + 6) * kOut + n); hi4.w = __ldg(W + (size_t)(ch * 8 + 7) * kOut + n);
        uint4 c0, c1, c2;
        split_chunk(lo4, hi4, c0, c1, c2);
        const uint32_t off = (uint32_t)((n >> 3) * kWSbo + ch * 128 + (n & 7) * 16);
        *reinterpret_cast<uint4 *>(w_st + off) = c0;
        *reinterpret_cast<uint4 *>(w_st + kWTermBytes + off) = c1;
        *reinterpret_cast<uint4 *>(w_st + 2 * kWTermBytes + off) = c2;
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
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
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

    // this thread's item of a stage: row = tid / 4, 8 consecutive columns q = tid % 4 (4 lanes = 128 contiguous bytes)
    const int prow = tid >> 2, pq = tid & 3;
    float4 pre[kStages * 2];
    auto fetch_chunk = [&](long long tile, int c, bool live) {
        const long long row = tile * kBM + prow;
        const int col = c * kStageK + pq * 8;  // column of the 128-wide [x | u] row
        pre[2 * c] = make_float4(0.f, 0.f, 0.f, 0.f);
        pre[2 * c + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && row < M) {
            const float *src = col < kOut ? x + row * kOut + col : u + row * kOut + (col - kOut);
#if ULTRA_TC3_PREFETCH == 0
            pre[2 * c] = __ldg(reinterpret_cast<const float4 *>(src));
            pre[2 * c + 1] = __ldg(reinterpret_cast<const float4 *>(src) + 1);
#else  // ordinary (coherent) loads: ptxas hoists ld.global.nc back above the membar + barrier (SASS identical to variant 0)
            asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(pre[2 * c].x), "=f"(pre[2 * c].y),
                         "=f"(pre[2 * c].z), "=f"(pre[2 * c].w) : "l"(src) : "memory");
            asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(pre[2 * c + 1].x), "=f"(pre[2 * c + 1].y),
                         "=f"(pre[2 * c + 1].z), "=f"(pre[2 * c + 1].w) : "l"(src + 4) : "memory");
#endif
        }
    };
#pragma unroll
    for (int c = 0; c < kStages; c++) fetch_chunk(first, c, n_my > 0);

    const uint64_t da_base = make_desc(smem_u32(a_st), kALbo, kASbo);
    const uint64_t dw_base = make_desc(smem_u32(w_st), 128, kWSbo);
    const uint32_t a_off = (uint32_t)((prow >> 3) * kASbo + pq * kALbo + (prow & 7) * 16);

    for (int i = 0; i <= n_my; i++) {
        const long long tile = first + (long long)i * step;
#if ULTRA_TC3_L2PF > 0
        if (tid < 2) {  // thread 0: x rows, thread 1: u rows of the tile ULTRA_TC3_L2PF ahead
            const long long pf_row = (tile + (long long)ULTRA_TC3_L2PF * step) * kBM;
            if (pf_row < M) {
                const long long rows = (M - pf_row) < kBM ? (M - pf_row) : kBM;
                const float *src = (tid == 0 ? x : u) + pf_row * kOut;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((unsigned)(rows * kOut * 4)) : "memory");
            }
        }
#endif
        // ================= producer + MMA issue for tile i =========================================================
        if (i < n_my) {
#pragma unroll
            for (int c = 0; c < kStages; c++) {
                // ring slot of this K-chunk; its previous user (chunk g - 3) must have been consumed by the tensor core
                const int g = i * kStages + c, slot = g % kSlots, use = g / kSlots;
                if (use >= 1) mbar_wait(bar0 + 8u * slot, (unsigned)(use - 1) & 1u);
                {
                    uint4 c0, c1, c2;
                    split_chunk(pre[2 * c], pre[2 * c + 1], c0, c1, c2);
                    unsigned char *dst = a_st + slot * kAStageBytes + a_off;
                    *reinterpret_cast<uint4 *>(dst) = c0;
                    *reinterpret_cast<uint4 *>(dst + kATermBytes) = c1;
                    *reinterpret_cast<uint4 *>(dst + 2 * kATermBytes) = c2;
                }
#if ULTRA_TC3_PREFETCH == 0
                fetch_chunk(tile + step, c, i + 1 < n_my);  // next tile's rows of this chunk: in flight for a whole tile
#endif
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();
#if ULTRA_TC3_PREFETCH == 1
                fetch_chunk(tile + step, c, i + 1 < n_my);  // after the membar + barrier, so neither waits for these loads
#elif ULTRA_TC3_PREFETCH == 2
                if (c == kStages - 1) {  // all of the next tile's rows at once: the epilogue below covers their latency
#pragma unroll
                    for (int cc = 0; cc < kStages; cc++) fetch_chunk(tile + step, cc, i + 1 < n_my);
                }
#endif
                if (tid == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d_tmem = tmem_base + (uint32_t)(i & 1) * kAccStride;
                    const uint64_t da0 = da_base + (uint64_t)(slot * (kAStageBytes >> 4));
#pragma unroll
                    for (int ks = 0; ks < kStageK / 16; ks++) {  // one MMA = 16 bf16 of K = 2 chunks
                        const uint64_t ka = da0 + (uint64_t)(ks * (2 * kALbo >> 4));
                        const uint64_t kb = dw_base + (uint64_t)((c * (kStageK / 16) + ks) * 16);
                        mma_bf16(d_tmem, ka, kb, idesc(3 * kOut), (c | ks) ? 1u : 0u);             // A0 x [W0;W1;W2]
                        mma_bf16(d_tmem, ka + (uint64_t)(kATermBytes >> 4), kb, idesc(2 * kOut), 1u);      // A1 x [W0;W1]
                        mma_bf16(d_tmem, ka + (uint64_t)(2 * kATermBytes >> 4), kb, idesc(kOut), 1u);      // A2 x [W0]
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
            // part A: thread = (row r, 16 columns): accumulator thirds -> bias -> LayerNorm -> ReLU -> staging
            mbar_wait(bar0 + 8u * (kSlots + (e & 1)), (unsigned)(e >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float v[16], v2[16];
            const uint32_t t_acc = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)((e & 1) * kAccStride + cb);
            tmem_ld16(t_acc + 2 * kOut, v);   // a0.w2            (smallest terms first)
            tmem_ld16(t_acc + kOut, v2);      // a0.w1 + a1.w1
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 16; c++) v[c] += v2[c];
            tmem_ld16(t_acc, v2);             // a0.w0 + a1.w0 + a2.w0
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // ordered before the next barrier
#pragma unroll
            for (int c = 0; c < 16; c++) v[c] += v2[c];
            const float4 *prm4 = reinterpret_cast<const float4 *>(prm);
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
                const float4 b = prm4[(cb + c) >> 2];
                v[c] += b.x; v[c + 1] += b.y; v[c + 2] += b.z; v[c + 3] += b.w;
                s += (v[c] + v[c + 1]) + (v[c + 2] + v[c + 3]);
            }
            if (z_out != nullptr && m0 + r < M) {  // training: keep z for the backward pass (64 contiguous bytes per thread)
#pragma unroll
                for (int c = 0; c < 16; c += 4)
                    *reinterpret_cast<float4 *>(z_out + (m0 + r) * kOut + cb + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
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
                if (stats_out != nullptr && grp == 0 && m0 + r < M)
                    *reinterpret_cast<float2 *>(stats_out + 2 * (m0 + r)) = make_float2(mean, rstd);
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
#pragma unroll
            for (int t = 0; t < kIt; t++) {
                const int idx = tid + t * kThreads, rr = idx >> 4, c4 = idx & 15;
                float4 o = *reinterpret_cast<const float4 *>(ostage + rr * 256 + ((c4 ^ (rr & 15)) << 4));
                o.x += xr[t].x; o.y += xr[t].y; o.z += xr[t].z; o.w += xr[t].w;
                if (m0 + rr < M) __stcs(reinterpret_cast<float4 *>(out + (m0 + rr) * kOut) + c4, o);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

}  // namespace tc3

}  // namespace ultra

namespace ultra {
namespace {

int launch_update(const float *x, const float *u, const float *weight_t, const float *bias, const float *ln_gamma,
                  const float *ln_beta, float *out, float *z_out, float *stats_out, int64_t num_rows, int64_t in_dim,
                  int64_t out_dim, float ln_eps, int layer_norm, int relu, int shortcut, int sm_limit, void *stream_) {
    using namespace ultra::tc3;
    if (in_dim != kK || out_dim != kOut) {
        set_error("nbfnet_update: only Linear(128 -> 64) (sum / mean / max aggregation, dim 64) is fused");
        return 1;
    }
    if (num_rows < 0) { set_error("nbfnet_update: negative num_rows"); return 1; }
    if (num_rows == 0) return 0;
    for (const void *p : {(const void *)x, (const void *)u, (const void *)weight_t, (const void *)out})
        if (p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) != 0) {
            set_error("nbfnet_update: x, u, weight_t, out must be non-null and 16-byte aligned");
            return 1;
        }
    if ((z_out == nullptr) != (stats_out == nullptr) || (z_out != nullptr && !layer_norm) ||
        (reinterpret_cast<uintptr_t>(z_out) & 15) != 0 || (reinterpret_cast<uintptr_t>(stats_out) & 7) != 0) {
        set_error("nbfnet_update (training): z_out and stats_out come together, aligned, and need layer_norm");
        return 1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    static thread_local int attr_dev = -1;
    int dev = 0, sms = 148;
    cudaError_t err = cudaGetDevice(&dev);
    if (err == cudaSuccess && attr_dev != dev) {
        err = cudaFuncSetAttribute(update_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        attr_dev = dev;
    }
    if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) { set_error(std::string("nbfnet_update: ") + cudaGetErrorString(err)); return 2; }
    if (sm_limit > 0 && sm_limit < sms) sms = sm_limit;  // leave SMs to a concurrent collective (per call, no global)
    const long long tiles = (num_rows + kBM - 1) / kBM;
    const int grid = (int)(tiles < sms ? tiles : sms);
    update_tc3_kernel<<<grid, kThreads, kSmemBytes, stream>>>(x, u, weight_t, bias, ln_gamma, ln_beta, out,
                                                              (long long)num_rows, ln_eps, layer_norm, relu, shortcut,
                                                              z_out, stats_out);
    err = cudaGetLastError();
    if (err != cudaSuccess) { set_error(std::string("nbfnet_update launch: ") + cudaGetErrorString(err)); return 2; }
    return 0;
}

// ---- backward of  y = relu( LayerNorm(z) * gamma + beta )  w.r.t. z, gamma, beta (and bias = sum of dz) -----------------
// One pass: read dout and z, write dz; 16 lanes own one 64-float row (a float4 each), a warp two rows per step; the row
// means go through 4 shuffle steps; the parameter gradients are accumulated per lane over its rows, reduced per CTA in
// shared memory in a fixed order and written as per-CTA partials (the caller sums `grid` rows: deterministic).
constexpr int kBwdThreads = 256;

__global__ void __launch_bounds__(kBwdThreads)
update_bwd_kernel(const float *__restrict__ dout, const float *__restrict__ z, const float *__restrict__ stats,
                  const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ dz,
                  float *__restrict__ partials, long long M, int relu) {
    __shared__ float4 red[3][kBwdThreads / 16][16];
    const int tid = threadIdx.x, l16 = tid & 15;
    const long long row0 = ((long long)blockIdx.x * kBwdThreads + tid) >> 4, stride = ((long long)gridDim.x * kBwdThreads) >> 4;
    const float4 ga = gamma ? __ldg(reinterpret_cast<const float4 *>(gamma) + l16) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 be = beta ? __ldg(reinterpret_cast<const float4 *>(beta) + l16) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg, dbias = dg;
    // every lane of a half-warp runs the same number of iterations (rows are per half-warp), so the shuffles are safe
    const long long iters = (M + stride - 1) / stride;
    for (long long it = 0; it < iters; it++) {
        const long long row = row0 + it * stride;
        const bool live = row < M;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f), zz = g;
        float2 st = make_float2(0.f, 1.f);
        if (live) {
            g = __ldcs(reinterpret_cast<const float4 *>(dout + row * 64) + l16);
            zz = __ldcs(reinterpret_cast<const float4 *>(z + row * 64) + l16);
            st = __ldg(reinterpret_cast<const float2 *>(stats) + row);
        }
        float zh[4] = {(zz.x - st.x) * st.y, (zz.y - st.x) * st.y, (zz.z - st.x) * st.y, (zz.w - st.x) * st.y};
        const float gam[4] = {ga.x, ga.y, ga.z, ga.w}, bet[4] = {be.x, be.y, be.z, be.w};
        float dy[4] = {g.x, g.y, g.z, g.w}, dzh[4];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float y = zh[k] * gam[k] + bet[k];
            if (relu && !(y > 0.f)) dy[k] = 0.f;  // relu'(y): 1 for y > 0 (torch convention: 0 at y == 0)
            dzh[k] = dy[k] * gam[k];
            s1 += dzh[k];
            s2 += dzh[k] * zh[k];
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        s1 *= (1.f / 64.f);
        s2 *= (1.f / 64.f);
        float o4[4];
#pragma unroll
        for (int k = 0; k < 4; k++) o4[k] = st.y * (dzh[k] - s1 - zh[k] * s2);
        if (live) {
            __stcs(reinterpret_cast<float4 *>(dz + row * 64) + l16, make_float4(o4[0], o4[1], o4[2], o4[3]));
            dg.x += dy[0] * zh[0]; dg.y += dy[1] * zh[1]; dg.z += dy[2] * zh[2]; dg.w += dy[3] * zh[3];
            db.x += dy[0]; db.y += dy[1]; db.z += dy[2]; db.w += dy[3];
            dbias.x += o4[0]; dbias.y += o4[1]; dbias.z += o4[2]; dbias.w += o4[3];
        }
    }
    red[0][tid >> 4][l16] = dg;
    red[1][tid >> 4][l16] = db;
    red[2][tid >> 4][l16] = dbias;
    __syncthreads();
    if (tid < 48) {  // 3 quantities x 16 column groups: fold the 16 row-slots in order
        const int qn = tid >> 4, cg = tid & 15;
        float4 a = red[qn][0][cg];
#pragma unroll
        for (int r = 1; r < kBwdThreads / 16; r++) {
            const float4 b = red[qn][r][cg];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        *reinterpret_cast<float4 *>(partials + ((size_t)blockIdx.x * 3 + qn) * 64 + cg * 4) = a;
    }
}

}  // namespace
}  // namespace ultra

extern "C" int ultra_nbfnet_update_f32(const float *x, const float *u, const float *weight_t, const float *bias,
                                       const float *ln_gamma, const float *ln_beta, float *out, int64_t num_rows,
                                       int64_t in_dim, int64_t out_dim, float ln_eps, int layer_norm, int relu,
                                       int shortcut, int sm_limit, void *stream_) {
    return ultra::launch_update(x, u, weight_t, bias, ln_gamma, ln_beta, out, nullptr, nullptr, num_rows, in_dim, out_dim,
                                ln_eps, layer_norm, relu, shortcut, sm_limit, stream_);
}

extern "C" int ultra_nbfnet_update_train_f32(const float *x, const float *u, const float *weight_t, const float *bias,
                                             const float *ln_gamma, const float *ln_beta, float *out, float *z_out,
                                             float *stats_out, int64_t num_rows, int64_t in_dim, int64_t out_dim,
                                             float ln_eps, int relu, int shortcut, int sm_limit, void *stream_) {
    if (z_out == nullptr || stats_out == nullptr) {
        ultra::set_error("nbfnet_update_train: z_out and stats_out are required");
        return 1;
    }
    return ultra::launch_update(x, u, weight_t, bias, ln_gamma, ln_beta, out, z_out, stats_out, num_rows, in_dim, out_dim,
                                ln_eps, 1, relu, shortcut, sm_limit, stream_);
}

extern "C" int ultra_nbfnet_update_bwd_grid(int64_t num_rows) {
    const long long want = (num_rows + 15) / 16 / (ultra::kBwdThreads / 16 * 4);  // >= 4 row-steps per thread
    return (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
}

extern "C" int ultra_nbfnet_update_bwd_f32(const float *out_grad, const float *z, const float *stats,
                                           const float *ln_gamma, const float *ln_beta, float *z_grad, float *partials,
                                           int64_t num_rows, int64_t dim, int relu, void *stream_) {
    using namespace ultra;
    if (dim != 64) { set_error("nbfnet_update_bwd: only dim == 64 is implemented"); return 1; }
    if (num_rows < 0) { set_error("nbfnet_update_bwd: negative num_rows"); return 1; }
    for (const void *p : {(const void *)out_grad, (const void *)z, (const void *)stats, (const void *)z_grad,
                          (const void *)partials})
        if (p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) != 0) {
            set_error("nbfnet_update_bwd: pointers must be non-null and 16-byte aligned");
            return 1;
        }
    const int grid = ultra_nbfnet_update_bwd_grid(num_rows);
    update_bwd_kernel<<<grid, kBwdThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
        out_grad, z, stats, ln_gamma, ln_beta, z_grad, partials, (long long)num_rows, relu);
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error(std::string("nbfnet_update_bwd launch: ") + cudaGetErrorString(err)); return 2; }
    return 0;
}
