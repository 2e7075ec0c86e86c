// fe_inst.cuh -- device entry point of the fused kernel and the per-variant host glue.
// Included by the per-configuration translation units fe_inst_*.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "fe_kernel.cuh"
#include "fe_pack.h"
#include "fe_variant.h"

namespace fe {

// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// ---- 2-D TMA tiles (cp.async.bulk.tensor; SASS: UTMALDG / UTMASTG) ----
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int x, int y, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(src) : "memory");
}
// ---- tcgen05 (UMMA) primitives; encodings validated on hardware by tools/tc_probe.cu ----
// K-major, SWIZZLE_NONE shared-memory matrix descriptor: element (row r, k) at
//   start + (r % 8) * 16 + (r / 8) * SBO + (k / 4) * LBO + (k % 4) * 4   bytes (tf32: 4 elements per 16 B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor: D f32, A/B tf32 (format 2) or f16 (format 0), both K-major, M = m (128 or 64), N = n
__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// kind::f16: A / B format 0 = fp16, 1 = bfloat16
__device__ __forceinline__ uint32_t umma_idesc_f16(int m, int n, int bf = 0) {
    return (1u << 4) | ((uint32_t)bf << 7) | ((uint32_t)bf << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {   // arrives on `bar` once every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {      // one lane of a converged warp (always the same one)
    uint32_t ok;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok));
    return ok != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int NT> __device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

template <class P> struct GpuCtx {
    float* sm; const float* blob; KParams prm; int s0; float* gs; int cta, ncta;
    int tid; unsigned seq_base; uint32_t bars;     // bars: full[STAGES] then empty[STAGES], 8 bytes each
    int h0, h1;             // hop range of the piece of work in progress (one item of a hop-sliced launch; Plan::SLICED variants only)
    __device__ __forceinline__ int hbeg() const { return P::SLICED ? h0 : 0; }
    __device__ __forceinline__ int hend() const { return P::SLICED ? h1 : prm.n_hops; }
    __device__ __forceinline__ void begin_range(int a, int b) { if constexpr (P::SLICED) { h0 = a; h1 = b; } }
    // hop-sliced launches: item `idx` (an earlier hop range of the same streams, possibly run by another CTA) has stored its state
    __device__ __forceinline__ void wait_item(int idx) const {
        if (tid == 0) {
            const long long t0 = clock64();
            int v = 0;
            do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(prm.slice_flags + idx) : "memory");
                if (v == 0 && clock64() - t0 > 20000000000LL) __trap();          // bounded: a scheduling bug is an error, not a hang
            } while (v == 0);
        }
        bar_consumers<P::NT>();
    }
    __device__ __forceinline__ void signal_item(int idx) const {
        __threadfence();
        bar_consumers<P::NT>();
        if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(prm.slice_flags + idx), "r"(1) : "memory");
    }
    int ci0, nci;           // weight chunks of one iteration: [ci0, ci0 + nci) (the whole frame, or one stage of the frame-parallel schedule)
    // ---- hop tiles by TMA (HOP_RING variants, prm.hop_tma): hop_full mbarrier behind the accumulator barrier ----
    // the thread that issues the hop tiles: first lane of the LAST consumer warp (warp 0 issues the MMAs and is the critical one)
    static constexpr int HOP_TID = P::NT - 32;
    static __device__ __forceinline__ uint32_t hop_full_bar(uint32_t bars_) { return bars_ + 8u * (2 * P::STAGES) + 16u; }
    // the input tile of hop `hop` has landed in the input ring
    __device__ __forceinline__ void hop_wait(int hop) const { mbar_wait(hop_full_bar(bars), (unsigned)hop & 1u); }
    // (after the barrier that ends a window phase / the one-time init) the ring tile that hop `hop` overwrites has been read by every
    // thread: consumer thread 0 issues its 2-D TMA tiles right away, a whole frame before they are needed -- nothing ever waits for a
    // free tile, so no thread has to poll for one.  Generic-proxy reads before an async-proxy write: proxy fence first.
    __device__ __forceinline__ void hop_prefetch(int hop) const {
        if (tid == HOP_TID && hop < prm.n_hops) {
            constexpr int H = P::Cf::HOP, HT = P::HT, N = P::Cf::N_FFT;
            fence_async_smem();
            mbar_expect_tx(hop_full_bar(bars), (uint32_t)(P::S * H * 4));
#pragma unroll
            for (int k = 0; k < H / HT; ++k) {
                const int pos = (hop * H + k * HT) & (N - 1);
                tma_load_2d(smem_u32(sm + P::SM_TIN + (pos / HT) * P::S * HT), static_cast<const CUtensorMap*>(prm.tmaps), hop * H + k * HT, s0,
                            hop_full_bar(bars));
            }
        }
    }
    // (after the barrier that ends an overlap-add phase) the output hop leaves from the overlap-add ring, one 2-D tile [S][HT] per ring tile
    __device__ __forceinline__ void hop_store(int hop) const {
        if (tid == HOP_TID) {
            fence_async_smem();
#pragma unroll
            for (int k = 0; k < P::Cf::HOP / P::HT; ++k) {
                const int pos = (hop * P::Cf::HOP + k * P::HT) & (P::Cf::N_FFT - 1);
                tma_store_2d(static_cast<const CUtensorMap*>(prm.tmaps) + 1, hop * P::Cf::HOP + k * P::HT, s0,
                             smem_u32(sm + P::SM_OLA + (pos / P::HT) * P::S * P::HT));
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    // the stores of the previous hop have read their tiles (before the ring is written again) / have completed (end of the launch)
    __device__ __forceinline__ void hop_store_wait(bool all) const {
        if (tid == HOP_TID) {
            if (all) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    __device__ __forceinline__ const float* acquire(int ci, int) const {
        const unsigned seq = seq_base + (unsigned)(P::TC ? ci : ci - ci0), stage = seq % P::STAGES, par = (seq / P::STAGES) & 1u;
        mbar_wait(bars + 8u * stage, par);
        return sm + P::SM_RING + stage * P::CHUNK;
    }
    __device__ __forceinline__ void release(int ci) const {
        const unsigned seq = seq_base + (unsigned)(P::TC ? ci : ci - ci0), stage = seq % P::STAGES;
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(bars + 8u * (P::STAGES + stage));
    }
    // ---- tensor-core state (TC variants) ----
    uint32_t tmem;          // TMEM base address (lane 0, first allocated column)
    uint32_t acc_bar;       // mbarrier: accumulators of the current layer are complete
    unsigned acc_uses;
    __device__ __forceinline__ void mma_fence() const { tc_fence_after(); }
    // K-major / no-swizzle operand descriptor split in two words: lo = start address | LBO, hi = SBO (128 B) | version.
    // Advancing the start address is one 32-bit add (addresses stay below 256 KB, so no carry into the LBO field).
    struct Desc { uint32_t lo, hi; };
    __device__ __forceinline__ Desc make_desc(const float* p, int lbo_floats) const {
        Desc d;
        d.lo = ((smem_u32(p) >> 4) & 0x3fffu) | ((((uint32_t)lbo_floats * 4u) >> 4) << 16);
        d.hi = (128u >> 4) | (1u << 14);
        return d;
    }
    // MN-major / no-swizzle operand (the N or M index is the contiguous one: 8 halves per 16-byte row, 8 k rows per 128-byte core
    // matrix): LBO = 128 B between groups of 8 k rows, SBO = pitch between the 16-byte groups along N (validated by tools/tc_probe_mn.cu)
    __device__ __forceinline__ Desc make_desc_mn(const float* p, int sbo_floats) const {
        Desc d;
        d.lo = ((smem_u32(p) >> 4) & 0x3fffu) | ((128u >> 4) << 16);
        d.hi = (((uint32_t)sbo_floats * 4u) >> 4) | (1u << 14);
        return d;
    }
    __device__ __forceinline__ Desc desc_add(Desc d, int floats) const { d.lo += (uint32_t)(floats >> 2); return d; }
    __device__ __forceinline__ Desc desc_set_lbo(Desc d, int lbo_floats) const {
        d.lo = (d.lo & 0x0000ffffu) | ((((uint32_t)lbo_floats * 4u) >> 4) << 16);
        return d;
    }
    // FE_ELECT_ONCE (default): tc_stream elects one lane of warp 0 per weight chunk and that lane alone runs the unrolled MMA
    // sequence; 0 = every lane runs it and each MMA elects (13 instead of ~8 instructions per MMA in the issue loop).
#ifndef FE_ELECT_ONCE
#define FE_ELECT_ONCE 0      // measured on B200 (B, 256 streams): 41.9 vs 39.8 us/hop -- a lone diverged lane issues MMAs more slowly
#endif
#if FE_ELECT_ONCE
#define FE_MMA_ELECT ".reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
#define FE_MMA_PRED ""
    __device__ __forceinline__ bool elect(int) const { return elect_one(); }
#else
#define FE_MMA_ELECT ".reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
#define FE_MMA_PRED "@q "
    __device__ __forceinline__ bool elect(int) const { return true; }
#endif
    __device__ __forceinline__ void warp_sync() const { __syncwarp(); }
    // M64: a 64-row MMA, whose accumulator row r lives in TMEM lane 32 * (r / 16) + r % 16 (tools/tc_bench2.cu probes the mapping).
    // FMT: 0 = kind::tf32; 1 / 2 = kind::f16 with fp16 / bfloat16 operands (8 per 16-byte row: K = 16 per MMA), same descriptors.
    static constexpr int FMT16 = P::BF16 ? 2 : 1;      // the 16-bit operand format of this variant's conv section
    // BMN: the B operand is MN-major (instruction-descriptor bit 16)
    template <bool M64 = false, int FMT = 0, bool BMN = false>
    __device__ __forceinline__ void mma(int /*tid*/, Desc a, Desc b, int np, int col, bool acc, int /*rows*/) const {
        const uint64_t da = ((uint64_t)a.hi << 32) | a.lo, db = ((uint64_t)b.hi << 32) | b.lo;
        static_assert(!BMN || FMT != 0, "MN-major operands: 16-bit kinds only");
        if constexpr (FMT != 0)
            asm volatile(
                "{\n\t" FE_MMA_ELECT
                FE_MMA_PRED "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem + (uint32_t)col), "l"(da), "l"(db), "r"(umma_idesc_f16(M64 ? 64 : 128, np, FMT == 2) | (BMN ? (1u << 16) : 0u)), "r"(acc ? 1u : 0u) : "memory");
        else
            asm volatile(
                "{\n\t" FE_MMA_ELECT
                FE_MMA_PRED "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem + (uint32_t)col), "l"(da), "l"(db), "r"(umma_idesc_tf32(M64 ? 64 : 128, np)), "r"(acc ? 1u : 0u) : "memory");
    }
    // same, A operand from tensor memory: columns a_col .. a_col + 7 (one fp32 / TF32 element per column, lane = row)
    // (F16: columns a_col .. a_col + 7 hold 16 halves, two per column, the even channel in the low half)
    template <bool F16 = false>
    __device__ __forceinline__ void mma_ts(int /*tid*/, int a_col, Desc b, int np, int col, bool acc, int /*rows*/) const {
        const uint64_t db = ((uint64_t)b.hi << 32) | b.lo;
        if constexpr (F16)
            asm volatile(
                "{\n\t" FE_MMA_ELECT
                FE_MMA_PRED "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                ::"r"(tmem + (uint32_t)col), "r"(tmem + (uint32_t)a_col), "l"(db), "r"(umma_idesc_f16(128, np)), "r"(acc ? 1u : 0u) : "memory");
        else
            asm volatile(
                "{\n\t" FE_MMA_ELECT
                FE_MMA_PRED "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                ::"r"(tmem + (uint32_t)col), "r"(tmem + (uint32_t)a_col), "l"(db), "r"(umma_idesc_tf32(128, np)), "r"(acc ? 1u : 0u) : "memory");
    }
    // ring stage release in a tensor-core layer: thread 0's arrival is a tcgen05.commit (fires when its MMAs, which read
    // the stage, are done); the other warps never touch the stage and arrive at once
    __device__ __forceinline__ void release_mma(int ci) const {
        const unsigned seq = seq_base + (unsigned)(P::TC ? ci : ci - ci0), stage = seq % P::STAGES;
        __syncwarp();
        if (tid < 32) {
            if (elect_one()) umma_commit(bars + 8u * (P::STAGES + stage));      // the lane that issued this warp's MMAs
        } else if ((tid & 31) == 0) mbar_arrive(bars + 8u * (P::STAGES + stage));
    }
    // FE_ACC_PARK: only warp 0 polls the accumulator-ready mbarrier; the other consumer warps park on a hardware named barrier
    // (no issue slots, no shared-memory polling next to the MMA issue) that warp 0 joins once the MMAs are done.
#ifndef FE_ACC_PARK
#define FE_ACC_PARK 0        // measured on B200 (B, 256 streams): 41.4 vs 39.9 us/hop -- the named barrier costs more than the polling
#endif
    __device__ __forceinline__ void acc_commit_wait() {
        if (tid < 32) { __syncwarp(); if (elect_one()) umma_commit(acc_bar); }
#if FE_ACC_PARK
        if (tid < 32) { mbar_wait(acc_bar, acc_uses & 1u); tc_fence_after(); tc_fence_before(); }   // observe, then publish across the barrier
        asm volatile("bar.sync 2, %0;" ::"n"(P::NT) : "memory");
#else
        mbar_wait(acc_bar, acc_uses & 1u);
#endif
        ++acc_uses;
        tc_fence_after();
    }
    __device__ __forceinline__ void tmem_ld4(int /*tid*/, int col, float* v) const {
        const uint32_t taddr = tmem + ((uint32_t)(((tid >> 5) & 3) << 5) << 16) + (uint32_t)col;   // this warp's lane quadrant
        uint32_t r0, r1, r2, r3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
        v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    }
    // 16 lanes x 256 bits: thread t of the warp gets (lane t/4, columns col + 2(t%4) + {0,1}) in v[0..1] and (lane t/4 + 8, same
    // columns) in v[2..3], lanes counted from the start of the warp's quadrant
    __device__ __forceinline__ void tmem_ld16(int /*tid*/, int col, float* v) const {
        const uint32_t taddr = tmem + ((uint32_t)(((tid >> 5) & 3) << 5) << 16) + (uint32_t)col;
        uint32_t r0, r1, r2, r3;
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
        v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    }
    // this thread's row (TMEM lane), 4 consecutive columns
    __device__ __forceinline__ void tmem_st4(int /*tid*/, int col, const float* v) const {
        const uint32_t taddr = tmem + ((uint32_t)(((tid >> 5) & 3) << 5) << 16) + (uint32_t)col;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                     ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])) : "memory");
    }
    __device__ __forceinline__ void tmem_st2(int /*tid*/, int col, const float* v) const {       // 2 consecutive columns
        const uint32_t taddr = tmem + ((uint32_t)(((tid >> 5) & 3) << 5) << 16) + (uint32_t)col;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};"
                     ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])) : "memory");
    }
    __device__ __forceinline__ void tmem_st2_row(int /*row*/, int col, const float* v) const { tmem_st2(0, col, v); }
    // same for callers that know their row instead of their thread id (the row must be the calling thread's own lane)
    __device__ __forceinline__ void tmem_st4_row(int /*row*/, int col, const float* v) const { tmem_st4(0, col, v); }
    // 16-byte asynchronous global -> shared copies (cp.async, L2 only: the source may have been written by another CTA of this launch),
    // one group per prefetch; async_wait_all before the phase barrier makes this thread's pieces visible to the CTA
    __device__ __forceinline__ void async_copy16(float* dst, const float* src) const {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    }
    __device__ __forceinline__ void async_commit() const { asm volatile("cp.async.commit_group;" ::: "memory"); }
    __device__ __forceinline__ void async_wait_all() const { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
    __device__ __forceinline__ void tmem_st_wait() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
    __device__ __forceinline__ void tmem_ld_wait() const { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
    // end of a phase: make generic-proxy shared-memory writes visible to the async proxy (MMA operand reads) and order
    // TMEM accesses across the barrier
    __device__ __forceinline__ void phase_sync() const {
        if constexpr (P::TC) { fence_async_smem(); tc_fence_before(); }
        bar_consumers<P::NT>();
        if constexpr (P::TC) tc_fence_after();
    }

    long long t_last, t_sub;
    int cur_phase;          // phase id being executed (profile only): sub-timers are also filed per phase
    __device__ __forceinline__ void sub_begin(int) { if (prm.prof != nullptr && cta == 0 && tid == 0) t_sub = clock64(); }
    __device__ __forceinline__ void sub_end(int, int id) {
        if (prm.prof != nullptr && cta == 0 && tid == 0) {
            const long long t = clock64();
            prm.prof[id] += t - t_sub;
            prm.prof[PH_COUNT + cur_phase * PH_NSUB + (id - PH_TC_WAITW)] += t - t_sub;
            t_sub = t;
        }
    }
    __device__ __forceinline__ void stamp(int id) {
        if (prm.prof != nullptr && cta == 0 && tid == 0) {
            const long long t = clock64();
            prm.prof[id] += t - t_last;
            t_last = t;
        }
    }
    template <class F> __device__ __forceinline__ void phase(int id, F&& f) { cur_phase = id; f(tid); phase_sync(); stamp(id); }
    template <class A, class F1, class F2> __device__ __forceinline__ void phase2(int id, F1&& f1, F2&& f2) {
        A a;
        cur_phase = id;
        f1(tid, a);
        phase_sync();
        f2(tid, a);
        phase_sync();
        stamp(id);
    }
    __device__ __forceinline__ void next_frame() { seq_base += (unsigned)(P::TC ? P::NCHUNK_FRAME : nci); }       // (staged schedule: fp32 family only)
    __device__ __forceinline__ void check_frame(int) const {}
};

template <class P>
__global__ void __launch_bounds__(P::NTHREADS, 1) fe_fused_kernel(const KParams prm)
{
    extern __shared__ __align__(128) float sm[];
    const uint32_t bars = smem_u32(sm + P::SM_BAR);
    if (threadIdx.x == 0) {
        for (int i = 0; i < P::STAGES; ++i) {
            mbar_init(bars + 8u * i, 1);                        // full: producer's expect_tx arrival
            mbar_init(bars + 8u * (P::STAGES + i), P::NW);      // empty: one arrival per consumer warp
        }
        if constexpr (P::TC) mbar_init(bars + 8u * (2 * P::STAGES), 1);   // accumulator-ready barrier
        mbar_init(GpuCtx<P>::hop_full_bar(bars), 1);                     // hop tile landed (expect_tx arrival of the issuing thread + bytes)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + P::SM_BAR + 4 * P::STAGES + 2);
    if constexpr (P::TC) {
        if (threadIdx.x < 32) {          // warp 0 owns the TMEM allocation (TMEMC columns x 128 lanes of fp32 accumulators)
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(P::TMEMC) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
    }
    __syncthreads();
    if constexpr (P::TC) tc_fence_after();
    if (threadIdx.x >= P::NT) {
        // ---------------- producer warp: stream the weights of every frame through the ring ----------------
        if (threadIdx.x == P::NT && prm.mode <= MODE_OFFLINE) {      // the STFT-only modes consume no weights
            constexpr auto A = P::make_aux();
            const int* table = reinterpret_cast<const int*>(prm.blob + A.table);
            const uint32_t ring = smem_u32(sm + P::SM_RING);
            unsigned seq = 0;
            int iters = prm.n_hops, c0 = 0, c1 = P::NCHUNK_FRAME;
            if (!P::TC && prm.tp_stage != 0) {       // frame-parallel offline schedule: one iteration per frame group of this CTA, one stage's chunks
                const long nf = (long)prm.n_streams * prm.n_hops;
                const int ngroups = (int)((nf + P::S - 1) / P::S);
                iters = ((int)blockIdx.x < ngroups) ? (ngroups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
                c0 = Frame<P>::tp_ci0(prm); c1 = Frame<P>::tp_ci1(prm);
            }
            if (P::SLICED && prm.slice_hops > 0 && prm.mode == MODE_STREAM) {      // hop-sliced launch: the hops of this CTA's items (Frame::run)
                const int ngrp = (prm.n_streams + P::S - 1) / P::S, nrange = (prm.n_hops + prm.slice_hops - 1) / prm.slice_hops;
                iters = 0;
                for (int item = blockIdx.x; item < ngrp * nrange; item += gridDim.x) {
                    const int h0 = (item / ngrp) * prm.slice_hops;
                    iters += (h0 + prm.slice_hops < prm.n_hops ? h0 + prm.slice_hops : prm.n_hops) - h0;
                }
            }
            for (int hop = 0; hop < iters; ++hop) {
                for (int ci = c0; ci < c1; ++ci, ++seq) {
                    const unsigned stage = seq % P::STAGES, use = seq / P::STAGES;
                    const int off = __ldg(table + 2 * ci), nfl = __ldg(table + 2 * ci + 1);
                    if (use > 0) mbar_wait(bars + 8u * (P::STAGES + stage), (use - 1) & 1u);
                    mbar_expect_tx(bars + 8u * stage, (uint32_t)nfl * 4u);
                    bulk_g2s(ring + stage * (P::CHUNK * 4u), prm.blob + off, (uint32_t)nfl * 4u, bars + 8u * stage);
                }
            }
        }
        return;
    }
    GpuCtx<P> x;
    x.sm = sm; x.blob = prm.blob; x.prm = prm; x.cta = blockIdx.x; x.ncta = gridDim.x; x.s0 = blockIdx.x * P::S;
    x.ci0 = 0; x.nci = P::NCHUNK_FRAME;
    if constexpr (!P::TC) { x.ci0 = Frame<P>::tp_ci0(prm); x.nci = Frame<P>::tp_ci1(prm) - x.ci0; }
    x.gs = prm.scratch + (size_t)blockIdx.x * P::GS_TOTAL;
    x.tid = threadIdx.x; x.seq_base = 0; x.bars = bars; x.t_last = clock64(); x.cur_phase = 0;
    x.h0 = 0; x.h1 = prm.n_hops;
    x.tmem = 0; x.acc_bar = bars + 8u * (2 * P::STAGES); x.acc_uses = 0;
    if constexpr (P::TC) x.tmem = *tmem_slot;
    Frame<P>::run(x);
    if constexpr (P::TC) {
        tc_fence_before();
        bar_consumers<P::NT>();
        if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(x.tmem), "n"(P::TMEMC) : "memory");
    }
}

template <class P> struct VariantImpl {
    static void pack(const float* canonical, std::vector<float>& blob) { Packer<P>(blob).pack(canonical); }
    static cudaError_t prepare() {
        return cudaFuncSetAttribute(fe_fused_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM_BYTES);
    }
    static cudaError_t launch(const KParams& prm, int grid, cudaStream_t stream) {
        fe_fused_kernel<P><<<grid, P::NTHREADS, P::SMEM_BYTES, stream>>>(prm);
        return cudaGetLastError();
    }
    static VariantOps ops(int cfg_id) {
        using C = typename P::Cf;
        VariantOps v{};
        v.cfg_id = cfg_id; v.S = P::S; v.tc = P::PREC;
        v.shape = ShapeKey{C::N_FFT, C::HOP, C::C1, C::E, C::C2, C::F2, C::K, C::NH};
        v.smem_bytes = P::SMEM_BYTES; v.nthreads = P::NTHREADS; v.gs_floats = P::GS_TOTAL; v.state_floats = C::STATE;
        v.tap_floats = Frame<P>::TAP_TOTAL; v.nchunk_frame = P::NCHUNK_FRAME; v.blob_floats = P::make_aux().total;
        v.hop_ring = P::HOP_RING; v.hop_tile = P::HT;
        v.tp_group = P::TP_GROUP; v.aux_window_sq = P::make_aux().window_sq;
        v.pack = &pack; v.prepare = &prepare; v.launch = &launch;
        return v;
    }
};

}  // namespace fe

#define FE_VARIANT_ENTRY(id, CFG, SV, TCV) fe::VariantImpl<fe::Plan<fe::CFG, SV, TCV>>::ops(id),
#define FE_DEFINE_VARIANTS(fn, LIST)                                    \
    namespace fe {                                                      \
    const VariantOps* fn(int* n) {                                      \
        static const VariantOps v[] = {LIST(FE_VARIANT_ENTRY)};         \
        *n = (int)(sizeof(v) / sizeof(v[0]));                           \
        return v;                                                       \
    }                                                                   \
    }
