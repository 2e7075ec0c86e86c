// fe_inst.cuh -- device entry point of the fused kernel and the per-variant host glue.
// Included by the per-configuration translation units fe_inst_*.cu.
#pragma once
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
template <int NT> __device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

template <class P> struct GpuCtx {
    float* sm; const float* blob; KParams prm; int s0; float* gs; int cta;
    int tid; unsigned seq_base; uint32_t bars;     // bars: full[STAGES] then empty[STAGES], 8 bytes each
    __device__ __forceinline__ const float* acquire(int ci, int) const {
        const unsigned seq = seq_base + (unsigned)ci, stage = seq % P::STAGES, par = (seq / P::STAGES) & 1u;
        mbar_wait(bars + 8u * stage, par);
        return sm + P::SM_RING + stage * P::CHUNK;
    }
    __device__ __forceinline__ void release(int ci) const {
        const unsigned seq = seq_base + (unsigned)ci, stage = seq % P::STAGES;
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(bars + 8u * (P::STAGES + stage));
    }
    long long t_last;
    __device__ __forceinline__ void stamp(int id) {
        if (prm.prof != nullptr && cta == 0 && tid == 0) {
            const long long t = clock64();
            prm.prof[id] += t - t_last;
            t_last = t;
        }
    }
    template <class F> __device__ __forceinline__ void phase(int id, F&& f) { f(tid); bar_consumers<P::NT>(); stamp(id); }
    template <class A, class F1, class F2> __device__ __forceinline__ void phase2(int id, F1&& f1, F2&& f2) {
        A a;
        f1(tid, a);
        bar_consumers<P::NT>();
        f2(tid, a);
        bar_consumers<P::NT>();
        stamp(id);
    }
    __device__ __forceinline__ void next_frame() { seq_base += P::NCHUNK_FRAME; }
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
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= P::NT) {
        // ---------------- producer warp: stream the weights of every frame through the ring ----------------
        if (threadIdx.x == P::NT) {
            constexpr auto A = P::make_aux();
            const int* table = reinterpret_cast<const int*>(prm.blob + A.table);
            const uint32_t ring = smem_u32(sm + P::SM_RING);
            unsigned seq = 0;
            for (int hop = 0; hop < prm.n_hops; ++hop) {
                for (int ci = 0; ci < P::NCHUNK_FRAME; ++ci, ++seq) {
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
    x.sm = sm; x.blob = prm.blob; x.prm = prm; x.cta = blockIdx.x; x.s0 = blockIdx.x * P::S;
    x.gs = prm.scratch + (size_t)blockIdx.x * P::GS_TOTAL;
    x.tid = threadIdx.x; x.seq_base = 0; x.bars = bars; x.t_last = clock64();
    Frame<P>::run(x);
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
        v.cfg_id = cfg_id; v.S = P::S;
        v.shape = ShapeKey{C::N_FFT, C::HOP, C::C1, C::E, C::C2, C::F2, C::K, C::NH};
        v.smem_bytes = P::SMEM_BYTES; v.nthreads = P::NTHREADS; v.gs_floats = P::GS_TOTAL; v.state_floats = C::STATE;
        v.tap_floats = Frame<P>::TAP_TOTAL; v.nchunk_frame = P::NCHUNK_FRAME; v.blob_floats = P::make_aux().total;
        v.pack = &pack; v.prepare = &prepare; v.launch = &launch;
        return v;
    }
};

}  // namespace fe

#define FE_VARIANT_ENTRY(id, CFG, SV) fe::VariantImpl<fe::Plan<fe::CFG, SV>>::ops(id),
#define FE_DEFINE_VARIANTS(fn, LIST)                                    \
    namespace fe {                                                      \
    const VariantOps* fn(int* n) {                                      \
        static const VariantOps v[] = {LIST(FE_VARIANT_ENTRY)};         \
        *n = (int)(sizeof(v) / sizeof(v[0]));                           \
        return v;                                                       \
    }                                                                   \
    }
