// tc_issue_bench.cu -- how fast can one warp issue small tcgen05.mma (M=128, N=48, K=8, tf32)?  Measures cycles from the
// first issue to the completion barrier for NMMA back-to-back MMAs, with 1..4 issuing warps (separate accumulators).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() { uint32_t ok; asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok)); return ok != 0; }

template <int NMMA, int NISS, int N>
__global__ void bench(long long* out)
{
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16384; i += blockDim.x) sm[i] = 0.001f * (i & 63);
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(NISS)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo0 = ((smem_u32(sm) >> 4) & 0x3fffu) | ((2112u >> 4) << 16);
    const uint32_t b_lo0 = ((smem_u32(sm + 8192) >> 4) & 0x3fffu) | (((uint32_t)N * 16u >> 4) << 16);
    long long t0 = clock64();
    if (warp < NISS) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i) {
            const uint64_t da = ((uint64_t)hi << 32) | (a_lo0 + (uint32_t)(i % 6) * 264u);
            const uint64_t db = ((uint64_t)hi << 32) | (b_lo0 + (uint32_t)(i % 8) * (uint32_t)(N * 2));
            asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem + (uint32_t)(warp * N)), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(i > 0)) : "memory");
        }
        long long t1 = clock64();
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        if (tid == 0) out[1] = t1 - t0;
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    long long t2 = clock64();
    if (tid == 0) out[0] = t2 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
template <int NMMA, int NISS, int N> int run(const char* nm) {
    long long* d; CK(cudaMalloc(&d, 16)); long long h[2];
    CK(cudaFuncSetAttribute(bench<NMMA, NISS, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int r = 0; r < 3; ++r) { bench<NMMA, NISS, N><<<1, 128, 65536>>>(d); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("ISSUE %-24s NMMA=%2d issuers=%d N=%3d : total %6lld cyc (%.1f / MMA per issuer), issue loop %6lld cyc (%.1f / MMA)\n", nm, NMMA, NISS, N, h[0], (double)h[0] / NMMA, h[1], (double)h[1] / NMMA);
    cudaFree(d); return 0;
}
int main() {
    run<1, 1, 48>("single"); run<6, 1, 48>("chain of 6"); run<18, 1, 48>("chain of 18"); run<36, 1, 48>("chain of 36");
    run<6, 3, 48>("3 issuers x 6"); run<6, 4, 48>("4 issuers x 6"); run<18, 2, 48>("2 issuers x 18");
    run<18, 1, 128>("chain of 18, N=128"); run<18, 1, 16>("chain of 18, N=16"); run<5, 4, 112>("4 issuers x 5, N=112");
    return 0;
}
