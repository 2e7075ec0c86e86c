// tc_probe.cu -- bring-up probe for the tcgen05 (UMMA) path: validates on a real B200 the shared-memory
// descriptor encoding this repo relies on (K-major, no swizzle, rows 16 B apart so that a 1-row shift is a
// 16-byte start-address offset), the instruction descriptor for kind::tf32, TMEM allocation and the
// tcgen05.ld 32x32b lane/column mapping.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe tc_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE ("interleave") shared-memory matrix descriptor.
// element (row r, k) lives at start + (r % 8) * 16 + (r / 8) * SBO + (k / T) * LBO + (k % T) * sizeof(elem),  T = 16 B / sizeof(elem)
__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;       // descriptor version (Blackwell)
    return d;                     // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
__host__ __device__ inline uint32_t make_idesc_tf32(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                 // D format f32
    d |= 2u << 7;                 // A format tf32
    d |= 2u << 10;                // B format tf32
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;                     // a_major = b_major = K
}

__global__ void probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int nk, int PS,
                      int row_shift, int a_floats, int b_floats)
{
    extern __shared__ __align__(128) float sm[];
    float* sA = sm;                          // [2*nk][PS][4]
    float* sB = sm + a_floats;               // [2*nk][N][4]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < a_floats; i += blockDim.x) sA[i] = A[i];
    for (int i = tid; i < b_floats; i += blockDim.x) sB[i] = B[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the MMA
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        for (int j = 0; j < nk; ++j) {
            const uint64_t da = make_desc(smem_u32(sA) + (uint32_t)(2 * j) * PS * 16 + row_shift * 16, PS * 16, 128);
            const uint64_t db = make_desc(smem_u32(sB) + (uint32_t)(2 * j) * N * 16, N * 16, 128);
            const uint32_t acc = j > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everybody waits for the MMAs
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w reads TMEM lanes 32w..32w+31; thread = one row, 16 columns at a time
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i)
            if (c0 + i < N) D[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

static float aval(int r, int k) { return (float)(((r * 3 + k * 5) % 17) - 8) / 16.0f; }
static float bval(int n, int k) { return (float)(((n * 7 + k * 3) % 13) - 6) / 8.0f; }

static int run_case(int N, int nk, int row_shift, const char* name)
{
    const int PS = 160, K = 8 * nk;
    const int a_floats = 2 * nk * PS * 4, b_floats = 2 * nk * N * 4;
    std::vector<float> A(a_floats), B(b_floats), D(128 * N, -1.f);
    // slot r of slab kc holds logical row (r) : A[r][4*kc .. 4*kc+3]
    for (int kc = 0; kc < 2 * nk; ++kc)
        for (int r = 0; r < PS; ++r)
            for (int e = 0; e < 4; ++e) A[(kc * PS + r) * 4 + e] = aval(r, 4 * kc + e);
    for (int kc = 0; kc < 2 * nk; ++kc)
        for (int n = 0; n < N; ++n)
            for (int e = 0; e < 4; ++e) B[(kc * N + n) * 4 + e] = bval(n, 4 * kc + e);
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, a_floats * 4)); CK(cudaMalloc(&dB, b_floats * 4)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, A.data(), a_floats * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), b_floats * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, 128 * N * 4));
    const int smem = (a_floats + b_floats) * 4;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe<<<1, 128, smem>>>(dA, dB, dD, N, nk, PS, row_shift, a_floats, b_floats);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)aval(m + row_shift, k) * bval(n, k);
            double e = fabs(ref - D[m * N + n]);
            if (e > maxerr) maxerr = e;
            if (e > 1e-5 && bad < 5) { printf("  mismatch m=%d n=%d got %f want %f\n", m, n, D[m * N + n], ref); ++bad; }
        }
    printf("PROBE %-28s N=%3d nk=%d shift=%d : max err %.3e %s\n", name, N, nk, row_shift, maxerr, maxerr < 1e-5 ? "OK" : "FAIL");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return maxerr < 1e-5 ? 0 : 1;
}

int main()
{
    int fails = 0;
    fails += run_case(48, 1, 0, "single MMA");
    fails += run_case(48, 1, 1, "A start +16 B (1 row)");
    fails += run_case(48, 1, 2, "A start +32 B (2 rows)");
    fails += run_case(48, 3, 0, "3 k-steps accumulate");
    fails += run_case(48, 6, 3, "6 k-steps, shift 3");
    fails += run_case(16, 2, 1, "N=16");
    fails += run_case(32, 2, 0, "N=32");
    fails += run_case(64, 2, 5, "N=64");
    fails += run_case(96, 2, 0, "N=96");
    fails += run_case(128, 2, 7, "N=128");
    fails += run_case(112, 2, 0, "N=112");
    printf("tc_probe: %d failing case(s)\n", fails);
    return fails ? 1 : 0;
}
