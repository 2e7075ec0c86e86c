// tc_probe_mn.cu -- bring-up probe: tcgen05.mma with an MN-major ("transposed"), un-swizzled B operand.
// The conv-section activation buffers of the fused kernel are [channels / 8][slot][8 halves]: for a contraction over SLOTS
// (the frequency-axis linears rf_pre.0 / rf_post.0) they are a B operand whose N index (channels) is the contiguous one.
// Canonical MN-major / no-swizzle layout (in 16-byte units): element (n, k) at  (n / T) * SBO + (k % 8) + (k / 8) * LBO,  T = 8 halves /
// 4 tf32 per unit.  The probe checks that reading -- and the swapped one -- against a host product, for kind::f16 and kind::tf32.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe_mn tc_probe_mn.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) |
           ((uint64_t)1 << 46);
}
// kind::f16 (fmt 0) / kind::tf32 (fmt 2); b_major = 1: B is MN-major
__host__ __device__ inline uint32_t make_idesc(int fmt, int M, int N, int b_mn) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// a_bytes of A (K-major, rows 16 B apart, k-chunks lbo_a apart), b_bytes of B; one MMA: M = 128, N, K = 32 bytes of elements
__global__ void probe(const uint8_t* A, const uint8_t* B, float* D, int a_bytes, int b_bytes, int N, int tf32, uint32_t lbo_a,
                      uint32_t lbo_b, uint32_t sbo_b)
{
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t* sA = sm;
    uint8_t* sB = sm + a_bytes;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < a_bytes; i += blockDim.x) sA[i] = A[i];
    for (int i = tid; i < b_bytes; i += blockDim.x) sB[i] = B[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint64_t da = make_desc(smem_u32(sA), lbo_a, 128), db = make_desc(smem_u32(sB), lbo_b, sbo_b);
        const uint32_t idesc = make_idesc(tf32 ? 2 : 0, 128, N, 1);
        if (tf32)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t ok = 0;
        long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
            if (clock64() - t0 > 2000000000LL) __trap();
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        const int row = warp * 32 + (tid & 31);
        for (int c = 0; c < N; c += 4) {
            uint32_t r0, r1, r2, r3;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                         : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            D[row * N + c] = __uint_as_float(r0); D[row * N + c + 1] = __uint_as_float(r1);
            D[row * N + c + 2] = __uint_as_float(r2); D[row * N + c + 3] = __uint_as_float(r3);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }

int run(int tf32, int swapped)
{
    const int M = 128, N = 48, ES = tf32 ? 4 : 2, T = 16 / ES, K = 32 / ES;          // T elements per 16-byte unit, K per MMA
    const int NG = N / T, SLOTS = 24;                                                 // B: [NG n-groups][SLOTS k rows][T]: only K of the SLOTS rows are read
    std::vector<float> a(M * K), b(K * N);
    for (auto& v : a) v = (float)((rand() % 17) - 8) / 8.f;
    for (auto& v : b) v = (float)((rand() % 13) - 6) / 4.f;
    // A K-major: chunk kc (T elements) of row r at kc * (M * 16) + r * 16
    std::vector<uint8_t> ha((K / T) * M * 16), hb(NG * SLOTS * 16, 0);
    for (int r = 0; r < M; ++r)
        for (int k = 0; k < K; ++k) {
            uint8_t* p = &ha[(k / T) * M * 16 + r * 16 + (k % T) * ES];
            if (tf32) { float v = a[r * K + k]; memcpy(p, &v, 4); } else { __half h = __float2half(a[r * K + k]); memcpy(p, &h, 2); }
        }
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            uint8_t* p = &hb[(n / T) * SLOTS * 16 + k * 16 + (n % T) * ES];
            if (tf32) { float v = b[k * N + n]; memcpy(p, &v, 4); } else { __half h = __float2half(b[k * N + n]); memcpy(p, &h, 2); }
        }
    uint8_t *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, ha.size())); CK(cudaMalloc(&dB, hb.size())); CK(cudaMalloc(&dD, M * N * 4));
    CK(cudaMemcpy(dA, ha.data(), ha.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hb.data(), hb.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, M * N * 4));
    const uint32_t kgroup = 128, ngroup = SLOTS * 16;      // stride between groups of 8 k rows / between n groups (bytes)
    const uint32_t lbo_b = swapped ? ngroup : kgroup, sbo_b = swapped ? kgroup : ngroup;
    probe<<<1, 128, ha.size() + hb.size()>>>(dA, dB, dD, (int)ha.size(), (int)hb.size(), N, tf32, M * 16, lbo_b, sbo_b);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  %s %s: launch failed: %s\n", tf32 ? "tf32" : "f16 ", swapped ? "LBO=n-group,SBO=k-group" : "LBO=k-group,SBO=n-group", cudaGetErrorString(e)); return -1; }
    std::vector<float> d(M * N);
    CK(cudaMemcpy(d.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (int r = 0; r < M; ++r)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)(tf32 ? tf32r(a[r * K + k]) : a[r * K + k]) * (tf32 ? tf32r(b[k * N + n]) : b[k * N + n]);
            worst = fmax(worst, fabs(ref - d[r * N + n]));
        }
    printf("  %s  B MN-major, %s: max |err| = %.3e  %s\n", tf32 ? "tf32" : "f16 ", swapped ? "LBO=n-group, SBO=k-group" : "LBO=k-group, SBO=n-group", worst, worst < 1e-3 ? "MATCH" : "no");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return worst < 1e-3;
}

int main()
{
    for (int tf32 = 0; tf32 < 2; ++tf32)
        for (int sw = 0; sw < 2; ++sw) run(tf32, sw);
    return 0;
}
