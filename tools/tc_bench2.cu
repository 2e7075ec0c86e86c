// tc_bench2.cu -- second tcgen05 micro-benchmark / probe (B200, sm_100a).  Answers, on hardware:
//   COST   cycles per MMA in a back-to-back chain (issue -> completion barrier) for the shapes the fused kernel could use:
//          tf32 K=8 vs f16 K=16, M = 128 vs 64, N = 16..256, A operand from shared memory vs from TMEM
//   F16    numeric check of kind::f16 with the K-major / no-swizzle layout (8 halves per 16-byte row)
//   M64    which TMEM lanes hold the 64 rows of an M = 64 accumulator
//   LD16   register mapping of tcgen05.ld.16x256b
//   ATMEM  numeric check of an MMA whose A operand was written to TMEM with tcgen05.st (tf32 and f16)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_bench2 tc_bench2.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() { uint32_t ok; asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok)); return ok != 0; }
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
// kind: 0 = tf32 (A/B format 2), 1 = f16 (format 0), 2 = bf16 (format 1); D = f32
__host__ __device__ inline uint32_t mk_idesc(int kind, int M, int N) {
    const uint32_t fmt = kind == 0 ? 2u : (kind == 1 ? 0u : 1u);
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int KIND> __device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if constexpr (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND> __device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    if constexpr (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
}
__device__ __forceinline__ void setup(uint64_t* bar, uint32_t* tmem_s, int tid) {
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (tid < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_s))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void teardown(uint32_t tmem, int tid) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// ---------------------------------------------------------------- COST
// ATM = 1: A operand from TMEM (columns 256..).  The chain mimics a 3-tap conv layer: descriptors move between MMAs.
template <int KIND, int M, int N, int NMMA, int ATM, int AOFF = 32, int LBO = 2112>
__global__ void cost_kernel(long long* out)
{
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x;
    for (int i = tid; i < 24576; i += blockDim.x) sm[i] = 0.f;
    setup(&bar, &tmem_s, tid);
    const uint32_t tmem = tmem_s, idesc = mk_idesc(KIND, M, N);
    const uint32_t abase = smem_u32(sm), bbase = smem_u32(sm + 12288);
    long long t0 = clock64(), t1 = t0;
    if (tid < 32) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i) {
            const uint64_t da = mk_desc(abase + (uint32_t)(i % 6) * (2u * LBO) + (uint32_t)(i % 3) * (uint32_t)AOFF, (uint32_t)LBO, 128u);
            const uint64_t db = mk_desc(bbase + (uint32_t)(i % 8) * (uint32_t)(N * 32), (uint32_t)N * 16u, 128u);
            if (elect_one()) {
                if constexpr (ATM) mma_ts<KIND>(tmem, tmem + 256u + (uint32_t)(i % 6) * 8u, db, idesc, (uint32_t)(i > 0));
                else mma_ss<KIND>(tmem, da, db, idesc, (uint32_t)(i > 0));
            }
            __syncwarp();
        }
        t1 = clock64();
        if (elect_one()) commit(smem_u32(&bar));
    }
    wait_bar(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (tid == 0) { out[0] = t2 - t0; out[1] = t1 - t0; }
    teardown(tmem, tid);
}
template <int KIND, int M, int N, int NMMA, int ATM, int AOFF = 32, int LBO = 2112> void cost(const char* nm) {
    long long* d; CK(cudaMalloc(&d, 16)); long long h[2];
    CK(cudaFuncSetAttribute(cost_kernel<KIND, M, N, NMMA, ATM, AOFF, LBO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
    for (int r = 0; r < 3; ++r) { cost_kernel<KIND, M, N, NMMA, ATM, AOFF, LBO><<<1, 128, 98304>>>(d); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("COST %-5s M=%3d N=%3d A=%s aoff=%3d lbo=%4d chain=%2d : total %5lld cyc = %5.1f / MMA   (issue loop %5lld = %5.1f / MMA)\n", nm, M, N, ATM ? "tmem" : "smem", AOFF, LBO, NMMA,
           h[0], (double)h[0] / NMMA, h[1], (double)h[1] / NMMA);
    cudaFree(d);
}

// ---------------------------------------------------------------- F16 numeric / M64 layout / ATMEM numeric
// A [M rows][K], B [N][K] small integers / 16; result dumped as D[lane][col] for all 128 lanes via 32x32b loads.
template <int KIND, int M, int N, int NK, int ATM>
__global__ void num_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D)
{
    constexpr int KE = KIND == 0 ? 8 : 16;            // K per MMA
    constexpr int K = KE * NK;
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // operand layouts: [K / T][rows][T], T = 4 (tf32) or 8 (halves) elements per 16-byte row
    float* sA = sm; float* sB = sm + 8192;
    if constexpr (KIND == 0) {
        for (int i = tid; i < 128 * K; i += blockDim.x) { int r = i / K, k = i % K; sA[((k / 4) * 128 + r) * 4 + k % 4] = r < M ? A[r * K + k] : 0.f; }
        for (int i = tid; i < N * K; i += blockDim.x) { int n = i / K, k = i % K; sB[((k / 4) * N + n) * 4 + k % 4] = B[n * K + k]; }
    } else {
        __half* hA = reinterpret_cast<__half*>(sA); __half* hB = reinterpret_cast<__half*>(sB);
        for (int i = tid; i < 128 * K; i += blockDim.x) { int r = i / K, k = i % K; hA[((k / 8) * 128 + r) * 8 + k % 8] = __float2half(r < M ? A[r * K + k] : 0.f); }
        for (int i = tid; i < N * K; i += blockDim.x) { int n = i / K, k = i % K; hB[((k / 8) * N + n) * 8 + k % 8] = __float2half(B[n * K + k]); }
    }
    setup(&bar, &tmem_s, tid);
    const uint32_t tmem = tmem_s;
    if constexpr (ATM) {
        // every thread writes its row of A to TMEM columns 256..: tf32 one element per column, f16 two (low half = even k)
        const int r = warp * 32 + lane;
        constexpr int NC = KIND == 0 ? K : K / 2;
        for (int c = 0; c < NC; ++c) {
            uint32_t v;
            if constexpr (KIND == 0) v = __float_as_uint(r < M ? A[r * K + c] : 0.f);
            else { __half2 h2 = __floats2half2_rn(r < M ? A[r * K + 2 * c] : 0.f, r < M ? A[r * K + 2 * c + 1] : 0.f); v = *reinterpret_cast<uint32_t*>(&h2); }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + 256u + (uint32_t)c), "r"(v) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (tid == 0) {
        const uint32_t idesc = mk_idesc(KIND, M, N);
        for (int j = 0; j < NK; ++j) {
            const uint64_t da = mk_desc(smem_u32(sA) + (uint32_t)(2 * j) * 128 * 16, 128 * 16, 128);
            const uint64_t db = mk_desc(smem_u32(sB) + (uint32_t)(2 * j) * N * 16, N * 16, 128);
            if constexpr (ATM) mma_ts<KIND>(tmem, tmem + 256u + (uint32_t)j * 8u, db, idesc, (uint32_t)(j > 0));
            else mma_ss<KIND>(tmem, da, db, idesc, (uint32_t)(j > 0));
        }
        commit(smem_u32(&bar));
    }
    wait_bar(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N; ++c) {
        uint32_t v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        D[(warp * 32 + lane) * N + c] = __uint_as_float(v);
    }
    teardown(tmem, tid);
}
static float aval(int r, int k) { return (float)(((r * 3 + k * 5) % 17) - 8) / 16.0f; }
static float bval(int n, int k) { return (float)(((n * 7 + k * 3) % 13) - 6) / 8.0f; }
template <int KIND, int M, int N, int NK, int ATM> int num(const char* nm) {
    constexpr int K = (KIND == 0 ? 8 : 16) * NK;
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    for (int r = 0; r < 128; ++r) for (int k = 0; k < K; ++k) A[r * K + k] = aval(r, k);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[n * K + k] = bval(n, k);
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, D.size() * 4));
    CK(cudaFuncSetAttribute(num_kernel<KIND, M, N, NK, ATM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    num_kernel<KIND, M, N, NK, ATM><<<1, 128, 65536>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("NUM %-22s : launch failed: %s\n", nm, cudaGetErrorString(e)); exit(1); }
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    // find, for every logical row, the TMEM lane that holds it
    std::vector<int> lane_of(M, -1);
    double maxerr = 0;
    for (int r = 0; r < M; ++r) {
        std::vector<double> ref(N);
        for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)aval(r, k) * bval(n, k); ref[n] = s; }
        double best = 1e30; int bl = -1;
        for (int l = 0; l < 128; ++l) { double err = 0; for (int n = 0; n < N; ++n) err = fmax(err, fabs(ref[n] - D[l * N + n])); if (err < best) { best = err; bl = l; } }
        lane_of[r] = bl; maxerr = fmax(maxerr, best);
    }
    printf("NUM %-22s kind=%d M=%3d N=%3d K=%3d A=%s : max err %.2e %s ; row->lane:", nm, KIND, M, N, K, ATM ? "tmem" : "smem", maxerr, maxerr < 1e-5 ? "OK" : "FAIL");
    for (int r = 0; r < M; r += 8) printf(" %d:%d", r, lane_of[r]);
    bool ident = true; for (int r = 0; r < M; ++r) ident = ident && lane_of[r] == r;
    printf(ident ? "  (identity)\n" : "\n");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return maxerr < 1e-5 ? 0 : 1;
}

// ---------------------------------------------------------------- LD16: register mapping of tcgen05.ld.16x256b.x1 / 16x128b / 16x64b
__global__ void ld16_kernel(float* out)
{
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    setup(&bar, &tmem_s, tid);
    const uint32_t tmem = tmem_s, tw = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 32; ++c) {
        const uint32_t v = __float_as_uint((float)((warp * 32 + lane) * 100 + c));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tw + (uint32_t)c), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(tw));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 4; ++i) out[(0 * 128 + tid) * 4 + i] = __uint_as_float(r[i]);
    // same, starting at lane 16 of the quadrant
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(tw + (16u << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 4; ++i) out[(1 * 128 + tid) * 4 + i] = __uint_as_float(r[i]);
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(tw));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[(2 * 128 + tid) * 4 + 0] = __uint_as_float(r[0]); out[(2 * 128 + tid) * 4 + 1] = __uint_as_float(r[1]);
    out[(2 * 128 + tid) * 4 + 2] = out[(2 * 128 + tid) * 4 + 3] = -1.f;
    teardown(tmem, tid);
}
static void ld16() {
    float* d; CK(cudaMalloc(&d, 3 * 128 * 4 * 4)); std::vector<float> h(3 * 128 * 4);
    ld16_kernel<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("LD16 failed: %s\n", cudaGetErrorString(e)); exit(1); }
    CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
    const char* nm[3] = {"16x256b.x1 @lane0 ", "16x256b.x1 @lane16", "16x128b.x1 @lane0 "};
    for (int s = 0; s < 3; ++s)
        for (int w = 0; w < 2; ++w) {
            printf("LD16 %s warp %d (value = tmem_lane*100 + col):", nm[s], w);
            for (int l = 0; l < 32; l += (l < 8 ? 1 : 8)) {
                printf(" t%d[", l);
                for (int i = 0; i < 4; ++i) printf("%s%d", i ? "," : "", (int)h[(s * 128 + w * 32 + l) * 4 + i]);
                printf("]");
            }
            printf("\n");
        }
    cudaFree(d);
}

int main(int argc, char** argv) {
    // separate processes per group (a faulting probe must not take the others down): ./tc_bench2 1 .. 5
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which == 1) {
        int f = 0;
        f += num<0, 128, 48, 2, 0>("tf32 baseline");
        f += num<1, 128, 48, 2, 0>("f16 K-major no-swz");
        f += num<1, 128, 16, 1, 0>("f16 N=16 single");
        f += num<0, 64, 48, 2, 0>("tf32 M=64");
        f += num<1, 64, 48, 2, 0>("f16 M=64");
        printf("tc_bench2 numeric (smem A): %d failing\n", f);
    }
    if (which == 2) {
        cost<0, 128, 48, 1, 0>("tf32"); cost<0, 128, 48, 18, 0>("tf32");
        cost<0, 128, 48, 18, 0, 0, 2112>("tf32"); cost<0, 128, 48, 18, 0, 0, 2048>("tf32"); cost<0, 128, 48, 18, 0, 32, 2048>("tf32");
        cost<0, 128, 48, 18, 0, 16, 2048>("tf32"); cost<0, 128, 48, 18, 0, 64, 2048>("tf32"); cost<0, 128, 48, 18, 0, 128, 2048>("tf32");
        cost<0, 128, 48, 18, 0, 0, 2176>("tf32"); cost<0, 128, 48, 18, 0, 0, 768>("tf32"); cost<0, 128, 48, 18, 0, 0, 3072>("tf32");
        cost<0, 128, 16, 18, 0, 0, 2048>("tf32"); cost<0, 128, 144, 18, 0, 0, 2048>("tf32");
        cost<1, 128, 48, 18, 0>("f16"); cost<1, 128, 48, 18, 0, 0, 2048>("f16"); cost<1, 128, 144, 18, 0, 0, 2048>("f16");
        cost<0, 64, 48, 18, 0>("tf32"); cost<0, 64, 48, 18, 0, 0, 2048>("tf32"); cost<1, 64, 48, 18, 0, 0, 2048>("f16"); cost<0, 64, 144, 18, 0, 0, 2048>("tf32");
    }
    if (which == 3) {
        int f = 0;
        f += num<0, 128, 48, 2, 1>("tf32 A in TMEM");
        f += num<1, 128, 48, 2, 1>("f16 A in TMEM");
        f += num<0, 64, 48, 2, 1>("tf32 M=64 A in TMEM");
        printf("tc_bench2 numeric (tmem A): %d failing\n", f);
    }
    if (which == 4) { cost<0, 128, 48, 18, 1>("tf32"); cost<1, 128, 48, 18, 1>("f16"); cost<0, 128, 144, 18, 1>("tf32"); cost<0, 64, 48, 18, 1>("tf32"); }
    if (which == 5) ld16();
    return 0;
}
