// fe_stft_gemm.cu -- the STFT as a tensor-core GEMM (the reference's ConvSTFT front end: models/fastenhancer/conv_stft/model.py:55-63,
// 110-114 -- F.conv1d of the waveform with the windowed DFT basis, stride = hop), an alternative to the packed-real radix-4 FFT of the
// fused kernel for callers that transform many frames at once.
//
//   spec[b][k][t] = sum_n wav[b][t H + n] * w[n] e^{-2 pi i k n / N}        k = 0 .. N/2,  t = 0 .. T-1
//
// GEMM view: rows = frames (b, t), contraction over the N samples of a frame, columns = the N real outputs of a real DFT
// packed as (re_0, re_{N/2}), (re_1, im_1), ..., (re_{N/2-1}, im_{N/2-1}) (the imaginary parts of DC and Nyquist are zero).
//   * The frames are never materialised: the A operand is a 3-D TMA tensor map over the waveform itself with OVERLAPPING rows
//     (dim 0 = sample in frame, stride 1; dim 1 = frame, stride H; dim 2 = utterance, stride ld), 128-byte swizzle, box 32 x 128.
//   * tcgen05.mma kind::tf32, M = 128 frames x N = 128 columns x K = 8 per instruction, accumulators in tensor memory (four of them,
//     one per quarter of the contraction, all 512 columns: summed in the epilogue, which keeps the truncating accumulation chains short).
//   * fp32-accurate mode (default): 3xTF32 -- a = a_hi + a_lo split in shared memory by the epilogue warps as the tiles land
//     (cvt.rna.tf32), the basis split on the host in double precision; a_hi b_hi + a_lo b_hi + a_hi b_lo into the same accumulator.
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2..5 = operand split, then epilogue
//     (tcgen05.ld, packed (re, im) stores); 3-stage mbarrier ring (full -> split-ready -> empty by tcgen05.commit).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "fe_stft_gemm.h"

namespace fe {
namespace {

constexpr int BM = 128, BN = 128, KC = 32, STAGES = 3;                 // frames x columns per CTA, samples per pipeline stage
constexpr int TILE_BYTES = BM * KC * 4;                                // 16 KB: one operand tile (128 rows x 128 bytes)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;                            // A (hi) | A lo | B hi | B lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment */ + 256 /* barriers */;
constexpr int NTHREADS = 192;
constexpr int NSEG = 4;                  // the contraction accumulates in NSEG tensor-memory accumulators (one per quarter of the samples), summed in the
                                         // epilogue with rounding fp32 adds: the tensor core's own accumulation truncates, and its error grows with the chain

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {      // bounded: a protocol bug traps instead of hanging the GPU
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_3d(uint32_t dst, const CUtensorMap* tm, int x, int y, int z, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* tm, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(bar) : "memory");
}
// K-major operand tile of 128-byte rows in the 128-byte swizzle (what the TMA box above writes): rows 128 B apart, groups of 8 rows
// 1024 B apart (SBO), swizzle mode 2 = SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp), descriptor version 1; a step of 8 tf32 along K is
// +32 bytes on the start address.
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float tf32_rna(float x) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return __uint_as_float(u); }

struct Params {
    float* out;             // [B][NB][T][2]
    int B, T, NB, nkc;      // utterances, frames per utterance, N/2 + 1, k-chunks (N / KC)
    int tiles_per_utt;      // ceil(T / BM)
    int x3;                 // 1 = fp32-accurate 3xTF32, 0 = single TF32 pass
};

__global__ void __launch_bounds__(NTHREADS, 1)
fe_stft_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_bh, const __grid_constant__ CUtensorMap tm_bl,
                    const Params prm)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    const uint32_t bar0 = s32(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto ready = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto empty = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
    const uint32_t accum = bar0 + 8u * (3 * STAGES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / prm.tiles_per_utt, t0 = (blockIdx.x % prm.tiles_per_utt) * BM, n0 = blockIdx.y * BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full(s), 1); mbar_init(ready(s), 128); mbar_init(empty(s), 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {         // the MMA warp owns the tensor-memory allocation: NSEG x 128 fp32 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "n"(NSEG * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            for (int kc = 0; kc < prm.nkc; ++kc) {
                const int s = kc % STAGES;
                if (kc >= STAGES) mbar_wait(empty(s), ((kc / STAGES) - 1) & 1);
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_expect_tx(full(s), (uint32_t)(TILE_BYTES * (prm.x3 ? 3 : 2)));
                tma_3d(s32(st), &tm_a, kc * KC, t0, b, full(s));                        // 128 frames x 32 samples, rows overlap in memory
                tma_2d(s32(st + 2 * TILE_BYTES), &tm_bh, kc * KC, n0, full(s));         // 128 basis rows x 32 samples
                if (prm.x3) tma_2d(s32(st + 3 * TILE_BYTES), &tm_bl, kc * KC, n0, full(s));
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        // instruction descriptor: D fp32, A / B tf32, both K-major, M = 128, N = BN
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (int kc = 0; kc < prm.nkc; ++kc) {
            const int s = kc % STAGES;
            mbar_wait(prm.x3 ? ready(s) : full(s), (kc / STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t a = s32(smem + s * STAGE_BYTES);
                const uint64_t da = desc_sw128(a), dal = desc_sw128(a + TILE_BYTES), dbh = desc_sw128(a + 2 * TILE_BYTES), dbl = desc_sw128(a + 3 * TILE_BYTES);
                const int per = prm.nkc / NSEG, seg = kc / per;              // accumulator of this quarter of the contraction
                const uint32_t d = tmem + (uint32_t)(seg * BN);
#pragma unroll
                for (int ks = 0; ks < KC / 8; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 32 >> 4);
                    mma_tf32(d, da + adv, dbh + adv, idesc, (kc % per > 0 || ks > 0) ? 1u : 0u);
                    if (prm.x3) {
                        mma_tf32(d, dal + adv, dbh + adv, idesc, 1u);
                        mma_tf32(d, da + adv, dbl + adv, idesc, 1u);
                    }
                }
                mma_commit(empty(s));                       // the stage is free once these MMAs have read it
                if (kc == prm.nkc - 1) mma_commit(accum);   // ... and after the last one the accumulator is complete
            }
            __syncwarp();
        }
    } else {
        // ---------------- operand split (fp32-accurate mode), then epilogue: 4 warps = 128 threads ----------------
        const int et = threadIdx.x - 64;
        if (prm.x3) {
            for (int kc = 0; kc < prm.nkc; ++kc) {
                const int s = kc % STAGES;
                mbar_wait(full(s), (kc / STAGES) & 1);
                float4* a = reinterpret_cast<float4*>(smem + s * STAGE_BYTES);
                float4* al = reinterpret_cast<float4*>(smem + s * STAGE_BYTES + TILE_BYTES);
#pragma unroll
                for (int i = 0; i < TILE_BYTES / 16 / 128; ++i) {       // elementwise, so the swizzled placement carries over to the lo tile
                    const float4 v = a[et + 128 * i];
                    float4 h, l;
                    h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
                    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                    a[et + 128 * i] = h; al[et + 128 * i] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> the MMA's async-proxy reads
                mbar_arrive(ready(s));
            }
        }
        mbar_wait(accum, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quad = warp & 3, row = quad * 32 + lane, t = t0 + row;       // a warp reads the TMEM lane quadrant warp % 4
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            float acc[32];
#pragma unroll
            for (int seg = 0; seg < NSEG; ++seg) {
                uint32_t r[32];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                             "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                               "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                               "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                               "=r"(r[30]), "=r"(r[31])
                             : "r"(taddr + (uint32_t)(seg * BN + c0)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = seg == 0 ? __uint_as_float(r[j]) : acc[j] + __uint_as_float(r[j]);
            }
            if (t < prm.T) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const int k = (n0 + c0 + j) >> 1;         // bin of this column pair
                    const float re = acc[j], im = acc[j + 1];
                    float2* o = reinterpret_cast<float2*>(prm.out + (((size_t)b * prm.NB + k) * prm.T + t) * 2);
                    if (k == 0) {                             // the pair (re_0, re_{N/2}): both bins are real
                        *o = make_float2(re, 0.f);
                        *reinterpret_cast<float2*>(prm.out + (((size_t)b * prm.NB + (prm.NB - 1)) * prm.T + t) * 2) = make_float2(im, 0.f);
                    } else *o = make_float2(re, im);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(NSEG * BN) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
float tf32_round_host(double v) {          // nearest value with a 10-bit stored significand (what cvt.rna.tf32 produces), as a float
    float f = (float)v;
    uint32_t u; std::memcpy(&u, &f, 4);
    u = (u + 0x1000u) & 0xffffe000u;
    std::memcpy(&f, &u, 4);
    return f;
}

}  // namespace

// windowed real-DFT basis [N][N] (row = packed output column, column = sample), split hi + lo for 3xTF32; `window` = the N analysis taps
void stft_gemm_basis(int n_fft, const float* window, std::vector<float>& hi, std::vector<float>& lo) {
    const int N = n_fft;
    hi.assign((size_t)N * N, 0.f); lo.assign((size_t)N * N, 0.f);
    for (int c = 0; c < N; ++c) {
        const int k = c >> 1;
        for (int n = 0; n < N; ++n) {
            double v;
            if (c == 0) v = 1.0;                                                 // re_0
            else if (c == 1) v = (n & 1) ? -1.0 : 1.0;                           // re_{N/2}
            else {
                const double ph = -2.0 * M_PI * (double)((long long)k * n % N) / N;
                v = (c & 1) ? std::sin(ph) : std::cos(ph);
            }
            v *= (double)window[n];
            const float h = tf32_round_host(v);
            hi[(size_t)c * N + n] = h;
            lo[(size_t)c * N + n] = tf32_round_host(v - (double)h);
        }
    }
}

int stft_gemm_smem_bytes() { return SMEM_BYTES; }

// 0 = ok, 1 = bad alignment / pitch for the tensor maps, 2 = tensor-map encoding failed, 3 = CUDA error (*cuda_err)
int stft_gemm_launch(const float* wav, long long ld, int B, int T, int n_fft, int hop, const float* basis_hi, const float* basis_lo, float* spec_out,
                     int accurate, cudaStream_t stream, cudaError_t* cuda_err) {
    *cuda_err = cudaSuccess;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return 2;
    if ((reinterpret_cast<size_t>(wav) & 15) != 0 || (ld % 4) != 0 || (hop % 4) != 0 || n_fft % BN != 0 || (n_fft / KC) % NSEG != 0) return 1;
    CUtensorMap tm_a, tm_bh, tm_bl;
    {
        const cuuint64_t gdim[3] = {(cuuint64_t)n_fft, (cuuint64_t)T, (cuuint64_t)B};
        const cuuint64_t gstr[2] = {(cuuint64_t)hop * 4, (cuuint64_t)ld * 4};          // frames overlap: row stride = hop samples
        const cuuint32_t box[3] = {KC, BM, 1}, estr[3] = {1, 1, 1};
        if (enc(&tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(wav), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 2;
    }
    for (int part = 0; part < 2; ++part) {
        const cuuint64_t gdim[2] = {(cuuint64_t)n_fft, (cuuint64_t)n_fft};
        const cuuint64_t gstr[1] = {(cuuint64_t)n_fft * 4};
        const cuuint32_t box[2] = {KC, BN}, estr[2] = {1, 1};
        if (enc(part ? &tm_bl : &tm_bh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(part ? basis_lo : basis_hi), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 2;
    }
    static bool prepared = false;
    if (!prepared) {
        *cuda_err = cudaFuncSetAttribute(fe_stft_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (*cuda_err != cudaSuccess) return 3;
        prepared = true;
    }
    Params prm{};
    prm.out = spec_out; prm.B = B; prm.T = T; prm.NB = n_fft / 2 + 1; prm.nkc = n_fft / KC; prm.tiles_per_utt = (T + BM - 1) / BM; prm.x3 = accurate ? 1 : 0;
    const dim3 grid((unsigned)(prm.tiles_per_utt * B), (unsigned)(n_fft / BN));
    fe_stft_gemm_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(tm_a, tm_bh, tm_bl, prm);
    *cuda_err = cudaGetLastError();
    return *cuda_err == cudaSuccess ? 0 : 3;
}

}  // namespace fe
