// fe_api.cu -- C ABI (include/fastenhancer_b200.h) over the fused per-hop kernel variants.
//
// Host-side responsibilities only: pick the (config, streams-per-CTA) variant, pack the canonical
// weights into the variant's blob once, own device buffers for weights / state / spill scratch,
// launch.  There is no CPU compute path: without a CUDA device every entry point fails loudly.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fastenhancer_b200.h"
#include "fe_variant.h"
#include "fe_kernel.cuh"
#include "fe_stft_gemm.h"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
int cuda_fail(cudaError_t e, const char* what) {
    return fail(FE_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
#define FE_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

std::vector<fe::VariantOps> all_variants() {
    std::vector<fe::VariantOps> v;
    const fe::VariantOps* (*fns[])(int*) = {fe::variants_16t, fe::variants_16b, fe::variants_16s, fe::variants_16m, fe::variants_16l,
                                            fe::variants_48t, fe::variants_48b, fe::variants_48s, fe::variants_48m, fe::variants_48l};
    for (auto fn : fns) { int n = 0; const fe::VariantOps* p = fn(&n); v.insert(v.end(), p, p + n); }
    return v;
}

bool same_shape(const fe::ShapeKey& k, const fe_config& c) {
    return k.n_fft == c.n_fft && k.hop == c.hop && k.c1 == c.c1 && k.n_enc == c.n_enc && k.c2 == c.c2 && k.f2 == c.f2 &&
           k.n_blocks == c.n_blocks && k.n_heads == c.n_heads;
}

size_t weight_count(const fe_config& c) {
    size_t C1 = c.c1, C2 = c.c2, F1 = c.n_fft / 8, F2 = c.f2, n = 0;
    n += C1 * 16 + C1;
    n += (size_t)c.n_enc * (C1 * C1 * 3 + C1);
    n += F2 * F1 + C2 * C1 + C2;
    n += (size_t)c.n_blocks * (2 * 3 * C2 * C2 + 2 * 3 * C2 + C2 * C2 + C2 + 3 * C2 * C2 + 3 * C2 + C2 * C2 + C2);
    n += F2 * C2;
    n += F1 * F2 + C1 * C2 + C1;
    n += (size_t)c.n_enc * (C1 * 2 * C1 + C1 + C1 * C1 * 3 + C1);
    n += C1 * 2 * C1 + C1 + C1 * 16 + 2;
    return n;
}

// Per-stream rows [cache_stft | cache_istft | h_0 [F2][C2] | ... | h_{K-1}] (the export layout of the C ABI) <-> the planes the
// kernels keep: cache_stft [B][CL] | cache_istft [B][CL] | h_k [B][F2][C2], k < K.  Element order inside every piece is the same.
__global__ void state_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_streams, int cl, int K, int hf,
                                       int to_planes)
{
    const int sf = 2 * cl + K * hf;
    const long total = (long)n_streams * sf;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int s = (int)(i / sf), r = (int)(i % sf);            // i indexes the per-stream-row form
        long pl;
        if (r < 2 * cl) pl = ((long)(r / cl) * n_streams + s) * cl + r % cl;
        else { const int q = r - 2 * cl; pl = 2L * n_streams * cl + ((long)(q / hf) * n_streams + s) * hf + q % hf; }
        if (to_planes) dst[pl] = src[i];
        else dst[i] = src[pl];
    }
}

// fp32 FMA-pipe peak microbenchmark (roofline denominator of the fp32 FMA-pipe kernel family): 16 independent FFMA chains per thread
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float a, float b)
{
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s;          // never true: keeps the chains alive
}

// ---- checkpoint folding on the device (fe_fold_device): one launch per rule of ONNXModel.remove_weight_reparameterizations
//      (models/fastenhancer/default/model.py:532-608, 215-258, 74-81); scale factors in double, rounded once, like the host oracle
//      fastenhancer_b200/fold.py ----
__device__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    __syncthreads();
    return t;
}
// one block per row (kinds 1, 2); one block for the whole tensor (kind 3); grid-stride copy (kind 0)
__global__ void __launch_bounds__(128) fold_kernel(fe_fold_op op, float* __restrict__ canon)
{
    __shared__ double red[4];
    float* dst = canon + op.dst;
    if (op.kind == FE_FOLD_COPY) {
        for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < (long)op.rows * op.cols; i += (long)gridDim.x * blockDim.x) dst[i] = op.w[i];
        return;
    }
    if (op.kind == FE_FOLD_FINAL_CONV) {          // scale * W / max(||W||_F, 1e-12)  (normalize_final_conv) or scale * W
        const long n = (long)op.rows * op.cols;
        double ss = 0.0;
        for (long i = threadIdx.x; i < n; i += blockDim.x) ss += (double)op.w[i] * (double)op.w[i];
        ss = block_sum(ss, red);
        double f = (double)op.a[0];
        if (op.flag) f /= fmax(sqrt(ss), 1e-12);
        for (long i = threadIdx.x; i < n; i += blockDim.x) dst[i] = (float)((double)op.w[i] * f);
        return;
    }
    const int r = blockIdx.x;
    const float* wr = op.w + (long)r * op.cols;
    double f;
    if (op.kind == FE_FOLD_WEIGHT_NORM) {         // g[r] * v[r][:] / ||v[r]||   (torch _weight_norm, dim = 0)
        double ss = 0.0;
        for (int i = threadIdx.x; i < op.cols; i += blockDim.x) ss += (double)wr[i] * (double)wr[i];
        f = (double)op.a[r] / sqrt(block_sum(ss, red));
    } else {                                      // eval-mode BatchNorm folded into the preceding bias-free conv / linear
        f = (double)op.a[r] / sqrt((double)op.d[r] + (double)op.eps);
        if (threadIdx.x == 0 && op.bias_dst >= 0) canon[op.bias_dst + r] = (float)((double)op.b[r] - (double)op.c[r] * f);
    }
    for (int i = threadIdx.x; i < op.cols; i += blockDim.x) dst[(long)r * op.cols + i] = (float)((double)wr[i] * f);
}

// ---- frame-parallel offline schedule: the two small kernels between the staged launches of the fused kernel ----
// The GRU recurrence of one RNNFormer block (models/fastenhancer/default/model.py:266-272: nn.GRU over time with the sub-bands as batch),
// the only sequential part of Model.forward: one CTA per (utterance, sub-band) row walks t = 0 .. T-1.  Thread (j, q) keeps the
// hidden-to-hidden weights of output channel j for the q-th part of the input channels in registers (3 gates x CQ weights), partial
// sums meet through a shuffle butterfly over the Q lanes, h goes through a double-buffered shared-memory vector (one barrier per
// step), the input-side pre-activations gx (computed frame-parallel by the previous stage) are prefetched four steps ahead.
//   gx [B*T][F2][3][C2] (W_ir x | W_iz x | W_in x, no bias), w_hh [3][C2][C2], b_ih / b_hh [3][C2]  ->  hseq [B*T][F2][C2]
template <int C2, int Q> __global__ void __launch_bounds__(((C2 * Q + 31) / 32) * 32)
fe_gru_scan_kernel(const float* __restrict__ gx, float* __restrict__ hseq, const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                   const float* __restrict__ b_hh, int T, int F2)
{
    constexpr int CQ = (C2 + Q - 1) / Q, NTH = ((C2 * Q + 31) / 32) * 32, HP = CQ * Q + 4;
    constexpr int NSTG = 8, GROW = 3 * C2;          // gx ring: the pre-activations of the next NSTG steps (cp.async, 16-byte pieces)
    static_assert((Q & (Q - 1)) == 0 && Q <= 32 && (CQ % 2 == 0 || Q == 1), "input parts: a power of two of lanes, 8-byte aligned slices of h");
    static_assert(GROW % 4 == 0, "a step's pre-activations are whole 16-byte pieces");
    __shared__ __align__(16) float hbuf[2][HP];
    __shared__ __align__(16) float gbuf[NSTG][GROW];
    const int tid = threadIdx.x, j = tid / Q, q = tid % Q, c0 = q * CQ;
    const bool live = j < C2;
    const int u = blockIdx.x / F2, f = blockIdx.x % F2;
    const float* g0 = gx + (((size_t)u * T) * F2 + f) * GROW;              // + t * F2 * GROW
    const size_t gstep = (size_t)F2 * GROW, hstep = (size_t)F2 * C2;
    static_assert(GROW / 4 <= NTH, "one 16-byte piece per thread and step");
    const float* gsrc = g0 + 4 * tid;                                       // this thread's piece of a step (threads < GROW / 4)
    const bool piece = tid < GROW / 4;
    const uint32_t gdst = (uint32_t)__cvta_generic_to_shared(&gbuf[0][0]) + 16u * tid;
    // every thread commits one (possibly empty) group per step: the waits below count groups
    auto issue = [&](int t, int slot) {
        if (piece && t < T) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(gdst + (uint32_t)(slot * GROW * 4)), "l"(gsrc + (size_t)t * gstep) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int t = 0; t < NSTG; ++t) issue(t, t);
    float wr[CQ], wz[CQ], wn[CQ];
#pragma unroll
    for (int i = 0; i < CQ; ++i) {
        const bool ok = live && c0 + i < C2;
        wr[i] = ok ? w_hh[(0 * C2 + j) * C2 + c0 + i] : 0.f;
        wz[i] = ok ? w_hh[(1 * C2 + j) * C2 + c0 + i] : 0.f;
        wn[i] = ok ? w_hh[(2 * C2 + j) * C2 + c0 + i] : 0.f;
    }
    const int jj = live ? j : 0;
    const float br = b_ih[jj] + b_hh[jj], bz = b_ih[C2 + jj] + b_hh[C2 + jj], bin = b_ih[2 * C2 + jj], bhn = b_hh[2 * C2 + jj];
    for (int i = tid; i < 2 * HP; i += NTH) (&hbuf[0][0])[i] = 0.f;
    asm volatile("cp.async.wait_group %0;" ::"n"(NSTG - 1) : "memory");      // step 0 has landed
    if (NTH > 32) __syncthreads(); else __syncwarp();
    float* hout = hseq + (((size_t)u * T) * F2 + f) * C2 + jj;
    float hj = 0.f;                // h[j] of this thread's channel (lanes q == 0)
    // NSTG steps per trip: the ring slot (t % NSTG) and the h buffer (t & 1) of every step are compile-time facts
    for (int t0 = 0; t0 < T; t0 += NSTG) {
#pragma unroll
        for (int d = 0; d < NSTG; ++d) {
            const int t = t0 + d;
            if (t < T) {       // uniform
                if (t > 0) issue(t - 1 + NSTG, (d + NSTG - 1) % NSTG);   // the slot of step t - 1 is free: every thread passed the barrier that ended it
                const float* hb = hbuf[d & 1];
                const float* gb = gbuf[d];
                const float gr = gb[jj], gz = gb[C2 + jj], gn = gb[2 * C2 + jj];
                float ar0 = 0.f, az0 = 0.f, an0 = 0.f, ar1 = 0.f, az1 = 0.f, an1 = 0.f;
#pragma unroll
                for (int i = 0; i + 1 < CQ; i += 2) {
                    const float2 hv = *reinterpret_cast<const float2*>(hb + c0 + i);
                    ar0 = fmaf(wr[i], hv.x, ar0); az0 = fmaf(wz[i], hv.x, az0); an0 = fmaf(wn[i], hv.x, an0);
                    ar1 = fmaf(wr[i + 1], hv.y, ar1); az1 = fmaf(wz[i + 1], hv.y, az1); an1 = fmaf(wn[i + 1], hv.y, an1);
                }
                if (CQ & 1) { const float hv = hb[c0 + CQ - 1]; ar0 = fmaf(wr[CQ - 1], hv, ar0); az0 = fmaf(wz[CQ - 1], hv, az0); an0 = fmaf(wn[CQ - 1], hv, an0); }
                float ar = ar0 + ar1, az = az0 + az1, an = an0 + an1;
#pragma unroll
                for (int o = 1; o < Q; o <<= 1) {
                    ar += __shfl_xor_sync(0xffffffffu, ar, o); az += __shfl_xor_sync(0xffffffffu, az, o); an += __shfl_xor_sync(0xffffffffu, an, o);
                }
                if (q == 0) {
                    const float r = fe::sigmoid_acc(gr + br + ar), z = fe::sigmoid_acc(gz + bz + az);
                    const float n = fe::tanh_acc(gn + bin + r * (an + bhn));
                    hj = (1.0f - z) * n + z * hj;
                    if (live) { hbuf[(d + 1) & 1][j] = hj; *hout = hj; }
                }
                hout += hstep;
                asm volatile("cp.async.wait_group %0;" ::"n"(NSTG - 2) : "memory");  // step t + 1 has landed (this thread's piece; the barrier publishes all)
                if (NTH > 32) __syncthreads(); else __syncwarp();
            }
        }
    }
}

// out[u][n] = sum_t y_t[n + N/2 - t H] / sum_t w^2[n + N/2 - t H], n < H (T - 1): torch.istft(center=True) over the windowed frames
// y [B*T][N] (functional/audio_modules.py:108-121); frames summed in ascending t like the sequential walk.
__global__ void fe_overlap_add_kernel(const float* __restrict__ frames, const float* __restrict__ wsq, float* __restrict__ out, int B, int T, int N, int H)
{
    const long len = (long)H * (T - 1), total = (long)B * len;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long u = i / len, n = i % len, npad = n + N / 2;
        long t0 = npad < N ? 0 : (npad - N + H) / H, t1 = npad / H;
        if (t1 > T - 1) t1 = T - 1;
        float v = 0.f, env = 0.f;
        for (long t = t0; t <= t1; ++t) { v += frames[((size_t)u * T + t) * N + (npad - t * H)]; env += __ldg(wsq + (npad - t * H)); }
        out[i] = v / env;
    }
}

struct Variant {
    fe::VariantOps ops;
    float* blob = nullptr;     // device
    bool prepared = false;
};

}  // namespace

struct fe_engine {
    fe_config cfg;
    int device = 0, num_sms = 148;
    std::vector<float> canonical;
    std::vector<Variant> variants;       // same shape, ascending S
    int forced_s = 0;
    int tc = 1;                          // kernel family (Plan::PREC): 0 fp32 FMA pipe, 1 TF32, 2 fp16, 3 bf16 conv section, 4 split fp16 (fp32-accurate)
    long long* prof = nullptr;           // optional per-phase cycle counters (device)
    int hop_slicing = 1;                 // multi-round streaming launches are cut into hop ranges on a persistent grid (FE_HOP_SLICING=0: off)
    int hop_tma = 1;                     // hop tiles by TMA where the variant supports it (FE_HOP_TMA=0 in the environment: plain loads / stores)
    std::atomic<long long> launches{0};
    std::mutex mu;
    cudaMemPool_t pool = nullptr;        // stream-ordered scratch of fe_offline: an engine-owned pool that keeps its memory between calls
    int offline_mode = 0;                // fe_offline: 0 = automatic, 1 = sequential walk (one CTA per stream group), 2 = frame-parallel schedule
    float* basis_dev = nullptr;          // windowed DFT basis of the tensor-core STFT (hi | lo, [2][N][N]), built on first use
    float* canon_dev = nullptr;          // canonical weights on the device (hidden-to-hidden GRU weights of the scan), uploaded on first use
};

// Everything a call mutates lives in the state (or on the call's stream), never in the engine: two states of one engine can be
// driven from two host threads / CUDA streams concurrently.
struct fe_state {
    fe_engine* e;
    int n_streams;
    float* data = nullptr;      // planes (fe_state_planes): cache_stft | cache_istft | h_0 .. h_{K-1}
    bool owns_data = true;      // false: the caller's buffer (fe_state_create_on)
    float* scratch = nullptr;   // spill scratch for the largest grid (S = 1)
    size_t scratch_floats = 0;
    CUtensorMap* tmaps = nullptr;   // device copy of the two hop-tile tensor maps of the launch in flight (fe_stream)
    int* flags = nullptr;           // item-done flags of a hop-sliced launch (zeroed in stream order before it)
    size_t flags_n = 0;
    // pipelined host path (fe_stream_host): staging buffers, streams and events of THIS state, created by fe_state_reserve_host
    cudaStream_t s_copy_in = nullptr, s_compute = nullptr, s_copy_out = nullptr;
    float* h_in[2] = {nullptr, nullptr}; float* h_out[2] = {nullptr, nullptr}; size_t h_floats = 0;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr}, ev_call = nullptr;
};

namespace {

// Which streams-per-CTA variant serves `n_streams` streams fastest.  The chain per hop is mostly fixed latency, so a CTA with S streams
// costs about r(S) = 1 / 1.35 / 2.4 (S = 1 / 2 / 4; measured: B 1.02-1.06 and M 1.37 for two streams, B f16 1.84 for four -- the upper
// ends, so that ties go to the smaller, lower-latency variant) of a one-stream CTA, and the launch takes rounds(S) x r(S): whole rounds
// of one CTA per SM, or -- where hop-sliced launches even out the last round (variants without hop-tiled rings) -- the fractional
// number.  E.g. B: 100 streams -> S = 1 (one round either way), 200 or 256 -> S = 2 (one round instead of two), 4096 (f16) -> S = 4.
int pick_variant(fe_engine* e, int n_streams) {
    // variants are sorted by (tc, S); only those of the engine's precision mode are eligible
    if (e->forced_s > 0) {
        for (size_t i = 0; i < e->variants.size(); ++i)
            if (e->variants[i].ops.tc == e->tc && e->variants[i].ops.S == e->forced_s) return (int)i;
    }
    int best = -1;
    double best_cost = 0.0;
    for (size_t i = 0; i < e->variants.size(); ++i) {
        if (e->variants[i].ops.tc != e->tc) continue;
        const int S = e->variants[i].ops.S;
        const double groups = (double)((n_streams + S - 1) / S);
        const bool sliced = e->hop_slicing && !e->variants[i].ops.hop_ring;
        const double rounds = (sliced && groups > e->num_sms) ? 1.02 * groups / e->num_sms : std::ceil(groups / e->num_sms);
        const double cost = rounds * (S == 1 ? 1.0 : (S == 2 ? 1.35 : 0.6 * S));
        if (best < 0 || cost < best_cost - 1e-9) { best = (int)i; best_cost = cost; }
    }
    return best;
}

int ensure_variant(fe_engine* e, int vi) {
    Variant& v = e->variants[vi];
    std::lock_guard<std::mutex> lk(e->mu);
    if (!v.blob) {
        std::vector<float> blob;
        try { v.ops.pack(e->canonical.data(), blob); }
        catch (const std::exception& ex) { return fail(FE_ERR_ARG, std::string("weight packing failed: ") + ex.what()); }
        if ((long)blob.size() != v.ops.blob_floats) return fail(FE_ERR_ARG, "packed blob size mismatch");
        FE_CUDA(cudaMalloc(&v.blob, blob.size() * sizeof(float)));
        FE_CUDA(cudaMemcpy(v.blob, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (!v.prepared) { FE_CUDA(v.ops.prepare()); v.prepared = true; }
    return FE_OK;
}

size_t scratch_need(const fe_engine* e, int n_streams) {
    size_t need = 4;
    for (const Variant& v : e->variants) {
        size_t grid = (size_t)(n_streams + v.ops.S - 1) / v.ops.S;
        need = std::max(need, grid * (size_t)v.ops.gs_floats);
    }
    return need;
}

// 2-D tensor map of a [n_streams][ld] float array restricted to its first `width` columns, box [S][tile]: the TMA descriptor of the
// input / output hop tiles.  cuTensorMapEncodeTiled comes from the driver through the runtime (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
bool hop_tensor_map(CUtensorMap* tm, const float* base, long long ld, long long width, int n_streams, int S, int tile) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn || (reinterpret_cast<size_t>(base) & 15) != 0 || (ld % 4) != 0 || width <= 0) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)width, (cuuint64_t)n_streams};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)tile, (cuuint32_t)S}, estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Hop-sliced streaming launches (fe_kernel.cuh::Frame::run): with more stream groups than SMs the last round of CTAs leaves SMs idle
// (256 groups on 148 SMs: 2 rounds for 1.73 rounds of work).  Cutting the launch into R hop ranges and dealing the items (range, group)
// round-robin to one persistent CTA per SM evens it out.  Returns the hops per range that minimises the simulated makespan (items wait
// for the previous range of their streams), or 0 when slicing gains less than 4 %.
int plan_slices(int ngrp, int n_hops, int num_sms) {
    if (ngrp <= num_sms || n_hops < 8) return 0;
    const double base = (double)((ngrp + num_sms - 1) / num_sms);
    double best = base * 0.96;
    int best_hops = 0;
    std::vector<double> fin;
    for (int R = 2; R <= 8; ++R) {
        const int hops = (n_hops + R - 1) / R;
        if (hops < 4) break;
        const int nr = (n_hops + hops - 1) / hops, total = ngrp * nr, G = std::min(num_sms, total);
        fin.assign(total, 0.0);
        std::vector<double> cta(G, 0.0);
        double span = 0.0;
        for (int i = 0; i < total; ++i) {
            const int r = i / ngrp, h0 = r * hops, len = std::min(hops, n_hops - h0);
            double start = cta[i % G];
            if (r > 0) start = std::max(start, fin[i - ngrp]);
            fin[i] = start + (double)len / n_hops + 0.003;        // + state hand-over and pipeline refill of an item
            cta[i % G] = fin[i];
            span = std::max(span, fin[i]);
        }
        if (span < best) { best = span; best_hops = hops; }
    }
    return best_hops;
}

int launch(fe_engine* e, fe::KParams prm, float* scratch, cudaStream_t st, CUtensorMap* tmaps_device = nullptr, fe_state* sliced = nullptr) {
    if (prm.n_streams <= 0 || prm.n_hops <= 0) return FE_OK;
    const int vi = pick_variant(e, prm.n_streams);
    int rc = ensure_variant(e, vi);
    if (rc) return rc;
    const Variant& v = e->variants[vi];
    prm.blob = v.blob;
    prm.scratch = scratch;
    prm.compression = e->cfg.compression;
    prm.prof = e->prof;
    int grid = (prm.n_streams + v.ops.S - 1) / v.ops.S;
    prm.slice_hops = 0;
    prm.slice_flags = nullptr;
    if (sliced && prm.mode == fe::MODE_STREAM && !prm.dbg && e->hop_slicing && !v.ops.hop_ring) {      // (compiled into the variants without hop-tiled rings)
        const int hops = plan_slices(grid, prm.n_hops, e->num_sms);
        if (hops > 0) {
            const size_t items = (size_t)grid * ((prm.n_hops + hops - 1) / hops);
            if (sliced->flags_n < items) {
                if (sliced->flags) { FE_CUDA(cudaStreamSynchronize(st)); FE_CUDA(cudaFree(sliced->flags)); sliced->flags = nullptr; sliced->flags_n = 0; }
                FE_CUDA(cudaMalloc(&sliced->flags, items * sizeof(int)));
                sliced->flags_n = items;
            }
            FE_CUDA(cudaMemsetAsync(sliced->flags, 0, items * sizeof(int), st));
            prm.slice_hops = hops;
            prm.slice_flags = sliced->flags;
            grid = (int)std::min<size_t>((size_t)e->num_sms, items);
        }
    }
    // streaming launches of the variants whose rings are hop-tiled: the input hop arrives and the output hop leaves as 2-D TMA tiles
    // [S streams][hop tile] (cp.async.bulk.tensor); plain loads / stores when the caller's arrays are not 16-byte aligned / pitched
    // (single-hop launches keep the plain path: nothing to prefetch, and the descriptors would cost an upload per hop)
    prm.hop_tma = 0;
    prm.tmaps = nullptr;
    if (prm.mode == fe::MODE_STREAM && v.ops.hop_ring && e->hop_tma && tmaps_device && prm.n_hops >= 2) {
        CUtensorMap tm[2];
        const long long width = (long long)prm.n_hops * e->cfg.hop;
        if (hop_tensor_map(&tm[0], prm.in, prm.ld_in, width, prm.n_streams, v.ops.S, v.ops.hop_tile) &&
            hop_tensor_map(&tm[1], prm.out, prm.ld_out, width, prm.n_streams, v.ops.S, v.ops.hop_tile)) {
            // pageable source: staged before the call returns; stream order puts it ahead of this launch and behind the previous one
            FE_CUDA(cudaMemcpyAsync(tmaps_device, tm, sizeof(tm), cudaMemcpyHostToDevice, st));
            prm.hop_tma = 1;
            prm.tmaps = tmaps_device;
        }
    }
    FE_CUDA(v.ops.launch(prm, grid, st));
    e->launches.fetch_add(1, std::memory_order_relaxed);
    return FE_OK;
}

cudaError_t pool_alloc(fe_engine* e, float** p, size_t floats, cudaStream_t st) {
    return e->pool ? cudaMallocFromPoolAsync((void**)p, floats * sizeof(float), e->pool, st) : cudaMallocAsync((void**)p, floats * sizeof(float), st);
}

// ---- Model.forward on few long utterances: the frame-parallel schedule ----
// float offsets of block k's GRU tensors in the canonical array (fe_pack.h::Canon)
struct GruOff { size_t w_hh, b_ih, b_hh; };
GruOff gru_offsets(const fe_config& c, int k) {
    const size_t C1 = c.c1, C2 = c.c2, F1 = c.n_fft / 8, F2 = c.f2;
    size_t o = C1 * 16 + C1 + (size_t)c.n_enc * (C1 * C1 * 3 + C1) + F2 * F1 + C2 * C1 + C2;
    for (int b = 0; b < k; ++b)
        o += 2 * 3 * C2 * C2 + 2 * 3 * C2 + C2 * C2 + C2 + (b == 0 ? F2 * C2 : 0) + 3 * C2 * C2 + 3 * C2 + C2 * C2 + C2;
    return GruOff{o + 3 * C2 * C2, o + 6 * C2 * C2, o + 6 * C2 * C2 + 3 * C2};
}
template <int C2, int Q> void scan_launch(const float* gx, float* h, const float* canon, const GruOff& g, int rows, int T, int F2, cudaStream_t st) {
    fe_gru_scan_kernel<C2, Q><<<rows, ((C2 * Q + 31) / 32) * 32, 0, st>>>(gx, h, canon + g.w_hh, canon + g.b_ih, canon + g.b_hh, T, F2);
}
int gru_scan(const fe_config& c, int k, const float* gx, float* h, const float* canon, int B, int T, cudaStream_t st) {
    const GruOff g = gru_offsets(c, k);
    const int rows = B * c.f2;
    switch (c.c2) {
        case 20: scan_launch<20, 1>(gx, h, canon, g, rows, T, c.f2, st); break;
        case 36: scan_launch<36, 2>(gx, h, canon, g, rows, T, c.f2, st); break;
        case 48: scan_launch<48, 2>(gx, h, canon, g, rows, T, c.f2, st); break;
        case 72: scan_launch<72, 4>(gx, h, canon, g, rows, T, c.f2, st); break;
        case 96: scan_launch<96, 4>(gx, h, canon, g, rows, T, c.f2, st); break;
        default: return fail(FE_ERR_UNSUPPORTED, "fe_offline: no GRU scan kernel for this rf_channels");
    }
    FE_CUDA(cudaGetLastError());
    return FE_OK;
}

// the fp32-family variant the frame-parallel schedule runs on: the largest S that still gives every SM a frame group, else the smallest
int pick_tp_variant(const fe_engine* e, long n_frames) {
    int first = -1;
    if (e->forced_s > 0)
        for (size_t i = 0; i < e->variants.size(); ++i)
            if (e->variants[i].ops.tc == 0 && e->variants[i].ops.S == e->forced_s) return (int)i;
    for (int i = (int)e->variants.size() - 1; i >= 0; --i) {
        if (e->variants[i].ops.tc != 0) continue;
        first = i;
        if ((n_frames + e->variants[i].ops.S - 1) / e->variants[i].ops.S >= e->num_sms) return i;
    }
    return first;
}

int offline_tp(fe_engine* e, const float* wav, int B, int L, float* wav_out, float* spec_out, cudaStream_t st) {
    const fe_config& c = e->cfg;
    const int T = 1 + L / c.hop;
    const long nf = (long)B * T;
    const int vi = pick_tp_variant(e, nf);
    if (vi < 0) return fail(FE_ERR_UNSUPPORTED, "fe_offline: this model has no fp32-family kernel variant for the frame-parallel schedule");
    if (int rc = ensure_variant(e, vi)) return rc;
    const Variant& v = e->variants[vi];
    {
        std::lock_guard<std::mutex> lk(e->mu);
        if (!e->canon_dev) {
            FE_CUDA(cudaMalloc(&e->canon_dev, e->canonical.size() * sizeof(float)));
            FE_CUDA(cudaMemcpy(e->canon_dev, e->canonical.data(), e->canonical.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
    }
    const int ngroups = (int)((nf + v.ops.S - 1) / v.ops.S), grid = std::min(ngroups, e->num_sms);
    const size_t n_scr = (size_t)ngroups * v.ops.tp_group, n_gx = (size_t)nf * c.f2 * 3 * c.c2, n_h = (size_t)nf * c.f2 * c.c2, n_fr = (size_t)nf * c.n_fft;
    float* buf = nullptr;
    FE_CUDA(pool_alloc(e, &buf, n_scr + n_gx + n_h + n_fr, st));
    fe::KParams prm{};
    prm.blob = v.blob; prm.in = wav; prm.out = wav_out; prm.spec_out = spec_out; prm.compression = c.compression; prm.prof = e->prof;
    prm.n_streams = B; prm.n_hops = T; prm.L = L; prm.mode = fe::MODE_OFFLINE; prm.dbg_hop = -1;
    prm.tp_scr = buf; prm.tp_gx = buf + n_scr; prm.tp_h = buf + n_scr + n_gx; prm.tp_frames = buf + n_scr + n_gx + n_h;
    int rc = FE_OK;
    cudaError_t ce = cudaSuccess;
    // FE_TP_TIMING=1 (debugging aid): CUDA-event time of every launch of the schedule on stderr
    static const bool timing = std::getenv("FE_TP_TIMING") && std::atoi(std::getenv("FE_TP_TIMING")) != 0;
    std::vector<cudaEvent_t> evs;
    auto mark = [&]() { if (timing) { cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, st); evs.push_back(ev); } };
    auto stage = [&](int stg, int blk) {
        if (rc != FE_OK || ce != cudaSuccess) return;
        prm.tp_stage = stg; prm.tp_blk = blk;
        ce = v.ops.launch(prm, grid, st);
        e->launches.fetch_add(1, std::memory_order_relaxed);
        mark();
    };
    mark();
    stage(1, 0);
    for (int k = 0; k < c.n_blocks; ++k) {
        if (rc == FE_OK && ce == cudaSuccess) { rc = gru_scan(c, k, prm.tp_gx, buf + n_scr + n_gx, e->canon_dev, B, T, st); e->launches.fetch_add(1, std::memory_order_relaxed); }
        mark();
        stage(2, k);
    }
    if (rc == FE_OK && ce == cudaSuccess) {
        const long total = (long)B * c.hop * (T - 1);
        fe_overlap_add_kernel<<<(int)std::min<long>((total + 255) / 256, 4096), 256, 0, st>>>(prm.tp_frames, v.blob + v.ops.aux_window_sq, wav_out, B, T,
                                                                                                c.n_fft, c.hop);
        ce = cudaGetLastError();
        e->launches.fetch_add(1, std::memory_order_relaxed);
    }
    cudaFreeAsync(buf, st);
    if (timing && !evs.empty()) {
        mark();
        cudaEventSynchronize(evs.back());
        std::string line = "fe_offline frame-parallel, ms per launch [stage A | scan, stage B per block | overlap-add]:";
        for (size_t i = 1; i < evs.size(); ++i) { float ms = 0.f; cudaEventElapsedTime(&ms, evs[i - 1], evs[i]); line += " " + std::to_string(ms); }
        std::fprintf(stderr, "%s\n", line.c_str());
        for (cudaEvent_t ev : evs) cudaEventDestroy(ev);
    }
    if (rc == FE_OK && ce != cudaSuccess) rc = cuda_fail(ce, "fe_offline (frame-parallel schedule)");
    return rc;
}

}  // namespace

#define FE_API __attribute__((visibility("default")))
extern "C" {

FE_API const char* fe_last_error(void) { return g_err.c_str(); }

FE_API size_t fe_weight_count(const fe_config* cfg) { return cfg ? weight_count(*cfg) : 0; }
FE_API size_t fe_state_floats(const fe_config* cfg) {
    return cfg ? 2 * (size_t)(cfg->n_fft - cfg->hop) + (size_t)cfg->n_blocks * cfg->f2 * cfg->c2 : 0;
}

FE_API int fe_create(const fe_config* cfg, const float* canonical, size_t n_floats, int device, fe_engine** out) {
    if (!cfg || !canonical || !out) return fail(FE_ERR_ARG, "fe_create: null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(FE_ERR_NO_DEVICE, "fe_create: no CUDA device visible (the engine has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(FE_ERR_ARG, "fe_create: bad device index");
    std::vector<Variant> vs;
    for (const fe::VariantOps& o : all_variants())
        if (same_shape(o.shape, *cfg)) { Variant v; v.ops = o; vs.push_back(v); }
    if (vs.empty())
        return fail(FE_ERR_UNSUPPORTED, "fe_create: model shape is not one of the shipped FastEnhancer configurations (T/B/S/M/L at 16 or 48 kHz)");
    if (n_floats != weight_count(*cfg)) return fail(FE_ERR_ARG, "fe_create: canonical weight array has the wrong length");
    if (!(cfg->compression > 0.f)) return fail(FE_ERR_ARG, "fe_create: compression must be positive");
    std::sort(vs.begin(), vs.end(), [](const Variant& a, const Variant& b) {
        return a.ops.tc != b.ops.tc ? a.ops.tc < b.ops.tc : a.ops.S < b.ops.S; });
    FE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    FE_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(FE_ERR_UNSUPPORTED, "fe_create: kernels are built for sm_100a (B200) only");
    fe_engine* e = new fe_engine();
    e->cfg = *cfg; e->device = device; e->num_sms = prop.multiProcessorCount;
    e->canonical.assign(canonical, canonical + n_floats);
    e->variants = std::move(vs);
    {
        // (the default pool hands its memory back to the driver at every synchronisation: a repeated Model.forward would pay the
        // allocation of its scratch again and again)
        cudaMemPoolProps pp{};
        pp.allocType = cudaMemAllocationTypePinned; pp.handleTypes = cudaMemHandleTypeNone;
        pp.location.type = cudaMemLocationTypeDevice; pp.location.id = device;
        if (cudaMemPoolCreate(&e->pool, &pp) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(e->pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else { e->pool = nullptr; cudaGetLastError(); }
    }
    if (const char* env = std::getenv("FE_STREAMS_PER_CTA")) e->forced_s = std::atoi(env);
    if (const char* env = std::getenv("FE_HOP_TMA")) e->hop_tma = std::atoi(env);
    if (const char* env = std::getenv("FE_HOP_SLICING")) e->hop_slicing = std::atoi(env);
    if (const char* env = std::getenv("FE_OFFLINE_MODE")) e->offline_mode = std::max(0, std::min(2, std::atoi(env)));
    // Default arithmetic = results identical to the fp32 reference: the fp32-accurate tensor-core family where the model has one
    // (split-fp16 operands, three MMAs per product), else the fp32 FMA pipe.  The faster reduced-precision families are opt-in.
    e->tc = 0;
    for (const Variant& v : e->variants) if (v.ops.tc == 4) e->tc = 4;
    if (const char* env = std::getenv("FE_PRECISION")) {
        const char* names[] = {"fp32", "tf32", "f16", "bf16", "fp32x3"};
        for (int want = 0; want < 5; ++want)
            if (std::strcmp(env, names[want]) == 0)
                for (const Variant& v : e->variants) if (v.ops.tc == want) e->tc = want;     // only if this model has such variants
    }
    *out = e;
    return FE_OK;
}

// Checkpoint ingestion on the device: `ops` (host array) lists one fold rule per tensor of the canonical array; all sources are device
// pointers (the reference's pre-fold parameters, e.g. torch tensors of ckpt['model'] moved to the GPU).  Replaces the host-side
// remove_weight_reparameterizations() call of scripts/export_onnx.py:78.
FE_API int fe_fold_device(const fe_fold_op* ops, int n_ops, float* canonical_device, void* cuda_stream) {
    if (!ops || n_ops <= 0 || !canonical_device) return fail(FE_ERR_ARG, "fe_fold_device: bad argument");
    for (int i = 0; i < n_ops; ++i) {
        const fe_fold_op& op = ops[i];
        if (op.kind < FE_FOLD_COPY || op.kind > FE_FOLD_FINAL_CONV || op.rows <= 0 || op.cols <= 0 || !op.w || op.dst < 0)
            return fail(FE_ERR_ARG, "fe_fold_device: malformed rule " + std::to_string(i));
        if ((op.kind == FE_FOLD_WEIGHT_NORM || op.kind == FE_FOLD_FINAL_CONV) && !op.a) return fail(FE_ERR_ARG, "fe_fold_device: missing gain / scale");
        if (op.kind == FE_FOLD_BATCH_NORM && (!op.a || !op.b || !op.c || !op.d)) return fail(FE_ERR_ARG, "fe_fold_device: missing BatchNorm statistics");
        const int grid = (op.kind == FE_FOLD_COPY) ? (int)std::min<long>(((long)op.rows * op.cols + 127) / 128, 1024) : (op.kind == FE_FOLD_FINAL_CONV ? 1 : op.rows);
        fold_kernel<<<grid, 128, 0, (cudaStream_t)cuda_stream>>>(op, canonical_device);
        FE_CUDA(cudaGetLastError());
    }
    return FE_OK;
}

// fe_create on a canonical array that already lives on `device` (the output of fe_fold_device).
FE_API int fe_create_from_device(const fe_config* cfg, const float* canonical_device, size_t n_floats, int device, fe_engine** out) {
    if (!cfg || !canonical_device || !out) return fail(FE_ERR_ARG, "fe_create_from_device: null argument");
    if (n_floats != weight_count(*cfg)) return fail(FE_ERR_ARG, "fe_create_from_device: canonical weight array has the wrong length");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(FE_ERR_NO_DEVICE, "fe_create_from_device: no CUDA device visible (the engine has no CPU fallback)");
    FE_CUDA(cudaSetDevice(device));
    std::vector<float> host(n_floats);        // the packer (fe_pack.h) lays the blob out on the host, once per kernel variant
    FE_CUDA(cudaMemcpy(host.data(), canonical_device, n_floats * sizeof(float), cudaMemcpyDeviceToHost));
    return fe_create(cfg, host.data(), n_floats, device, out);
}

FE_API void fe_destroy(fe_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (Variant& v : e->variants) if (v.blob) cudaFree(v.blob);
    if (e->canon_dev) cudaFree(e->canon_dev);
    if (e->basis_dev) cudaFree(e->basis_dev);
    if (e->pool) cudaMemPoolDestroy(e->pool);
    delete e;
}

static int state_create(fe_engine* e, int n_streams, float* external, fe_state** out) {
    if (!e || !out || n_streams <= 0) return fail(FE_ERR_ARG, "fe_state_create: bad argument");
    *out = nullptr;
    FE_CUDA(cudaSetDevice(e->device));
    fe_state* s = new fe_state();
    s->e = e; s->n_streams = n_streams;
    const size_t sf = fe_state_floats(&e->cfg);
    cudaError_t ce = cudaSuccess;
    if (external) { s->data = external; s->owns_data = false; }
    else {
        ce = cudaMalloc(&s->data, (size_t)n_streams * sf * sizeof(float));
        if (ce == cudaSuccess) ce = cudaMemset(s->data, 0, (size_t)n_streams * sf * sizeof(float));
    }
    s->scratch_floats = scratch_need(e, n_streams);
    if (ce == cudaSuccess) ce = cudaMalloc(&s->scratch, s->scratch_floats * sizeof(float));
    if (ce == cudaSuccess) ce = cudaMalloc(&s->tmaps, 2 * sizeof(CUtensorMap));
    if (ce != cudaSuccess) { fe_state_destroy(s); return cuda_fail(ce, "fe_state_create"); }
    *out = s;
    return FE_OK;
}
FE_API int fe_state_create(fe_engine* e, int n_streams, fe_state** out) { return state_create(e, n_streams, nullptr, out); }
FE_API int fe_state_create_on(fe_engine* e, int n_streams, float* planes_device, fe_state** out) {
    if (!planes_device) return fail(FE_ERR_ARG, "fe_state_create_on: null buffer");
    if ((reinterpret_cast<size_t>(planes_device) & 15) != 0) return fail(FE_ERR_ARG, "fe_state_create_on: buffer must be 16-byte aligned");
    return state_create(e, n_streams, planes_device, out);
}
FE_API float* fe_state_planes(fe_state* s) { return s ? s->data : nullptr; }

FE_API void fe_state_destroy(fe_state* s) {
    if (!s) return;
    cudaSetDevice(s->e->device);
    if (s->data && s->owns_data) cudaFree(s->data);
    if (s->scratch) cudaFree(s->scratch);
    if (s->tmaps) cudaFree(s->tmaps);
    if (s->flags) cudaFree(s->flags);
    for (int i = 0; i < 2; ++i) {
        if (s->h_in[i]) cudaFree(s->h_in[i]);
        if (s->h_out[i]) cudaFree(s->h_out[i]);
        if (s->ev_in[i]) cudaEventDestroy(s->ev_in[i]);
        if (s->ev_k[i]) cudaEventDestroy(s->ev_k[i]);
        if (s->ev_out[i]) cudaEventDestroy(s->ev_out[i]);
    }
    if (s->ev_call) cudaEventDestroy(s->ev_call);
    if (s->s_copy_in) cudaStreamDestroy(s->s_copy_in);
    if (s->s_compute) cudaStreamDestroy(s->s_compute);
    if (s->s_copy_out) cudaStreamDestroy(s->s_copy_out);
    delete s;
}

FE_API int fe_state_reset(fe_state* s, void* cuda_stream) {
    if (!s) return fail(FE_ERR_ARG, "fe_state_reset: null state");
    FE_CUDA(cudaMemsetAsync(s->data, 0, (size_t)s->n_streams * fe_state_floats(&s->e->cfg) * sizeof(float), (cudaStream_t)cuda_stream));
    return FE_OK;
}

static int state_xpose(fe_state* s, const float* src, float* dst, int to_planes, void* cuda_stream) {
    const fe_config& c = s->e->cfg;
    const long total = (long)s->n_streams * (long)fe_state_floats(&c);
    const int blocks = (int)std::min<long>((total + 255) / 256, 4096);
    state_transpose_kernel<<<blocks, 256, 0, (cudaStream_t)cuda_stream>>>(src, dst, s->n_streams, c.n_fft - c.hop, c.n_blocks, c.f2 * c.c2, to_planes);
    FE_CUDA(cudaGetLastError());
    return FE_OK;
}
FE_API int fe_state_export(fe_state* s, float* dst_device, void* cuda_stream) {
    if (!s || !dst_device) return fail(FE_ERR_ARG, "fe_state_export: null argument");
    return state_xpose(s, s->data, dst_device, 0, cuda_stream);
}
FE_API int fe_state_import(fe_state* s, const float* src_device, void* cuda_stream) {
    if (!s || !src_device) return fail(FE_ERR_ARG, "fe_state_import: null argument");
    return state_xpose(s, src_device, s->data, 1, cuda_stream);
}

FE_API int fe_stream_taps(fe_engine* e, fe_state* s, const float* wav_in, float* wav_out, int n_hops, long long ld_in,
                   long long ld_out, float* taps_device, int tap_hop, void* cuda_stream) {
    if (!e || !s || s->e != e || !wav_in || !wav_out) return fail(FE_ERR_ARG, "fe_stream: null / mismatched argument");
    if (n_hops < 0 || ld_in < (long long)n_hops * e->cfg.hop || ld_out < (long long)n_hops * e->cfg.hop)
        return fail(FE_ERR_ARG, "fe_stream: leading dimension smaller than n_hops*hop");
    FE_CUDA(cudaSetDevice(e->device));
    fe::KParams prm{};
    prm.state = s->data; prm.in = wav_in; prm.out = wav_out; prm.ld_in = ld_in; prm.ld_out = ld_out;
    prm.n_streams = s->n_streams; prm.n_hops = n_hops; prm.mode = fe::MODE_STREAM;
    prm.dbg = taps_device; prm.dbg_hop = tap_hop;
    return launch(e, prm, s->scratch, (cudaStream_t)cuda_stream, s->tmaps, s);
}

FE_API int fe_stream(fe_engine* e, fe_state* s, const float* wav_in, float* wav_out, int n_hops, long long ld_in,
              long long ld_out, void* cuda_stream) {
    return fe_stream_taps(e, s, wav_in, wav_out, n_hops, ld_in, ld_out, nullptr, -1, cuda_stream);
}

FE_API int fe_spec(fe_engine* e, fe_state* s, const float* spec_in, float* spec_out, int T, void* cuda_stream) {
    if (!e || !s || s->e != e || !spec_in || !spec_out || T < 0) return fail(FE_ERR_ARG, "fe_spec: bad argument");
    FE_CUDA(cudaSetDevice(e->device));
    fe::KParams prm{};
    prm.state = s->data; prm.in = spec_in; prm.out = spec_out;
    prm.n_streams = s->n_streams; prm.n_hops = T; prm.mode = fe::MODE_SPEC; prm.dbg_hop = -1;
    return launch(e, prm, s->scratch, (cudaStream_t)cuda_stream);
}

FE_API int fe_stft(fe_engine* e, fe_state* s, const float* wav_in, float* spec_out, int n_hops, long long ld_in, void* cuda_stream) {
    if (!e || !s || s->e != e || !wav_in || !spec_out) return fail(FE_ERR_ARG, "fe_stft: null / mismatched argument");
    if (n_hops < 0 || ld_in < (long long)n_hops * e->cfg.hop) return fail(FE_ERR_ARG, "fe_stft: leading dimension smaller than n_hops*hop");
    FE_CUDA(cudaSetDevice(e->device));
    fe::KParams prm{};
    prm.state = s->data; prm.in = wav_in; prm.out = spec_out; prm.ld_in = ld_in;
    prm.n_streams = s->n_streams; prm.n_hops = n_hops; prm.mode = fe::MODE_STFT; prm.dbg_hop = -1;
    return launch(e, prm, s->scratch, (cudaStream_t)cuda_stream);
}

FE_API int fe_istft(fe_engine* e, fe_state* s, const float* spec_in, float* wav_out, int n_hops, long long ld_out, void* cuda_stream) {
    if (!e || !s || s->e != e || !spec_in || !wav_out) return fail(FE_ERR_ARG, "fe_istft: null / mismatched argument");
    if (n_hops < 0 || ld_out < (long long)n_hops * e->cfg.hop) return fail(FE_ERR_ARG, "fe_istft: leading dimension smaller than n_hops*hop");
    FE_CUDA(cudaSetDevice(e->device));
    fe::KParams prm{};
    prm.state = s->data; prm.in = spec_in; prm.out = wav_out; prm.ld_out = ld_out;
    prm.n_streams = s->n_streams; prm.n_hops = n_hops; prm.mode = fe::MODE_ISTFT; prm.dbg_hop = -1;
    return launch(e, prm, s->scratch, (cudaStream_t)cuda_stream);
}

FE_API int fe_offline(fe_engine* e, const float* wav, int B, int L, float* wav_out, float* spec_out, void* cuda_stream) {
    if (!e || !wav || !wav_out || B <= 0) return fail(FE_ERR_ARG, "fe_offline: bad argument");
    if (L <= e->cfg.n_fft / 2) return fail(FE_ERR_ARG, "fe_offline: input shorter than n_fft/2 + 1 samples (reflect padding needs more)");
    FE_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    // Few long utterances leave the one-CTA-per-stream-group walk with most SMs idle: outside the GRU recurrence the frames of an utterance
    // are independent, so the fp32-accurate families switch to the frame-parallel schedule (same kernels, same weights, frames instead of
    // streams in the CTA slots; the recurrence runs as a scan between the launches).  Automatic when the walk would fill less than half
    // of the SMs and the scratch (about 100 KB per frame) stays below 2 GB.
    {
        const long T = 1 + L / e->cfg.hop, nf = (long)B * T;
        const int vi = pick_tp_variant(e, nf);
        bool tp = false;
        if (vi >= 0 && T >= 2 && (e->tc == 0 || e->tc == 4)) {
            const int S_walk = e->variants[pick_variant(e, B)].ops.S;
            const double scr_bytes = (double)((nf + e->variants[vi].ops.S - 1) / e->variants[vi].ops.S) * e->variants[vi].ops.tp_group * 4.0;
            tp = e->offline_mode == 2 || (e->offline_mode == 0 && (B + S_walk - 1) / S_walk * 2 <= e->num_sms && scr_bytes < 2e9);
        }
        if (e->offline_mode == 2 && !tp) return fail(FE_ERR_UNSUPPORTED, "fe_offline: the frame-parallel schedule needs an fp32-accurate precision mode");
        if (tp) return offline_tp(e, wav, B, L, wav_out, spec_out, st);
    }
    // zero state + spill scratch of THIS call, allocated and freed in stream order: concurrent calls on other streams share nothing
    const size_t sf = fe_state_floats(&e->cfg), need = (size_t)B * sf, sneed = scratch_need(e, B);
    float *off_state = nullptr, *off_scratch = nullptr;
    FE_CUDA(pool_alloc(e, &off_state, need, st));
    cudaError_t ce = pool_alloc(e, &off_scratch, sneed, st);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(off_state, 0, need * sizeof(float), st);
    int rc = FE_OK;
    if (ce != cudaSuccess) rc = cuda_fail(ce, "fe_offline: scratch allocation");
    else {
        fe::KParams prm{};
        prm.state = off_state; prm.in = wav; prm.out = wav_out; prm.spec_out = spec_out;
        prm.n_streams = B; prm.n_hops = 1 + L / e->cfg.hop; prm.L = L; prm.mode = fe::MODE_OFFLINE; prm.dbg_hop = -1;
        rc = launch(e, prm, off_scratch, st);
    }
    cudaFreeAsync(off_state, st);
    if (off_scratch) cudaFreeAsync(off_scratch, st);
    return rc;
}

// Staging buffers / streams / events of the pipelined host path for pieces of up to `hops_per_chunk` hops: call it once up front to
// keep every allocation out of fe_stream_host (which otherwise reserves on first use / when a larger piece is asked for).
FE_API int fe_state_reserve_host(fe_state* s, int hops_per_chunk) {
    if (!s || hops_per_chunk <= 0) return fail(FE_ERR_ARG, "fe_state_reserve_host: bad argument");
    FE_CUDA(cudaSetDevice(s->e->device));
    if (!s->s_compute) {
        FE_CUDA(cudaStreamCreateWithFlags(&s->s_copy_in, cudaStreamNonBlocking));
        FE_CUDA(cudaStreamCreateWithFlags(&s->s_compute, cudaStreamNonBlocking));
        FE_CUDA(cudaStreamCreateWithFlags(&s->s_copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            FE_CUDA(cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming));
            FE_CUDA(cudaEventCreateWithFlags(&s->ev_k[i], cudaEventDisableTiming));
            FE_CUDA(cudaEventCreateWithFlags(&s->ev_out[i], cudaEventDisableTiming));
        }
        FE_CUDA(cudaEventCreateWithFlags(&s->ev_call, cudaEventDisableTiming));
    }
    const size_t need = (size_t)s->n_streams * hops_per_chunk * s->e->cfg.hop;
    if (s->h_floats < need) {
        FE_CUDA(cudaStreamSynchronize(s->s_copy_out));
        FE_CUDA(cudaStreamSynchronize(s->s_compute));
        for (int i = 0; i < 2; ++i) {
            if (s->h_in[i]) FE_CUDA(cudaFree(s->h_in[i]));
            if (s->h_out[i]) FE_CUDA(cudaFree(s->h_out[i]));
            s->h_in[i] = s->h_out[i] = nullptr;
        }
        s->h_floats = 0;
        for (int i = 0; i < 2; ++i) {
            FE_CUDA(cudaMalloc(&s->h_in[i], need * sizeof(float)));
            FE_CUDA(cudaMalloc(&s->h_out[i], need * sizeof(float)));
        }
        s->h_floats = need;
    }
    return FE_OK;
}

// Host buffers: [copy-in | kernel | copy-out] pipelined over `hops_per_chunk`-hop pieces on three streams with
// double-buffered device staging; the GRU / overlap state carries from piece to piece in fe_state.
// Ordered after the work already queued on `cuda_stream` (a state reset / import issued there); returns when the last copy-out
// has completed.  On an error every copy already in flight into the caller's buffers is drained before returning.
FE_API int fe_stream_host(fe_engine* e, fe_state* s, const float* wav_in_host, float* wav_out_host, int n_hops, long long ld_in,
                   long long ld_out, int hops_per_chunk, void* cuda_stream) {
    if (!e || !s || s->e != e || !wav_in_host || !wav_out_host) return fail(FE_ERR_ARG, "fe_stream_host: null / mismatched argument");
    const int H = e->cfg.hop, B = s->n_streams;
    if (n_hops < 0 || ld_in < (long long)n_hops * H || ld_out < (long long)n_hops * H)
        return fail(FE_ERR_ARG, "fe_stream_host: leading dimension smaller than n_hops*hop");
    if (n_hops == 0) return FE_OK;
    FE_CUDA(cudaSetDevice(e->device));
    if (hops_per_chunk <= 0) hops_per_chunk = 64;
    hops_per_chunk = std::min(hops_per_chunk, n_hops);
    if (int rc = fe_state_reserve_host(s, hops_per_chunk)) return rc;
    FE_CUDA(cudaEventRecord(s->ev_call, (cudaStream_t)cuda_stream));
    FE_CUDA(cudaStreamWaitEvent(s->s_compute, s->ev_call, 0));
    int rc = FE_OK;
    cudaError_t ce = cudaSuccess;
#define FE_TRY(call) do { if (rc == FE_OK && ce == cudaSuccess) { ce = (call); if (ce != cudaSuccess) rc = cuda_fail(ce, #call); } } while (0)
    int piece = 0;
    for (int h0 = 0; h0 < n_hops && rc == FE_OK; h0 += hops_per_chunk, ++piece) {
        const int nh = std::min(hops_per_chunk, n_hops - h0), b = piece & 1;
        const size_t w = (size_t)nh * H;
        if (piece >= 2) {   // staging buffers b are free once piece-2 finished its kernel (in) / its copy-out (out)
            FE_TRY(cudaStreamWaitEvent(s->s_copy_in, s->ev_k[b], 0));
            FE_TRY(cudaStreamWaitEvent(s->s_compute, s->ev_out[b], 0));
        }
        FE_TRY(cudaMemcpy2DAsync(s->h_in[b], w * sizeof(float), wav_in_host + (size_t)h0 * H, (size_t)ld_in * sizeof(float),
                                 w * sizeof(float), B, cudaMemcpyHostToDevice, s->s_copy_in));
        FE_TRY(cudaEventRecord(s->ev_in[b], s->s_copy_in));
        FE_TRY(cudaStreamWaitEvent(s->s_compute, s->ev_in[b], 0));
        if (rc == FE_OK) rc = fe_stream(e, s, s->h_in[b], s->h_out[b], nh, (long long)w, (long long)w, s->s_compute);
        FE_TRY(cudaEventRecord(s->ev_k[b], s->s_compute));
        FE_TRY(cudaStreamWaitEvent(s->s_copy_out, s->ev_k[b], 0));
        FE_TRY(cudaMemcpy2DAsync(wav_out_host + (size_t)h0 * H, (size_t)ld_out * sizeof(float), s->h_out[b], w * sizeof(float),
                                 w * sizeof(float), B, cudaMemcpyDeviceToHost, s->s_copy_out));
        FE_TRY(cudaEventRecord(s->ev_out[b], s->s_copy_out));
    }
#undef FE_TRY
    // drain (also on the error path: nothing may still be writing into the caller's buffers once we return)
    const std::string err_keep = g_err;
    cudaError_t c1 = cudaStreamSynchronize(s->s_copy_in), c2 = cudaStreamSynchronize(s->s_compute), c3 = cudaStreamSynchronize(s->s_copy_out);
    if (rc != FE_OK) { g_err = err_keep; return rc; }
    if (c1 != cudaSuccess) return cuda_fail(c1, "fe_stream_host: copy-in stream");
    if (c2 != cudaSuccess) return cuda_fail(c2, "fe_stream_host: compute stream");
    if (c3 != cudaSuccess) return cuda_fail(c3, "fe_stream_host: copy-out stream");
    return FE_OK;
}

FE_API int fe_streams_per_cta(fe_engine* e, int n_streams) {
    if (!e || n_streams <= 0) return fail(FE_ERR_ARG, "fe_streams_per_cta: bad argument");
    return e->variants[pick_variant(e, n_streams)].ops.S;
}
FE_API int fe_set_streams_per_cta(fe_engine* e, int s) {
    if (!e) return fail(FE_ERR_ARG, "fe_set_streams_per_cta: null engine");
    if (s != 0) {
        bool ok = false;
        for (const Variant& v : e->variants) ok = ok || (v.ops.S == s && v.ops.tc == e->tc);
        if (!ok) return fail(FE_ERR_UNSUPPORTED, "fe_set_streams_per_cta: no such variant for this model");
    }
    e->forced_s = s;
    return FE_OK;
}
FE_API int fe_set_precision(fe_engine* e, int mode) {
    if (!e) return fail(FE_ERR_ARG, "fe_set_precision: null engine");
    if (mode < 0 || mode > 4) return fail(FE_ERR_ARG, "fe_set_precision: mode must be 0 (tf32), 1 (fp32), 2 (f16), 3 (bf16) or 4 (fp32x3)");
    const int tc = mode == 1 ? 0 : (mode == 0 ? 1 : mode);
    bool ok = false;
    for (const Variant& v : e->variants) ok = ok || (v.ops.tc == tc && (e->forced_s == 0 || v.ops.S == e->forced_s));
    if (!ok) return fail(FE_ERR_UNSUPPORTED, "fe_set_precision: this model has no kernel variant for the requested mode");
    e->tc = tc;
    return FE_OK;
}
// The STFT of many frames at once as a tensor-core GEMM (fe_stft_gemm.cu): frame t of utterance b = wav[b][t*hop .. t*hop + n_fft - 1]
// (no centering and no cache: a streaming caller prepends its n_fft - hop cached samples), periodic Hann window of the model.
FE_API int fe_stft_gemm(fe_engine* e, const float* wav, int B, long long ld, int T, float* spec_out, int accurate, void* cuda_stream) {
    if (!e || !wav || !spec_out || B <= 0 || T <= 0) return fail(FE_ERR_ARG, "fe_stft_gemm: bad argument");
    const int N = e->cfg.n_fft, H = e->cfg.hop;
    if (ld < (long long)(T - 1) * H + N) return fail(FE_ERR_ARG, "fe_stft_gemm: leading dimension smaller than (T-1)*hop + n_fft");
    FE_CUDA(cudaSetDevice(e->device));
    {
        std::lock_guard<std::mutex> lk(e->mu);
        if (!e->basis_dev) {
            std::vector<float> win(N), hi, lo;
            for (int i = 0; i < N; ++i) win[i] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * i / N));       // periodic Hann, as fe_pack.h
            fe::stft_gemm_basis(N, win.data(), hi, lo);
            FE_CUDA(cudaMalloc(&e->basis_dev, 2 * hi.size() * sizeof(float)));
            FE_CUDA(cudaMemcpy(e->basis_dev, hi.data(), hi.size() * sizeof(float), cudaMemcpyHostToDevice));
            FE_CUDA(cudaMemcpy(e->basis_dev + hi.size(), lo.data(), lo.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
    }
    cudaError_t ce = cudaSuccess;
    const int rc = fe::stft_gemm_launch(wav, ld, B, T, N, H, e->basis_dev, e->basis_dev + (size_t)N * N, spec_out, accurate, (cudaStream_t)cuda_stream, &ce);
    if (rc == 1) return fail(FE_ERR_ARG, "fe_stft_gemm: wav must be 16-byte aligned with ld % 4 == 0 (TMA tensor map over the waveform)");
    if (rc == 2) return fail(FE_ERR_UNSUPPORTED, "fe_stft_gemm: cuTensorMapEncodeTiled rejected the tensor map");
    if (rc == 3) return cuda_fail(ce, "fe_stft_gemm");
    return FE_OK;
}

// The slicing decision on its own (pure host arithmetic, no device needed): hops per range for a streaming launch of `n_groups` stream
// groups x `n_hops` hops on `num_sms` SMs, 0 = do not slice.
FE_API int fe_plan_hop_slices(int n_groups, int n_hops, int num_sms) {
    if (n_groups <= 0 || n_hops <= 0 || num_sms <= 0) return 0;
    return plan_slices(n_groups, n_hops, num_sms);
}
FE_API int fe_set_hop_slicing(fe_engine* e, int on) {
    if (!e) return fail(FE_ERR_ARG, "fe_set_hop_slicing: null engine");
    e->hop_slicing = on ? 1 : 0;
    return FE_OK;
}
FE_API int fe_set_offline_mode(fe_engine* e, int mode) {
    if (!e || mode < 0 || mode > 2) return fail(FE_ERR_ARG, "fe_set_offline_mode: mode must be 0 (automatic), 1 (sequential walk) or 2 (frame-parallel)");
    e->offline_mode = mode;
    return FE_OK;
}
FE_API int fe_get_precision(fe_engine* e) {
    return e ? (e->tc == 0 ? 1 : (e->tc == 1 ? 0 : e->tc)) : fail(FE_ERR_ARG, "fe_get_precision: null engine");
}
// Measured fp32 FMA throughput of `device` in TFLOP/s (2 FLOP per FFMA): the roofline denominator bench.py uses for the fp32 FMA-pipe
// kernel family instead of the nominal 148 SM x 128 lanes x 2 x clock.
FE_API int fe_microbench_fma(int device, double* tflops) {
    if (!tflops) return fail(FE_ERR_ARG, "fe_microbench_fma: null argument");
    FE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    FE_CUDA(cudaGetDeviceProperties(&prop, device));
    float* out = nullptr;
    FE_CUDA(cudaMalloc(&out, 4));
    cudaEvent_t e0, e1;
    FE_CUDA(cudaEventCreate(&e0)); FE_CUDA(cudaEventCreate(&e1));
    const int grid = prop.multiProcessorCount * 8, iters = 1 << 16;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        FE_CUDA(cudaEventRecord(e0));
        fma_peak_kernel<<<grid, 256>>>(out, iters, 1.0000001f, 1e-7f);
        FE_CUDA(cudaEventRecord(e1));
        FE_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        FE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 16.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *tflops = best;
    return FE_OK;
}
FE_API int fe_profile_slots(void) { return (int)(fe::PH_COUNT + fe::PH_COUNT * fe::PH_NSUB); }
FE_API int fe_set_profile(fe_engine* e, long long* counters_device) {
    if (!e) return fail(FE_ERR_ARG, "fe_set_profile: null engine");
    e->prof = counters_device;
    return FE_OK;
}
FE_API long long fe_kernel_launches(fe_engine* e) { return e ? e->launches.load(std::memory_order_relaxed) : 0; }
FE_API int fe_tap_floats(fe_engine* e) { return e ? e->variants[0].ops.tap_floats : 0; }

}  // extern "C"
