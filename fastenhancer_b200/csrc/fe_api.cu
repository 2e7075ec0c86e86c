// fe_api.cu -- C ABI (include/fastenhancer_b200.h) over the fused per-hop kernel variants.
//
// Host-side responsibilities only: pick the (config, streams-per-CTA) variant, pack the canonical
// weights into the variant's blob once, own device buffers for weights / state / spill scratch,
// launch.  There is no CPU compute path: without a CUDA device every entry point fails loudly.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fastenhancer_b200.h"
#include "fe_variant.h"
#include "fe_kernel.cuh"

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

// h [K][F2][C2] (reference cache layout) <-> h [K][C2][F2] (engine layout); caches copied through.
__global__ void state_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_streams, int cl2,
                                       int K, int F2, int C2, int to_native)
{
    const int sf = cl2 + K * F2 * C2;
    const long total = (long)n_streams * sf;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int s = (int)(i / sf), r = (int)(i % sf);
        if (r < cl2) { dst[i] = src[i]; continue; }
        const int q = r - cl2, k = q / (F2 * C2), e = q % (F2 * C2);
        int f, c;
        if (to_native) { c = e / F2; f = e % F2; dst[i] = src[(long)s * sf + cl2 + k * F2 * C2 + f * C2 + c]; }
        else { f = e / C2; c = e % C2; dst[i] = src[(long)s * sf + cl2 + k * F2 * C2 + c * F2 + f]; }
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
    int tc = 1;                          // 1: conv-type contractions on tcgen05 (TF32 operands, fp32 accumulate); 0: all fp32 FMA
    long long* prof = nullptr;           // optional per-phase cycle counters (device)
    long long launches = 0;
    std::mutex mu;
    // offline-mode scratch state (zeroed before every call)
    float* off_state = nullptr; size_t off_state_floats = 0;
    float* off_scratch = nullptr; size_t off_scratch_floats = 0;
    // pipelined host path
    cudaStream_t s_copy_in = nullptr, s_compute = nullptr, s_copy_out = nullptr;
    float* h_in[2] = {nullptr, nullptr}; float* h_out[2] = {nullptr, nullptr}; size_t h_floats = 0;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
};

struct fe_state {
    fe_engine* e;
    int n_streams;
    float* data = nullptr;      // [n_streams][STATE] engine layout
    float* scratch = nullptr;   // spill scratch for the largest grid (S = 1)
    size_t scratch_floats = 0;
};

namespace {

int pick_variant(fe_engine* e, int n_streams) {
    // variants are sorted by (tc, S); only those of the engine's precision mode are eligible
    int first = -1;
    if (e->forced_s > 0) {
        for (size_t i = 0; i < e->variants.size(); ++i)
            if (e->variants[i].ops.tc == e->tc && e->variants[i].ops.S == e->forced_s) return (int)i;
    }
    // largest S that still gives most SMs a CTA; otherwise the smallest S
    for (int i = (int)e->variants.size() - 1; i >= 0; --i) {
        if (e->variants[i].ops.tc != e->tc) continue;
        first = i;
        const int S = e->variants[i].ops.S;
        if ((n_streams + S - 1) / S >= (e->num_sms * 4) / 5) return i;
    }
    return first;
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

int launch(fe_engine* e, fe::KParams prm, float* scratch, cudaStream_t st) {
    if (prm.n_streams <= 0 || prm.n_hops <= 0) return FE_OK;
    const int vi = pick_variant(e, prm.n_streams);
    int rc = ensure_variant(e, vi);
    if (rc) return rc;
    const Variant& v = e->variants[vi];
    prm.blob = v.blob;
    prm.scratch = scratch;
    prm.compression = e->cfg.compression;
    prm.prof = e->prof;
    const int grid = (prm.n_streams + v.ops.S - 1) / v.ops.S;
    FE_CUDA(v.ops.launch(prm, grid, st));
    ++e->launches;
    return FE_OK;
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
    if (const char* env = std::getenv("FE_STREAMS_PER_CTA")) e->forced_s = std::atoi(env);
    if (const char* env = std::getenv("FE_PRECISION")) {
        e->tc = std::strcmp(env, "fp32") == 0 ? 0 : 1;
        if (std::strcmp(env, "f16") == 0) {      // only if this model has fp16 variants
            for (const Variant& v : e->variants) if (v.ops.tc == 2) e->tc = 2;
        }
    }
    *out = e;
    return FE_OK;
}

FE_API void fe_destroy(fe_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (Variant& v : e->variants) if (v.blob) cudaFree(v.blob);
    if (e->off_state) cudaFree(e->off_state);
    if (e->off_scratch) cudaFree(e->off_scratch);
    for (int i = 0; i < 2; ++i) {
        if (e->h_in[i]) cudaFree(e->h_in[i]);
        if (e->h_out[i]) cudaFree(e->h_out[i]);
        if (e->ev_in[i]) cudaEventDestroy(e->ev_in[i]);
        if (e->ev_k[i]) cudaEventDestroy(e->ev_k[i]);
        if (e->ev_out[i]) cudaEventDestroy(e->ev_out[i]);
    }
    if (e->s_copy_in) cudaStreamDestroy(e->s_copy_in);
    if (e->s_compute) cudaStreamDestroy(e->s_compute);
    if (e->s_copy_out) cudaStreamDestroy(e->s_copy_out);
    delete e;
}

FE_API int fe_state_create(fe_engine* e, int n_streams, fe_state** out) {
    if (!e || !out || n_streams <= 0) return fail(FE_ERR_ARG, "fe_state_create: bad argument");
    *out = nullptr;
    FE_CUDA(cudaSetDevice(e->device));
    fe_state* s = new fe_state();
    s->e = e; s->n_streams = n_streams;
    const size_t sf = fe_state_floats(&e->cfg);
    cudaError_t ce = cudaMalloc(&s->data, (size_t)n_streams * sf * sizeof(float));
    if (ce == cudaSuccess) ce = cudaMemset(s->data, 0, (size_t)n_streams * sf * sizeof(float));
    s->scratch_floats = scratch_need(e, n_streams);
    if (ce == cudaSuccess) ce = cudaMalloc(&s->scratch, s->scratch_floats * sizeof(float));
    if (ce != cudaSuccess) { fe_state_destroy(s); return cuda_fail(ce, "fe_state_create"); }
    *out = s;
    return FE_OK;
}

FE_API void fe_state_destroy(fe_state* s) {
    if (!s) return;
    cudaSetDevice(s->e->device);
    if (s->data) cudaFree(s->data);
    if (s->scratch) cudaFree(s->scratch);
    delete s;
}

FE_API int fe_state_reset(fe_state* s, void* cuda_stream) {
    if (!s) return fail(FE_ERR_ARG, "fe_state_reset: null state");
    FE_CUDA(cudaMemsetAsync(s->data, 0, (size_t)s->n_streams * fe_state_floats(&s->e->cfg) * sizeof(float), (cudaStream_t)cuda_stream));
    return FE_OK;
}

static int state_xpose(fe_state* s, const float* src, float* dst, int to_native, void* cuda_stream) {
    const fe_config& c = s->e->cfg;
    const long total = (long)s->n_streams * (long)fe_state_floats(&c);
    const int blocks = (int)std::min<long>((total + 255) / 256, 4096);
    state_transpose_kernel<<<blocks, 256, 0, (cudaStream_t)cuda_stream>>>(src, dst, s->n_streams, 2 * (c.n_fft - c.hop), c.n_blocks, c.f2, c.c2, to_native);
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
    return launch(e, prm, s->scratch, (cudaStream_t)cuda_stream);
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
    const size_t sf = fe_state_floats(&e->cfg), need = (size_t)B * sf, sneed = scratch_need(e, B);
    if (e->off_state_floats < need) {
        if (e->off_state) FE_CUDA(cudaFree(e->off_state));
        e->off_state = nullptr; e->off_state_floats = 0;
        FE_CUDA(cudaMalloc(&e->off_state, need * sizeof(float)));
        e->off_state_floats = need;
    }
    if (e->off_scratch_floats < sneed) {
        if (e->off_scratch) FE_CUDA(cudaFree(e->off_scratch));
        e->off_scratch = nullptr; e->off_scratch_floats = 0;
        FE_CUDA(cudaMalloc(&e->off_scratch, sneed * sizeof(float)));
        e->off_scratch_floats = sneed;
    }
    FE_CUDA(cudaMemsetAsync(e->off_state, 0, need * sizeof(float), st));
    fe::KParams prm{};
    prm.state = e->off_state; prm.in = wav; prm.out = wav_out; prm.spec_out = spec_out;
    prm.n_streams = B; prm.n_hops = 1 + L / e->cfg.hop; prm.L = L; prm.mode = fe::MODE_OFFLINE; prm.dbg_hop = -1;
    return launch(e, prm, e->off_scratch, st);
}

// Host buffers: [copy-in | kernel | copy-out] pipelined over `hops_per_chunk`-hop pieces on three streams with
// double-buffered device staging; the GRU / overlap state carries from piece to piece in fe_state.
FE_API int fe_stream_host(fe_engine* e, fe_state* s, const float* wav_in_host, float* wav_out_host, int n_hops, long long ld_in,
                   long long ld_out, int hops_per_chunk) {
    if (!e || !s || s->e != e || !wav_in_host || !wav_out_host) return fail(FE_ERR_ARG, "fe_stream_host: null / mismatched argument");
    const int H = e->cfg.hop, B = s->n_streams;
    if (n_hops < 0 || ld_in < (long long)n_hops * H || ld_out < (long long)n_hops * H)
        return fail(FE_ERR_ARG, "fe_stream_host: leading dimension smaller than n_hops*hop");
    if (n_hops == 0) return FE_OK;
    FE_CUDA(cudaSetDevice(e->device));
    if (hops_per_chunk <= 0) hops_per_chunk = 64;
    hops_per_chunk = std::min(hops_per_chunk, n_hops);
    if (!e->s_compute) {
        FE_CUDA(cudaStreamCreateWithFlags(&e->s_copy_in, cudaStreamNonBlocking));
        FE_CUDA(cudaStreamCreateWithFlags(&e->s_compute, cudaStreamNonBlocking));
        FE_CUDA(cudaStreamCreateWithFlags(&e->s_copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            FE_CUDA(cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming));
            FE_CUDA(cudaEventCreateWithFlags(&e->ev_k[i], cudaEventDisableTiming));
            FE_CUDA(cudaEventCreateWithFlags(&e->ev_out[i], cudaEventDisableTiming));
        }
    }
    const size_t need = (size_t)B * hops_per_chunk * H;
    if (e->h_floats < need) {
        for (int i = 0; i < 2; ++i) {
            if (e->h_in[i]) FE_CUDA(cudaFree(e->h_in[i]));
            if (e->h_out[i]) FE_CUDA(cudaFree(e->h_out[i]));
            e->h_in[i] = e->h_out[i] = nullptr;
        }
        e->h_floats = 0;
        for (int i = 0; i < 2; ++i) {
            FE_CUDA(cudaMalloc(&e->h_in[i], need * sizeof(float)));
            FE_CUDA(cudaMalloc(&e->h_out[i], need * sizeof(float)));
        }
        e->h_floats = need;
    }
    // order against work already queued on the default stream (state reset / import)
    FE_CUDA(cudaStreamSynchronize(nullptr));
    int piece = 0;
    for (int h0 = 0; h0 < n_hops; h0 += hops_per_chunk, ++piece) {
        const int nh = std::min(hops_per_chunk, n_hops - h0), b = piece & 1;
        const size_t w = (size_t)nh * H;
        if (piece >= 2) {   // staging buffers b are free once piece-2 finished its kernel (in) / its copy-out (out)
            FE_CUDA(cudaStreamWaitEvent(e->s_copy_in, e->ev_k[b], 0));
            FE_CUDA(cudaStreamWaitEvent(e->s_compute, e->ev_out[b], 0));
        }
        FE_CUDA(cudaMemcpy2DAsync(e->h_in[b], w * sizeof(float), wav_in_host + (size_t)h0 * H, (size_t)ld_in * sizeof(float),
                                  w * sizeof(float), B, cudaMemcpyHostToDevice, e->s_copy_in));
        FE_CUDA(cudaEventRecord(e->ev_in[b], e->s_copy_in));
        FE_CUDA(cudaStreamWaitEvent(e->s_compute, e->ev_in[b], 0));
        int rc = fe_stream(e, s, e->h_in[b], e->h_out[b], nh, (long long)w, (long long)w, e->s_compute);
        if (rc) return rc;
        FE_CUDA(cudaEventRecord(e->ev_k[b], e->s_compute));
        FE_CUDA(cudaStreamWaitEvent(e->s_copy_out, e->ev_k[b], 0));
        FE_CUDA(cudaMemcpy2DAsync(wav_out_host + (size_t)h0 * H, (size_t)ld_out * sizeof(float), e->h_out[b], w * sizeof(float),
                                  w * sizeof(float), B, cudaMemcpyDeviceToHost, e->s_copy_out));
        FE_CUDA(cudaEventRecord(e->ev_out[b], e->s_copy_out));
    }
    FE_CUDA(cudaStreamSynchronize(e->s_copy_out));
    FE_CUDA(cudaStreamSynchronize(e->s_compute));
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
    if (mode < 0 || mode > 2) return fail(FE_ERR_ARG, "fe_set_precision: mode must be 0 (tf32), 1 (fp32) or 2 (f16)");
    const int tc = mode == 1 ? 0 : (mode == 2 ? 2 : 1);
    bool ok = false;
    for (const Variant& v : e->variants) ok = ok || (v.ops.tc == tc && (e->forced_s == 0 || v.ops.S == e->forced_s));
    if (!ok) return fail(FE_ERR_UNSUPPORTED, "fe_set_precision: this model has no kernel variant for the requested mode");
    e->tc = tc;
    return FE_OK;
}
FE_API int fe_get_precision(fe_engine* e) {
    return e ? (e->tc == 0 ? 1 : (e->tc == 2 ? 2 : 0)) : fail(FE_ERR_ARG, "fe_get_precision: null engine");
}
FE_API int fe_profile_slots(void) { return (int)(fe::PH_COUNT + fe::PH_COUNT * fe::PH_NSUB); }
FE_API int fe_set_profile(fe_engine* e, long long* counters_device) {
    if (!e) return fail(FE_ERR_ARG, "fe_set_profile: null engine");
    e->prof = counters_device;
    return FE_OK;
}
FE_API long long fe_kernel_launches(fe_engine* e) { return e ? e->launches : 0; }
FE_API int fe_tap_floats(fe_engine* e) { return e ? e->variants[0].ops.tap_floats : 0; }

}  // extern "C"
