// fe_emu.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build (-DFE_EMU) of the fused kernel body: every barrier-separated phase of
// fastenhancer_b200/csrc/fe_kernel.cuh runs as `for tid in 0..NT`, CTAs run one after another and
// the weight ring is read straight from the packed blob.  It checks the packer, the tile / layout
// index math and the buffer aliasing plan against the oracle on the build container (no GPU).
// It is never linked into the product library.
#define FE_EMU 1
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../fastenhancer_b200/csrc/fe_configs.h"
#include "../../fastenhancer_b200/csrc/fe_kernel.cuh"
#include "../../fastenhancer_b200/csrc/fe_pack.h"

namespace fe {

template <class P> struct EmuCtx {
    float* sm; const float* blob; KParams prm; int s0; float* gs; int cta, ncta = 1;
    long frames_done = 0;
    int h0 = 0, h1 = 0;
    int hbeg() const { return P::SLICED ? h0 : 0; }
    int hend() const { return P::SLICED ? h1 : prm.n_hops; }
    void begin_range(int a, int b) { h0 = a; h1 = b; hops_loaded = a; }
    void wait_item(int) const {}        // hop-sliced launches: the emulation runs the items in order on one CTA
    void signal_item(int) const {}
    const int* table;
    const float* acquire(int ci, int expect_floats) const {
        if (ci < 0 || ci >= P::NCHUNK_FRAME) throw std::runtime_error("emu: chunk index out of range");
        if (table[2 * ci + 1] != expect_floats) throw std::runtime_error("emu: chunk size mismatch at chunk " + std::to_string(ci));
        return blob + table[2 * ci];
    }
    void release(int) const {}
    // ---- hop tiles by TMA (HOP_RING variants): consumer thread 0 issues the tile of hop t right after the window phase of frame t - 1
    //      (when its ring tile has been read): emulated at that point, zero-filled rows for streams past the end like the hardware's
    //      out-of-bounds fill ----
    int hops_loaded = 0;
    void hop_prefetch(int hop) {
        if (hop != hops_loaded) throw std::runtime_error("emu: hop tiles prefetched out of order");
        ++hops_loaded;
        if (hop >= prm.n_hops) return;
        constexpr int H = P::Cf::HOP, HT = P::HT, N = P::Cf::N_FFT;
        for (int k = 0; k < H / HT; ++k) {
            const int pos = (hop * H + k * HT) & (N - 1);
            float* dst = sm + P::SM_TIN + (pos / HT) * P::S * HT;
            for (int s = 0; s < P::S; ++s)
                for (int i = 0; i < HT; ++i)
                    dst[s * HT + i] = (s0 + s < prm.n_streams) ? prm.in[(size_t)(s0 + s) * prm.ld_in + (size_t)hop * H + k * HT + i] : 0.f;
        }
    }
    void hop_wait(int hop) const { if (hop >= hops_loaded) throw std::runtime_error("emu: hop tile awaited before it was prefetched"); }
    void hop_store(int hop) {
        constexpr int H = P::Cf::HOP, HT = P::HT, N = P::Cf::N_FFT;
        for (int k = 0; k < H / HT; ++k) {
            const int pos = (hop * H + k * HT) & (N - 1);
            const float* src = sm + P::SM_OLA + (pos / HT) * P::S * HT;
            for (int s = 0; s < P::S; ++s)
                if (s0 + s < prm.n_streams)
                    for (int i = 0; i < HT; ++i) prm.out[(size_t)(s0 + s) * prm.ld_out + (size_t)hop * H + k * HT + i] = src[s * HT + i];
        }
    }
    void hop_store_wait(bool) const {}
    template <class F> void phase(int, F&& f) { for (int t = 0; t < P::NT; ++t) f(t); }
    template <class A, class F1, class F2> void phase2(int, F1&& f1, F2&& f2) {
        std::vector<A> acc(P::NT);
        for (int t = 0; t < P::NT; ++t) f1(t, acc[t]);
        for (int t = 0; t < P::NT; ++t) f2(t, acc[t]);
    }
    // ---- tensor-core emulation: TMEM as a [128 lanes][512 columns] array; operands are read through the
    //      canonical K-major / no-swizzle formula the descriptors encode (row r at r*16 B, k-chunk at LBO) ----
    std::vector<float> tmem = std::vector<float>(128 * 512, std::nanf(""));
    static float tf32_trunc(float x) { uint32_t u; std::memcpy(&u, &x, 4); u &= 0xffffe000u; std::memcpy(&x, &u, 4); return x; }
    void mma_fence() const {}
    bool elect(int tid) const { return tid == 0; }
    void warp_sync() const {}
    void sub_begin(int) const {}
    void sub_end(int, int) const {}
    struct Desc { const float* p; int lbo; int sbo = 32; };     // float units; K-major operands have SBO = 128 bytes
    Desc make_desc(const float* p, int lbo_floats) const { return Desc{p, lbo_floats}; }
    Desc make_desc_mn(const float* p, int sbo_floats) const { return Desc{p, 32, sbo_floats}; }      // LBO = 128 bytes
    Desc desc_add(Desc d, int floats) const { return Desc{d.p + floats, d.lbo, d.sbo}; }
    Desc desc_set_lbo(Desc d, int lbo_floats) const { return Desc{d.p, lbo_floats, d.sbo}; }
    // element (n, k) of an MN-major / no-swizzle 16-bit operand: 8 n per 16-byte row, 8 k rows per 128-byte core matrix, groups of 8 k
    // LBO apart, groups of 8 n SBO apart (canonical layout; tools/tc_probe_mn.cu checks it on hardware)
    template <bool BF = false> static float h16mn(const Desc& d, int n, int k) {
        const uint16_t* h = reinterpret_cast<const uint16_t*>(d.p + (n / 8) * d.sbo + (k / 8) * d.lbo + (k % 8) * 4);
        return h16_bits_to_f32<BF>(h[n % 8]);
    }
    // element (row r, k) of a K-major / no-swizzle fp16 operand: 8 halves per 16-byte row, k-chunks of 8 LBO apart
    template <bool BF = false> static float h16(const float* p, int lbo, int r, int k) {
        const uint16_t* h = reinterpret_cast<const uint16_t*>(p + (k / 8) * lbo + r * 4);
        return h16_bits_to_f32<BF>(h[k % 8]);
    }
    static constexpr int FMT16 = P::BF16 ? 2 : 1;
    template <bool M64 = false, int FMT = 0, bool BMN = false>
    void mma(int tid, Desc a, Desc b, int NP, int col, bool acc, int rows) {
        constexpr bool F16 = FMT != 0, BF = FMT == 2;
        if (tid != 0) return;                       // one elected lane of warp 0 issues
        if (M64 && rows > 64) throw std::runtime_error("emu: M = 64 MMA with more than 64 rows");
        for (int m = 0; m < rows; ++m) {
            const int lane = M64 ? 32 * (m / 16) + m % 16 : m;      // M = 64: 16 rows per TMEM lane quadrant
            for (int n = 0; n < NP; ++n) {
                float sum = acc ? tmem[lane * 512 + col + n] : 0.f;
                if (F16) {
                    for (int k = 0; k < 16; ++k) sum += h16<BF>(a.p, a.lbo, m, k) * (BMN ? h16mn<BF>(b, n, k) : h16<BF>(b.p, b.lbo, n, k));
                } else {
                    for (int k = 0; k < 8; ++k)
                        sum += tf32_trunc(a.p[(k / 4) * a.lbo + m * 4 + (k % 4)]) * tf32_trunc(b.p[(k / 4) * b.lbo + n * 4 + (k % 4)]);
                }
                tmem[lane * 512 + col + n] = sum;
            }
        }
    }
    template <bool F16 = false>
    void mma_ts(int tid, int a_col, Desc b, int NP, int col, bool acc, int rows) {      // A operand from tensor memory
        if (tid != 0) return;
        for (int m = 0; m < rows; ++m)
            for (int n = 0; n < NP; ++n) {
                float sum = acc ? tmem[m * 512 + col + n] : 0.f;
                if (F16) {          // 16 halves in 8 columns, the even k in the low half
                    for (int k = 0; k < 16; ++k) {
                        uint32_t u; std::memcpy(&u, &tmem[m * 512 + a_col + k / 2], 4);
                        sum += f16_bits_to_f32((uint16_t)((k & 1) ? (u >> 16) : (u & 0xffffu))) * h16(b.p, b.lbo, n, k);
                    }
                } else {
                    for (int k = 0; k < 8; ++k) sum += tf32_trunc(tmem[m * 512 + a_col + k]) * tf32_trunc(b.p[(k / 4) * b.lbo + n * 4 + (k % 4)]);
                }
                tmem[m * 512 + col + n] = sum;
            }
    }
    void tmem_st2(int tid, int col, const float* v) {
        const int row = (((tid >> 5) & 3) << 5) + (tid & 31);
        tmem[row * 512 + col] = v[0]; tmem[row * 512 + col + 1] = v[1];
    }
    void tmem_st2_row(int row, int col, const float* v) { tmem[row * 512 + col] = v[0]; tmem[row * 512 + col + 1] = v[1]; }
    void tmem_st4(int tid, int col, const float* v) {
        const int row = (((tid >> 5) & 3) << 5) + (tid & 31);
        for (int e = 0; e < 4; ++e) tmem[row * 512 + col + e] = v[e];
    }
    void tmem_st4_row(int row, int col, const float* v) { for (int e = 0; e < 4; ++e) tmem[row * 512 + col + e] = v[e]; }
    void async_copy16(float* dst, const float* src) const { std::memcpy(dst, src, 16); }
    void async_commit() const {}
    void async_wait_all() const {}
    void tmem_st_wait() const {}
    void tmem_ld16(int tid, int col, float* v) const {      // tcgen05.ld.16x256b.x1
        const int q = (tid >> 5) & 3, t = tid & 31, l0 = 32 * q + t / 4, c = col + 2 * (t % 4);
        v[0] = tmem[l0 * 512 + c]; v[1] = tmem[l0 * 512 + c + 1];
        v[2] = tmem[(l0 + 8) * 512 + c]; v[3] = tmem[(l0 + 8) * 512 + c + 1];
    }
    void release_mma(int) const {}
    void acc_commit_wait() const {}
    void tmem_ld4(int tid, int col, float* v) const {
        const int row = (((tid >> 5) & 3) << 5) + (tid & 31);
        for (int e = 0; e < 4; ++e) v[e] = tmem[row * 512 + col + e];
    }
    void tmem_ld_wait() const {}
    void next_frame() { ++frames_done; }
    void check_frame(int ci) const { if (ci != P::NCHUNK_FRAME) throw std::runtime_error("emu: frame consumed wrong number of chunks"); }
};

template <class P> int run_variant(const float* canonical, KParams prm) {
    std::vector<float> blob = pack_blob<P>(canonical);
    constexpr auto A = P::make_aux();
    int grid = (prm.n_streams + P::S - 1) / P::S;
    std::vector<float> sm(P::SM_TOTAL), gs((size_t)P::GS_TOTAL * (P::SLICED && prm.slice_hops > 0 ? grid : 1));
    if (P::SLICED && prm.slice_hops > 0) { prm.scratch = gs.data(); grid = 1; }        // one emulated CTA takes every item, in order
    for (int cta = 0; cta < grid; ++cta) {
        // poison shared memory so that reads of never-written locations show up
        for (auto& v : sm) v = std::nanf("");
        EmuCtx<P> x;
        x.sm = sm.data(); x.blob = blob.data(); x.prm = prm; x.prm.blob = blob.data();
        x.prm.hop_tma = (P::HOP_RING && prm.mode == MODE_STREAM && !std::getenv("FE_EMU_NO_HOP_TMA")) ? 1 : 0;
        x.s0 = cta * P::S; x.gs = gs.data(); x.cta = cta; x.ncta = grid;
        x.table = reinterpret_cast<const int*>(blob.data() + A.table);
        Frame<P>::run(x);
    }
    return 0;
}

// Frame-parallel offline schedule (fe_api.cu::offline_tp) on the host: the staged launches of the fused kernel body with a plain C
// restatement of the two small kernels that run between them (fe_gru_scan_kernel, fe_overlap_add_kernel) -- checks the stage
// boundaries, the group scratch and the chunk ranges of the weight stream.
template <class P> int run_offline_tp(const float* canonical, KParams prm, int grid) {
    if constexpr (P::TC) { (void)canonical; (void)prm; (void)grid; return -3; }
    else {
        using C = typename P::Cf;
        constexpr int C2 = C::C2, F2 = C::F2, N = C::N_FFT, H = C::HOP;
        std::vector<float> blob = pack_blob<P>(canonical);
        constexpr auto A = P::make_aux();
        const int B = prm.n_streams, T = prm.n_hops;
        const long nf = (long)B * T;
        const int ngroups = (int)((nf + P::S - 1) / P::S);
        std::vector<float> scr((size_t)ngroups * P::TP_GROUP, std::nanf("")), gx((size_t)nf * F2 * 3 * C2, std::nanf("")), hs((size_t)nf * F2 * C2, std::nanf("")),
            frames((size_t)nf * N, std::nanf("")), sm(P::SM_TOTAL);
        auto stage = [&](int st, int blk) {
            for (int cta = 0; cta < grid; ++cta) {
                for (auto& v : sm) v = std::nanf("");
                EmuCtx<P> x;
                x.sm = sm.data(); x.blob = blob.data(); x.prm = prm; x.prm.blob = blob.data(); x.prm.hop_tma = 0;
                x.prm.tp_stage = st; x.prm.tp_blk = blk; x.prm.tp_scr = scr.data(); x.prm.tp_gx = gx.data(); x.prm.tp_h = hs.data(); x.prm.tp_frames = frames.data();
                x.s0 = cta * P::S; x.gs = nullptr; x.cta = cta; x.ncta = grid;
                x.table = reinterpret_cast<const int*>(blob.data() + A.table);
                Frame<P>::run(x);
            }
        };
        Canon<C> cw(canonical);
        auto sig = [](float v) { return 1.0f / (1.0f + std::exp(-v)); };
        stage(1, 0);
        for (int k = 0; k < C::K; ++k) {
            const auto& b = cw.blk[k];
            for (int u = 0; u < B; ++u)
                for (int f = 0; f < F2; ++f) {
                    std::vector<float> h(C2, 0.f), hn(C2);
                    for (int t = 0; t < T; ++t) {
                        const long q = (long)u * T + t;
                        const float* g = gx.data() + ((q * F2 + f) * 3) * C2;
                        for (int j = 0; j < C2; ++j) {
                            float ar = 0.f, az = 0.f, an = 0.f;
                            for (int c = 0; c < C2; ++c) {
                                ar += b.w_hh[(0 * C2 + j) * C2 + c] * h[c]; az += b.w_hh[(1 * C2 + j) * C2 + c] * h[c]; an += b.w_hh[(2 * C2 + j) * C2 + c] * h[c];
                            }
                            const float r = sig(g[j] + b.b_ih[j] + b.b_hh[j] + ar), z = sig(g[C2 + j] + b.b_ih[C2 + j] + b.b_hh[C2 + j] + az);
                            const float n = std::tanh(g[2 * C2 + j] + b.b_ih[2 * C2 + j] + r * (an + b.b_hh[2 * C2 + j]));
                            hn[j] = (1.0f - z) * n + z * h[j];
                        }
                        h = hn;
                        for (int j = 0; j < C2; ++j) hs[(q * F2 + f) * C2 + j] = h[j];
                    }
                }
            stage(2, k);
        }
        const float* wsq = blob.data() + A.window_sq;
        for (int u = 0; u < B; ++u)
            for (long n = 0; n < (long)H * (T - 1); ++n) {
                const long npad = n + N / 2;
                long t0 = npad < N ? 0 : (npad - N + H) / H, t1 = npad / H;
                if (t1 > T - 1) t1 = T - 1;
                float v = 0.f, env = 0.f;
                for (long t = t0; t <= t1; ++t) { v += frames[((size_t)u * T + t) * N + (npad - t * H)]; env += wsq[npad - t * H]; }
                prm.out[(size_t)u * H * (T - 1) + n] = v / env;
            }
        return 0;
    }
}

}  // namespace fe

// The variants are instantiated in ten groups (one per model configuration) so that emu.py can compile them in parallel:
//   -DFE_EMU_GROUP=g   : fee_run_g<g> and fee_tap_g<g> for the variants of configuration g
//   -DFE_EMU_MAIN      : fee_run / fee_tap_total dispatching over the groups
#define FE_CAT2(a, b) a##b
#define FE_CAT(a, b) FE_CAT2(a, b)
#ifdef FE_EMU_GROUP
#if FE_EMU_GROUP == 0
#define FE_GROUP_VARIANTS FE_VARIANTS_16T
#elif FE_EMU_GROUP == 1
#define FE_GROUP_VARIANTS FE_VARIANTS_16B
#elif FE_EMU_GROUP == 2
#define FE_GROUP_VARIANTS FE_VARIANTS_16S
#elif FE_EMU_GROUP == 3
#define FE_GROUP_VARIANTS FE_VARIANTS_16M
#elif FE_EMU_GROUP == 4
#define FE_GROUP_VARIANTS FE_VARIANTS_16L
#elif FE_EMU_GROUP == 5
#define FE_GROUP_VARIANTS FE_VARIANTS_48T
#elif FE_EMU_GROUP == 6
#define FE_GROUP_VARIANTS FE_VARIANTS_48B
#elif FE_EMU_GROUP == 7
#define FE_GROUP_VARIANTS FE_VARIANTS_48S
#elif FE_EMU_GROUP == 8
#define FE_GROUP_VARIANTS FE_VARIANTS_48M
#else
#define FE_GROUP_VARIANTS FE_VARIANTS_48L
#endif
extern "C" int FE_CAT(fee_run_g, FE_EMU_GROUP)(const fe::ShapeKey* key, int S, int tc, const float* canonical, const fe::KParams* prm)
{
#define X(id, CFG, SV, TCV) if (fe::shape_matches<fe::CFG>(*key) && S == SV && tc == (int)(TCV)) return fe::run_variant<fe::Plan<fe::CFG, SV, TCV>>(canonical, *prm);
    FE_GROUP_VARIANTS(X)
#undef X
    return -1;
}
extern "C" int FE_CAT(fee_tp_g, FE_EMU_GROUP)(const fe::ShapeKey* key, int S, const float* canonical, const fe::KParams* prm, int grid)
{
#define X(id, CFG, SV, TCV) if (fe::shape_matches<fe::CFG>(*key) && S == SV && (int)(TCV) == 0) return fe::run_offline_tp<fe::Plan<fe::CFG, SV, 0>>(canonical, *prm, grid);
    FE_GROUP_VARIANTS(X)
#undef X
    return -1;
}
extern "C" int FE_CAT(fee_tap_g, FE_EMU_GROUP)(const fe::ShapeKey* key)
{
#define X(id, CFG, SV, TCV) if (fe::shape_matches<fe::CFG>(*key)) return fe::Frame<fe::Plan<fe::CFG, SV, TCV>>::TAP_TOTAL;
    FE_GROUP_VARIANTS(X)
#undef X
    return -1;
}
#endif

#ifdef FE_EMU_MAIN
#define FE_GROUPS(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9)
#define X(g) extern "C" int fee_run_g##g(const fe::ShapeKey*, int, int, const float*, const fe::KParams*); extern "C" int fee_tap_g##g(const fe::ShapeKey*); \
    extern "C" int fee_tp_g##g(const fe::ShapeKey*, int, const float*, const fe::KParams*, int);
FE_GROUPS(X)
#undef X
extern "C" int fee_run(int n_fft, int hop, int c1, int n_enc, int c2, int f2, int n_blocks, int n_heads, int S, int tc,
                       const float* canonical, int mode, float* state, const float* in, float* out, float* spec_out,
                       int n_streams, int n_hops, int L, long long ld_in, long long ld_out, float* dbg, int dbg_hop,
                       float compression, int slice_hops)
{
    fe::ShapeKey key{n_fft, hop, c1, n_enc, c2, f2, n_blocks, n_heads};
    fe::KParams prm{};
    prm.state = state; prm.in = in; prm.out = out; prm.spec_out = spec_out; prm.dbg = dbg;
    prm.ld_in = ld_in; prm.ld_out = ld_out; prm.n_streams = n_streams; prm.n_hops = n_hops; prm.mode = mode; prm.L = L;
    prm.dbg_hop = dbg_hop; prm.compression = compression; prm.slice_hops = slice_hops;
    try {
#define X(g) { const int rc = fee_run_g##g(&key, S, tc, canonical, &prm); if (rc != -1) return rc; }
        FE_GROUPS(X)
#undef X
    } catch (const std::exception& e) {
        std::fprintf(stderr, "fee_run: %s\n", e.what());
        return -2;
    }
    return -1;
}

// Model.forward on [B, L] through the frame-parallel schedule (fp32 family), `grid` emulated CTAs
extern "C" int fee_offline_tp(int n_fft, int hop, int c1, int n_enc, int c2, int f2, int n_blocks, int n_heads, int S, const float* canonical,
                              const float* in, float* out, float* spec_out, int B, int L, int grid, float compression)
{
    fe::ShapeKey key{n_fft, hop, c1, n_enc, c2, f2, n_blocks, n_heads};
    fe::KParams prm{};
    prm.in = in; prm.out = out; prm.spec_out = spec_out; prm.n_streams = B; prm.n_hops = 1 + L / hop; prm.mode = fe::MODE_OFFLINE; prm.L = L;
    prm.dbg_hop = -1; prm.compression = compression;
    try {
#define X(g) { const int rc = fee_tp_g##g(&key, S, canonical, &prm, grid); if (rc != -1) return rc; }
        FE_GROUPS(X)
#undef X
    } catch (const std::exception& e) {
        std::fprintf(stderr, "fee_offline_tp: %s\n", e.what());
        return -2;
    }
    return -1;
}

extern "C" int fee_tap_total(int n_fft, int hop, int c1, int n_enc, int c2, int f2, int n_blocks, int n_heads)
{
    fe::ShapeKey key{n_fft, hop, c1, n_enc, c2, f2, n_blocks, n_heads};
#define X(g) { const int rc = fee_tap_g##g(&key); if (rc != -1) return rc; }
    FE_GROUPS(X)
#undef X
    return -1;
}
#endif
