// fe_pack.h -- host-side packer: canonical folded weights -> the blob the fused kernel consumes.
//
// The canonical array (include/fastenhancer_b200.h, fastenhancer_b200/schema.py::canonical_schema)
// is what the reference's remove_weight_reparameterizations leaves behind
// (models/fastenhancer/default/model.py:532-608).  The blob is
//     [ aux: windows, FFT twiddles, biases, positional embedding | chunk table | ring section ]
// where the ring section is every layer's weights in execution order, laid out as the rows the
// PosGemm / RowGemm tiles of fe_plan.h read (one contiguous slice per ring chunk).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "fe_half.h"
#include "fe_plan.h"

namespace fe {

template <class C> struct Canon {
    struct Blk { const float *w_ih, *w_hh, *b_ih, *b_hh, *fc_w, *fc_b, *pe, *qkv_w, *qkv_b, *afc_w, *afc_b; };
    const float *enc_pre_w, *enc_pre_b, *enc_w[8], *enc_b[8], *rf_pre_lin, *rf_pre_w, *rf_pre_b;
    Blk blk[16];
    const float *rf_post_lin, *rf_post_w, *rf_post_b, *dec_w1[8], *dec_b1[8], *dec_w2[8], *dec_b2[8];
    const float *dp_w, *dp_b, *dp_wt, *dp_bt;
    explicit Canon(const float* p) {
        constexpr int C1 = C::C1, C2 = C::C2, F1 = C::F1, F2 = C::F2;
        auto take = [&p](long n) { const float* r = p; p += n; return r; };
        enc_pre_w = take(C1 * 16); enc_pre_b = take(C1);
        for (int i = 0; i < C::E; ++i) { enc_w[i] = take(C1 * C1 * 3); enc_b[i] = take(C1); }
        rf_pre_lin = take(F2 * F1); rf_pre_w = take(C2 * C1); rf_pre_b = take(C2);
        for (int k = 0; k < C::K; ++k) {
            Blk& b = blk[k];
            b.w_ih = take(3 * C2 * C2); b.w_hh = take(3 * C2 * C2); b.b_ih = take(3 * C2); b.b_hh = take(3 * C2);
            b.fc_w = take(C2 * C2); b.fc_b = take(C2);
            b.pe = (k == 0) ? take(F2 * C2) : nullptr;
            b.qkv_w = take(3 * C2 * C2); b.qkv_b = take(3 * C2);
            b.afc_w = take(C2 * C2); b.afc_b = take(C2);
        }
        rf_post_lin = take(F1 * F2); rf_post_w = take(C1 * C2); rf_post_b = take(C1);
        for (int i = 0; i < C::E; ++i) {
            dec_w1[i] = take(C1 * 2 * C1); dec_b1[i] = take(C1); dec_w2[i] = take(C1 * C1 * 3); dec_b2[i] = take(C1);
        }
        dp_w = take(C1 * 2 * C1); dp_b = take(C1); dp_wt = take(C1 * 16); dp_bt = take(2);
    }
};

template <class P> class Packer {
    using C = typename P::Cf;
    std::vector<float>& blob_;
    std::vector<int> table_;
    long off_;

    // w(set, co, k, tap)
    template <class L, class W> void pos(W w) {
        for (int pass = 0; pass < L::NPASS; ++pass) {
            for (int c = 0; c < L::NCHUNK_PASS; ++c) {
                int rows = cmin(L::KC, L::K - c * L::KC);
                table_.push_back((int)(off_ + ((long)pass * L::K + (long)c * L::KC) * L::ROW));
                table_.push_back(rows * L::ROW);
            }
            for (int k = 0; k < L::K; ++k)
                for (int cgp = 0; cgp < L::NCGP; ++cgp)
                    for (int cl = 0; cl < L::CL; ++cl) {
                        float* d = &blob_[off_ + (((long)pass * L::K + k) * L::NCGP + cgp) * L::CL * L::RW + (long)cl * L::RW];
                        for (int set = 0; set < L::SETS; ++set)
                            for (int t = 0; t < L::TAPS; ++t)
                                for (int i = 0; i < L::CT; ++i) {
                                    int co = ((pass * L::NCGP + cgp) * L::CL + cl) * L::CT + i;
                                    d[(set * L::TAPS + t) * L::CT + i] = co < L::COUT ? w(set, co, k, t) : 0.f;
                                }
                    }
        }
        off_ += L::FLOATS;
    }
    // w(o, k)
    template <class L, class W> void row(W w) {
        for (int c = 0; c < L::NCHUNK; ++c) {
            int rows = cmin(L::KC, L::K4 - c * L::KC);
            table_.push_back((int)(off_ + (long)c * L::KC * L::ROW));
            table_.push_back(rows * L::ROW);
        }
        for (int k4 = 0; k4 < L::K4; ++k4)
            for (int og = 0; og < L::NOG; ++og)
                for (int j = 0; j < L::NO; ++j)
                    for (int t = 0; t < 4; ++t) {
                        int o = og * L::NO + j;
                        blob_[off_ + ((long)k4 * L::NOG + og) * L::NO * 4 + j * 4 + t] = o < L::NOUT ? w(o, 4 * k4 + t) : 0.f;
                    }
        off_ += L::FLOATS;
    }

    // round-to-nearest (ties away) to TF32: what cvt.rna.tf32.f32 does; the tensor core then reads the value exactly
    static float tf32_rna(float x) {
        uint32_t u;
        std::memcpy(&u, &x, 4);
        u = (u + 0x1000u) & 0xffffe000u;
        std::memcpy(&x, &u, 4);
        return x;
    }
    // tensor-core tiles, w(n, k, tap): tile (tap, k-step j) = [2][NP][4] holding W[n][8j + 4*kc2 + e][tap] (TF32), or
    // [2][NP][8 halves] holding W[n][16j + 8*kc2 + e][tap] (fp16 layers: KE = 16)
    template <class L, class W> void tc(W w, float scale = 1.f) {
        for (int c = 0; c < L::NCHUNK; ++c) {
            int tiles = cmin(L::TPC, L::NTILE - c * L::TPC);
            table_.push_back((int)(off_ + (long)c * L::TPC * L::TILE));
            table_.push_back(tiles * L::TILE);
        }
        for (int tile = 0; tile < L::NTILE; ++tile) {
            const int t = tile / L::NKS, j = tile % L::NKS;
            if constexpr (L::KE == 16) {
                // hi parts, then (split layers) the tile of the remainders w - fp16(w)
                uint16_t* h = reinterpret_cast<uint16_t*>(&blob_[off_ + (long)tile * L::TILE]);
                uint16_t* lo = reinterpret_cast<uint16_t*>(&blob_[off_ + (long)tile * L::TILE + L::TILE1]);
                for (int kc2 = 0; kc2 < 2; ++kc2)
                    for (int n = 0; n < L::NP; ++n)
                        for (int e = 0; e < 8; ++e) {
                            const int k = 16 * j + 8 * kc2 + e;
                            const float v = (n < L::N && k < L::K) ? scale * w(n, k, t) : 0.f;
                            const uint16_t hb = f32_to_h16_bits<P::BF16>(v);
                            h[(kc2 * L::NP + n) * 8 + e] = hb;
                            if constexpr (L::PARTS == 2) lo[(kc2 * L::NP + n) * 8 + e] = f32_to_f16_bits(v - f16_bits_to_f32(hb));
                        }
            } else {
                static_assert(L::PARTS == 1, "split layers have 16-bit operands");
                for (int kc2 = 0; kc2 < 2; ++kc2)
                    for (int n = 0; n < L::NP; ++n)
                        for (int e = 0; e < 4; ++e) {
                            const int k = 8 * j + 4 * kc2 + e;
                            blob_[off_ + (long)tile * L::TILE + (kc2 * L::NP + n) * 4 + e] = (n < L::N && k < L::K) ? tf32_rna(scale * w(n, k, t)) : 0.f;
                        }
            }
        }
        off_ += L::FLOATS;
    }
    // frequency-axis linear on the tensor cores (TcLin), w(f_out, f_in): tile (M tile mt, k-step j) = [2][128][8 halves] holding
    // W[row 128 mt + r][slot 16 j + 8 kc2 + e] with row = f_out * S + s, slot = f_in * S + s' and W = (s == s') ? w(f_out, f_in) : 0
    template <class L, class W> void lin(W w) {
        for (int c = 0; c < L::NCHUNK; ++c) {
            int tiles = cmin(L::TPC, L::NTILE - c * L::TPC);
            table_.push_back((int)(off_ + (long)c * L::TPC * L::TILE));
            table_.push_back(tiles * L::TILE);
        }
        for (int tile = 0; tile < L::NTILE; ++tile) {
            const int mt = tile / L::NKS, j = tile % L::NKS;
            uint16_t* h = reinterpret_cast<uint16_t*>(&blob_[off_ + (long)tile * L::TILE]);
            uint16_t* lo = reinterpret_cast<uint16_t*>(&blob_[off_ + (long)tile * L::TILE + L::TILE1]);
            for (int kc2 = 0; kc2 < 2; ++kc2)
                for (int r = 0; r < 128; ++r)
                    for (int e = 0; e < 8; ++e) {
                        const int row = 128 * mt + r, k = 16 * j + 8 * kc2 + e;
                        const float v = (row < L::NPOS && row % P::S == k % P::S) ? w(row / P::S, k / P::S) : 0.f;
                        const uint16_t hb = f32_to_h16_bits<P::BF16>(v);
                        h[(kc2 * 128 + r) * 8 + e] = hb;
                        if constexpr (L::PARTS == 2) lo[(kc2 * 128 + r) * 8 + e] = f32_to_f16_bits(v - f16_bits_to_f32(hb));
                    }
        }
        off_ += L::FLOATS;
    }
    // fused GRU tiles, w(set, c, k) with sets W_ir W_iz W_in W_hr W_hz W_hn: tile (inp, j) = [R|Z: [2][2*NPG][4] | N: [2][NPG][4]]
    template <class L, class W> void gru(W w) {
        for (int c = 0; c < L::NCHUNK; ++c) {
            int tiles = cmin(L::TPC, L::NTILE - c * L::TPC);
            table_.push_back((int)(off_ + (long)c * L::TPC * L::TILE));
            table_.push_back(tiles * L::TILE);
        }
        constexpr int KH = L::KE / 2;           // k's per 16-byte row chunk: 4 floats (TF32) or 8 halves
        for (int tile = 0; tile < L::NTILE; ++tile) {
            float* t = &blob_[off_ + (long)tile * L::TILE];
            // element (k-chunk kc2, row n of a part with `rows` rows starting `base` floats into the tile, e) <- value
            auto put = [&](int base, int rows, int kc2, int n, int e, float v) {
                if constexpr (L::KE == 16) {
                    const uint16_t hb = f32_to_f16_bits(v);
                    reinterpret_cast<uint16_t*>(t + base)[(kc2 * rows + n) * 8 + e] = hb;
                    if constexpr (L::PARTS == 2)
                        reinterpret_cast<uint16_t*>(t + L::TILE1 + base)[(kc2 * rows + n) * 8 + e] = f32_to_f16_bits(v - f16_bits_to_f32(hb));
                } else t[base + (kc2 * rows + n) * 4 + e] = tf32_rna(v);
            };
            if constexpr (L::MERGED) {
                // h tiles first; rows [ W_in | W_r | W_z | W_hn ], the set that does not belong to this input is zero
                const int inp = tile < L::NKS ? 1 : 0, j = tile % L::NKS;
                for (int kc2 = 0; kc2 < 2; ++kc2)
                    for (int e = 0; e < KH; ++e) {
                        const int k = L::KE * j + KH * kc2 + e;
                        for (int n = 0; n < 4 * L::NPG; ++n) {
                            const int blk = n / L::NPG, c = n % L::NPG;             // 0: in, 1: r, 2: z, 3: hn
                            const int set = blk == 0 ? 2 : (blk == 1 ? inp * 3 + 0 : (blk == 2 ? inp * 3 + 1 : 5));
                            const bool mine = (blk == 0) ? inp == 0 : (blk == 3 ? inp == 1 : true);
                            put(0, 4 * L::NPG, kc2, n, e, (mine && c < L::N && k < L::K) ? w(set, c, k) : 0.f);
                        }
                    }
            } else {
                const int inp = tile / L::NKS, j = tile % L::NKS;
                for (int kc2 = 0; kc2 < 2; ++kc2)
                    for (int e = 0; e < KH; ++e) {
                        const int k = L::KE * j + KH * kc2 + e;
                        for (int n = 0; n < 2 * L::NPG; ++n) {
                            const int set = inp * 3 + n / L::NPG, c = n % L::NPG;
                            put(0, 2 * L::NPG, kc2, n, e, (c < L::N && k < L::K) ? w(set, c, k) : 0.f);
                        }
                        for (int n = 0; n < L::NPG; ++n)
                            put(2 * L::NPG * 8, L::NPG, kc2, n, e, (n < L::N && k < L::K) ? w(inp * 3 + 2, n, k) : 0.f);
                    }
            }
        }
        off_ += L::FLOATS;
    }
    // w(o, k), one ring row per k: [og][NO]
    template <class L, class W> void rowk1(W w) {
        for (int c = 0; c < L::NCHUNK; ++c) {
            int rows = cmin(L::KC, L::K - c * L::KC);
            table_.push_back((int)(off_ + (long)c * L::KC * L::ROW));
            table_.push_back(rows * L::ROW);
        }
        for (int k = 0; k < L::K; ++k)
            for (int o = 0; o < L::NOG * L::NO; ++o) blob_[off_ + (long)k * L::ROW + o] = o < L::NOUT ? w(o, k) : 0.f;
        off_ += L::FLOATS;
    }

public:
    explicit Packer(std::vector<float>& blob) : blob_(blob), off_(0) {}

    void pack(const float* canonical) {
        constexpr int C1 = C::C1, C2 = C::C2, F1 = C::F1, F2 = C::F2, N = C::N_FFT, H = C::HOP, M = C::M;
        constexpr auto A = P::make_aux();
        Canon<C> cw(canonical);
        blob_.assign((size_t)A.total, 0.f);
        // ---- tables (periodic Hann: functional/audio_modules.py:213-214; window_istft :221-234) ----
        float* win = &blob_[A.window];
        for (int i = 0; i < N; ++i) win[i] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * i / N));
        for (int i = 0; i < N; ++i) {
            float s = 0.f;
            for (int j = i % H; j < N; j += H) s += win[j] * win[j];
            blob_[A.window_istft + i] = win[i] / s;
            blob_[A.window_sq + i] = win[i] * win[i];
        }
        for (int t = 0; t < M / 2; ++t) {
            blob_[A.tw + 2 * t] = (float)std::cos(-2.0 * M_PI * t / M);
            blob_[A.tw + 2 * t + 1] = (float)std::sin(-2.0 * M_PI * t / M);
        }
        for (int k = 0; k < M; ++k) {
            blob_[A.twn + 2 * k] = (float)std::cos(-2.0 * M_PI * k / N);
            blob_[A.twn + 2 * k + 1] = (float)std::sin(-2.0 * M_PI * k / N);
        }
        // ---- biases ----
        auto cp = [&](int dst, const float* src, int n) { std::memcpy(&blob_[dst], src, n * sizeof(float)); };
        cp(A.enc_pre_b, cw.enc_pre_b, C1);
        for (int i = 0; i < C::E; ++i) cp(A.enc_b(i), cw.enc_b[i], C1);
        cp(A.rf_pre_b, cw.rf_pre_b, C2);
        for (int k = 0; k < C::K; ++k) {
            const auto& b = cw.blk[k];
            const auto q = A.blk(k);
            for (int c = 0; c < C2; ++c) {
                blob_[q.b_r + c] = b.b_ih[c] + b.b_hh[c];
                blob_[q.b_z + c] = b.b_ih[C2 + c] + b.b_hh[C2 + c];
                blob_[q.b_in + c] = b.b_ih[2 * C2 + c];
                blob_[q.b_hn + c] = b.b_hh[2 * C2 + c];
            }
            cp(q.fc_b, b.fc_b, C2);
            if (b.pe)        // fp32 variants read [C2][F2] (4 frequencies per thread), tensor-core variants [F2][C2] (4 channels)
                for (int f = 0; f < F2; ++f)
                    for (int c = 0; c < C2; ++c) blob_[q.pe + (P::TC ? f * C2 + c : c * F2 + f)] = b.pe[f * C2 + c];
            if constexpr (P::TC) {       // padded per-head layout [head][q|k|v][HDP]
                for (int h = 0; h < C::NH; ++h)
                    for (int w3 = 0; w3 < 3; ++w3)
                        for (int d = 0; d < C::HD; ++d)
                            blob_[q.qkv_b + (h * 3 + w3) * P::HDP + d] = b.qkv_b[(h * 3 + w3) * C::HD + d];
            } else {
                cp(q.qkv_b, b.qkv_b, 3 * C2);
            }
            cp(q.afc_b, b.afc_b, C2);
        }
        {
            bool any = false;
            for (int k = 0; k < C::K; ++k)
                for (int i = 0; i < 3 * C2; ++i) any = any || cw.blk[k].qkv_b[i] != 0.f;
            blob_[A.flags] = any ? 1.f : 0.f;
        }
        cp(A.rf_post_b, cw.rf_post_b, C1);
        for (int i = 0; i < C::E; ++i) { cp(A.dec1_b(i), cw.dec_b1[i], C1); cp(A.dec2_b(i), cw.dec_b2[i], C1); }
        cp(A.dp_b, cw.dp_b, C1);
        // Tensor-core variants: layers followed by SiLU carry weights and bias pre-multiplied by 1/2 (exact), so the accumulator is
        // h = x / 2 and the epilogue computes silu(x) = h + h tanh(h) without the extra multiply.
        constexpr float HS = P::FAST_ACT ? 0.5f : 1.f;
        if constexpr (P::FAST_ACT) {
            auto half = [&](int dst, int n) { for (int i = 0; i < n; ++i) blob_[dst + i] *= 0.5f; };
            half(A.enc_pre_b, C1);
            for (int i = 0; i < C::E; ++i) { half(A.enc_b(i), C1); half(A.dec1_b(i), C1); half(A.dec2_b(i), C1); }
            half(A.dp_b, C1);
        }
        for (int vo = 0; vo < 8; ++vo) blob_[A.convt_b + vo] = cw.dp_bt[vo / 4];   // entries 8..15 stay zero (tensor-core N padding)

        // ---- ring section, execution order (must match fe_kernel.cuh::frame) ----
        off_ = A.ring;
        table_.clear();
        auto w_enc_pre = [&](int co, int v, int t) {
            // enc_pre as a 3-tap conv over 8 virtual channels v = c*4 + q holding x[c][4m + q]:
            // original tap k = 4*dj + q + 2 (model.py:15-59, weight index [co][(k%4)*2 + c][k/4])
            int c = v / 4, q = v % 4, k = 4 * (t - 1) + q + 2;
            return (k >= 0 && k < 8) ? cw.enc_pre_w[(co * 8 + (k % 4) * 2 + c) * 2 + k / 4] : 0.f;
        };
        auto w_convt = [&](int vo, int ci, int t) {
            // transposed conv as a 3-tap conv to 8 virtual output channels vo = o*4 + q (bin 4m + q):
            // original tap k = -4*dj + q + 2 (model.py:62-95)
            int o = vo / 4, q = vo % 4, k = -4 * (t - 1) + q + 2;
            return (k >= 0 && k < 8) ? cw.dp_wt[(ci * 2 + o) * 8 + k] : 0.f;
        };
        auto blocks = [&]() {
            for (int k = 0; k < C::K; ++k) {
                const auto& b = cw.blk[k];
                pos<typename P::Gru>([&](int set, int c, int ci, int) {
                    return set < 3 ? b.w_ih[(set * C2 + c) * C2 + ci] : b.w_hh[((set - 3) * C2 + c) * C2 + ci];
                });
                pos<typename P::Fc>([&](int, int co, int ci, int) { return b.fc_w[co * C2 + ci]; });
                for (int g = 0; g < P::NQG; ++g)
                    pos<typename P::Qkv>([&](int, int co, int ci, int) { return b.qkv_w[(g * 3 * C::HD * P::HG + co) * C2 + ci]; });
                pos<typename P::Fc>([&](int, int co, int ci, int) { return b.afc_w[co * C2 + ci]; });
            }
        };
        if constexpr (P::TC) {
            tc<typename P::TEncPre>([&](int co, int v, int t) { return v < 8 ? w_enc_pre(co, v, t) : 0.f; }, HS);
            for (int i = 0; i < C::E; ++i)
                tc<typename P::TConv3>([&](int co, int ci, int t) { return ci < C1 ? cw.enc_w[i][(co * C1 + ci) * 3 + t] : 0.f; }, HS);
            if constexpr (P::LIN_TC) lin<typename P::TLinPre>([&](int o, int k) { return cw.rf_pre_lin[o * F1 + k]; });
            else rowk1<typename P::LinPreT>([&](int o, int k) { return cw.rf_pre_lin[o * F1 + k]; });
            tc<typename P::TRfPre>([&](int co, int ci, int) { return ci < C1 ? cw.rf_pre_w[co * C1 + ci] : 0.f; });
            for (int k = 0; k < C::K; ++k) {
                const auto& b = cw.blk[k];
                gru<typename P::TGru>([&](int set, int c, int ci) {
                    return set < 3 ? b.w_ih[(set * C2 + c) * C2 + ci] : b.w_hh[((set - 3) * C2 + c) * C2 + ci];
                });
                // (ci >= C2: channels that pad the contraction to a whole fp16 k-step)
                tc<typename P::TFc>([&](int co, int ci, int) { return ci < C2 ? b.fc_w[co * C2 + ci] : 0.f; });
                for (int g = 0; g < P::NQG; ++g)
                    tc<typename P::TQkv>([&](int co, int ci, int) {      // co = (head-in-round * 3 + q|k|v) * HDP + d, zero for d >= HD
                        const int d = co % P::HDP, hw = co / P::HDP;
                        return (d < C::HD && ci < C2) ? b.qkv_w[((g * P::HG * 3 + hw) * C::HD + d) * C2 + ci] : 0.f;
                    });
                tc<typename P::TFc>([&](int co, int ci, int) { return ci < C2 ? b.afc_w[co * C2 + ci] : 0.f; });
            }
            if constexpr (P::LIN_TC) lin<typename P::TLinPost>([&](int o, int k) { return cw.rf_post_lin[o * F2 + k]; });
            else rowk1<typename P::LinPostT>([&](int o, int k) { return cw.rf_post_lin[o * F2 + k]; });
            tc<typename P::TRfPost>([&](int co, int ci, int) { return ci < C2 ? cw.rf_post_w[co * C2 + ci] : 0.f; });
            for (int i = 0; i < C::E; ++i) {
                // cat([x, skip]) with each half padded to C1P channels: k < C1P is x channel k, k >= C1P is skip channel k - C1P
                auto cat = [&](const float* w1, int co, int k) {
                    const int half = k / P::C1P, c = k % P::C1P;
                    return c < C1 ? w1[co * 2 * C1 + half * C1 + c] : 0.f;
                };
                tc<typename P::TPwCat>([&](int co, int k, int) { return cat(cw.dec_w1[i], co, k); }, HS);
                tc<typename P::TConv3>([&](int co, int ci, int t) { return ci < C1 ? cw.dec_w2[i][(co * C1 + ci) * 3 + t] : 0.f; }, HS);
            }
            tc<typename P::TPwCat>([&](int co, int k, int) {
                const int half = k / P::C1P, c = k % P::C1P;
                return c < C1 ? cw.dp_w[co * 2 * C1 + half * C1 + c] : 0.f; }, HS);
            tc<typename P::TConvT>([&](int vo, int ci, int t) { return ci < C1 ? w_convt(vo, ci, t) : 0.f; });
        } else {
            pos<typename P::EncPre>([&](int, int co, int v, int t) { return w_enc_pre(co, v, t); });
            for (int i = 0; i < C::E; ++i)
                pos<typename P::Conv3>([&](int, int co, int ci, int t) { return cw.enc_w[i][(co * C1 + ci) * 3 + t]; });
            row<typename P::LinPre>([&](int o, int k) { return cw.rf_pre_lin[o * F1 + k]; });
            pos<typename P::RfPre>([&](int, int co, int ci, int) { return cw.rf_pre_w[co * C1 + ci]; });
            blocks();
            row<typename P::LinPost>([&](int o, int k) { return cw.rf_post_lin[o * F2 + k]; });
            pos<typename P::RfPost>([&](int, int co, int ci, int) { return cw.rf_post_w[co * C2 + ci]; });
            for (int i = 0; i < C::E; ++i) {
                pos<typename P::PwCat>([&](int, int co, int ci, int) { return cw.dec_w1[i][co * 2 * C1 + ci]; });
                pos<typename P::Conv3>([&](int, int co, int ci, int t) { return cw.dec_w2[i][(co * C1 + ci) * 3 + t]; });
            }
            pos<typename P::PwCat>([&](int, int co, int ci, int) { return cw.dp_w[co * 2 * C1 + ci]; });
            pos<typename P::ConvT>([&](int, int vo, int ci, int t) { return w_convt(vo, ci, t); });
        }
        if (off_ != A.total || (int)table_.size() != 2 * P::NCHUNK_FRAME) throw std::runtime_error("fe_pack: schedule mismatch");
        std::memcpy(&blob_[A.table], table_.data(), table_.size() * sizeof(int));
    }
};

template <class P> std::vector<float> pack_blob(const float* canonical) {
    std::vector<float> blob;
    Packer<P>(blob).pack(canonical);
    return blob;
}

}  // namespace fe
