/*
 * fe_oracle.c -- CPU restatement of FastEnhancer's per-frame hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA engine in
 * fastenhancer_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load it; the product path never does.
 *
 * Parity status: PINNED.  tools/gen_golden.py runs the reference itself (imported from
 * /root/reference on the build container) on seeded inputs and commits the outputs under
 * tests/golden/; tests/test_oracle.py checks this restatement against them.
 *
 * Plain C, float32 arithmetic like the reference's inference scripts.  Every function cites the
 * reference file:line (relative to /root/reference) whose behaviour it restates.  The weights are
 * the "canonical folded order" declared in include/fastenhancer_b200.h (BN / weight-norm already
 * folded: models/fastenhancer/default/model.py:532-608).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FEO_MAX_ENC 8
#define FEO_MAX_BLK 16

typedef struct {
    int n_fft, hop, c1, n_enc, c2, f2, n_blocks, n_heads;
    float compression;
} feo_config;

typedef struct {
    const float *w_ih, *w_hh, *b_ih, *b_hh, *fc_w, *fc_b, *pe, *qkv_w, *qkv_b, *afc_w, *afc_b;
} feo_block;

typedef struct feo_model {
    feo_config c;
    int fin, f1, hd;
    float *blob;                 /* owned copy of the canonical array */
    const float *enc_pre_w, *enc_pre_b;
    const float *enc_w[FEO_MAX_ENC], *enc_b[FEO_MAX_ENC];
    const float *rf_pre_lin, *rf_pre_w, *rf_pre_b;
    feo_block blk[FEO_MAX_BLK];
    const float *rf_post_lin, *rf_post_w, *rf_post_b;
    const float *dec_w1[FEO_MAX_ENC], *dec_b1[FEO_MAX_ENC], *dec_w2[FEO_MAX_ENC], *dec_b2[FEO_MAX_ENC];
    const float *dp_w, *dp_b, *dp_wt, *dp_bt;
    float *window, *window_istft;     /* [N] */
    float *tw_re, *tw_im;             /* FFT twiddles exp(-2 pi i k / N), k < N/2 */
    int *bitrev;                      /* [N] */
} feo_model;

/* ------------------------------------------------------------------------------------------ */
size_t feo_weight_count(const feo_config *c)
{
    size_t C1 = c->c1, C2 = c->c2, F1 = c->n_fft / 8, F2 = c->f2, n = 0;
    n += C1 * 16 + C1;
    n += (size_t)c->n_enc * (C1 * C1 * 3 + C1);
    n += F2 * F1 + C2 * C1 + C2;
    n += (size_t)c->n_blocks * (2 * 3 * C2 * C2 + 2 * 3 * C2 + C2 * C2 + C2 + 3 * C2 * C2 + 3 * C2 + C2 * C2 + C2);
    n += F2 * C2;                                   /* pe, block 0 only */
    n += F1 * F2 + C1 * C2 + C1;
    n += (size_t)c->n_enc * (C1 * 2 * C1 + C1 + C1 * C1 * 3 + C1);
    n += C1 * 2 * C1 + C1 + C1 * 16 + 2;
    return n;
}

size_t feo_state_floats(const feo_config *c)
{
    return 2 * (size_t)(c->n_fft - c->hop) + (size_t)c->n_blocks * c->f2 * c->c2;
}

void feo_destroy(feo_model *m)
{
    if (!m) return;
    free(m->blob); free(m->window); free(m->window_istft); free(m->tw_re); free(m->tw_im); free(m->bitrev);
    free(m);
}

feo_model *feo_create(const feo_config *c, const float *canonical, size_t n)
{
    if (n != feo_weight_count(c) || c->n_enc > FEO_MAX_ENC || c->n_blocks > FEO_MAX_BLK) return NULL;
    if (c->n_fft & (c->n_fft - 1)) return NULL;      /* radix-2 FFT below */
    feo_model *m = (feo_model *)calloc(1, sizeof(*m));
    m->c = *c;
    m->fin = c->n_fft / 2; m->f1 = m->fin / 4; m->hd = c->c2 / c->n_heads;
    m->blob = (float *)malloc(n * sizeof(float));
    memcpy(m->blob, canonical, n * sizeof(float));
    const float *p = m->blob;
    size_t C1 = c->c1, C2 = c->c2, F1 = m->f1, F2 = c->f2;
#define TAKE(dst, cnt) do { (dst) = p; p += (cnt); } while (0)
    TAKE(m->enc_pre_w, C1 * 16); TAKE(m->enc_pre_b, C1);
    for (int i = 0; i < c->n_enc; ++i) { TAKE(m->enc_w[i], C1 * C1 * 3); TAKE(m->enc_b[i], C1); }
    TAKE(m->rf_pre_lin, F2 * F1); TAKE(m->rf_pre_w, C2 * C1); TAKE(m->rf_pre_b, C2);
    for (int k = 0; k < c->n_blocks; ++k) {
        feo_block *b = &m->blk[k];
        TAKE(b->w_ih, 3 * C2 * C2); TAKE(b->w_hh, 3 * C2 * C2); TAKE(b->b_ih, 3 * C2); TAKE(b->b_hh, 3 * C2);
        TAKE(b->fc_w, C2 * C2); TAKE(b->fc_b, C2);
        if (k == 0) TAKE(b->pe, F2 * C2); else b->pe = NULL;
        TAKE(b->qkv_w, 3 * C2 * C2); TAKE(b->qkv_b, 3 * C2);
        TAKE(b->afc_w, C2 * C2); TAKE(b->afc_b, C2);
    }
    TAKE(m->rf_post_lin, F1 * F2); TAKE(m->rf_post_w, C1 * C2); TAKE(m->rf_post_b, C1);
    for (int i = 0; i < c->n_enc; ++i) {
        TAKE(m->dec_w1[i], C1 * 2 * C1); TAKE(m->dec_b1[i], C1); TAKE(m->dec_w2[i], C1 * C1 * 3); TAKE(m->dec_b2[i], C1);
    }
    TAKE(m->dp_w, C1 * 2 * C1); TAKE(m->dp_b, C1); TAKE(m->dp_wt, C1 * 16); TAKE(m->dp_bt, 2);
#undef TAKE
    /* periodic Hann window: torch.hann_window(win_size) -- functional/audio_modules.py:213-214 */
    int N = c->n_fft, H = c->hop;
    m->window = (float *)malloc(N * sizeof(float));
    m->window_istft = (float *)malloc(N * sizeof(float));
    for (int i = 0; i < N; ++i) m->window[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * i / N));
    /* window_istft = window / steady-state sum of shifted window^2 -- audio_modules.py:221-234 */
    for (int i = 0; i < N; ++i) {
        float s = 0.f;
        for (int j = i % H; j < N; j += H) s += m->window[j] * m->window[j];
        m->window_istft[i] = m->window[i] / s;
    }
    m->tw_re = (float *)malloc(N / 2 * sizeof(float));
    m->tw_im = (float *)malloc(N / 2 * sizeof(float));
    for (int k = 0; k < N / 2; ++k) { m->tw_re[k] = (float)cos(-2.0 * M_PI * k / N); m->tw_im[k] = (float)sin(-2.0 * M_PI * k / N); }
    m->bitrev = (int *)malloc(N * sizeof(int));
    int bits = 0; while ((1 << bits) < N) ++bits;
    for (int i = 0; i < N; ++i) { int r = 0; for (int b = 0; b < bits; ++b) if (i & (1 << b)) r |= 1 << (bits - 1 - b); m->bitrev[i] = r; }
    return m;
}

/* ---- complex radix-2 FFT of size N (torch.fft.rfft / ifft are the library calls restated) -- */
static void fft_inplace(const feo_model *m, float *re, float *im, int inverse)
{
    int N = m->c.n_fft;
    for (int i = 0; i < N; ++i) {
        int j = m->bitrev[i];
        if (j > i) { float t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
    }
    for (int len = 2; len <= N; len <<= 1) {
        int half = len >> 1, step = N / len;
        for (int i = 0; i < N; i += len)
            for (int k = 0; k < half; ++k) {
                float wr = m->tw_re[k * step], wi = inverse ? -m->tw_im[k * step] : m->tw_im[k * step];
                float xr = re[i + k + half], xi = im[i + k + half];
                float tr = xr * wr - xi * wi, ti = xr * wi + xi * wr;
                re[i + k + half] = re[i + k] - tr; im[i + k + half] = im[i + k] - ti;
                re[i + k] += tr; im[i + k] += ti;
            }
    }
}

/* windowed real FFT of one frame -> N/2+1 bins.  ONNXSTFT.forward: audio_modules.py:250-251 */
static void frame_rfft(const feo_model *m, const float *frame, float *spec_re, float *spec_im, float *wr, float *wi)
{
    int N = m->c.n_fft;
    for (int i = 0; i < N; ++i) { wr[i] = frame[i] * m->window[i]; wi[i] = 0.f; }
    fft_inplace(m, wr, wi, 0);
    for (int k = 0; k <= N / 2; ++k) { spec_re[k] = wr[k]; spec_im[k] = wi[k]; }
}

/* irfft(n=N) of N/2+1 bins, imag of DC and Nyquist ignored.  ONNXSTFT.inverse:
 * audio_modules.py:285-296 (zero-padded ifft + correction == irfft) and torch.istft's irfft. */
static void frame_irfft(const feo_model *m, const float *spec_re, const float *spec_im, float *out, float *wr, float *wi)
{
    int N = m->c.n_fft;
    wr[0] = spec_re[0]; wi[0] = 0.f;
    wr[N / 2] = spec_re[N / 2]; wi[N / 2] = 0.f;
    for (int k = 1; k < N / 2; ++k) { wr[k] = spec_re[k]; wi[k] = spec_im[k]; wr[N - k] = spec_re[k]; wi[N - k] = -spec_im[k]; }
    fft_inplace(m, wr, wi, 1);
    float inv = 1.0f / (float)N;
    for (int i = 0; i < N; ++i) out[i] = wr[i] * inv;
}

/* ---- elementwise helpers ---------------------------------------------------------------- */
static inline float silu(float x) { return x / (1.0f + expf(-x)); }        /* nn.SiLU */
static inline float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

/* power-law compression of N/2 bins (Nyquist dropped).
 * ONNXModel.forward head: model.py:684-690; CompressedSTFT.forward: audio_modules.py:150-154 */
static void compress(const feo_model *m, const float *re, const float *im, float *xc /* [2][fin] */)
{
    int fin = m->fin; float e = m->c.compression - 1.0f;
    for (int k = 0; k < fin; ++k) {
        float mag = sqrtf(re[k] * re[k] + im[k] * im[k]);
        if (mag < 1.0e-5f) mag = 1.0e-5f;
        float g = powf(mag, e);
        xc[k] = re[k] * g; xc[fin + k] = im[k] * g;
    }
}

/* Conv1d(k=3, pad=1) + bias (+SiLU) along frequency.  model.py:446-456 (encoder), :499-504 */
static void conv3(const float *w, const float *b, const float *x, float *y, int cin, int cout, int F, int act)
{
    for (int co = 0; co < cout; ++co) {
        float *yr = y + (size_t)co * F;
        for (int f = 0; f < F; ++f) yr[f] = b[co];
        for (int ci = 0; ci < cin; ++ci) {
            const float *xr = x + (size_t)ci * F;
            const float *wk = w + ((size_t)co * cin + ci) * 3;
            float w0 = wk[0], w1 = wk[1], w2 = wk[2];
            for (int f = 1; f < F; ++f) yr[f] += w0 * xr[f - 1];
            for (int f = 0; f < F; ++f) yr[f] += w1 * xr[f];
            for (int f = 0; f < F - 1; ++f) yr[f] += w2 * xr[f + 1];
        }
        if (act) for (int f = 0; f < F; ++f) yr[f] = silu(yr[f]);
    }
}

/* Conv1d(k=1) + bias (+SiLU) over the channel-concatenation [xa ; xb].  model.py:496, :517,
 * torch.cat at :663 / :669 (x channels first, skip second). */
static void conv1_cat(const float *w, const float *b, const float *xa, int ca, const float *xb, int cb,
                      float *y, int cout, int F, int act)
{
    int cin = ca + cb;
    for (int co = 0; co < cout; ++co) {
        float *yr = y + (size_t)co * F;
        for (int f = 0; f < F; ++f) yr[f] = b ? b[co] : 0.f;
        for (int ci = 0; ci < cin; ++ci) {
            const float *xr = ci < ca ? xa + (size_t)ci * F : xb + (size_t)(ci - ca) * F;
            float wv = w[(size_t)co * cin + ci];
            for (int f = 0; f < F; ++f) yr[f] += wv * xr[f];
        }
        if (act) for (int f = 0; f < F; ++f) yr[f] = silu(yr[f]);
    }
}

/* enc_pre: StridedConv1d(2 -> C1, k=8, s=4, pad=2) + folded BN + SiLU.  model.py:15-59, :436-443.
 * Restated literally: pad, fold the stride into channels (index phase*2 + c), conv k=2. */
static void enc_pre(const feo_model *m, const float *xc, float *y, float *xf /* [8][f1+1] */)
{
    int fin = m->fin, F1 = m->f1, C1 = m->c.c1, M = F1 + 1;
    for (int phase = 0; phase < 4; ++phase)
        for (int c = 0; c < 2; ++c)
            for (int mm = 0; mm < M; ++mm) {
                int i = 4 * mm + phase - 2;                 /* index into the unpadded input */
                xf[(phase * 2 + c) * M + mm] = (i >= 0 && i < fin) ? xc[c * fin + i] : 0.f;
            }
    for (int co = 0; co < C1; ++co) {
        float *yr = y + (size_t)co * F1;
        for (int j = 0; j < F1; ++j) yr[j] = m->enc_pre_b[co];
        for (int q = 0; q < 8; ++q)
            for (int t = 0; t < 2; ++t) {
                float wv = m->enc_pre_w[(co * 8 + q) * 2 + t];
                const float *xr = xf + q * M + t;
                for (int j = 0; j < F1; ++j) yr[j] += wv * xr[j];
            }
        for (int j = 0; j < F1; ++j) yr[j] = silu(yr[j]);
    }
}

/* dec_post tail: ConvTranspose1d(C1 -> 2, k=8, s=4, pad=2) + bias.  model.py:62-95, :510-515 */
static void conv_transpose(const feo_model *m, const float *x, float *mask /* [2][fin] */)
{
    int fin = m->fin, F1 = m->f1, C1 = m->c.c1;
    for (int o = 0; o < 2; ++o) for (int n = 0; n < fin; ++n) mask[o * fin + n] = m->dp_bt[o];
    for (int ci = 0; ci < C1; ++ci)
        for (int j = 0; j < F1; ++j) {
            float xv = x[(size_t)ci * F1 + j];
            for (int o = 0; o < 2; ++o)
                for (int k = 0; k < 8; ++k) {
                    int n = 4 * j + k - 2;
                    if (n >= 0 && n < fin) mask[o * fin + n] += xv * m->dp_wt[(ci * 2 + o) * 8 + k];
                }
        }
}

/* y[r][c] = b[r]? ... plain dense layer on channels-last rows: y[f][co] = b[co] + sum_ci W[co][ci] x[f][ci] */
static void linear_rows(const float *w, const float *b, const float *x, float *y, int rows, int cin, int cout)
{
    for (int f = 0; f < rows; ++f)
        for (int co = 0; co < cout; ++co) {
            float s = b ? b[co] : 0.f;
            const float *wr = w + (size_t)co * cin, *xr = x + (size_t)f * cin;
            for (int ci = 0; ci < cin; ++ci) s += wr[ci] * xr[ci];
            y[(size_t)f * cout + co] = s;
        }
}

/* One RNNFormer block, one frame, one stream.  x: [F2][C2] channels-last, h: [F2][C2].
 * RNNFormerBlock.forward: model.py:266-291; GRU = torch nn.GRU single step, gate order r,z,n;
 * Attention.forward: model.py:142-152 (per-head interleaved q|k|v rows, SDPA scale hd^-0.5). */
static void rf_block(const feo_model *m, const feo_block *b, float *x, float *h, float *ws)
{
    int F2 = m->c.f2, C2 = m->c.c2, NH = m->c.n_heads, hd = m->hd;
    float *gi = ws, *gh = gi + (size_t)F2 * 3 * C2, *t0 = gh + (size_t)F2 * 3 * C2, *qkv = t0 + (size_t)F2 * C2;
    float *att = qkv + (size_t)F2 * 3 * C2, *sc = att + (size_t)F2 * C2;
    linear_rows(b->w_ih, b->b_ih, x, gi, F2, C2, 3 * C2);
    linear_rows(b->w_hh, b->b_hh, h, gh, F2, C2, 3 * C2);
    for (int f = 0; f < F2; ++f)
        for (int c = 0; c < C2; ++c) {
            const float *a = gi + (size_t)f * 3 * C2, *g = gh + (size_t)f * 3 * C2;
            float r = sigmoidf(a[c] + g[c]);
            float z = sigmoidf(a[C2 + c] + g[C2 + c]);
            float n = tanhf(a[2 * C2 + c] + r * g[2 * C2 + c]);
            float hp = h[(size_t)f * C2 + c];
            h[(size_t)f * C2 + c] = (1.0f - z) * n + z * hp;
        }
    linear_rows(b->fc_w, b->fc_b, h, t0, F2, C2, C2);                 /* rnn_fc + folded BN */
    for (int i = 0; i < F2 * C2; ++i) x[i] += t0[i];                   /* residual, model.py:277 */
    if (b->pe) for (int i = 0; i < F2 * C2; ++i) x[i] += b->pe[i];     /* model.py:279-280 */
    linear_rows(b->qkv_w, b->qkv_b, x, qkv, F2, C2, 3 * C2);
    float scale = 1.0f / sqrtf((float)hd);
    for (int hh = 0; hh < NH; ++hh) {
        int base = hh * 3 * hd;
        for (int i = 0; i < F2; ++i) {
            const float *q = qkv + (size_t)i * 3 * C2 + base;
            float mx = -INFINITY;
            for (int j = 0; j < F2; ++j) {
                const float *kk = qkv + (size_t)j * 3 * C2 + base + hd;
                float s = 0.f;
                for (int d = 0; d < hd; ++d) s += q[d] * kk[d];
                sc[j] = s * scale; if (sc[j] > mx) mx = sc[j];
            }
            float den = 0.f;
            for (int j = 0; j < F2; ++j) { sc[j] = expf(sc[j] - mx); den += sc[j]; }
            for (int d = 0; d < hd; ++d) {
                float s = 0.f;
                for (int j = 0; j < F2; ++j) s += sc[j] * qkv[(size_t)j * 3 * C2 + base + 2 * hd + d];
                att[(size_t)i * C2 + hh * hd + d] = s / den;
            }
        }
    }
    linear_rows(b->afc_w, b->afc_b, att, t0, F2, C2, C2);             /* attn_fc + folded BN */
    for (int i = 0; i < F2 * C2; ++i) x[i] += t0[i];                   /* residual, model.py:290 */
}

/* ---- tap layout (debug / stage-by-stage parity) -------------------------------------------- */
size_t feo_tap_floats(const feo_config *c)
{
    size_t fin = c->n_fft / 2, F1 = fin / 4, C1 = c->c1, C2 = c->c2, F2 = c->f2;
    return 2 * fin + (1 + c->n_enc) * C1 * F1 + F2 * C2 + (size_t)c->n_blocks * 3 * F2 * C2 + C1 * F1
           + (size_t)c->n_enc * C1 * F1 + 2 * fin + 2 * fin;
}

/* Per-frame model core: compressed spectrum [2][fin] -> mask [2][fin], GRU state h [K][F2][C2]
 * updated in place.  ONNXModel.model_forward with T=1: model.py:620-675.
 * taps (may be NULL) receives, in order: spec_c, enc_pre, enc[i].., rf_pre (channels-last),
 * per block {x after GRU half, x after attention half, h_new}, rf_post, dec[i].., mask. */
static float *core(const feo_model *m, const float *xc, float *h, float *mask, float *ws, float *taps)
{
    int fin = m->fin, F1 = m->f1, C1 = m->c.c1, C2 = m->c.c2, F2 = m->c.f2, E = m->c.n_enc;
    size_t act = (size_t)C1 * F1;
    float *skips = ws;                       /* (E+1) x [C1][F1] */
    float *a = skips + (E + 1) * act, *bb = a + act;
    float *xr = bb + act;                    /* [F2][C2] */
    float *tmp = xr + (size_t)F2 * C2;       /* [max(C1,C2)][max(F1,F2)] */
    size_t tmpn = (size_t)(C1 > C2 ? C1 : C2) * (F1 > F2 ? F1 : F2);
    float *tmp2 = tmp + tmpn;
    float *blkws = tmp2 + tmpn;
#define TAP(src, n) do { if (taps) { memcpy(taps, (src), (n) * sizeof(float)); taps += (n); } } while (0)
    TAP(xc, 2 * (size_t)fin);
    enc_pre(m, xc, skips, blkws);
    TAP(skips, act);
    for (int i = 0; i < E; ++i) {
        conv3(m->enc_w[i], m->enc_b[i], skips + i * act, skips + (i + 1) * act, C1, C1, F1, 1);
        TAP(skips + (i + 1) * act, act);
    }
    /* rf_pre: Linear(F1->F2) on the frequency axis, Conv1d(C1->C2,1)+BN.  model.py:459-465, :646 */
    const float *xe = skips + E * act;
    for (int c = 0; c < C1; ++c)
        for (int f2 = 0; f2 < F2; ++f2) {
            float s = 0.f;
            for (int f = 0; f < F1; ++f) s += m->rf_pre_lin[(size_t)f2 * F1 + f] * xe[(size_t)c * F1 + f];
            tmp[(size_t)c * F2 + f2] = s;
        }
    conv1_cat(m->rf_pre_w, m->rf_pre_b, tmp, C1, NULL, 0, tmp2, C2, F2, 0);
    for (int c = 0; c < C2; ++c) for (int f = 0; f < F2; ++f) xr[(size_t)f * C2 + c] = tmp2[(size_t)c * F2 + f];   /* model.py:647-650 */
    TAP(xr, (size_t)F2 * C2);
    for (int k = 0; k < m->c.n_blocks; ++k) {
        float *hk = h + (size_t)k * F2 * C2;
        float *xin_copy = tmp;                                 /* tmp is free here */
        memcpy(xin_copy, xr, (size_t)F2 * C2 * sizeof(float));
        rf_block(m, &m->blk[k], xr, hk, blkws);
        if (taps) {
            /* mid-block tap (x after the GRU half) = x_in + rnn_fc(h_new) (+pe), model.py:273-280 */
            float *t0 = tmp2;
            linear_rows(m->blk[k].fc_w, m->blk[k].fc_b, hk, t0, F2, C2, C2);
            for (int i = 0; i < F2 * C2; ++i) {
                float v = xin_copy[i] + t0[i];
                if (m->blk[k].pe) v += m->blk[k].pe[i];
                taps[i] = v;
            }
            taps += (size_t)F2 * C2;
            TAP(xr, (size_t)F2 * C2);
            TAP(hk, (size_t)F2 * C2);
        }
    }
    /* rf_post: Linear(F2->F1), Conv1d(C2->C1,1)+BN.  model.py:486-490, :654-656 */
    for (int c = 0; c < C2; ++c)
        for (int f = 0; f < F1; ++f) {
            float s = 0.f;
            for (int f2 = 0; f2 < F2; ++f2) s += m->rf_post_lin[(size_t)f * F2 + f2] * xr[(size_t)f2 * C2 + c];
            tmp[(size_t)c * F1 + f] = s;
        }
    conv1_cat(m->rf_post_w, m->rf_post_b, tmp, C2, NULL, 0, a, C1, F1, 0);
    TAP(a, act);
    /* decoder: cat([x, skip.pop()]) -> 1x1 + SiLU -> k3 + SiLU.  model.py:493-506, :661-666 */
    for (int i = 0; i < E; ++i) {
        conv1_cat(m->dec_w1[i], m->dec_b1[i], a, C1, skips + (size_t)(E - i) * act, C1, bb, C1, F1, 1);
        conv3(m->dec_w2[i], m->dec_b2[i], bb, a, C1, C1, F1, 1);
        TAP(a, act);
    }
    /* dec_post: cat([x, enc_pre_out]) -> 1x1 + SiLU -> transposed conv.  model.py:508-521, :669-670 */
    conv1_cat(m->dp_w, m->dp_b, a, C1, skips, C1, bb, C1, F1, 1);
    conv_transpose(m, bb, mask);
    TAP(mask, 2 * (size_t)fin);
#undef TAP
    return taps;
}

static size_t core_ws_floats(const feo_config *c)
{
    size_t fin = c->n_fft / 2, F1 = fin / 4, C1 = c->c1, C2 = c->c2, F2 = c->f2;
    size_t tmpn = (C1 > C2 ? C1 : C2) * (F1 > F2 ? F1 : F2);
    size_t blk = F2 * 3 * C2 * 3 + F2 * C2 * 2 + F2 + 8 * (F1 + 1);
    return (c->n_enc + 3) * C1 * F1 + F2 * C2 + 2 * tmpn + blk + 64;
}

/* compressed-domain mask application + decompression.
 * streaming: model.py:694-709 ; offline: model.py:732-733 + audio_modules.py:160-163 */
static void apply_mask(const feo_model *m, const float *xc, const float *mask, float *yc /* [2][fin] compressed */,
                       float *out_re, float *out_im /* [fin+1] decompressed, Nyquist 0 */)
{
    int fin = m->fin; float e = 1.0f / m->c.compression - 1.0f;
    for (int k = 0; k < fin; ++k) {
        float xr = xc[k], xi = xc[fin + k], mr = mask[k], mi = mask[fin + k];
        float yr = xr * mr - xi * mi, yi = xr * mi + xi * mr;
        yc[k] = yr; yc[fin + k] = yi;
        float g = powf(sqrtf(yr * yr + yi * yi), e);
        out_re[k] = yr * g; out_im[k] = yi * g;
    }
    out_re[fin] = 0.f; out_im[fin] = 0.f;
}

/* ---- streaming wav -> wav (scripts/export_onnx.py:48-58 composed with the loop at :130-136) -- */
void feo_stream(const feo_model *m, float *state, const float *wav_in, float *wav_out,
                int n_streams, int n_hops, int ld_in, int ld_out, int n_threads, float *taps_stream0)
{
    int N = m->c.n_fft, H = m->c.hop, CL = N - H, fin = m->fin;
    size_t sf = feo_state_floats(&m->c), wsn = core_ws_floats(&m->c);
    size_t tapn = feo_tap_floats(&m->c);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(static)
#endif
    for (int s = 0; s < n_streams; ++s) {
        float *ws = (float *)malloc((wsn + 6 * (size_t)N + 8 * (size_t)fin + 16) * sizeof(float));
        float *frame = ws + wsn, *wr = frame + N, *wi = wr + N, *sre = wi + N, *sim = sre + N, *y = sim + N;
        float *xc = y + N, *mask = xc + 2 * fin, *yc = mask + 2 * fin, *ore = yc + 2 * fin, *oim = ore + fin + 1;
        float *cache_stft = state + (size_t)s * sf, *cache_istft = cache_stft + CL, *h = cache_istft + CL;
        for (int hop = 0; hop < n_hops; ++hop) {
            const float *xin = wav_in + (size_t)s * ld_in + (size_t)hop * H;
            /* ONNXSTFT.forward: audio_modules.py:248-251 */
            memcpy(frame, cache_stft, CL * sizeof(float));
            memcpy(frame + CL, xin, H * sizeof(float));
            memcpy(cache_stft, frame + H, CL * sizeof(float));
            frame_rfft(m, frame, sre, sim, wr, wi);
            compress(m, sre, sim, xc);
            float *t = (taps_stream0 && s == 0) ? taps_stream0 + (size_t)hop * tapn : NULL;
            t = core(m, xc, h, mask, ws, t);
            apply_mask(m, xc, mask, yc, ore, oim);
            if (t) memcpy(t, yc, 2 * (size_t)fin * sizeof(float));
            /* ONNXSTFT.inverse: audio_modules.py:285-303 */
            frame_irfft(m, ore, oim, y, wr, wi);
            for (int i = 0; i < N; ++i) y[i] *= m->window_istft[i];
            for (int i = 0; i < CL; ++i) y[i] += cache_istft[i];
            memcpy(wav_out + (size_t)s * ld_out + (size_t)hop * H, y, H * sizeof(float));
            memcpy(cache_istft, y + H, CL * sizeof(float));
        }
        free(ws);
    }
}

/* ---- spec -> spec (ONNXModel.forward, model.py:677-710), T frames, h: [B][K][F2][C2] -------- */
void feo_spec(const feo_model *m, float *h, const float *spec_in, float *spec_out, int n_streams, int T, int n_threads)
{
    int fin = m->fin, NB = fin + 1;
    size_t hs = (size_t)m->c.n_blocks * m->c.f2 * m->c.c2, wsn = core_ws_floats(&m->c);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(static)
#endif
    for (int s = 0; s < n_streams; ++s) {
        float *ws = (float *)malloc((wsn + 10 * (size_t)NB) * sizeof(float));
        float *sre = ws + wsn, *sim = sre + NB, *xc = sim + NB, *mask = xc + 2 * fin, *yc = mask + 2 * fin, *ore = yc + 2 * fin, *oim = ore + NB;
        for (int t = 0; t < T; ++t) {
            for (int k = 0; k < NB; ++k) {       /* layout [B, N/2+1, T, 2] */
                sre[k] = spec_in[(((size_t)s * NB + k) * T + t) * 2];
                sim[k] = spec_in[(((size_t)s * NB + k) * T + t) * 2 + 1];
            }
            compress(m, sre, sim, xc);
            core(m, xc, h + (size_t)s * hs, mask, ws, NULL);
            apply_mask(m, xc, mask, yc, ore, oim);
            for (int k = 0; k < NB; ++k) {
                spec_out[(((size_t)s * NB + k) * T + t) * 2] = ore[k];
                spec_out[(((size_t)s * NB + k) * T + t) * 2 + 1] = oim[k];
            }
        }
        free(ws);
    }
}

/* ---- offline wav -> wav (Model.forward, model.py:728-735) ----------------------------------
 * STFT: torch.stft(center=True, pad_mode='reflect') audio_modules.py:78-80, T = 1 + L/H frames;
 * iSTFT: torch.istft(center=True): irfft * window, overlap-add, divide by the summed window^2,
 * drop N/2 samples at both ends -> length H*(T-1)  (audio_modules.py:115-119).
 * spec_out (may be NULL): compressed masked spectrum [B, N/2, T, 2]. */
void feo_offline(const feo_model *m, const float *wav, int n_streams, int L, float *wav_out, float *spec_out, int n_threads)
{
    int N = m->c.n_fft, H = m->c.hop, fin = m->fin, T = 1 + L / H;
    size_t hs = (size_t)m->c.n_blocks * m->c.f2 * m->c.c2, wsn = core_ws_floats(&m->c);
    size_t full = (size_t)N + (size_t)H * (T - 1);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(static)
#endif
    for (int s = 0; s < n_streams; ++s) {
        float *ws = (float *)malloc((wsn + 6 * (size_t)N + 8 * (size_t)fin + 16 + hs + 2 * full) * sizeof(float));
        float *frame = ws + wsn, *wr = frame + N, *wi = wr + N, *sre = wi + N, *sim = sre + N, *y = sim + N;
        float *xc = y + N, *mask = xc + 2 * fin, *yc = mask + 2 * fin, *ore = yc + 2 * fin, *oim = ore + fin + 1;
        float *h = oim + fin + 1 + 8, *acc = h + hs, *env = acc + full;
        memset(h, 0, hs * sizeof(float));
        memset(acc, 0, 2 * full * sizeof(float));
        const float *x = wav + (size_t)s * L;
        for (int t = 0; t < T; ++t) {
            for (int i = 0; i < N; ++i) {
                long j = (long)t * H + i - N / 2;
                if (j < 0) j = -j;
                if (j >= L) j = 2L * (L - 1) - j;
                frame[i] = x[j];
            }
            frame_rfft(m, frame, sre, sim, wr, wi);
            compress(m, sre, sim, xc);
            core(m, xc, h, mask, ws, NULL);
            apply_mask(m, xc, mask, yc, ore, oim);
            if (spec_out)
                for (int k = 0; k < fin; ++k) {
                    spec_out[(((size_t)s * fin + k) * T + t) * 2] = yc[k];
                    spec_out[(((size_t)s * fin + k) * T + t) * 2 + 1] = yc[fin + k];
                }
            frame_irfft(m, ore, oim, y, wr, wi);
            for (int i = 0; i < N; ++i) {
                acc[(size_t)t * H + i] += y[i] * m->window[i];
                env[(size_t)t * H + i] += m->window[i] * m->window[i];
            }
        }
        for (size_t i = 0; i < (size_t)H * (T - 1); ++i)
            wav_out[(size_t)s * H * (T - 1) + i] = acc[N / 2 + i] / env[N / 2 + i];
        free(ws);
    }
}

/* standalone front/back-end pieces for unit tests of the engine's STFT shims */
void feo_stft_frame(const feo_model *m, const float *frame, float *spec /* [N/2+1][2] */)
{
    int N = m->c.n_fft;
    float *b = (float *)malloc(4 * (size_t)N * sizeof(float));
    frame_rfft(m, frame, b + 2 * N, b + 3 * N, b, b + N);
    for (int k = 0; k <= N / 2; ++k) { spec[2 * k] = b[2 * N + k]; spec[2 * k + 1] = b[3 * N + k]; }
    free(b);
}

void feo_get_windows(const feo_model *m, float *window, float *window_istft)
{
    memcpy(window, m->window, m->c.n_fft * sizeof(float));
    memcpy(window_istft, m->window_istft, m->c.n_fft * sizeof(float));
}
