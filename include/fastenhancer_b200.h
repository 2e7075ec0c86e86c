/*
 * fastenhancer_b200.h -- C ABI of the B200-native FastEnhancer streaming engine.
 *
 * The reference (aask1357/fastenhancer) has no native ABI: its hot path is Python/PyTorch
 * (models/fastenhancer/default/model.py, functional/audio_modules.py) and its streaming boundary is
 * the ONNX graph of scripts/export_onnx.py:48-58 driven hop by hop from scripts/test_onnx.py:44-49.
 * Each entry point below names the reference interface it replaces.  Plain pointers and sizes
 * only; all `device` pointers are CUDA device memory of the engine's device, `cuda_stream` is a
 * cudaStream_t passed as void* (NULL = default stream).
 *
 * Every function returns 0 on success and a negative code on failure; fe_last_error() returns a
 * thread-local description.  One engine per device.  An engine is immutable after fe_create apart from
 * its settings (fe_set_*); everything a call mutates lives in the fe_state it is given or on the call's
 * CUDA stream, so distinct states of one engine may be driven concurrently from different host threads
 * and streams.  Calls on ONE state are serialised by the caller.  No global state.
 */
#ifndef FASTENHANCER_B200_H
#define FASTENHANCER_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Shape of one FastEnhancer model: the `model_kwargs` block of the reference YAMLs
 * (configs/fastenhancer/b.yaml:1-29) reduced to what the hot path needs. */
typedef struct fe_config {
    int n_fft, hop;         /* STFT: n_fft == win_size, hop_size */
    int c1, n_enc;          /* encoder channels, number of k=3 encoder blocks (len(kernel_size) - 1) */
    int c2, f2;             /* rnnformer channels / freq */
    int n_blocks, n_heads;  /* rnnformer num_blocks / num_heads */
    float compression;      /* input_compression (0.3) */
} fe_config;

typedef struct fe_engine fe_engine;   /* weights + kernels of one model on one device */
typedef struct fe_state fe_state;     /* recurrent + overlap state of n_streams independent streams */

enum {
    FE_OK = 0,
    FE_ERR_ARG = -1,          /* bad argument */
    FE_ERR_UNSUPPORTED = -2,  /* model shape not among the shipped FastEnhancer configurations */
    FE_ERR_CUDA = -3,         /* CUDA runtime error (see fe_last_error) */
    FE_ERR_NO_DEVICE = -4     /* no CUDA device: there is no CPU fallback */
};

const char* fe_last_error(void);

/* Number of floats of the canonical folded weight array / of the per-stream state for `cfg`.
 * Canonical order = what ONNXModel.remove_weight_reparameterizations leaves behind
 * (model.py:532-608), flattened as fastenhancer_b200/schema.py::canonical_schema lists it. */
size_t fe_weight_count(const fe_config* cfg);
size_t fe_state_floats(const fe_config* cfg);   /* 2*(n_fft-hop) + n_blocks*f2*c2 */

/* Replaces: building ONNXModel(**model_kwargs) + load_state_dict + remove_weight_reparameterizations
 * (scripts/export_onnx.py:60-78).  `canonical` is HOST memory, copied. */
int fe_create(const fe_config* cfg, const float* canonical, size_t n_floats, int device, fe_engine** out);
void fe_destroy(fe_engine* e);

/* Checkpoint ingestion on the device.  Replaces: ModelWrapper.load() + remove_weight_reparameterizations()
 * (wrappers/ns.py:308-321, model.py:532-608, block part :215-258, final transposed conv :74-81) without the reference classes.
 * One rule per tensor of the canonical array; every source pointer is DEVICE memory holding a pre-fold parameter of
 * ckpt['model'] (element order of the canonical tensor = element order of the source weight):
 *   FE_FOLD_COPY         dst = w                                               (filterbanks, biases, positional embedding)
 *   FE_FOLD_WEIGHT_NORM  dst[r][:] = a[r] * w[r][:] / ||w[r]||                 (w = original1, a = original0: GRU weights, qkv)
 *   FE_FOLD_BATCH_NORM   dst[r][:] = w[r][:] * a[r] / sqrt(d[r] + eps);  canonical[bias_dst + r] = b[r] - c[r] * a[r] / sqrt(d[r] + eps)
 *                        (a = gamma, b = beta, c = running_mean, d = running_var; bias_dst < 0: no bias output)
 *   FE_FOLD_FINAL_CONV   dst = a[0] * w / max(||w||_F, 1e-12) when flag != 0 (normalize_final_conv), else a[0] * w
 * Scale factors are computed in double and the product rounded once (fastenhancer_b200/fold.py is the host oracle). */
enum { FE_FOLD_COPY = 0, FE_FOLD_WEIGHT_NORM = 1, FE_FOLD_BATCH_NORM = 2, FE_FOLD_FINAL_CONV = 3 };
typedef struct fe_fold_op {
    int kind, rows, cols, flag;
    const float *w, *a, *b, *c, *d;
    long long dst, bias_dst;      /* float offsets into the canonical array */
    float eps;
} fe_fold_op;
int fe_fold_device(const fe_fold_op* ops, int n_ops, float* canonical_device, void* cuda_stream);
int fe_create_from_device(const fe_config* cfg, const float* canonical_device, size_t n_floats, int device, fe_engine** out);

/* Replaces: ONNXSTFT.initialize_cache + ONNXModel.initialize_cache (audio_modules.py:238-241,
 * model.py:614-618): zeroed cache_stft [n, n_fft-hop], cache_istft [n, n_fft-hop], h_k [n*f2, c2] x K. */
int fe_state_create(fe_engine* e, int n_streams, fe_state** out);
/* Same on a caller-owned device buffer of n_streams * fe_state_floats floats (16-byte aligned, zeroed by the caller or by
 * fe_state_reset).  The kernels keep the state there as PLANES in the reference's own cache shapes,
 *   cache_stft [n][n_fft-hop] | cache_istft [n][n_fft-hop] | h_0 [n][f2][c2] | ... | h_{K-1} [n][f2][c2],
 * so a host wrapper can hand the pieces out as zero-copy tensors (`cache_in_* / cache_out_*` of scripts/test_onnx.py:34-49)
 * and a streaming step needs no cache import / export.  fe_state_planes returns that buffer (owned or not). */
int fe_state_create_on(fe_engine* e, int n_streams, float* planes_device, fe_state** out);
float* fe_state_planes(fe_state* s);
void fe_state_destroy(fe_state* s);
int fe_state_reset(fe_state* s, void* cuda_stream);
/* Reference cache layout per stream: [cache_stft | cache_istft | h_0 [f2][c2] | ... | h_{K-1}], so ORT-style
 * callers that round-trip `cache_in_* / cache_out_*` tensors (scripts/test_onnx.py:34-49) still can. */
int fe_state_export(fe_state* s, float* dst_device, void* cuda_stream);
int fe_state_import(fe_state* s, const float* src_device, void* cuda_stream);

/* Replaces: `n_hops` iterations of the streaming graph export_onnx.Model.forward
 * (scripts/export_onnx.py:48-58; loop at scripts/test_onnx.py:44-49) for n_streams independent streams.
 * wav_in / wav_out: device, [n_streams][ld] floats, the first n_hops*hop columns are read / written.
 * Output sample n of hop i is input time i*hop + n - (n_fft - hop) (docs/docs/onnx.md:37-72).
 * One persistent kernel launch; state stays on chip between the hops of the call. */
int fe_stream(fe_engine* e, fe_state* s, const float* wav_in, float* wav_out, int n_hops,
              long long ld_in, long long ld_out, void* cuda_stream);

/* Same, HOST buffers (pinned or pageable): host->device copy, kernel and device->host copy are pipelined in
 * `hops_per_chunk`-hop pieces (0 = default) on streams owned by `s`.  Ordered after the work already queued on
 * `cuda_stream`; returns once the last device->host copy has completed (also drains on error).
 * fe_state_reserve_host pre-allocates the staging buffers / streams / events so that no allocation happens inside
 * fe_stream_host (otherwise they are created on first use).  This is the end-to-end path bench.py times as `e2e`. */
int fe_state_reserve_host(fe_state* s, int hops_per_chunk);
int fe_stream_host(fe_engine* e, fe_state* s, const float* wav_in_host, float* wav_out_host, int n_hops,
                   long long ld_in, long long ld_out, int hops_per_chunk, void* cuda_stream);

/* Replaces: ONNXModel.forward(spec_noisy, *h) (model.py:677-710; the spec2spec graph of
 * scripts/export_onnx_spec.py:135-142).  spec_in / spec_out: device, [n_streams][n_fft/2+1][T][2]. */
int fe_spec(fe_engine* e, fe_state* s, const float* spec_in, float* spec_out, int T, void* cuda_stream);

/* Replace: ONNXSTFT.forward(x, cache) / ONNXSTFT.inverse(spec, cache) (functional/audio_modules.py:243-303) on their own,
 * for callers that compose the streaming graph themselves (scripts/export_onnx.py:53-57).  They use the cache_stft /
 * cache_istft parts of `s`.  wav: [n_streams][ld] device; spec: [n_streams][n_fft/2+1][n_hops][2] device
 * (fe_stft writes all n_fft/2+1 bins; fe_istft ignores the imaginary parts of the DC and Nyquist bins, like irfft). */
int fe_stft(fe_engine* e, fe_state* s, const float* wav_in, float* spec_out, int n_hops, long long ld_in, void* cuda_stream);
int fe_istft(fe_engine* e, fe_state* s, const float* spec_in, float* wav_out, int n_hops, long long ld_out, void* cuda_stream);

/* The STFT as a tensor-core GEMM -- the reference's ConvSTFT front end (models/fastenhancer/conv_stft/model.py:55-63, 110-114:
 * F.conv1d of the waveform with the windowed DFT basis, stride = hop) for callers that transform many frames at once:
 * spec_out [B][n_fft/2+1][T][2] (device), frame t of utterance b = wav[b][t*hop .. t*hop + n_fft - 1] (wav [B][ld] device, 16-byte
 * aligned, ld % 4 == 0, ld >= (T-1)*hop + n_fft; no centering, no cache), periodic Hann window.  accurate != 0: fp32-accurate 3xTF32
 * (hi / lo split of both operands); 0: one TF32 pass.  An alternative to fe_stft's in-kernel FFT, not used by the fused path. */
int fe_stft_gemm(fe_engine* e, const float* wav, int B, long long ld, int T, float* spec_out, int accurate, void* cuda_stream);

/* Replaces: Model.forward(noisy) (model.py:728-735): wav [B][L] -> wav_out [B][hop*(L/hop)] and (optional)
 * the compressed masked spectrum spec_out [B][n_fft/2][1 + L/hop][2].  Zero initial GRU state. */
int fe_offline(fe_engine* e, const float* wav, int B, int L, float* wav_out, float* spec_out, void* cuda_stream);
/* Schedule of fe_offline.  0 (default) = automatic; 1 = sequential walk: one CTA per group of utterances steps through the frames
 * (best for large batches); 2 = frame-parallel: outside the GRU recurrence (nn.GRU over time, model.py:266-272) the frames of an
 * utterance are independent, so the CTAs take groups of FRAMES -- front end, encoder, rf_pre and the input half of the GRU for all
 * frames at once, the recurrence as a scan (one CTA per (utterance, sub-band) row), then attention / decoder / inverse FFT for all
 * frames at once, then one overlap-add pass -- which is what Model.forward(noisy [B, L]) on a few long utterances needs to fill the GPU
 * (reference: models/fastenhancer/default/model.py:711-735).  fp32-accurate precision modes only (it runs the fp32 kernel family);
 * automatic mode picks it when the walk would occupy fewer than half of the SMs.  Results equal the walk's up to fp32 summation order. */
int fe_set_offline_mode(fe_engine* e, int mode);

/* Arithmetic of the channel contractions (conv-type layers, RNNFormer linears, GRU matrix products).
 * The DEFAULT of a new engine reproduces the fp32 reference: mode 4 where the model has such kernels (T / B), else mode 1.
 *   mode = 4: "fp32x3" -- fp32-accurate on the tcgen05 tensor cores: every operand is held as two fp16 parts
 *             (hi = fp16(v), lo = fp16(v - hi): 22 significand bits) and every product is three kind::f16 MMAs
 *             (hi*hi + lo*hi + hi*lo) into one fp32 TMEM accumulator; waveform error vs the fp32 reference ~7e-8 RMS,
 *             the same as mode 1.  FE_ERR_UNSUPPORTED for models without such a kernel variant.
 *   mode = 1: every multiply-add on the fp32 FMA pipe (~6e-8 RMS);
 *   mode = 0: tcgen05 tensor cores, TF32 operands (round-to-nearest), fp32 accumulation in TMEM --
 *             what PyTorch itself does for cuDNN convolutions by default (allow_tf32); waveform error vs the
 *             fp32 reference ~7e-6 RMS on the synthetic checkpoint whose mask is near identity and up to ~1e-3
 *             relative when the mask is network-dominated (tests/test_gpu_parity.py), against the 1e-4 RMS bar;
 *   mode = 2: as 0, with the activations and weights of the conv section (encoder, decoder, 1x1 convs, mask head)
 *             stored as fp16 -- the same 11-bit significand as TF32, twice the contraction length per MMA and half the
 *             shared memory; where the RNNFormer operands live in tensor memory (T/B/S) they are fp16 as well, beside an
 *             fp32 master of the GRU state; the RNNFormer of M/L stays TF32.  Same waveform error as mode 0.
 *             FE_ERR_UNSUPPORTED for models without such a kernel variant.
 *   mode = 3: "bf16 conv / fp32 GRU" (BASELINE.json configs[2]): as 2 with bfloat16 (8-bit significand) activations and
 *             weights in the conv section; the RNNFormer (GRU, linears, qkv) keeps TF32 operands, the GRU state is fp32.
 *             Waveform error ~3e-5 RMS on the near-identity-mask checkpoint.  FE_ERR_UNSUPPORTED without such a variant.
 * Attention, FFTs, (de)compression, the residual stream and all state are fp32 in every mode. */
int fe_set_precision(fe_engine* e, int mode);
int fe_get_precision(fe_engine* e);               /* 0 TF32, 1 fp32 FMA pipe, 2 fp16, 3 bf16 conv section, 4 fp32x3 */

/* Introspection used by the host wrapper, tests and bench. */
/* Hop-sliced streaming launches (default on): when fe_stream has more stream groups than the GPU has SMs, the launch is cut into
 * hop ranges and the items (range, stream group) run on one persistent CTA per SM, the state of a group handed from range to range
 * through global memory -- the last round of CTAs no longer leaves SMs idle (256 groups on 148 SMs: 1.77 instead of 2 rounds).
 * Available in the kernel variants without hop-tiled rings (M / L: the configs with one or two streams per CTA).
 * Results are bit-identical to the unsliced launch. */
int fe_set_hop_slicing(fe_engine* e, int on);
int fe_plan_hop_slices(int n_groups, int n_hops, int num_sms);   /* hops per range the launch would use (0 = unsliced); no device needed */
int fe_streams_per_cta(fe_engine* e, int n_streams);            /* kernel variant the engine would pick */
int fe_set_streams_per_cta(fe_engine* e, int s);                /* force a variant (0 = automatic) */
long long fe_kernel_launches(fe_engine* e);                     /* kernels of this library launched so far (fused kernel, GRU scan, overlap-add) */
int fe_tap_floats(fe_engine* e);
/* Test hook: like fe_stream, additionally dumps the per-stage tensors of stream 0 at hop `tap_hop`
 * (layout of oracle/fe_oracle.c::core) into taps_device [fe_tap_floats]. */
int fe_stream_taps(fe_engine* e, fe_state* s, const float* wav_in, float* wav_out, int n_hops,
                   long long ld_in, long long ld_out, float* taps_device, int tap_hop, void* cuda_stream);

/* The audio front door of the directory-level runner, on the device.  Replace: librosa.load(path, sr=wrapper.sr, mono=True) and
 * soundfile.write(path, enhanced, fs) around the model in scripts/test_pytorch.py:29,37 --
 *   fe_pcm16_to_float   interleaved int16 PCM [n_frames][n_channels] -> mono float32 [n_frames] (mean of the channels / 32768)
 *   fe_resample_poly    rational-ratio polyphase FIR resampling, out[m] = sum_j in[j] * taps[m*down - j*up + (n_taps-1)/2]
 *                       (scipy.signal.resample_poly semantics; n_taps odd, taps already scaled by `up`)
 *   fe_float_to_pcm16   float32 -> int16 PCM, round to nearest, clipped (soundfile's default WAV subtype)
 * All pointers are device memory. */
int fe_pcm16_to_float(const short* pcm_device, long long n_frames, int n_channels, float* wav_device, void* cuda_stream);
int fe_resample_poly(const float* in_device, long long n_in, int up, int down, const float* taps_device, int n_taps,
                     float* out_device, long long n_out, void* cuda_stream);
int fe_float_to_pcm16(const float* wav_device, long long n, short* pcm_device, void* cuda_stream);

/* Measured fp32 FMA-pipe throughput of `device` in TFLOP/s (a short FFMA microbenchmark): the roofline denominator of the fp32
 * FMA-pipe kernel family (mode 1) in bench.py. */
int fe_microbench_fma(int device, double* tflops);

/* Profiling hook: when set, CTA 0 of every fused-kernel launch accumulates the SM cycles it spends in each
 * phase of the frame (ids: enum PhaseId in fastenhancer_b200/csrc/fe_kernel.cuh) into counters_device
 * [fe_profile_slots()] (int64).  NULL switches it off. */
int fe_profile_slots(void);
int fe_set_profile(fe_engine* e, long long* counters_device);

#ifdef __cplusplus
}
#endif
#endif /* FASTENHANCER_B200_H */
