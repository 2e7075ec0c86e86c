// fe_frontdoor.cu -- the audio "front door" of the directory-level runner on the device: what librosa.load(path, sr=wrapper.sr,
// mono=True) and soundfile.write do around the model in the reference's scripts/test_pytorch.py:29,37 --
//   int16 PCM (interleaved channels) -> mono float32 in [-1, 1)          fe_pcm16_to_float
//   rational-ratio polyphase resampling (e.g. 48 kHz -> 16 kHz)           fe_resample_poly
//   float32 -> int16 PCM (soundfile's default subtype for .wav)           fe_float_to_pcm16
// so that a WAV file goes host bytes -> GPU -> enhanced bytes without a host-side sample loop.
// Resampling semantics = scipy.signal.resample_poly (zero-phase FIR, Kaiser beta 5 by default -- the oracle in tests/test_frontdoor.py;
// the taps are designed on the host once per ratio, fastenhancer_b200/frontdoor.py).  librosa itself resamples with soxr, which is
// not in this image: files at the model's rate (the reference's own fixtures) are bit-exact, resampled ones follow the polyphase oracle.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/fastenhancer_b200.h"

namespace {
// y[m] = sum_j x[j] * h[m * down - j * up + center],  the "upfirdn" form: only taps congruent to m * down (mod up) contribute
__global__ void resample_poly_kernel(const float* __restrict__ x, long n_in, const float* __restrict__ h, int n_taps, int up, int down,
                                     int center, float* __restrict__ y, long n_out)
{
    for (long m = blockIdx.x * (long)blockDim.x + threadIdx.x; m < n_out; m += (long)gridDim.x * blockDim.x) {
        const long t = m * down + center;                 // index into the up-sampled, filtered signal
        long j_hi = t / up;                               // largest j with t - j * up >= 0
        if (j_hi > n_in - 1) j_hi = n_in - 1;
        long j_lo = (t - (n_taps - 1) + up - 1) / up;     // smallest j with t - j * up <= n_taps - 1
        if (t - (n_taps - 1) < 0) j_lo = 0;
        float acc = 0.f;
        for (long j = j_lo; j <= j_hi; ++j) acc = fmaf(x[j], h[t - j * up], acc);
        y[m] = acc;
    }
}
__global__ void pcm16_to_float_kernel(const int16_t* __restrict__ pcm, long n_frames, int n_channels, float* __restrict__ out)
{
    const float inv = 1.0f / (32768.0f * (float)n_channels);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n_frames; i += (long)gridDim.x * blockDim.x) {
        int s = 0;
        for (int c = 0; c < n_channels; ++c) s += pcm[i * n_channels + c];
        out[i] = (float)s * inv;                           // mono mix-down = mean of the channels, exact in fp32 for <= 256 channels
    }
}
__global__ void float_to_pcm16_kernel(const float* __restrict__ x, long n, int16_t* __restrict__ pcm)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float v = x[i] * 32768.0f;
        v = fminf(fmaxf(v, -32768.0f), 32767.0f);
        pcm[i] = (int16_t)__float2int_rn(v);
    }
}
int grid_for(long n) { long g = (n + 255) / 256; return (int)(g < 1 ? 1 : (g > 4096 ? 4096 : g)); }
}  // namespace

#define FE_API __attribute__((visibility("default")))
extern "C" {

FE_API int fe_pcm16_to_float(const short* pcm_device, long long n_frames, int n_channels, float* wav_device, void* cuda_stream) {
    if (!pcm_device || !wav_device || n_frames < 0 || n_channels < 1 || n_channels > 256) return FE_ERR_ARG;
    if (n_frames == 0) return FE_OK;
    pcm16_to_float_kernel<<<grid_for(n_frames), 256, 0, (cudaStream_t)cuda_stream>>>((const int16_t*)pcm_device, n_frames, n_channels, wav_device);
    return cudaGetLastError() == cudaSuccess ? FE_OK : FE_ERR_CUDA;
}

FE_API int fe_float_to_pcm16(const float* wav_device, long long n, short* pcm_device, void* cuda_stream) {
    if (!pcm_device || !wav_device || n < 0) return FE_ERR_ARG;
    if (n == 0) return FE_OK;
    float_to_pcm16_kernel<<<grid_for(n), 256, 0, (cudaStream_t)cuda_stream>>>(wav_device, n, (int16_t*)pcm_device);
    return cudaGetLastError() == cudaSuccess ? FE_OK : FE_ERR_CUDA;
}

FE_API int fe_resample_poly(const float* in_device, long long n_in, int up, int down, const float* taps_device, int n_taps,
                            float* out_device, long long n_out, void* cuda_stream) {
    if (!in_device || !out_device || !taps_device || n_in <= 0 || n_out <= 0 || up < 1 || down < 1 || n_taps < 1 || (n_taps & 1) == 0) return FE_ERR_ARG;
    resample_poly_kernel<<<grid_for(n_out), 256, 0, (cudaStream_t)cuda_stream>>>(in_device, n_in, taps_device, n_taps, up, down, (n_taps - 1) / 2,
                                                                                 out_device, n_out);
    return cudaGetLastError() == cudaSuccess ? FE_OK : FE_ERR_CUDA;
}

}  // extern "C"
