// fe_stft_gemm.h -- host interface of the tensor-core (DFT-as-GEMM) STFT kernel, fe_stft_gemm.cu.
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace fe {
// windowed real-DFT basis [N][N] (row = packed output column, column = sample in frame), split hi + lo (both TF32-exact) for 3xTF32
void stft_gemm_basis(int n_fft, const float* window, std::vector<float>& hi, std::vector<float>& lo);
// spec_out [B][n_fft/2+1][T][2] = STFT of wav [B][ld] (frame t = samples t*hop .. t*hop + n_fft - 1; device pointers).
// 0 = ok, 1 = bad alignment / pitch, 2 = tensor-map encoding failed, 3 = CUDA error (*cuda_err)
int stft_gemm_launch(const float* wav, long long ld, int B, int T, int n_fft, int hop, const float* basis_hi, const float* basis_lo, float* spec_out,
                     int accurate, cudaStream_t stream, cudaError_t* cuda_err);
int stft_gemm_smem_bytes();
}  // namespace fe
