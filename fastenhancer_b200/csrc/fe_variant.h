// fe_variant.h -- type-erased handle on one instantiated (config, streams-per-CTA) kernel variant.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <vector>

#include "fe_configs.h"

namespace fe {

struct VariantOps {
    int cfg_id, S, tc;
    ShapeKey shape;
    int smem_bytes, nthreads, gs_floats, state_floats, tap_floats, nchunk_frame;
    long blob_floats;
    int tp_group;                                                  // floats of per-group scratch of the frame-parallel offline schedule (fp32 family)
    int aux_window_sq;                                             // float offset of window^2 [N] in the blob
    int hop_ring, hop_tile;                                        // hop tiles by 2-D TMA (cp.async.bulk.tensor): supported, tile width
    void (*pack)(const float* canonical, std::vector<float>& blob);
    cudaError_t (*prepare)();                                      // one-time function attributes
    cudaError_t (*launch)(const KParams& prm, int grid, cudaStream_t stream);
};

// one per translation unit (fe_inst_*.cu)
const VariantOps* variants_16t(int* n);
const VariantOps* variants_16b(int* n);
const VariantOps* variants_16s(int* n);
const VariantOps* variants_16m(int* n);
const VariantOps* variants_16l(int* n);
const VariantOps* variants_48t(int* n);
const VariantOps* variants_48b(int* n);
const VariantOps* variants_48s(int* n);
const VariantOps* variants_48m(int* n);
const VariantOps* variants_48l(int* n);

}  // namespace fe
