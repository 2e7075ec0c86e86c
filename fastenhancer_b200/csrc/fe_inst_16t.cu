// fused-kernel instantiations for the 16T configuration (one translation unit per config so they build in parallel)
#include "fe_inst.cuh"
FE_DEFINE_VARIANTS(variants_16t, FE_VARIANTS_16T)
