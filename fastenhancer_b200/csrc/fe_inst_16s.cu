// fused-kernel instantiations for the 16S configuration (one translation unit per config so they build in parallel)
#include "fe_inst.cuh"
FE_DEFINE_VARIANTS(variants_16s, FE_VARIANTS_16S)
