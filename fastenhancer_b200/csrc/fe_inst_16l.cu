// fused-kernel instantiations for the 16L configuration (one translation unit per config so they build in parallel)
#include "fe_inst.cuh"
FE_DEFINE_VARIANTS(variants_16l, FE_VARIANTS_16L)
