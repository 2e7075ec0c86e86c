// fused-kernel instantiations for the 16M configuration (one translation unit per config so they build in parallel)
#include "fe_inst.cuh"
FE_DEFINE_VARIANTS(variants_16m, FE_VARIANTS_16M)
