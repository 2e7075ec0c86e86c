// fused-kernel instantiations for the 48B configuration (one translation unit per config so they build in parallel)
#include "fe_inst.cuh"
FE_DEFINE_VARIANTS(variants_48b, FE_VARIANTS_48B)
