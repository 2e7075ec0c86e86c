// fused-kernel instantiations for the 48M configuration (one translation unit per config so they build in parallel)
#include "fe_inst.cuh"
FE_DEFINE_VARIANTS(variants_48m, FE_VARIANTS_48M)
