// fused-kernel instantiations for the 48L configuration (one translation unit per config so they build in parallel)
#include "fe_inst.cuh"
FE_DEFINE_VARIANTS(variants_48l, FE_VARIANTS_48L)
