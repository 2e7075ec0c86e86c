// fe_configs.h -- the shipped FastEnhancer configurations and the (config, streams-per-CTA)
// variants the engine instantiates.  Shapes: configs/fastenhancer/{t,b,s,m,l}.yaml:1-29 and
// configs/fastenhancer_48khz/{t,b,s,m,l}.yaml:1-29 of the reference.
#pragma once
#include "fe_plan.h"

namespace fe {
//                 N_FFT HOP  C1  E  C2  F2  K
using C16T = Cfg<  512, 256,  24, 2, 20, 16, 2>;
using C16B = Cfg<  512, 256,  48, 2, 36, 24, 3>;
using C16S = Cfg<  512, 256,  64, 3, 48, 36, 3>;
using C16M = Cfg<  512, 160,  96, 3, 72, 48, 4>;
using C16L = Cfg<  512, 100, 128, 4, 96, 64, 5>;
using C48T = Cfg< 1024, 512,  24, 2, 20, 24, 2>;
using C48B = Cfg< 1024, 512,  48, 2, 36, 36, 3>;
using C48S = Cfg< 1024, 512,  64, 3, 48, 48, 3>;
using C48M = Cfg< 1024, 320,  96, 3, 72, 72, 4>;
using C48L = Cfg< 1024, 200, 128, 4, 96, 96, 5>;

// Per-(config, S) tuning overrides (measured on B200, profiles/r01/alt_ring_{tf32,f16}.txt).
// T, 1 or 2 streams per CTA: its layers are short, so the weight producer runs further ahead with a third ring stage
// (2 streams: 21.9 -> 20.8 us/hop f16, 23.1 -> 21.7 tf32; 1 stream: 19.0 -> 18.6 f16; 4 streams: 209.4 -> 210.5, kept at two);
// B loses 3 % with the same change and keeps two.
#if FE_STAGES == 2
template <> struct Tune<C16T, 1> : TuneBase<C16T, 1> { static constexpr int STAGES = 3; };
template <> struct Tune<C16T, 2> : TuneBase<C16T, 2> { static constexpr int STAGES = 3; };
#endif
// Two streams per CTA for the larger configs (M in f16 / bf16, S in tf32 / bf16, 48 kHz B / S): the chain per hop is mostly fixed latency, so
// the second stream costs ~40 % (M, 512 streams: 313 -> 215 us / hop, 1.64 -> 2.38 M frames/s) even though the skip tensors of M then
// spill to the L2-resident scratch and the front / back end overlap no longer fits.
// fp32 FMA family of the wide configs (M / L, one stream per CTA): register tiles of the RNNFormer layers sized so that every layer is one
// pass with all eight warps busy -- rnn_fc / attn_fc / qkv with up to 12 channels per lane (8 left L's 96-channel linears with a
// half-empty second pass), the fused GRU tile 4 positions x 4 (M: 5) channels per lane instead of 2 x 6 (13 instead of 6.5
// multiply-adds per shared-memory load).  The tensor-core families of the same (config, S) do not use these fields.
template <> struct Tune<C16M, 1> : TuneBase<C16M, 1> { static constexpr int CT_RF = 12, PT_GRU = 4, CT_GRU = 5; };
template <> struct Tune<C48M, 1> : TuneBase<C48M, 1> { static constexpr int CT_RF = 12, PT_GRU = 4, CT_GRU = 5; };
#ifndef FE_L_CT_GRU
#define FE_L_CT_GRU 4
#endif
template <> struct Tune<C16L, 1> : TuneBase<C16L, 1> { static constexpr int CT_RF = 12, PT_GRU = 4, CT_GRU = FE_L_CT_GRU; };
template <> struct Tune<C48L, 1> : TuneBase<C48L, 1> { static constexpr int CT_RF = 12, PT_GRU = 4, CT_GRU = FE_L_CT_GRU; };
// B, 4 streams per CTA (throughput variant for thousands of streams: the RNNFormer tiles carry 96 of 128 rows instead of 48, the conv
// section runs two M tiles per layer): the activations of four streams leave room for a 2 x 16 KB weight ring only.
template <> struct Tune<C16B, 4> : TuneBase<C16B, 4> { static constexpr int CHUNK = 4096; };
}  // namespace fe

// (B with 4 streams per CTA exists only with the tensor-core frequency-axis linears: the FMA-pipe form does not tile 4 streams)
#ifndef FE_LIN_TC
#define FE_LIN_TC 1
#endif
#if FE_LIN_TC
#define FE_IF_LIN_TC(x) x
#else
#define FE_IF_LIN_TC(x)
#endif
// X(config id, Cfg type, S, PREC)   PREC: false / 0 = everything on the fp32 FMA pipe; true / 1 = contractions on tcgen05 (TF32);
// 2 = as 1 with the conv section's operands in fp16 (K = 16 per MMA, half the shared memory); 3 = bfloat16 conv section, TF32 RNNFormer
// (BASELINE config 3: "bf16 conv / fp32 GRU"); 4 = fp32-accurate split-fp16 operands, three MMAs per product (fe_plan.h)
#define FE_VARIANTS_16T(X) X(0, C16T, 1, false) X(0, C16T, 2, false) X(0, C16T, 4, false) X(0, C16T, 1, true) X(0, C16T, 2, true) X(0, C16T, 4, true) \
    X(0, C16T, 1, 2) X(0, C16T, 2, 2) X(0, C16T, 4, 2) X(0, C16T, 2, 3) X(0, C16T, 1, 4) X(0, C16T, 2, 4)
#define FE_VARIANTS_16B(X) X(1, C16B, 1, false) X(1, C16B, 2, false) X(1, C16B, 1, true) X(1, C16B, 2, true) X(1, C16B, 1, 2) X(1, C16B, 2, 2) \
    X(1, C16B, 2, 3) X(1, C16B, 1, 4) X(1, C16B, 2, 4) FE_IF_LIN_TC(X(1, C16B, 4, 2))
#define FE_VARIANTS_16S(X) X(2, C16S, 1, false) X(2, C16S, 1, true) X(2, C16S, 2, true) X(2, C16S, 1, 2) X(2, C16S, 2, 2) X(2, C16S, 1, 3) X(2, C16S, 2, 3)
#define FE_VARIANTS_16M(X) X(3, C16M, 1, false) X(3, C16M, 1, true) X(3, C16M, 1, 2) X(3, C16M, 1, 3) X(3, C16M, 2, 2) X(3, C16M, 2, 3)
#define FE_VARIANTS_16L(X) X(4, C16L, 1, false) X(4, C16L, 1, true) X(4, C16L, 1, 2) X(4, C16L, 1, 3)
#define FE_VARIANTS_48T(X) X(5, C48T, 1, false) X(5, C48T, 2, false) X(5, C48T, 1, true) X(5, C48T, 2, true) X(5, C48T, 1, 2) X(5, C48T, 2, 2) X(5, C48T, 1, 4)
#define FE_VARIANTS_48B(X) X(6, C48B, 1, false) X(6, C48B, 1, true) X(6, C48B, 2, true) X(6, C48B, 1, 2) X(6, C48B, 2, 2) X(6, C48B, 1, 4) X(6, C48B, 2, 4)
#define FE_VARIANTS_48S(X) X(7, C48S, 1, false) X(7, C48S, 1, true) X(7, C48S, 1, 2) X(7, C48S, 2, 2)
#define FE_VARIANTS_48M(X) X(8, C48M, 1, false) X(8, C48M, 1, true) X(8, C48M, 1, 2) X(8, C48M, 1, 3)
#define FE_VARIANTS_48L(X) X(9, C48L, 1, false) X(9, C48L, 1, true) X(9, C48L, 1, 2) X(9, C48L, 1, 3)
#define FE_ALL_VARIANTS(X) FE_VARIANTS_16T(X) FE_VARIANTS_16B(X) FE_VARIANTS_16S(X) FE_VARIANTS_16M(X) FE_VARIANTS_16L(X) \
    FE_VARIANTS_48T(X) FE_VARIANTS_48B(X) FE_VARIANTS_48S(X) FE_VARIANTS_48M(X) FE_VARIANTS_48L(X)

namespace fe {
struct ShapeKey { int n_fft, hop, c1, n_enc, c2, f2, n_blocks, n_heads; };
template <class C> constexpr bool shape_matches(const ShapeKey& k) {
    return k.n_fft == C::N_FFT && k.hop == C::HOP && k.c1 == C::C1 && k.n_enc == C::E && k.c2 == C::C2 && k.f2 == C::F2 &&
           k.n_blocks == C::K && k.n_heads == C::NH;
}
}  // namespace fe
