// fe_plan.h -- compile-time shape, tile, shared-memory and weight-stream plan of the fused per-hop kernel.
//
// One `Cfg` per shipped FastEnhancer configuration (reference YAMLs:
// configs/fastenhancer/{t,b,s,m,l}.yaml:1-29, configs/fastenhancer_48khz/*.yaml:1-29) and one
// `Plan<Cfg, S>` per (configuration, streams-per-CTA).  The device kernel (fe_kernel.cuh), the
// host-side weight packer (fe_pack.h) and the CPU emulation build used by the tests
// (tests/emu/fe_emu.cpp) are all instantiated from the same Plan, so the packed layout has a single
// source of truth.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define FE_HD __host__ __device__
#else
#define FE_HD
#endif

namespace fe {

constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }
constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }
constexpr int pow2ceil(int a) { int p = 1; while (p < a) p <<= 1; return p; }

// ----------------------------------------------------------------------------------------------
// Model shape (SURVEY.md section 8 notation).  kernel_size = [8, 3 x E], stride 4, 4 heads, SiLU.
// ----------------------------------------------------------------------------------------------
template <int N_FFT_, int HOP_, int C1_, int E_, int C2_, int F2_, int K_, int NH_ = 4>
struct Cfg {
    static constexpr int N_FFT = N_FFT_, HOP = HOP_, C1 = C1_, E = E_, C2 = C2_, F2 = F2_, K = K_, NH = NH_;
    static constexpr int FIN = N_FFT / 2;      // bins kept (Nyquist dropped)
    static constexpr int NB = FIN + 1;         // bins of the rFFT
    static constexpr int F1 = FIN / 4;         // encoder frequency positions
    static constexpr int HD = C2 / NH;         // attention head dim
    static constexpr int CL = N_FFT - HOP;     // overlap / cache length
    static constexpr int M = N_FFT / 2;        // complex FFT size of the packed real FFT
    static constexpr int HS = K * F2 * C2;     // floats of GRU state per stream
    static constexpr int STATE = 2 * CL + HS;  // floats of state per stream
    static_assert(C2 % NH == 0, "heads");
    static_assert((N_FFT & (N_FFT - 1)) == 0, "n_fft must be a power of two");
    static_assert(F2 % 4 == 0 && F1 % 4 == 0, "frequency axes must be multiples of 4");
};

// Number of floats of the canonical folded weight array (include/fastenhancer_b200.h).
template <class C> constexpr long canonical_floats() {
    long n = 0;
    n += C::C1 * 16 + C::C1;
    n += (long)C::E * (C::C1 * C::C1 * 3 + C::C1);
    n += C::F2 * C::F1 + C::C2 * C::C1 + C::C2;
    n += (long)C::K * (2 * 3 * C::C2 * C::C2 + 2 * 3 * C::C2 + C::C2 * C::C2 + C::C2 + 3 * C::C2 * C::C2 + 3 * C::C2 + C::C2 * C::C2 + C::C2);
    n += C::F2 * C::C2;
    n += C::F1 * C::F2 + C::C1 * C::C2 + C::C1;
    n += (long)C::E * (C::C1 * 2 * C::C1 + C::C1 + C::C1 * C::C1 * 3 + C::C1);
    n += C::C1 * 2 * C::C1 + C::C1 + C::C1 * 16 + 2;
    return n;
}

// ----------------------------------------------------------------------------------------------
// "Position GEMM": Y[co][p] = sum_{k, tap} W[co][k][tap] * X[k][p + tap - TAPS/2], activations
// channel-major in shared memory with the positions contiguous.  A warp owns a tile of
// (PL*PT positions) x (CL*CT output channels): lane = cl*PL + pl, each lane PT consecutive
// positions x CT channels in registers.  Weights arrive through the shared-memory ring as rows
//     row k of pass p : [cgp][cl][RW],  RW = round_up(SETS*TAPS*CT, 4),  entry [set][tap][i]
// so a lane reads RW contiguous floats per k (float4 loads, broadcast across the pl lanes).
// SETS = 6 is the fused GRU tile (W_ir, W_iz, W_in, W_hr, W_hz, W_hn).
// If the layer has more tiles than warps it runs in NPASS passes; the stream is pass-major.
// ----------------------------------------------------------------------------------------------
template <int NPOS_, int F_, int COUT_, int K_, int TAPS_, int PT_, int CTMAX_, int NW_, int CHUNK_, int SETS_ = 1>
struct PosGemm {
    static constexpr int NPOS = NPOS_, F = F_, COUT = COUT_, K = K_, TAPS = TAPS_, PT = PT_, NW = NW_, SETS = SETS_;
    static_assert(F % PT == 0, "a thread's positions must not straddle streams");
    static constexpr int PL = cmin(32, pow2ceil(cdiv(NPOS, PT)));
    static constexpr int CL = 32 / PL;
    static constexpr int NPG = pow2ceil(cdiv(NPOS, PL * PT));
    static_assert(NPG <= NW, "too many positions for one pass: lower S");
    static constexpr int NCGP = NW / NPG;                       // channel groups per pass
    static constexpr int CT = cmax(1, cmin(CTMAX_, cdiv(COUT, CL * NCGP)));
    static constexpr int NCG = cdiv(COUT, CL * CT);
    static constexpr int NPASS = cdiv(NCG, NCGP);
    static constexpr int RW = round_up(SETS * TAPS * CT, 4);
    static constexpr int ROW = NCGP * CL * RW;                  // floats per k row of one pass
    static_assert(ROW <= CHUNK_, "one weight row must fit a ring chunk");
    static constexpr int KC = cmax(1, cmin(K, CHUNK_ / ROW));   // rows per chunk
    static constexpr int NCHUNK_PASS = cdiv(K, KC);
    static constexpr int NCHUNK = NPASS * NCHUNK_PASS;
    static constexpr int FLOATS = NPASS * K * ROW;
};

// ----------------------------------------------------------------------------------------------
// "Row GEMM": Y[r][o] = sum_k W[o][k] * X[r][k] with k the contiguous axis of X (the two
// frequency-axis linears rf_pre.0 / rf_post.0).  Lanes own rows (row pitch = 4*odd floats, so the
// float4 loads of 8 consecutive lanes hit 8 distinct bank groups), a warp owns NO outputs.
// Ring rows: one per 4 k's: [og][NO][4].
// ----------------------------------------------------------------------------------------------
template <int NROWS_, int K_, int NOUT_, int NW_, int CHUNK_>
struct RowGemm {
    static constexpr int NROWS = NROWS_, K = K_, NOUT = NOUT_, NW = NW_;
    static_assert(K % 4 == 0, "contraction axis must be a multiple of 4");
    static constexpr int RT = cdiv(NROWS, 32);
    static constexpr int NO = round_up(cdiv(NOUT, NW), 4);
    static constexpr int NOG = cdiv(NOUT, NO);
    static_assert(NOG <= NW && RT * NO <= 64, "row gemm tile too large: lower S");
    static constexpr int K4 = K / 4;
    static constexpr int ROW = NOG * NO * 4;
    static_assert(ROW <= CHUNK_, "one weight row must fit a ring chunk");
    static constexpr int KC = cmax(1, cmin(K4, CHUNK_ / ROW));
    static constexpr int NCHUNK = cdiv(K4, KC);
    static constexpr int FLOATS = K4 * ROW;
};

// ----------------------------------------------------------------------------------------------
// "Tensor-core GEMM" (tcgen05.mma kind::tf32, accumulators in TMEM): D[pos][n] = sum_{k,tap} X[k][pos+tap-1] W[n][k][tap].
// Activations live in shared memory as K-major, un-swizzled UMMA operands:  X[k/4][slot][k%4]  with
// slot = (f + 1) * S + s  (streams interleaved, S zero slots before and after the data), so rows are 16 B
// apart and a shift by one frequency position is a start-address offset of S*16 bytes: the three taps
// of a k=3 conv are three MMAs on the same buffer.  M = 128 rows per MMA (NMT tiles of 128 positions).
// Weights stream through the ring as tiles of one (tap, 8-wide k-step):  [2][NP][4]  (K-major B operand),
// NP = N rounded up to 16, zero rows / zero k beyond the real sizes, values pre-rounded to TF32.
// ----------------------------------------------------------------------------------------------
// KE = 16: kind::f16 -- the operands are halves, 8 per 16-byte row, so a tile [2][NP][8 halves] has the same bytes and covers K = 16.
// PARTS = 2 (split variants): every tile is followed by the tile of the weights' low parts (w - fp16(w) as fp16), and the layer
// issues three MMAs per tile -- (x_hi, w_hi), (x_lo, w_hi), (x_hi, w_lo) -- for an fp32-accurate product.
template <int NPOS_, int N_, int K_, int TAPS_, int CHUNK_, int TMEMC_ = 256, int KE_ = 8, int PARTS_ = 1>
struct TcGemm {
    static constexpr int PARTS = PARTS_;
    static constexpr int NPOS = NPOS_, N = N_, K = K_, TAPS = TAPS_;   // TAPS doubles as "weight sets" for the GRU (6)
    static constexpr int KE = KE_;                     // contraction length of one MMA
    static constexpr int NP = round_up(N, 16), KP = round_up(K, KE);
    static constexpr int NKS = KP / KE;                // k-steps per tap
    static constexpr int NTILE = TAPS * NKS;
    static constexpr int TILE1 = NP * 8;               // floats per tile part (32 bytes per output row)
    static constexpr int WLBO = NP * 4;                // floats between the two k-chunks of a weight tile
    static constexpr int TILE = TILE1 * PARTS;
    static_assert(TILE <= CHUNK_, "one weight tile must fit a ring chunk");
    static constexpr int TPC = cmax(1, cmin(NTILE, CHUNK_ / TILE));
    static constexpr int NCHUNK = cdiv(NTILE, TPC);
    static constexpr int FLOATS = NTILE * TILE;
    static constexpr int NMT = cdiv(NPOS, 128);        // 128-row M tiles
    static constexpr int NG = cdiv(N, 4);              // output channel groups of 4
    static constexpr int NSPLIT = cdiv(NP, 256);       // one tcgen05.mma covers at most 256 accumulator columns
    static constexpr int NPS = NP / NSPLIT;
    static_assert(NPS * NSPLIT == NP && NPS % 16 == 0 && NPS <= 256, "N split");
    static_assert(NMT * NP <= TMEMC_, "accumulators exceed the TMEM allocation");
};

// Fused GRU on the tensor cores.
// MERGED (3 * NPG <= 256: T, B, S, M): the accumulator columns are [ NX | R | Z | NH ] and every (input, k-step) is ONE MMA with N = 3 NPG
// on a tile [2][4 NPG][4] whose rows are [ W_in | W_r | W_z | W_hn ] (x tiles: W_hn rows zero, h tiles: W_in rows zero): the h tiles
// come first -- k-step 0 with N = 4 NPG and accumulate = 0, which also zeroes NX -- then accumulate onto columns NPG .. 4 NPG - 1
// (rows NPG ..), and the x tiles accumulate onto columns 0 .. 3 NPG - 1 (rows 0 ..).  Half the MMAs of the split form.
// Split form (L): per (input x | h, k-step) one tile [ R|Z part: [2][2 NPG][4] | N part: [2][NPG][4] ], i.e. two MMAs: N = 2 NPG into
// the R|Z accumulator columns (x and h accumulate together) and N = NPG into NX or NH; x tiles first; columns [ R | Z | NX | NH ].
// KE = 16: fp16 operands (x and h as packed halves), tiles [2][rows][8 halves].
template <int NPOS_, int C2_, int CHUNK_, int KE_ = 8, int PARTS_ = 1>
struct TcGru {
    static constexpr int NPOS = NPOS_, N = C2_, K = C2_, KE = KE_, PARTS = PARTS_;
    static constexpr int NPG = round_up(C2_, 16), KP = round_up(C2_, KE_), NKS = KP / KE_;
#ifndef FE_GRU_MERGE
#define FE_GRU_MERGE 1
#endif
    static constexpr bool MERGED = FE_GRU_MERGE && 3 * NPG <= 256;
    // WIDE (M: 3 NPG <= 256 < 4 NPG): no single MMA covers all four accumulator blocks, so the first h tile writes [ R | Z | NH ]
    // (N = 3 NPG, accumulate = 0) and the first x tile is two MMAs -- NX alone with accumulate = 0, then [ R | Z ] accumulating --;
    // every other tile is one N = 3 NPG MMA as in the narrow form: 2 NKS + 1 MMAs per block instead of the split form's 4 NKS.
    static constexpr bool WIDE = MERGED && 4 * NPG > 256;
    static constexpr int NP = MERGED ? 4 * NPG : 2 * NPG;              // rows (LBO) of the tile / of its R|Z part
    static constexpr int NTILE = 2 * NKS;              // MERGED: h tiles, then x tiles; split: x tiles, then h tiles
    static constexpr int TILE1 = (MERGED ? 4 : 3) * NPG * 8;
    static constexpr int WLBO = NP * 4;
    static constexpr int TILE = TILE1 * PARTS;
    static constexpr int COL_NX = MERGED ? 0 : 2 * NPG, COL_R = MERGED ? NPG : 0, COL_Z = MERGED ? 2 * NPG : NPG, COL_NH = 3 * NPG;
    static_assert(TILE <= CHUNK_ && (MERGED ? 3 * NPG : NP) <= 256, "GRU tile");
    static constexpr int TPC = cmax(1, cmin(NTILE, CHUNK_ / TILE));
    static constexpr int NCHUNK = cdiv(NTILE, TPC);
    static constexpr int FLOATS = NTILE * TILE;
};

// Frequency-axis linear on the tensor cores (16-bit variants): D[row][c] = sum_slot W[row][slot] X[slot][c], a contraction over the SLOTS
// of an activation buffer [c / 8][slot][8 halves].  That buffer is the B operand, MN-major (its N index, the channel, is the contiguous
// one: tools/tc_probe_mn.cu validated the descriptor on hardware -- LBO = 128 B between groups of 8 slots, SBO = the slab pitch).  The
// weights are the A operand, K-major, 128 rows per M tile: row = output slot (frequency-major, streams interleaved), k = input slot,
// W[(f_out, s)][(f_in, s')] = (s == s') * w[f_out][f_in] (streams share a tile, so the other streams' columns are zero).
// Ring tiles, one per (M tile, k-step of 16 slots): [2][128][8 halves] (+ the tile of the low parts in the split variants).
template <int ROWS_, int N_, int K_, int CHUNK_, int PARTS_ = 1>
struct TcLin {
    static constexpr int NPOS = ROWS_, N = N_, K = K_, KE = 16, PARTS = PARTS_, TAPS = 1;
    static_assert(K % 16 == 0 && N % 16 == 0, "slot contraction: whole k-steps, N a multiple of 16");
    static constexpr int NP = N, NKS = K / 16, NMT = cdiv(ROWS_, 128);
    static constexpr int NTILE = NMT * NKS;             // M-tile major
    static constexpr int TILE1 = 128 * 8, TILE = TILE1 * PARTS, WLBO = 128 * 4;
    static_assert(TILE <= CHUNK_, "one weight tile must fit a ring chunk");
    static constexpr int TPC = cmax(1, cmin(NTILE, CHUNK_ / TILE));
    static constexpr int NCHUNK = cdiv(NTILE, TPC);
    static constexpr int FLOATS = NTILE * TILE;
    static constexpr int NG = N / 4, NSPLIT = 1, NPS = N;
    static_assert(N <= 256, "N");
};

// Row GEMM reading one k per step (frequency-axis linear on a tensor-core-layout activation, where
// consecutive frequencies are not contiguous).  Ring rows: one per k: [og][NO].
template <int NROWS_, int K_, int NOUT_, int NW_, int CHUNK_>
struct RowGemmK1 {
    static constexpr int NROWS = NROWS_, K = K_, NOUT = NOUT_, NW = NW_;
    static constexpr int RT = cdiv(NROWS, 32);
    static constexpr int NO = round_up(cdiv(NOUT, NW), 4);
    static constexpr int NOG = cdiv(NOUT, NO);
    static_assert(NOG <= NW && RT * NO <= 64, "row gemm tile too large: lower S");
    static constexpr int ROW = NOG * NO;
#ifndef FE_LIN_KC_MAX
#define FE_LIN_KC_MAX 1000000      // experiment switch: cap the rows per ring chunk (more, smaller chunks)
#endif
    static constexpr int KC = cmax(1, cmin(cmin(K, CHUNK_ / ROW), FE_LIN_KC_MAX));
    static constexpr int NCHUNK = cdiv(K, KC);
    static constexpr int FLOATS = K * ROW;
};

// ----------------------------------------------------------------------------------------------
// Per-(config, S) tuning.  Primary template = generic heuristics; specialise to override.
// ----------------------------------------------------------------------------------------------
template <class C, int S> struct TuneBase {
    static constexpr int NW = 8;                      // consumer warps (one more warp streams weights)
#ifndef FE_CHUNK
#define FE_CHUNK 8192
#endif
#ifndef FE_STAGES
#define FE_STAGES 2
#endif
    static constexpr int CHUNK = FE_CHUNK;            // floats per ring chunk
    static constexpr int STAGES = FE_STAGES;
    static constexpr int CT_CONV = 16;                // max output channels per lane, conv / 1x1 layers
    static constexpr int CT_RF = 8;                   // ... RNNFormer linears
    static constexpr int PT_GRU = 2, CT_GRU = 6;      // GRU tile: 4 accumulators per (channel, position)
    // heads per attention round: the largest that fits the work region (see Plan static_asserts)
    static constexpr int HG = (C::C1 >= 96) ? 1 : C::NH;
    // upper bound on skip tensors kept in shared memory (the rest round-trip through L2);
    // the Plan lowers it to what fits in 227 KB
    static constexpr int SKIP_SMEM_MAX = C::E + 1;
};
template <class C, int S> struct Tune : TuneBase<C, S> {};

// PREC: 0 = everything on the fp32 FMA pipe; 1 = contractions on tcgen05 with TF32 operands; 2 = as 1, with the conv section's
// activations and weights stored as fp16 (kind::f16: 11-bit significand like TF32, K = 16 per MMA, half the shared memory);
// 3 = as 2 with bfloat16 instead of fp16 in the conv section and a TF32 RNNFormer ("bf16 conv / fp32 GRU", BASELINE config 3);
// 4 = fp32-accurate tensor-core mode: every MMA operand is stored as two fp16 parts, hi = fp16(v) and lo = fp16(v - hi)
//     (22 significand bits together), and every product is three kind::f16 MMAs into the same fp32 accumulator:
//     hi*hi + lo*hi + hi*lo (the dropped lo*lo term is ~2^-22 relative).  Same shared-memory footprint as the TF32 variant.
template <class C, int S_, int PREC_ = 0>
struct Plan {
    using Cf = C;
    static constexpr int S = S_;
    static constexpr int PREC = PREC_;
    static constexpr bool TC = PREC_ != 0;             // conv-type contractions on tcgen05 instead of the FMA pipe
    static constexpr bool H16 = PREC_ >= 2;            // conv-section operands are 16-bit
    static constexpr bool BF16 = PREC_ == 3;           // ... bfloat16 instead of fp16
    static constexpr bool SPLIT = PREC_ == 4;          // ... stored as hi + lo fp16 parts
    static constexpr int NPART = SPLIT ? 2 : 1;
    // tanh.approx-based SiLU / GRU gates (error ~2^-11, the size of the operand rounding that follows) in the reduced-precision tensor-core
    // variants; the fp32-accurate split variants use the ex2 / rcp forms (~1e-7) and un-halved SiLU layers
    static constexpr bool FAST_ACT = TC && !SPLIT;
    static constexpr int CG = H16 ? 8 : 4;             // channels per 16-byte row of a conv-section operand buffer
    static constexpr int KEC = H16 ? 16 : 8;           // contraction length of one conv-section MMA
    static constexpr int C1P = H16 ? round_up(C::C1, 16) : C::C1;                 // conv channels padded to a k-step
    static constexpr int C2Z = H16 ? round_up(C::C2, 16) : round_up(C::C2, 8);    // rf_post conv input channels padded to a k-step
    using T = Tune<C, S_>;
    static constexpr int NW = T::NW, NT = NW * 32, NTHREADS = NT + 32;
    static constexpr int CHUNK = T::CHUNK, STAGES = T::STAGES;
    // ---- conv geometry ("Geo1"): [C1][S][P1], data at column 4, zero columns 0..3; the 4 zero
    //      columns of the next row are the right halo, +4 floats after the very last row ----
    static constexpr int P1 = C::F1 + 4;
    static constexpr int CP1 = S * P1;                 // channel pitch
    // ---- tensor-core geometry ("GeoT"): [C/4][SLOTS][4], slot = (f+1)*S + s ----
    static constexpr int SLOTS = (C::F1 + 2) * S;
    static constexpr int SLABF = SLOTS * 4;            // floats per 4-channel slab
    static constexpr int C2P = round_up(C::C2, 8);     // rf_post conv input channels padded to a k-step
    static constexpr int ACT1 = (C1P / CG) * SLABF;    // one part of a conv-section operand buffer (TC variants); the low parts follow at + ACT1
    static constexpr int ACT = TC ? NPART * ACT1 : C::C1 * CP1 + 4;
    // compressed spectrum: 8 virtual channels (c*4+q), fp32 in every variant; split variants add an fp16 operand copy for enc_pre:
    // [hi slab | zero slab | lo slab | zero slab] of 8 halves per slot (a zero slab is the second k-chunk of a part: K = 8 padded to 16)
    static constexpr int SPECF = TC ? 2 * SLABF + (SPLIT ? 4 * SLABF : 0) : 8 * CP1 + 4;
    static constexpr int O_SPECH = 2 * SLABF;
    static constexpr int ZB1 = (C2Z / CG) * SLABF;
    static constexpr int ZBF = TC ? NPART * ZB1 : C::C2 * CP1;   // rf_post linear output
    static_assert(C::C1 % 8 == 0, "C1 must be a multiple of 8");
    static_assert(2 * SLABF <= ACT, "the fp32 mask (two 4-channel slabs) must fit an activation buffer");
    // ---- RNNFormer geometry: [C2][S][F2P] channel-major, F2P = 4*odd ----
    static constexpr int F2P = ((C::F2 / 4) % 2 == 1) ? C::F2 : C::F2 + 4;
    static constexpr int PR = S * F2P;
    static constexpr int XRS = C::C2 * PR;
    // heads per attention round.  fp32 variants: Tune::HG.  Tensor-core variants: as many as the shared-memory budget
    // allows (more parallel work for the thread-per-query attention), trading resident skip tensors for scratch.
    static constexpr int hg_tc() {
        for (int hg = C::NH; hg > 1; hg /= 2) {
            if (C::NH % hg) continue;
            const int act = NPART * (C1P / CG) * (C::F1 + 2) * S * 4, xts = (round_up(C::C2, 8) / 4) * S * C::F2 * 4;
            const int f2p = ((C::F2 / 4) % 2 == 1) ? C::F2 : C::F2 + 4;
            const int qn = hg * 3 * round_up(C::HD, 4), qrow = ((qn / 4) % 2 == 1) ? qn : qn + 4;
            const int need1 = cmax(S * C::F2 * qrow + xts, NPART * (C1P / CG) * S * C::F2 * 4) + xts;
            if (f2p < 0) return 0;
            const int rest = 2 * (C::F1 + 2) * S * 4 + 2 * S * C::N_FFT + T::STAGES * T::CHUNK + 4 * T::STAGES + 4;
            if (cmax(2, cdiv(need1, act)) * act + rest <= 227 * 256) return hg;
        }
        return 1;
    }
    static constexpr int HG = TC ? hg_tc() : T::HG, NQG = C::NH / HG;
    static_assert(C::NH % HG == 0, "HG must divide NH");
    // fp32 variants: q/k/v rows channel-major [3*HD*HG][PR].  Tensor-core variants: position-major [RSLOTS][QROW] with every
    // head's q | k | v padded to HDP = round_up(HD, 4) (the padding rows of the packed weights are zero), so the MMA epilogue
    // stores and the attention loads are float4; QROW = 4*odd floats keeps both conflict-free.
    static constexpr int HDP = round_up(C::HD, 4);
    static constexpr int QN = HG * 3 * HDP;
    static constexpr int QROW_PAD = ((QN / 4) % 2 == 1) ? QN : QN + 4;
    // ... unless that padding alone would cost another work buffer (16 kHz B: 192 floats over)
    static constexpr int rf_need2(int qrow) {      // [QKV | ATT] or Y1T, then XT (padded channels), then XR (real channels only)
        return cmax(S * C::F2 * qrow + (round_up(C::C2, 8) / 4) * S * C::F2 * 4, NPART * (C1P / CG) * S * C::F2 * 4) +
               (round_up(C::C2, 8) / 4) * S * C::F2 * 4 + (C::C2 / 4) * S * C::F2 * 4;
    }
    static constexpr int act_tc() { return NPART * (C1P / CG) * (C::F1 + 2) * S * 4; }
    static constexpr int QROW = (cdiv(rf_need2(QROW_PAD), act_tc()) > cdiv(rf_need2(QN), act_tc())) ? QN : QROW_PAD;
    static constexpr int QKVS = TC ? S * C::F2 * QROW : 3 * C::HD * HG * PR;
    // ---- RNNFormer tensor-core geometry ("GeoR", TC variants): [C2P/4][RSLOTS][4], slot = f2*S + s ----
    static constexpr int RSLOTS = S * C::F2;
    static constexpr int RSLABF = RSLOTS * 4;
    static constexpr int XTS = (C2P / 4) * RSLABF;     // one RNNFormer activation in GeoR (x, h, attention output)
    static constexpr int Y1T1 = (C1P / CG) * RSLABF;
    static constexpr int Y1TS = NPART * Y1T1;          // rf_pre linear output in GeoR (16-bit in the H16 variants; low parts at + Y1T1)
    static constexpr int NPG = round_up(C::C2, 16);    // accumulator columns per GRU gate
    static_assert(RSLOTS <= 128, "RNNFormer positions exceed one M tile: lower S");
    // RNNFormer MMAs with M = 64 when the positions fit: a 64-row accumulator occupies 16 lanes in each of the four TMEM lane
    // quadrants, so (with 16x256b loads, where all 32 threads of a warp share 16 rows) the epilogues spread over three or four
    // schedulers instead of the one or two that own lanes 0..RSLOTS-1 of an M = 128 accumulator.
#ifndef FE_RM64
#define FE_RM64 0      // measured on B200 (B, 256 streams): 49.7 vs 47.2 us/hop -- the 2-channel granularity of the 16x256b mapping doubles
#endif                 // the per-element epilogue overhead, which outweighs the extra warps; kept as an experiment switch
    static constexpr bool RM64 = TC && FE_RM64 && RSLOTS <= 64;
    // work region AB = [W0 | W1 (| W2)]; RNNFormer: XR at the tail, ATT/HB at 0, G/QKV at XRS.
    // W2 exists only when the RNNFormer scratch needs the room (16 kHz L).
    // TC variants: [QKV | Y1T | Zb from 0 ... | ATT (= h scratch when h is not resident) | XT | XR at the tail].
    // XT is a TF32-rounded copy of x for the MMAs; when it does not fit (48 kHz L) the MMAs read the fp32 master
    // (the tensor core then truncates instead of rounding).
    static constexpr int SM_REST = SPECF + 32 + 2 * S * C::N_FFT + T::STAGES * T::CHUNK + 4 * T::STAGES + 8;   // (+32: TMA alignment of the rings)
    static constexpr int XRT = (C::C2 / 4) * RSLABF;   // fp32 master of x when a separate rounded copy exists: no K-padding group
    static constexpr int RF_NEED1 = cmax(QKVS + XTS, Y1TS) + XTS;
    static constexpr int RF_NEED2 = RF_NEED1 + XRT;
    static constexpr int NWORK2 = cmax(2, cdiv(RF_NEED2, ACT));
    static constexpr bool XT_COPY = TC && (NWORK2 * ACT + SM_REST <= 227 * 256);
    static constexpr int NWORK = TC ? (XT_COPY ? NWORK2 : cmax(2, cdiv(RF_NEED1, ACT))) : ((2 * XRS + cmax(XRS, QKVS) > 2 * ACT) ? 3 : 2);
    static constexpr int AB = NWORK * ACT;
    static constexpr int O_XR = AB - (TC ? (XT_COPY ? XRT : XTS) : XRS);
    static constexpr int O_XT = XT_COPY ? O_XR - XTS : O_XR;
    static constexpr int O_ATT_T = O_XT - XTS, O_HB_T = O_ATT_T;           // TC variants
    static constexpr int O_HB = 0, O_ATT = 0, O_G = XRS, O_QKV = TC ? 0 : XRS;
    static constexpr int O_Y1 = 0;                     // rf_pre linear output
    static constexpr int O_Z = 0;                      // rf_post linear output
    static_assert(TC || XRS + cmax(XRS, QKVS) <= O_XR, "RNNFormer scratch does not fit: lower Tune::HG");
    static_assert(!SPLIT || XT_COPY, "split variants keep the lo parts of the attention output in the (otherwise unused) XT region");
    static_assert(!TC || (QKVS <= O_ATT_T && Y1TS <= O_XT && Y1TS <= (NWORK - 1) * ACT), "RNNFormer tensor-core scratch does not fit");
    static_assert(TC || (C::C1 * PR <= O_XR && C::C1 * PR <= (NWORK - 1) * ACT), "rf_pre scratch does not fit");
    static_assert(ZBF <= O_XR, "rf_post scratch does not fit");
    static_assert(S * C::N_FFT <= ACT, "FFT buffers do not fit");
    // ---- shared memory map (float offsets) ----
    static constexpr int NSK = C::E + 1;
    static constexpr int SM_FIXED = AB + SM_REST;
    static_assert(SM_FIXED <= 227 * 256, "shared memory plan exceeds 227 KB even with every skip tensor spilled");
    static constexpr int SKIP_SMEM = cmax(0, cmin(cmin(NSK, T::SKIP_SMEM_MAX), (227 * 256 - SM_FIXED) / ACT));
    // ---- tensor memory (TC variants): accumulators, and -- when the columns fit -- the MMA A operands of the RNNFormer:
    //      the TF32 copy of x and the GRU state of all K blocks (one fp32 column per channel, TMEM lane = position).  An MMA whose
    //      A operand comes from TMEM costs ~33 cycles against ~100 from shared memory (tools/tc_bench2.cu), the epilogue threads
    //      already own the rows (lanes) they write, and TMEM is never aliased, so the K-padding columns are zeroed once. ----
    static constexpr int ACCW = cmax(cmax(32, 4 * NPG), cmax(cdiv(S * C::F1, 128) * round_up(C::C1, 16), round_up(QN, 16)));
#ifndef FE_HTMEM
#define FE_HTMEM 1
#endif
#ifndef FE_RF16
#define FE_RF16 1
#endif
    // fp16 variants with TMEM operands run the RNNFormer MMAs on fp16 too (RF16): x and h as packed halves (two channels per column,
    // K = 16 per MMA), beside an fp32 master of h for the state update.
    static constexpr int C2H = round_up(C::C2, 16);
    static constexpr bool WANT_RF16 = H16 && !BF16 && FE_RF16;     // the bf16 variants keep a TF32 RNNFormer ("fp32 GRU")
    static constexpr int XH = NPART * (C2H / 2);       // columns of one packed-halves operand (x, or the state of one block): hi part, then (split) lo part
    static constexpr int TM_COLS_TF32 = ACCW + C2P * (C::K + 1);
    static constexpr int TM_COLS_F16 = ACCW + XH + C::K * (C2P + XH);
    static constexpr bool H_TMEM = TC && FE_HTMEM && !RM64 && (WANT_RF16 ? TM_COLS_F16 : TM_COLS_TF32) <= 512;
    static constexpr bool RF16 = WANT_RF16 && H_TMEM;
    static_assert(!SPLIT || RF16, "split variants exist only where the RNNFormer operands fit tensor memory");
    static constexpr int TM_XT = ACCW;                 // x as MMA operand: C2P columns (TF32) or XH columns (packed halves)
    static constexpr int TM_H = ACCW + (RF16 ? XH : C2P);          // K x C2P columns: fp32 GRU state, resident for the whole launch
    static constexpr int TM_H16 = TM_H + C::K * C2P;   // RF16: K x XH columns: the state as packed halves (MMA operand)
    static constexpr int TM_COLS = RF16 ? TM_COLS_F16 : TM_COLS_TF32;
    // TC variants keep the GRU state of all K blocks on chip across hops: in TMEM, else in shared memory when it fits
    static constexpr bool H_RES = TC && (H_TMEM || SM_FIXED + SKIP_SMEM * ACT + C::K * XTS <= 227 * 256);
    static constexpr int SM_SK = 0;
    // position `pos` (0 .. N-1) of stream s in an input / overlap-add ring
    FE_HD static constexpr int ring_off(int s, int pos) {
        return HOP_RING ? ((pos / HT) * S + s) * HT + pos % HT : s * C::N_FFT + pos;
    }
    static constexpr int SM_HST = SM_SK + SKIP_SMEM * ACT;          // [K][XTS] resident GRU state (GeoR) unless it lives in TMEM
    static constexpr int SM_W = SM_HST + ((H_RES && !H_TMEM) ? C::K * XTS : 0);
    static constexpr int SM_SPEC = SM_W + AB;
    // Input / overlap-add rings: the last N input samples and the N-sample overlap-add accumulator of every stream, circular.
    // HOP_RING (hop divides n_fft: T / B / S): the rings are tiled [N / HT][S][HT] so that one hop of all S streams is a dense [S][HT]
    // box (HT = hop, or 256 when the hop is longer: TMA boxes hold at most 256 elements per dimension) -- the 2-D TMA tile
    // (cp.async.bulk.tensor) of the input hop lands straight in the input ring and the output hop leaves straight from the
    // overlap-add ring.  Otherwise (M / L: hop does not divide n_fft) the rings are [S][N] and the hop moves with plain loads / stores.
#ifndef FE_HOP_RING
#define FE_HOP_RING 1
#endif
    static constexpr bool HOP_RING = TC && FE_HOP_RING && C::N_FFT % C::HOP == 0;
    static constexpr bool SLICED = !HOP_RING;          // hop-sliced streaming launches (fe_kernel.cuh::Frame::run) are compiled in
    static constexpr int HT = C::HOP > 256 ? 256 : C::HOP;
    static_assert(!HOP_RING || (C::HOP % HT == 0 && (HT * 4) % 16 == 0), "hop tile");
    static constexpr int SM_TIN = round_up(SM_SPEC + SPECF, 32);      // 128-byte aligned (TMA destination)
    static constexpr int SM_OLA = SM_TIN + S * C::N_FFT;
    // Streaming launches overlap the back end of hop t (mask, inverse FFT, overlap-add: half of the threads) with the front end of
    // hop t + 1 (window, FFT, compression: the other half) when two more FFT buffers fit: half as many barrier-separated phases.
#ifndef FE_FB_OVL
#define FE_FB_OVL 1
#endif
    static constexpr int SM_FF = SM_OLA + S * C::N_FFT;        // [2][S][N] front-end FFT buffers of the overlapped schedule
    static constexpr bool FB_OVL = TC && FE_FB_OVL &&
        (SM_FF + 2 * S * C::N_FFT + STAGES * CHUNK + 4 * STAGES + 8) * 4 <= 227 * 1024;
    static constexpr int SM_RING = SM_FF + (FB_OVL ? 2 * S * C::N_FFT : 0);
    static constexpr int SM_BAR = SM_RING + STAGES * CHUNK;    // 2*STAGES mbarriers (8 bytes each)
    static constexpr int SM_TOTAL = SM_BAR + 4 * STAGES + 8;   // + accumulator-ready mbarrier, TMEM base slot, hop-tile full / empty mbarriers
    static_assert(SM_RING % 4 == 0 && SM_BAR % 2 == 0, "alignment");
    static constexpr int SMEM_BYTES = SM_TOTAL * 4;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory plan exceeds 227 KB");
    // global scratch per CTA: spilled skip tensors
    static constexpr int GS_TOTAL = (NSK - SKIP_SMEM) * ACT + 4;
    static_assert(C1P == C::C1 || SKIP_SMEM == NSK, "a channel-padding slab and spilled skip tensors do not mix (the spill never holds the zero slab)");
    // frame-parallel offline schedule (fp32 family): per-group scratch between the stages = [spilled skips | spectrum | resident skips | x]
    static constexpr int TP_O_SPEC = round_up(GS_TOTAL, 4), TP_O_SK = TP_O_SPEC + round_up(SPECF, 4), TP_O_XR = TP_O_SK + SKIP_SMEM * round_up(ACT, 4);
    static constexpr int TP_GROUP = TP_O_XR + round_up(XRS, 4);

    // ---- layers (weight-stream order) ----
    using EncPre = PosGemm<S * C::F1, C::F1, C::C1, 8, 3, 4, T::CT_CONV, NW, CHUNK>;
    using Conv3 = PosGemm<S * C::F1, C::F1, C::C1, C::C1, 3, 4, T::CT_CONV, NW, CHUNK>;
    using PwCat = PosGemm<S * C::F1, C::F1, C::C1, 2 * C::C1, 1, 4, T::CT_CONV, NW, CHUNK>;
    using ConvT = PosGemm<S * C::F1, C::F1, 8, C::C1, 3, 4, 1, NW, CHUNK>;
    using LinPre = RowGemm<C::C1 * S, C::F1, C::F2, NW, CHUNK>;
    using RfPre = PosGemm<S * C::F2, C::F2, C::C2, C::C1, 1, 4, T::CT_RF, NW, CHUNK>;
    using Gru = PosGemm<S * C::F2, C::F2, C::C2, C::C2, 1, T::PT_GRU, T::CT_GRU, NW, CHUNK, 6>;
    using Fc = PosGemm<S * C::F2, C::F2, C::C2, C::C2, 1, 4, T::CT_RF, NW, CHUNK>;
    using Qkv = PosGemm<S * C::F2, C::F2, 3 * C::HD * HG, C::C2, 1, 4, T::CT_RF, NW, CHUNK>;
    using LinPost = RowGemm<C::C2 * S, C::F2, C::F1, NW, CHUNK>;
    using RfPost = PosGemm<S * C::F1, C::F1, C::C1, C::C2, 1, 4, T::CT_CONV, NW, CHUNK>;
    static_assert(PwCat::NPASS == 1, "the concatenating 1x1 conv stores in place after a barrier: one pass only");

    // tensor-core versions of the conv-type layers (TC variants only)
    // (K = padded input channels: the packer fills the padding with zero weights)
    // enc_pre reads the fp32 spectrum as TF32 operands; the split variants read its fp16 hi / lo copy (K = 8 padded to one k-step of 16)
    using TEncPre = TcGemm<S * C::F1, C::C1, SPLIT ? 16 : 8, 3, CHUNK, 256, SPLIT ? 16 : 8, NPART>;
    using TConv3 = TcGemm<S * C::F1, C::C1, C1P, 3, CHUNK, 256, KEC, NPART>;
    using TPwCat = TcGemm<S * C::F1, C::C1, 2 * C1P, 1, CHUNK, 256, KEC, NPART>;
    using TConvT = TcGemm<S * C::F1, 8, C1P, 3, CHUNK, 256, KEC, NPART>;
    using TRfPost = TcGemm<S * C::F1, C::C1, C2Z, 1, CHUNK, 256, KEC, NPART>;
    using LinPreT = RowGemmK1<C::C1 * S, C::F1, C::F2, NW, CHUNK>;
#ifndef FE_LIN_TC
#define FE_LIN_TC 1
#endif
    // frequency-axis linears on the tensor cores (16-bit variants; slot counts must be whole k-steps)
    static constexpr bool LIN_TC = H16 && FE_LIN_TC && (S * C::F2) % 16 == 0 && (S * C::F1) % 16 == 0 && S * C::F2 <= 128;
    using TLinPre = TcLin<S * C::F2, LIN_TC ? C1P : 16, S * C::F1, CHUNK, NPART>;
    using TLinPost = TcLin<S * C::F1, LIN_TC ? C2Z : 16, LIN_TC ? S * C::F2 : 16, CHUNK, NPART>;
    static_assert(!LIN_TC || (TLinPre::NMT * TLinPre::N <= ACCW && TLinPost::NMT * TLinPost::N <= ACCW), "frequency-axis linear accumulators exceed the TMEM window");
    static constexpr int TMEMC = pow2ceil(H_TMEM ? TM_COLS : ACCW);
    static_assert(TMEMC <= 512, "TMEM columns");
    using TRfPre = TcGemm<S * C::F2, C::C2, C1P, 1, CHUNK, 512, KEC, NPART>;
    static constexpr int KER = RF16 ? 16 : 8;          // contraction length of one RNNFormer MMA
    using TGru = TcGru<S * C::F2, C::C2, CHUNK, KER, NPART>;
    using TFc = TcGemm<S * C::F2, C::C2, RF16 ? C2H : C::C2, 1, CHUNK, 512, KER, NPART>;
    using TQkv = TcGemm<S * C::F2, QN, RF16 ? C2H : C::C2, 1, CHUNK, 512, KER, NPART>;
    using LinPostT = RowGemmK1<C::C2 * S, C::F2, C::F1, NW, CHUNK>;
    // (An experimental variant of these two layers that stored every weight twice gave intermittently wrong rf_pre outputs on the GPU
    // for 48 kHz L while the CPU emulation was exact.  It was slower anyway and is gone; the ring itself is not the cause -- capping
    // the rows per chunk with FE_LIN_KC_MAX, which gives these layers 4-5 chunks on the 2-stage ring, is bit-identical on the GPU.)
    static constexpr int BLK_CHUNKS = Gru::NCHUNK + 2 * Fc::NCHUNK + NQG * Qkv::NCHUNK;
    static constexpr long BLK_FLOATS = (long)Gru::FLOATS + 2 * Fc::FLOATS + NQG * Qkv::FLOATS;
    static constexpr int TBLK_CHUNKS = TGru::NCHUNK + 2 * TFc::NCHUNK + NQG * TQkv::NCHUNK;
    static constexpr long TBLK_FLOATS = (long)TGru::FLOATS + 2 * TFc::FLOATS + NQG * TQkv::FLOATS;
    static constexpr int NCHUNK_FRAME = TC
        ? TEncPre::NCHUNK + C::E * TConv3::NCHUNK + (LIN_TC ? TLinPre::NCHUNK : LinPreT::NCHUNK) + TRfPre::NCHUNK + C::K * TBLK_CHUNKS +
              (LIN_TC ? TLinPost::NCHUNK : LinPostT::NCHUNK) +
              TRfPost::NCHUNK + C::E * (TPwCat::NCHUNK + TConv3::NCHUNK) + TPwCat::NCHUNK + TConvT::NCHUNK
        : EncPre::NCHUNK + C::E * Conv3::NCHUNK + LinPre::NCHUNK + RfPre::NCHUNK + C::K * BLK_CHUNKS + LinPost::NCHUNK +
              RfPost::NCHUNK + C::E * (PwCat::NCHUNK + Conv3::NCHUNK) + PwCat::NCHUNK + ConvT::NCHUNK;
    static constexpr long RING_FLOATS = TC
        ? (long)TEncPre::FLOATS + (long)C::E * TConv3::FLOATS + (LIN_TC ? TLinPre::FLOATS : LinPreT::FLOATS) + TRfPre::FLOATS + (long)C::K * TBLK_FLOATS +
              (LIN_TC ? TLinPost::FLOATS : LinPostT::FLOATS) + TRfPost::FLOATS + (long)C::E * (TPwCat::FLOATS + TConv3::FLOATS) + TPwCat::FLOATS + TConvT::FLOATS
        : (long)EncPre::FLOATS + (long)C::E * Conv3::FLOATS + LinPre::FLOATS + RfPre::FLOATS + (long)C::K * BLK_FLOATS +
              LinPost::FLOATS + RfPost::FLOATS + (long)C::E * (PwCat::FLOATS + Conv3::FLOATS) + PwCat::FLOATS + ConvT::FLOATS;

    // ---- blob layout (float offsets): [aux tables + biases | chunk table | ring section] ----
    // Per-layer arrays are affine (base + index * stride) so that device code never indexes a
    // constexpr array at run time (which would materialise the table on the stack).
    struct Blk { int b_r, b_z, b_in, b_hn, fc_b, pe, qkv_b, afc_b; };
    struct Aux {
        int window, window_istft, window_sq;   // [N] each
        int tw, twn;                           // float2[M/2] : exp(-2 pi i t / M) ; float2[M] : exp(-2 pi i k / N)
        int enc_pre_b, enc_b0, enc_bs;         // encoder[i] bias at enc_b0 + i*enc_bs
        int rf_pre_b;
        int blk0, blks;                        // block k at blk0 + k*blks + rel.*
        Blk rel;
        int rf_post_b;
        int dec_b0, dec_bs, dec2_rel;          // decoder[i]: 1x1 bias at dec_b0 + i*dec_bs, k=3 bias dec2_rel after it
        int dp_b, convt_b;
        int zeros;                             // 4 zeros (stand-in for the positional embedding in blocks > 0)
        int flags;                             // [4]: flags[0] != 0 <=> some attention qkv bias is non-zero (attn_bias: False in every shipped config)
        int table;                             // int32[2*NCHUNK_FRAME] : (float offset from blob start, floats)
        int ring;                              // start of the ring section
        int total;
        constexpr int enc_b(int i) const { return enc_b0 + i * enc_bs; }
        constexpr int dec1_b(int i) const { return dec_b0 + i * dec_bs; }
        constexpr int dec2_b(int i) const { return dec_b0 + i * dec_bs + dec2_rel; }
        constexpr Blk blk(int k) const {
            const int o = blk0 + k * blks;
            return Blk{o + rel.b_r, o + rel.b_z, o + rel.b_in, o + rel.b_hn, o + rel.fc_b, o + rel.pe, o + rel.qkv_b, o + rel.afc_b};
        }
    };
    static constexpr Aux make_aux() {
        Aux a{};
        int o = 0;
        auto take = [&o](int n) { int r = o; o += round_up(n, 4); return r; };
        a.window = take(C::N_FFT); a.window_istft = take(C::N_FFT); a.window_sq = take(C::N_FFT);
        a.tw = take(C::M); a.twn = take(2 * C::M);
        a.enc_pre_b = take(C::C1);
        a.enc_bs = round_up(C::C1, 4); a.enc_b0 = take(C::E * a.enc_bs);
        a.rf_pre_b = take(C::C2);
        {
            int r = 0;
            auto rel = [&r](int n) { int q = r; r += round_up(n, 4); return q; };
            a.rel.b_r = rel(C::C2); a.rel.b_z = rel(C::C2); a.rel.b_in = rel(C::C2); a.rel.b_hn = rel(C::C2);
            a.rel.fc_b = rel(C::C2); a.rel.pe = rel(C::C2 * C::F2); a.rel.qkv_b = rel(C::NH * 3 * round_up(C::HD, 4)); a.rel.afc_b = rel(C::C2);
            a.blks = r; a.blk0 = take(C::K * r);
        }
        a.rf_post_b = take(C::C1);
        a.dec2_rel = round_up(C::C1, 4); a.dec_bs = 2 * a.dec2_rel; a.dec_b0 = take(C::E * a.dec_bs);
        a.dp_b = take(C::C1); a.convt_b = take(16);
        a.zeros = take(4);
        a.flags = take(4);
        a.table = take(2 * NCHUNK_FRAME);
        a.ring = o;
        a.total = o + (int)RING_FLOATS;
        return a;
    }
};

// kernel parameters (plain data, passed by value)
struct KParams {
    const float* blob;        // packed weights + tables (device)
    float* state;             // planes: cache_stft [n_streams][N-H] | cache_istft [n_streams][N-H] | h_k [n_streams][F2][C2], k < K
    const float* in;          // mode 0/3: wav [n_streams][ld_in]; mode 1/4: spec [B][NB][T][2]; mode 2: wav [B][L]
    float* out;               // mode 0/4: wav [n_streams][ld_out]; mode 1/3: spec [B][NB][T][2]; mode 2: wav [B][H*(T-1)]
    float* spec_out;          // mode 2 (optional): compressed masked spectrum [B][FIN][T][2]
    float* scratch;           // global scratch [grid][GS_TOTAL]
    float* dbg;               // optional tap dump of stream 0 (oracle tap layout), frame `dbg_hop`
    long long* prof;          // optional [PH_COUNT] cycle counters of CTA 0 (accumulated over the launch)
    long long ld_in, ld_out;
    int n_streams, n_hops;    // n_hops = T frames in modes 1, 2
    int mode, L, dbg_hop;
    int hop_tma;              // streaming launches of HOP_RING variants: the input hop arrives / the output hop leaves as 2-D TMA tiles
    const void* tmaps;        // ... described by two CUtensorMap (input, output) in global memory (device; 64-byte aligned)
    float compression;
    // ---- hop-sliced streaming launches: items (hop range, stream group) on a persistent grid (fe_kernel.cuh::run) ----
    int slice_hops;           // hops per range; 0 = one CTA per stream group walks all hops
    int* slice_flags;         // [ranges][groups] item-done flags (zeroed before the launch)
    // ---- frame-parallel offline schedule (MODE_OFFLINE, fp32 family; fe_api.cu::offline_tp): the CTAs take groups of S FRAMES (slot =
    //      frame q = group * S + s of the n_streams * n_hops frames; utterance q / n_hops, frame q % n_hops) instead of S streams ----
    int tp_stage;             // 0 = off; 1 = stage A: front end, encoder, rf_pre, input half of GRU 0; 2 = stage B of block tp_blk: rnn_fc,
                              // attention, attn_fc, then the input half of GRU tp_blk + 1 or (last block) rf_post .. inverse FFT
    int tp_blk;
    float* tp_scr;            // [groups][Plan::TP_GROUP] spectrum, skip tensors, residual stream between the stages
    float* tp_gx;             // [frames][F2][3][C2] input-side gate pre-activations (W_ir x | W_iz x | W_in x, no bias)
    const float* tp_h;        // [frames][F2][C2] hidden states of block tp_blk (output of the scan)
    float* tp_frames;         // [frames][N] windowed output frames (input of the overlap-add)
};

// MODE_STFT / MODE_ISTFT: the front / back end alone (ONNXSTFT.forward / .inverse, functional/audio_modules.py:243-303);
// they consume no weights, so the producer warp stays idle.
enum { MODE_STREAM = 0, MODE_SPEC = 1, MODE_OFFLINE = 2, MODE_STFT = 3, MODE_ISTFT = 4 };

}  // namespace fe
