// fe_half.h -- host-side IEEE binary16 / bfloat16 conversion (round to nearest even), shared by the weight packer and the CPU emulation.
#pragma once
#include <cstdint>
#include <cstring>

namespace fe {

inline uint16_t f32_to_f16_bits(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0u));     // inf / nan
    if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);                                      // rounds to >= 65520: inf
    if (x < 0x38800000u) {                                                                        // subnormal half (or zero)
        if (x < 0x33000000u) return (uint16_t)sign;                                               // < 2^-25: zero
        const int e = (int)(x >> 23);                                                             // biased float exponent, 102..112
        const uint32_t m = (x & 0x7fffffu) | 0x800000u;                                           // 24-bit significand
        const int shift = 126 - e;                                                                // half = m * 2^(e-150) / 2^-24 = m >> (126 - e)
        uint32_t h = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (h & 1u))) ++h;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((x - 0x38000000u) >> 13);                                                       // rebias 127 -> 15, 10-bit mantissa
    const uint32_t rem = x & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;                                       // may carry into the exponent: still right
    return (uint16_t)(sign | h);
}

inline float f16_bits_to_f32(uint16_t h) {
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) x = sign;
        else {                                                                                    // subnormal: normalise
            int s = 0;
            uint32_t mm = m;
            while (!(mm & 0x400u)) { mm <<= 1; ++s; }
            x = sign | ((uint32_t)(113 - s) << 23) | ((mm & 0x3ffu) << 13);
        }
    } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e + 112u) << 23) | (m << 13);
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}

// bfloat16: the top 16 bits of an fp32, round to nearest even (what cvt.rn.bf16x2.f32 does); finite inputs only
inline uint16_t f32_to_bf16_bits(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    if ((x & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((x >> 16) | 0x40u);                     // nan stays nan
    x += 0x7fffu + ((x >> 16) & 1u);
    return (uint16_t)(x >> 16);
}
inline float bf16_bits_to_f32(uint16_t h) {
    const uint32_t x = (uint32_t)h << 16;
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}
// 16-bit operand format of a tensor-core variant: fp16 (FMT 1) or bfloat16 (FMT 2)
template <bool BF> inline uint16_t f32_to_h16_bits(float f) { return BF ? f32_to_bf16_bits(f) : f32_to_f16_bits(f); }
template <bool BF> inline float h16_bits_to_f32(uint16_t h) { return BF ? bf16_bits_to_f32(h) : f16_bits_to_f32(h); }

}  // namespace fe
