// Small host-side helpers shared by the kernel launchers: per-thread error text and fp16 bits.
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>

namespace oidnb200 {

void set_error(const std::string& msg);

// IEEE binary16 bits -> float, exact (same value set as common/half.cpp:37-99 in the reference).
inline float half_bits_to_float(uint16_t h)
{
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1Fu;
  uint32_t man = h & 0x3FFu;
  uint32_t bits;
  if (exp == 0)
  {
    if (man == 0)
      bits = sign;
    else
    {
      int e = -1;
      do { man <<= 1; ++e; } while (!(man & 0x400u));
      man &= 0x3FFu;
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
    }
  }
  else if (exp == 31)
    bits = sign | 0x7F800000u | (man << 13);
  else
    bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

// float -> binary16 bits, round to nearest even.
inline uint16_t float_to_half_bits(float f)
{
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7FFFFFFFu;
  if (x >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | ((x > 0x7F800000u) ? 0x200u : 0));
  if (x >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u); // overflow -> inf
  if (x < 0x38800000u)
  {
    // subnormal half or zero
    if (x < 0x33000000u) return (uint16_t)sign;
    const int shift = 113 - (int)(x >> 23);
    uint32_t man = (x & 0x7FFFFFu) | 0x800000u;
    const uint32_t r = man >> (shift + 13);
    const uint32_t rem = man & ((1u << (shift + 13)) - 1);
    const uint32_t half = 1u << (shift + 12);
    uint32_t out = r;
    if (rem > half || (rem == half && (r & 1))) out++;
    return (uint16_t)(sign | out);
  }
  uint32_t out = ((x - 0x38000000u) >> 13);
  const uint32_t rem = x & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (out & 1))) out++;
  return (uint16_t)(sign | out);
}

} // namespace oidnb200
