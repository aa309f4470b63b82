/*
 * drv_r11g11b10.h — the R11F_G11F_B10F texel format of the specular environment-map atlas
 * (renderer.cpp:282; OpenGL 4.5 section 2.3.4.3/4: unsigned 11- and 10-bit floats, 5 exponent bits, bias 15,
 * 6 / 5 mantissa bits; texel = R | G << 11 | B << 22). One definition for host and device code.
 * Policy where GL leaves a choice: float -> small float rounds to nearest even; negative values and NaN store 0;
 * values above the largest finite small float (65024 / 64512) store that value.
 */
#ifndef DRV_R11G11B10_H
#define DRV_R11G11B10_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define DRV_HD __host__ __device__ inline
#else
#define DRV_HD static inline
#endif

/* MBITS = 6 (11-bit) or 5 (10-bit) */
DRV_HD uint32_t drv_float_to_small(float f, int mbits) {
  uint32_t x;
  memcpy(&x, &f, 4);
  if (x & 0x80000000u) return 0u;              /* negative (and -0) */
  if (x > 0x7f800000u) return 0u;              /* NaN */
  const uint32_t maxv = (30u << mbits) | ((1u << mbits) - 1u); /* largest finite */
  const int e = (int)(x >> 23) - 127;
  if (e > 15) return maxv;                     /* >= 65536 or inf */
  if (e < -15 - mbits) return 0u;              /* below half the smallest subnormal... rounds to 0 (checked below for ties) */
  uint32_t man = (x & 0x7fffffu) | 0x800000u;  /* 24-bit significand */
  int shift;                                    /* bits dropped */
  uint32_t hexp;
  if (e < -14) { shift = (23 - mbits) + (-14 - e); hexp = 0u; }
  else { shift = 23 - mbits; hexp = (uint32_t)(e + 15); }
  if (shift > 24) return 0u;
  uint32_t q = man >> shift;
  const uint32_t rem = man & ((1u << shift) - 1u), half = 1u << (shift - 1);
  if (rem > half || (rem == half && (q & 1u))) q++;
  uint32_t out = hexp == 0u ? q : ((hexp << mbits) + (q - (1u << mbits))); /* a carry runs into the exponent */
  return out > maxv ? maxv : out;
}
DRV_HD float drv_small_to_float(uint32_t v, int mbits) {
  const uint32_t e = v >> mbits, m = v & ((1u << mbits) - 1u);
  float r;
  if (e == 0u) {
    r = (float)m * (1.0f / (float)(1u << mbits)) * 6.103515625e-05f; /* m / 2^mbits * 2^-14 */
  } else if (e == 31u) {
    const uint32_t bits = 0x7f800000u | (m << (23 - mbits));
    memcpy(&r, &bits, 4);
  } else {
    const uint32_t bits = ((e + 112u) << 23) | (m << (23 - mbits));
    memcpy(&r, &bits, 4);
  }
  return r;
}
DRV_HD uint32_t drv_pack_r11g11b10(float r, float g, float b) {
  return drv_float_to_small(r, 6) | (drv_float_to_small(g, 6) << 11) | (drv_float_to_small(b, 5) << 22);
}
DRV_HD void drv_unpack_r11g11b10(uint32_t t, float* r, float* g, float* b) {
  *r = drv_small_to_float(t & 0x7ffu, 6);
  *g = drv_small_to_float((t >> 11) & 0x7ffu, 6);
  *b = drv_small_to_float(t >> 22, 5);
}
#endif
