/*
 * glsl_scalar.h — the GLSL built-ins the path's shaders use, as scalar C++
 * with the arithmetic policy of oracle.h. TEST INFRASTRUCTURE ONLY.
 */
#ifndef DRV_ORACLE_GLSL_SCALAR_H
#define DRV_ORACLE_GLSL_SCALAR_H

#include <cmath>
#include <cstdint>
#include <cstring>
#include <climits>
#include <thread>
#include <vector>
#include <functional>
#include <algorithm>

namespace orc {

struct vec3 {
  float x, y, z;
};
inline vec3 V3(float x, float y, float z) { vec3 r = {x, y, z}; return r; }
inline vec3 V3(const float* p) { vec3 r = {p[0], p[1], p[2]}; return r; }
inline vec3 operator+(vec3 a, vec3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(vec3 a) { return V3(-a.x, -a.y, -a.z); }
inline vec3 operator*(vec3 a, vec3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(vec3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator/(vec3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator/(vec3 a, vec3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline vec3 normalize(vec3 a) { return a * inversesqrt(dot(a, a)); }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3 cross(vec3 a, vec3 b) {
  return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float saturate(float x) { return std::fmin(std::fmax(x, 0.0f), 1.0f); }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

/* float -> int: truncate toward zero, saturate, NaN -> 0. */
inline int trunc_to_int(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.0f) return INT_MAX;
  if (f <= -2147483648.0f) return INT_MIN;
  return static_cast<int>(f);
}
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* GLSL `vec4(v) * M` on the raw row-major ei bytes == ei M * v:
 * component j = dot(row j, v), summed left to right. */
inline void mul_row_major(const float* m, const float v[4], float out[4]) {
  for (int j = 0; j < 4; ++j) {
    const float* r = m + j * 4;
    out[j] = ((r[0] * v[0] + r[1] * v[1]) + r[2] * v[2]) + r[3] * v[3];
  }
}

/* IEEE binary16 <-> binary32, exact / round-to-nearest-even. */
inline float half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu;
  uint32_t man = h & 0x3ffu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else {
      int e = -1;
      do { e++; man <<= 1; } while ((man & 0x400u) == 0);
      man &= 0x3ffu;
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7f800000u | (man << 13);
  } else {
    bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
  }
  float f;
  std::memcpy(&f, &bits, 4);
  return f;
}

inline uint16_t float_to_half(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t absx = x & 0x7fffffffu;
  if (absx >= 0x7f800000u) { /* inf / nan */
    return (uint16_t)(sign | 0x7c00u | ((absx > 0x7f800000u) ? 0x200u : 0u));
  }
  if (absx >= 0x477ff000u) { /* rounds to >= 65520 -> inf */
    return (uint16_t)(sign | 0x7c00u);
  }
  if (absx < 0x33000001u) { /* < 2^-25 (or exactly 2^-25: ties to even -> 0) */
    return (uint16_t)sign;
  }
  int e = (int)(absx >> 23) - 127;
  uint32_t man = (absx & 0x7fffffu) | 0x800000u;
  int shift;
  uint32_t hexp;
  if (e < -14) { /* subnormal half */
    shift = 13 + (-14 - e);
    hexp = 0;
  } else {
    shift = 13;
    hexp = (uint32_t)(e + 15);
  }
  uint32_t q = man >> shift;
  uint32_t rem = man & ((1u << shift) - 1u);
  uint32_t halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (q & 1u))) q++;
  uint32_t out;
  if (hexp == 0) {
    out = q; /* may carry into exponent 1: correct */
  } else {
    out = ((hexp << 10) + (q - 0x400u)); /* q has the implicit bit; carry propagates */
  }
  return (uint16_t)(sign | out);
}

/* cacheLightingRSM.comp:46-62. */
inline void morton_decode(uint32_t morton, uint32_t& cx, uint32_t& cy) {
  uint32_t x = morton, y = morton >> 1;
  x &= 0x55555555u; y &= 0x55555555u;
  x |= x >> 1; y |= y >> 1;
  x &= 0x33333333u; y &= 0x33333333u;
  x |= x >> 2; y |= y >> 2;
  x &= 0x0f0f0f0fu; y &= 0x0f0f0f0fu;
  x |= x >> 4; y |= y >> 4;
  x &= 0x00ff00ffu; y &= 0x00ff00ffu;
  x |= x >> 8; y |= y >> 8;
  cx = x & 0xffffu; cy = y & 0xffffu;
}

constexpr float GLSL_PI = 3.14159265358979f; /* utils.glsl:1 */

/* utils.glsl:44-58. */
inline vec3 unpack_normal16i(int16_t px, int16_t py) {
  float a = (float)px * (GLSL_PI / 32768.0f);
  float z = (float)py * (1.0f / 32768.0f);
  float sinPhi = std::sqrt(1.0f - z * z);
  return normalize(V3(std::cos(a) * sinPhi, std::sin(a) * sinPhi, z));
}

/* utils.glsl:62-89; the int16 render target clamps. */
inline void pack_normal16i(vec3 n, int16_t& ox, int16_t& oy) {
  float px = (n.x == 0.0f) ? (signf(n.y) * GLSL_PI / 2) : std::atan2(n.y, n.x);
  float py = n.z;
  int ix = trunc_to_int(px * (32768.0f / GLSL_PI));
  int iy = trunc_to_int(py * 32768.0f);
  ox = (int16_t)clampi(ix, -32768, 32767);
  oy = (int16_t)clampi(iy, -32768, 32767);
}

/* The exact piecewise sRGB EOTF the sampler applies to an SRGB8 texel
 * (gbuffer.glsl:1, renderer.cpp:468), evaluated in double, rounded once. */
inline float srgb8_to_linear(uint8_t v) {
  double c = (double)v / 255.0;
  double l = (c <= 0.04045) ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4);
  return (float)l;
}

inline int default_threads() {
  unsigned n = std::thread::hardware_concurrency();
  return n == 0 ? 1 : (int)n;
}

/* Static partition of [0,n) over `threads` workers. */
inline void parallel_for(int64_t n, int threads, const std::function<void(int64_t, int64_t, int)>& fn) {
  if (threads <= 0) threads = default_threads();
  if (n <= 0) return;
  if (threads > n) threads = (int)n;
  if (threads <= 1) { fn(0, n, 0); return; }
  std::vector<std::thread> pool;
  pool.reserve(threads);
  int64_t chunk = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    int64_t b = t * chunk, e = std::min<int64_t>(n, b + chunk);
    if (b >= e) break;
    pool.emplace_back(fn, b, e, t);
  }
  for (auto& th : pool) th.join();
}

} // namespace orc
#endif
