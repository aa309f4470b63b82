// device_math.cuh — arithmetic helpers for the sm_100a kernels.
//
// Two kinds of maths live on this path:
//  * DECISION maths — anything that feeds a float->int truncation, a
//    comparison against a box, or a loop trip count (world position -> CAV
//    cell, cascade choice, voxel coverage, cone-march distances). These use
//    the ex_* helpers: IEEE round-to-nearest, every multiply and add rounded
//    on its own (never contracted into FMA), fixed left-to-right summation,
//    so the result is bit-identical to the shader-transcribing oracle
//    (DESIGN.md "Parity policy").
//  * CONTINUOUS maths — the radiance / SH / filtering arithmetic, where the
//    1e-3 / 1e-5 gate applies; these are free to use FMA, MUFU approximations
//    and re-association.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace drvk {

__device__ __forceinline__ float ex_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float ex_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float ex_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ex_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float ex_sqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float ex_rsqrt(float a) { return __fdiv_rn(1.0f, __fsqrt_rn(a)); } // inversesqrt policy
__device__ __forceinline__ float ex_dot3(float ax, float ay, float az, float bx, float by, float bz) {
  return ex_add(ex_add(ex_mul(ax, bx), ex_mul(ay, by)), ex_mul(az, bz));
}
__device__ __forceinline__ float ex_dot4(const float* r, float v0, float v1, float v2, float v3) {
  return ex_add(ex_add(ex_add(ex_mul(r[0], v0), ex_mul(r[1], v1)), ex_mul(r[2], v2)), ex_mul(r[3], v3));
}
__device__ __forceinline__ float ex_mix(float a, float b, float t) { // a*(1-t) + b*t
  return ex_add(ex_mul(a, ex_sub(1.0f, t)), ex_mul(b, t));
}
// float -> int: truncate, saturate, NaN -> 0 (cvt.rzi.s32.f32 semantics)
__device__ __forceinline__ int ex_trunc(float f) { return __float2int_rz(f); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ float saturatef(float x) { return __saturatef(x); }

struct F3 { float x, y, z; };

// `vec4(v,1) * M` on the raw row-major ei bytes, then the perspective divide
// (cacheGather.comp:118-119, cacheApply.frag:134-135).
// row . (v, 1): the last product is r[3] * 1.0f, which is r[3] bit for bit (NaN payloads aside) — no multiply issued
__device__ __forceinline__ float ex_dot4_w1(const float* r, float v0, float v1, float v2) {
  return ex_add(ex_add(ex_add(ex_mul(r[0], v0), ex_mul(r[1], v1)), ex_mul(r[2], v2)), r[3]);
}
__device__ __forceinline__ F3 ex_unproject(const float* m, float x, float y, float z) {
  float w0 = ex_dot4_w1(m + 0, x, y, z);
  float w1 = ex_dot4_w1(m + 4, x, y, z);
  float w2 = ex_dot4_w1(m + 8, x, y, z);
  float w3 = ex_dot4_w1(m + 12, x, y, z);
  F3 r = {ex_div(w0, w3), ex_div(w1, w3), ex_div(w2, w3)};
  return r;
}

// Morton_2D_Decode_16bit, cacheLightingRSM.comp:46-62.
__device__ __forceinline__ uint32_t compact_bits(uint32_t v) {
  v &= 0x55555555u;
  v = (v | (v >> 1)) & 0x33333333u;
  v = (v | (v >> 2)) & 0x0f0f0f0fu;
  v = (v | (v >> 4)) & 0x00ff00ffu;
  v = (v | (v >> 8)) & 0x0000ffffu;
  return v;
}
__device__ __forceinline__ void morton_decode(uint32_t k, uint32_t& x, uint32_t& y) {
  x = compact_bits(k);
  y = compact_bits(k >> 1);
}

__device__ __forceinline__ float half_bits_to_float(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ uint16_t float_to_half_bits(float f) { return __half_as_ushort(__float2half_rn(f)); }

#define DRV_GLSL_PI 3.14159265358979f

// UnpackNormal16I, utils.glsl:44-58 (continuous maths).
__device__ __forceinline__ F3 unpack_normal16i(int px, int py) {
  float a = (float)px * (DRV_GLSL_PI / 32768.0f);
  float z = (float)py * (1.0f / 32768.0f);
  float sinPhi = sqrtf(1.0f - z * z);
  float s, c;
  sincosf(a, &s, &c);
  float x = c * sinPhi, y = s * sinPhi;
  float inv = rsqrtf(x * x + y * y + z * z);
  F3 r = {x * inv, y * inv, z * inv};
  return r;
}

// The same with MUFU sin/cos (|a| <= pi: absolute error ~1e-6) for the per-pixel apply stage.
__device__ __forceinline__ F3 unpack_normal16i_fast(int px, int py) {
  float a = (float)px * (DRV_GLSL_PI / 32768.0f);
  float z = (float)py * (1.0f / 32768.0f);
  float sinPhi = sqrtf(fmaxf(1.0f - z * z, 0.0f));
  float s, c;
  __sincosf(a, &s, &c);
  float x = c * sinPhi, y = s * sinPhi;
  float inv = rsqrtf(x * x + y * y + z * z);
  F3 r = {x * inv, y * inv, z * inv};
  return r;
}

// PackNormal16I, utils.glsl:62-89, int16 clamp.
__device__ __forceinline__ void pack_normal16i(float nx, float ny, float nz, int& ox, int& oy) {
  float sgn = ny > 0.0f ? 1.0f : (ny < 0.0f ? -1.0f : 0.0f);
  float px = (nx == 0.0f) ? (sgn * DRV_GLSL_PI / 2) : atan2f(ny, nx);
  ox = clampi(__float2int_rz(px * (32768.0f / DRV_GLSL_PI)), -32768, 32767);
  oy = clampi(__float2int_rz(nz * 32768.0f), -32768, 32767);
}

} // namespace drvk
