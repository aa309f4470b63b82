// voxel_sample.cuh — reading the voxel chain through its gather-ready records: footprint of a sample in one mip
// level, one 64-bit load per footprint, trilinear filter in packed FP32x2. Shared by the cone pass of the gather
// (gather.cu) and the cone-traced ambient occlusion (adjacent.cu).
#pragma once

#include "ctx.h"
#include "device_math.cuh"

namespace drvk {

// ---------------------------------------------------------------- cone trace
// The voxel chain is read through its gather-ready copy (voxel.cu: voxel_records_kernel): one 64-bit load
// returns the eight clamp-to-edge texels of a trilinear footprint, instead of eight byte loads with per-corner
// address clamping. Records of level l start at rec_offset[l]; a level holds (r + 1 + 2 pad)^3 of them, indexed by the
// footprint's lower corner + 1 + pad (kVoxelRecordPad, ctx.h).
struct VoxelVol {
  const uint2* rec;
  const uint32_t* rec_offset; // GatherParams::rec_offset in the kernel's constant bank (LDC with a register index)
  int res, levels;
  float vmin[3];
  float voxel_size;
};

constexpr float kMagic = 12582912.0f; // 1.5 * 2^23: (v + kMagic) - kMagic rounds v to the nearest integer

// The cone march is issue-bound (ncu: issue slots 75 % busy, FMA pipe 38 %), so everything that comes in x / y
// pairs is evaluated with packed FP32x2 instructions (FFMA2 / FADD2: one issue slot for two IEEE-identical
// results): positions, texel coordinates, floor / fraction, and the x- and y-lerps of the trilinear filter.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }

// floor(v) without the XU pipe (FRND / F2I / I2F are quarter rate). The caller passes w = v - 0.5; w + kMagic
// rounds w to the nearest integer = floor(v) (on exact integers ties-to-even may pick v - 1, in which case
// frac = 1 and the trilinear result is the same: the filter is continuous across texel boundaries).
// i = floor as int, f = v - floor.
__device__ __forceinline__ void floor_frac_w(float w, int& i, float& f) {
  const float m = w + kMagic;
  i = __float_as_int(m) - 0x4B400000;
  f = (w - (m - kMagic)) + 0.5f;
}
__device__ __forceinline__ void floor_frac_w2(float2 w, int& ix, int& iy, float2& f) {
  const float2 m = __fadd2_rn(w, f2(kMagic));
  ix = __float_as_int(m.x) - 0x4B400000;
  iy = __float_as_int(m.y) - 0x4B400000;
  f = __fadd2_rn(__fadd2_rn(w, __fadd2_rn(f2(kMagic), f2(-m.x, -m.y))), f2(0.5f)); // (w - (m - magic)) + .5
}

// byte `sel` (0..3) of w as a float, exactly: build 2^23 + byte with one PRMT, subtract 2^23.
template <int SEL>
__device__ __forceinline__ float byte_as_float_biased(uint32_t w) {
  return __int_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | SEL)); // 8388608 + byte
}

struct Footprint {
  uint32_t index; // record index
  float2 txy;
  float tz;
};

// footprint of the sample at q in level `l`; q is given in level-0 texel units with the half-texel shift
// already applied (q = p * res - 0.5 for p in [0,1]^3), so level l sees q * 2^-l + (2^-(l+1) - 0.5)
__device__ __forceinline__ Footprint footprint(const VoxelVol& V, int l, float2 qxy, float qz) {
  const int r = V.res >> l;
  Footprint F;
  int x, y, z;
  if (l == 0) {
    floor_frac_w2(__fadd2_rn(qxy, f2(-0.5f)), x, y, F.txy);
    floor_frac_w(qz - 0.5f, z, F.tz);
  } else {
    const float sc = __int_as_float(0x3f800000 - (l << 23));  // 2^-l
    const float of = fmaf(sc, 0.5f, -1.0f);                    // the level's half-texel shift, minus the 0.5 of floor
    floor_frac_w2(__ffma2_rn(qxy, f2(sc), f2(of)), x, y, F.txy);
    floor_frac_w(fmaf(qz, sc, of), z, F.tz);
  }
  // clamp the lower corner to [-1, r-1]: outside that range both taps of the axis are the same edge texel
  x = min(max(x, -1), r - 1) + 1;
  y = min(max(y, -1), r - 1) + 1;
  z = min(max(z, -1), r - 1) + 1;
  constexpr int P = (int)kVoxelRecordPad; // the record grid is padded by P corners on every side (ctx.h)
  const int D = r + 1 + 2 * P;
  F.index = V.rec_offset[l] + (uint32_t)((x + P) + D * ((y + P) + D * (z + P)));
  return F;
}

// rec.x = texels (x,y,z) 000 100 010 110, rec.y = 001 101 011 111 (one byte each). The pairs are the two z
// planes, so the x-lerp and the y-lerp are packed and only the z-lerp is scalar.
__device__ __forceinline__ float trilinear(uint2 rec, float2 txy, float tz) {
  const float B = 8388608.0f;
  const float2 a00 = f2(byte_as_float_biased<0>(rec.x), byte_as_float_biased<0>(rec.y)); // (x0,y0) at z0 | z1
  const float2 a10 = f2(byte_as_float_biased<1>(rec.x), byte_as_float_biased<1>(rec.y)); // (x1,y0)
  const float2 a01 = f2(byte_as_float_biased<2>(rec.x), byte_as_float_biased<2>(rec.y)); // (x0,y1)
  const float2 a11 = f2(byte_as_float_biased<3>(rec.x), byte_as_float_biased<3>(rec.y)); // (x1,y1)
  const float2 tx = f2(txy.x), ty = f2(txy.y), one = f2(1.0f), mone = f2(-1.0f), mB = f2(-B);
  // (a1 - a0) is exact on the biased values; only the base needs un-biasing
  const float2 cy0 = __ffma2_rn(tx, __ffma2_rn(a00, mone, a10), __fadd2_rn(a00, mB)); // x-lerp at y0
  const float2 cy1 = __ffma2_rn(tx, __ffma2_rn(a01, mone, a11), __fadd2_rn(a01, mB)); // x-lerp at y1
  const float2 c = __ffma2_rn(ty, __ffma2_rn(cy0, mone, cy1), cy0);                   // y-lerp: (z0, z1)
  (void)one;
  return fmaf(tz, c.y - c.x, c.x) * (1.0f / 255.0f);
}


// D.0 sampler (sampler3D, linear / mip-linear / clamp to edge): p in [0,1]^3 volume coordinates, lod as passed to
// textureLod — clamped to [0, levels - 1]; NaN / -inf / negative select level 0.
__device__ __forceinline__ float sample_voxel_records(const VoxelVol& V, float px, float py, float pz, float lod) {
  const float res = (float)V.res;
  const float2 qxy = __ffma2_rn(f2(px, py), f2(res), f2(-0.5f));
  const float qz = fmaf(pz, res, -0.5f);
  const float l = fminf(fmaxf(lod, 0.0f), (float)(V.levels - 1));
  int l0; float t;
  floor_frac_w(l - 0.5f, l0, t);
  if (!(lod > 0.0f)) { l0 = 0; t = 0.0f; }
  const Footprint f0 = footprint(V, l0, qxy, qz);
  float o = trilinear(__ldg(V.rec + f0.index), f0.txy, f0.tz);
  if (t != 0.0f) {
    const Footprint f1 = footprint(V, min(l0 + 1, V.levels - 1), qxy, qz);
    o = fmaf(t, trilinear(__ldg(V.rec + f1.index), f1.txy, f1.tz) - o, o);
  }
  return o;
}

} // namespace drvk
