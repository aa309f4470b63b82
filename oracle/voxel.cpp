/*
 * voxel.cpp — oracle for stage 3 (voxelise, blend, mip chain).
 * TEST INFRASTRUCTURE ONLY. Restates shader/voxelize.vert:15-23,
 * voxelize.geom:19-112, voxelize.frag:21-58, voxelblend.comp:8-19,
 * voxelmipmap.comp:8-13 and rendering/voxelization.cpp:90-176.
 *
 * The hardware rasteriser's fill rule cannot be reproduced bit for bit in
 * software; following SURVEY D.3 the coverage test below DEFINES the voxel
 * set: a pixel of the res x res viewport is covered when its centre lies on
 * the inner side (inclusive) of the three edge lines of the triangle after
 * they were pushed outwards by half a pixel diagonal (geom:56-80) and inside
 * the half-voxel-dilated bounding box (geom:59-61, frag:24-27).
 */
#include "oracle.h"
#include "glsl_scalar.h"

using namespace orc;

namespace {

struct Plane { float x, y, z; };

inline Plane cross_h(float ax, float ay, float az, float bx, float by, float bz) {
  Plane p = {ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx};
  return p;
}

inline void set_voxel(uint8_t* vol, int res, int side, int px, int py, int pz) {
  /* UnswizzlePos, voxelize.frag:17-20 */
  int x, y, z;
  if (side == 0) { x = pz; y = py; z = px; }
  else if (side == 1) { x = px; y = pz; z = py; }
  else { x = px; y = py; z = pz; }
  if (x < 0 || y < 0 || z < 0 || x >= res || y >= res || z >= res) return; /* OOB imageStore is dropped */
  vol[(size_t)x + (size_t)res * ((size_t)y + (size_t)res * z)] = 255;
}

void voxelize_triangle(const drv_volume_info* vi, int res, const float* tri, const float* world, uint8_t* vol) {
  const float fres = (float)res;
  vec3 vmin = V3(vi->VolumeWorldMin), vmax = V3(vi->VolumeWorldMax);
  vec3 clip[3];
  for (int i = 0; i < 3; ++i) { /* voxelize.vert:20-22 */
    float v[4] = {tri[i * 3 + 0], tri[i * 3 + 1], tri[i * 3 + 2], 1.0f};
    float w4[4];
    mul_row_major(world, v, w4);
    vec3 wp = V3(w4[0], w4[1], w4[2]);
    vec3 q = (wp - vmin) / (vmax - vmin);
    clip[i] = V3(q.x * 2.0f - 1.0f, q.y * 2.0f - 1.0f, q.z * 2.0f - 1.0f);
  }
  /* voxelize.geom:21-26 */
  vec3 nrm = normalize(cross(clip[1] - clip[0], clip[2] - clip[0]));
  float an[3] = {std::fabs(nrm.x), std::fabs(nrm.y), std::fabs(nrm.z)};
  int side = an[0] > an[1] ? 0 : 1;
  side = an[side] > an[2] ? side : 2;
  vec3 r[3];
  for (int i = 0; i < 3; ++i) { /* geom:31-53 */
    if (side == 0) r[i] = V3(clip[i].z, clip[i].y, clip[i].x);
    else if (side == 1) r[i] = V3(clip[i].x, clip[i].z, clip[i].y);
    else r[i] = clip[i];
  }
  const float h = 1.0f / fres; /* geom:56 */
  /* geom:59-61, in pixels */
  float aabb[4] = {std::fmin(std::fmin(r[0].x, r[1].x), r[2].x) - h, std::fmin(std::fmin(r[0].y, r[1].y), r[2].y) - h,
                   std::fmax(std::fmax(r[0].x, r[1].x), r[2].x) + h, std::fmax(std::fmax(r[0].y, r[1].y), r[2].y) + h};
  for (int i = 0; i < 4; ++i) aabb[i] = (aabb[i] * 0.5f + 0.5f) * fres;
  /* geom:64-69 */
  float ax = r[0].x - r[2].x, ay = r[0].y - r[2].y;
  float bx = r[1].x - r[0].x, by = r[1].y - r[0].y;
  Plane pl[3];
  pl[0] = cross_h(ax, ay, 0.0f, r[2].x, r[2].y, 1.0f);
  pl[1] = cross_h(bx, by, 0.0f, r[0].x, r[0].y, 1.0f);
  pl[2] = cross_h(r[2].x - r[1].x, r[2].y - r[1].y, 0.0f, r[1].x, r[1].y, 1.0f);
  float winding = signf(ax * by - bx * ay); /* geom:72 */
  if (winding == 0.0f) return;               /* degenerate in projection: the rasteriser emits nothing */
  for (int i = 0; i < 3; ++i) {
    pl[i].x *= winding; pl[i].y *= winding; pl[i].z *= winding;
    pl[i].z -= h * std::fabs(pl[i].x) + h * std::fabs(pl[i].y); /* geom:78-80 */
  }
  /* geom:96-106: dilated vertices = pairwise intersections, z of the original vertex */
  float vx[3], vy[3], vz[3];
  for (int i = 0; i < 3; ++i) {
    const Plane& p0 = pl[i];
    const Plane& p1 = pl[(i + 1) % 3];
    Plane c = cross_h(p0.x, p0.y, p0.z, p1.x, p1.y, p1.z);
    vx[i] = (c.x / c.z * 0.5f + 0.5f) * fres; /* window x, pixels */
    vy[i] = (c.y / c.z * 0.5f + 0.5f) * fres;
    vz[i] = (r[i].z * 0.5f + 0.5f) * fres;    /* gl_FragCoord.z * VoxelResolution, frag:33 */
  }
  /* affine depth over the dilated triangle: Z(X,Y) = vz0 + gx (X - vx0) + gy (Y - vy0) */
  float e1x = vx[1] - vx[0], e1y = vy[1] - vy[0], e1z = vz[1] - vz[0];
  float e2x = vx[2] - vx[0], e2y = vy[2] - vy[0], e2z = vz[2] - vz[0];
  float det = e1x * e2y - e2x * e1y;
  if (det == 0.0f || det != det) return;
  float gx = (e1z * e2y - e2z * e1y) / det; /* dFdx of voxelPosSwizzled.z, frag:39 */
  float gy = (e1x * e2z - e2x * e1z) / det; /* dFdy */
  float maxChange = std::sqrt(gx * gx + gy * gy) * 1.414f; /* frag:41 */
  int x0 = std::max(0, trunc_to_int(std::floor(aabb[0] - 0.5f)));
  int y0 = std::max(0, trunc_to_int(std::floor(aabb[1] - 0.5f)));
  int x1 = std::min(res - 1, trunc_to_int(std::floor(aabb[2])));
  int y1 = std::min(res - 1, trunc_to_int(std::floor(aabb[3])));
  for (int py = y0; py <= y1; ++py)
    for (int px = x0; px <= x1; ++px) {
      float fx = (float)px + 0.5f, fy = (float)py + 0.5f; /* gl_FragCoord.xy */
      if (fx < aabb[0] || fy < aabb[1] || fx > aabb[2] || fy > aabb[3]) continue; /* frag:24-27 */
      float cx = fx / fres * 2.0f - 1.0f, cy = fy / fres * 2.0f - 1.0f;
      bool inside = true;
      for (int i = 0; i < 3; ++i)
        if ((pl[i].x * cx + pl[i].y * cy) + pl[i].z > 0.0f) inside = false;
      if (!inside) continue;
      float zv = vz[0] + (gx * (fx - vx[0]) + gy * (fy - vy[0])); /* frag:33 */
      if (zv < 0.0f || zv > fres) continue; /* near/far clip */
      int zi = trunc_to_int(zv);            /* frag:34 */
      set_voxel(vol, res, side, px, py, zi);
      if (zi != trunc_to_int(zv - maxChange)) set_voxel(vol, res, side, px, py, zi - 1); /* frag:46-51 */
      if (zi != trunc_to_int(zv + maxChange)) set_voxel(vol, res, side, px, py, zi + 1); /* frag:52-57 */
    }
}

} // namespace

extern "C" void orc_voxelize(const drv_volume_info* vi, uint32_t res, const float* tri_pos, uint32_t num_tris,
                             const float world[16], uint8_t* target) {
  for (uint32_t t = 0; t < num_tris; ++t) voxelize_triangle(vi, (int)res, tri_pos + (size_t)t * 9, world, target);
}

/* voxelblend.comp:16 in UNORM8: old + sign(target - old) * k/255, clamped by
 * the R8 store. Equal to the float evaluation for every (old, target, k)
 * (tests/test_oracle_kat.py checks the 2^24 cases). */
extern "C" void orc_voxel_blend(uint8_t* volume, const uint8_t* target, uint32_t res, float adaption) {
  int k = (int)std::floor(adaption * 255.0f + 0.5f);
  size_t n = (size_t)res * res * res;
  for (size_t i = 0; i < n; ++i) {
    int o = volume[i], t = target[i];
    int s = (t > o) - (t < o);
    volume[i] = (uint8_t)clampi(o + s * k, 0, 255);
  }
}

/* voxelmipmap.comp:11-12: linear fetch at the centre of 8 children = their
 * mean; the R8 store rounds to nearest (ties up: (sum + 4) >> 3). */
extern "C" void orc_voxel_mips(uint8_t* chain, uint32_t res) {
  uint8_t* src = chain;
  while (res > 1) {
    uint32_t h = res / 2;
    uint8_t* dst = src + (size_t)res * res * res;
    for (uint32_t z = 0; z < h; ++z)
      for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < h; ++x) {
          int sum = 0;
          for (int dz = 0; dz < 2; ++dz)
            for (int dy = 0; dy < 2; ++dy)
              for (int dx = 0; dx < 2; ++dx)
                sum += src[(size_t)(2 * x + dx) + (size_t)res * ((size_t)(2 * y + dy) + (size_t)res * (2 * z + dz))];
          dst[(size_t)x + (size_t)h * ((size_t)y + (size_t)h * z)] = (uint8_t)((sum + 4) >> 3);
        }
    src = dst;
    res = h;
  }
}
