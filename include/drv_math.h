/*
 * drv_math.h — host-side maths that feeds the uniform blocks of drv_gi.h.
 *
 * Restates (does not copy) the handful of `ei` (epsilon) functions the
 * reference's packers use: row-major 4x4 matrices acting on column vectors
 * (M * v); `camera`, `perspectiveDX`, LU-with-pivoting `invert`
 * (dependencies/epsilon/include/ei/details/matrix.inl:1707-1725, 1767-1776,
 * 1090-1170) and round-half-even (details/elementary.inl:96-108), followed by
 * the four packers of rendering/renderer.cpp:290-431 and 664-725.
 * C++11, header-only, no dependencies beyond drv_gi.h.
 */
#ifndef DRV_MATH_H
#define DRV_MATH_H

#include <cmath>
#include <cstring>
#include "drv_gi.h"

namespace drv {

constexpr float kPi = 3.14159265358979323846f; // ei::PI

struct Vec3 {
  float x, y, z;
  Vec3() : x(0), y(0), z(0) {}
  Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  explicit Vec3(float s) : x(s), y(s), z(s) {}
  explicit Vec3(const float* p) : x(p[0]), y(p[1]), z(p[2]) {}
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator-(Vec3 a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 operator*(Vec3 a, float s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator/(Vec3 a, float s) { return Vec3(a.x / s, a.y / s, a.z / s); }
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) {
  return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline Vec3 normalize(Vec3 a) { return a / std::sqrt(dot(a, a)); }
inline void store3(float* dst, Vec3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

/* Row-major, m[r*4+c]; acts on column vectors. */
struct Mat4 {
  float m[16];
  float& operator()(int r, int c) { return m[r * 4 + c]; }
  float operator()(int r, int c) const { return m[r * 4 + c]; }
};

inline Mat4 identity4() {
  Mat4 r;
  for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  return r;
}

inline Mat4 mul(const Mat4& a, const Mat4& b) {
  Mat4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = 0.0f;
      for (int k = 0; k < 4; ++k) s += a(i, k) * b(k, j);
      r(i, j) = s;
    }
  return r;
}

/* ei::camera(pos, target, up) = lookAtH(target - pos, up) * translation(-pos);
 * lookAt rows = x, y, z axes with z = normalize(dir), x = normalize(cross(z, up)),
 * y = cross(x, z)   (matrix.inl:1707-1725). */
inline Mat4 camera(Vec3 position, Vec3 target, Vec3 up = Vec3(0.0f, 1.0f, 0.0f)) {
  Vec3 zAxis = normalize(target - position);
  Vec3 xAxis = normalize(cross(zAxis, up));
  Vec3 yAxis = cross(xAxis, zAxis);
  Mat4 look = identity4();
  look(0, 0) = xAxis.x; look(0, 1) = xAxis.y; look(0, 2) = xAxis.z;
  look(1, 0) = yAxis.x; look(1, 1) = yAxis.y; look(1, 2) = yAxis.z;
  look(2, 0) = zAxis.x; look(2, 1) = zAxis.y; look(2, 2) = zAxis.z;
  Mat4 tr = identity4();
  tr(0, 3) = -position.x; tr(1, 3) = -position.y; tr(2, 3) = -position.z;
  return mul(look, tr);
}

/* ei::perspectiveDX(fovY, aspect, n, f) (matrix.inl:1767-1776). The camera
 * passes (far, near) swapped for reversed-Z (camera/camera.hpp:28); lights
 * pass (2*halfAngle, 1, farPlane, nearPlane) (renderer.cpp:688). */
inline Mat4 perspectiveDX(float fovY, float aspect, float n, float f) {
  float h = std::tan(kPi * 0.5f - fovY / 2.0f);
  float w = h / aspect;
  Mat4 r;
  std::memset(r.m, 0, sizeof(r.m));
  r(0, 0) = w;
  r(1, 1) = h;
  r(2, 2) = f / (f - n);
  r(2, 3) = -n * f / (f - n);
  r(3, 2) = 1.0f;
  return r;
}

/* LUP decomposition + solve against the identity (matrix.inl:1090-1170);
 * returns identity for a singular matrix like ei::invert does. */
inline Mat4 invert(const Mat4& a) {
  Mat4 lu = a;
  unsigned p[4] = {0, 1, 2, 3};
  for (unsigned r = 0; r < 3; ++r) {
    float pivot = 0.0f;
    unsigned pr = 0;
    for (unsigned i = r; i < 4; ++i)
      if (std::fabs(lu(i, r)) > pivot) { pivot = std::fabs(lu(i, r)); pr = i; }
    if (pivot == 0.0f) return identity4();
    if (pr != r) {
      unsigned t = p[r]; p[r] = p[pr]; p[pr] = t;
      for (unsigned c = 0; c < 4; ++c) { float tt = lu(r, c); lu(r, c) = lu(pr, c); lu(pr, c) = tt; }
    }
    for (unsigned i = r + 1; i < 4; ++i) {
      lu(i, r) /= lu(r, r);
      for (unsigned j = r + 1; j < 4; ++j) lu(i, j) -= lu(i, r) * lu(r, j);
    }
  }
  if (lu(3, 3) == 0.0f) return identity4();
  Mat4 x;
  for (unsigned n = 0; n < 4; ++n) {
    for (unsigned i = 0; i < 4; ++i) {
      float sum = 0.0f;
      for (unsigned j = 0; j < i; ++j) sum += lu(i, j) * x(j, n);
      x(i, n) = ((p[i] == n) ? 1.0f : 0.0f) - sum;
    }
    for (int i = 3; i >= 0; --i) {
      float sum = 0.0f;
      for (unsigned j = i + 1; j < 4; ++j) sum += lu(i, j) * x(j, n);
      x(i, n) = (x(i, n) - sum) / lu(i, i);
    }
  }
  return x;
}

/* ei::round: nearest integer, ties to even (elementary.inl:96-108). */
inline float roundHalfEven(float v) {
  float r = std::floor(v);
  float f = v - r;
  if (f < 0.5f) return r;
  if (f > 0.5f) return r + 1.0f;
  return (static_cast<long long>(r) & 1) ? r + 1.0f : r;
}

/* ---- the scene-side inputs of the packers ------------------------------ */

/* camera/camera.hpp:15-39 (+ application.cpp:51-52 defaults). */
struct Camera {
  Vec3 position{0.0f, 2.5f, 5.0f};
  Vec3 direction{0.0f, -2.5f, -5.0f}; /* normalised by the packer like Camera::SetDirection */
  Vec3 up{0.0f, 1.0f, 0.0f};
  float hfovDegrees = 60.0f;
  float aspectRatio = 16.0f / 9.0f;
  float nearPlane = 0.1f;
  float farPlane = 1000.0f;
};

/* scene/light.hpp:8-55 + scene/scene.cpp:6-7. */
struct Light {
  Vec3 intensity{10.0f, 10.0f, 10.0f};
  Vec3 position{0.0f, 0.0f, 0.0f};
  Vec3 direction{0.0f, 0.0f, 1.0f};
  float halfAngle = 0.5f;
  unsigned rsmResolution = 1024;
  unsigned rsmReadLod = 4;
  float normalOffsetShadowBias = 0.01f;
  float shadowBias = 0.0001f;
  unsigned indirectShadowComputationLod = 2;
  float nearPlane = 0.1f;
  float farPlane = 10000.0f;
};

/* Renderer::UpdateConstantUBO, renderer.cpp:290-322 (specular fields zeroed:
 * INDIRECT_SPECULAR is out of scope). Note the signs: ShCosLobeFactor1 is
 * uploaded positive and ShCosLobeFactor2n2_p1_n1 negative. */
inline void packConstant(drv_constant* c, int width, int height, int voxelRes, int cavRes,
                         int cavCascades, unsigned maxCaches) {
  std::memset(c, 0, sizeof(*c));
  c->ShCosLobeFactor0 = sqrtf(kPi) / 2.0f;
  c->ShCosLobeFactor1 = sqrtf(kPi / 3.0f);
  c->ShCosLobeFactor2n2_p1_n1 = -sqrtf(15.0f * kPi) / 8.0f;
  c->ShCosLobeFactor20 = sqrtf(5.0f * kPi) / 16.0f;
  c->ShCosLobeFactor2p2 = sqrtf(15.0f * kPi) / 16.0f;
  c->ShEvaFactor0 = 1.0f / (2.0f * sqrtf(kPi));
  c->ShEvaFactor1 = sqrtf(3.0f) / (2.0f * sqrtf(kPi));
  c->ShEvaFactor2n2_p1_n1 = sqrtf(15.0f / (4.0f * kPi));
  c->ShEvaFactor20 = sqrtf(5.0f / (16.0f * kPi));
  c->ShEvaFactor2p2 = sqrtf(15.0f / (16.0f * kPi));
  c->BackbufferResolution[0] = width;
  c->BackbufferResolution[1] = height;
  c->VoxelResolution = voxelRes;
  c->AddressVolumeResolution = cavRes;
  c->NumAddressVolumeCascades = cavCascades;
  c->MaxNumLightCaches = maxCaches;
}

/* The four specular fields of the Constant block: Renderer::AllocateCacheData sizes the environment-map atlas
 * (renderer.cpp:253: 2^ceil(log2(ceil(sqrt(maxCaches)) * perCacheSize)), capped at GL_MAX_TEXTURE_SIZE — 16384 here)
 * and UpdateConstantUBO uploads the derived values (renderer.cpp:316-319). */
inline void packSpecular(drv_constant* c, unsigned maxCaches, unsigned perCacheSize) {
  int size = (int)powf(2.0f, ceilf(log2f(ceilf(sqrtf((float)maxCaches)) * (float)perCacheSize)));
  if (size > 16384) size = 16384;
  c->SpecularEnvmapTotalSize = size;
  c->SpecularEnvmapPerCacheSize_Texel = (int)perCacheSize;
  c->SpecularEnvmapPerCacheSize_Texcoord = (float)perCacheSize / (float)size;
  c->SpecularEnvmapNumCachesPerDimension = size / (int)perCacheSize;
}

/* Renderer::UpdatePerFrameUBO, renderer.cpp:324-344. */
inline void packPerFrame(drv_per_frame* f, const Camera& cam, float passedTime) {
  std::memset(f, 0, sizeof(*f));
  Vec3 dir = normalize(cam.direction);
  Mat4 view = camera(cam.position, cam.position + dir, cam.up);
  Mat4 proj = perspectiveDX(cam.hfovDegrees * (kPi / 180.0f), cam.aspectRatio, cam.farPlane, cam.nearPlane);
  Mat4 viewProj = mul(proj, view);
  Mat4 invView = invert(view);
  Mat4 invViewProj = invert(viewProj);
  std::memcpy(f->Projection, proj.m, 64);
  std::memcpy(f->ViewProjection, viewProj.m, 64);
  std::memcpy(f->InverseView, invView.m, 64);
  std::memcpy(f->InverseViewProjection, invViewProj.m, 64);
  store3(f->CameraPosition, cam.position);
  store3(f->CameraDirection, dir);
  f->PassedTime = passedTime;
}

/* Renderer::UpdateVolumeUBO, renderer.cpp:346-431 ("simple version"). */
inline void packVolumeInfo(drv_volume_info* v, const Camera& cam, Vec3 sceneMin, Vec3 sceneMax,
                           int voxelRes, int cavRes, int cavCascades, const float* cascadeWorldSize,
                           float transitionZoneSize) {
  std::memset(v, 0, sizeof(*v));
  Vec3 vmin = sceneMin - Vec3(0.001f);
  Vec3 vmax = sceneMax + Vec3(0.001f);
  Vec3 extent = vmax - vmin;
  float largest = std::fmax(extent.x, std::fmax(extent.y, extent.z));
  vmax = vmax + (Vec3(largest) - extent);
  store3(v->VolumeWorldMin, vmin);
  v->VoxelSizeInWorld = (vmax.x - vmin.x) / static_cast<float>(voxelRes);
  store3(v->VolumeWorldMax, vmax);
  v->CAVTransitionZoneSize = transitionZoneSize;
  for (int i = 0; i < cavCascades && i < DRV_MAX_CASCADES; ++i) {
    float size = cascadeWorldSize[i];
    float voxel = size / static_cast<float>(cavRes);
    Vec3 q = cam.position / voxel;
    Vec3 snapped = Vec3(roundHalfEven(q.x), roundHalfEven(q.y), roundHalfEven(q.z)) * voxel;
    drv_cav_cascade& c = v->AddressVolumeCascades[i];
    store3(c.Min, snapped - Vec3(size * 0.5f));
    c.WorldVoxelSize = voxel;
    store3(c.Max, snapped + Vec3(size * 0.5f));
    store3(c.DecisionMin, cam.position - Vec3(size * 0.5f) + Vec3(voxel * 1.5f));
    store3(c.DecisionMax, cam.position + Vec3(size * 0.5f) - Vec3(voxel * 1.5f));
  }
}

/* Default cascade sizes of Renderer::SetCAVCascades, renderer.cpp:1181-1187. */
inline void defaultCascadeWorldSizes(float* sizes, int cavCascades, float first = 4.0f) {
  for (int i = 0; i < cavCascades; ++i) sizes[i] = (i == 0) ? first : sizes[i - 1] * 2.0f;
}

/* Renderer::PrepareLights, renderer.cpp:664-725. */
inline void packSpotLight(drv_spot_light* s, const Light& l) {
  std::memset(s, 0, sizeof(*s));
  store3(s->LightIntensity, l.intensity);
  s->ShadowNormalOffset = l.normalOffsetShadowBias;
  s->ShadowBias = l.shadowBias;
  store3(s->LightPosition, l.position);
  store3(s->LightDirection, normalize(l.direction));
  s->LightCosHalfAngle = cosf(l.halfAngle);
  Mat4 view = camera(l.position, l.position + l.direction);
  Mat4 proj = perspectiveDX(l.halfAngle * 2.0f, 1.0f, l.farPlane, l.nearPlane);
  Mat4 viewProj = mul(proj, view);
  Mat4 inv = invert(viewProj);
  std::memcpy(s->LightViewProjection, viewProj.m, 64);
  std::memcpy(s->InverseLightViewProjection, inv.m, 64);
  int pow2 = 1 << static_cast<int>(std::ceil(std::log2(static_cast<double>(l.rsmResolution))));
  s->RSMRenderResolution = pow2;
  int readRes = static_cast<int>(pow2 / std::pow(2.0, static_cast<double>(l.rsmReadLod)));
  s->RSMReadResolution = readRes;
  float clipPlaneWidth = sinf(l.halfAngle) * l.nearPlane * 2.0f;
  float valAreaFactor = clipPlaneWidth * clipPlaneWidth / (l.nearPlane * l.nearPlane * readRes * readRes);
  s->ValAreaFactor = valAreaFactor;
  s->IndirectShadowComputationLod = static_cast<float>(l.indirectShadowComputationLod);
  float block = static_cast<float>(1 << l.indirectShadowComputationLod);
  s->IndirectShadowComputationBlockSize = block;
  s->IndirectShadowComputationSampleInterval = static_cast<int>(block * block);
  s->IndirectShadowComputationSuperValWidth = sqrtf(valAreaFactor) * block;
  s->IndirectShadowSamplingOffset = 0.5f + sqrtf(2.0f) * block / 2.0f;
}

} // namespace drv
#endif /* DRV_MATH_H */
