/*
 * apply.cpp — oracle for stage 5 (per-pixel cache interpolation).
 * TEST INFRASTRUCTURE ONLY. Restates shader/cacheApply.frag:28-195 and
 * lightcache.glsl:109-183 (INDIRECT_SPECULAR / SHOW_ADDRESSVOL_CASCADES off).
 */
#include "oracle.h"
#include "glsl_scalar.h"

using namespace orc;

namespace {

struct ApplyParams {
  const drv_constant* cb;
  const drv_per_frame* pf;
  const drv_volume_info* vi;
  int W, H, R, C;
  bool transitions;
  int order;
  const uint32_t* atlas;
  const uint8_t* entries;
  uint32_t stride, maxCaches;
};

inline int compute_cascade(const ApplyParams& p, vec3 wp) { /* lightcache.glsl:109-122 */
  int c = 0;
  for (; c < p.C - 1; ++c) {
    const drv_cav_cascade& k = p.vi->AddressVolumeCascades[c];
    if (wp.x <= k.DecisionMax[0] && wp.y <= k.DecisionMax[1] && wp.z <= k.DecisionMax[2] &&
        wp.x >= k.DecisionMin[0] && wp.y >= k.DecisionMin[1] && wp.z >= k.DecisionMin[2])
      break;
  }
  return c;
}

inline float cascade_transition(const ApplyParams& p, vec3 wp, int c) { /* lightcache.glsl:125-134 */
  const drv_cav_cascade& k = p.vi->AddressVolumeCascades[c];
  vec3 toMax = V3(k.DecisionMax) - wp;
  vec3 toMin = wp - V3(k.DecisionMin);
  float minDist = std::fmin(std::fmin(std::fmin(toMax.x, toMax.y), toMax.z),
                            std::fmin(std::fmin(toMin.x, toMin.y), toMin.z));
  return saturate(1.0f - minDist / (k.WorldVoxelSize * p.vi->CAVTransitionZoneSize));
}

/* lightcache.glsl:137-183. A missing cache (atlas 0 -> address 0xFFFFFFFF, or
 * an address past the buffer) reads zeros. */
inline vec3 sample_cache_irradiance(const ApplyParams& p, uint32_t address, vec3 n) {
  if (address >= p.maxCaches) return V3(0, 0, 0);
  const float* E = (const float*)(p.entries + (size_t)address * p.stride);
  const drv_constant& k = *p.cb;
  vec3 irr = V3(E[7], E[11], E[15]) * k.ShCosLobeFactor0;
  irr = irr - V3(E + 4) * (k.ShCosLobeFactor1 * n.y);
  irr = irr + V3(E + 8) * (k.ShCosLobeFactor1 * n.z);
  irr = irr - V3(E + 12) * (k.ShCosLobeFactor1 * n.x);
  if (p.order == 2) {
    irr = irr - V3(E + 16) * (k.ShCosLobeFactor2n2_p1_n1 * n.x * n.y);
    irr = irr + V3(E + 20) * (k.ShCosLobeFactor2n2_p1_n1 * n.y * n.z);
    irr = irr + V3(E[19], E[23], E[27]) * (k.ShCosLobeFactor20 * (n.z * n.z * 3.0f - 1.0f));
    irr = irr + V3(E + 24) * (k.ShCosLobeFactor2n2_p1_n1 * n.x * n.z);
    irr = irr + V3(E + 28) * (k.ShCosLobeFactor2p2 * (n.x * n.x - n.y * n.y));
  }
  return V3(std::fmax(irr.x, 0.0f), std::fmax(irr.y, 0.0f), std::fmax(irr.z, 0.0f));
}

/* cacheApply.frag:28-118. */
inline vec3 lighting_from_caches(const ApplyParams& p, vec3 wp, vec3 n, int c, vec3 diffuse) {
  const drv_cav_cascade& k = p.vi->AddressVolumeCascades[c];
  vec3 a = (wp - V3(k.Min)) / k.WorldVoxelSize;
  int bx = trunc_to_int(a.x), by = trunc_to_int(a.y), bz = trunc_to_int(a.z);
  vec3 f = V3(a.x - (float)bx, a.y - (float)by, a.z - (float)bz);
  vec3 g = V3(1.0f - f.x, 1.0f - f.y, 1.0f - f.z);
  static const int off[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {1, 1, 0},
                                {0, 0, 1}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
  float w[8] = {g.x * g.y * g.z, f.x * g.y * g.z, g.x * f.y * g.z, f.x * f.y * g.z,
                g.x * g.y * f.z, f.x * g.y * f.z, g.x * f.y * f.z, f.x * f.y * f.z};
  const int atlasW = p.R * p.C;
  vec3 sum = V3(0, 0, 0);
  for (int i = 0; i < 8; ++i) {
    int x = bx + off[i][0] + p.R * c, y = by + off[i][1], z = bz + off[i][2];
    uint32_t address = 0; /* texelFetch outside the texture returns 0 */
    if (x >= 0 && x < atlasW && y >= 0 && y < p.R && z >= 0 && z < p.R)
      address = p.atlas[(size_t)x + (size_t)atlasW * ((size_t)y + (size_t)p.R * z)];
    address -= 1u;
    vec3 irr = sample_cache_irradiance(p, address, n);
    sum = sum + irr * w[i];
  }
  return sum * diffuse / GLSL_PI;
}

void shade_rows(const ApplyParams& p, const float* depth, const int16_t* normal, const uint8_t* diffuse,
                float* out, int64_t y0, int64_t y1) {
  for (int64_t y = y0; y < y1; ++y)
    for (int x = 0; x < p.W; ++x) {
      size_t t = (size_t)y * p.W + x;
      float* o = out + t * 4;
      o[0] = o[1] = o[2] = o[3] = 0.0f;
      float d = depth[t];
      if (d < 0.00001f) continue; /* :128 discard */
      float px = (float)x + 0.5f, py = (float)y + 0.5f; /* gl_FragCoord.xy */
      float ndc[4] = {px / (float)p.W * 2.0f - 1.0f, py / (float)p.H * 2.0f - 1.0f, d, 1.0f};
      float w4[4];
      mul_row_major(p.pf->InverseViewProjection, ndc, w4);
      vec3 wp = V3(w4[0] / w4[3], w4[1] / w4[3], w4[2] / w4[3]);
      int c = compute_cascade(p, wp);
      vec3 n = unpack_normal16i(normal[t * 2], normal[t * 2 + 1]);
      vec3 albedo = V3(srgb8_to_linear(diffuse[t * 4]), srgb8_to_linear(diffuse[t * 4 + 1]),
                       srgb8_to_linear(diffuse[t * 4 + 2]));
      vec3 color = lighting_from_caches(p, wp, n, c, albedo);
      if (p.transitions) { /* :172-184 */
        float tr = cascade_transition(p, wp, c);
        if (tr > 0.0f && c < p.C - 1) {
          vec3 second = lighting_from_caches(p, wp, n, c + 1, albedo);
          color = V3(mixf(color.x, second.x, tr), mixf(color.y, second.y, tr), mixf(color.z, second.z, tr));
        }
      }
      o[0] = color.x; o[1] = color.y; o[2] = color.z; o[3] = 1.0f;
    }
}

} // namespace

extern "C" void orc_apply_caches(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi,
                                 int transitions, int sh_order, const float* depth, const int16_t* normal,
                                 const uint8_t* diffuse, const uint32_t* atlas, const void* entries,
                                 uint32_t entry_stride, uint32_t max_caches, float* out_rgba, int threads) {
  ApplyParams p;
  p.cb = cb; p.pf = pf; p.vi = vi;
  p.W = cb->BackbufferResolution[0]; p.H = cb->BackbufferResolution[1];
  p.R = cb->AddressVolumeResolution; p.C = cb->NumAddressVolumeCascades;
  p.transitions = transitions != 0; p.order = sh_order;
  p.atlas = atlas; p.entries = (const uint8_t*)entries; p.stride = entry_stride; p.maxCaches = max_caches;
  parallel_for(p.H, threads, [&](int64_t b, int64_t e, int) { shade_rows(p, depth, normal, diffuse, out_rgba, b, e); });
}
