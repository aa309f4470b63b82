/*
 * light.cpp — oracle for stages 2+4 (VPL generation, cone-traced visibility,
 * cache x VPL SH gather) and the RSM mip rule. TEST INFRASTRUCTURE ONLY.
 * Restates shader/cacheLightingRSM.comp:83-374 (INDIRECT_SPECULAR undefined)
 * and shader/downsamplersm.frag:15-33.
 */
#include "oracle.h"
#include "glsl_scalar.h"

using namespace orc;

namespace {

inline int voxel_levels(uint32_t res) {
  int l = 1;
  while (res > 1) { res >>= 1; ++l; }
  return l;
}
inline size_t voxel_level_offset(uint32_t res, int level) {
  size_t off = 0;
  for (int l = 0; l < level; ++l) { off += (size_t)res * res * res; res >>= 1; }
  return off;
}

/* D.0 tri(l): clamp-to-edge trilinear fetch of one level. */
inline float trilinear_level(const uint8_t* lvl, int r, vec3 p) {
  float fx = p.x * (float)r - 0.5f, fy = p.y * (float)r - 0.5f, fz = p.z * (float)r - 0.5f;
  float flx = std::floor(fx), fly = std::floor(fy), flz = std::floor(fz);
  float tx = fx - flx, ty = fy - fly, tz = fz - flz;
  int x0 = trunc_to_int(flx), y0 = trunc_to_int(fly), z0 = trunc_to_int(flz);
  int x1 = clampi(x0 + 1, 0, r - 1), y1 = clampi(y0 + 1, 0, r - 1), z1 = clampi(z0 + 1, 0, r - 1);
  x0 = clampi(x0, 0, r - 1); y0 = clampi(y0, 0, r - 1); z0 = clampi(z0, 0, r - 1);
  auto T = [&](int x, int y, int z) { return (float)lvl[(size_t)x + (size_t)r * ((size_t)y + (size_t)r * z)] / 255.0f; };
  float c00 = mixf(T(x0, y0, z0), T(x1, y0, z0), tx);
  float c10 = mixf(T(x0, y1, z0), T(x1, y1, z0), tx);
  float c01 = mixf(T(x0, y0, z1), T(x1, y0, z1), tx);
  float c11 = mixf(T(x0, y1, z1), T(x1, y1, z1), tx);
  float c0 = mixf(c00, c10, ty);
  float c1 = mixf(c01, c11, ty);
  return mixf(c0, c1, tz);
}

/* D.0 trilinearClampMip3D (sampler renderer.cpp:29-30,916). */
inline float sample_voxel(const uint8_t* chain, uint32_t res, vec3 p, float lod) {
  int L = voxel_levels(res);
  float maxLod = (float)(L - 1);
  if (!(lod > 0.0f)) lod = 0.0f; /* also catches NaN / -inf */
  if (lod > maxLod) lod = maxLod;
  float fl = std::floor(lod);
  int l0 = (int)fl;
  float t = lod - fl;
  int l1 = std::min(l0 + 1, L - 1);
  float a = trilinear_level(chain + voxel_level_offset(res, l0), (int)(res >> l0), p);
  if (t == 0.0f) return a;
  float b = trilinear_level(chain + voxel_level_offset(res, l1), (int)(res >> l1), p);
  return mixf(a, b, t);
}

/* cacheLightingRSM.comp:195-230. */
inline float cone_trace(const drv_volume_info* vi, const uint8_t* chain, uint32_t res, vec3 worldPosition,
                        const drv_shadow_block& blk) {
  float fres = (float)res;
  vec3 voxelPos = (worldPosition - V3(vi->VolumeWorldMin)) / (vi->VoxelSizeInWorld * fres); /* :104 */
  float distToSphereRad = blk.DistToSphereRad;
  vec3 toAverageVal = V3(blk.AverageValPos) - worldPosition;
  float lightDist = length(toAverageVal);
  toAverageVal = toAverageVal / lightDist;
  vec3 dirInVoxel = toAverageVal / fres;
  vec3 currentPosVoxel = voxelPos + dirInVoxel * 2.0f;
  float occlusion = 0.0f;
  float stepSize = 1.0f;
  float dist = 0.0f;
  float goalDist = lightDist / vi->VoxelSizeInWorld - 2.0f;
  float sphereRadiusToStepSize = 2.0f / (1.0f - distToSphereRad);
  for (int s = 0; s < 32; ++s) {
    currentPosVoxel = currentPosVoxel + dirInVoxel * stepSize;
    dist += stepSize;
    float currentSphereRadius = dist * distToSphereRad;
    float newOcclusion = sample_voxel(chain, res, currentPosVoxel, std::log2(currentSphereRadius));
    occlusion += (1.0f - occlusion) * newOcclusion;
    if (dist >= goalDist) break;
    stepSize = std::fmax(1.0f, currentSphereRadius * sphereRadiusToStepSize);
  }
  return saturate(1.0f - occlusion);
}

/* unproject a light-clip-space uv at depth 0 and push it out to distance d
 * along the ray from the light (cacheLightingRSM.comp:154-155, 175-176). */
inline vec3 rsm_world_position(const drv_spot_light* L, float u, float v, float d) {
  float clip[4] = {u * 2.0f - 1.0f, v * 2.0f - 1.0f, 0.0f, 1.0f};
  float w4[4];
  mul_row_major(L->InverseLightViewProjection, clip, w4);
  vec3 lp = V3(L->LightPosition);
  vec3 ws = V3(w4[0] / w4[3], w4[1] / w4[3], w4[2] / w4[3]);
  return lp + normalize(ws - lp) * d;
}

template <typename ACC, int ORDER, bool SHADOW>
void light_range(const drv_constant* cb, const drv_volume_info* vi, const drv_spot_light* lights,
                 uint32_t num_lights, const drv_vpl* const* vpls, const drv_shadow_block* const* blocks,
                 const uint8_t* chain, uint8_t* entries, uint32_t stride, int64_t b, int64_t e) {
  const float f0 = cb->ShEvaFactor0, f1 = cb->ShEvaFactor1, f2 = cb->ShEvaFactor2n2_p1_n1,
              f20 = cb->ShEvaFactor20, f22 = cb->ShEvaFactor2p2;
  const uint32_t vres = (uint32_t)cb->VoxelResolution;
  for (int64_t id = b; id < e; ++id) {
    float* E = (float*)(entries + (size_t)id * stride);
    vec3 worldPosition = V3(E);
    for (uint32_t li = 0; li < num_lights; ++li) {
      const drv_spot_light& L = lights[li];
      const drv_vpl* V = vpls[li];
      ACC sh[9][3];
      for (int i = 0; i < 9; ++i) sh[i][0] = sh[i][1] = sh[i][2] = (ACC)0;
      float shadowing = 1.0f;
      const uint32_t total = (uint32_t)(L.RSMReadResolution * L.RSMReadResolution);
      const uint32_t interval = (uint32_t)L.IndirectShadowComputationSampleInterval;
      for (uint32_t k = 0; k < total; ++k) {
        if (SHADOW && (k % interval) == 0)
          shadowing = cone_trace(vi, chain, vres, worldPosition, blocks[li][k / interval]);
        const drv_vpl& v = V[k];
        vec3 toVal = V3(v.Position) - worldPosition;              /* :249 */
        float lightDistanceSq = dot(toVal, toVal);                /* :252 */
        toVal = toVal * inversesqrt(lightDistanceSq);             /* :253 */
        float fluxToIntensity = saturate(dot(V3(v.Normal), -toVal)); /* :256 */
        fluxToIntensity *= shadowing;                             /* :258 */
        float s = fluxToIntensity / (lightDistanceSq + v.DiscArea); /* :262 */
        float rad[3] = {v.Flux[0] * s, v.Flux[1] * s, v.Flux[2] * s};
        float b1y = f1 * toVal.y, b1z = f1 * toVal.z, b1x = f1 * toVal.x;
        for (int c = 0; c < 3; ++c) {
          sh[0][c] += (ACC)(f0 * rad[c]);   /* SH00    :267 */
          sh[1][c] -= (ACC)(b1y * rad[c]);  /* SH1neg1 :268 */
          sh[2][c] += (ACC)(b1z * rad[c]);  /* SH10    :269 */
          sh[3][c] -= (ACC)(b1x * rad[c]);  /* SH1pos1 :270 */
        }
        if (ORDER == 2) {
          float b2n2 = f2 * toVal.x * toVal.y;                     /* :273 */
          float b2n1 = f2 * toVal.y * toVal.z;                     /* :274 */
          float b20 = f20 * (toVal.z * toVal.z * 3.0f - 1.0f);     /* :275 */
          float b2p1 = f2 * toVal.x * toVal.z;                     /* :276 */
          float b2p2 = f22 * (toVal.x * toVal.x - toVal.y * toVal.y); /* :277 */
          for (int c = 0; c < 3; ++c) {
            sh[4][c] -= (ACC)(b2n2 * rad[c]);
            sh[5][c] += (ACC)(b2n1 * rad[c]);
            sh[6][c] += (ACC)(b20 * rad[c]);
            sh[7][c] += (ACC)(b2p1 * rad[c]);
            sh[8][c] += (ACC)(b2p2 * rad[c]);
          }
        }
      }
      /* :358-373 — entry layout lightcache.glsl:33-57 */
      for (int c = 0; c < 3; ++c) {
        E[4 + c] += (float)sh[1][c];   /* SH1neg1 */
        E[8 + c] += (float)sh[2][c];   /* SH10 */
        E[12 + c] += (float)sh[3][c];  /* SH1pos1 */
      }
      E[7] += (float)sh[0][0]; E[11] += (float)sh[0][1]; E[15] += (float)sh[0][2]; /* SH00_r/g/b */
      if (ORDER == 2) {
        for (int c = 0; c < 3; ++c) {
          E[16 + c] += (float)sh[4][c]; /* SH2neg2 */
          E[20 + c] += (float)sh[5][c]; /* SH2neg1 */
          E[24 + c] += (float)sh[7][c]; /* SH2pos1 */
          E[28 + c] += (float)sh[8][c]; /* SH2pos2 */
        }
        E[19] += (float)sh[6][0]; E[23] += (float)sh[6][1]; E[27] += (float)sh[6][2]; /* SH20_r/g/b */
      }
    }
  }
}

} // namespace

extern "C" float orc_sample_voxel(const uint8_t* chain, uint32_t res, const float p[3], float lod) {
  return sample_voxel(chain, res, V3(p), lod);
}

extern "C" float orc_cone_trace(const drv_volume_info* vi, const uint8_t* chain, uint32_t res,
                                const float cache_pos[3], const drv_shadow_block* block) {
  return cone_trace(vi, chain, res, V3(cache_pos), *block);
}

extern "C" void orc_generate_vpls(const drv_spot_light* L, const uint16_t* flux, const int16_t* normal,
                                  const uint16_t* depth, drv_vpl* out) {
  const uint32_t R = (uint32_t)L->RSMReadResolution;
  for (uint32_t k = 0; k < R * R; ++k) {
    uint32_t x, y;
    morton_decode(k, x, y);                                   /* :142 */
    float u = ((float)x + 0.5f) / (float)R, v = ((float)y + 0.5f) / (float)R; /* :144 */
    size_t t = (size_t)y * R + x;
    drv_vpl o;
    std::memset(&o, 0, sizeof(o));
    o.Flux[0] = half_to_float(flux[t * 4 + 0]);               /* :147 */
    o.Flux[1] = half_to_float(flux[t * 4 + 1]);
    o.Flux[2] = half_to_float(flux[t * 4 + 2]);
    float d = half_to_float(depth[t * 2 + 0]);                /* :150 texel centre => exact texel */
    o.DiscArea = d * d * L->ValAreaFactor;                    /* :151 */
    vec3 p = rsm_world_position(L, u, v, d);                  /* :154-155 */
    o.Position[0] = p.x; o.Position[1] = p.y; o.Position[2] = p.z;
    vec3 n = unpack_normal16i(normal[t * 2 + 0], normal[t * 2 + 1]); /* :158 */
    o.Normal[0] = n.x; o.Normal[1] = n.y; o.Normal[2] = n.z;
    out[k] = o;
  }
}

extern "C" void orc_shadow_blocks(const drv_spot_light* L, const uint16_t* depth_lod, drv_shadow_block* out) {
  const uint32_t R = (uint32_t)L->RSMReadResolution;
  const uint32_t interval = (uint32_t)L->IndirectShadowComputationSampleInterval;
  const int lod = (int)L->IndirectShadowComputationLod;
  const int Rl = (int)(R >> lod);
  for (uint32_t blk = 0; blk < R * R / interval; ++blk) {
    uint32_t x, y;
    morton_decode(blk * interval, x, y);                                       /* :171 */
    float u = ((float)x + L->IndirectShadowSamplingOffset) / (float)R;         /* :172 */
    float v = ((float)y + L->IndirectShadowSamplingOffset) / (float)R;
    /* :174 bilinear, clamp to edge, mip `lod` (D.0 bilinearClamp2D) */
    float fx = u * (float)Rl - 0.5f, fy = v * (float)Rl - 0.5f;
    float flx = std::floor(fx), fly = std::floor(fy);
    float tx = fx - flx, ty = fy - fly;
    int x0 = trunc_to_int(flx), y0 = trunc_to_int(fly);
    int x1 = clampi(x0 + 1, 0, Rl - 1), y1 = clampi(y0 + 1, 0, Rl - 1);
    x0 = clampi(x0, 0, Rl - 1); y0 = clampi(y0, 0, Rl - 1);
    float dd[2];
    for (int c = 0; c < 2; ++c) {
      auto T = [&](int xx, int yy) { return half_to_float(depth_lod[((size_t)yy * Rl + xx) * 2 + c]); };
      dd[c] = mixf(mixf(T(x0, y0), T(x1, y0), tx), mixf(T(x0, y1), T(x1, y1), tx), ty);
    }
    vec3 avg = rsm_world_position(L, u, v, dd[0]);                             /* :175-176 */
    float depthVariance = dd[1] - dd[0] * dd[0];                               /* :181 */
    if (!(depthVariance > 0.0f)) depthVariance = 0.0f;                         /* SURVEY B.7 policy */
    float distToSphereRad = std::fmax(L->IndirectShadowComputationSuperValWidth,
                                      std::sqrt(depthVariance) * 2.0f / dd[0]); /* :192 */
    if (distToSphereRad != distToSphereRad) distToSphereRad = L->IndirectShadowComputationSuperValWidth;
    drv_shadow_block o = {{avg.x, avg.y, avg.z}, distToSphereRad};
    out[blk] = o;
  }
}

extern "C" void orc_light_caches(const drv_constant* cb, const drv_volume_info* vi, const drv_spot_light* lights,
                                 uint32_t num_lights, const drv_vpl* const* vpls,
                                 const drv_shadow_block* const* blocks, const uint8_t* voxel_chain, void* entries,
                                 uint32_t entry_stride, uint32_t first, uint32_t count, int sh_order,
                                 int indirect_shadow, int accumulate_fp64, int threads) {
  uint8_t* E = (uint8_t*)entries;
  parallel_for((int64_t)count, threads, [&](int64_t b, int64_t e, int) {
    b += first; e += first;
#define DRV_GO(ACC, ORD, SH) light_range<ACC, ORD, SH>(cb, vi, lights, num_lights, vpls, blocks, voxel_chain, E, entry_stride, b, e)
    if (accumulate_fp64) {
      if (sh_order == 2) { if (indirect_shadow) DRV_GO(double, 2, true); else DRV_GO(double, 2, false); }
      else               { if (indirect_shadow) DRV_GO(double, 1, true); else DRV_GO(double, 1, false); }
    } else {
      if (sh_order == 2) { if (indirect_shadow) DRV_GO(float, 2, true); else DRV_GO(float, 2, false); }
      else               { if (indirect_shadow) DRV_GO(float, 1, true); else DRV_GO(float, 1, false); }
    }
#undef DRV_GO
  });
}

extern "C" void orc_rsm_downsample(const uint16_t* flux_src, const int16_t* normal_src, const uint16_t* depth_src,
                                   uint32_t res, uint16_t* flux_dst, int16_t* normal_dst, uint16_t* depth_dst) {
  const uint32_t h = res / 2;
  /* rows are independent: all host threads for the large levels (the reference arm of bench.py times this) */
  parallel_for((int64_t)h, h >= 128 ? 0 : 1, [&](int64_t y0, int64_t y1, int) {
  for (uint32_t y = (uint32_t)y0; y < (uint32_t)y1; ++y)
    for (uint32_t x = 0; x < h; ++x) {
      /* textureGather at the shared corner of the 2x2 footprint: texels
       * (2x,2y+1) (2x+1,2y+1) (2x+1,2y) (2x,2y) in .xyzw order */
      size_t t[4] = {(size_t)(2 * y + 1) * res + 2 * x, (size_t)(2 * y + 1) * res + 2 * x + 1,
                     (size_t)(2 * y) * res + 2 * x + 1, (size_t)(2 * y) * res + 2 * x};
      size_t o = (size_t)y * h + x;
      for (int c = 0; c < 3; ++c) { /* :17-22 flux = sum of 4 */
        float s = half_to_float(flux_src[t[0] * 4 + c]) + half_to_float(flux_src[t[1] * 4 + c]) +
                  half_to_float(flux_src[t[2] * 4 + c]) + half_to_float(flux_src[t[3] * 4 + c]);
        flux_dst[o * 4 + c] = float_to_half(s);
      }
      flux_dst[o * 4 + 3] = 0;
      vec3 n = unpack_normal16i(normal_src[t[0] * 2], normal_src[t[0] * 2 + 1]); /* :24-30 */
      n = n + unpack_normal16i(normal_src[t[1] * 2], normal_src[t[1] * 2 + 1]);
      n = n + unpack_normal16i(normal_src[t[2] * 2], normal_src[t[2] * 2 + 1]);
      n = n + unpack_normal16i(normal_src[t[3] * 2], normal_src[t[3] * 2 + 1]);
      pack_normal16i(normalize(n), normal_dst[o * 2], normal_dst[o * 2 + 1]);
      for (int c = 0; c < 2; ++c) { /* :32 bilinear at the footprint centre = mean of 4 */
        float a = mixf(half_to_float(depth_src[t[3] * 2 + c]), half_to_float(depth_src[t[2] * 2 + c]), 0.5f);
        float b = mixf(half_to_float(depth_src[t[0] * 2 + c]), half_to_float(depth_src[t[1] * 2 + c]), 0.5f);
        depth_dst[o * 2 + c] = float_to_half(mixf(a, b, 0.5f));
      }
    }
  });
}

extern "C" float orc_half_to_float(uint16_t h) { return half_to_float(h); }
extern "C" uint16_t orc_float_to_half(float f) { return float_to_half(f); }
extern "C" void orc_pack_normal16i(const float n[3], int16_t out[2]) { pack_normal16i(V3(n), out[0], out[1]); }
extern "C" void orc_unpack_normal16i(const int16_t in[2], float n[3]) {
  vec3 v = unpack_normal16i(in[0], in[1]);
  n[0] = v.x; n[1] = v.y; n[2] = v.z;
}
extern "C" float orc_srgb8_to_linear(uint8_t v) { return srgb8_to_linear(v); }
extern "C" uint32_t orc_morton_decode_x(uint32_t k) { uint32_t x, y; morton_decode(k, x, y); return x; }
extern "C" uint32_t orc_morton_decode_y(uint32_t k) { uint32_t x, y; morton_decode(k, x, y); return y; }
