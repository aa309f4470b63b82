/*
 * specular.cpp — oracle for SURVEY.md 8(f) row f4: indirect specular via per-cache hemispherical environment maps.
 * TEST INFRASTRUCTURE ONLY (see oracle.h). Restates, with INDIRECT_SPECULAR and DIRECT_SPECULAR_MAP_WRITE defined,
 *   shader/cacheLightingRSM.comp:87-99, 127-129, 239-243, 281-335 (+ lightcache.glsl:97-106, utils.glsl:103-201,
 *   lightingfunctions.glsl:33-36), shader/specularenvmap_mipmap.frag + specularenvmap.vert:11-22,
 *   shader/specularenvmap_fillholes.frag:10-52, shader/cacheApply.frag:28-118, 136-157 (+ lightingfunctions.glsl:25-48)
 * and the host side Renderer::PrepareSpecularEnvmaps (renderer.cpp:994-1045).
 *
 * The atlas is R11F_G11F_B10F (renderer.cpp:282): every imageStore rounds to 6 / 5 mantissa bits, so the result
 * depends on the ORDER of the per-texel read-add-write. The reference's order is its execution order: one dispatch
 * per light; work groups of 64 caches; inside a group every invocation works through the 64 VPLs staged between two
 * barriers before the next tile is staged. A map's texel coordinate can be SPECULARENVMAP_PERCACHESIZE (the +0.5 of
 * :312 gives size + 1 bins), which lands in the next cache's map — a data race on the GPU; here, as in the pin
 * (oracle/_ref run on one thread, invocations in index order), it is the serialisation light -> group -> tile ->
 * invocation -> VPL. Invocations past the cache count (padding of the last group) read zeros from the entry buffer
 * (robust access) and still store into their own maps (:323-327 is not guarded), as in the shader.
 */
#include "oracle.h"
#include "glsl_scalar.h"
#include "../include/drv_r11g11b10.h"

using namespace orc;

namespace {

struct Mat3 { vec3 c[3]; }; /* columns */
inline vec3 mul_vm(vec3 v, const Mat3& m) { return V3(dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2])); } /* v * M */

/* lightcache.glsl:97-106 */
inline Mat3 local_view_space(const float* cameraPosition, vec3 wp) {
  Mat3 m;
  m.c[2] = normalize(V3(cameraPosition) - wp);
  m.c[0] = normalize(V3(m.c[2].z, 0.0f, -m.c[2].x));
  m.c[1] = cross(m.c[2], m.c[0]);
  return m;
}

/* utils.glsl:103-171, HEMIPROJECTION_LAMBERT_CONCENTRICQUAD, the "better perf" branch */
inline void hemispherical_projection(vec3 d, float& px, float& py) {
  float r = std::sqrt(1.0f - d.z) * 0.5f;
  float phi = std::atan2(d.y, d.x) * (4.0f / GLSL_PI);
  if (phi < -1.0f) phi += 8.0f;
  float x, y;
  if (phi < 3.0f) {
    if (phi < 1.0f) { x = r; y = phi * r; }
    else { x = -(phi - 2.0f) * r; y = r; }
  } else {
    if (phi < 5.0f) { x = -r; y = -(phi - 4.0f) * r; }
    else { x = (phi - 6.0f) * r; y = -r; }
  }
  px = x + 0.5f;
  py = y + 0.5f;
}

inline void image_add(uint32_t* atlas, int size, int x, int y, vec3 v) { /* :324-327 */
  if (x < 0 || y < 0 || x >= size || y >= size) return; /* out-of-range image accesses are dropped */
  uint32_t* t = atlas + (size_t)y * size + x;
  float r, g, b;
  drv_unpack_r11g11b10(*t, &r, &g, &b);
  *t = drv_pack_r11g11b10(r + v.x, g + v.y, b + v.z);
}

inline vec3 fetch_rgb(const uint32_t* lvl, int size, int x, int y) {
  float r, g, b;
  drv_unpack_r11g11b10(lvl[(size_t)y * size + x], &r, &g, &b);
  return V3(r, g, b);
}
/* linear / clamp-to-edge sample of one level (D.0 bilinearClamp2D) */
inline vec3 bilinear_rgb(const uint32_t* lvl, int size, float u, float v) {
  float fx = u * (float)size - 0.5f, fy = v * (float)size - 0.5f;
  float flx = std::floor(fx), fly = std::floor(fy), tx = fx - flx, ty = fy - fly;
  int x0 = trunc_to_int(flx), y0 = trunc_to_int(fly);
  int x1 = clampi(x0 + 1, 0, size - 1), y1 = clampi(y0 + 1, 0, size - 1);
  x0 = clampi(x0, 0, size - 1); y0 = clampi(y0, 0, size - 1);
  vec3 a = fetch_rgb(lvl, size, x0, y0) * (1.0f - tx) + fetch_rgb(lvl, size, x1, y0) * tx;
  vec3 b = fetch_rgb(lvl, size, x0, y1) * (1.0f - tx) + fetch_rgb(lvl, size, x1, y1) * tx;
  return a * (1.0f - ty) + b * ty;
}
inline const uint32_t* level_ptr(const uint32_t* mips, int total, int l) {
  const uint32_t* p = mips;
  for (int i = 0, r = total; i < l; ++i, r >>= 1) p += (size_t)r * r;
  return p;
}
/* textureLod(CacheSpecularEnvmap, uv, lod): linear, mip-linear, clamp (renderer.cpp:1070-1071) */
inline vec3 sample_envmap(const uint32_t* mips, int total, int levels, float u, float v, float lod) {
  float maxLod = (float)(levels - 1);
  if (!(lod > 0.0f)) lod = 0.0f;
  if (lod > maxLod) lod = maxLod;
  float fl = std::floor(lod), f = lod - fl;
  int l0 = (int)fl, l1 = std::min(l0 + 1, levels - 1);
  vec3 a = bilinear_rgb(level_ptr(mips, total, l0), total >> l0, u, v);
  if (f == 0.0f) return a;
  vec3 b = bilinear_rgb(level_ptr(mips, total, l1), total >> l1, u, v);
  return a * (1.0f - f) + b * f;
}

} // namespace

/* The cone trace of light.cpp */
extern "C" float orc_cone_trace(const drv_volume_info* vi, const uint8_t* chain, uint32_t res, const float cache_pos[3],
                                const drv_shadow_block* block);

/* cacheLightingRSM.comp with INDIRECT_SPECULAR + DIRECT_SPECULAR_MAP_WRITE: SH (with the :241 early-out) AND the
 * environment-map atlas (cleared here, renderer.cpp:903). `count` entries; the last group is padded to 64. */
extern "C" void orc_light_caches_specular(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi,
                                          const drv_spot_light* lights, uint32_t num_lights, const drv_vpl* const* vpls,
                                          const drv_shadow_block* const* blocks, const uint8_t* voxel_chain, void* entries,
                                          uint32_t entry_stride, uint32_t count, int sh_order, int indirect_shadow,
                                          uint32_t* atlas) {
  const int total = cb->SpecularEnvmapTotalSize, S = cb->SpecularEnvmapPerCacheSize_Texel;
  const int perDim = cb->SpecularEnvmapNumCachesPerDimension;
  std::memset(atlas, 0, (size_t)total * total * sizeof(uint32_t));
  const float f0 = cb->ShEvaFactor0, f1 = cb->ShEvaFactor1, f2 = cb->ShEvaFactor2n2_p1_n1, f20 = cb->ShEvaFactor20,
              f22 = cb->ShEvaFactor2p2;
  const float baseExp = (float)S * (float)S - 1.0f;                  /* utils.glsl:172-183 */
  const float baseNorm = (baseExp + 8.0f) / (8.0f * GLSL_PI);         /* lightingfunctions.glsl:33-36 */
  const uint32_t vres = (uint32_t)cb->VoxelResolution;
  const uint32_t groups = (count + 63u) / 64u;
  struct Inv { vec3 wp; Mat3 view; float sh[9][3]; float shadowing; int ox, oy; };
  std::vector<Inv> inv(64);
  uint8_t* E = (uint8_t*)entries;
  for (uint32_t li = 0; li < num_lights; ++li) {
    const drv_spot_light& L = lights[li];
    const drv_vpl* V = vpls[li];
    const uint32_t totalVpls = (uint32_t)(L.RSMReadResolution * L.RSMReadResolution);
    const uint32_t interval = (uint32_t)L.IndirectShadowComputationSampleInterval;
    for (uint32_t g = 0; g < groups; ++g) {
      for (uint32_t f = 0; f < 64; ++f) {
        const uint32_t id = g * 64 + f;
        Inv& I = inv[f];
        I.wp = id < count ? V3((const float*)(E + (size_t)id * entry_stride)) : V3(0, 0, 0); /* :83, robust access */
        I.view = local_view_space(pf->CameraPosition, I.wp);                                   /* :127-129 */
        for (int i = 0; i < 9; ++i) I.sh[i][0] = I.sh[i][1] = I.sh[i][2] = 0.0f;
        I.shadowing = 1.0f;
        I.ox = (int)(id % (uint32_t)perDim) * S;                                               /* :87-88 */
        I.oy = (int)(id / (uint32_t)perDim) * S;
      }
      for (uint32_t tile = 0; tile < totalVpls; tile += 64) {
        for (uint32_t f = 0; f < 64; ++f) {
          Inv& I = inv[f];
          for (uint32_t k = tile; k < tile + 64 && k < totalVpls; ++k) {
            if (indirect_shadow && (k % interval) == 0) {
              const float p[3] = {I.wp.x, I.wp.y, I.wp.z};
              I.shadowing = orc_cone_trace(vi, voxel_chain, vres, p, &blocks[li][k / interval]);
            }
            const drv_vpl& v = V[k];
            if (v.Flux[0] + v.Flux[1] + v.Flux[2] < 0.001f) continue;                         /* :241 */
            vec3 toVal = V3(v.Position) - I.wp;
            float d2 = dot(toVal, toVal);
            toVal = toVal * inversesqrt(d2);
            float fluxToIntensity = saturate(dot(V3(v.Normal), -toVal));
            fluxToIntensity *= I.shadowing;
            float s = fluxToIntensity / (d2 + v.DiscArea);
            vec3 rad = V3(v.Flux[0] * s, v.Flux[1] * s, v.Flux[2] * s);
            float b1y = f1 * toVal.y, b1z = f1 * toVal.z, b1x = f1 * toVal.x;
            const float radc[3] = {rad.x, rad.y, rad.z};
            for (int c = 0; c < 3; ++c) {
              I.sh[0][c] += f0 * radc[c];
              I.sh[1][c] -= b1y * radc[c];
              I.sh[2][c] += b1z * radc[c];
              I.sh[3][c] -= b1x * radc[c];
            }
            if (sh_order == 2) {
              float b2n2 = f2 * toVal.x * toVal.y, b2n1 = f2 * toVal.y * toVal.z;
              float b20 = f20 * (toVal.z * toVal.z * 3.0f - 1.0f), b2p1 = f2 * toVal.x * toVal.z;
              float b2p2 = f22 * (toVal.x * toVal.x - toVal.y * toVal.y);
              for (int c = 0; c < 3; ++c) {
                I.sh[4][c] -= b2n2 * radc[c];
                I.sh[5][c] += b2n1 * radc[c];
                I.sh[6][c] += b20 * radc[c];
                I.sh[7][c] += b2p1 * radc[c];
                I.sh[8][c] += b2p2 * radc[c];
              }
            }
            vec3 local = mul_vm(toVal, I.view);                                                /* :282 */
            vec3 h = normalize(local + V3(0.0f, 0.0f, 1.0f));                                  /* :311 */
            float px, py;
            hemispherical_projection(h, px, py);
            int tx = trunc_to_int(px * (float)S + 0.5f), ty = trunc_to_int(py * (float)S + 0.5f); /* :312 */
            vec3 out = rad * saturate(dot(h, local));                                          /* :319 */
            out = out * baseNorm;                                                              /* :323 */
            image_add(atlas, total, tx + I.ox, ty + I.oy, out);                                /* :324-327 */
          }
        }
      }
      for (uint32_t f = 0; f < 64; ++f) { /* :341-374 */
        const uint32_t id = g * 64 + f;
        if (id >= count) continue;
        float* e = (float*)(E + (size_t)id * entry_stride);
        const Inv& I = inv[f];
        for (int c = 0; c < 3; ++c) { e[4 + c] += I.sh[1][c]; e[8 + c] += I.sh[2][c]; e[12 + c] += I.sh[3][c]; }
        e[7] += I.sh[0][0]; e[11] += I.sh[0][1]; e[15] += I.sh[0][2];
        if (sh_order == 2) {
          for (int c = 0; c < 3; ++c) {
            e[16 + c] += I.sh[4][c]; e[20 + c] += I.sh[5][c]; e[24 + c] += I.sh[7][c]; e[28 + c] += I.sh[8][c];
          }
          e[19] += I.sh[6][0]; e[23] += I.sh[6][1]; e[27] += I.sh[6][2];
        }
      }
    }
  }
}

/* specularenvmap_mipmap.frag through specularenvmap.vert (renderer.cpp:1008-1020). mips: level 0 first. */
extern "C" void orc_specular_mips(const drv_constant* cb, uint32_t count, uint32_t* mips) {
  const int total = cb->SpecularEnvmapTotalSize, per = cb->SpecularEnvmapPerCacheSize_Texel;
  const int perDim = cb->SpecularEnvmapNumCachesPerDimension;
  const float numUsedRows = std::ceil((float)count / (float)perDim);
  const float rowPercentage = numUsedRows / (float)perDim;
  uint32_t* src = mips;
  int r = total;
  for (int p = per; p > 1; p >>= 1, r >>= 1) {
    uint32_t* dst = src + (size_t)r * r;
    const int h = r / 2;
    for (int y = 0; y < h; ++y) {
      if (((float)y + 0.5f) / (float)h > rowPercentage) continue;
      for (int x = 0; x < h; ++x) {
        vec3 c = bilinear_rgb(src, r, ((float)x + 0.5f) / (float)h, ((float)y + 0.5f) / (float)h);
        dst[(size_t)y * h + x] = drv_pack_r11g11b10(c.x, c.y, c.z);
      }
    }
    src = dst;
  }
}

/* specularenvmap_fillholes.frag (renderer.cpp:1022-1044): push level i down into the empty texels of level i-1. */
extern "C" void orc_specular_fill_holes(const drv_constant* cb, uint32_t count, uint32_t max_level, uint32_t* mips) {
  const int total = cb->SpecularEnvmapTotalSize;
  const int perDim = cb->SpecularEnvmapNumCachesPerDimension;
  const float numUsedRows = std::ceil((float)count / (float)perDim);
  const float rowPercentage = numUsedRows / (float)perDim;
  for (int i = (int)max_level; i > 0; --i) {
    const int rs = total >> i, rd = total >> (i - 1);
    uint32_t* src = const_cast<uint32_t*>(level_ptr(mips, total, i));
    uint32_t* dst = const_cast<uint32_t*>(level_ptr(mips, total, i - 1));
    for (int y = 0; y < rs; ++y) {
      if (((float)y + 0.5f) / (float)rs > rowPercentage) continue;
      for (int x = 0; x < rs; ++x) {
        float u = ((float)x + 0.5f) / (float)rs, v = ((float)y + 0.5f) / (float)rs;
        int sx = trunc_to_int(u * (float)rs), sy = trunc_to_int(v * (float)rs);              /* :15 */
        if (sx < 0 || sy < 0 || sx >= rs || sy >= rs) continue;
        vec3 sc = fetch_rgb(src, rs, sx, sy);
        if (sc.x + sc.y + sc.z < 0.0001f) continue;                                            /* :19-20 */
        int dx[4] = {sx * 2, sx * 2 + 1, sx * 2, sx * 2 + 1}, dy[4] = {sy * 2, sy * 2, sy * 2 + 1, sy * 2 + 1};
        vec3 dc[4];
        vec3 sum = V3(0, 0, 0);
        for (int k = 0; k < 4; ++k) {
          dc[k] = (dx[k] < rd && dy[k] < rd) ? fetch_rgb(dst, rd, dx[k], dy[k]) : V3(0, 0, 0);
          if (dc[k].x + dc[k].y + dc[k].z == 0.0f) dc[k] = sc;                                 /* :36-37 */
          sum = sum + dc[k];
        }
        sum = V3(sum.x + 0.00001f, sum.y + 0.00001f, sum.z + 0.00001f);                        /* :43 */
        vec3 norm = (sc * 4.0f) / sum;                                                         /* :44 */
        for (int k = 0; k < 4; ++k) {
          if (dx[k] >= rd || dy[k] >= rd) continue;
          vec3 o = dc[k] * norm;
          dst[(size_t)dy[k] * rd + dx[k]] = drv_pack_r11g11b10(o.x, o.y, o.z);
        }
      }
    }
  }
}

namespace {
/* the cascade helpers of apply.cpp, restated (lightcache.glsl:109-134) */
inline int cascade_of(const drv_volume_info* vi, int C, vec3 wp) {
  int c = 0;
  for (; c < C - 1; ++c) {
    const drv_cav_cascade& k = vi->AddressVolumeCascades[c];
    if (wp.x <= k.DecisionMax[0] && wp.y <= k.DecisionMax[1] && wp.z <= k.DecisionMax[2] && wp.x >= k.DecisionMin[0] &&
        wp.y >= k.DecisionMin[1] && wp.z >= k.DecisionMin[2])
      break;
  }
  return c;
}
inline float transition_of(const drv_volume_info* vi, vec3 wp, int c) {
  const drv_cav_cascade& k = vi->AddressVolumeCascades[c];
  vec3 toMax = V3(k.DecisionMax) - wp, toMin = wp - V3(k.DecisionMin);
  float minDist = std::fmin(std::fmin(std::fmin(toMax.x, toMax.y), toMax.z), std::fmin(std::fmin(toMin.x, toMin.y), toMin.z));
  return saturate(1.0f - minDist / (k.WorldVoxelSize * vi->CAVTransitionZoneSize));
}
struct SpecApply {
  const drv_constant* cb; const drv_volume_info* vi;
  int R, C, order;
  const uint32_t* atlas; const uint8_t* entries; uint32_t stride, maxCaches;
  const uint32_t* spec; int total, levels;
};
inline vec3 irradiance_of(const SpecApply& p, uint32_t address, vec3 n) { /* lightcache.glsl:137-183 */
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
/* cacheApply.frag:28-118 with INDIRECT_SPECULAR */
inline vec3 lighting(const SpecApply& p, vec3 wp, vec3 n, int c, vec3 diffuse, float lx, float ly, vec3 specColor, float lod) {
  const drv_cav_cascade& k = p.vi->AddressVolumeCascades[c];
  vec3 a = (wp - V3(k.Min)) / k.WorldVoxelSize;
  int bx = trunc_to_int(a.x), by = trunc_to_int(a.y), bz = trunc_to_int(a.z);
  vec3 f = V3(a.x - (float)bx, a.y - (float)by, a.z - (float)bz), g = V3(1.0f - f.x, 1.0f - f.y, 1.0f - f.z);
  static const int off[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {1, 1, 0}, {0, 0, 1}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
  float w[8] = {g.x * g.y * g.z, f.x * g.y * g.z, g.x * f.y * g.z, f.x * f.y * g.z,
                g.x * g.y * f.z, f.x * g.y * f.z, g.x * f.y * f.z, f.x * f.y * f.z};
  const int atlasW = p.R * p.C;
  const int perDim = p.cb->SpecularEnvmapNumCachesPerDimension;
  vec3 sum = V3(0, 0, 0), spec = V3(0, 0, 0);
  for (int i = 0; i < 8; ++i) {
    int x = bx + off[i][0] + p.R * c, y = by + off[i][1], z = bz + off[i][2];
    uint32_t address = 0;
    if (x >= 0 && x < atlasW && y >= 0 && y < p.R && z >= 0 && z < p.R)
      address = p.atlas[(size_t)x + (size_t)atlasW * ((size_t)y + (size_t)p.R * z)];
    address -= 1u;
    /* :102-106 — also for a missing cache (address 0xFFFFFFFF): the lookup lands wherever that leads, clamped */
    float ox = (float)(address % (uint32_t)perDim), oy = (float)(address / (uint32_t)perDim);
    float u = (lx + ox) * p.cb->SpecularEnvmapPerCacheSize_Texcoord, v = (ly + oy) * p.cb->SpecularEnvmapPerCacheSize_Texcoord;
    spec = spec + sample_envmap(p.spec, p.total, p.levels, u, v, lod) * w[i];
    sum = sum + irradiance_of(p, address, n) * w[i];
  }
  return sum * diffuse / GLSL_PI + spec * specColor; /* :116 */
}
} // namespace

/* cacheApply.frag with INDIRECT_SPECULAR (renderer.cpp:1047-1079, envmap bound with the linear-clamp sampler). */
extern "C" void orc_apply_caches_specular(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi,
                                          int transitions, int sh_order, const float* depth, const int16_t* normal,
                                          const uint8_t* diffuse, const uint8_t* roughness_metallic, const uint32_t* atlas,
                                          const void* entries, uint32_t entry_stride, uint32_t max_caches,
                                          const uint32_t* specular_mips, float* out_rgba, int threads) {
  SpecApply p;
  p.cb = cb; p.vi = vi; p.R = cb->AddressVolumeResolution; p.C = cb->NumAddressVolumeCascades; p.order = sh_order;
  p.atlas = atlas; p.entries = (const uint8_t*)entries; p.stride = entry_stride; p.maxCaches = max_caches;
  p.spec = specular_mips; p.total = cb->SpecularEnvmapTotalSize;
  p.levels = 0;
  for (int per = cb->SpecularEnvmapPerCacheSize_Texel; per >= 1; per >>= 1) p.levels++;
  const int W = cb->BackbufferResolution[0], H = cb->BackbufferResolution[1];
  parallel_for(H, threads, [&](int64_t y0, int64_t y1, int) {
    for (int64_t y = y0; y < y1; ++y)
      for (int x = 0; x < W; ++x) {
        size_t t = (size_t)y * W + x;
        float* o = out_rgba + t * 4;
        o[0] = o[1] = o[2] = o[3] = 0.0f;
        float d = depth[t];
        if (d < 0.00001f) continue;
        float ndc[4] = {((float)x + 0.5f) / (float)W * 2.0f - 1.0f, ((float)y + 0.5f) / (float)H * 2.0f - 1.0f, d, 1.0f};
        float w4[4];
        mul_row_major(pf->InverseViewProjection, ndc, w4);
        vec3 wp = V3(w4[0] / w4[3], w4[1] / w4[3], w4[2] / w4[3]);
        int c = cascade_of(vi, p.C, wp);
        vec3 n = unpack_normal16i(normal[t * 2], normal[t * 2 + 1]);
        vec3 base = V3(srgb8_to_linear(diffuse[t * 4]), srgb8_to_linear(diffuse[t * 4 + 1]), srgb8_to_linear(diffuse[t * 4 + 2]));
        float roughness = (float)roughness_metallic[t * 2] / 255.0f, metallic = (float)roughness_metallic[t * 2 + 1] / 255.0f;
        /* lightingfunctions.glsl:37-48 */
        vec3 diffuseColor = V3(mixf(base.x, 0.02f, metallic), mixf(base.y, 0.02f, metallic), mixf(base.z, 0.02f, metallic));
        vec3 specularColor = V3(mixf(0.04f, base.x, metallic), mixf(0.04f, base.y, metallic), mixf(0.04f, base.z, metallic));
        float rsq = roughness * roughness;
        float blinnExponent = 2.0f / (rsq * rsq + 0.0005f);                                   /* lightingfunctions.glsl:25-31 */
        float S = (float)cb->SpecularEnvmapPerCacheSize_Texel;
        float lod = std::fmax(0.0f, std::log2(S * S / (1.0f + blinnExponent)) * 0.5f);       /* utils.glsl:185-193 */
        float maxHalf = 0.5f / (std::pow(2.0f, -std::ceil(lod)) * S);                         /* :148 */
        Mat3 view = local_view_space(pf->CameraPosition, wp);
        vec3 vn = mul_vm(n, view);                                                            /* :150 */
        float lx, ly;
        hemispherical_projection(vn, lx, ly);
        lx = std::fmin(std::fmax(lx, maxHalf), 1.0f - maxHalf);                               /* :152 */
        ly = std::fmin(std::fmax(ly, maxHalf), 1.0f - maxHalf);
        vec3 color = lighting(p, wp, n, c, diffuseColor, lx, ly, specularColor, lod);
        if (transitions) {
          float tr = transition_of(vi, wp, c);
          if (tr > 0.0f && c < p.C - 1) {
            vec3 second = lighting(p, wp, n, c + 1, diffuseColor, lx, ly, specularColor, lod);
            color = V3(mixf(color.x, second.x, tr), mixf(color.y, second.y, tr), mixf(color.z, second.z, tr));
          }
        }
        o[0] = color.x; o[1] = color.y; o[2] = color.z; o[3] = 1.0f;
      }
  });
}
