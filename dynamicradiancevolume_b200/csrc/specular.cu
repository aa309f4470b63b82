// specular.cu — SURVEY.md 8(f) row f4: indirect specular through per-cache hemispherical environment maps
// (the reference's INDIRECT_SPECULAR + DIRECT_SPECULAR_MAP_WRITE build: shader/cacheLightingRSM.comp:87-99, 127-129,
// 239-243, 281-335; shader/specularenvmap_mipmap.frag, specularenvmap_fillholes.frag, specularenvmap.vert;
// shader/cacheApply.frag:28-118, 136-157; Renderer::PrepareSpecularEnvmaps, renderer.cpp:994-1045). Off by default
// (drv_config.indirect_specular), as in the reference.
//
// The atlas is R11F_G11F_B10F: every read-add-write of a texel rounds to 6 / 5 mantissa bits, so a texel's value
// depends on the ORDER of its contributions. The reference's order is one thread per cache walking the VPLs in
// Morton order, light after light. Here a WARP owns a cache: its lanes evaluate 32 consecutive VPLs at once, and
// contributions that hit the same texel are applied in lane (= VPL) order, round by round (__match_any_sync), into
// the cache's private (S+1)^2 patch in shared memory — the reference's order for every texel of the cache's own map,
// with 32x the parallelism. Texel coordinate S (the +0.5 of :312 yields S+1 bins) belongs to the neighbouring
// cache's map: the reference races there; here the patch keeps such spills apart and a merge pass adds them to the
// neighbour's texel in a fixed order (own, left, lower, diagonal).
#include "ctx.h"
#include "device_math.cuh"
#include "../../include/drv_r11g11b10.h"

using namespace drvk;

namespace {

constexpr int kSpecWarps = 8;          // caches per block
constexpr int kSpecTile = kSpecWarps * 32; // VPLs staged per tile
constexpr int kMaxS = 16;

struct SpecLight {
  const float4* vpls;      // live list: (pos, area) (normal, shadow-block index) (flux, -)
  const uint32_t* live;
  uint32_t num_vpls;
  uint32_t block_offset;   // first row of this light in the visibility table
};
struct SpecParams {
  SpecLight lights[DRV_MAX_LIGHTS];
  uint32_t num_lights;
  const uint8_t* entries;
  uint32_t entry_stride;
  const drv_cache_counter* counter;
  float cam[3];
  int S, per_dim, total;
  float base_norm;
  const float* shadow_table; // null: unshadowed
  uint32_t shadow_stride;
  uint32_t* patches;         // [cache][(S+1)^2]
};

struct M3 { F3 c0, c1, c2; };
__device__ __forceinline__ F3 ex_normalize(F3 v) {
  const float inv = ex_rsqrt(ex_dot3(v.x, v.y, v.z, v.x, v.y, v.z));
  F3 r = {ex_mul(v.x, inv), ex_mul(v.y, inv), ex_mul(v.z, inv)};
  return r;
}
// lightcache.glsl:97-106
__device__ __forceinline__ M3 local_view_space(const float* cam, F3 wp) {
  M3 m;
  F3 d = {ex_sub(cam[0], wp.x), ex_sub(cam[1], wp.y), ex_sub(cam[2], wp.z)};
  m.c2 = ex_normalize(d);
  F3 x = {m.c2.z, 0.0f, -m.c2.x};
  m.c0 = ex_normalize(x);
  // cross(c2, c0)
  m.c1.x = ex_sub(ex_mul(m.c2.y, m.c0.z), ex_mul(m.c2.z, m.c0.y));
  m.c1.y = ex_sub(ex_mul(m.c2.z, m.c0.x), ex_mul(m.c2.x, m.c0.z));
  m.c1.z = ex_sub(ex_mul(m.c2.x, m.c0.y), ex_mul(m.c2.y, m.c0.x));
  return m;
}
__device__ __forceinline__ F3 mul_vm(F3 v, const M3& m) { // v * M: component j = dot(v, column j)
  F3 r = {ex_dot3(v.x, v.y, v.z, m.c0.x, m.c0.y, m.c0.z), ex_dot3(v.x, v.y, v.z, m.c1.x, m.c1.y, m.c1.z),
          ex_dot3(v.x, v.y, v.z, m.c2.x, m.c2.y, m.c2.z)};
  return r;
}
// utils.glsl:103-171 (HEMIPROJECTION_LAMBERT_CONCENTRICQUAD, "better perf" branch)
__device__ __forceinline__ void hemispherical_projection(F3 d, float& px, float& py) {
  const float r = ex_mul(ex_sqrt(ex_sub(1.0f, d.z)), 0.5f);
  float phi = ex_mul(atan2f(d.y, d.x), 4.0f / DRV_GLSL_PI);
  if (phi < -1.0f) phi = ex_add(phi, 8.0f);
  float x, y;
  if (phi < 3.0f) {
    if (phi < 1.0f) { x = r; y = ex_mul(phi, r); }
    else { x = ex_mul(-ex_sub(phi, 2.0f), r); y = r; }
  } else {
    if (phi < 5.0f) { x = -r; y = ex_mul(-ex_sub(phi, 4.0f), r); }
    else { x = ex_mul(ex_sub(phi, 6.0f), r); y = -r; }
  }
  px = ex_add(x, 0.5f);
  py = ex_add(y, 0.5f);
}

template <bool SHADOW>
__global__ void __launch_bounds__(kSpecWarps * 32) specular_accumulate_kernel(const __grid_constant__ SpecParams p) {
  __shared__ uint32_t s_patch[kSpecWarps][(kMaxS + 1) * (kMaxS + 1)];
  __shared__ float4 s_vpl[kSpecTile * 3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t count = (uint32_t)max(p.counter->TotalLightCacheCount, 0);
  if (blockIdx.x * kSpecWarps >= count) return; // block-uniform
  const uint32_t id = blockIdx.x * kSpecWarps + warp;
  const bool have = id < count;
  const int S1 = p.S + 1, texels = S1 * S1;
  for (int i = lane; i < texels; i += 32) s_patch[warp][i] = 0u;
  F3 wp = {0.f, 0.f, 0.f};
  if (have) {
    const float4 e = *reinterpret_cast<const float4*>(p.entries + (size_t)id * p.entry_stride);
    wp.x = e.x; wp.y = e.y; wp.z = e.z;
  }
  const M3 view = local_view_space(p.cam, wp);
  for (uint32_t li = 0; li < p.num_lights; ++li) {
    const SpecLight& L = p.lights[li];
    const uint32_t n = min(__ldg(L.live), L.num_vpls);
    for (uint32_t base = 0; base < n; base += kSpecTile) {
      __syncthreads(); // previous tile consumed
      {
        const uint32_t v = base + threadIdx.x;
        const bool ok = v < n;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        s_vpl[threadIdx.x * 3 + 0] = ok ? __ldg(L.vpls + (size_t)v * 3) : z;
        s_vpl[threadIdx.x * 3 + 1] = ok ? __ldg(L.vpls + (size_t)v * 3 + 1) : z;
        s_vpl[threadIdx.x * 3 + 2] = ok ? __ldg(L.vpls + (size_t)v * 3 + 2) : z;
      }
      __syncthreads();
      if (!have) continue; // warp-uniform; the barriers above are outside
      const uint32_t m = min((uint32_t)kSpecTile, n - base);
      for (uint32_t c0 = 0; c0 < m; c0 += 32) {
        const uint32_t i = c0 + lane;
        int key = -1;
        float vr = 0.f, vg = 0.f, vb = 0.f;
        if (i < m) {
          const float4 va = s_vpl[i * 3], vn = s_vpl[i * 3 + 1], vf = s_vpl[i * 3 + 2];
          // cacheLightingRSM.comp:249-262, decision-maths operators throughout: the 6-bit mantissa of the target
          // turns every last-bit difference of a contribution into a visible one
          F3 t = {ex_sub(va.x, wp.x), ex_sub(va.y, wp.y), ex_sub(va.z, wp.z)};
          const float d2 = ex_dot3(t.x, t.y, t.z, t.x, t.y, t.z);
          const float inv = ex_rsqrt(d2);
          t.x = ex_mul(t.x, inv); t.y = ex_mul(t.y, inv); t.z = ex_mul(t.z, inv);
          float fti = saturatef(ex_dot3(vn.x, vn.y, vn.z, -t.x, -t.y, -t.z));
          if (SHADOW) {
            const uint32_t blk = __float_as_uint(vn.w);
            fti = ex_mul(fti, __ldg(p.shadow_table + (size_t)(L.block_offset + blk) * p.shadow_stride + id));
          }
          const float s = ex_div(fti, ex_add(d2, va.w));
          const F3 rad = {ex_mul(vf.x, s), ex_mul(vf.y, s), ex_mul(vf.z, s)};
          const F3 local = mul_vm(t, view);                                                     // :282
          F3 hsum = {local.x, local.y, ex_add(local.z, 1.0f)};
          const F3 h = ex_normalize(hsum);                                                      // :311
          float px, py;
          hemispherical_projection(h, px, py);
          const int tx = ex_trunc(ex_add(ex_mul(px, (float)p.S), 0.5f));                        // :312
          const int ty = ex_trunc(ex_add(ex_mul(py, (float)p.S), 0.5f));
          const float w = saturatef(ex_dot3(h.x, h.y, h.z, local.x, local.y, local.z));                  // :319
          vr = ex_mul(ex_mul(rad.x, w), p.base_norm);                                           // :323
          vg = ex_mul(ex_mul(rad.y, w), p.base_norm);
          vb = ex_mul(ex_mul(rad.z, w), p.base_norm);
          if (tx >= 0 && ty >= 0 && tx <= p.S && ty <= p.S) key = ty * S1 + tx; // anything else is outside the atlas side this cache can reach
        }
        // contributions to the same texel in lane (= VPL) order
        const uint32_t same = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(same & ((1u << lane) - 1u));
        int rounds = key >= 0 ? rank + 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, o));
        for (int r = 0; r < rounds; ++r) {
          if (key >= 0 && rank == r) {                                                          // :324-327
            float cr, cg, cb;
            drv_unpack_r11g11b10(s_patch[warp][key], &cr, &cg, &cb);
            s_patch[warp][key] = drv_pack_r11g11b10(ex_add(cr, vr), ex_add(cg, vg), ex_add(cb, vb));
          }
          __syncwarp();
        }
      }
    }
  }
  if (have) {
    uint32_t* dst = p.patches + (size_t)id * texels;
    for (int i = lane; i < texels; i += 32) dst[i] = s_patch[warp][i];
  }
}

__device__ __forceinline__ uint32_t r11_add(uint32_t a, uint32_t b) {
  float ar, ag, ab, br, bg, bb;
  drv_unpack_r11g11b10(a, &ar, &ag, &ab);
  drv_unpack_r11g11b10(b, &br, &bg, &bb);
  return drv_pack_r11g11b10(ex_add(ar, br), ex_add(ag, bg), ex_add(ab, bb));
}

// Level 0 of the atlas from the patches: also the per-frame ClearToZero of renderer.cpp:903 (unused maps store 0).
__global__ void specular_merge_kernel(const uint32_t* __restrict__ patches, const drv_cache_counter* __restrict__ counter,
                                      int S, int per_dim, int total, uint32_t* __restrict__ atlas) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y;
  if (X >= total) return;
  const uint32_t count = (uint32_t)max(counter->TotalLightCacheCount, 0);
  const int cx = X / S, cy = Y / S, lx = X - cx * S, ly = Y - cy * S, S1 = S + 1, texels = S1 * S1;
  const uint32_t id = (uint32_t)(cy * per_dim + cx);
  uint32_t v = id < count ? __ldg(patches + (size_t)id * texels + ly * S1 + lx) : 0u;
  if (lx == 0 && cx > 0 && id - 1u < count) v = r11_add(v, __ldg(patches + (size_t)(id - 1u) * texels + ly * S1 + S));
  if (ly == 0 && cy > 0 && id - (uint32_t)per_dim < count)
    v = r11_add(v, __ldg(patches + (size_t)(id - (uint32_t)per_dim) * texels + S * S1 + lx));
  if (lx == 0 && ly == 0 && cx > 0 && cy > 0 && id - (uint32_t)per_dim - 1u < count)
    v = r11_add(v, __ldg(patches + (size_t)(id - (uint32_t)per_dim - 1u) * texels + S * S1 + S));
  atlas[(size_t)Y * total + X] = v;
}

__device__ __forceinline__ F3 fetch_rgb(const uint32_t* __restrict__ lvl, int size, int x, int y) {
  F3 r;
  drv_unpack_r11g11b10(__ldg(lvl + (size_t)y * size + x), &r.x, &r.y, &r.z);
  return r;
}
// linear, clamp-to-edge sample of one level, SURVEY D.0 (mix(a, b, t) = a * (1 - t) + b * t)
__device__ __forceinline__ F3 bilinear_rgb(const uint32_t* __restrict__ lvl, int size, float u, float v) {
  const float fx = ex_sub(ex_mul(u, (float)size), 0.5f), fy = ex_sub(ex_mul(v, (float)size), 0.5f);
  const float flx = floorf(fx), fly = floorf(fy), tx = ex_sub(fx, flx), ty = ex_sub(fy, fly);
  int x0 = ex_trunc(flx), y0 = ex_trunc(fly);
  const int x1 = clampi(x0 + 1, 0, size - 1), y1 = clampi(y0 + 1, 0, size - 1);
  x0 = clampi(x0, 0, size - 1); y0 = clampi(y0, 0, size - 1);
  const F3 a00 = fetch_rgb(lvl, size, x0, y0), a10 = fetch_rgb(lvl, size, x1, y0), a01 = fetch_rgb(lvl, size, x0, y1),
           a11 = fetch_rgb(lvl, size, x1, y1);
  F3 r;
  r.x = ex_mix(ex_mix(a00.x, a10.x, tx), ex_mix(a01.x, a11.x, tx), ty);
  r.y = ex_mix(ex_mix(a00.y, a10.y, tx), ex_mix(a01.y, a11.y, tx), ty);
  r.z = ex_mix(ex_mix(a00.z, a10.z, tx), ex_mix(a01.z, a11.z, tx), ty);
  return r;
}

// specularenvmap_mipmap.frag through specularenvmap.vert: dst (h x h) from src (2h x 2h); rows outside the shifted
// triangle are left at zero.
__global__ void specular_mip_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int h,
                                    const drv_cache_counter* __restrict__ counter, int per_dim) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= h) return;
  const uint32_t count = (uint32_t)max(counter->TotalLightCacheCount, 0);
  const float rowPct = ex_div(ceilf(ex_div((float)count, (float)per_dim)), (float)per_dim);
  const float v = ex_div((float)y + 0.5f, (float)h);
  uint32_t o = 0u;
  if (!(v > rowPct)) {
    const F3 c = bilinear_rgb(src, 2 * h, ex_div((float)x + 0.5f, (float)h), v);
    o = drv_pack_r11g11b10(c.x, c.y, c.z);
  }
  dst[(size_t)y * h + x] = o;
}

// specularenvmap_fillholes.frag:10-52: one thread per source texel pushes into its 2x2 destination block.
__global__ void specular_fill_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int rs,
                                     const drv_cache_counter* __restrict__ counter, int per_dim) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= rs) return;
  const uint32_t count = (uint32_t)max(counter->TotalLightCacheCount, 0);
  const float rowPct = ex_div(ceilf(ex_div((float)count, (float)per_dim)), (float)per_dim);
  if (ex_div((float)y + 0.5f, (float)rs) > rowPct) return;
  const F3 sc = fetch_rgb(src, rs, x, y);
  if (ex_add(ex_add(sc.x, sc.y), sc.z) < 0.0001f) return;
  const int rd = rs * 2;
  F3 dc[4];
  F3 sum = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    dc[k] = fetch_rgb(dst, rd, 2 * x + (k & 1), 2 * y + (k >> 1));
    if (ex_add(ex_add(dc[k].x, dc[k].y), dc[k].z) == 0.0f) dc[k] = sc;
    sum.x = ex_add(sum.x, dc[k].x); sum.y = ex_add(sum.y, dc[k].y); sum.z = ex_add(sum.z, dc[k].z);
  }
  sum.x = ex_add(sum.x, 0.00001f); sum.y = ex_add(sum.y, 0.00001f); sum.z = ex_add(sum.z, 0.00001f);
  const F3 nrm = {ex_div(ex_mul(sc.x, 4.0f), sum.x), ex_div(ex_mul(sc.y, 4.0f), sum.y), ex_div(ex_mul(sc.z, 4.0f), sum.z)};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    dst[(size_t)(2 * y + (k >> 1)) * rd + 2 * x + (k & 1)] =
        drv_pack_r11g11b10(ex_mul(dc[k].x, nrm.x), ex_mul(dc[k].y, nrm.y), ex_mul(dc[k].z, nrm.z));
}

// ---------------------------------------------------------------------------------------------- apply
struct SpecApplyParams {
  int W, H, R, C, transitions, order;
  float zone;
  float ivp[16];
  float cam[3];
  drv_cav_cascade casc[DRV_MAX_CASCADES];
  float g0, g1, g2, g20, g22;
  uint32_t max_caches;
  int S, per_dim, total, levels;
  float per_texcoord;
  uint32_t level_offset[8];
};

__device__ __forceinline__ F3 sample_envmap(const SpecApplyParams& p, const uint32_t* __restrict__ mips, float u, float v,
                                            float lod) {
  const float maxLod = (float)(p.levels - 1);
  if (!(lod > 0.0f)) lod = 0.0f;
  if (lod > maxLod) lod = maxLod;
  const float fl = floorf(lod), f = ex_sub(lod, fl);
  const int l0 = (int)fl, l1 = min(l0 + 1, p.levels - 1);
  F3 a = bilinear_rgb(mips + p.level_offset[l0], p.total >> l0, u, v);
  if (f == 0.0f) return a;
  const F3 b = bilinear_rgb(mips + p.level_offset[l1], p.total >> l1, u, v);
  a.x = ex_mix(a.x, b.x, f); a.y = ex_mix(a.y, b.y, f); a.z = ex_mix(a.z, b.z, f);
  return a;
}

__device__ __forceinline__ F3 irradiance_of(const SpecApplyParams& p, const uint8_t* __restrict__ entries, uint32_t stride,
                                            uint32_t address, F3 n) { // lightcache.glsl:137-183
  F3 z = {0.f, 0.f, 0.f};
  if (address >= p.max_caches) return z;
  const float4* e = reinterpret_cast<const float4*>(entries + (size_t)address * stride);
  const float4 q1 = __ldg(e + 1), q2 = __ldg(e + 2), q3 = __ldg(e + 3);
  float ir = q1.w * p.g0, ig = q2.w * p.g0, ib = q3.w * p.g0;
  const float b1y = p.g1 * n.y, b1z = p.g1 * n.z, b1x = p.g1 * n.x;
  ir -= q1.x * b1y; ig -= q1.y * b1y; ib -= q1.z * b1y;
  ir += q2.x * b1z; ig += q2.y * b1z; ib += q2.z * b1z;
  ir -= q3.x * b1x; ig -= q3.y * b1x; ib -= q3.z * b1x;
  if (p.order == 2) {
    const float4 q4 = __ldg(e + 4), q5 = __ldg(e + 5), q6 = __ldg(e + 6), q7 = __ldg(e + 7);
    const float b2xy = p.g2 * n.x * n.y, b2yz = p.g2 * n.y * n.z, b20 = p.g20 * (n.z * n.z * 3.0f - 1.0f);
    const float b2xz = p.g2 * n.x * n.z, b2dd = p.g22 * (n.x * n.x - n.y * n.y);
    ir -= q4.x * b2xy; ig -= q4.y * b2xy; ib -= q4.z * b2xy;
    ir += q5.x * b2yz; ig += q5.y * b2yz; ib += q5.z * b2yz;
    ir += q4.w * b20;  ig += q5.w * b20;  ib += q6.w * b20;
    ir += q6.x * b2xz; ig += q6.y * b2xz; ib += q6.z * b2xz;
    ir += q7.x * b2dd; ig += q7.y * b2dd; ib += q7.z * b2dd;
  }
  F3 r = {fmaxf(ir, 0.0f), fmaxf(ig, 0.0f), fmaxf(ib, 0.0f)};
  return r;
}

// cacheApply.frag:28-118 with INDIRECT_SPECULAR
__device__ __forceinline__ F3 lighting(const SpecApplyParams& p, const uint32_t* __restrict__ atlas,
                                       const uint8_t* __restrict__ entries, uint32_t stride, const uint32_t* __restrict__ mips,
                                       F3 wp, F3 n, int c, F3 diffuse, float lx, float ly, F3 specColor, float lod) {
  const drv_cav_cascade& k = p.casc[c];
  const float ax = ex_div(ex_sub(wp.x, k.Min[0]), k.WorldVoxelSize), ay = ex_div(ex_sub(wp.y, k.Min[1]), k.WorldVoxelSize),
              az = ex_div(ex_sub(wp.z, k.Min[2]), k.WorldVoxelSize);
  const int bx = ex_trunc(ax), by = ex_trunc(ay), bz = ex_trunc(az);
  const float fx = ax - (float)bx, fy = ay - (float)by, fz = az - (float)bz, gx = 1.0f - fx, gy = 1.0f - fy, gz = 1.0f - fz;
  const float w[8] = {gx * gy * gz, fx * gy * gz, gx * fy * gz, fx * fy * gz, gx * gy * fz, fx * gy * fz, gx * fy * fz, fx * fy * fz};
  const int atlasW = p.R * p.C;
  F3 sum = {0.f, 0.f, 0.f}, spec = {0.f, 0.f, 0.f};
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    const int x = bx + (i & 1) + p.R * c, y = by + ((i >> 1) & 1), z = bz + (i >> 2);
    uint32_t address = 0u;
    if (x >= 0 && x < atlasW && y >= 0 && y < p.R && z >= 0 && z < p.R) address = __ldg(atlas + (size_t)x + (size_t)atlasW * ((size_t)y + (size_t)p.R * z));
    address -= 1u;
    // :102-106 (the lookup of a missing cache, address 0xFFFFFFFF, lands wherever that leads: clamped to the edge)
    const float ox = (float)(address % (uint32_t)p.per_dim), oy = (float)(address / (uint32_t)p.per_dim);
    const float u = ex_mul(ex_add(lx, ox), p.per_texcoord), v = ex_mul(ex_add(ly, oy), p.per_texcoord);
    const F3 sv = sample_envmap(p, mips, u, v, lod);
    spec.x += sv.x * w[i]; spec.y += sv.y * w[i]; spec.z += sv.z * w[i];
    const F3 ir = irradiance_of(p, entries, stride, address, n);
    sum.x += ir.x * w[i]; sum.y += ir.y * w[i]; sum.z += ir.z * w[i];
  }
  const float inv_pi = 1.0f / DRV_GLSL_PI;
  F3 r = {sum.x * diffuse.x * inv_pi + spec.x * specColor.x, sum.y * diffuse.y * inv_pi + spec.y * specColor.y,
          sum.z * diffuse.z * inv_pi + spec.z * specColor.z}; // :116
  return r;
}

__device__ __forceinline__ int cascade_of(const SpecApplyParams& p, F3 wp) {
  int c = 0;
  for (; c < p.C - 1; ++c) {
    const drv_cav_cascade& k = p.casc[c];
    if (wp.x <= k.DecisionMax[0] && wp.y <= k.DecisionMax[1] && wp.z <= k.DecisionMax[2] && wp.x >= k.DecisionMin[0] &&
        wp.y >= k.DecisionMin[1] && wp.z >= k.DecisionMin[2])
      break;
  }
  return c;
}

__global__ void __launch_bounds__(256) apply_specular_kernel(const __grid_constant__ SpecApplyParams p, const float* __restrict__ depth,
                                                             const int* __restrict__ normal, const uchar4* __restrict__ diffuse,
                                                             const uchar2* __restrict__ rough_metal, const uint32_t* __restrict__ atlas,
                                                             const uint8_t* __restrict__ entries, uint32_t stride,
                                                             const uint32_t* __restrict__ mips, const float* __restrict__ srgb_lut,
                                                             const float* __restrict__ ndc_xy, void* __restrict__ out, int format,
                                                             int y_begin, int y_end) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = y_begin + blockIdx.y * 8 + threadIdx.y;
  if (x >= p.W || y >= y_end) return;
  const uint32_t t = (uint32_t)y * p.W + x;
  const float d = __ldg(depth + t);
  float r = 0.f, g = 0.f, b = 0.f;
  const bool discard = d < 0.00001f;
  if (!discard) {
    const F3 wp = ex_unproject(p.ivp, __ldg(ndc_xy + x), __ldg(ndc_xy + p.W + y), d);
    const int c = cascade_of(p, wp);
    const int pn = __ldg(normal + t);
    const F3 n = unpack_normal16i((int)(short)(pn & 0xffff), (int)(short)((uint32_t)pn >> 16));
    const uchar4 dc = __ldg(diffuse + t);
    const F3 base = {__ldg(srgb_lut + dc.x), __ldg(srgb_lut + dc.y), __ldg(srgb_lut + dc.z)};
    const uchar2 rm = __ldg(rough_metal + t);
    const float roughness = (float)rm.x / 255.0f, metallic = (float)rm.y / 255.0f;
    // lightingfunctions.glsl:25-48
    const F3 diffuseColor = {ex_mix(base.x, 0.02f, metallic), ex_mix(base.y, 0.02f, metallic), ex_mix(base.z, 0.02f, metallic)};
    const F3 specColor = {ex_mix(0.04f, base.x, metallic), ex_mix(0.04f, base.y, metallic), ex_mix(0.04f, base.z, metallic)};
    const float rsq = ex_mul(roughness, roughness);
    const float blinn = ex_div(2.0f, ex_add(ex_mul(rsq, rsq), 0.0005f));
    const float Sf = (float)p.S;
    const float lod = fmaxf(0.0f, ex_mul(log2f(ex_div(ex_mul(Sf, Sf), ex_add(1.0f, blinn))), 0.5f)); // utils.glsl:185-193
    const float maxHalf = ex_div(0.5f, ex_mul(exp2f(-ceilf(lod)), Sf));                              // :148
    const M3 view = local_view_space(p.cam, wp);
    const F3 vn = mul_vm(n, view);                                                                   // :150
    float lx, ly;
    hemispherical_projection(vn, lx, ly);
    lx = fminf(fmaxf(lx, maxHalf), ex_sub(1.0f, maxHalf));                                           // :152
    ly = fminf(fmaxf(ly, maxHalf), ex_sub(1.0f, maxHalf));
    F3 col = lighting(p, atlas, entries, stride, mips, wp, n, c, diffuseColor, lx, ly, specColor, lod);
    if (p.transitions && c < p.C - 1) {
      const drv_cav_cascade& k = p.casc[c];
      const float ax = ex_sub(k.DecisionMax[0], wp.x), ay = ex_sub(k.DecisionMax[1], wp.y), az = ex_sub(k.DecisionMax[2], wp.z);
      const float bx = ex_sub(wp.x, k.DecisionMin[0]), by = ex_sub(wp.y, k.DecisionMin[1]), bz = ex_sub(wp.z, k.DecisionMin[2]);
      const float minDist = fminf(fminf(fminf(ax, ay), az), fminf(fminf(bx, by), bz));
      const float tr = saturatef(ex_sub(1.0f, ex_div(minDist, ex_mul(k.WorldVoxelSize, p.zone))));
      if (tr > 0.0f) {
        const F3 c2 = lighting(p, atlas, entries, stride, mips, wp, n, c + 1, diffuseColor, lx, ly, specColor, lod);
        col.x = ex_mix(col.x, c2.x, tr); col.y = ex_mix(col.y, c2.y, tr); col.z = ex_mix(col.z, c2.z, tr);
      }
    }
    r = col.x; g = col.y; b = col.z;
  }
  if (format == DRV_HDR_RGBA32F_WRITE) {
    reinterpret_cast<float4*>(out)[t] = discard ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(r, g, b, 1.0f);
  } else if (format == DRV_HDR_RGBA16F_WRITE) {
    __half2 nrg = __floats2half2_rn(r, g), nba = __floats2half2_rn(b, 0.0f);
    uint2 nw;
    nw.x = *reinterpret_cast<uint32_t*>(&nrg);
    nw.y = *reinterpret_cast<uint32_t*>(&nba);
    reinterpret_cast<uint2*>(out)[t] = nw;
  } else if (!discard) {
    uint2* o = reinterpret_cast<uint2*>(out) + t;
    uint2 old = *o;
    float2 rg = __half22float2(*reinterpret_cast<__half2*>(&old.x)), ba = __half22float2(*reinterpret_cast<__half2*>(&old.y));
    __half2 nrg = __floats2half2_rn(rg.x + r, rg.y + g), nba = __floats2half2_rn(ba.x + b, ba.y);
    uint2 nw;
    nw.x = *reinterpret_cast<uint32_t*>(&nrg);
    nw.y = *reinterpret_cast<uint32_t*>(&nba);
    *o = nw;
  }
}

} // namespace

// ------------------------------------------------------------------------------------------------ host side
drv_status drv_impl_specular_light(drv_ctx* ctx) {
  if (!ctx->cfg.indirect_specular) return DRV_OK;
  if (!ctx->have_constant || !ctx->have_per_frame) return ctx->fail(DRV_ERR_NOT_BOUND, "indirect specular: uniform blocks not set");
  const int S = ctx->constant.SpecularEnvmapPerCacheSize_Texel, total = ctx->constant.SpecularEnvmapTotalSize,
            per_dim = ctx->constant.SpecularEnvmapNumCachesPerDimension;
  if (S != (int)ctx->spec_S || total != (int)ctx->spec_total || per_dim * S != total)
    return ctx->fail(DRV_ERR_INVALID, "indirect specular: the Constant block's specular fields disagree with the context "
                                      "(drv_pack_specular with max_cache_count and the configured per-cache size)");
  if (ctx->shard_world > 1) return ctx->fail(DRV_ERR_INVALID, "indirect specular is not sharded: use one GPU");
  const bool shadow = ctx->cfg.indirect_shadow != 0;
  if (shadow && ctx->shadow_chunks != 1)
    return ctx->fail(DRV_ERR_CAPACITY, "indirect specular with indirect shadows needs max_cache_count <= 262144 (one visibility-table chunk)");
  SpecParams p;
  memset(&p, 0, sizeof(p));
  p.num_lights = ctx->num_lights;
  for (uint32_t l = 0; l < ctx->num_lights; ++l) {
    LightState& L = ctx->lights[l];
    p.lights[l].vpls = (const float4*)L.vpls_live;
    p.lights[l].live = ctx->live_counts + l;
    p.lights[l].num_vpls = L.num_vpls;
    p.lights[l].block_offset = ctx->shadow_block_offset[l];
  }
  p.entries = ctx->entries;
  p.entry_stride = ctx->entry_stride;
  p.counter = ctx->counter;
  memcpy(p.cam, ctx->per_frame.CameraPosition, sizeof(p.cam));
  p.S = S; p.per_dim = per_dim; p.total = total;
  const float base_exp = (float)S * (float)S - 1.0f;      // utils.glsl:172-183
  p.base_norm = (base_exp + 8.0f) / (8.0f * DRV_GLSL_PI); // lightingfunctions.glsl:33-36
  p.shadow_table = shadow ? ctx->shadow_table : nullptr;
  p.shadow_stride = ctx->shadow_stride;
  p.patches = ctx->spec_patches;
  const uint32_t blocks = (ctx->cfg.max_cache_count + kSpecWarps - 1) / kSpecWarps; // blocks past the live count exit at once
  if (shadow) specular_accumulate_kernel<true><<<blocks, kSpecWarps * 32, 0, ctx->stream>>>(p);
  else specular_accumulate_kernel<false><<<blocks, kSpecWarps * 32, 0, ctx->stream>>>(p);
  DRV_LAUNCH_CHECK();
  dim3 grid((total + 255) / 256, total);
  specular_merge_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->spec_patches, ctx->counter, S, per_dim, total, ctx->spec_mips);
  DRV_LAUNCH_CHECK();
  return DRV_OK;
}

drv_status drv_impl_prepare_specular(drv_ctx* ctx) {
  if (!ctx->cfg.indirect_specular) return DRV_OK;
  const int total = (int)ctx->spec_total, per_dim = ctx->constant.SpecularEnvmapNumCachesPerDimension;
  for (uint32_t l = 1; l < ctx->spec_levels; ++l) { // renderer.cpp:1008-1020
    const int h = total >> l;
    dim3 grid((h + 127) / 128, h);
    specular_mip_kernel<<<grid, 128, 0, ctx->stream>>>(ctx->spec_mips + ctx->spec_level_offset[l - 1],
                                                       ctx->spec_mips + ctx->spec_level_offset[l], h, ctx->counter, per_dim);
    DRV_LAUNCH_CHECK();
  }
  for (int i = (int)ctx->cfg.specular_fill_holes_level; i > 0; --i) { // renderer.cpp:1032-1042
    const int rs = total >> i;
    dim3 grid((rs + 127) / 128, rs);
    specular_fill_kernel<<<grid, 128, 0, ctx->stream>>>(ctx->spec_mips + ctx->spec_level_offset[i],
                                                        ctx->spec_mips + ctx->spec_level_offset[i - 1], rs, ctx->counter, per_dim);
    DRV_LAUNCH_CHECK();
  }
  return DRV_OK;
}

drv_status drv_impl_apply_specular(drv_ctx* ctx, void* out, uint32_t format, uint32_t y_begin, uint32_t y_end, bool timed,
                                   const float* srgb_lut_dev) {
  if (!ctx->gb_rough_metal) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_apply_caches: indirect specular needs drv_bind_gbuffer_material");
  SpecApplyParams p;
  memset(&p, 0, sizeof(p));
  p.W = ctx->constant.BackbufferResolution[0];
  p.H = ctx->constant.BackbufferResolution[1];
  p.R = ctx->constant.AddressVolumeResolution;
  p.C = ctx->constant.NumAddressVolumeCascades;
  p.transitions = ctx->cfg.cascade_transitions ? 1 : 0;
  p.order = (int)ctx->cfg.sh_order;
  p.zone = ctx->volume.CAVTransitionZoneSize;
  memcpy(p.ivp, ctx->per_frame.InverseViewProjection, sizeof(p.ivp));
  memcpy(p.cam, ctx->per_frame.CameraPosition, sizeof(p.cam));
  memcpy(p.casc, ctx->volume.AddressVolumeCascades, sizeof(p.casc));
  p.g0 = ctx->constant.ShCosLobeFactor0; p.g1 = ctx->constant.ShCosLobeFactor1; p.g2 = ctx->constant.ShCosLobeFactor2n2_p1_n1;
  p.g20 = ctx->constant.ShCosLobeFactor20; p.g22 = ctx->constant.ShCosLobeFactor2p2;
  p.max_caches = ctx->cfg.max_cache_count;
  p.S = (int)ctx->spec_S; p.per_dim = ctx->constant.SpecularEnvmapNumCachesPerDimension; p.total = (int)ctx->spec_total;
  p.levels = (int)ctx->spec_levels;
  p.per_texcoord = ctx->constant.SpecularEnvmapPerCacheSize_Texcoord;
  for (uint32_t l = 0; l < ctx->spec_levels && l < 8; ++l) p.level_offset[l] = ctx->spec_level_offset[l];
  if (y_end > (uint32_t)p.H) y_end = (uint32_t)p.H;
  if (y_begin >= y_end) return DRV_OK;
  if (timed) ctx->stage_begin(DRV_STAGE_APPLY_CACHES);
  dim3 block(32, 8), grid((p.W + 31) / 32, (y_end - y_begin + 7) / 8);
  apply_specular_kernel<<<grid, block, 0, ctx->stream>>>(p, ctx->gb_depth, (const int*)ctx->gb_normal, (const uchar4*)ctx->gb_diffuse,
                                                         (const uchar2*)ctx->gb_rough_metal, ctx->atlas, ctx->entries,
                                                         ctx->entry_stride, ctx->spec_mips, srgb_lut_dev, ctx->ndc_xy, out,
                                                         (int)format, (int)y_begin, (int)y_end);
  DRV_LAUNCH_CHECK();
  if (timed) ctx->stage_end(DRV_STAGE_APPLY_CACHES);
  return DRV_OK;
}
