// apply.cu — stage 5: per-pixel SH cache interpolation
// (≙ Renderer::ApplyCaches, rendering/renderer.cpp:1047-1079;
// shader/cacheApply.frag:28-195; shader/lightcache.glsl:109-183).
//
// One thread per pixel on 32x8 blocks (a warp = 32 consecutive pixels of a
// row, so depth / normal / diffuse loads and the HDR read-modify-write are
// fully coalesced). The eight (sixteen in a transition zone) cache entries
// are fetched with 128-bit loads; neighbouring pixels hit the same entries,
// which L1 absorbs, and the whole entry list + atlas is L2-resident.
// Cascade choice and cell selection are DECISION maths (device_math.cuh) so
// they agree with the allocation stage and with the oracle bit for bit.
#include "ctx.h"
#include "device_math.cuh"

using namespace drvk;

namespace {

__constant__ float c_srgb_lut[256]; // exact piecewise sRGB EOTF per 8-bit code (gbuffer.glsl:1, renderer.cpp:468)

struct ApplyParams {
  int W, H, R, C;
  int transitions;
  float zone;
  float ivp[16];
  drv_cav_cascade casc[DRV_MAX_CASCADES];
  float g0, g1, g2, g20, g22; // ShCosLobeFactor*
  uint32_t max_caches;
  float inv_voxel[DRV_MAX_CASCADES]; // 1 / WorldVoxelSize when that is a power of two (exact), else 0
};

__device__ __forceinline__ int compute_cascade(const ApplyParams& p, F3 wp) { // lightcache.glsl:109-122
  int c = 0;
  for (; c < p.C - 1; ++c) {
    const drv_cav_cascade& k = p.casc[c];
    // six compares and one branch (bitwise &: no short-circuit jumps; NaN fails every compare either way)
    if ((wp.x <= k.DecisionMax[0]) & (wp.y <= k.DecisionMax[1]) & (wp.z <= k.DecisionMax[2]) &
        (wp.x >= k.DecisionMin[0]) & (wp.y >= k.DecisionMin[1]) & (wp.z >= k.DecisionMin[2]))
      break;
  }
  return c;
}

__device__ __forceinline__ float cascade_transition(const ApplyParams& p, F3 wp, int c) { // lightcache.glsl:125-134
  const drv_cav_cascade& k = p.casc[c];
  float ax = ex_sub(k.DecisionMax[0], wp.x), ay = ex_sub(k.DecisionMax[1], wp.y), az = ex_sub(k.DecisionMax[2], wp.z);
  float bx = ex_sub(wp.x, k.DecisionMin[0]), by = ex_sub(wp.y, k.DecisionMin[1]), bz = ex_sub(wp.z, k.DecisionMin[2]);
  float minDist = fminf(fminf(fminf(ax, ay), az), fminf(fminf(bx, by), bz));
  // Outside the transition zone — nearly every pixel — the division is not needed: for a positive finite d and
  // minDist >= d the correctly rounded quotient is >= 1, so saturate(1 - q) is exactly +0 (NaN saturates to 0 too).
  const float d = ex_mul(k.WorldVoxelSize, p.zone);
  if (d > 0.0f && d <= 3.0e38f && !(minDist < d)) return 0.0f;
  return saturatef(ex_sub(1.0f, ex_div(minDist, d)));
}

// SampleCacheIrradiance, lightcache.glsl:137-183. nb* = the per-pixel normal
// factors, hoisted out of the 8-corner loop.
template <int ORDER>
struct NormalBasis {
  float b1y, b1z, b1x;                 // g1 * n.{y,z,x}
  float b2xy, b2yz, b20, b2xz, b2dd;   // band 2
};

// One corner: entry `address` weighted by w. Branch-free, so the 24 (56) entry loads of a pixel are independent of
// each other and of the atlas contents and can all be in flight together. A corner without a cache (`valid` false):
//  * SH2 (ZERO_SLOT false): reads entry 0 — always inside the buffer — and selects zero irradiance;
//  * SH1 (ZERO_SLOT true): `address` is max_cache_count. The LightCacheBuffer is max_cache_count x 128 bytes whatever
//    the mode (renderer.cpp:266-269), SH1 entries are 64 bytes, so slot max_cache_count lies in the half of the
//    buffer no stage ever writes: zeros since drv_create. Its irradiance is max(+0, 0) = 0 and fma(0, w, r) == r bit
//    for bit — no selects, and the address is one unsigned min instead of a compare + select.
template <int ORDER, bool ZERO_SLOT>
__device__ __forceinline__ void accumulate_corner(const ApplyParams& p, const uint8_t* __restrict__ entries,
                                                  uint32_t address, bool valid, const NormalBasis<ORDER>& nb, float w,
                                                  float& r, float& g, float& b) {
  constexpr uint32_t STRIDE = ORDER == 2 ? 128 : 64;
  const float4* e = reinterpret_cast<const float4*>(entries + address * STRIDE); // < 2^31 bytes: checked at create
  float4 q1 = __ldg(e + 1), q2 = __ldg(e + 2), q3 = __ldg(e + 3); // (SH1neg1,SH00_r) (SH10,SH00_g) (SH1pos1,SH00_b)
  float ir = q1.w * p.g0, ig = q2.w * p.g0, ib = q3.w * p.g0;
  ir = fmaf(-q1.x, nb.b1y, ir); ig = fmaf(-q1.y, nb.b1y, ig); ib = fmaf(-q1.z, nb.b1y, ib);
  ir = fmaf(q2.x, nb.b1z, ir);  ig = fmaf(q2.y, nb.b1z, ig);  ib = fmaf(q2.z, nb.b1z, ib);
  ir = fmaf(-q3.x, nb.b1x, ir); ig = fmaf(-q3.y, nb.b1x, ig); ib = fmaf(-q3.z, nb.b1x, ib);
  if (ORDER == 2) {
    float4 q4 = __ldg(e + 4), q5 = __ldg(e + 5), q6 = __ldg(e + 6), q7 = __ldg(e + 7);
    // (SH2neg2,SH20_r) (SH2neg1,SH20_g) (SH2pos1,SH20_b) (SH2pos2,-)
    ir = fmaf(-q4.x, nb.b2xy, ir); ig = fmaf(-q4.y, nb.b2xy, ig); ib = fmaf(-q4.z, nb.b2xy, ib);
    ir = fmaf(q5.x, nb.b2yz, ir);  ig = fmaf(q5.y, nb.b2yz, ig);  ib = fmaf(q5.z, nb.b2yz, ib);
    ir = fmaf(q4.w, nb.b20, ir);   ig = fmaf(q5.w, nb.b20, ig);   ib = fmaf(q6.w, nb.b20, ib);
    ir = fmaf(q6.x, nb.b2xz, ir);  ig = fmaf(q6.y, nb.b2xz, ig);  ib = fmaf(q6.z, nb.b2xz, ib);
    ir = fmaf(q7.x, nb.b2dd, ir);  ig = fmaf(q7.y, nb.b2dd, ig);  ib = fmaf(q7.z, nb.b2dd, ib);
  }
  // max(irradiance, 0) then * weight, lightcache.glsl:178, cacheApply.frag:110; no cache -> exactly zero
  if (ZERO_SLOT) {
    r = fmaf(fmaxf(ir, 0.0f), w, r);
    g = fmaf(fmaxf(ig, 0.0f), w, g);
    b = fmaf(fmaxf(ib, 0.0f), w, b);
  } else {
    r = fmaf(valid ? fmaxf(ir, 0.0f) : 0.0f, w, r);
    g = fmaf(valid ? fmaxf(ig, 0.0f) : 0.0f, w, g);
    b = fmaf(valid ? fmaxf(ib, 0.0f) : 0.0f, w, b);
  }
}

// The SH1 corner with red and green riding in one packed register pair: the .xy of every coefficient's float4 IS an
// aligned pair, the per-pixel basis factor enters as a 32-bit broadcast operand (FFMA2 R.F32 form). The same fused
// operations in the same order as the scalar form — bit-identical — in 13 instead of 18 FP32-pipe instructions.
__device__ __forceinline__ void accumulate_corner_sh1_packed(const ApplyParams& p, const uint8_t* __restrict__ entries,
                                                             uint32_t address, const NormalBasis<1>& nb, float w,
                                                             float2& rg, float& b) {
  const float4* e = reinterpret_cast<const float4*>(entries + address * 64u);
  const float4 q1 = __ldg(e + 1), q2 = __ldg(e + 2), q3 = __ldg(e + 3);
  float2 irg = make_float2(q1.w * p.g0, q2.w * p.g0);
  float ib = q3.w * p.g0;
  // fma(-a, b, c) == fma(a, -b, c) exactly: the sign moves to the broadcast factor
  irg = __ffma2_rn(make_float2(q1.x, q1.y), make_float2(-nb.b1y, -nb.b1y), irg); ib = fmaf(-q1.z, nb.b1y, ib);
  irg = __ffma2_rn(make_float2(q2.x, q2.y), make_float2(nb.b1z, nb.b1z), irg);   ib = fmaf(q2.z, nb.b1z, ib);
  irg = __ffma2_rn(make_float2(q3.x, q3.y), make_float2(-nb.b1x, -nb.b1x), irg); ib = fmaf(-q3.z, nb.b1x, ib);
  irg.x = fmaxf(irg.x, 0.0f); irg.y = fmaxf(irg.y, 0.0f);
  rg = __ffma2_rn(irg, make_float2(w, w), rg);
  b = fmaf(fmaxf(ib, 0.0f), w, b);
}

// ComputeLightingFromCaches, cacheApply.frag:28-118 (before the * diffuse / PI).
template <int ORDER, bool PACK>
__device__ __forceinline__ void lighting_from_caches(const ApplyParams& p, const uint32_t* __restrict__ atlas,
                                                     const uint8_t* __restrict__ entries, F3 wp,
                                                     const NormalBasis<ORDER>& nb, int c, float& r, float& g, float& b) {
  const drv_cav_cascade& k = p.casc[c];
  const float inv = p.inv_voxel[c];
  float ax = ex_sub(wp.x, k.Min[0]), ay = ex_sub(wp.y, k.Min[1]), az = ex_sub(wp.z, k.Min[2]);
  if (inv != 0.0f) { ax = ex_mul(ax, inv); ay = ex_mul(ay, inv); az = ex_mul(az, inv); } // power-of-two voxel: exact
  else { ax = ex_div(ax, k.WorldVoxelSize); ay = ex_div(ay, k.WorldVoxelSize); az = ex_div(az, k.WorldVoxelSize); }
  const int bx = ex_trunc(ax), by = ex_trunc(ay), bz = ex_trunc(az);
  const float fx = ax - (float)bx, fy = ay - (float)by, fz = az - (float)bz;
  const float gx = 1.0f - fx, gy = 1.0f - fy, gz = 1.0f - fz;
  const int atlasW = p.R * p.C;
  const int x0 = bx + p.R * c;
  r = g = b = 0.0f;
  // first all eight address fetches (independent loads in flight together), then the entries
  uint32_t addr[8];
  const bool interior = bx >= 0 && bx + 1 < p.R && by >= 0 && by + 1 < p.R && bz >= 0 && bz + 1 < p.R;
  if (interior) {
    const uint32_t* a0 = atlas + (uint32_t)(x0 + atlasW * (by + p.R * bz));
    const uint32_t sy = (uint32_t)atlasW, sz = (uint32_t)(atlasW * p.R);
#pragma unroll
    for (int i = 0; i < 8; ++i) addr[i] = __ldg(a0 + (i & 1) + ((i >> 1) & 1) * sy + (i >> 2) * sz);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) { // texelFetch outside the texture returns 0
      const int x = x0 + (i & 1), y = by + ((i >> 1) & 1), z = bz + (i >> 2);
      addr[i] = 0u;
      if (x >= 0 && x < atlasW && y >= 0 && y < p.R && z >= 0 && z < p.R)
        addr[i] = __ldg(atlas + (uint32_t)(x + atlasW * (y + p.R * z)));
    }
  }
  const float wxy[4] = {gx * gy, fx * gy, gx * fy, fx * fy};
  float2 rg = make_float2(0.0f, 0.0f);
#pragma unroll
  for (int i = 0; i < 8; ++i) { // offsets in cacheApply.frag:43-54 order: x fastest, then y, then z
    const float w = wxy[i & 3] * ((i >> 2) ? fz : gz);
    const uint32_t address = addr[i] - 1u; // atlas 0 -> 0xFFFFFFFF: no cache, contributes zero
    if constexpr (ORDER == 1 && PACK) {
      accumulate_corner_sh1_packed(p, entries, min(address, p.max_caches), nb, w, rg, b);
    } else if (ORDER == 1) { // the never-written slot max_caches of the SH1 layout stands in for "no cache"
      accumulate_corner<ORDER, true>(p, entries, min(address, p.max_caches), true, nb, w, r, g, b);
    } else {
      const bool valid = address < p.max_caches;
      accumulate_corner<ORDER, false>(p, entries, valid ? address : 0u, valid, nb, w, r, g, b);
    }
  }
  if constexpr (ORDER == 1 && PACK) { r = rg.x; g = rg.y; }
}

// Pixels per thread: a thread walks kApplyIter rows (8 apart... see the kernel) and fetches the G-buffer texels of
// its next pixel before it shades the current one, so the DRAM latency of the depth / normal / albedo reads
// overlaps the atlas + entry gathers of the previous pixel.
constexpr int kApplyIter = 4;

template <int ORDER, bool PACK>
__device__ __forceinline__ void shade_pixel(const ApplyParams& p, const float* s_srgb, const uint32_t* __restrict__ atlas,
                                            const uint8_t* __restrict__ entries, const float* __restrict__ ndc_xy,
                                            void* __restrict__ out, int format, int x, int y, float d, int pn, uchar4 dc) {
  const uint32_t t = (uint32_t)y * p.W + x;
  if (d < 0.00001f) { // :128 discard
    if (format == DRV_HDR_RGBA32F_WRITE) reinterpret_cast<float4*>(out)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    else if (format == DRV_HDR_RGBA16F_WRITE) reinterpret_cast<uint2*>(out)[t] = make_uint2(0u, 0u);
    return;
  }
  // gl_FragCoord.xy -> NDC (:134) through the per-context tables
  F3 wp = ex_unproject(p.ivp, __ldg(ndc_xy + (uint32_t)x), __ldg(ndc_xy + (uint32_t)(p.W + y)), d);
  const int c = compute_cascade(p, wp); // :138
  F3 n = unpack_normal16i_fast((int)(short)(pn & 0xffff), (int)(short)((uint32_t)pn >> 16)); // :140
  const float albr = s_srgb[dc.x], albg = s_srgb[dc.y], albb = s_srgb[dc.z]; // :144
  NormalBasis<ORDER> nb;
  nb.b1y = p.g1 * n.y; nb.b1z = p.g1 * n.z; nb.b1x = p.g1 * n.x;
  if (ORDER == 2) {
    nb.b2xy = p.g2 * n.x * n.y;
    nb.b2yz = p.g2 * n.y * n.z;
    nb.b20 = p.g20 * (n.z * n.z * 3.0f - 1.0f);
    nb.b2xz = p.g2 * n.x * n.z;
    nb.b2dd = p.g22 * (n.x * n.x - n.y * n.y);
  }
  float r, g, b;
  lighting_from_caches<ORDER, PACK>(p, atlas, entries, wp, nb, c, r, g, b);
  if (p.transitions && c < p.C - 1) { // :172-184
    float tr = cascade_transition(p, wp, c);
    if (tr > 0.0f) {
      float r2, g2, b2;
      lighting_from_caches<ORDER, PACK>(p, atlas, entries, wp, nb, c + 1, r2, g2, b2);
      r = fmaf(r2 - r, tr, r); g = fmaf(g2 - g, tr, g); b = fmaf(b2 - b, tr, b);
    }
  }
  const float inv_pi = 1.0f / DRV_GLSL_PI;
  r = r * albr * inv_pi; g = g * albg * inv_pi; b = b * albb * inv_pi; // :114
  if (format == DRV_HDR_RGBA32F_WRITE) {
    reinterpret_cast<float4*>(out)[t] = make_float4(r, g, b, 1.0f);
  } else if (format == DRV_HDR_RGBA16F_WRITE) { // cleared target (0,0,0,0) + additive blend, in one store
    __half2 nrg = __floats2half2_rn(0.0f + r, 0.0f + g);
    __half2 nba = __floats2half2_rn(0.0f + b, 0.0f);
    uint2 nw;
    nw.x = *reinterpret_cast<uint32_t*>(&nrg);
    nw.y = *reinterpret_cast<uint32_t*>(&nba);
    reinterpret_cast<uint2*>(out)[t] = nw;
  } else { // additive blend GL_ONE, GL_ONE into RGBA16F (renderer.cpp:119, 480, 1053)
    uint2* o = reinterpret_cast<uint2*>(out) + t;
    uint2 old = *o;
    float2 rg = __half22float2(*reinterpret_cast<__half2*>(&old.x));
    float2 ba = __half22float2(*reinterpret_cast<__half2*>(&old.y));
    __half2 nrg = __floats2half2_rn(rg.x + r, rg.y + g);
    __half2 nba = __floats2half2_rn(ba.x + b, ba.y); // the shader outputs a vec3: alpha is left untouched
    uint2 nw;
    nw.x = *reinterpret_cast<uint32_t*>(&nrg);
    nw.y = *reinterpret_cast<uint32_t*>(&nba);
    *o = nw;
  }
}

// Block = 32 x 8 threads = a 32-wide column strip; it walks ITER consecutive 8-row groups.
template <int ORDER, int MINB, int ITER, bool PACK = false>
__global__ void __launch_bounds__(256, MINB) apply_kernel(ApplyParams p, const float* __restrict__ depth,
                                                          const int* __restrict__ normal, const uchar4* __restrict__ diffuse,
                                                          const uint32_t* __restrict__ atlas, const uint8_t* __restrict__ entries,
                                                          const float* __restrict__ ndc_xy, void* __restrict__ out, int format,
                                                          int y_begin, int y_end) {
  __shared__ float s_srgb[256]; // sRGB8 -> linear; shared memory serves divergent indices, constant memory would serialise
  s_srgb[threadIdx.y * 32 + threadIdx.x] = c_srgb_lut[threadIdx.y * 32 + threadIdx.x];
  __syncthreads();
  const int x = blockIdx.x * 32 + threadIdx.x;
  if (x >= p.W) return;
  int y = y_begin + blockIdx.y * (8 * ITER) + threadIdx.y;
  float d = 0.f; int pn = 0; uchar4 dc = make_uchar4(0, 0, 0, 0);
  if (y < y_end) {
    const uint32_t t = (uint32_t)y * p.W + x;
    d = __ldg(depth + t); pn = __ldg(normal + t); dc = __ldg(diffuse + t);
  }
#pragma unroll 1
  for (int it = 0; it < ITER; ++it, y += 8) {
    if (y >= y_end) return;
    const float d0 = d; const int pn0 = pn; const uchar4 dc0 = dc;
    if (it + 1 < ITER && y + 8 < y_end) { // next pixel's texels, in flight while this one is shaded
      const uint32_t t = (uint32_t)(y + 8) * p.W + x;
      d = __ldg(depth + t); pn = __ldg(normal + t); dc = __ldg(diffuse + t);
    }
    shade_pixel<ORDER, PACK>(p, s_srgb, atlas, entries, ndc_xy, out, format, x, y, d0, pn0, dc0);
  }
}

} // namespace

void drv_impl_upload_srgb_lut() {
  float lut[256];
  for (int v = 0; v < 256; ++v) {
    double c = (double)v / 255.0;
    double l = (c <= 0.04045) ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4);
    lut[v] = (float)l;
  }
  cudaMemcpyToSymbol(c_srgb_lut, lut, sizeof(lut));
}

drv_status drv_impl_apply(drv_ctx* ctx, void* out, uint32_t format) {
  return drv_impl_apply_rows(ctx, out, format, 0, ctx->cfg.backbuffer_height, true);
}

// Rows [y_begin, y_end) only: the host-frame pipeline applies a band as soon as its normals / albedo have arrived.
drv_status drv_impl_apply_rows(drv_ctx* ctx, void* out, uint32_t format, uint32_t y_begin, uint32_t y_end, bool timed) {
  if (!ctx->have_constant || !ctx->have_per_frame || !ctx->have_volume)
    return ctx->fail(DRV_ERR_NOT_BOUND, "drv_apply_caches: uniform blocks not set");
  if (!ctx->gb_depth || !ctx->gb_normal || !ctx->gb_diffuse)
    return ctx->fail(DRV_ERR_NOT_BOUND, "drv_apply_caches: g-buffer not bound");
  if (!out) return ctx->fail(DRV_ERR_INVALID, "drv_apply_caches: null output");
  if (format != DRV_HDR_RGBA16F_ADD && format != DRV_HDR_RGBA32F_WRITE && format != DRV_HDR_RGBA16F_WRITE)
    return ctx->fail(DRV_ERR_INVALID, "drv_apply_caches: unknown output format");
  ApplyParams p;
  p.W = ctx->constant.BackbufferResolution[0];
  p.H = ctx->constant.BackbufferResolution[1];
  p.R = ctx->constant.AddressVolumeResolution;
  p.C = ctx->constant.NumAddressVolumeCascades;
  if (p.W != (int)ctx->gb_w || p.H != (int)ctx->gb_h || p.R != (int)ctx->cfg.cav_resolution ||
      p.C != (int)ctx->cfg.cav_cascades)
    return ctx->fail(DRV_ERR_INVALID, "drv_apply_caches: Constant block disagrees with the context configuration");
  p.transitions = ctx->cfg.cascade_transitions ? 1 : 0;
  p.zone = ctx->volume.CAVTransitionZoneSize;
  memcpy(p.ivp, ctx->per_frame.InverseViewProjection, sizeof(p.ivp));
  memcpy(p.casc, ctx->volume.AddressVolumeCascades, sizeof(p.casc));
  p.g0 = ctx->constant.ShCosLobeFactor0;
  p.g1 = ctx->constant.ShCosLobeFactor1;
  p.g2 = ctx->constant.ShCosLobeFactor2n2_p1_n1;
  p.g20 = ctx->constant.ShCosLobeFactor20;
  p.g22 = ctx->constant.ShCosLobeFactor2p2;
  p.max_caches = ctx->cfg.max_cache_count;
  for (int c = 0; c < DRV_MAX_CASCADES; ++c) {
    int e = 0;
    const float v = p.casc[c].WorldVoxelSize;
    p.inv_voxel[c] = (v > 1e-6f && v < 1e6f && frexpf(v, &e) == 0.5f) ? 1.0f / v : 0.0f;
  }
  if (ctx->cfg.indirect_specular) // cacheApply.frag with INDIRECT_SPECULAR: its own kernel (specular.cu)
    return drv_impl_apply_specular(ctx, out, format, y_begin, y_end, timed, ctx->srgb_lut_dev);
  if (y_end > (uint32_t)p.H) y_end = (uint32_t)p.H;
  if (y_begin >= y_end) return DRV_OK;
  if (timed) ctx->stage_begin(DRV_STAGE_APPLY_CACHES);
  const uint32_t tune = (ctx->cfg.gather_variant >> 8) & 0xFu;
  const uint32_t iter_sel = (ctx->cfg.gather_variant >> 12) & 0xFu; // tuning sweeps: 1, 2 or 8 rows per thread
  const uint32_t iter = (iter_sel == 1 || iter_sel == 2 || iter_sel == 8) ? iter_sel : (uint32_t)kApplyIter;
  dim3 block(32, 8), grid((p.W + 31) / 32, (y_end - y_begin + 8 * iter - 1) / (8 * iter));
  // resident blocks per SM the kernel is compiled for (register budget 64 / 80 / 128 per thread): more registers
  // keep more of a pixel's entry loads in flight. drv_config.gather_variant bits 8..11 override it (tuning sweeps).
  // (bits 12..15 = 1: one pixel per thread, no prefetch loop)
#define DRV_APPLY(ORD, MB, IT)                                                                                       \
  apply_kernel<ORD, MB, IT><<<grid, block, 0, ctx->stream>>>(p, ctx->gb_depth, (const int*)ctx->gb_normal,            \
                                                             (const uchar4*)ctx->gb_diffuse, ctx->atlas, ctx->entries, \
                                                             ctx->ndc_xy, out, (int)format, (int)y_begin, (int)y_end)
#define DRV_APPLY_MB(ORD, IT)                                                                                   \
  do {                                                                                                          \
    if (tune == 4) DRV_APPLY(ORD, 4, IT); else if (tune == 6) DRV_APPLY(ORD, 6, IT); else DRV_APPLY(ORD, 5, IT); \
  } while (0)
  if (tune == 7 && ctx->cfg.sh_order == 1 && iter == (uint32_t)kApplyIter) { // SH1 with red / green packed (A/B switch)
    apply_kernel<1, 5, kApplyIter, true><<<grid, block, 0, ctx->stream>>>(p, ctx->gb_depth, (const int*)ctx->gb_normal,
                                                                          (const uchar4*)ctx->gb_diffuse, ctx->atlas, ctx->entries,
                                                                          ctx->ndc_xy, out, (int)format, (int)y_begin, (int)y_end);
  } else
  if (ctx->cfg.sh_order == 2) {
    if (iter == 1) DRV_APPLY_MB(2, 1); else if (iter == 2) DRV_APPLY_MB(2, 2); else if (iter == 8) DRV_APPLY_MB(2, 8);
    else DRV_APPLY_MB(2, kApplyIter);
  } else {
    if (iter == 1) DRV_APPLY_MB(1, 1); else if (iter == 2) DRV_APPLY_MB(1, 2); else if (iter == 8) DRV_APPLY_MB(1, 8);
    else DRV_APPLY_MB(1, kApplyIter);
  }
#undef DRV_APPLY_MB
#undef DRV_APPLY
  DRV_LAUNCH_CHECK();
  if (timed) ctx->stage_end(DRV_STAGE_APPLY_CACHES);
  return DRV_OK;
}
