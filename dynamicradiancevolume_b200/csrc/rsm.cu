// rsm.cu — stage 2: RSM mip chain and VPL generation.
//
//  drv_prepare_rsm      ≙ Renderer::ShadowMap::PrepareRSM (renderer.cpp:1300-1339)
//                         running shader/downsamplersm.frag:15-33 per level.
//  drv_impl_generate_vpls  the VPL load phase of shader/cacheLightingRSM.comp:137-163
//                         and the cache-independent half of its indirect-shadow
//                         sample (:168-192), hoisted out of the gather: the
//                         reference redoes both in every 64-cache work group
//                         (N/64 x R^2 texture fetches); here they run once per
//                         light per frame and the gather streams a 48-byte VPL
//                         record + a 16-byte block record instead.
#include "ctx.h"
#include "device_math.cuh"

using namespace drvk;

namespace {

struct RsmTexel {
  uint2 flux;     // 4 halfs (r,g,b,x)
  int normal;     // 2 int16
  uint32_t depth; // 2 halfs (dist, dist^2)
};

// downsamplersm.frag:15-33 for one destination texel from its 2x2 footprint, given in textureGather order:
// t[0] = (x0,y1), t[1] = (x1,y1), t[2] = (x1,y0), t[3] = (x0,y0). Results are rounded to the storage formats
// (half / int16) exactly as a render to the next mip level does, so chained levels see quantised inputs.
__device__ __forceinline__ RsmTexel rsm_downsample4(const RsmTexel (&t)[4]) {
  float fr[4], fg[4], fb[4], d0[4], d1[4];
  F3 n = {0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    fr[i] = half_bits_to_float((uint16_t)(t[i].flux.x & 0xffffu));
    fg[i] = half_bits_to_float((uint16_t)(t[i].flux.x >> 16));
    fb[i] = half_bits_to_float((uint16_t)(t[i].flux.y & 0xffffu));
    d0[i] = half_bits_to_float((uint16_t)(t[i].depth & 0xffffu));
    d1[i] = half_bits_to_float((uint16_t)(t[i].depth >> 16));
    F3 u = unpack_normal16i((int)(short)(t[i].normal & 0xffff), (int)(short)((uint32_t)t[i].normal >> 16));
    n.x += u.x; n.y += u.y; n.z += u.z;
  }
  RsmTexel o;
  // flux: sum of the four (:17-22); the RGB16F store rounds to nearest even
  float sr = ex_add(ex_add(ex_add(fr[0], fr[1]), fr[2]), fr[3]);
  float sg = ex_add(ex_add(ex_add(fg[0], fg[1]), fg[2]), fg[3]);
  float sb = ex_add(ex_add(ex_add(fb[0], fb[1]), fb[2]), fb[3]);
  o.flux = make_uint2((uint32_t)float_to_half_bits(sr) | ((uint32_t)float_to_half_bits(sg) << 16),
                      (uint32_t)float_to_half_bits(sb));
  // normal: mean direction, renormalised, repacked (:24-30)
  float inv = rsqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
  int ox, oy;
  pack_normal16i(n.x * inv, n.y * inv, n.z * inv, ox, oy);
  o.normal = (int)(((uint32_t)ox & 0xffffu) | ((uint32_t)oy << 16));
  // depthLinSq: linear fetch at the footprint centre = mean of the four (:32)
  float a0 = ex_mix(d0[3], d0[2], 0.5f), b0 = ex_mix(d0[0], d0[1], 0.5f);
  float a1 = ex_mix(d1[3], d1[2], 0.5f), b1 = ex_mix(d1[0], d1[1], 0.5f);
  o.depth = (uint32_t)float_to_half_bits(ex_mix(a0, b0, 0.5f)) | ((uint32_t)float_to_half_bits(ex_mix(a1, b1, 0.5f)) << 16);
  return o;
}

// Several mip levels per launch (the reference renders one full-screen pass per level, renderer.cpp:1328-1338).
// A block owns a 32x32 tile of the source level and produces its 16x16, 8x8, ... footprint in up to
// `levels` (<= 5) successive levels, passing each level to the next through shared memory. With
// src_res <= 32 a single block walks the rest of the chain. Level l (>= 1) of the chain lives at texel
// offset `off[l - first_level]`.
struct RsmMipArgs {
  const uint2* flux_src; const int* normal_src; const uint32_t* depth_src;
  uint2* flux_mips; int* normal_mips; uint32_t* depth_mips;
  uint32_t off[5]; // texel offsets of the destination levels
  int src_res;     // resolution of the source level
  int levels;      // destination levels to produce
};

__global__ void __launch_bounds__(256) rsm_mip_chain_kernel(RsmMipArgs A) {
  __shared__ RsmTexel s_lvl[2][256];
  const int tile = min(A.src_res, 32);
  const int tx0 = blockIdx.x * tile, ty0 = blockIdx.y * tile;
  int src = A.src_res, span = tile; // span = this block's extent in the current source level
  for (int l = 0; l < A.levels; ++l) {
    const int h = src >> 1, hs = span >> 1; // destination resolution / this block's extent in it
    const int ox0 = (tx0 >> (l + 1)), oy0 = (ty0 >> (l + 1));
    RsmTexel* cur = s_lvl[l & 1];
    const RsmTexel* prev = s_lvl[(l & 1) ^ 1];
    for (int i = threadIdx.x; i < hs * hs; i += blockDim.x) {
      const int lx = i % hs, ly = i / hs;
      RsmTexel t[4];
      if (l == 0) {
        const int gx = tx0 + 2 * lx, gy = ty0 + 2 * ly;
        const size_t q[4] = {(size_t)(gy + 1) * src + gx, (size_t)(gy + 1) * src + gx + 1, (size_t)gy * src + gx + 1,
                             (size_t)gy * src + gx};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          t[k].flux = __ldg(A.flux_src + q[k]);
          t[k].normal = __ldg(A.normal_src + q[k]);
          t[k].depth = __ldg(A.depth_src + q[k]);
        }
      } else {
        const int q[4] = {(2 * ly + 1) * span + 2 * lx, (2 * ly + 1) * span + 2 * lx + 1, (2 * ly) * span + 2 * lx + 1,
                          (2 * ly) * span + 2 * lx};
#pragma unroll
        for (int k = 0; k < 4; ++k) t[k] = prev[q[k]];
      }
      const RsmTexel o = rsm_downsample4(t);
      cur[i] = o;
      const size_t g = A.off[l] + (size_t)(oy0 + ly) * h + (ox0 + lx);
      A.flux_mips[g] = o.flux;
      A.normal_mips[g] = o.normal;
      A.depth_mips[g] = o.depth;
    }
    __syncthreads();
    src = h;
    span = hs;
  }
}

// cacheLightingRSM.comp:154-155 / :175-176 in decision maths (the block record
// feeds the cone-march trip count).
__device__ __forceinline__ F3 rsm_world_position(const drv_spot_light& L, float u, float v, float d) {
  F3 ws = ex_unproject(L.InverseLightViewProjection, ex_sub(ex_mul(u, 2.0f), 1.0f), ex_sub(ex_mul(v, 2.0f), 1.0f), 0.0f);
  float dx = ex_sub(ws.x, L.LightPosition[0]), dy = ex_sub(ws.y, L.LightPosition[1]), dz = ex_sub(ws.z, L.LightPosition[2]);
  float inv = ex_rsqrt(ex_dot3(dx, dy, dz, dx, dy, dz));
  F3 r = {ex_add(L.LightPosition[0], ex_mul(ex_mul(dx, inv), d)), ex_add(L.LightPosition[1], ex_mul(ex_mul(dy, inv), d)),
          ex_add(L.LightPosition[2], ex_mul(ex_mul(dz, inv), d))};
  return r;
}

// Threads [0, R^2) write VPLs; threads [R^2, R^2 + numBlocks) write shadow-block records.
__global__ void vplgen_kernel(drv_spot_light L, const uint2* __restrict__ flux, const int* __restrict__ normal,
                              const uint32_t* __restrict__ depth, const uint32_t* __restrict__ depth_lod,
                              int with_blocks, float4* __restrict__ vpls, float4* __restrict__ blocks,
                              uint32_t* __restrict__ chunk_counts, uint8_t* __restrict__ block_live, int specular) {
  const uint32_t R = (uint32_t)L.RSMReadResolution;
  const uint32_t total = R * R;
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  // live-VPL bookkeeping of the compaction (see vpl_compact_kernel), fused in: per-256-chunk counts + block flags
  {
    bool live = false;
    if (k < total) {
      uint32_t x, y;
      morton_decode(k, x, y);
      const uint2 f = __ldg(flux + (size_t)y * R + x);
      // half bits: zero flux = all three halfs are +-0. With INDIRECT_SPECULAR the shader itself skips a VPL whose
      // flux sums to less than 0.001 — for the SH as well (cacheLightingRSM.comp:241)
      live = ((f.x & 0x7fff7fffu) | (f.y & 0x7fffu)) != 0u;
      if (specular)
        live = !(ex_add(ex_add(half_bits_to_float((uint16_t)(f.x & 0xffffu)), half_bits_to_float((uint16_t)(f.x >> 16))),
                        half_bits_to_float((uint16_t)(f.y & 0xffffu))) < 0.001f);
      if (live && with_blocks) block_live[k / (uint32_t)L.IndirectShadowComputationSampleInterval] = 1;
    }
    const int c = __syncthreads_count(live);
    if (threadIdx.x == 0 && blockIdx.x * blockDim.x < total) chunk_counts[blockIdx.x] = (uint32_t)c;
  }
  if (k < total) {
    uint32_t x, y;
    morton_decode(k, x, y);                                             // :142
    float u = ex_div((float)x + 0.5f, (float)R), v = ex_div((float)y + 0.5f, (float)R); // :144
    size_t t = (size_t)y * R + x;
    uint2 f = __ldg(flux + t);                                          // :147
    float d = half_bits_to_float((uint16_t)(__ldg(depth + t) & 0xffffu)); // :150
    F3 p = rsm_world_position(L, u, v, d);                              // :154-155
    int pn = __ldg(normal + t);
    F3 n = unpack_normal16i((int)(short)(pn & 0xffff), (int)(short)((uint32_t)pn >> 16)); // :158
    float4* o = vpls + (size_t)k * 3;
    o[0] = make_float4(p.x, p.y, p.z, ex_mul(ex_mul(d, d), L.ValAreaFactor)); // :151
    o[1] = make_float4(n.x, n.y, n.z, 0.0f);
    o[2] = make_float4(half_bits_to_float((uint16_t)(f.x & 0xffffu)), half_bits_to_float((uint16_t)(f.x >> 16)),
                       half_bits_to_float((uint16_t)(f.y & 0xffffu)), 0.0f);
    return;
  }
  if (!with_blocks) return;
  const uint32_t interval = (uint32_t)L.IndirectShadowComputationSampleInterval;
  uint32_t b = k - total;
  if (b >= total / interval) return;
  const int lod = (int)L.IndirectShadowComputationLod;
  const int Rl = (int)(R >> lod);
  uint32_t x, y;
  morton_decode(b * interval, x, y);                                                // :171
  float u = ex_div(ex_add((float)x, L.IndirectShadowSamplingOffset), (float)R);     // :172
  float v = ex_div(ex_add((float)y, L.IndirectShadowSamplingOffset), (float)R);
  // :174 — bilinear, clamp to edge, at mip `lod` of the read level
  float fx = ex_sub(ex_mul(u, (float)Rl), 0.5f), fy = ex_sub(ex_mul(v, (float)Rl), 0.5f);
  float flx = floorf(fx), fly = floorf(fy);
  float tx = ex_sub(fx, flx), ty = ex_sub(fy, fly);
  int x0 = ex_trunc(flx), y0 = ex_trunc(fly);
  int x1 = clampi(x0 + 1, 0, Rl - 1), y1 = clampi(y0 + 1, 0, Rl - 1);
  x0 = clampi(x0, 0, Rl - 1); y0 = clampi(y0, 0, Rl - 1);
  uint32_t t00 = __ldg(depth_lod + (size_t)y0 * Rl + x0), t10 = __ldg(depth_lod + (size_t)y0 * Rl + x1);
  uint32_t t01 = __ldg(depth_lod + (size_t)y1 * Rl + x0), t11 = __ldg(depth_lod + (size_t)y1 * Rl + x1);
  float m1 = ex_mix(ex_mix(half_bits_to_float((uint16_t)(t00 & 0xffffu)), half_bits_to_float((uint16_t)(t10 & 0xffffu)), tx),
                    ex_mix(half_bits_to_float((uint16_t)(t01 & 0xffffu)), half_bits_to_float((uint16_t)(t11 & 0xffffu)), tx), ty);
  float m2 = ex_mix(ex_mix(half_bits_to_float((uint16_t)(t00 >> 16)), half_bits_to_float((uint16_t)(t10 >> 16)), tx),
                    ex_mix(half_bits_to_float((uint16_t)(t01 >> 16)), half_bits_to_float((uint16_t)(t11 >> 16)), tx), ty);
  F3 avg = rsm_world_position(L, u, v, m1);                                          // :175-176
  float var = ex_sub(m2, ex_mul(m1, m1));                                            // :181
  if (!(var > 0.0f)) var = 0.0f;
  float k2 = ex_div(ex_mul(ex_sqrt(var), 2.0f), m1);
  float kc = fmaxf(L.IndirectShadowComputationSuperValWidth, k2);                    // :192 (fmaxf drops a NaN k2)
  if (kc != kc) kc = L.IndirectShadowComputationSuperValWidth;
  blocks[b] = make_float4(avg.x, avg.y, avg.z, kc);
}

// ---- live-VPL compaction ------------------------------------------------------------------------------
// A VPL whose flux is (+-)0 in all three channels adds exactly zero to every accumulator of every cache
// (cacheLightingRSM.comp:262-277: every term carries the factor Flux), so the gather never needs to see it:
// RSM texels outside the spot cone or past the scene are such VPLs. Two small kernels keep the order of the
// survivors (deterministic sums): per-256-chunk counts, then an ordered scatter. NaN flux compares != 0 and
// stays in. Each survivor is tagged with its shadow-block index so a compacted list still finds its visibility.
constexpr int kCompactThreads = 256;

__device__ __forceinline__ bool vpl_is_live(const float4 flux, int specular) {
  if (specular) return !(ex_add(ex_add(flux.x, flux.y), flux.z) < 0.001f); // cacheLightingRSM.comp:241
  return flux.x != 0.0f || flux.y != 0.0f || flux.z != 0.0f;
}

__global__ void __launch_bounds__(kCompactThreads) vpl_count_kernel(const float4* __restrict__ vpls, uint32_t n,
                                                                   uint32_t interval, uint32_t* __restrict__ chunk_counts,
                                                                   uint8_t* __restrict__ block_live, int specular) {
  const uint32_t k = blockIdx.x * kCompactThreads + threadIdx.x;
  const bool live = k < n && vpl_is_live(__ldg(vpls + (size_t)k * 3 + 2), specular);
  if (live && block_live) block_live[k / interval] = 1; // benign race: every writer stores 1
  const int c = __syncthreads_count(live);
  if (threadIdx.x == 0) chunk_counts[blockIdx.x] = (uint32_t)c;
}

__global__ void __launch_bounds__(kCompactThreads) vpl_compact_kernel(const float4* __restrict__ vpls, uint32_t n,
                                                                     uint32_t interval, const uint32_t* __restrict__ chunk_counts,
                                                                     float4* __restrict__ out, uint32_t* __restrict__ live_count,
                                                                     int specular) {
  __shared__ uint32_t s_warp[kCompactThreads / 32];
  __shared__ uint32_t s_base;
  // survivors in the chunks before this one
  uint32_t part = 0;
  for (uint32_t c = threadIdx.x; c < blockIdx.x; c += kCompactThreads) part += __ldg(chunk_counts + c);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_warp[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t b = 0;
    for (int w = 0; w < kCompactThreads / 32; ++w) b += s_warp[w];
    s_base = b;
  }
  __syncthreads();
  const uint32_t base = s_base;
  const uint32_t k = blockIdx.x * kCompactThreads + threadIdx.x;
  float4 q0, q1, q2 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k < n) q2 = __ldg(vpls + (size_t)k * 3 + 2);
  const bool live = k < n && vpl_is_live(q2, specular);
  const uint32_t ballot = __ballot_sync(0xffffffffu, live);
  __syncthreads(); // s_warp is reused
  if (lane == 0) s_warp[warp] = __popc(ballot);
  __syncthreads();
  uint32_t pos = base + __popc(ballot & ((1u << lane) - 1u));
  for (uint32_t w = 0; w < warp; ++w) pos += s_warp[w];
  if (live) {
    q0 = __ldg(vpls + (size_t)k * 3);
    q1 = __ldg(vpls + (size_t)k * 3 + 1);
    q1.w = __uint_as_float(k / interval); // shadow-block index of this VPL (cacheLightingRSM.comp:169)
    float4* o = out + (size_t)pos * 3;
    o[0] = q0; o[1] = q1; o[2] = q2;
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    uint32_t total = base;
    for (int w = 0; w < kCompactThreads / 32; ++w) total += s_warp[w];
    *live_count = total;
  }
}

} // namespace

drv_status drv_impl_compact_vpls(drv_ctx* ctx, uint32_t li, bool counted) {
  LightState& S = ctx->lights[li];
  const uint32_t n = S.num_vpls;
  if (n == 0) {
    DRV_CUDA(cudaMemsetAsync(ctx->live_counts + li, 0, sizeof(uint32_t), ctx->stream));
    return DRV_OK;
  }
  const bool shadow = ctx->cfg.indirect_shadow != 0 && !S.vpls_external;
  const int spec = ctx->cfg.indirect_specular ? 1 : 0;
  uint32_t interval = shadow ? (uint32_t)S.block.IndirectShadowComputationSampleInterval : 1u;
  if (interval == 0) interval = 1;
  const uint32_t chunks = (n + kCompactThreads - 1) / kCompactThreads;
  if (!counted) { // external lists (drv_set_vpls); vplgen_kernel does this itself
    if (shadow) DRV_CUDA(cudaMemsetAsync(S.block_live, 0, (n + interval - 1) / interval, ctx->stream));
    vpl_count_kernel<<<chunks, kCompactThreads, 0, ctx->stream>>>((const float4*)S.vpls, n, interval, S.chunk_counts,
                                                                 shadow ? S.block_live : nullptr, spec);
    DRV_LAUNCH_CHECK();
  }
  vpl_compact_kernel<<<chunks, kCompactThreads, 0, ctx->stream>>>((const float4*)S.vpls, n, interval, S.chunk_counts,
                                                                 (float4*)S.vpls_live, ctx->live_counts + li, spec);
  DRV_LAUNCH_CHECK();
  return DRV_OK;
}

// `only_consumed`: stop after the last level this frame's consumers read (the VPL read level and, with indirect
// shadows, the shadow-sample level below it) instead of walking the chain down to 2x2.
drv_status drv_impl_prepare_rsm(drv_ctx* ctx, uint32_t li, bool only_consumed) {
  LightState& S = ctx->lights[li];
  if (!S.rsm_bound) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_prepare_rsm: RSM not bound");
  ctx->stage_begin(DRV_STAGE_PREPARE_RSM);
  uint32_t res = S.rsm_res;
  const uint2* fs = (const uint2*)S.flux0;
  const int* ns = (const int*)S.normal0;
  const uint32_t* ds = (const uint32_t*)S.depth0;
  // levels 1 .. log2(res)-1: the 1x1 top level is never rendered (renderer.cpp:1293-1297)
  uint32_t last = 0;
  while ((res >> (last + 1)) >= 2) ++last; // last level to produce
  if (only_consumed && S.block_set && S.block.RSMReadResolution > 0 && (uint32_t)S.block.RSMRenderResolution == res) {
    uint32_t need = 0;
    while ((res >> need) > (uint32_t)S.block.RSMReadResolution) ++need; // TEXTURE_BASE_LEVEL = rsmReadLod
    if (ctx->cfg.indirect_shadow) need += (uint32_t)S.block.IndirectShadowComputationLod;
    if (need < last) last = need;
  }
  uint32_t level = 1;                      // next level to produce
  while (level <= last) {
    RsmMipArgs A;
    memset(&A, 0, sizeof(A));
    const uint32_t src_res = res >> (level - 1);
    A.flux_src = level == 1 ? fs : (const uint2*)S.flux_mips + rsm_level_offset_texels(res, level - 1);
    A.normal_src = level == 1 ? ns : (const int*)S.normal_mips + rsm_level_offset_texels(res, level - 1);
    A.depth_src = level == 1 ? ds : (const uint32_t*)S.depth_mips + rsm_level_offset_texels(res, level - 1);
    A.flux_mips = (uint2*)S.flux_mips; A.normal_mips = (int*)S.normal_mips; A.depth_mips = (uint32_t*)S.depth_mips;
    A.src_res = (int)src_res;
    // a 32x32 source tile yields 5 levels (16..1); a single block (src_res <= 32) walks down to the 2x2 level
    uint32_t n = src_res > 32 ? 5u : 31u;
    if (n > last - level + 1) n = last - level + 1;
    if (n > 5) n = 5;
    A.levels = (int)n;
    for (uint32_t k = 0; k < n; ++k) A.off[k] = (uint32_t)rsm_level_offset_texels(res, level + k);
    const uint32_t tiles = src_res > 32 ? src_res / 32 : 1;
    rsm_mip_chain_kernel<<<dim3(tiles, tiles), 256, 0, ctx->stream>>>(A);
    DRV_LAUNCH_CHECK();
    level += n;
  }
  ctx->stage_end(DRV_STAGE_PREPARE_RSM);
  return DRV_OK;
}

drv_status drv_impl_generate_vpls(drv_ctx* ctx, uint32_t li) {
  LightState& S = ctx->lights[li];
  if (!S.block_set) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_light_caches: SpotLight block not set");
  if (!S.rsm_bound) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_light_caches: RSM not bound");
  const drv_spot_light& L = S.block;
  const uint32_t R = (uint32_t)L.RSMReadResolution;
  if ((uint32_t)L.RSMRenderResolution != S.rsm_res || R == 0 || (S.rsm_res % R) != 0 || R > ctx->cfg.max_rsm_resolution)
    return ctx->fail(DRV_ERR_INVALID, "drv_light_caches: SpotLight block disagrees with the bound RSM");
  uint32_t readLod = 0;
  while ((S.rsm_res >> readLod) > R) ++readLod; // TEXTURE_BASE_LEVEL = rsmReadLod, renderer.cpp:807-809
  auto level_ptr = [&](const void* l0, const void* mips, uint32_t level, size_t texel_bytes) -> const void* {
    if (level == 0) return l0;
    return (const uint8_t*)mips + rsm_level_offset_texels(S.rsm_res, level) * texel_bytes;
  };
  const uint2* flux = (const uint2*)level_ptr(S.flux0, S.flux_mips, readLod, 8);
  const int* normal = (const int*)level_ptr(S.normal0, S.normal_mips, readLod, 4);
  const uint32_t* depth = (const uint32_t*)level_ptr(S.depth0, S.depth_mips, readLod, 4);
  int with_blocks = ctx->cfg.indirect_shadow ? 1 : 0;
  const uint32_t* depth_lod = depth;
  uint32_t nblocks = 0;
  if (with_blocks) {
    uint32_t slod = (uint32_t)L.IndirectShadowComputationLod;
    uint32_t interval = (uint32_t)L.IndirectShadowComputationSampleInterval;
    // the sampled level must have been rendered (block < read resolution)
    if (interval == 0 || (R >> slod) < 2 || interval != (1u << (2 * slod)))
      return ctx->fail(DRV_ERR_INVALID, "drv_light_caches: indirect shadow LOD out of range for this RSM");
    depth_lod = (const uint32_t*)level_ptr(S.depth0, S.depth_mips, readLod + slod, 4);
    nblocks = R * R / interval;
  }
  uint32_t threads = R * R + nblocks;
  static_assert(kCompactThreads == 256, "vplgen_kernel's chunks are the compaction's chunks");
  if (with_blocks) DRV_CUDA(cudaMemsetAsync(S.block_live, 0, nblocks, ctx->stream));
  vplgen_kernel<<<(threads + 255) / 256, 256, 0, ctx->stream>>>(L, flux, normal, depth, depth_lod, with_blocks,
                                                               (float4*)S.vpls, (float4*)S.blocks, S.chunk_counts,
                                                               S.block_live, ctx->cfg.indirect_specular ? 1 : 0);
  DRV_LAUNCH_CHECK();
  S.num_vpls = R * R;
  return drv_impl_compact_vpls(ctx, li, true);
}
