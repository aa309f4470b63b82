// gather.cu — stage 4: the cache-entry x VPL irradiance gather projected into
// SH1 / SH2 with voxel-cone-traced indirect visibility
// (≙ Renderer::LightCachesRSM, rendering/renderer.cpp:899-933;
// shader/cacheLightingRSM.comp:83-374, INDIRECT_SPECULAR undefined).
//
// Reference shape: one thread per cache, 64 caches per group, every group
// re-derives all R^2 VPLs from three textures and walks them serially, once
// per light — parallelism = N_cache only (≈600 warps for a 40k-cache frame).
//
// B200 shape (this file):
//  * VPLs and shadow-block records are materialised once per light (rsm.cu)
//    and the VPLs with zero flux — which add exactly zero — are dropped there;
//    the kernels below stream the live list and read its length on the device.
//  * N-body tiling over BOTH axes. The work is the grid of UNITS
//    (cache tile of NT*CPT entries) x (32 VPLs of one light),
//    VPL unit fastest. A persistent grid (SM count x resident CTAs) splits the
//    flattened unit list into equal contiguous ranges ("stream-K"), so every
//    CTA gets the same number of units +-1 whatever N_cache and N_vpl are, and
//    the range bounds are computed on the device from the live cache counter —
//    no host read-back between allocation and lighting.
//  * each thread keeps CPT caches in registers (positions + 12/27 raw
//    accumulators) so every 48-byte VPL fetched from shared memory (three
//    broadcast 128-bit loads) feeds CPT pair evaluations.
//  * VPL tiles are staged global->shared with a register-prefetch pipeline
//    (default) or TMA bulk copies completing on an mbarrier (variant 1).
//  * the packed variants evaluate two caches per instruction with FP32x2
//    maths (FFMA2/FMUL2/FADD2, new on sm_100) to halve issue-slot pressure.
//  * SH basis constants are folded out of the loop: the kernel accumulates raw
//    moments (rad, rad*t, rad*t_i*t_j) and applies ShEvaFactor* once in the
//    epilogue: 28 FP32 ops + 2 MUFU per SH1 pair instead of 32 + 2.
//  * a CTA whose range covers a cache tile completely adds straight into the
//    entries (`entry.SH += acc`, :358-373). Ranges that start or end inside a
//    tile write their partial sums to a per-CTA scratch slot; after a
//    grid-wide barrier (cooperative launch) all CTAs add those in a fixed order
//    (deterministic — no float atomics) and, when peers are mapped, store the
//    finished entry to every other GPU over NVLink (fused all-gather).
//  * indirect shadows: cone_kernel first traces every (cache, live shadow
//    block) cone into a visibility table (device-side work queue, packed
//    FP32x2 texel maths), the pair kernel then reads one value per block.
#include "ctx.h"
#include "device_math.cuh"
#include "voxel_sample.cuh"

#include <cooperative_groups.h>

#include <algorithm>

using namespace drvk;
namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 128; // default CTA size; a shared-memory VPL tile holds one VPL per thread (NT)

struct GatherLight {
  const float4* vpls;   // the LIVE list (rsm.cu): 3 x float4 per VPL: (pos, area) (normal, shadow-block index) (flux, -)
  const float4* blocks; // (avgPos, distToSphereRad) per shadow block
  const uint32_t* live; // device: number of VPLs in the live list
  const uint8_t* block_live; // 1 = the block has a live VPL
  uint32_t num_vpls;    // VPLs before compaction (R^2)
  uint32_t interval;    // IndirectShadowComputationSampleInterval
};
__device__ __forceinline__ uint32_t live_vpls(const GatherLight& L) { return min(__ldg(L.live), L.num_vpls); }

struct GatherParams {
  GatherLight lights[DRV_MAX_LIGHTS];
  uint32_t num_lights;
  uint32_t granule;      // VPLs per scheduling unit (a power of two)
  uint8_t* entries;
  const drv_cache_counter* counter;
  float* partials;       // [cta][2][coef][tile cache] floats
  uint32_t shard_rank, shard_world, shard_interleave;
  uint32_t grid;         // CTAs of the gather launch (the finalize kernel needs it too)
  uint32_t fused_finalize; // 1: cooperative launch — the gather kernel adds the partial segments itself after a grid barrier
  float f0, f1, f2, f20, f22;
  // indirect shadows: the visibility table written by cone_kernel for the current chunk of caches,
  // table[(block_offset[light] + k / interval) * shadow_stride + chunk-local cache index]
  const float* shadow_table;
  uint32_t shadow_stride;
  uint32_t block_offset[DRV_MAX_LIGHTS];
  uint32_t chunk_first, chunk_cap; // the gather works on entries [chunk_first, chunk_first + chunk_cap) of the shard
  // fused all-gather: peer copies of the entries buffer (NVLink P2P)
  uint8_t* peers[8];
  uint32_t num_peers;
  uint32_t overwrite;    // 1: entry.SH = acc instead of += (sharded frames: the allocation did not clear the SH)
  uint32_t* tickets;     // warp-split kernel: one arrival counter per cache tile (zero between launches)
  unsigned long long* trace; // diagnostics (gather_variant bit 18): %globaltimer at 4 points of every CTA, else null
};
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_extra(const GatherParams& p, int k) { // words 5..7: inside the fix-up
  if (p.trace && threadIdx.x == 0) p.trace[(size_t)blockIdx.x * 8 + k] = global_ns();
}
__device__ __forceinline__ void trace_point(const GatherParams& p, int k) {
  if (p.trace && threadIdx.x == 0) {
    p.trace[(size_t)blockIdx.x * 8 + k] = global_ns();
    if (k == 0) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.trace[(size_t)blockIdx.x * 8 + 4] = smid;
    }
  }
}

struct Schedule {
  uint32_t first, count;      // contiguous shard: first entry / entries of this shard's current chunk
  uint32_t lo;                // offset of the current chunk inside the shard
  uint32_t rank, world, inter; // interleaved shard: 64-entry group g of the cell-ordered list belongs to rank g % world
  uint32_t tiles;             // cache tiles
  uint32_t units_per_tile;    // VPL units over all lights
  unsigned long long units;   // tiles * units_per_tile
};

// The part of the n active entries this rank lights, restricted to the current chunk: `count` entries with local
// indices [0, count); entry_of() maps a local index to the entry. Contiguous: the range of drv_shard_range (64-entry
// boundaries of the cell-ordered list). Interleaved: every world-th 64-entry group, so that regions of the scene
// whose cones are expensive (or cheap) are spread over all ranks.
__device__ __forceinline__ void shard_span(uint32_t n, uint32_t rank, uint32_t world, uint32_t inter, uint32_t chunk_first,
                                           uint32_t chunk_cap, uint32_t& first, uint32_t& lo_out, uint32_t& count) {
  const uint32_t groups = (n + 63u) / 64u;
  uint32_t total;
  if (inter) {
    const uint32_t owned = groups > rank ? (groups - rank + world - 1u) / world : 0u;
    total = owned * 64u;
    if (owned && (groups - 1u) % world == rank) total -= groups * 64u - n; // the ragged last group is this rank's
    first = 0;
  } else {
    const uint32_t g0 = (uint32_t)(((unsigned long long)groups * rank) / world);
    const uint32_t g1 = (uint32_t)(((unsigned long long)groups * (rank + 1)) / world);
    first = min(g0 * 64u, n);
    total = min(g1 * 64u, n) - first;
  }
  const uint32_t lo = min(total, chunk_first), hi = min(total, chunk_first + min(chunk_cap, 0xFFFFFFFFu - chunk_first));
  first += lo;
  lo_out = lo;
  count = hi - lo;
}
__device__ __forceinline__ uint32_t entry_of(uint32_t first, uint32_t lo, uint32_t rank, uint32_t world, uint32_t inter,
                                             uint32_t local) {
  if (!inter) return first + local;
  const uint32_t L = lo + local;
  return (((L >> 6) * world + rank) << 6) | (L & 63u);
}

// Same result in every thread of the gather and finalize kernels.
__device__ __forceinline__ Schedule make_schedule(const GatherParams& p, int tile_caches) {
  Schedule s;
  uint32_t n = (uint32_t)max(p.counter->TotalLightCacheCount, 0);
  s.rank = p.shard_rank; s.world = p.shard_world; s.inter = p.shard_interleave;
  shard_span(n, p.shard_rank, p.shard_world, p.shard_interleave, p.chunk_first, p.chunk_cap, s.first, s.lo, s.count);
  s.tiles = (s.count + tile_caches - 1) / tile_caches;
  uint32_t upt = 0;
  for (uint32_t l = 0; l < p.num_lights; ++l) upt += (live_vpls(p.lights[l]) + p.granule - 1) / p.granule;
  s.units_per_tile = upt;
  s.units = (unsigned long long)s.tiles * upt;
  return s;
}
__device__ __forceinline__ uint32_t entry_of(const Schedule& S, uint32_t local) {
  return entry_of(S.first, S.lo, S.rank, S.world, S.inter, local);
}
__device__ __forceinline__ unsigned long long range_begin(const Schedule& s, uint32_t grid, uint32_t cta) {
  return (s.units * cta) / grid;
}
// the CTA whose range contains unit u
__device__ __forceinline__ uint32_t owner_of(const Schedule& s, uint32_t grid, unsigned long long u) {
  return (uint32_t)(((u + 1ull) * grid - 1ull) / s.units);
}

__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---------------------------------------------------------------- cone trace
// (the record reader — VoxelVol, footprint, trilinear — lives in voxel_sample.cuh)
// One sample of the march with its loads in flight.
struct ConeSample {
  uint2 r0, r1;          // records of level l0 and l0 + 1 (r1 only when t != 0)
  float2 txy0; float tz0; // trilinear fractions in level l0
  float2 txy1; float tz1; // ... in level l0 + 1
  float t;               // mip fraction: 0 = level l0 only
  float dist, radius;
};

// cacheLightingRSM.comp:195-230 for one cache and one shadow block. Distances, step sizes and the break test
// are decision maths (the trip count equals the oracle's); positions, filtering and occlusion are continuous
// maths. The march is software-pipelined: where a cone goes next depends only on the distance travelled, not
// on what it sampled, so the fetch of step s+1 is issued before step s is filtered and the L1/L2 latency of
// the dependent chain position -> index -> load -> filter -> occlusion overlaps with useful work.
// A cone also stops once occlusion reached 1: later samples would add (1 - 1) * x = 0.
__device__ __forceinline__ float cone_trace(const VoxelVol& V, float wx, float wy, float wz, float4 blk, uint32_t& steps) {
  // :104 voxelPos, kept in level-0 texel units minus the half-texel shift (q = voxelPos * res - 0.5), so that a
  // step of `stepSize` voxels along the unit direction is q += dir * stepSize (:197-198, 213)
  const float inv_voxel = 1.0f / V.voxel_size;
  const float maxLod = (float)(V.levels - 1);
  const float kk = blk.w;
  float tx = ex_sub(blk.x, wx), ty = ex_sub(blk.y, wy), tz = ex_sub(blk.z, wz);     // :195
  const float lightDist = ex_sqrt(ex_dot3(tx, ty, tz, tx, ty, tz));                 // :196
  const float inv = 1.0f / lightDist;
  const float2 dxy = f2(tx * inv, ty * inv);                                        // :197-198 (dirInVoxel * res)
  const float dz = tz * inv;
  float2 cxy = __ffma2_rn(dxy, f2(2.0f), f2(fmaf(wx - V.vmin[0], inv_voxel, -0.5f), fmaf(wy - V.vmin[1], inv_voxel, -0.5f))); // :201
  float cz = fmaf(dz, 2.0f, fmaf(wz - V.vmin[2], inv_voxel, -0.5f));
  const float goal = ex_sub(ex_div(lightDist, V.voxel_size), 2.0f);                 // :206
  const float radToStep = ex_div(2.0f, ex_sub(1.0f, kk));                           // :209
  float dist = 0.0f;

  // advance by `stepSize` (:213-216) and issue the fetches of that sample (:219)
  auto fetch = [&](float stepSize) -> ConeSample {
    ConeSample S;
    ++steps;
    cxy = __ffma2_rn(dxy, f2(stepSize), cxy); cz = fmaf(dz, stepSize, cz);          // :213
    dist = ex_add(dist, stepSize);                                                  // :214
    S.dist = dist;
    S.radius = ex_mul(dist, kk);                                                    // :216
    S.t = 0.0f;
    int l0 = 0;
    if (S.radius > 1.0f) {
      // lod = log2(radius) clamped to the chain. radius <= 1 — the first stretch of every cone, a VAL block
      // subtends ~1/32 rad — is level 0 exactly, without the log. fmaxf(NaN, 0) = 0.
      const float l = fminf(__log2f(S.radius), maxLod);
      floor_frac_w(l - 0.5f, l0, S.t);
    }
    const Footprint f0 = footprint(V, l0, cxy, cz);
    S.r0 = __ldg(V.rec + f0.index);
    S.txy0 = f0.txy; S.tz0 = f0.tz;
    if (S.t != 0.0f) { // mip-linear: also the next coarser level
      const Footprint f1 = footprint(V, min(l0 + 1, V.levels - 1), cxy, cz);
      S.r1 = __ldg(V.rec + f1.index);
      S.txy1 = f1.txy; S.tz1 = f1.tz;
    }
    return S;
  };

  auto filter = [&](const ConeSample& S) -> float {
    float o = trilinear(S.r0, S.txy0, S.tz0);
    if (S.t != 0.0f) o = fmaf(S.t, trilinear(S.r1, S.txy1, S.tz1) - o, o);
    return o;
  };
  // two steps per trip so that the in-flight sample alternates between A and B without register copies
  float occ = 0.0f;
  ConeSample A = fetch(1.0f), B;
#pragma unroll 1
  for (int s = 0; s < 32; s += 2) {
    const bool lastA = A.dist >= goal;                                              // :222 (s <= 30 here)
    if (!lastA) B = fetch(fmaxf(1.0f, ex_mul(A.radius, radToStep)));                // :225, one step ahead
    occ = fmaf(1.0f - occ, filter(A), occ);                                         // :220
    if (lastA || occ >= 1.0f) break;
    const bool lastB = (B.dist >= goal) || (s == 30);                               // :211
    if (!lastB) A = fetch(fmaxf(1.0f, ex_mul(B.radius, radToStep)));
    occ = fmaf(1.0f - occ, filter(B), occ);
    if (lastB || occ >= 1.0f) break;
  }
  return saturatef(1.0f - occ);                                                     // :230
}

// The same march without the manual software pipeline: one sample per trip, one branch for the second mip level,
// no ping-pong state. With the cone pass on its own (8+ CTAs of 4 warps per SM) the other warps of the SM hide the
// L1/L2 latency of the dependent chain, and the loop spends a quarter fewer instructions on control flow and
// register moves (ncu: BRA + BSSY + BSYNC + MOV were 18 % of the pipelined kernel's instructions).
//
// The pass is issue-bound, so the march is written for instruction count:
//  * the position is carried as w = (level-0 texel coordinate) - 1: w + kMagic rounds to the lower corner of the
//    footprint directly, and level l sees w * 2^-l + (2^-l - 1) — one FFMA, exact for l = 0;
//  * the "+1" of the record index (records are addressed by lower corner + 1) and the start of the level are one
//    constant per level (ConeTables::rec_k);
//  * INSIDE: the record grid is padded by kVoxelRecordPad corners on every side (clamp to edge baked in), so a cone
//    whose first and last possible sample lie within the padded range needs no clamp at any level: the index is formed
//    from the raw float bits of the rounded coordinates (the 0x4B400000 bias of all three is folded into
//    ConeTables::rec_kb). A warp takes this path when all its cones are inside — nearly always: caches and VAL
//    blocks are surface points of the scene the volume encloses;
//  * an empty footprint (most of a scene's volume) filters to exactly 0 and occ = fma(1 - occ, 0, occ) is occ, bit
//    for bit, so unpack + trilinear + blend are skipped.
struct ConeTables {
  uint32_t rec_k[16];  // rec_offset[l] + (1 + pad) (1 + S + S^2),        S = (res >> l) + 1 + 2 pad (kVoxelRecordPad)
  uint32_t rec_kb[16]; // rec_k[l] - 0x4B400000 * (1 + S + S^2)  (mod 2^32)
};

template <bool INSIDE>
__device__ __forceinline__ uint32_t record_index(const ConeTables& T, int l, int r, float2 mxy, float mz) {
  const int S = r + 1 + 2 * (int)kVoxelRecordPad;
  if (INSIDE) // raw bits: 0x4B400000 + corner each; the bias is in rec_kb (unsigned: the sum wraps by design)
    return __float_as_uint(mxy.x) + (uint32_t)S * (__float_as_uint(mxy.y) + (uint32_t)S * __float_as_uint(mz)) + T.rec_kb[l];
  const int x = min(max(__float_as_int(mxy.x) - 0x4B400000, -1), r - 1);
  const int y = min(max(__float_as_int(mxy.y) - 0x4B400000, -1), r - 1);
  const int z = min(max(__float_as_int(mz) - 0x4B400000, -1), r - 1);
  return (uint32_t)(x + S * (y + S * z)) + T.rec_k[l];
}
// fractions of w around its rounded value m: (w - (m - kMagic)) + 0.5
__device__ __forceinline__ float2 frac_of2(float2 w, float2 m) {
  return __fadd2_rn(__fadd2_rn(w, __fadd2_rn(f2(kMagic), f2(-m.x, -m.y))), f2(0.5f));
}
__device__ __forceinline__ float frac_of(float w, float m) { return (w - (m - kMagic)) + 0.5f; }

template <bool INSIDE>
__device__ __forceinline__ float cone_march(const VoxelVol& V, const ConeTables& T, float2 dxy, float dz, float2 wxy,
                                            float wz, float kk, float goal, float radToStep, uint32_t& steps) {
  const float maxLod = (float)(V.levels - 1);
  float dist = 0.0f, occ = 0.0f, stepSize = 1.0f;
  int s = 0;
  // First stretch: while the sphere radius is <= 1 voxel, log2(radius) <= 0 selects mip level 0 alone
  // and the step stays max(1, r * g): no log2, no second level, no level arithmetic. A VAL block subtends ~1/32
  // rad (SuperValWidth), so this loop is where nearly every sample of a frame is taken.
#pragma unroll 1
  for (; s < 32; ++s) {
    const float nd = ex_add(dist, stepSize);
    const float radius = ex_mul(nd, kk);
    if (radius > 1.0f) break;                                                       // continue in the general loop
    wxy = __ffma2_rn(dxy, f2(stepSize), wxy); wz = fmaf(dz, stepSize, wz);          // :213
    dist = nd;                                                                      // :214
    ++steps;
    const float2 mxy = __fadd2_rn(wxy, f2(kMagic));
    const float mz = wz + kMagic;
    const uint2 r0 = __ldg(V.rec + record_index<INSIDE>(T, 0, V.res, mxy, mz));
    if ((r0.x | r0.y) != 0u) {
      occ = fmaf(1.0f - occ, trilinear(r0, frac_of2(wxy, mxy), frac_of(wz, mz)), occ); // :220
      if (occ >= 1.0f) return saturatef(1.0f - occ);                                // :222
    }
    if (dist >= goal) return saturatef(1.0f - occ);                                 // :222
    stepSize = fmaxf(1.0f, ex_mul(radius, radToStep));                              // :225
  }
#pragma unroll 1
  for (; s < 32; ++s) {                                                             // :211
    wxy = __ffma2_rn(dxy, f2(stepSize), wxy); wz = fmaf(dz, stepSize, wz);          // :213
    dist = ex_add(dist, stepSize);                                                  // :214
    ++steps;
    const float radius = ex_mul(dist, kk);                                          // :216
    // lod = log2(radius) clamped to the chain; radius <= 1 is level 0 exactly. fmaxf(NaN, 0) = 0.
    const float l = fminf(fmaxf(__log2f(radius), 0.0f), maxLod);
    int l0; float t;
    floor_frac_w(l - 0.5f, l0, t);
    if (!(radius > 1.0f)) { l0 = 0; t = 0.0f; }
    const float sc0 = __int_as_float(0x3f800000 - (l0 << 23));                      // 2^-l0
    const float2 w0xy = __ffma2_rn(wxy, f2(sc0), f2(sc0 - 1.0f));
    const float w0z = fmaf(wz, sc0, sc0 - 1.0f);
    const float2 m0xy = __fadd2_rn(w0xy, f2(kMagic));
    const float m0z = w0z + kMagic;
    const uint2 r0 = __ldg(V.rec + record_index<INSIDE>(T, l0, V.res >> l0, m0xy, m0z));
    float o = 0.0f;
    if (t != 0.0f) { // mip-linear: also the next coarser level (:219)
      const int l1 = min(l0 + 1, V.levels - 1);
      const float sc1 = __int_as_float(0x3f800000 - (l1 << 23));
      const float2 w1xy = __ffma2_rn(wxy, f2(sc1), f2(sc1 - 1.0f));
      const float w1z = fmaf(wz, sc1, sc1 - 1.0f);
      const float2 m1xy = __fadd2_rn(w1xy, f2(kMagic));
      const float m1z = w1z + kMagic;
      const uint2 r1 = __ldg(V.rec + record_index<INSIDE>(T, l1, V.res >> l1, m1xy, m1z));
      if ((r0.x | r0.y | r1.x | r1.y) != 0u) { // else both footprints empty: 0 + t * (0 - 0), exactly
        o = trilinear(r0, frac_of2(w0xy, m0xy), frac_of(w0z, m0z));
        o = fmaf(t, trilinear(r1, frac_of2(w1xy, m1xy), frac_of(w1z, m1z)) - o, o);
      }
    } else if ((r0.x | r0.y) != 0u) {
      o = trilinear(r0, frac_of2(w0xy, m0xy), frac_of(w0z, m0z));
    }
    occ = fmaf(1.0f - occ, o, occ);                                                 // :220
    if (dist >= goal || occ >= 1.0f) break;                                         // :222
    stepSize = fmaxf(1.0f, ex_mul(radius, radToStep));                              // :225
  }
  return saturatef(1.0f - occ);                                                     // :230
}

__device__ __forceinline__ float cone_trace_simple(const VoxelVol& V, const ConeTables& T, float wx, float wy, float wz,
                                                   float4 blk, uint32_t& steps) {
  const float inv_voxel = 1.0f / V.voxel_size;
  const float kk = blk.w;
  float tx = ex_sub(blk.x, wx), ty = ex_sub(blk.y, wy), tz = ex_sub(blk.z, wz);     // :195
  const float lightDist = ex_sqrt(ex_dot3(tx, ty, tz, tx, ty, tz));                 // :196
  const float inv = 1.0f / lightDist;
  const float2 dxy = f2(tx * inv, ty * inv);                                        // :197-198 (dirInVoxel * res)
  const float dz = tz * inv;
  // the cache in level-0 texel units, minus one (see above); the march starts two voxels along the cone (:201)
  const float sx = fmaf(wx - V.vmin[0], inv_voxel, -1.0f), sy = fmaf(wy - V.vmin[1], inv_voxel, -1.0f);
  const float sz = fmaf(wz - V.vmin[2], inv_voxel, -1.0f);
  const float2 cxy = __ffma2_rn(dxy, f2(2.0f), f2(sx, sy));
  const float cz = fmaf(dz, 2.0f, sz);
  const float distVox = ex_div(lightDist, V.voxel_size);
  const float goal = ex_sub(distVox, 2.0f);                                         // :206
  const float radToStep = ex_div(2.0f, ex_sub(1.0f, kk));                           // :209
  // Inside test. Every sample lies on the segment from the start position to less than one step beyond the light
  // (the march starts two voxels in and ends at the first distance >= lightDist - 2), and no step is longer than
  // max(1, lightDist * kk * radToStep). The record grid holds the lower corners [-1 - pad, res - 1 + pad] (ctx.h), a
  // corner is round(w), and level l sees (w + 1) 2^-l - 1: both ends of the segment within that range, less half
  // a voxel for the rounding => every corner of every level is. NaN / inf fail the comparisons.
  const float over = fmaxf(1.0f, distVox * kk * radToStep);
  const float lo = -0.5f - (float)kVoxelRecordPad, hi = (float)V.res - 1.5f + (float)kVoxelRecordPad;
  const float ex = fmaf(dxy.x, distVox + over, sx), ey = fmaf(dxy.y, distVox + over, sy), ez = fmaf(dz, distVox + over, sz);
  const bool inside = cxy.x >= lo && cxy.x <= hi && cxy.y >= lo && cxy.y <= hi && cz >= lo && cz <= hi &&
                      ex >= lo && ex <= hi && ey >= lo && ey <= hi && ez >= lo && ez <= hi;
  if (__all_sync(__activemask(), inside))
    return cone_march<true>(V, T, dxy, dz, cxy, cz, kk, goal, radToStep, steps);
  return cone_march<false>(V, T, dxy, dz, cxy, cz, kk, goal, radToStep, steps);
}

// ------------------------------------------------------------ epilogue helpers
template <int ORDER>
constexpr int num_coefs() { return ORDER == 2 ? 27 : 12; }

// Raw moments (a0,ax,ay,az,xy,yz,zz,xz,dd; rgb each) -> the values to ADD at
// float offsets 4.. of the entry, i.e. in lightcache.glsl:33-57 order.
template <int ORDER>
__device__ __forceinline__ void coef_values(const GatherParams& p, const float* raw, float* out /*12 or 28*/) {
  const float* a0 = raw; const float* ax = raw + 3; const float* ay = raw + 6; const float* az = raw + 9;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    out[0 + c] = -p.f1 * ay[c];  // SH1neg1 -= (f1 * t.y) * rad   :268
    out[4 + c] = p.f1 * az[c];   // SH10    += (f1 * t.z) * rad   :269
    out[8 + c] = -p.f1 * ax[c];  // SH1pos1 -= (f1 * t.x) * rad   :270
  }
  out[3] = p.f0 * a0[0]; out[7] = p.f0 * a0[1]; out[11] = p.f0 * a0[2]; // SH00 :267
  if (ORDER == 2) {
    const float* xy = raw + 12; const float* yz = raw + 15; const float* zz = raw + 18;
    const float* xz = raw + 21; const float* dd = raw + 24;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      out[12 + c] = -p.f2 * xy[c];  // SH2neg2 :273
      out[16 + c] = p.f2 * yz[c];   // SH2neg1 :274
      out[20 + c] = p.f2 * xz[c];   // SH2pos1 :276
      out[24 + c] = p.f22 * dd[c];  // SH2pos2 :277
    }
    out[15] = p.f20 * (3.0f * zz[0] - a0[0]); // SH20 :275 = sum (3 z^2 - 1) rad
    out[19] = p.f20 * (3.0f * zz[1] - a0[1]);
    out[23] = p.f20 * (3.0f * zz[2] - a0[2]);
    out[27] = 0.0f;
  }
}

template <int ORDER>
__device__ __forceinline__ void add_to_entry(const GatherParams& p, uint32_t entry, const float* vals) {
  constexpr int STRIDE = ORDER == 2 ? 128 : 64;
  constexpr int NQ = ORDER == 2 ? 7 : 3;
  float4* e = reinterpret_cast<float4*>(p.entries + (size_t)entry * STRIDE);
  float4 outq[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    float4 o = p.overwrite ? make_float4(0.f, 0.f, 0.f, 0.f) : e[1 + q];
    o.x += vals[q * 4 + 0]; o.y += vals[q * 4 + 1]; o.z += vals[q * 4 + 2]; o.w += vals[q * 4 + 3];
    e[1 + q] = o; // entry.SH += acc, :358-373
    outq[q] = o;
  }
  // fused all-gather: the finished entry goes straight to every peer's copy over NVLink
  for (uint32_t r = 0; r < p.num_peers; ++r) {
    if (!p.peers[r]) continue;
    float4* pe = reinterpret_cast<float4*>(p.peers[r] + (size_t)entry * STRIDE);
#pragma unroll
    for (int q = 0; q < NQ; ++q) pe[1 + q] = outq[q];
  }
}

// ------------------------------------------------------------ scalar pair maths
template <int ORDER>
struct Acc {
  float a0[3], ax[3], ay[3], az[3];
  float xy[3], yz[3], zz[3], xz[3], dd[3]; // used only when ORDER == 2
};

// One cache x VPL pair, cacheLightingRSM.comp:249-277 with the basis
// constants factored out: s = sat(dot(N,-t^))*shadow / (d^2 + A);
// a0 += F s; a{x,y,z} += F s t^; (SH2) second moments of t^.
template <int ORDER, bool SHADOW>
__device__ __forceinline__ void pair_eval(Acc<ORDER>& A, float px, float py, float pz, float4 va, float4 vb, float4 vc,
                                          float shadow) {
  float tx = va.x - px, ty = va.y - py, tz = va.z - pz;      // :249
  float d2 = fmaf(tz, tz, fmaf(ty, ty, tx * tx));            // :252
  float inv = rsqrt_approx(d2);                              // :253
  float cr = fmaf(vb.z, tz, fmaf(vb.y, ty, vb.x * tx));      // dot(N, t)
  float cosv = __saturatef(-cr * inv);                       // :256
  float s = cosv * rcp_approx(d2 + va.w);                    // :262
  if (SHADOW) s *= shadow;                                   // :258
  float w = s * inv;
  float ux = tx * w, uy = ty * w, uz = tz * w;               // s * t^
  const float fl[3] = {vc.x, vc.y, vc.z};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    A.a0[c] = fmaf(fl[c], s, A.a0[c]);
    A.ax[c] = fmaf(fl[c], ux, A.ax[c]);
    A.ay[c] = fmaf(fl[c], uy, A.ay[c]);
    A.az[c] = fmaf(fl[c], uz, A.az[c]);
  }
  if (ORDER == 2) {
    float nx = tx * inv, ny = ty * inv, nz = tz * inv;       // t^
    float qxy = nx * uy, qyz = ny * uz, qzz = nz * uz, qxz = nx * uz;
    float qdd = fmaf(nx, ux, -ny * uy);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      A.xy[c] = fmaf(fl[c], qxy, A.xy[c]);
      A.yz[c] = fmaf(fl[c], qyz, A.yz[c]);
      A.zz[c] = fmaf(fl[c], qzz, A.zz[c]);
      A.xz[c] = fmaf(fl[c], qxz, A.xz[c]);
      A.dd[c] = fmaf(fl[c], qdd, A.dd[c]);
    }
  }
}

// Policy: CPT caches per thread, scalar FP32 maths, 3 float4 per VPL in shared memory.
template <int ORDER, bool SHADOW, int CPT_>
struct ScalarMath {
  static constexpr int CPT = CPT_;
  static constexpr int kSmemPerVpl = 3;
  float px[CPT], py[CPT], pz[CPT];
  float shadow[CPT];
  bool live[CPT];
  Acc<ORDER> A[CPT];

  __device__ __forceinline__ void begin(int j, bool alive, float4 pos) {
    live[j] = alive;
    px[j] = pos.x; py[j] = pos.y; pz[j] = pos.z;
    shadow[j] = 1.0f; // :132
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      A[j].a0[c] = A[j].ax[c] = A[j].ay[c] = A[j].az[c] = 0.0f;
      A[j].xy[c] = A[j].yz[c] = A[j].zz[c] = A[j].xz[c] = A[j].dd[c] = 0.0f;
    }
  }
  static __device__ __forceinline__ void stage(float4* slot, float4 r0, float4 r1, float4 r2) {
    slot[0] = r0; slot[1] = r1; slot[2] = r2;
  }
  // CPT cones (one per cache of this thread) x up to NB consecutive shadow blocks, marched in lock step
  __device__ __forceinline__ void set_shadow(const float (&v)[CPT]) {
#pragma unroll
    for (int j = 0; j < CPT; ++j) shadow[j] = v[j];
  }
  __device__ __forceinline__ void eval(const float4* v) {
    float4 va = v[0], vb = v[1], vc = v[2];
#pragma unroll
    for (int j = 0; j < CPT; ++j) pair_eval<ORDER, SHADOW>(A[j], px[j], py[j], pz[j], va, vb, vc, shadow[j]);
  }
  __device__ __forceinline__ void raw(int j, float* r) const {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      r[c] = A[j].a0[c]; r[3 + c] = A[j].ax[c]; r[6 + c] = A[j].ay[c]; r[9 + c] = A[j].az[c];
      r[12 + c] = A[j].xy[c]; r[15 + c] = A[j].yz[c]; r[18 + c] = A[j].zz[c];
      r[21 + c] = A[j].xz[c]; r[24 + c] = A[j].dd[c];
    }
  }
};

// ------------------------------------------------------------ packed (FFMA2) pair maths
// Policy: PAIRS x 2 caches per thread; the two caches of a pair ride in the
// .x/.y halves of 64-bit registers. The shared-memory VPL record duplicates
// every scalar into both halves so it feeds FFMA2 without register shuffles:
//  q0 = (x,x,y,y) q1 = (z,z,area,area) q2 = (nx,nx,ny,ny) q3 = (nz,nz,fr,fr) q4 = (fg,fg,fb,fb)
// BCAST: the shared-memory record is the plain 48-byte VPL (three float4) and every per-VPL scalar enters the packed
// instructions as a 32-bit BROADCAST operand — sm_100's FFMA2 / FADD2 / FMUL2 take `R.F32` next to `R.F32x2` (nvcc
// emits it for make_float2(s, s)). Against the duplicated record: 3 instead of 5 LDS.128 per VPL, 10 instead of 20
// registers per VPL in flight, and 19 of the 27 packed instructions of a pair read five registers instead of six.
template <int ORDER, bool SHADOW, int PAIRS, bool BCAST = false>
struct PackedMath {
  static constexpr int CPT = PAIRS * 2;
  static constexpr int kSmemPerVpl = BCAST ? 3 : 5;
  float2 npx[PAIRS], npy[PAIRS], npz[PAIRS]; // NEGATED cache positions
  float2 shadow[PAIRS];
  bool live[CPT];
  float2 a0[PAIRS][3], ax[PAIRS][3], ay[PAIRS][3], az[PAIRS][3];
  float2 xy[PAIRS][3], yz[PAIRS][3], zz[PAIRS][3], xz[PAIRS][3], dd[PAIRS][3];

  __device__ __forceinline__ void begin(int j, bool alive, float4 pos) {
    live[j] = alive;
    const int pr = j >> 1;
    if (j & 1) { npx[pr].y = -pos.x; npy[pr].y = -pos.y; npz[pr].y = -pos.z; shadow[pr].y = 1.0f; }
    else       { npx[pr].x = -pos.x; npy[pr].x = -pos.y; npz[pr].x = -pos.z; shadow[pr].x = 1.0f; }
    const float2 z = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      a0[pr][c] = ax[pr][c] = ay[pr][c] = az[pr][c] = z;
      xy[pr][c] = yz[pr][c] = zz[pr][c] = xz[pr][c] = dd[pr][c] = z;
    }
  }
  static __device__ __forceinline__ void stage(float4* d, float4 r0, float4 r1, float4 r2) {
    if (BCAST) { d[0] = r0; d[1] = r1; d[2] = r2; return; }
    d[0] = make_float4(r0.x, r0.x, r0.y, r0.y);
    d[1] = make_float4(r0.z, r0.z, r0.w, r0.w);
    d[2] = make_float4(r1.x, r1.x, r1.y, r1.y);
    d[3] = make_float4(r1.z, r1.z, r2.x, r2.x);
    d[4] = make_float4(r2.y, r2.y, r2.z, r2.z);
  }
  __device__ __forceinline__ void set_shadow(const float (&v)[CPT]) {
#pragma unroll
    for (int j = 0; j < PAIRS; ++j) { shadow[j].x = v[2 * j]; shadow[j].y = v[2 * j + 1]; }
  }
  __device__ __forceinline__ void eval(const float4* q) {
    float2 vx, vy, vz, ar, nx, ny, nz, fl[3];
    if (BCAST) {
      const float4 r0 = q[0], r1 = q[1], r2 = q[2]; // (x, y, z, area) (nx, ny, nz, block) (fr, fg, fb, -)
      vx = make_float2(r0.x, r0.x); vy = make_float2(r0.y, r0.y); vz = make_float2(r0.z, r0.z);
      ar = make_float2(r0.w, r0.w);
      nx = make_float2(r1.x, r1.x); ny = make_float2(r1.y, r1.y); nz = make_float2(r1.z, r1.z);
      fl[0] = make_float2(r2.x, r2.x); fl[1] = make_float2(r2.y, r2.y); fl[2] = make_float2(r2.z, r2.z);
    } else {
      const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4];
      vx = make_float2(q0.x, q0.y); vy = make_float2(q0.z, q0.w); vz = make_float2(q1.x, q1.y);
      ar = make_float2(q1.z, q1.w);
      nx = make_float2(q2.x, q2.y); ny = make_float2(q2.z, q2.w); nz = make_float2(q3.x, q3.y);
      fl[0] = make_float2(q3.z, q3.w); fl[1] = make_float2(q4.x, q4.y); fl[2] = make_float2(q4.z, q4.w);
    }
#pragma unroll
    for (int j = 0; j < PAIRS; ++j) {
      float2 tx = __fadd2_rn(vx, npx[j]), ty = __fadd2_rn(vy, npy[j]), tz = __fadd2_rn(vz, npz[j]);
      float2 d2 = __ffma2_rn(tz, tz, __ffma2_rn(ty, ty, __fmul2_rn(tx, tx)));
      float2 inv = make_float2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
      float2 cr = __ffma2_rn(nz, tz, __ffma2_rn(ny, ty, __fmul2_rn(nx, tx)));
      float2 cosv = make_float2(__saturatef(-cr.x * inv.x), __saturatef(-cr.y * inv.y));
      float2 den = __fadd2_rn(d2, ar);
      float2 s = __fmul2_rn(cosv, make_float2(rcp_approx(den.x), rcp_approx(den.y)));
      if (SHADOW) s = __fmul2_rn(s, shadow[j]);
      float2 w = __fmul2_rn(s, inv);
      float2 ux = __fmul2_rn(tx, w), uy = __fmul2_rn(ty, w), uz = __fmul2_rn(tz, w);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        a0[j][c] = __ffma2_rn(fl[c], s, a0[j][c]);
        ax[j][c] = __ffma2_rn(fl[c], ux, ax[j][c]);
        ay[j][c] = __ffma2_rn(fl[c], uy, ay[j][c]);
        az[j][c] = __ffma2_rn(fl[c], uz, az[j][c]);
      }
      if (ORDER == 2) {
        float2 hx = __fmul2_rn(tx, inv), hy = __fmul2_rn(ty, inv), hz = __fmul2_rn(tz, inv);
        float2 qxy = __fmul2_rn(hx, uy), qyz = __fmul2_rn(hy, uz), qzz = __fmul2_rn(hz, uz), qxz = __fmul2_rn(hx, uz);
        float2 nhy = make_float2(-hy.x, -hy.y);
        float2 qdd = __ffma2_rn(hx, ux, __fmul2_rn(nhy, uy));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          xy[j][c] = __ffma2_rn(fl[c], qxy, xy[j][c]);
          yz[j][c] = __ffma2_rn(fl[c], qyz, yz[j][c]);
          zz[j][c] = __ffma2_rn(fl[c], qzz, zz[j][c]);
          xz[j][c] = __ffma2_rn(fl[c], qxz, xz[j][c]);
          dd[j][c] = __ffma2_rn(fl[c], qdd, dd[j][c]);
        }
      }
    }
  }
  __device__ __forceinline__ void raw(int j, float* r) const {
    const int pr = j >> 1;
    const bool hi = (j & 1) != 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#define DRV_H(v) (hi ? (v).y : (v).x)
      r[c] = DRV_H(a0[pr][c]); r[3 + c] = DRV_H(ax[pr][c]); r[6 + c] = DRV_H(ay[pr][c]); r[9 + c] = DRV_H(az[pr][c]);
      r[12 + c] = DRV_H(xy[pr][c]); r[15 + c] = DRV_H(yz[pr][c]); r[18 + c] = DRV_H(zz[pr][c]);
      r[21 + c] = DRV_H(xz[pr][c]); r[24 + c] = DRV_H(dd[pr][c]);
#undef DRV_H
    }
  }
};

// ------------------------------------------------------------ TMA helpers (variant 1)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------ finalize: add the partial segments in VPL order
// One block iteration per 32 consecutive caches of a tile (block-stride loop: the cache count lives on the device).
// The CTAs that own pieces of the tile are the same for all 32 caches, so their list is derived once per
// chunk. Thread (y, x) then sums, for cache x, all coefficients over the owners y, y + 8, y + 16, ... — every
// load of a thread is independent, so a tile split between ~25 CTAs costs one or two L2 round trips instead of
// a serial chain — and the eight slices are added in a fixed order through shared memory (deterministic — no
// float atomics). 32 threads apply the SH factors and add the result into the entries with 128-bit accesses
// (+ the peer stores of the fused all-gather).
constexpr int kFinChunk = 32;

template <int ORDER, int NTHREADS>
struct FinalizeSmem {
  static constexpr int SLICES = NTHREADS / kFinChunk;
  float part[SLICES][num_coefs<ORDER>()][kFinChunk];
  uint32_t list[NTHREADS];
};

// `first_block` / `num_blocks`: this block's index and the number of blocks that share the chunk loop.
template <int ORDER, int NTHREADS>
__device__ __forceinline__ void finalize_phase(const GatherParams& p, int tile_caches, FinalizeSmem<ORDER, NTHREADS>& sm,
                                               uint32_t first_block, uint32_t num_blocks) {
  constexpr int NC = num_coefs<ORDER>();
  constexpr int SLICES = NTHREADS / kFinChunk;
  static_assert(NTHREADS % kFinChunk == 0 && SLICES >= 1, "one cache per lane, NTHREADS / 32 owner slices");
  const Schedule S = make_schedule(p, tile_caches);
  if (S.units == 0) return;
  const uint32_t G = p.grid;
  const uint32_t chunks = (S.count + kFinChunk - 1) / kFinChunk;
  const uint32_t x = threadIdx.x % kFinChunk, y = threadIdx.x / kFinChunk;
  for (uint32_t chunk = first_block; chunk < chunks; chunk += num_blocks) {
    const uint32_t local0 = chunk * kFinChunk;
    const uint32_t tile = local0 / tile_caches, in_tile0 = local0 - tile * tile_caches;
    const unsigned long long ua = (unsigned long long)tile * S.units_per_tile, ub = ua + S.units_per_tile;
    const uint32_t c_lo = owner_of(S, G, ua), c_hi = owner_of(S, G, ub - 1);
    if (c_lo == c_hi) continue; // one CTA covered the whole tile and already wrote it (block-uniform)
    float acc[NC];
#pragma unroll
    for (int q = 0; q < NC; ++q) acc[q] = 0.0f;
    for (uint32_t cbase = c_lo; cbase <= c_hi; cbase += NTHREADS) {
      const uint32_t c = cbase + threadIdx.x;
      uint32_t entry = 0xFFFFFFFFu;
      if (c <= c_hi) {
        const unsigned long long cb = range_begin(S, G, c), ce = range_begin(S, G, c + 1);
        // slot 0 = the CTA's range starts inside this tile (only c_lo can start before it)
        if (cb < ce) entry = c * 2u + (cb >= ua ? 0u : 1u);
      }
      __syncthreads(); // the previous batch has been consumed
      sm.list[threadIdx.x] = entry;
      __syncthreads();
      const int cnt = (int)min((uint32_t)NTHREADS, c_hi - cbase + 1u);
#pragma unroll 2
      for (int i = (int)y; i < cnt; i += SLICES) {
        const uint32_t e = sm.list[i];
        if (e == 0xFFFFFFFFu) continue;
        const float* src = p.partials + (size_t)e * NC * tile_caches + in_tile0 + x;
#pragma unroll
        for (int q = 0; q < NC; ++q) acc[q] += __ldcg(src + (size_t)q * tile_caches);
      }
    }
    __syncthreads(); // previous chunk's readers are done with sm.part
#pragma unroll
    for (int q = 0; q < NC; ++q) sm.part[y][q][x] = acc[q];
    __syncthreads();
    // slice-sum in ascending slice order: item = (coefficient, cache)
    for (int item = threadIdx.x; item < NC * kFinChunk; item += NTHREADS) {
      const int q = item / kFinChunk, cx = item % kFinChunk;
      float v = sm.part[0][q][cx];
#pragma unroll
      for (int sl = 1; sl < SLICES; ++sl) v += sm.part[sl][q][cx];
      sm.part[0][q][cx] = v;
    }
    __syncthreads();
    if (threadIdx.x < kFinChunk && local0 + threadIdx.x < S.count) {
      float raw[27];
#pragma unroll
      for (int q = 0; q < 27; ++q) raw[q] = q < NC ? sm.part[0][q < NC ? q : 0][threadIdx.x] : 0.0f;
      float vals[28];
      coef_values<ORDER>(p, raw, vals);
      add_to_entry<ORDER>(p, entry_of(S, local0 + threadIdx.x), vals);
    }
  }
}

constexpr int kFinThreads = 256;
template <int ORDER>
__global__ void __launch_bounds__(kFinThreads) gather_finalize_kernel(GatherParams p, int tile_caches) {
  __shared__ FinalizeSmem<ORDER, kFinThreads> sm;
  finalize_phase<ORDER, kFinThreads>(p, tile_caches, sm, blockIdx.x, gridDim.x);
}

// ------------------------------------------------------------ the gather kernel
// A cursor over the shared-memory tiles of a CTA's unit range: runs of
// consecutive units that share (cache tile, light) are contiguous VPL ranges.
struct Cursor {
  unsigned long long u, u_next; // first unit of the current run / of the next run
  uint32_t tile, light;
  uint32_t base, v_end;         // current shared-memory tile starts at VPL `base`; the run ends at v_end
  bool valid;
};
__device__ __forceinline__ void start_run(const GatherParams& p, const Schedule& S, Cursor& c, unsigned long long u,
                                          unsigned long long u1) {
  c.u = u;
  c.valid = u < u1;
  if (!c.valid) return;
  c.tile = (uint32_t)(u / S.units_per_tile);
  uint32_t j = (uint32_t)(u - (unsigned long long)c.tile * S.units_per_tile), l = 0, ul = 0, nv = 0;
  for (;; ++l) {
    nv = live_vpls(p.lights[l]);
    ul = (nv + p.granule - 1) / p.granule;
    if (j < ul || l + 1 >= p.num_lights) break;
    j -= ul;
  }
  c.light = l;
  const unsigned long long tile_end = (unsigned long long)(c.tile + 1) * S.units_per_tile;
  const unsigned long long avail = (u1 < tile_end ? u1 : tile_end) - u;
  const uint32_t take = (uint32_t)min((unsigned long long)(ul - j), avail);
  c.base = j * p.granule;
  c.v_end = min((j + take) * p.granule, nv);
  c.u_next = u + take;
}
__device__ __forceinline__ void advance(const GatherParams& p, const Schedule& S, Cursor& c, unsigned long long u1,
                                        uint32_t vpl_tile) {
  c.base += vpl_tile;
  if (c.base >= c.v_end) start_run(p, S, c, c.u_next, u1);
}

// SHADOW: every pair is scaled by the visibility of its (cache, VAL block), read from the table cone_kernel wrote.
template <int ORDER, bool SHADOW, typename Math, bool USE_TMA, int NT, int UNROLL>
__device__ __forceinline__ void gather_main(const GatherParams& p) {
  constexpr int CPT = Math::CPT;
  constexpr int TILE = NT * CPT;
  constexpr int kVplTile = NT; // VPLs staged per shared-memory tile: one per thread
  constexpr int NC = num_coefs<ORDER>();
  constexpr int STRIDE = ORDER == 2 ? 128 : 64;
  constexpr int SPV = Math::kSmemPerVpl;
  constexpr int STAGES = USE_TMA ? 2 : 1;
  __shared__ __align__(128) float4 s_vpl[STAGES][kVplTile * SPV];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ uint32_t s_blk[SHADOW ? kVplTile : 1]; // shadow-block index of every staged VPL

  const Schedule S = make_schedule(p, TILE);
  if (S.units == 0) return;
  const unsigned long long u0 = range_begin(S, gridDim.x, blockIdx.x), u1 = range_begin(S, gridDim.x, blockIdx.x + 1);
  if (u0 >= u1) return;
  uint32_t phase[2] = {0u, 0u};
  if (USE_TMA) {
    if (threadIdx.x == 0) {
      mbar_init(&s_bar[0], 1);
      mbar_init(&s_bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }

  Math M;
  float4 r0, r1, r2; // register-prefetched VPL of the NEXT shared-memory tile
  r0 = r1 = r2 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto prefetch = [&](const Cursor& c) {
    const GatherLight& L = p.lights[c.light];
    uint32_t v = c.base + threadIdx.x;
    bool ok = v < c.v_end;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    r0 = ok ? __ldg(L.vpls + (size_t)v * 3) : make_float4(0.f, 0.f, 0.f, 1.f);
    r1 = ok ? __ldg(L.vpls + (size_t)v * 3 + 1) : z;
    r2 = ok ? __ldg(L.vpls + (size_t)v * 3 + 2) : z;
  };
  auto tma_issue = [&](const Cursor& c, int st) {
    const GatherLight& L = p.lights[c.light];
    uint32_t n = min((uint32_t)kVplTile, c.v_end - c.base);
    mbar_expect_tx(&s_bar[st], n * 48u);
    tma_bulk_g2s(&s_vpl[st][0], L.vpls + (size_t)c.base * 3, n * 48u, &s_bar[st]);
  };

  Cursor cur;
  start_run(p, S, cur, u0, u1);
  if (USE_TMA) {
    if (threadIdx.x == 0) tma_issue(cur, 0);
  } else {
    prefetch(cur);
  }
  bool seg_open = false;
  uint32_t seg_first_j = 0;
  uint32_t step = 0;
  while (cur.valid) {
    if (!seg_open) { // (re)load this thread's caches and clear the accumulators
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        uint32_t local = cur.tile * TILE + j * NT + threadIdx.x;
        bool alive = local < S.count;
        float4 pos = alive ? *reinterpret_cast<const float4*>(p.entries + (size_t)entry_of(S, local) * STRIDE)
                           : make_float4(1e30f, 1e30f, 1e30f, 0.f); // :83
        M.begin(j, alive, pos);
      }
      seg_first_j = (uint32_t)(cur.u - (unsigned long long)cur.tile * S.units_per_tile);
      seg_open = true;
    }
    Cursor nxt = cur;
    advance(p, S, nxt, u1, kVplTile);
    const int st = USE_TMA ? (int)(step & 1u) : 0;
    if (USE_TMA) {
      if (threadIdx.x == 0 && nxt.valid) tma_issue(nxt, st ^ 1); // stage st^1 was released by the last __syncthreads
      mbar_wait(&s_bar[st], phase[st]);
      phase[st] ^= 1u;
    } else {
      __syncthreads(); // previous tile fully consumed
      Math::stage(&s_vpl[0][threadIdx.x * SPV], r0, r1, r2);
      if (SHADOW) s_blk[threadIdx.x] = __float_as_uint(r1.w);
      __syncthreads();
      if (nxt.valid) prefetch(nxt);
    }
    const int n = (int)min((uint32_t)kVplTile, cur.v_end - cur.base);
    const float4* sv = &s_vpl[st][0];
    if constexpr (SHADOW) {
      // :169 — a new shadow value every `interval` VPLs: one table read per cache per block. The
      // live list keeps the VPL order, so the VPLs of a block are still consecutive; each carries its block index.
      const float* col = p.shadow_table + (size_t)p.block_offset[cur.light] * p.shadow_stride + cur.tile * TILE + threadIdx.x;
      auto load_shadow = [&](uint32_t blk) {
        float v[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          v[j] = (cur.tile * TILE + j * NT + threadIdx.x < S.count) ? __ldg(col + (size_t)blk * p.shadow_stride + j * NT) : 0.0f;
        M.set_shadow(v);
      };
      if constexpr (USE_TMA) {
        uint32_t cur_blk = 0xFFFFFFFFu;
#pragma unroll UNROLL
        for (int i = 0; i < n; ++i) {
          const uint32_t blk = __float_as_uint(sv[i * SPV + 1].w);
          if (blk != cur_blk) { cur_blk = blk; load_shadow(blk); } // warp-uniform
          M.eval(sv + i * SPV);
        }
      } else {
        // Runs of VPLs that share a block, found once per 32 VPLs with a ballot (every warp derives the same masks
        // from shared memory): the pair loop of a run is then the same branch-free, unrolled loop as without
        // shadows — a block test inside it (one LDS + compare + branch per VPL) cost the shadowed kernel its
        // instruction-level parallelism across VPLs.
        const int lane = threadIdx.x & 31;
#pragma unroll 1
        for (int c0 = 0; c0 < n; c0 += 32) {
          const int idx = c0 + lane;
          const bool start = idx < n && (lane == 0 || s_blk[idx] != s_blk[idx - 1]);
          uint32_t m = __ballot_sync(0xffffffffu, start); // bit 0 is always set: a run may continue from the last chunk
          const int c1 = min(n, c0 + 32);
          int i = c0;
          while (m) {
            m &= m - 1u;                                   // drop this run's start bit
            const int next = m ? c0 + (__ffs(m) - 1) : c1; // the next run's start, or the end of the chunk
            load_shadow(s_blk[i]);
#pragma unroll UNROLL
            for (; i < next; ++i) M.eval(sv + i * SPV);
          }
        }
      }
    } else {
#pragma unroll UNROLL
      for (int i = 0; i < n; ++i) M.eval(sv + i * SPV);
    }
    if (USE_TMA) __syncthreads(); // stage st consumed by every warp
    ++step;
    if (!nxt.valid || nxt.tile != cur.tile) { // segment end
      const uint32_t last_j = (uint32_t)(cur.u_next - 1ull - (unsigned long long)cur.tile * S.units_per_tile);
      const bool full = seg_first_j == 0 && last_j == S.units_per_tile - 1;
      // partial slot 0: the segment in the tile that contains u0; slot 1: a later (the last) one
      const int slot = (cur.tile == (uint32_t)(u0 / S.units_per_tile)) ? 0 : 1;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        uint32_t in_tile = j * NT + threadIdx.x;
        uint32_t local = cur.tile * TILE + in_tile;
        float raw[27];
        M.raw(j, raw);
        if (full) {
          if (local < S.count) {
            float vals[28];
            coef_values<ORDER>(p, raw, vals);
            add_to_entry<ORDER>(p, entry_of(S, local), vals);
          }
        } else {
          float* dst = p.partials + ((size_t)blockIdx.x * 2 + slot) * NC * TILE + in_tile;
#pragma unroll
          for (int q = 0; q < NC; ++q) dst[(size_t)q * TILE] = raw[q];
        }
      }
      seg_open = false;
    }
    cur = nxt;
  }
}

// The kernel: the pair loop over this CTA's unit range and — when launched cooperatively (fused_finalize) — after a
// grid-wide barrier the addition of the partial segments by all CTAs together (the finalize pass without a second
// launch, its prologue and the launch gap; every CTA of the persistent grid is resident by construction).
template <int ORDER, bool SHADOW, typename Math, bool USE_TMA, int MINB = 1, int NT = kThreads, int UNROLL = 4>
__global__ void __launch_bounds__(NT, MINB) gather_kernel(const __grid_constant__ GatherParams p) {
  trace_point(p, 0);
  gather_main<ORDER, SHADOW, Math, USE_TMA, NT, UNROLL>(p);
  trace_point(p, 1);
  if (p.fused_finalize) {
    __shared__ FinalizeSmem<ORDER, NT> fin;
    cg::this_grid().sync();
    trace_point(p, 2);
    finalize_phase<ORDER, NT>(p, NT * Math::CPT, fin, blockIdx.x, gridDim.x);
  }
  trace_point(p, 3);
}

// Sum of the partial segments the contributors c_lo..c_hi published for cache `t` of a tile, in CTA order (fixed =>
// deterministic). A function of its own (not inlined) so that it gets a register allocation of its own: inside the
// kernel the pair loop's live state left room for a dozen loads in flight and every contributor cost a full L2
// round trip (a tile of a small shard is split between ~30 CTAs: 12 us); here FU contributors x NC coefficients are
// loaded before the first one is consumed. A contributor past the end is clamped onto the last one and weighted 0.
template <int NC>
__device__ __noinline__ void fixup_sum(const float* __restrict__ partials, uint32_t tile_caches, uint32_t t, uint32_t c_lo,
                                       uint32_t c_hi, uint32_t sl_lo, float* __restrict__ raw) {
  constexpr int FU = NC > 12 ? 6 : 14;
  float acc[NC];
#pragma unroll
  for (int q = 0; q < NC; ++q) acc[q] = 0.0f;
  for (uint32_t c = c_lo; c <= c_hi; c += FU) {
    float v[FU][NC];
#pragma unroll
    for (int u = 0; u < FU; ++u) {
      const uint32_t cc = min(c + u, c_hi);
      const uint32_t sl = cc == c_lo ? sl_lo : 0u;
      const float* src = partials + ((size_t)cc * 2 + sl) * NC * tile_caches + t;
#pragma unroll
      for (int q = 0; q < NC; ++q) v[u][q] = __ldcg(src + (size_t)q * tile_caches);
    }
#pragma unroll
    for (int u = 0; u < FU; ++u) {
      const float wgt = (c + u <= c_hi) ? 1.0f : 0.0f;
#pragma unroll
      for (int q = 0; q < NC; ++q) acc[q] = fmaf(wgt, v[u][q], acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < NC; ++q) raw[q] = acc[q];
}

// ------------------------------------------------------------ the warp-split gather kernel
// Same pair maths, staging and stream-K unit schedule as gather_kernel, but the cache tile is what ONE warp holds
// in registers (32 lanes x CPT entries: 128 for SH1, 64 for SH2) and the NW warps of a CTA split the VPLs of every
// shared-memory tile between them instead of splitting the caches. What that buys at frame-sized cache counts:
//  * the tile granularity drops from 512 to 128 entries, so the padded slots of the last tile are ~1 % of a
//    6 k-cache frame instead of 7 %;
//  * a tile is split between ~grid / tiles CTAs (6 at 1080p) instead of ~23, each CTA first folds its NW warps
//    through shared memory, so a partial segment is 6 KB instead of 24 KB;
//  * the cross-CTA fix-up needs no grid-wide barrier and no cooperative launch: every contributor of a split tile
//    publishes its partial sums and takes a ticket; whoever draws the last ticket adds all partial segments in CTA
//    order (deterministic — no float atomics) and writes the entry (+ the peer stores of the fused all-gather).
template <int ORDER, bool SHADOW, typename Math, int NW, int UNROLL, int KVT = 1>
__global__ void __launch_bounds__(NW * 32, NW >= 8 ? 1 : 256 / (NW * 32)) gather_ws_kernel(const __grid_constant__ GatherParams p) {
  constexpr int CPT = Math::CPT;
  constexpr int TILE = 32 * CPT;
  constexpr int NT = NW * 32;
  constexpr int KV = NT * KVT; // VPLs staged per shared-memory tile: KVT per thread
  constexpr int NC = num_coefs<ORDER>();
  constexpr int STRIDE = ORDER == 2 ? 128 : 64;
  constexpr int SPV = Math::kSmemPerVpl;
  constexpr int FOLD = (TILE + NT - 1) / NT; // caches of the tile a thread folds at a segment end
  // the staged VPL tile and the cross-warp fold buffer share one allocation (the fold runs between two tiles)
  constexpr int RW = (NW < 4 ? NW : 4) * NC * TILE * 4 > 40960 ? 2 : (NW < 4 ? NW : 4); // warps folded per round
  constexpr int kVplBytes = KV * SPV * 16, kRedBytes = RW * NC * TILE * 4;
  __shared__ __align__(16) unsigned char s_raw[kVplBytes > kRedBytes ? kVplBytes : kRedBytes];
  float4* const s_vpl = reinterpret_cast<float4*>(s_raw);
  float (*const s_red)[NC][TILE] = reinterpret_cast<float (*)[NC][TILE]>(s_raw);
  __shared__ uint32_t s_blk[SHADOW ? KV : 1];
  __shared__ uint32_t s_last;

  trace_point(p, 0);
  Schedule S = make_schedule(p, TILE);
  if (S.units == 0) {
    // no live VPL at all: `entry.SH += 0`. In overwrite mode the SH was never cleared, so the zeros are written
    if (p.overwrite) {
      float vals[28];
#pragma unroll
      for (int q = 0; q < 28; ++q) vals[q] = 0.0f;
      for (uint32_t i = blockIdx.x * NT + threadIdx.x; i < S.count; i += gridDim.x * NT) add_to_entry<ORDER>(p, entry_of(S, i), vals);
    }
    return;
  }
  // every CTA of the (effective) grid owns a non-empty unit range
  const uint32_t G = (uint32_t)min((unsigned long long)gridDim.x, S.units);
  if (blockIdx.x >= G) return;
  const unsigned long long u0 = range_begin(S, G, blockIdx.x), u1 = range_begin(S, G, blockIdx.x + 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  Math M;
  float4 r0[KVT], r1[KVT], r2[KVT];
#pragma unroll
  for (int k = 0; k < KVT; ++k) r0[k] = r1[k] = r2[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto prefetch = [&](const Cursor& c) {
    const GatherLight& L = p.lights[c.light];
#pragma unroll
    for (int k = 0; k < KVT; ++k) {
      const uint32_t v = c.base + k * NT + threadIdx.x;
      const bool ok = v < c.v_end;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      r0[k] = ok ? __ldg(L.vpls + (size_t)v * 3) : make_float4(0.f, 0.f, 0.f, 1.f);
      r1[k] = ok ? __ldg(L.vpls + (size_t)v * 3 + 1) : z;
      r2[k] = ok ? __ldg(L.vpls + (size_t)v * 3 + 2) : z;
    }
  };
  Cursor cur;
  start_run(p, S, cur, u0, u1);
  prefetch(cur);
  bool seg_open = false;
  uint32_t seg_first_j = 0;
  bool traced = false;
  while (cur.valid) {
    if (!seg_open) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const uint32_t local = cur.tile * TILE + j * 32 + lane;
        const bool alive = local < S.count;
        const float4 pos = alive ? *reinterpret_cast<const float4*>(p.entries + (size_t)entry_of(S, local) * STRIDE)
                                 : make_float4(1e30f, 1e30f, 1e30f, 0.f);
        M.begin(j, alive, pos);
      }
      seg_first_j = (uint32_t)(cur.u - (unsigned long long)cur.tile * S.units_per_tile);
      seg_open = true;
    }
    Cursor nxt = cur;
    advance(p, S, nxt, u1, KV);
    __syncthreads(); // previous tile (or fold) fully consumed
#pragma unroll
    for (int k = 0; k < KVT; ++k) {
      Math::stage(&s_vpl[(k * NT + threadIdx.x) * SPV], r0[k], r1[k], r2[k]);
      if (SHADOW) s_blk[k * NT + threadIdx.x] = __float_as_uint(r1[k].w);
    }
    __syncthreads();
    if (nxt.valid) prefetch(nxt);
    if (!traced) { trace_point(p, 1); traced = true; } // prologue done: the first tile is staged
    // this warp's share of the staged VPLs
    const int n = (int)min((uint32_t)KV, cur.v_end - cur.base);
    const int share = (n + NW - 1) / NW;
    const int i0 = min(n, warp * share), i1 = min(n, i0 + share);
    if constexpr (SHADOW) {
      const float* col = p.shadow_table + (size_t)p.block_offset[cur.light] * p.shadow_stride + cur.tile * TILE + lane;
      auto load_shadow = [&](uint32_t blk) {
        float v[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          v[j] = (cur.tile * TILE + j * 32 + lane < S.count) ? __ldg(col + (size_t)blk * p.shadow_stride + j * 32) : 0.0f;
        M.set_shadow(v);
      };
      // runs of VPLs sharing a shadow block, found 32 VPLs at a time with one ballot
#pragma unroll 1
      for (int c0 = i0; c0 < i1; c0 += 32) {
        const int idx = c0 + lane, c1 = min(i1, c0 + 32);
        const bool start = idx < c1 && (lane == 0 || s_blk[idx] != s_blk[idx - 1]);
        uint32_t m = __ballot_sync(0xffffffffu, start);
        int i = c0;
        while (m) {
          m &= m - 1u;
          const int next = m ? c0 + (__ffs(m) - 1) : c1;
          load_shadow(s_blk[i]);
#pragma unroll UNROLL
          for (; i < next; ++i) M.eval(s_vpl + i * SPV);
        }
      }
    } else {
#pragma unroll UNROLL
      for (int i = i0; i < i1; ++i) M.eval(s_vpl + i * SPV);
    }
    if (!nxt.valid) trace_point(p, 2); // pair loop done
    if (!nxt.valid || nxt.tile != cur.tile) { // segment end: fold the warps, then write or publish
      const uint32_t last_j = (uint32_t)(cur.u_next - 1ull - (unsigned long long)cur.tile * S.units_per_tile);
      const bool full = seg_first_j == 0 && last_j == S.units_per_tile - 1;
      // fold the warps' accumulators through shared memory, RW warps per round, in warp order (deterministic)
      float folded[FOLD][NC];
#pragma unroll
      for (int k = 0; k < FOLD; ++k)
#pragma unroll
        for (int q = 0; q < NC; ++q) folded[k][q] = 0.0f;
#pragma unroll
      for (int w0 = 0; w0 < NW; w0 += RW) {
        __syncthreads(); // the staged tile (which the fold buffer aliases) / the previous round has been consumed
        if (warp >= w0 && warp < w0 + RW) {
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            float raw[27];
            M.raw(j, raw);
#pragma unroll
            for (int q = 0; q < NC; ++q) s_red[warp - w0][q][j * 32 + lane] = raw[q];
          }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < FOLD; ++k) {
          const uint32_t t = threadIdx.x + k * NT;
          if (t < (uint32_t)TILE) {
#pragma unroll
            for (int q = 0; q < NC; ++q)
#pragma unroll
              for (int w = 0; w < RW; ++w) folded[k][q] += s_red[w][q][t];
          }
        }
      }
      // slot 0: the segment in the tile that contains u0; slot 1: the (later) tile this range ends in
      const unsigned long long ua = (unsigned long long)cur.tile * S.units_per_tile, ub = ua + S.units_per_tile;
      const int slot = u0 >= ua ? 0 : 1;
#pragma unroll
      for (int k = 0; k < FOLD; ++k) { // thread t owns cache t (+ NT, ...) of the tile
        const uint32_t t = threadIdx.x + k * NT;
        if (t >= (uint32_t)TILE) break;
        if (full) {
          if (cur.tile * TILE + t < S.count) {
            float raw[27];
#pragma unroll
            for (int q = 0; q < 27; ++q) raw[q] = q < NC ? folded[k][q < NC ? q : 0] : 0.0f;
            float vals[28];
            coef_values<ORDER>(p, raw, vals);
            add_to_entry<ORDER>(p, entry_of(S, cur.tile * TILE + t), vals);
          }
        } else {
          float* dst = p.partials + ((size_t)blockIdx.x * 2 + slot) * NC * TILE + t;
#pragma unroll
          for (int q = 0; q < NC; ++q) __stcg(dst + (size_t)q * TILE, folded[k][q]);
        }
      }
      if (!full) {
        const uint32_t c_lo = owner_of(S, G, ua), c_hi = owner_of(S, G, ub - 1);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(p.tickets + cur.tile, 1u) == c_hi - c_lo) ? 1u : 0u;
        __syncthreads();
        trace_extra(p, 5);
        if (s_last) { // block-uniform: every other contributor's partial segment is visible
          __threadfence();
          trace_extra(p, 6);
          for (uint32_t t = threadIdx.x; t < (uint32_t)TILE; t += NT) {
            if (cur.tile * TILE + t >= S.count) break;
            float raw[27];
#pragma unroll
            for (int q = 0; q < 27; ++q) raw[q] = 0.0f;
            // only the first contributor can have started before this tile (slot 1)
            const uint32_t sl_lo = range_begin(S, G, c_lo) >= ua ? 0u : 1u;
            fixup_sum<NC>(p.partials, (uint32_t)TILE, t, c_lo, c_hi, sl_lo, raw);
            float vals[28];
            coef_values<ORDER>(p, raw, vals);
            add_to_entry<ORDER>(p, entry_of(S, cur.tile * TILE + t), vals);
          }
          trace_extra(p, 7);
          if (threadIdx.x == 0) p.tickets[cur.tile] = 0u; // rewound for the next launch
        }
      }
      seg_open = false;
    }
    cur = nxt;
  }
  trace_point(p, 3);
}

// ------------------------------------------------------------ pass 1: the visibility table
// cacheLightingRSM.comp:167-232 hoisted out of the pair loop: the reference traces a cone for every
// (cache, VAL block) inside its VPL loop; here all cones of a chunk of caches are traced first, by a kernel
// whose register budget is the cone march alone (so ~3x the warps per SM of a fused kernel hide the L1/L2
// latency of the dependent fetches), into table[block][cache] — 25 MB for the 1080p / 16 k-VPL frame, resident
// in B200's 126 MB L2 until the pair kernel (pass 2) reads every value exactly once, coalesced. Larger frames
// (configs[3]: 4096 blocks x 21 k caches = 346 MB) spill to HBM: one write and one read of every value, ~0.1 ms of
// a 7 ms pass at HBM speed.
struct ConeParams {
  GatherLight lights[DRV_MAX_LIGHTS];
  uint32_t block_offset[DRV_MAX_LIGHTS + 1]; // prefix sum of the lights' block counts
  uint32_t num_lights;
  const uint8_t* entries;
  uint32_t entry_stride;
  const drv_cache_counter* counter;
  uint32_t shard_rank, shard_world, shard_interleave;
  uint32_t chunk_first, chunk_cap;
  float* table;
  uint32_t stride;
  const uint2* rec;
  uint32_t rec_offset[16];
  ConeTables tables;
  int vres, vlevels;
  float vmin[3];
  float voxel_size;
  uint32_t* work; // [0] next item, [1] CTAs done — a device-side queue: items differ in cost (dead blocks, trip counts);
                  // [2..3] 64-bit count of the samples taken since drv_debug_cone_steps last read it
};

constexpr int kConeThreads = 128;
constexpr int kBlocksPerItem = 2; // small items: the queue balances the tail to ~1 / 24 of a CTA's share at 1080p

template <bool SIMPLE, int MINB>
__global__ void __launch_bounds__(kConeThreads, MINB) cone_kernel(const __grid_constant__ ConeParams p) {
  // this shard's chunk, as make_schedule derives it
  const uint32_t n = (uint32_t)max(p.counter->TotalLightCacheCount, 0);
  uint32_t first, lo, count;
  shard_span(n, p.shard_rank, p.shard_world, p.shard_interleave, p.chunk_first, p.chunk_cap, first, lo, count);
  if (count == 0) return;
  VoxelVol V;
  V.rec = p.rec; V.rec_offset = p.rec_offset; V.res = p.vres; V.levels = p.vlevels; V.voxel_size = p.voxel_size;
  V.vmin[0] = p.vmin[0]; V.vmin[1] = p.vmin[1]; V.vmin[2] = p.vmin[2];
  const uint32_t total_blocks = p.block_offset[p.num_lights];
  const uint32_t cache_groups = (count + kConeThreads - 1) / kConeThreads;
  const uint32_t block_groups = (total_blocks + kBlocksPerItem - 1) / kBlocksPerItem;
  const uint32_t items = cache_groups * block_groups; // <= 2048 x 2048
  __shared__ uint32_t s_item;
  uint32_t steps = 0; // samples this thread takes: the unit of the cone pass's roofline (bench.py)
  for (;;) {
    // items are handed out through an atomic counter: 13 % of the SM cycles were idle with a static round-robin
    if (threadIdx.x == 0) s_item = atomicAdd(p.work, 1u);
    __syncthreads();
    const uint32_t item = s_item;
    __syncthreads();
    if (item >= items) break;
    const uint32_t cg = item / block_groups, bg = item - cg * block_groups;
    const uint32_t local = cg * kConeThreads + threadIdx.x;
    const bool alive = local < count;
    float4 pos = make_float4(0.f, 0.f, 0.f, 0.f);
    if (alive) pos = *reinterpret_cast<const float4*>(p.entries + (size_t)entry_of(first, lo, p.shard_rank, p.shard_world, p.shard_interleave, local) * p.entry_stride);
    const uint32_t b_end = min(total_blocks, (bg + 1) * kBlocksPerItem);
    uint32_t light = 0;
    for (uint32_t b = bg * kBlocksPerItem; b < b_end; ++b) {
      while (b >= p.block_offset[light + 1]) ++light;
      if (!__ldg(p.lights[light].block_live + (b - p.block_offset[light]))) continue; // no live VPL reads this entry
      const float4 blk = __ldg(p.lights[light].blocks + (b - p.block_offset[light]));
      if (alive)
        p.table[(size_t)b * p.stride + local] = SIMPLE ? cone_trace_simple(V, p.tables, pos.x, pos.y, pos.z, blk, steps)
                                                       : cone_trace(V, pos.x, pos.y, pos.z, blk, steps);
    }
  }
  steps = __reduce_add_sync(0xffffffffu, steps);
  if ((threadIdx.x & 31) == 0 && steps) atomicAdd(reinterpret_cast<unsigned long long*>(p.work + 2), (unsigned long long)steps);
  // the last CTA to leave rewinds the queue for the next launch
  if (threadIdx.x == 0 && atomicAdd(p.work + 1, 1u) == gridDim.x - 1) {
    p.work[0] = 0u;
    p.work[1] = 0u;
  }
}

} // namespace

// -------------------------------------------------------------------------------- host side
namespace {

using GatherFn = void (*)(const GatherParams);

drv_status launch_gather(drv_ctx* ctx, GatherFn kernel, GatherParams& p, int tile_caches, int order, int threads, bool ws) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
  if (per_sm < 1) per_sm = 1;
  const int grid = ctx->num_sms * per_sm;
  const size_t need = (size_t)grid * 2 * (order == 2 ? 27 : 12) * tile_caches;
  if (need > ctx->partial_slots) { // grow the per-CTA partial scratch (first launch of a variant)
    cudaStreamSynchronize(ctx->stream);
    if (ctx->partials) cudaFree(ctx->partials);
    ctx->partials = nullptr;
    ctx->partial_slots = 0;
    DRV_CUDA(cudaMalloc(&ctx->partials, need * sizeof(float)));
    ctx->partial_slots = need;
  }
  p.partials = ctx->partials;
  p.grid = (uint32_t)grid;
  p.tickets = ctx->gather_tickets;
  p.trace = nullptr;
  if (ctx->cfg.gather_variant & 0x40000u) { // diagnostics: per-CTA timestamps, read with drv_debug_gather_trace
    if (!ctx->gather_trace) {
      DRV_CUDA(cudaMalloc(&ctx->gather_trace, (size_t)8192 * 8 * sizeof(unsigned long long)));
      DRV_CUDA(cudaMemsetAsync(ctx->gather_trace, 0, (size_t)8192 * 8 * sizeof(unsigned long long), ctx->stream));
    }
    if (grid <= 8192) p.trace = ctx->gather_trace;
    ctx->gather_trace_ctas = (uint32_t)grid;
  }
  if (ws) { // warp-split kernel: ticketed fix-up inside the one ordinary launch
    p.fused_finalize = 0u;
    kernel<<<grid, threads, 0, ctx->stream>>>(p);
    DRV_LAUNCH_CHECK();
    return DRV_OK;
  }
  // one cooperative launch (pair loop, grid barrier, finalize) when the device can do it; gather_variant bit 16
  // forces the two-kernel form
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
  p.fused_finalize = (coop && !(ctx->cfg.gather_variant & 0x10000u)) ? 1u : 0u;
  if (p.fused_finalize) {
    void* args[] = {(void*)&p};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(threads), args, 0, ctx->stream);
    if (e != cudaSuccess) return ctx->fail(DRV_ERR_CUDA, std::string("cooperative gather launch: ") + cudaGetErrorString(e));
    DRV_LAUNCH_CHECK();
    return DRV_OK;
  }
  kernel<<<grid, threads, 0, ctx->stream>>>(p);
  DRV_LAUNCH_CHECK();
  const int fin_grid = ctx->num_sms * 8;
  if (order == 1) gather_finalize_kernel<1><<<fin_grid, kFinThreads, 0, ctx->stream>>>(p, tile_caches);
  else gather_finalize_kernel<2><<<fin_grid, kFinThreads, 0, ctx->stream>>>(p, tile_caches);
  DRV_LAUNCH_CHECK();
  return DRV_OK;
}

// Kernel variants (drv_config.gather_variant; measurements in profiles/):
//   0  = 31: warp-split kernel, one 8-warp CTA per SM, 6 caches / thread (SH1) or 4 (SH2) — default
//   30 the round-1 default: cooperative stream-K kernel, packed FP32x2 maths, 4 caches / thread (SH1) or 2 (SH2)
//   1  scalar maths, VPL tiles staged with TMA bulk copies + mbarrier, double buffered
//   3  packed, one pair of caches per thread
//   4  scalar, 4 caches / thread (SH1)
//   12 scalar, 2 caches / thread, register-prefetch staging
//   6  as 0 for SH1 but compiled for 3 resident CTAs per SM (168 registers)
//   7  as 0 with 64-thread CTAs (cache tiles of 256 / 128 entries: less padding when the frame has few caches)
//   8  as 0 with 32-thread CTAs (cache tiles of 128 / 64 entries)
//   9, 10, 11  the packed kernel with the VPL loop unrolled 2x / 8x / 4x (default: 8x for SH1, 4x for SH2)
// With indirect shadows the same kernels additionally scale every pair by its table visibility.
//   20 warp-split kernel (gather_ws_kernel): 4 warps share a 128- (SH1) / 64-entry (SH2) tile and split the VPLs,
//      ticketed fix-up, ordinary launch; 21 the same with 2-warp CTAs; 22 / 23 as 20 with the VPL loop unrolled 4x / 2x;
//      24 / 25 as 20 with 256 / 512 VPLs per shared-memory tile; 26 / 27 ONE 8-warp CTA per SM (256 / 512 VPLs per
//      tile): the warps of a CTA advance in lock step through the tile barriers, so no SM is left with half its
//      warps while a co-resident CTA that the scheduler favoured has already finished; 28 one 12-warp CTA per SM
template <bool SH>
GatherFn select_kernel(int order, uint32_t variant, int* tile, int* threads, bool* ws) {
  *threads = kThreads;
  *ws = false;
  if (variant == 0) variant = 35; // the default: fastest at every measured size (profiles/r2_gather_variants.md)
  if (variant == 31 || variant == 32) { // three (SH1) / two (SH2) packed pairs per thread, one 8-warp CTA per SM
    using P1 = PackedMath<1, SH, 3>; using P2 = PackedMath<2, SH, 2>;
    *ws = true;
    *tile = 32 * (order == 1 ? P1::CPT : P2::CPT);
    *threads = 256;
    if (order == 1) return variant == 31 ? gather_ws_kernel<1, SH, P1, 8, 8, 2> : gather_ws_kernel<1, SH, P1, 8, 4, 2>;
    return variant == 31 ? gather_ws_kernel<2, SH, P2, 8, 4, 2> : gather_ws_kernel<2, SH, P2, 8, 2, 2>;
  }
  if (variant >= 35 && variant <= 38) { // as 31, the per-VPL scalars as 32-bit broadcast operands (PackedMath BCAST)
    using P1 = PackedMath<1, SH, 3, true>; using P2 = PackedMath<2, SH, 2, true>;
    using Q1 = PackedMath<1, SH, 4, true>; using Q2 = PackedMath<2, SH, 3, true>;
    *ws = true;
    *threads = 256;
    if (variant == 37 || variant == 38) { // four (SH1) / three (SH2) packed pairs per thread
      *tile = 32 * (order == 1 ? Q1::CPT : Q2::CPT);
      if (order == 1) return variant == 37 ? gather_ws_kernel<1, SH, Q1, 8, 4, 2> : gather_ws_kernel<1, SH, Q1, 8, 8, 2>;
      return variant == 37 ? gather_ws_kernel<2, SH, Q2, 8, 2, 2> : gather_ws_kernel<2, SH, Q2, 8, 4, 2>;
    }
    *tile = 32 * (order == 1 ? P1::CPT : P2::CPT);
    if (order == 1) return variant == 35 ? gather_ws_kernel<1, SH, P1, 8, 8, 2> : gather_ws_kernel<1, SH, P1, 8, 4, 2>;
    return variant == 35 ? gather_ws_kernel<2, SH, P2, 8, 8, 2> : gather_ws_kernel<2, SH, P2, 8, 4, 2>;
  }
  if (variant == 39 || variant == 40) { // as 35 with 768 VPLs per staged tile (39) / the VPL loop unrolled 16x / 8x, 256 per tile (40)
    using P1 = PackedMath<1, SH, 3, true>; using P2 = PackedMath<2, SH, 2, true>;
    *ws = true;
    *threads = 256;
    *tile = 32 * (order == 1 ? P1::CPT : P2::CPT);
    if (order == 1) return variant == 39 ? gather_ws_kernel<1, SH, P1, 8, 8, 3> : gather_ws_kernel<1, SH, P1, 8, 16, 2>;
    return variant == 39 ? gather_ws_kernel<2, SH, P2, 8, 8, 3> : gather_ws_kernel<2, SH, P2, 8, 8, 1>;
  }
  if (variant == 33 || variant == 34) { // four packed pairs per thread (SH1), 256-entry tiles
    using P1 = PackedMath<1, SH, 4>; using P2 = PackedMath<2, SH, 2>;
    *ws = true;
    *tile = 32 * (order == 1 ? P1::CPT : P2::CPT);
    *threads = 256;
    if (order == 1) return variant == 33 ? gather_ws_kernel<1, SH, P1, 8, 4, 2> : gather_ws_kernel<1, SH, P1, 8, 2, 2>;
    return variant == 33 ? gather_ws_kernel<2, SH, P2, 8, 4, 1> : gather_ws_kernel<2, SH, P2, 8, 8, 2>;
  }
  if (variant >= 20 && variant <= 28) {
    using P1 = PackedMath<1, SH, 2>; using P2 = PackedMath<2, SH, 1>;
    *ws = true;
    *tile = 32 * (order == 1 ? P1::CPT : P2::CPT);
    *threads = variant == 21 ? 64 : variant == 28 ? 384 : variant >= 26 ? 256 : 128;
    if (order == 1) {
      switch (variant) {
        case 21: return gather_ws_kernel<1, SH, P1, 2, 8>;
        case 22: return gather_ws_kernel<1, SH, P1, 4, 4>;
        case 23: return gather_ws_kernel<1, SH, P1, 4, 2>;
        case 24: return gather_ws_kernel<1, SH, P1, 4, 8, 2>;
        case 25: return gather_ws_kernel<1, SH, P1, 4, 8, 4>;
        case 26: return gather_ws_kernel<1, SH, P1, 8, 8, 1>;
        case 27: return gather_ws_kernel<1, SH, P1, 8, 8, 2>;
        case 28: return gather_ws_kernel<1, SH, P1, 12, 8, 1>;
        default: return gather_ws_kernel<1, SH, P1, 4, 8>;
      }
    }
    switch (variant) {
      case 21: return gather_ws_kernel<2, SH, P2, 2, 4>;
      case 22: return gather_ws_kernel<2, SH, P2, 4, 8>;
      case 23: return gather_ws_kernel<2, SH, P2, 4, 2>;
      case 24: return gather_ws_kernel<2, SH, P2, 4, 4, 2>;
      case 25: return gather_ws_kernel<2, SH, P2, 4, 4, 4>;
      case 26: return gather_ws_kernel<2, SH, P2, 8, 4, 1>;
      case 27: return gather_ws_kernel<2, SH, P2, 8, 4, 2>;
      case 28: return gather_ws_kernel<2, SH, P2, 12, 4, 1>;
      default: return gather_ws_kernel<2, SH, P2, 4, 4>;
    }
  }
#define DRV_PICK(ORD, MATH, TMA) do { *tile = kThreads * MATH::CPT; return gather_kernel<ORD, SH, MATH, TMA>; } while (0)
  using S1c2 = ScalarMath<1, SH, 2>; using S1c4 = ScalarMath<1, SH, 4>; using S2c2 = ScalarMath<2, SH, 2>;
  using P1p1 = PackedMath<1, SH, 1>; using P1p2 = PackedMath<1, SH, 2>; using P2p1 = PackedMath<2, SH, 1>;
  if (order == 1) {
    switch (variant) {
      case 1: DRV_PICK(1, S1c2, true);
      case 3: DRV_PICK(1, P1p1, false);
      case 4: DRV_PICK(1, S1c4, false);
      case 12: DRV_PICK(1, S1c2, false);
      case 6: *tile = kThreads * P1p2::CPT; return gather_kernel<1, SH, P1p2, false, 3>; // 168 registers: 3 CTAs / SM
      case 7: *threads = 64; *tile = 64 * P1p2::CPT; return gather_kernel<1, SH, P1p2, false, 1, 64>; // 2-warp CTAs, 256-cache tiles
      case 8: *threads = 32; *tile = 32 * P1p2::CPT; return gather_kernel<1, SH, P1p2, false, 1, 32>; // 1-warp CTAs, 128-cache tiles
      case 9: *tile = kThreads * P1p2::CPT; return gather_kernel<1, SH, P1p2, false, 1, kThreads, 2>;  // VPL loop unrolled 2x
      case 10: *tile = kThreads * P1p2::CPT; return gather_kernel<1, SH, P1p2, false, 1, kThreads, 8>; // ... 8x
      case 11: DRV_PICK(1, P1p2, false); // VPL loop unrolled 4x
      default: *tile = kThreads * P1p2::CPT; return gather_kernel<1, SH, P1p2, false, 1, kThreads, 8>; // 8x: +1.5 % (sweep)
    }
  }
  switch (variant) {
    case 1: DRV_PICK(2, S2c2, true);
    case 4: case 12: DRV_PICK(2, S2c2, false);
    case 7: *threads = 64; *tile = 64 * P2p1::CPT; return gather_kernel<2, SH, P2p1, false, 1, 64>;
    case 8: *threads = 32; *tile = 32 * P2p1::CPT; return gather_kernel<2, SH, P2p1, false, 1, 32>;
    case 9: *tile = kThreads * P2p1::CPT; return gather_kernel<2, SH, P2p1, false, 1, kThreads, 2>;
    case 10: *tile = kThreads * P2p1::CPT; return gather_kernel<2, SH, P2p1, false, 1, kThreads, 8>;
    default: DRV_PICK(2, P2p1, false);
  }
#undef DRV_PICK
}

// caches per visibility-table chunk: every chunk costs three launches whether it holds caches or not (the count
// lives on the device), so chunks are as large as a table of at most kShadowTableBytes allows
constexpr uint32_t kShadowChunk = 262144;
constexpr size_t kShadowTableBytes = 6ull << 30;

} // namespace

// ---- cross-GPU barrier over NVLink peer memory ---------------------------------------------------------
// Every rank stores the new epoch into slot [rank] of every peer's flag block (which sits behind that peer's
// entries buffer), then spins on its own block until all peers' slots carry the epoch. Kernels launched earlier
// on the stream — including their peer stores — have completed before this kernel starts, so once a rank leaves
// the barrier every peer's previous stage is visible to it. Epochs only grow: no reset, no ABA.
namespace {
struct BarrierArgs {
  uint32_t* own;
  uint32_t* peer[8];
  uint32_t rank, world;
};
// flags[0..7] = last epoch each rank announced to this GPU, flags[8] = time-out marker (the epoch of the barrier
// that gave up; read by drv_peer_status), flags[10] = the rank it was waiting for, flags[9] = this GPU's own epoch
// counter. The epoch lives on the device so that the launch has no per-call argument and a recorded
// frame graph can replay it; all ranks call the barrier equally often, so their counters advance in lock step.
__global__ void peer_barrier_kernel(BarrierArgs a) {
  const uint32_t t = threadIdx.x;
  uint32_t epoch = 0;
  if (t == 0) { epoch = a.own[9] + 1u; a.own[9] = epoch; }
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  const bool active = t < a.world && t != a.rank;
  __threadfence_system();
  if (active) *reinterpret_cast<volatile uint32_t*>(a.peer[t] + a.rank) = epoch;
  __syncwarp(); // all announcements are on their way before anybody starts to wait
  if (!active) return;
  // a barrier that timed out earlier poisons the ones behind it: they still announce (so that healthy peers do not
  // hang on this rank and the epoch counters stay in lock step) but no longer wait — the frame drains at once and
  // the host sees DRV_ERR_PEER from drv_peer_status / drv_active_cache_count instead of a half-gathered image
  if (*reinterpret_cast<volatile uint32_t*>(a.own + 8) != 0u) return;
  const long long t0 = clock64();
  while ((int32_t)(*reinterpret_cast<volatile uint32_t*>(a.own + t) - epoch) < 0) {
    if (clock64() - t0 > 8000000000ll) { // ~4 s at 2 GHz: a peer died; do not hang the GPU
      a.own[8] = epoch;
      a.own[10] = t; // the rank that did not arrive
      break;
    }
  }
  __threadfence_system();
}
} // namespace

drv_status drv_impl_peer_barrier(drv_ctx* ctx) {
  BarrierArgs a;
  memset(&a, 0, sizeof(a));
  a.own = ctx->sync_flags;
  const size_t off = (size_t)ctx->cfg.max_cache_count * 128;
  for (uint32_t r = 0; r < ctx->shard_world && r < 8; ++r)
    a.peer[r] = r == ctx->shard_rank ? ctx->sync_flags : reinterpret_cast<uint32_t*>((uint8_t*)ctx->peer_entries[r] + off);
  a.rank = ctx->shard_rank;
  a.world = ctx->shard_world;
  peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(a);
  DRV_LAUNCH_CHECK();
  return DRV_OK;
}

drv_status drv_impl_gather(drv_ctx* ctx, bool overwrite) {
  if (!ctx->have_constant) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_light_caches: Constant block not set");
  const bool shadow = ctx->cfg.indirect_shadow != 0;
  if (shadow && !ctx->have_volume) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_light_caches: VolumeInfo not set");
  if (ctx->num_lights == 0) return DRV_OK;
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.num_lights = ctx->num_lights;
  const uint32_t granule = 32; // scheduling quantum in VPLs (a shared-memory tile still holds up to kVplTile)
  for (uint32_t l = 0; l < ctx->num_lights; ++l) {
    LightState& S = ctx->lights[l];
    p.lights[l].vpls = (const float4*)S.vpls_live;
    p.lights[l].blocks = (const float4*)S.blocks;
    p.lights[l].live = ctx->live_counts + l;
    p.lights[l].block_live = S.block_live;
    p.lights[l].num_vpls = S.num_vpls;
    uint32_t interval = shadow ? (uint32_t)S.block.IndirectShadowComputationSampleInterval : 1u;
    if (interval == 0 || (interval & (interval - 1)) != 0)
      return ctx->fail(DRV_ERR_INVALID, "drv_light_caches: shadow sample interval must be a power of two");
    if (shadow && S.vpls_external)
      return ctx->fail(DRV_ERR_INVALID, "drv_light_caches: drv_set_vpls cannot be combined with indirect shadows");
    p.lights[l].interval = interval;
  }
  p.granule = granule;
  p.entries = ctx->entries;
  p.counter = ctx->counter;
  p.shard_rank = ctx->shard_rank;
  p.shard_world = ctx->shard_world;
  p.shard_interleave = (ctx->shard_interleave && ctx->shard_world > 1) ? 1u : 0u;
  p.overwrite = overwrite ? 1u : 0u;
  p.f0 = ctx->constant.ShEvaFactor0;
  p.f1 = ctx->constant.ShEvaFactor1;
  p.f2 = ctx->constant.ShEvaFactor2n2_p1_n1;
  p.f20 = ctx->constant.ShEvaFactor20;
  p.f22 = ctx->constant.ShEvaFactor2p2;
  if (drv_peers_complete(ctx)) {
    p.num_peers = ctx->shard_world;
    for (uint32_t r = 0; r < ctx->shard_world && r < 8; ++r)
      p.peers[r] = (r == ctx->shard_rank) ? nullptr : (uint8_t*)ctx->peer_entries[r];
  }
  const int order = (int)ctx->cfg.sh_order;
  const uint32_t variant = ctx->cfg.gather_variant & 0xFFu; // bits 8.. tune other kernels
  int tile = 0, threads = kThreads;
  bool ws = false;
  GatherFn kernel = shadow ? select_kernel<true>(order, variant, &tile, &threads, &ws)
                           : select_kernel<false>(order, variant, &tile, &threads, &ws);
  if (overwrite && !ws) // only the warp-split kernel writes the zeros of a frame without live VPLs: take the default
    kernel = shadow ? select_kernel<true>(order, 0, &tile, &threads, &ws) : select_kernel<false>(order, 0, &tile, &threads, &ws);
  if (ws) p.granule = 8; // the warps split every staged tile evenly, so a finer quantum only improves the balance
  p.chunk_first = 0;
  p.chunk_cap = 0xFFFFFFFFu;
  if (!shadow) {
    ctx->stage_begin(DRV_STAGE_GATHER_KERNEL);
    drv_status st = launch_gather(ctx, kernel, p, tile, order, threads, ws);
    ctx->stage_end(DRV_STAGE_GATHER_KERNEL);
    return st;
  }

  // ---- indirect shadows: pass 1 traces every (cache, VAL block) cone of a chunk of caches into the visibility
  // table, pass 2 is the pair kernel reading it. The cache count lives on the device, so the host walks all
  // chunks up to max_cache_count; kernels of empty chunks return at once.
  ConeParams c;
  memset(&c, 0, sizeof(c));
  uint32_t total_blocks = 0;
  for (uint32_t l = 0; l < ctx->num_lights; ++l) {
    c.lights[l] = p.lights[l];
    c.block_offset[l] = p.block_offset[l] = total_blocks;
    total_blocks += (p.lights[l].num_vpls + p.lights[l].interval - 1) / p.lights[l].interval;
  }
  c.block_offset[ctx->num_lights] = total_blocks;
  c.num_lights = ctx->num_lights;
  c.entries = ctx->entries;
  c.entry_stride = ctx->entry_stride;
  c.counter = ctx->counter;
  c.shard_rank = ctx->shard_rank;
  c.shard_world = ctx->shard_world;
  c.shard_interleave = p.shard_interleave;
  c.rec = ctx->voxel_records;
  for (int l = 0; l < 16; ++l) {
    c.rec_offset[l] = ctx->voxel_record_offset[l];
    const uint32_t S = ((uint32_t)ctx->cfg.voxel_resolution >> l) + 1u + 2u * kVoxelRecordPad, K = 1u + S + S * S;
    c.tables.rec_k[l] = c.rec_offset[l] + (1u + kVoxelRecordPad) * K;
    c.tables.rec_kb[l] = c.tables.rec_k[l] - 0x4B400000u * K;
  }
  c.vres = (int)ctx->cfg.voxel_resolution;
  c.vlevels = (int)ctx->voxel_levels;
  memcpy(c.vmin, ctx->volume.VolumeWorldMin, 12);
  c.voxel_size = ctx->volume.VoxelSizeInWorld;
  uint32_t chunk = std::min<uint32_t>(kShadowChunk, (ctx->cfg.max_cache_count + 511u) & ~511u);
  while (chunk > 8192 && (size_t)total_blocks * chunk * sizeof(float) > kShadowTableBytes) chunk = (chunk / 2 + 511u) & ~511u;
  const size_t table_floats = (size_t)total_blocks * chunk;
  if (table_floats > ctx->shadow_table_floats) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->shadow_table) cudaFree(ctx->shadow_table);
    ctx->shadow_table = nullptr;
    ctx->shadow_table_floats = 0;
    DRV_CUDA(cudaMalloc(&ctx->shadow_table, table_floats * sizeof(float)));
    ctx->shadow_table_floats = table_floats;
  }
  c.table = ctx->shadow_table;
  c.stride = chunk;
  ctx->shadow_stride = chunk; // what the specular pass needs to read the table after this call
  ctx->shadow_chunks = (ctx->cfg.max_cache_count + chunk - 1) / chunk;
  for (uint32_t l = 0; l < ctx->num_lights; ++l) ctx->shadow_block_offset[l] = p.block_offset[l];
  c.work = ctx->cone_work;
  p.shadow_table = ctx->shadow_table;
  p.shadow_stride = chunk;
  // gather_variant bit 17: the software-pipelined march instead of the plain loop
  using ConeFn = void (*)(const ConeParams);
  const ConeFn cone_fn = (ctx->cfg.gather_variant & 0x20000u) ? (ConeFn)cone_kernel<false, 8>
                         : (ctx->cfg.gather_variant & 0x80000u) ? (ConeFn)cone_kernel<true, 8> : (ConeFn)cone_kernel<true, 10>;
  int cone_per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cone_per_sm, cone_fn, kConeThreads, 0);
  if (cone_per_sm < 1) cone_per_sm = 1;
  const int cone_grid = ctx->num_sms * cone_per_sm;
  ctx->stage_begin(DRV_STAGE_GATHER_KERNEL);
  for (uint32_t first = 0; first < ctx->cfg.max_cache_count; first += chunk) {
    c.chunk_first = p.chunk_first = first;
    c.chunk_cap = p.chunk_cap = chunk;
    if (first == 0) ctx->stage_begin(DRV_STAGE_CONE_KERNEL); // the chunk that holds the caches (<= 262 144 of them)
    cone_fn<<<cone_grid, kConeThreads, 0, ctx->stream>>>(c);
    DRV_LAUNCH_CHECK();
    if (first == 0) ctx->stage_end(DRV_STAGE_CONE_KERNEL);
    drv_status st = launch_gather(ctx, kernel, p, tile, order, threads, ws);
    if (st != DRV_OK) return st;
  }
  ctx->stage_end(DRV_STAGE_GATHER_KERNEL);
  return DRV_OK;
}

// Diagnostics: what the last gather launch wrote, 8 words per CTA: %globaltimer (ns) at kernel entry, prologue done /
// pair loop done (the two kernels differ, see trace_point calls), exit; then %smid and clock64 at the last three
// points. Needs gather_variant bit 18.
extern "C" drv_status drv_debug_gather_trace(drv_ctx* ctx, uint64_t* out, uint32_t capacity_ctas, uint32_t* num_ctas) {
  if (!ctx) return DRV_ERR_INVALID;
  if (!ctx->gather_trace) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_debug_gather_trace: gather_variant bit 18 is not set");
  const uint32_t n = ctx->gather_trace_ctas < capacity_ctas ? ctx->gather_trace_ctas : capacity_ctas;
  DRV_CUDA(cudaMemcpyAsync(out, ctx->gather_trace, (size_t)n * 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  DRV_CUDA(cudaStreamSynchronize(ctx->stream));
  if (num_ctas) *num_ctas = n;
  return DRV_OK;
}

// Diagnostics: the number of voxel samples (cone steps) cone_kernel has taken since the last call; resets it.
extern "C" drv_status drv_debug_cone_steps(drv_ctx* ctx, uint64_t* steps) {
  if (!ctx || !steps) return DRV_ERR_INVALID;
  cudaSetDevice(ctx->device);
  DRV_CUDA(cudaMemcpyAsync(steps, ctx->cone_work + 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  DRV_CUDA(cudaMemsetAsync(ctx->cone_work + 2, 0, sizeof(uint64_t), ctx->stream));
  DRV_CUDA(cudaStreamSynchronize(ctx->stream));
  return DRV_OK;
}
