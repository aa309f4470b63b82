/*
 * alloc.cpp — oracle for stage 1 (cache allocation). TEST INFRASTRUCTURE ONLY.
 * Restates shader/cacheGather.comp:20-164, lightcache.glsl:109-134 and
 * cachePrepareLighting.comp:8-14 (see oracle.h for the arithmetic policy).
 */
#include "oracle.h"
#include "glsl_scalar.h"

using namespace orc;

namespace {

struct AllocParams {
  int W, H, R, C;
  const float* ivp;
  const drv_cav_cascade* casc;
  float zone;
  bool transitions;
};

/* cacheGather.comp:113-119 (identical in cacheApply.frag:134-135). */
inline vec3 world_position(const AllocParams& p, int x, int y, float depth) {
  float px = (float)x + 0.5f, py = (float)y + 0.5f;
  float ndc[4] = {px / (float)p.W * 2.0f - 1.0f, py / (float)p.H * 2.0f - 1.0f, depth, 1.0f};
  float w4[4];
  mul_row_major(p.ivp, ndc, w4);
  return V3(w4[0] / w4[3], w4[1] / w4[3], w4[2] / w4[3]);
}

/* lightcache.glsl:109-122. */
inline int compute_cascade(const AllocParams& p, vec3 wp) {
  int c = 0;
  for (; c < p.C - 1; ++c) {
    const drv_cav_cascade& k = p.casc[c];
    if (wp.x <= k.DecisionMax[0] && wp.y <= k.DecisionMax[1] && wp.z <= k.DecisionMax[2] &&
        wp.x >= k.DecisionMin[0] && wp.y >= k.DecisionMin[1] && wp.z >= k.DecisionMin[2])
      break;
  }
  return c;
}

/* lightcache.glsl:125-134. */
inline float cascade_transition(const AllocParams& p, vec3 wp, int c) {
  const drv_cav_cascade& k = p.casc[c];
  vec3 toMax = V3(k.DecisionMax) - wp;
  vec3 toMin = wp - V3(k.DecisionMin);
  float minDist = std::fmin(std::fmin(std::fmin(toMax.x, toMax.y), toMax.z),
                            std::fmin(std::fmin(toMin.x, toMin.y), toMin.z));
  return saturate(1.0f - minDist / (k.WorldVoxelSize * p.zone));
}

/* cacheGather.comp:20-30. */
inline int cache_1d_coord(const AllocParams& p, vec3 wp, int c) {
  const drv_cav_cascade& k = p.casc[c];
  vec3 g = (wp - V3(k.Min)) / k.WorldVoxelSize;
  int gx = clampi(trunc_to_int(g.x), 0, p.R - 1);
  int gy = clampi(trunc_to_int(g.y), 0, p.R - 1);
  int gz = clampi(trunc_to_int(g.z), 0, p.R - 1);
  return gx + gy * p.R + gz * p.R * p.R + c * p.R * p.R * p.R;
}

/* cacheGather.comp:32-91 with the index assignment deferred: marks the 8
 * corner cells. Corners with a component >= R are skipped. */
inline void mark_corners(const AllocParams& p, int coord, int c, uint8_t* flags, uint32_t* oob) {
  static const int off[8][3] = {{0, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1},
                                {1, 0, 0}, {1, 1, 0}, {1, 0, 1}, {1, 1, 1}};
  const int R = p.R, R2 = R * R, R3 = R2 * R;
  int local = coord - R3 * c;
  int bz = local / R2;
  int by = (local - bz * R2) / R;
  int bx = local % R;
  for (int i = 0; i < 8; ++i) {
    int x = bx + off[i][0], y = by + off[i][1], z = bz + off[i][2];
    if (x >= R || y >= R || z >= R) {
      __atomic_fetch_add(oob, 1u, __ATOMIC_RELAXED);
      continue;
    }
    __atomic_store_n(&flags[(size_t)x + (size_t)y * R + (size_t)z * R2 + (size_t)c * R3], (uint8_t)1,
                     __ATOMIC_RELAXED);
  }
}

/* cacheGather.comp:93-164 for one 16x16 work group. */
void run_tile(const AllocParams& p, const float* depth, int tx, int ty, uint8_t* flags, uint32_t* oob) {
  int T1[16][16], T2[16][16], CS[16][16]; /* [local x][local y] like cacheList[x][y] */
  for (int ly = 0; ly < 16; ++ly)
    for (int lx = 0; lx < 16; ++lx) {
      T1[lx][ly] = -1; T2[lx][ly] = -1; CS[lx][ly] = -1;
      int x = tx * 16 + lx, y = ty * 16 + ly;
      if (x >= p.W || y >= p.H) continue;
      float d = depth[(size_t)y * p.W + x];
      if (!(d > 0.0001f)) continue;
      vec3 wp = world_position(p, x, y, d);
      int c = compute_cascade(p, wp);
      CS[lx][ly] = c;
      T1[lx][ly] = cache_1d_coord(p, wp, c);
      if (p.transitions) {
        float t = cascade_transition(p, wp, c);
        if (t > 0.0f && c < p.C - 1) T2[lx][ly] = cache_1d_coord(p, wp, c + 1);
      }
    }
  for (int ly = 0; ly < 16; ++ly)
    for (int lx = 0; lx < 16; ++lx) {
      int own = T1[lx][ly];
      int ax = std::max(0, lx - 1), ay = std::max(0, ly - 1);
      if (((T1[lx][ay] != own && T1[ax][ly] != own && T1[ax][ay] != own) || (ax == lx && ay == ly)) &&
          own != -1)
        mark_corners(p, own, CS[lx][ly], flags, oob);
      if (p.transitions) {
        int own2 = T2[lx][ly];
        int bx = std::min(15, lx + 1), by = std::min(15, ly + 1);
        if (((T2[lx][by] != own2 && T2[bx][ly] != own2 && T2[bx][by] != own2) || (bx == 15 && by == 15)) &&
            own2 != -1)
          mark_corners(p, own2, CS[lx][ly] + 1, flags, oob);
      }
    }
}

void mark_all(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi, int transitions,
              const float* depth, std::vector<uint8_t>& flags, uint32_t* oob, int threads, AllocParams& p) {
  p.W = cb->BackbufferResolution[0];
  p.H = cb->BackbufferResolution[1];
  p.R = cb->AddressVolumeResolution;
  p.C = cb->NumAddressVolumeCascades;
  p.ivp = pf->InverseViewProjection;
  p.casc = vi->AddressVolumeCascades;
  p.zone = vi->CAVTransitionZoneSize;
  p.transitions = transitions != 0;
  flags.assign((size_t)p.R * p.R * p.R * p.C, 0);
  int tilesX = (p.W + 15) / 16, tilesY = (p.H + 15) / 16;
  uint8_t* f = flags.data();
  parallel_for((int64_t)tilesX * tilesY, threads, [&](int64_t b, int64_t e, int) {
    for (int64_t t = b; t < e; ++t) run_tile(p, depth, (int)(t % tilesX), (int)(t / tilesX), f, oob);
  });
}

} // namespace

extern "C" int orc_default_threads(void) { return default_threads(); }

extern "C" int orc_allocated_cell_ids(const drv_constant* cb, const drv_per_frame* pf,
                                      const drv_volume_info* vi, int transitions, const float* depth,
                                      int32_t* ids, uint32_t max_ids, int threads) {
  std::vector<uint8_t> flags;
  AllocParams p;
  uint32_t oob = 0;
  mark_all(cb, pf, vi, transitions, depth, flags, &oob, threads, p);
  uint32_t n = 0;
  for (size_t i = 0; i < flags.size(); ++i)
    if (flags[i]) {
      if (ids && n < max_ids) ids[n] = (int32_t)i;
      ++n;
    }
  return (int)n;
}

extern "C" int orc_allocate_caches(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi,
                                   int transitions, const float* depth, uint32_t* atlas, void* entries,
                                   uint32_t entry_stride, uint32_t max_caches, drv_cache_counter* counter,
                                   uint32_t* overflow, uint32_t* oob_corners, int threads) {
  std::vector<uint8_t> flags;
  AllocParams p;
  uint32_t oob = 0;
  mark_all(cb, pf, vi, transitions, depth, flags, &oob, threads, p);
  const int R = p.R, R2 = R * R, R3 = R2 * R;
  const size_t atlasW = (size_t)R * p.C;
  std::memset(atlas, 0, atlasW * R * R * sizeof(uint32_t));
  uint32_t n = 0, dropped = 0;
  for (size_t id = 0; id < flags.size(); ++id) {
    if (!flags[id]) continue;
    int c = (int)(id / R3);
    int local = (int)(id - (size_t)c * R3);
    int z = local / R2, y = (local - z * R2) / R, x = local % R;
    if (n >= max_caches) { ++dropped; continue; } /* SURVEY B.5 */
    uint8_t* e = (uint8_t*)entries + (size_t)n * entry_stride;
    std::memset(e, 0, entry_stride);
    const drv_cav_cascade& k = vi->AddressVolumeCascades[c];
    float pos[3] = {(float)x * k.WorldVoxelSize + k.Min[0], (float)y * k.WorldVoxelSize + k.Min[1],
                    (float)z * k.WorldVoxelSize + k.Min[2]}; /* cacheGather.comp:65 */
    std::memcpy(e, pos, 12);
    atlas[(size_t)x + (size_t)c * R + atlasW * ((size_t)y + (size_t)R * z)] = n + 1; /* :88 */
    ++n;
  }
  if (counter) { /* cachePrepareLighting.comp:8-14 */
    counter->NumCacheLightingThreadGroupsX = (n + DRV_LIGHTING_THREADS_PER_GROUP - 1) / DRV_LIGHTING_THREADS_PER_GROUP;
    counter->NumCacheLightingThreadGroupsY = 1;
    counter->NumCacheLightingThreadGroupsZ = 1;
    counter->TotalLightCacheCount = (int32_t)n;
  }
  if (overflow) *overflow = dropped;
  if (oob_corners) *oob_corners = oob;
  return (int)n;
}
