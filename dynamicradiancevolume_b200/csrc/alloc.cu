// alloc.cu — stage 1: cache allocation (≙ Renderer::AllocateCaches,
// rendering/renderer.cpp:951-992; shader/cacheGather.comp:93-164;
// shader/cachePrepareLighting.comp:8-14).
//
// B200-first redesign: the reference locks atlas texels with CAS and takes
// entry indices from one global atomic counter (nondeterministic order, two
// contended atomics per new cell). Here:
//   mark    one thread per pixel on the reference's own 16x16 tiles evaluates
//           the identical trigger predicate (including its tile-edge quirk,
//           SURVEY B.1) and sets a byte flag per corner cell — plain
//           idempotent stores, no atomics;
//   count   per 2048-cell block popcount of the flags;
//   scan    one block: exclusive scan of the block counts, writes the
//           LightCacheCounter / indirect args (cachePrepareLighting.comp);
//   compact per cell: index = #flagged cells with smaller linear id; writes
//           the atlas texel (0 or index+1, which also replaces the reference's
//           per-frame atlas clear), the entry position and zeroed SH.
// The index order (ascending linear cell id) is deterministic and identical
// on every GPU, so multi-GPU runs replicate this stage without communication.
#include "ctx.h"
#include "device_math.cuh"

#include <algorithm>

using namespace drvk;

namespace {

struct AllocParams {
  int W, H, R, C;
  int transitions;
  float zone;
  float ivp[16];
  drv_cav_cascade casc[DRV_MAX_CASCADES];
  // 1 / WorldVoxelSize when the voxel size is a power of two (then x / size == x * (1 / size) bit for bit),
  // else 0: take the IEEE division
  float inv_voxel[DRV_MAX_CASCADES];
};

// scan + compact kernel: 256 threads, GROUPS 8-cell groups per thread. 2 (4096 cells per tile) for frame-sized
// grids; 4 / 8 when 4096-cell tiles would be more than the GPU holds at once (a second wave of tiles can only start
// when first-wave blocks retire, and those wait for their own look-back: 4 x 128^3 cells = 2048 tiles took 47 us)
constexpr int kScanThreads = 256;

// ndc tables: ((i + .5) / N) * 2 - 1, cacheGather.comp:113-116
__global__ void ndc_table_kernel(float* __restrict__ out, int W, int H) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W) out[i] = ex_sub(ex_mul(ex_div((float)i + 0.5f, (float)W), 2.0f), 1.0f);
  else if (i < W + H) out[i] = ex_sub(ex_mul(ex_div((float)(i - W) + 0.5f, (float)H), 2.0f), 1.0f);
}

// lightcache.glsl:109-122
__device__ __forceinline__ int compute_cascade(const AllocParams& p, F3 wp) {
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

// lightcache.glsl:125-134, as the allocation uses it (cacheGather.comp:133-134): only whether the transition is > 0.
// saturate(1 - minDist / d) > 0  <=>  fl(minDist / d) < 1  <=>  minDist < d  for a correctly rounded division by a
// positive finite d (the largest float below d divided by d rounds to at most 1 - 2^-24; 1 - q is then exact and
// positive) — so the division is only evaluated when d is not a positive number (zone <= 0, NaN).
__device__ __forceinline__ bool in_cascade_transition(const AllocParams& p, F3 wp, int c) {
  const drv_cav_cascade& k = p.casc[c];
  float ax = ex_sub(k.DecisionMax[0], wp.x), ay = ex_sub(k.DecisionMax[1], wp.y), az = ex_sub(k.DecisionMax[2], wp.z);
  float bx = ex_sub(wp.x, k.DecisionMin[0]), by = ex_sub(wp.y, k.DecisionMin[1]), bz = ex_sub(wp.z, k.DecisionMin[2]);
  float minDist = fminf(fminf(fminf(ax, ay), az), fminf(fminf(bx, by), bz));
  const float d = ex_mul(k.WorldVoxelSize, p.zone);
  if (d > 0.0f && d <= 3.0e38f) return minDist < d;
  return saturatef(ex_sub(1.0f, ex_div(minDist, d))) > 0.0f;
}

// cacheGather.comp:20-30; the cell is kept as (x,y,z) packed 10:10:10 next to its linear id
struct Cell { int id; int x, y, z; };
__device__ __forceinline__ Cell cache_cell(const AllocParams& p, F3 wp, int c) {
  const drv_cav_cascade& k = p.casc[c];
  const float inv = p.inv_voxel[c];
  float fx = ex_sub(wp.x, k.Min[0]), fy = ex_sub(wp.y, k.Min[1]), fz = ex_sub(wp.z, k.Min[2]);
  if (inv != 0.0f) { fx = ex_mul(fx, inv); fy = ex_mul(fy, inv); fz = ex_mul(fz, inv); }
  else { fx = ex_div(fx, k.WorldVoxelSize); fy = ex_div(fy, k.WorldVoxelSize); fz = ex_div(fz, k.WorldVoxelSize); }
  Cell o;
  o.x = clampi(ex_trunc(fx), 0, p.R - 1);
  o.y = clampi(ex_trunc(fy), 0, p.R - 1);
  o.z = clampi(ex_trunc(fz), 0, p.R - 1);
  o.id = o.x + o.y * p.R + o.z * p.R * p.R + c * p.R * p.R * p.R;
  return o;
}

// The peers' flag arrays of a sharded frame (every rank marks only its band of pixel rows): a flag store is
// idempotent, so the union of all ranks' stores over NVLink IS the all-reduce of the mark phase.
struct MarkTargets {
  uint8_t* flags[8];
  int n;
};

// cacheGather.comp:32-91 with the index assignment deferred to the scan.
__device__ __forceinline__ void mark_corners(const AllocParams& p, const Cell& cell, int c, uint8_t* __restrict__ flags,
                                             uint32_t* __restrict__ oob_accum) {
  const int R = p.R, R2 = R * R;
  uint8_t* f = flags + (uint32_t)(c * R2 * R + cell.x + cell.y * R + cell.z * R2);
  if (cell.x + 1 < R && cell.y + 1 < R && cell.z + 1 < R) {
    // offsets (0,0,0)(0,1,0)(0,0,1)(0,1,1)(1,0,0)(1,1,0)(1,0,1)(1,1,1), cacheGather.comp:34-44. Plain idempotent
    // byte stores: no read-before-write, nothing on the critical path waits for memory.
    f[0] = 1; f[R] = 1; f[R2] = 1; f[R2 + R] = 1;
    f[1] = 1; f[R + 1] = 1; f[R2 + 1] = 1; f[R2 + R + 1] = 1;
    return;
  }
  uint32_t oob = 0; // out-of-range +1 corners are skipped and counted
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ox = i >> 2, oy = i & 1, oz = (i >> 1) & 1;
    if (cell.x + ox >= R || cell.y + oy >= R || cell.z + oz >= R) { ++oob; continue; }
    f[ox + oy * R + oz * R2] = 1;
  }
  atomicAdd(oob_accum, oob);
}

// tile_y0: first 16-row tile of this launch (a sharded frame marks one band of tile rows per rank; the tiles are
// the reference's own whatever the band, so the dedupe predicate sees the same neighbours)
__global__ void __launch_bounds__(256) mark_kernel(AllocParams p, const float* __restrict__ depth,
                                                   const float* __restrict__ ndc_xy, uint8_t* __restrict__ flags,
                                                   uint32_t* __restrict__ oob_accum, int tile_y0) {
  __shared__ int T1[16][17]; // [local x][local y] like cacheList[x][y]; padded against bank conflicts
  __shared__ int T2[16][17];
  const int lx = threadIdx.x, ly = threadIdx.y;
  const int x = blockIdx.x * 16 + lx, y = (blockIdx.y + tile_y0) * 16 + ly;
  Cell own = {-1, 0, 0, 0}, own2 = {-1, 0, 0, 0};
  int casc = -1;
  if (x < p.W && y < p.H) {
    float d = __ldg(depth + (uint32_t)(y * p.W + x)); // < 2^31 pixels: checked at create
    if (d > 0.0001f) {
      F3 wp = ex_unproject(p.ivp, __ldg(ndc_xy + (uint32_t)x), __ldg(ndc_xy + (uint32_t)(p.W + y)), d);
      casc = compute_cascade(p, wp);
      own = cache_cell(p, wp, casc);
      if (p.transitions && casc < p.C - 1) {
        if (in_cascade_transition(p, wp, casc)) own2 = cache_cell(p, wp, casc + 1);
      }
    }
  }
  T1[lx][ly] = own.id;
  if (p.transitions) T2[lx][ly] = own2.id;
  __syncthreads();
  { // cacheGather.comp:142-146, conditions kept literally. Bitwise logic: no short-circuit jumps
    const int ax = max(0, lx - 1), ay = max(0, ly - 1), id = own.id;
    const bool fresh = ((T1[lx][ay] != id) & (T1[ax][ly] != id) & (T1[ax][ay] != id)) | ((ax == lx) & (ay == ly));
    if (fresh & (id != -1)) mark_corners(p, own, casc, flags, oob_accum);
  }
  if (p.transitions) { // cacheGather.comp:147-150 (the mirrored neighbours)
    const int bx = min(15, lx + 1), by = min(15, ly + 1), id = own2.id;
    // (bx == 15 && by == 15) holds for lx >= 14 && ly >= 14 — four threads, not one: the shader's own condition, kept literally
    const bool fresh = ((T2[lx][by] != id) & (T2[bx][ly] != id) & (T2[bx][by] != id)) | ((bx == 15) & (by == 15));
    if (fresh & (id != -1)) mark_corners(p, own2, casc + 1, flags, oob_accum);
  }
}

__device__ __forceinline__ uint32_t nonzero_bytes(uint2 v) { // number of non-zero bytes among 8 flags
  uint32_t a = __vcmpne4(v.x, 0u), b = __vcmpne4(v.y, 0u); // 0xff per non-zero byte
  return (__popc(a) + __popc(b)) >> 3;
}

// ---- scan + compact in one pass ---------------------------------------------------------------------------
// Every block counts the flagged cells of its 2048-cell slice, publishes the count, obtains the number of
// flagged cells in all earlier slices by decoupled look-back over the published states (aggregate / inclusive
// prefix), and writes its atlas texels and entries — count, scan and compact of the classic three-kernel scheme
// in one launch. State words carry a frame epoch kept on the device (so a recorded frame graph replays without
// per-launch arguments, and nothing needs clearing): a word from the previous frame is simply "not yet there".
// The kernel also zeroes the flags it has consumed (the next frame's mark kernel starts from a clean slate
// without a memset) and moves the out-of-range-corner statistic out of its accumulator.
struct ScanState {
  unsigned long long* words; // per block: value | (epoch << 2 | flag) << 32
  uint32_t* epoch;           // [0] epoch of the current frame, [1] blocks done, [2] oob accumulator, [3] tile ticket
  uint32_t* oob_accum;       // mark_kernel's accumulator
};
constexpr uint32_t kFlagAggregate = 1u, kFlagPrefix = 2u;

template <int STRIDE, bool ZERO_SH, int kGroups>
__global__ void __launch_bounds__(kScanThreads) scan_compact_kernel(AllocParams p, uint8_t* __restrict__ flags,
                                                                    uint32_t num_cells, ScanState st, uint32_t max_caches,
                                                                    uint32_t* __restrict__ atlas, uint8_t* __restrict__ entries,
                                                                    drv_cache_counter* __restrict__ counter,
                                                                    uint32_t* __restrict__ stats) {
  constexpr int NW = kScanThreads / 32;
  constexpr int kCellsPerThread = 8 * kGroups;
  __shared__ uint32_t warp_tot[NW];
  __shared__ uint32_t s_prefix, s_tile;
  __shared__ uint32_t lb_sum[NW];   // look-back: per warp, the values up to (and including) its nearest inclusive prefix
  __shared__ uint32_t lb_found[NW]; // ... and whether the warp saw one
  // Tiles are handed out through an atomic ticket, not blockIdx: a block only ever waits for tiles with SMALLER
  // tickets, whose blocks are already running — forward progress does not depend on the dispatch order of CTAs.
  if (threadIdx.x == 0) s_tile = atomicAdd(st.epoch + 3, 1u);
  __syncthreads();
  const uint32_t tile = s_tile, num_tiles = gridDim.x;
  const uint32_t epoch_raw = *reinterpret_cast<volatile uint32_t*>(st.epoch);
  const uint32_t epoch = epoch_raw & 0x3FFFFFFFu; // 30 bits fit beside the 2 flag bits
  const uint32_t cell0 = (tile * kScanThreads + threadIdx.x) * kCellsPerThread;
  uint2 f[kGroups];
  uint32_t mine = 0;
#pragma unroll
  for (int g = 0; g < kGroups; ++g) {
    f[g] = make_uint2(0u, 0u);
    const uint32_t c0 = cell0 + g * 8;
    if (c0 < num_cells) {
      f[g] = *reinterpret_cast<const uint2*>(flags + c0);
      if (f[g].x | f[g].y) *reinterpret_cast<uint2*>(flags + c0) = make_uint2(0u, 0u); // consumed
    }
    mine += nonzero_bytes(f[g]);
  }
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += t;
  }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint32_t warp_base = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    if (w < (int)(threadIdx.x >> 5)) warp_base += warp_tot[w];
    block_total += warp_tot[w];
  }
  // publish the aggregate, then look back over the predecessors, kScanThreads of them per round (thread 0 = the
  // nearest): a 1080p frame's 128 tiles resolve in a single round of polls instead of a chain of 32-wide windows
  volatile unsigned long long* W = st.words;
  const unsigned long long tag = (unsigned long long)(epoch << 2) << 32;
  if (threadIdx.x == 0)
    W[tile] = (unsigned long long)block_total | tag | ((unsigned long long)(tile == 0 ? kFlagPrefix : kFlagAggregate) << 32);
  uint32_t prefix = 0;
  for (int base = (int)tile - 1; base >= 0; base -= kScanThreads) {
    const int i = base - (int)threadIdx.x;
    uint32_t flag = kFlagPrefix, val = 0; // below tile 0 there is nothing: an inclusive prefix of 0
    if (i >= 0) {
      unsigned long long w;
      do { w = W[i]; } while ((uint32_t)(w >> 34) != epoch || ((uint32_t)(w >> 32) & 3u) == 0u);
      flag = (uint32_t)(w >> 32) & 3u;
      val = (uint32_t)w;
    }
    const uint32_t pm = __ballot_sync(0xffffffffu, flag == kFlagPrefix);
    const int first = __ffs(pm) - 1; // nearest lane of this warp holding an inclusive prefix (-1: none)
    const uint32_t take = (first < 0 || (int)(threadIdx.x & 31) <= first) ? val : 0u;
    const uint32_t wsum = __reduce_add_sync(0xffffffffu, take);
    if ((threadIdx.x & 31) == 0) { lb_sum[threadIdx.x >> 5] = wsum; lb_found[threadIdx.x >> 5] = first >= 0 ? 1u : 0u; }
    __syncthreads();
    bool found = false;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      if (!found) { prefix += lb_sum[w]; found = lb_found[w] != 0u; }
    }
    __syncthreads(); // lb_* are rewritten by the next round
    if (found) break;
  }
  if (threadIdx.x == 0 && tile != 0)
    W[tile] = (unsigned long long)(prefix + block_total) | tag | ((unsigned long long)kFlagPrefix << 32);
  uint32_t index = prefix + warp_base + incl - mine;

  if (tile == num_tiles - 1 && threadIdx.x == 0) { // the last slice knows the total: cachePrepareLighting.comp:8-14
    const uint32_t total = prefix + block_total;
    const uint32_t n = total > max_caches ? max_caches : total; // clamp, report
    stats[0] = total - n;
    counter->NumCacheLightingThreadGroupsX = (n + DRV_LIGHTING_THREADS_PER_GROUP - 1) / DRV_LIGHTING_THREADS_PER_GROUP;
    counter->NumCacheLightingThreadGroupsY = 1;
    counter->NumCacheLightingThreadGroupsZ = 1;
    counter->TotalLightCacheCount = (int)n;
  }
  if (threadIdx.x == 0) { // frame bookkeeping: the block that finishes last opens the next epoch
    __threadfence();
    if (atomicAdd(st.epoch + 1, 1u) == num_tiles - 1) {
      stats[1] = *st.oob_accum;
      *st.oob_accum = 0u;
      st.epoch[1] = 0u;
      st.epoch[3] = 0u; // the ticket counter
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(st.epoch) = (epoch_raw + 1u) == 0u ? 1u : epoch_raw + 1u;
    }
  }

  const int R = p.R, R2 = R * R, R3 = R2 * R;
  const uint32_t atlasW = (uint32_t)(R * p.C);
#pragma unroll
  for (int g = 0; g < kGroups; ++g) {
    const uint32_t c0 = cell0 + g * 8;
    if (c0 >= num_cells) break;
    // R is a multiple of 8 (checked at create), so the 8 cells of a group share (c, z, y)
    const int c = c0 / R3;
    const int local = c0 - c * R3;
    const int z = local / R2;
    const int y = (local - z * R2) / R;
    const int x0 = local - z * R2 - y * R;
    const drv_cav_cascade& k = p.casc[c];
    uint32_t out[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t word = i < 4 ? f[g].x : f[g].y;
      bool set = ((word >> ((i & 3) * 8)) & 0xffu) != 0u;
      uint32_t v = 0u;
      if (set) {
        if (index < max_caches) {
          v = index + 1u; // +1 since zero means "cleared", cacheGather.comp:88
          float4* e = reinterpret_cast<float4*>(entries + (size_t)index * STRIDE);
          // Position = cell * WorldVoxelSize + Min, cacheGather.comp:65 (mul and add rounded separately)
          e[0] = make_float4(ex_add(ex_mul((float)(x0 + i), k.WorldVoxelSize), k.Min[0]),
                             ex_add(ex_mul((float)y, k.WorldVoxelSize), k.Min[1]),
                             ex_add(ex_mul((float)z, k.WorldVoxelSize), k.Min[2]), 0.0f);
          if (ZERO_SH) { // cacheGather.comp:68-83 (a sharded frame leaves it to the gather, which overwrites)
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 1; q < STRIDE / 16; ++q) e[q] = zero;
          }
        }
        ++index;
      }
      out[i] = v;
    }
    uint4* dst = reinterpret_cast<uint4*>(atlas + (size_t)(x0 + c * R) + (size_t)atlasW * ((size_t)y + (size_t)R * z));
    dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
    dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
  }
}

// drv_set_synthetic_entries: positions -> entries with zeroed SH.
template <int STRIDE>
__global__ void synthetic_entries_kernel(const float4* __restrict__ pos, uint32_t n, uint8_t* __restrict__ entries,
                                         drv_cache_counter* __restrict__ counter) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    counter->NumCacheLightingThreadGroupsX = (n + 63) / 64;
    counter->NumCacheLightingThreadGroupsY = 1;
    counter->NumCacheLightingThreadGroupsZ = 1;
    counter->TotalLightCacheCount = (int)n;
  }
  if (i >= n) return;
  float4 p = pos[i];
  p.w = 0.0f;
  float4* e = reinterpret_cast<float4*>(entries + (size_t)i * STRIDE);
  e[0] = p;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int q = 1; q < STRIDE / 16; ++q) e[q] = zero;
}

} // namespace

static drv_status alloc_params(drv_ctx* ctx, AllocParams& p) {
  if (!ctx->have_constant || !ctx->have_per_frame || !ctx->have_volume)
    return ctx->fail(DRV_ERR_NOT_BOUND, "drv_allocate_caches: uniform blocks not set");
  if (!ctx->gb_depth) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_allocate_caches: g-buffer not bound");
  p.W = ctx->constant.BackbufferResolution[0];
  p.H = ctx->constant.BackbufferResolution[1];
  p.R = ctx->constant.AddressVolumeResolution;
  p.C = ctx->constant.NumAddressVolumeCascades;
  if (p.W != (int)ctx->gb_w || p.H != (int)ctx->gb_h || p.R != (int)ctx->cfg.cav_resolution ||
      p.C != (int)ctx->cfg.cav_cascades)
    return ctx->fail(DRV_ERR_INVALID, "drv_allocate_caches: Constant block disagrees with the context configuration");
  p.transitions = ctx->cfg.cascade_transitions ? 1 : 0;
  p.zone = ctx->volume.CAVTransitionZoneSize;
  memcpy(p.ivp, ctx->per_frame.InverseViewProjection, sizeof(p.ivp));
  memcpy(p.casc, ctx->volume.AddressVolumeCascades, sizeof(p.casc));
  for (int c = 0; c < DRV_MAX_CASCADES; ++c) {
    int e = 0;
    const float v = p.casc[c].WorldVoxelSize;
    // exact power of two well inside the normal range: division == multiplication by the reciprocal
    p.inv_voxel[c] = (v > 1e-6f && v < 1e6f && frexpf(v, &e) == 0.5f) ? 1.0f / v : 0.0f;
  }
  return DRV_OK;
}

// Sharded frames: after a rank has marked its band into its own flag array, the flags that are set — a few
// thousand bytes — are stored into every peer's array (8-byte scan, one remote byte store per set flag and peer).
// Marking straight into the peers costs a remote store per triggering PIXEL instead and was slower than not
// sharding at all. Flags a peer has already pushed here are pushed again: harmless, the stores are idempotent.
// The push ends in the cross-GPU barrier of the mark phase (the same epoch protocol as peer_barrier_kernel, gather.cu):
// every block fences its remote stores and takes a ticket; the block that draws the last one announces this rank's
// new epoch to all peers and waits for theirs — one launch instead of push + barrier.
struct PushBarrier {
  uint32_t* own;      // this rank's sync block: [0..7] epochs announced by the ranks, [8] time-out marker, [9] own epoch
  uint32_t* peer[8];  // every rank's sync block (own included)
  uint32_t rank, world;
  uint32_t* done;     // blocks-done ticket (zero between launches)
};
__global__ void __launch_bounds__(256) push_flags_kernel(const uint8_t* __restrict__ flags, uint32_t num_cells, MarkTargets peers,
                                                         PushBarrier B) {
  const uint32_t words = num_cells / 8;
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
    const uint2 f = __ldcg(reinterpret_cast<const uint2*>(flags + (size_t)w * 8));
    if (!(f.x | f.y)) continue;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t word = i < 4 ? f.x : f.y;
      if (((word >> ((i & 3) * 8)) & 0xffu) == 0u) continue;
      for (int t = 0; t < peers.n; ++t) peers.flags[t][(size_t)w * 8 + i] = 1;
    }
  }
  __threadfence_system(); // this thread's remote flag stores are ordered before the ticket
  __shared__ uint32_t s_last;
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(B.done, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last || threadIdx.x >= 32) return;
  // the last block: every block's stores are ordered before this point
  const uint32_t t = threadIdx.x;
  uint32_t epoch = 0;
  if (t == 0) { *B.done = 0u; epoch = B.own[9] + 1u; B.own[9] = epoch; }
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  const bool active = t < B.world && t != B.rank;
  __threadfence_system();
  if (active) *reinterpret_cast<volatile uint32_t*>(B.peer[t] + B.rank) = epoch;
  __syncwarp();
  if (!active) return;
  if (*reinterpret_cast<volatile uint32_t*>(B.own + 8) != 0u) return; // an earlier barrier timed out: do not wait
  const long long t0 = clock64();
  while ((int32_t)(*reinterpret_cast<volatile uint32_t*>(B.own + t) - epoch) < 0) {
    if (clock64() - t0 > 8000000000ll) { B.own[8] = epoch; B.own[10] = t; break; }
  }
  __threadfence_system();
}

// Mark phase. `sharded`: only this rank's band of 16-row tiles, then the set flags are pushed to every peer.
drv_status drv_impl_allocate_mark(drv_ctx* ctx, bool sharded) {
  AllocParams p;
  drv_status st = alloc_params(ctx, p);
  if (st != DRV_OK) return st;
  ctx->stage_begin(DRV_STAGE_ALLOCATE_CACHES);
  // ≙ m_lightCacheCounter->ClearToZero() and the atlas clear (renderer.cpp:969-970): both folded into the
  // scan + compact kernel, which also leaves the cell flags zeroed for the next frame — no memset in the frame
  MarkTargets P;
  memset(&P, 0, sizeof(P));
  int tiles_y = (p.H + 15) / 16, tile_y0 = 0;
  if (sharded) {
    const size_t off = (size_t)ctx->cfg.max_cache_count * 128 + kSyncBytes; // the flags sit behind entries + sync block
    for (uint32_t r = 0; r < ctx->shard_world && r < 8; ++r)
      if (r != ctx->shard_rank) P.flags[P.n++] = (uint8_t*)ctx->peer_entries[r] + off;
    const int band = (tiles_y + (int)ctx->shard_world - 1) / (int)ctx->shard_world;
    tile_y0 = std::min(tiles_y, (int)ctx->shard_rank * band);
    tiles_y = std::min(tiles_y, tile_y0 + band) - tile_y0;
  }
  if (tiles_y > 0) {
    dim3 grid((p.W + 15) / 16, tiles_y); // renderer.cpp:981-985
    mark_kernel<<<grid, dim3(16, 16), 0, ctx->stream>>>(p, ctx->gb_depth, ctx->ndc_xy, ctx->cell_flags, ctx->scan_epoch + 2,
                                                        tile_y0);
    DRV_LAUNCH_CHECK();
  }
  if (sharded && P.n > 0) {
    const uint32_t words = ctx->num_cells / 8;
    const uint32_t blocks = std::min<uint32_t>((words + 255) / 256, (uint32_t)ctx->num_sms * 8);
    PushBarrier B;
    memset(&B, 0, sizeof(B));
    B.own = ctx->sync_flags;
    const size_t sync_off = (size_t)ctx->cfg.max_cache_count * 128;
    for (uint32_t r = 0; r < ctx->shard_world && r < 8; ++r)
      B.peer[r] = r == ctx->shard_rank ? ctx->sync_flags : reinterpret_cast<uint32_t*>((uint8_t*)ctx->peer_entries[r] + sync_off);
    B.rank = ctx->shard_rank;
    B.world = ctx->shard_world;
    B.done = ctx->sync_flags + 12; // a word of the sync block the barrier protocol does not use
    // ... and the cross-GPU barrier that follows the mark phase rides in the same launch
    push_flags_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->cell_flags, ctx->num_cells, P, B);
    DRV_LAUNCH_CHECK();
  }
  return DRV_OK;
}

// Scan + compact phase (replicated on every rank of a sharded frame: deterministic, identical indices).
// zero_sh = false: a sharded drv_draw_frame, whose gather OVERWRITES every entry's SH on every rank — zeroing here
// would race with a faster peer's stores (and would need a barrier of its own).
drv_status drv_impl_allocate_compact(drv_ctx* ctx, bool zero_sh) {
  AllocParams p;
  drv_status st0 = alloc_params(ctx, p);
  if (st0 != DRV_OK) return st0;
  ScanState st;
  st.words = ctx->scan_words;
  st.epoch = ctx->scan_epoch;
  st.oob_accum = ctx->scan_epoch + 2;
#define DRV_SCAN_G(ST, Z, G) scan_compact_kernel<ST, Z, G><<<(ctx->num_cells + 2048u * G - 1u) / (2048u * G), kScanThreads, 0, ctx->stream>>>( \
    p, ctx->cell_flags, ctx->num_cells, st, ctx->cfg.max_cache_count, ctx->atlas, ctx->entries, ctx->counter, ctx->stats)
  // tiles per launch <= what one wave of resident blocks covers (8 blocks of 256 threads per SM). Measured (stage
  // AllocateCaches): 4 x 128^3 cells 154 -> 144 us with 8192-cell tiles (152 with 16384), 2 x 256^3 cells 169 -> 115 us
  // with 16384-cell tiles (119 with 32768)
  const uint32_t wave = (uint32_t)ctx->num_sms * 8u;
  const int groups = ctx->num_scan_blocks <= wave ? 2 : (ctx->num_scan_blocks <= 2 * wave ? 4 : 8);
#define DRV_SCAN(ST, Z) do { if (groups == 2) DRV_SCAN_G(ST, Z, 2); else if (groups == 4) DRV_SCAN_G(ST, Z, 4); else DRV_SCAN_G(ST, Z, 8); } while (0)
  if (ctx->entry_stride == 64) { if (zero_sh) DRV_SCAN(64, true); else DRV_SCAN(64, false); }
  else { if (zero_sh) DRV_SCAN(128, true); else DRV_SCAN(128, false); }
#undef DRV_SCAN_G
#undef DRV_SCAN
  DRV_LAUNCH_CHECK();
  ctx->stage_end(DRV_STAGE_ALLOCATE_CACHES);
  return DRV_OK;
}

drv_status drv_impl_allocate(drv_ctx* ctx) {
  drv_status st = drv_impl_allocate_mark(ctx, false);
  if (st != DRV_OK) return st;
  return drv_impl_allocate_compact(ctx, true);
}

drv_status drv_impl_set_synthetic_entries(drv_ctx* ctx, const float* pos, uint32_t n) {
  if (n > ctx->cfg.max_cache_count) return ctx->fail(DRV_ERR_CAPACITY, "drv_set_synthetic_entries: n > max_cache_count");
  uint32_t blocks = (n + 255) / 256;
  if (blocks == 0) blocks = 1;
  if (ctx->entry_stride == 64)
    synthetic_entries_kernel<64><<<blocks, 256, 0, ctx->stream>>>((const float4*)pos, n, ctx->entries, ctx->counter);
  else
    synthetic_entries_kernel<128><<<blocks, 256, 0, ctx->stream>>>((const float4*)pos, n, ctx->entries, ctx->counter);
  DRV_LAUNCH_CHECK();
  return DRV_OK;
}

drv_status drv_impl_build_ndc_tables(drv_ctx* ctx) {
  const int W = (int)ctx->cfg.backbuffer_width, H = (int)ctx->cfg.backbuffer_height;
  ndc_table_kernel<<<(W + H + 255) / 256, 256, 0, ctx->stream>>>(ctx->ndc_xy, W, H);
  DRV_LAUNCH_CHECK();
  return DRV_OK;
}
