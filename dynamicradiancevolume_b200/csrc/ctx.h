// ctx.h — internal context of libdrv_gi (not part of the C-ABI).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/drv_gi.h"

#define DRV_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      return ctx->fail(DRV_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    }                                                                                    \
  } while (0)

#define DRV_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      return ctx->fail(DRV_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
    }                                                                                    \
    ctx->launches++;                                                                     \
  } while (0)

// Per-light device state.
struct LightState {
  bool block_set = false;   // drv_set_spot_light seen
  bool rsm_bound = false;   // drv_bind_rsm / drv_upload_rsm seen
  bool vpls_external = false; // drv_set_vpls: skip VPL generation
  uint32_t num_vpls = 0;
  drv_spot_light block{};
  const uint16_t* flux0 = nullptr;   // level 0 (borrowed or staging)
  const int16_t* normal0 = nullptr;
  const uint16_t* depth0 = nullptr;
  uint32_t rsm_res = 0;
  uint16_t* flux_mips = nullptr;     // levels >= 1, owned
  int16_t* normal_mips = nullptr;
  uint16_t* depth_mips = nullptr;
  drv_vpl* vpls = nullptr;           // owned, max_rsm_resolution^2 (Morton order; the list parity tests read)
  drv_shadow_block* blocks = nullptr;
  // what the gather streams: the VPLs with non-zero flux, order preserved, each tagged with its shadow-block
  // index (Normal.w); RSM texels that saw no surface carry zero flux and would add exactly zero to every cache
  drv_vpl* vpls_live = nullptr;
  uint32_t* chunk_counts = nullptr;  // live VPLs per 256-VPL chunk (compaction scratch)
  uint8_t* block_live = nullptr;     // 1 = the shadow block has at least one live VPL (cones of dead blocks are skipped)
  // staging for drv_upload_rsm
  uint16_t* st_flux = nullptr;
  int16_t* st_normal = nullptr;
  uint16_t* st_depth = nullptr;
};

struct drv_ctx {
  drv_config cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 0;
  std::string last_error;
  uint64_t launches = 0;

  // uniform blocks (host copies; passed to kernels by value)
  drv_constant constant{};
  drv_per_frame per_frame{};
  drv_volume_info volume{};
  bool have_constant = false, have_per_frame = false, have_volume = false;
  uint32_t num_lights = 0;
  LightState lights[DRV_MAX_LIGHTS];

  // g-buffer (borrowed or staging)
  const float* gb_depth = nullptr;
  const int16_t* gb_normal = nullptr;
  const uint8_t* gb_diffuse = nullptr;
  uint32_t gb_w = 0, gb_h = 0;
  float* st_depth = nullptr;
  int16_t* st_normal = nullptr;
  uint8_t* st_diffuse = nullptr;
  const uint8_t* gb_rough_metal = nullptr; // RG8 plane (drv_bind_gbuffer_material), read only with indirect specular
  void* hdr16 = nullptr; // owned RGBA16F target for drv_draw_to_host
  // ndc_x[x] = ((x + .5) / W) * 2 - 1 and ndc_y[y] likewise (cacheGather.comp:113-116), evaluated once per
  // context with the decision-maths operators instead of two IEEE divisions per pixel per stage
  float* ndc_xy = nullptr; // W floats then H floats

  // allocation
  uint32_t entry_stride = 64;
  uint8_t* entries = nullptr;        // max_cache_count * 128 B (renderer.cpp:266-269)
  drv_cache_counter* counter = nullptr;
  uint32_t* stats = nullptr;         // [0] overflow, [1] oob corners
  uint32_t* atlas = nullptr;
  uint8_t* cell_flags = nullptr;     // one byte per CAV cell, linear-cell-id order; lives behind the entries + sync
                                     // block in the SAME allocation, so a peer that mapped the entries can store flags
  unsigned long long* scan_words = nullptr; // decoupled look-back states of the scan + compact kernel
  uint32_t* scan_epoch = nullptr;           // [0] frame epoch (starts at 1), [1] blocks done, [2] oob-corner accumulator, [3] tile ticket
  uint32_t num_cells = 0, num_scan_blocks = 0;

  // voxels
  uint8_t* voxel_chain = nullptr;
  uint8_t* voxel_target = nullptr;
  uint32_t voxel_levels = 0;
  uint64_t voxel_chain_bytes = 0;
  // gather-ready copy of the chain for the cone tracer: per level (r+1)^3 records of 8 bytes, record
  // (x,y,z), x,y,z in [-1, r-1], = the 2x2x2 clamp-to-edge texel neighbourhood whose lower corner is (x,y,z).
  // One 64-bit load fetches a whole trilinear footprint.
  uint2* voxel_records = nullptr;
  uint32_t voxel_record_offset[16] = {0};
  uint64_t voxel_record_count = 0;

  // gather
  uint32_t* live_counts = nullptr;   // [DRV_MAX_LIGHTS] live VPLs per light (device; the gather reads it there)
  uint32_t* cone_work = nullptr;     // cone_kernel's work queue: [0] next item, [1] CTAs done
  float* shadow_table = nullptr;     // visibility of (VAL block, cache) for the current chunk of caches (cone_kernel)
  size_t shadow_table_floats = 0;
  // layout of the visibility table the last shadowed gather used (the specular pass reads it)
  uint32_t shadow_stride = 0, shadow_chunks = 0, shadow_block_offset[DRV_MAX_LIGHTS] = {0};
  // indirect specular (specular.cu): per-cache patches, the R11F_G11F_B10F atlas with its mip levels
  uint32_t spec_S = 0, spec_total = 0, spec_levels = 0, spec_level_offset[8] = {0};
  uint32_t* spec_patches = nullptr;
  uint32_t* spec_mips = nullptr;
  float* srgb_lut_dev = nullptr;
  uint32_t* gather_tickets = nullptr; // warp-split gather: arrival counter per cache tile (zero between launches)
  unsigned long long* gather_trace = nullptr; // diagnostics (gather_variant bit 18)
  uint32_t gather_trace_ctas = 0;
  float* partials = nullptr;         // split-VPL partial sums
  uint64_t partial_slots = 0;        // capacity in cache slots
  uint32_t shard_rank = 0, shard_world = 1;
  bool shard_interleave = false;     // 64-entry groups dealt round-robin to the ranks instead of contiguous ranges
  void* peer_entries[8] = {nullptr};
  void* peer_hdr[8] = {nullptr};     // peers' context-owned RGBA16F targets (only rank 0's is used)
  bool peers_open = false;
  // cross-GPU barrier flags live right behind the entries in the same allocation (one IPC handle maps both):
  // flags[r] = last epoch rank r has announced to this GPU; flags[8] = time-out marker
  uint32_t* sync_flags = nullptr;

  // host-frame pipeline (drv_draw_host_frame): copy streams + events
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t ev_rsm[DRV_MAX_LIGHTS]{}, ev_depth = nullptr, ev_band_in[32]{}, ev_band_done[32]{}, ev_frame_start = nullptr;
  cudaEvent_t ev_band_out[32]{}, ev_lit = nullptr; // timeline of the last host frame (drv_debug_host_frame_timeline)
  uint32_t host_frame_bands = 0;
  bool host_timeline = false; // the events above carry time stamps (stage timers were on when they were created)

  // drv_draw_frame: light-side stream, fork / join events, recorded frame graph
  cudaStream_t side = nullptr, side2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;
  const float* scene_tris = nullptr; // drv_bind_scene
  uint32_t scene_num_tris = 0;
  float scene_world[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  float scene_adaption = 1.0f;
  uint64_t state_gen = 1;            // bumped by every call that changes a kernel argument
  uint64_t shape_gen = 1;            // ... and this one only by calls that can change launch shapes / scratch sizes
  uint64_t graph_updates = 0, graph_instantiations = 0; // cudaGraphExecUpdate patches / fresh instantiations
  cudaGraphExec_t frame_graph = nullptr;
  uint64_t graph_gen = 0;            // state_gen the graph was recorded at
  void* graph_out = nullptr;
  uint32_t graph_format = 0, graph_flags = 0;
  uint64_t graph_launches = 0;       // kernels per replay
  uint64_t warm_gen = 0;             // shape_gen of the last eager frame (scratch buffers are sized)

  // timers
  bool timers = false;
  cudaEvent_t ev_begin[DRV_STAGE_COUNT]{}, ev_end[DRV_STAGE_COUNT]{};
  bool ev_valid[DRV_STAGE_COUNT]{};

  drv_status fail(drv_status code, const std::string& msg) {
    last_error = msg;
    return code;
  }
  void stage_begin(drv_stage s) {
    if (timers) cudaEventRecord(ev_begin[s], stream);
  }
  void stage_end(drv_stage s) {
    if (timers) { cudaEventRecord(ev_end[s], stream); ev_valid[s] = true; }
  }
};

// A sharded frame stores through every peer's mapping (mark flags, barrier words, finished entries): all of them
// must have been imported.
inline bool drv_peers_complete(const drv_ctx* ctx) {
  if (ctx->shard_world <= 1) return false;
  for (uint32_t r = 0; r < ctx->shard_world && r < 8; ++r)
    if (r != ctx->shard_rank && !ctx->peer_entries[r]) return false;
  return true;
}

// stage implementations (one .cu each)
drv_status drv_impl_allocate(drv_ctx* ctx);
drv_status drv_impl_allocate_mark(drv_ctx* ctx, bool sharded);
drv_status drv_impl_allocate_compact(drv_ctx* ctx, bool zero_sh = true);
drv_status drv_impl_prepare_rsm(drv_ctx* ctx, uint32_t light, bool only_consumed = false);
drv_status drv_impl_generate_vpls(drv_ctx* ctx, uint32_t light);
drv_status drv_impl_compact_vpls(drv_ctx* ctx, uint32_t light, bool counted = false);
drv_status drv_impl_gather(drv_ctx* ctx, bool overwrite = false);
drv_status drv_impl_apply(drv_ctx* ctx, void* out, uint32_t format);
drv_status drv_impl_apply_rows(drv_ctx* ctx, void* out, uint32_t format, uint32_t y_begin, uint32_t y_end, bool timed);
drv_status drv_impl_voxelize(drv_ctx* ctx, const float* tris, uint32_t n, const float* world, float adaption,
                             uint32_t flags);
drv_status drv_impl_set_synthetic_entries(drv_ctx* ctx, const float* pos, uint32_t n);
drv_status drv_impl_set_voxel_volume(drv_ctx* ctx, const uint8_t* level0);
drv_status drv_impl_peer_barrier(drv_ctx* ctx);
drv_status drv_impl_specular_light(drv_ctx* ctx);
drv_status drv_impl_prepare_specular(drv_ctx* ctx);
drv_status drv_impl_apply_specular(drv_ctx* ctx, void* out, uint32_t format, uint32_t y_begin, uint32_t y_end, bool timed,
                                   const float* srgb_lut_dev);
constexpr size_t kSyncBytes = 256;
// Gather-ready voxel records (voxel.cu / voxel_sample.cuh): a level of resolution r holds (r + 1 + 2 kVoxelRecordPad)^3
// records, one per footprint lower corner in [-1 - pad, r - 1 + pad]^3 (clamp to edge baked in), so that a cone whose
// samples stray a few voxels outside the volume — caches and lights ON its boundary walls — needs no index clamp.
// A multiple of 4: the record build reads aligned words.
constexpr uint32_t kVoxelRecordPad = 4;
void drv_impl_upload_srgb_lut();
drv_status drv_impl_build_ndc_tables(drv_ctx* ctx);

inline uint64_t rsm_level_offset_texels(uint32_t res, uint32_t level) {
  uint64_t off = 0;
  for (uint32_t l = 1; l < level; ++l) off += (uint64_t)(res >> l) * (res >> l);
  return off;
}
inline uint64_t voxel_level_offset_bytes(uint32_t res, uint32_t level) {
  uint64_t off = 0;
  for (uint32_t l = 0; l < level; ++l) { off += (uint64_t)res * res * res; res >>= 1; }
  return off;
}
