// ctx.cu — the C-ABI of libdrv_gi (include/drv_gi.h): context lifetime,
// resource ownership and the stage order of the DYN_RADIANCE_VOLUME case of
// Renderer::Draw (rendering/renderer.cpp:539-570). The context plays the role
// glhelper's buffer/texture objects + Renderer's members play in the
// reference (renderer.hpp:227-352).
#include "ctx.h"
#include "../../include/drv_math.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

namespace {
std::string g_create_error;
std::once_flag g_lut_once[16];

uint32_t ilog2(uint32_t v) {
  uint32_t l = 0;
  while (v > 1) { v >>= 1; ++l; }
  return l;
}
bool is_pow2(uint32_t v) { return v && !(v & (v - 1)); }

template <typename T>
cudaError_t dmalloc(T** p, size_t bytes) {
  return cudaMalloc(reinterpret_cast<void**>(p), bytes ? bytes : 16);
}
} // namespace

#define CREATE_CUDA(expr)                                                          \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      g_create_error = std::string(#expr) + ": " + cudaGetErrorString(_e);         \
      drv_destroy(ctx);                                                            \
      return DRV_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

extern "C" const char* drv_version(void) { return "libdrv_gi 0.1 sm_100a"; }

extern "C" const char* drv_last_error(const drv_ctx* ctx) {
  return ctx ? ctx->last_error.c_str() : g_create_error.c_str();
}

extern "C" drv_status drv_create(const drv_config* cfg, drv_ctx** out) {
  if (!cfg || !out) { g_create_error = "drv_create: null argument"; return DRV_ERR_INVALID; }
  *out = nullptr;
  const drv_config& c = *cfg;
  if (c.max_cache_count == 0 || c.max_cache_count > (1u << 24) || c.cav_resolution > 256 || c.cav_cascades < 1 || c.cav_cascades > DRV_MAX_CASCADES || c.cav_resolution < 8 ||
      (c.cav_resolution % 8) != 0 || (c.sh_order != 1 && c.sh_order != 2) || c.backbuffer_width == 0 ||
      c.backbuffer_height == 0 || (uint64_t)c.backbuffer_width * c.backbuffer_height > (1ull << 30) /* 32-bit pixel indices */ ||
      c.max_lights > DRV_MAX_LIGHTS || !is_pow2(c.voxel_resolution) ||
      c.voxel_resolution < 16 || (c.max_lights > 0 && !is_pow2(c.max_rsm_resolution))) {
    g_create_error = "drv_create: invalid configuration";
    return DRV_ERR_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    g_create_error = "drv_create: no CUDA device (libdrv_gi has no CPU fallback)";
    return DRV_ERR_NO_DEVICE;
  }
  if (c.device < 0 || c.device >= ndev) { g_create_error = "drv_create: bad device ordinal"; return DRV_ERR_INVALID; }
  drv_ctx* ctx = new drv_ctx();
  ctx->cfg = c;
  ctx->device = c.device;
  CREATE_CUDA(cudaSetDevice(c.device));
  cudaDeviceProp prop;
  CREATE_CUDA(cudaGetDeviceProperties(&prop, c.device));
  ctx->num_sms = prop.multiProcessorCount;
  if (c.stream) {
    ctx->stream = reinterpret_cast<cudaStream_t>(c.stream);
  } else {
    CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  std::call_once(g_lut_once[c.device & 15], drv_impl_upload_srgb_lut);

  ctx->entry_stride = c.sh_order == 2 ? 128 : 64;
  // "Maximum size per cache": the buffer is max * 128 B whatever the mode (renderer.cpp:266-269)
  ctx->num_cells = c.cav_cascades * c.cav_resolution * c.cav_resolution * c.cav_resolution;
  const size_t entries_bytes = (size_t)c.max_cache_count * 128 + kSyncBytes + ctx->num_cells + 16;
  CREATE_CUDA(dmalloc(&ctx->entries, entries_bytes));
  CREATE_CUDA(cudaMemsetAsync(ctx->entries, 0, entries_bytes, ctx->stream));
  ctx->sync_flags = reinterpret_cast<uint32_t*>(ctx->entries + (size_t)c.max_cache_count * 128);
  ctx->cell_flags = ctx->entries + (size_t)c.max_cache_count * 128 + kSyncBytes;
  CREATE_CUDA(dmalloc(&ctx->counter, sizeof(drv_cache_counter)));
  CREATE_CUDA(cudaMemsetAsync(ctx->counter, 0, sizeof(drv_cache_counter), ctx->stream));
  CREATE_CUDA(dmalloc(&ctx->stats, 2 * sizeof(uint32_t)));
  CREATE_CUDA(cudaMemsetAsync(ctx->stats, 0, 2 * sizeof(uint32_t), ctx->stream));
  ctx->num_cells = c.cav_cascades * c.cav_resolution * c.cav_resolution * c.cav_resolution;
  CREATE_CUDA(dmalloc(&ctx->atlas, (size_t)ctx->num_cells * sizeof(uint32_t))); // renderer.cpp:1179
  CREATE_CUDA(cudaMemsetAsync(ctx->atlas, 0, (size_t)ctx->num_cells * sizeof(uint32_t), ctx->stream));
  ctx->num_scan_blocks = (ctx->num_cells + 4095) / 4096;
  CREATE_CUDA(dmalloc(&ctx->scan_words, (size_t)ctx->num_scan_blocks * sizeof(unsigned long long)));
  CREATE_CUDA(cudaMemsetAsync(ctx->scan_words, 0, (size_t)ctx->num_scan_blocks * sizeof(unsigned long long), ctx->stream));
  CREATE_CUDA(dmalloc(&ctx->scan_epoch, 4 * sizeof(uint32_t)));
  {
    const uint32_t init[4] = {1u, 0u, 0u, 0u};
    CREATE_CUDA(cudaMemcpyAsync(ctx->scan_epoch, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    CREATE_CUDA(cudaStreamSynchronize(ctx->stream));
  }

  // voxel volumes (voxelization.cpp:56-66): target (1 level) + persistent (full chain)
  const uint32_t vr = c.voxel_resolution;
  ctx->voxel_levels = ilog2(vr) + 1; // glhelper/texture.cpp:54-71
  ctx->voxel_chain_bytes = voxel_level_offset_bytes(vr, ctx->voxel_levels);
  CREATE_CUDA(dmalloc(&ctx->voxel_chain, ctx->voxel_chain_bytes + 16));
  CREATE_CUDA(cudaMemsetAsync(ctx->voxel_chain, 0, ctx->voxel_chain_bytes + 16, ctx->stream));
  CREATE_CUDA(dmalloc(&ctx->voxel_target, (size_t)vr * vr * vr));
  CREATE_CUDA(cudaMemsetAsync(ctx->voxel_target, 0, (size_t)vr * vr * vr, ctx->stream));
  if (ctx->voxel_levels > 15) { g_create_error = "drv_create: voxel_resolution too large"; drv_destroy(ctx); return DRV_ERR_INVALID; }
  for (uint32_t l = 0, r = vr; l < ctx->voxel_levels; ++l, r >>= 1) {
    ctx->voxel_record_offset[l] = (uint32_t)ctx->voxel_record_count;
    const uint64_t d = (uint64_t)r + 1 + 2 * kVoxelRecordPad;
    ctx->voxel_record_count += d * d * d;
  }
  if (ctx->voxel_record_count >= (1ull << 31)) { g_create_error = "drv_create: voxel_resolution too large"; drv_destroy(ctx); return DRV_ERR_INVALID; }
  if (c.indirect_shadow) {
    CREATE_CUDA(dmalloc(&ctx->voxel_records, ctx->voxel_record_count * sizeof(uint2)));
    CREATE_CUDA(cudaMemsetAsync(ctx->voxel_records, 0, ctx->voxel_record_count * sizeof(uint2), ctx->stream));
  }

  for (uint32_t l = 0; l < c.max_lights; ++l) {
    LightState& S = ctx->lights[l];
    const size_t texels = (size_t)c.max_rsm_resolution * c.max_rsm_resolution;
    const size_t mip_texels = texels / 3 + 16; // sum of levels >= 1 < texels / 3
    CREATE_CUDA(dmalloc(&S.flux_mips, mip_texels * 8));
    CREATE_CUDA(dmalloc(&S.normal_mips, mip_texels * 4));
    CREATE_CUDA(dmalloc(&S.depth_mips, mip_texels * 4));
    CREATE_CUDA(dmalloc(&S.vpls, texels * sizeof(drv_vpl)));
    CREATE_CUDA(dmalloc(&S.blocks, texels * sizeof(drv_shadow_block)));
    CREATE_CUDA(dmalloc(&S.vpls_live, texels * sizeof(drv_vpl)));
    CREATE_CUDA(dmalloc(&S.chunk_counts, (texels / 256 + 2) * sizeof(uint32_t)));
    CREATE_CUDA(dmalloc(&S.block_live, texels));
  }
  if (c.indirect_specular) { // environment-map atlas (Renderer::AllocateCacheData, renderer.cpp:253, 282)
    uint32_t S = c.specular_per_cache_size ? c.specular_per_cache_size : 16u;
    if (!is_pow2(S) || S < 2 || S > 16 || c.specular_fill_holes_level > ilog2(S)) {
      g_create_error = "drv_create: specular_per_cache_size must be a power of two in 2..16 and specular_fill_holes_level <= log2 of it";
      drv_destroy(ctx);
      return DRV_ERR_INVALID;
    }
    drv_constant tmp;
    memset(&tmp, 0, sizeof(tmp));
    drv::packSpecular(&tmp, c.max_cache_count, S);
    if ((uint64_t)tmp.SpecularEnvmapNumCachesPerDimension * tmp.SpecularEnvmapNumCachesPerDimension < c.max_cache_count) {
      g_create_error = "drv_create: max_cache_count does not fit a 16384^2 specular environment-map atlas (renderer.cpp:256-263)";
      drv_destroy(ctx);
      return DRV_ERR_CAPACITY;
    }
    ctx->spec_S = S;
    ctx->spec_total = (uint32_t)tmp.SpecularEnvmapTotalSize;
    ctx->spec_levels = ilog2(S) + 1; // renderer.cpp:282
    size_t texels = 0;
    for (uint32_t l = 0, r = ctx->spec_total; l < ctx->spec_levels; ++l, r >>= 1) { ctx->spec_level_offset[l] = (uint32_t)texels; texels += (size_t)r * r; }
    CREATE_CUDA(dmalloc(&ctx->spec_mips, texels * sizeof(uint32_t)));
    CREATE_CUDA(cudaMemsetAsync(ctx->spec_mips, 0, texels * sizeof(uint32_t), ctx->stream));
    CREATE_CUDA(dmalloc(&ctx->spec_patches, (size_t)c.max_cache_count * (S + 1) * (S + 1) * sizeof(uint32_t)));
    float lut[256];
    for (int v = 0; v < 256; ++v) {
      const double cc = (double)v / 255.0;
      lut[v] = (float)((cc <= 0.04045) ? cc / 12.92 : pow((cc + 0.055) / 1.055, 2.4));
    }
    CREATE_CUDA(dmalloc(&ctx->srgb_lut_dev, sizeof(lut)));
    CREATE_CUDA(cudaMemcpy(ctx->srgb_lut_dev, lut, sizeof(lut), cudaMemcpyHostToDevice));
  }
  CREATE_CUDA(dmalloc(&ctx->cone_work, 4 * sizeof(uint32_t)));
  CREATE_CUDA(cudaMemsetAsync(ctx->cone_work, 0, 4 * sizeof(uint32_t), ctx->stream));
  CREATE_CUDA(dmalloc(&ctx->gather_tickets, ((size_t)c.max_cache_count / 64 + 2) * sizeof(uint32_t)));
  CREATE_CUDA(cudaMemsetAsync(ctx->gather_tickets, 0, ((size_t)c.max_cache_count / 64 + 2) * sizeof(uint32_t), ctx->stream));
  CREATE_CUDA(dmalloc(&ctx->live_counts, DRV_MAX_LIGHTS * sizeof(uint32_t)));
  CREATE_CUDA(cudaMemsetAsync(ctx->live_counts, 0, DRV_MAX_LIGHTS * sizeof(uint32_t), ctx->stream));
  CREATE_CUDA(dmalloc(&ctx->ndc_xy, ((size_t)c.backbuffer_width + c.backbuffer_height) * sizeof(float)));
  if (drv_impl_build_ndc_tables(ctx) != DRV_OK) { g_create_error = ctx->last_error; drv_destroy(ctx); return DRV_ERR_CUDA; }
  for (int s = 0; s < DRV_STAGE_COUNT; ++s) {
    CREATE_CUDA(cudaEventCreate(&ctx->ev_begin[s]));
    CREATE_CUDA(cudaEventCreate(&ctx->ev_end[s]));
  }
  CREATE_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = ctx;
  return DRV_OK;
}

extern "C" void drv_destroy(drv_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < 8; ++r) {
    if (ctx->peer_entries[r]) cudaIpcCloseMemHandle(ctx->peer_entries[r]);
    if (ctx->peer_hdr[r]) cudaIpcCloseMemHandle(ctx->peer_hdr[r]);
  }
  cudaFree(ctx->entries); cudaFree(ctx->counter); cudaFree(ctx->stats); cudaFree(ctx->atlas);
  cudaFree(ctx->scan_words); cudaFree(ctx->scan_epoch); cudaFree(ctx->voxel_chain); cudaFree(ctx->voxel_target); cudaFree(ctx->voxel_records);
  cudaFree(ctx->partials); cudaFree(ctx->shadow_table); cudaFree(ctx->st_depth); cudaFree(ctx->st_normal); cudaFree(ctx->st_diffuse);
  cudaFree(ctx->hdr16); cudaFree(ctx->ndc_xy);
  for (auto& S : ctx->lights) {
    cudaFree(S.flux_mips); cudaFree(S.normal_mips); cudaFree(S.depth_mips); cudaFree(S.vpls); cudaFree(S.blocks);
    cudaFree(S.vpls_live); cudaFree(S.chunk_counts); cudaFree(S.block_live);
    cudaFree(S.st_flux); cudaFree(S.st_normal); cudaFree(S.st_depth);
  }
  for (int s = 0; s < DRV_STAGE_COUNT; ++s) {
    if (ctx->ev_begin[s]) cudaEventDestroy(ctx->ev_begin[s]);
    if (ctx->ev_end[s]) cudaEventDestroy(ctx->ev_end[s]);
  }
  cudaFree(ctx->live_counts); cudaFree(ctx->cone_work); cudaFree(ctx->gather_tickets); cudaFree(ctx->gather_trace);
  cudaFree(ctx->spec_mips); cudaFree(ctx->spec_patches); cudaFree(ctx->srgb_lut_dev);
  if (ctx->frame_graph) cudaGraphExecDestroy(ctx->frame_graph);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  if (ctx->side2) cudaStreamDestroy(ctx->side2);
  if (ctx->ev_join2) cudaEventDestroy(ctx->ev_join2);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
  if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
  for (auto& e : ctx->ev_rsm) if (e) cudaEventDestroy(e);
  for (auto& e : ctx->ev_band_in) if (e) cudaEventDestroy(e);
  for (auto& e : ctx->ev_band_done) if (e) cudaEventDestroy(e);
  for (auto& e : ctx->ev_band_out) if (e) cudaEventDestroy(e);
  if (ctx->ev_lit) cudaEventDestroy(ctx->ev_lit);
  if (ctx->ev_depth) cudaEventDestroy(ctx->ev_depth);
  if (ctx->ev_frame_start) cudaEventDestroy(ctx->ev_frame_start);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

#define NEED_CTX() do { if (!ctx) return DRV_ERR_INVALID; cudaSetDevice(ctx->device); } while (0)
// calls that change something a kernel takes as an argument invalidate the recorded frame graph: MUTATES() for
// anything that can change a launch shape or a scratch-buffer size (the next graph frame runs eagerly first),
// MUTATES_UNIFORMS() for the per-frame uniform blocks (moving camera): kernel arguments only, so the recorded
// graph is re-captured and patched in place with cudaGraphExecUpdate — no eager frame, no re-instantiation
#define MUTATES() do { ctx->state_gen++; ctx->shape_gen++; } while (0)
#define MUTATES_UNIFORMS() do { ctx->state_gen++; } while (0)

extern "C" drv_status drv_set_constant(drv_ctx* ctx, const drv_constant* b) {
  NEED_CTX();
  MUTATES_UNIFORMS();
  if (!b) return ctx->fail(DRV_ERR_INVALID, "drv_set_constant: null block");
  ctx->constant = *b;
  ctx->have_constant = true;
  return DRV_OK;
}
extern "C" drv_status drv_set_per_frame(drv_ctx* ctx, const drv_per_frame* b) {
  NEED_CTX();
  MUTATES_UNIFORMS();
  if (!b) return ctx->fail(DRV_ERR_INVALID, "drv_set_per_frame: null block");
  ctx->per_frame = *b;
  ctx->have_per_frame = true;
  return DRV_OK;
}
extern "C" drv_status drv_set_volume_info(drv_ctx* ctx, const drv_volume_info* b) {
  NEED_CTX();
  MUTATES_UNIFORMS();
  if (!b) return ctx->fail(DRV_ERR_INVALID, "drv_set_volume_info: null block");
  ctx->volume = *b;
  ctx->have_volume = true;
  return DRV_OK;
}
extern "C" drv_status drv_set_light_count(drv_ctx* ctx, uint32_t n) {
  NEED_CTX();
  MUTATES();
  if (n > ctx->cfg.max_lights) return ctx->fail(DRV_ERR_INVALID, "drv_set_light_count: more lights than max_lights");
  ctx->num_lights = n;
  return DRV_OK;
}
extern "C" drv_status drv_set_spot_light(drv_ctx* ctx, uint32_t light, const drv_spot_light* b) {
  NEED_CTX();
  MUTATES();
  if (!b || light >= ctx->cfg.max_lights) return ctx->fail(DRV_ERR_INVALID, "drv_set_spot_light: bad light index");
  ctx->lights[light].block = *b;
  ctx->lights[light].block_set = true;
  return DRV_OK;
}

extern "C" drv_status drv_bind_gbuffer(drv_ctx* ctx, const float* depth, const int16_t* normal, const uint8_t* diffuse,
                                       uint32_t w, uint32_t h) {
  NEED_CTX();
  MUTATES();
  if (!depth || w != ctx->cfg.backbuffer_width || h != ctx->cfg.backbuffer_height)
    return ctx->fail(DRV_ERR_INVALID, "drv_bind_gbuffer: resolution differs from the configured backbuffer");
  ctx->gb_depth = depth; ctx->gb_normal = normal; ctx->gb_diffuse = diffuse;
  ctx->gb_w = w; ctx->gb_h = h;
  return DRV_OK;
}

extern "C" drv_status drv_bind_gbuffer_material(drv_ctx* ctx, const uint8_t* roughness_metallic_rg8) {
  NEED_CTX();
  MUTATES();
  ctx->gb_rough_metal = roughness_metallic_rg8;
  return DRV_OK;
}

extern "C" drv_status drv_prepare_specular_envmaps(drv_ctx* ctx) {
  NEED_CTX();
  return drv_impl_prepare_specular(ctx);
}

extern "C" drv_status drv_bind_rsm(drv_ctx* ctx, uint32_t light, const uint16_t* flux, const int16_t* normal,
                                   const uint16_t* depth, uint32_t res) {
  NEED_CTX();
  MUTATES();
  if (light >= ctx->cfg.max_lights || !flux || !normal || !depth || !is_pow2(res) || res > ctx->cfg.max_rsm_resolution)
    return ctx->fail(DRV_ERR_INVALID, "drv_bind_rsm: bad light index or resolution");
  LightState& S = ctx->lights[light];
  S.flux0 = flux; S.normal0 = normal; S.depth0 = depth; S.rsm_res = res;
  S.rsm_bound = true;
  S.vpls_external = false;
  return DRV_OK;
}

extern "C" drv_status drv_prepare_rsm(drv_ctx* ctx, uint32_t light) {
  NEED_CTX();
  if (light >= ctx->cfg.max_lights) return ctx->fail(DRV_ERR_INVALID, "drv_prepare_rsm: bad light index");
  return drv_impl_prepare_rsm(ctx, light);
}

extern "C" drv_status drv_voxelize(drv_ctx* ctx, const float* tri_pos, uint32_t num_tris, const float world[16],
                                   float adaption, uint32_t flags) {
  NEED_CTX();
  if ((num_tris && !tri_pos) || !world) return ctx->fail(DRV_ERR_INVALID, "drv_voxelize: null argument");
  return drv_impl_voxelize(ctx, tri_pos, num_tris, world, adaption, flags);
}

extern "C" drv_status drv_set_voxel_volume(drv_ctx* ctx, const uint8_t* level0) {
  NEED_CTX();
  if (!level0) return ctx->fail(DRV_ERR_INVALID, "drv_set_voxel_volume: null volume");
  return drv_impl_set_voxel_volume(ctx, level0);
}

extern "C" drv_status drv_allocate_caches(drv_ctx* ctx) {
  NEED_CTX();
  return drv_impl_allocate(ctx);
}

extern "C" drv_status drv_light_caches(drv_ctx* ctx) {
  NEED_CTX();
  ctx->stage_begin(DRV_STAGE_LIGHT_CACHES);
  for (uint32_t l = 0; l < ctx->num_lights; ++l) {
    drv_status st = ctx->lights[l].vpls_external ? drv_impl_compact_vpls(ctx, l) : drv_impl_generate_vpls(ctx, l);
    if (st != DRV_OK) return st;
  }
  drv_status st = drv_impl_gather(ctx);
  if (st == DRV_OK) st = drv_impl_specular_light(ctx); // INDIRECT_SPECULAR: the environment-map atlas (same dispatch in the reference)
  ctx->stage_end(DRV_STAGE_LIGHT_CACHES);
  return st;
}

extern "C" drv_status drv_apply_caches(drv_ctx* ctx, void* hdr_out, uint32_t format) {
  NEED_CTX();
  return drv_impl_apply(ctx, hdr_out, format);
}

extern "C" drv_status drv_apply_caches_rows(drv_ctx* ctx, void* hdr_out, uint32_t format, uint32_t y0, uint32_t y1) {
  NEED_CTX();
  return drv_impl_apply_rows(ctx, hdr_out, format, y0, y1, true);
}

extern "C" drv_status drv_draw(drv_ctx* ctx, void* hdr_out, uint32_t format) {
  NEED_CTX();
  drv_status st = drv_impl_allocate(ctx); // renderer.cpp:550
  if (st != DRV_OK) return st;
  st = drv_light_caches(ctx);             // renderer.cpp:556
  if (st != DRV_OK) return st;
  st = drv_impl_prepare_specular(ctx);    // renderer.cpp:557-558
  if (st != DRV_OK) return st;
  return drv_impl_apply(ctx, hdr_out, format); // renderer.cpp:570
}

// The frame in GPU order: light side (RSM mips, VPLs) on ctx->side || camera side (allocate) on the main
// stream, join, gather, apply.
static drv_status frame_body(drv_ctx* ctx, void* hdr_out, uint32_t format, uint32_t flags) {
  cudaStream_t main_stream = ctx->stream;
  DRV_CUDA(cudaEventRecord(ctx->ev_fork, main_stream));
  drv_status st = DRV_OK;
  const bool vox = (flags & DRV_FRAME_VOXELIZE) && ctx->cfg.indirect_shadow;
  if (vox) { // VoxelizeScene on its own stream
    DRV_CUDA(cudaStreamWaitEvent(ctx->side2, ctx->ev_fork, 0));
    ctx->stream = ctx->side2;
    st = drv_impl_voxelize(ctx, ctx->scene_tris, ctx->scene_num_tris, ctx->scene_world, ctx->scene_adaption,
                           DRV_VOXELIZE_CLEAR | DRV_VOXELIZE_FINISH);
    cudaError_t e2 = cudaEventRecord(ctx->ev_join2, ctx->side2);
    ctx->stream = main_stream;
    if (st != DRV_OK) return st;
    if (e2 != cudaSuccess) return ctx->fail(DRV_ERR_CUDA, "drv_draw_frame: event record failed");
  }
  DRV_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
  ctx->stream = ctx->side; // the stage implementations launch on ctx->stream
  for (uint32_t l = 0; l < ctx->num_lights && st == DRV_OK; ++l) {
    if (ctx->lights[l].vpls_external) { st = drv_impl_compact_vpls(ctx, l); continue; }
    if (flags & DRV_FRAME_PREPARE_RSM) st = drv_impl_prepare_rsm(ctx, l, true);
    if (st == DRV_OK) st = drv_impl_generate_vpls(ctx, l);
  }
  cudaError_t e = cudaEventRecord(ctx->ev_join, ctx->side);
  ctx->stream = main_stream;
  if (st != DRV_OK) return st;
  if (e != cudaSuccess) return ctx->fail(DRV_ERR_CUDA, "drv_draw_frame: event record failed");
  // renderer.cpp:550. Sharded frame: every rank marks its band of pixel rows and stores the flags to all ranks
  // (idempotent byte stores over NVLink = the all-reduce of the mark phase); after the barrier every rank holds the
  // complete flag set and runs the deterministic scan + compact itself (identical indices, no communication)
  const bool sharded = drv_peers_complete(ctx);
  st = drv_impl_allocate_mark(ctx, sharded); // sharded: + flag push + the cross-GPU barrier, in one launch
  if (st != DRV_OK) return st;
  // sharded: no SH clear here — every entry's SH is overwritten on every rank by its owner's gather epilogue, so
  // a fast peer's stores cannot be wiped by this rank's (later) compaction and no second barrier is needed
  st = drv_impl_allocate_compact(ctx, !sharded);
  if (st != DRV_OK) return st;
  DRV_CUDA(cudaStreamWaitEvent(main_stream, ctx->ev_join, 0));
  if (vox) DRV_CUDA(cudaStreamWaitEvent(main_stream, ctx->ev_join2, 0));
  ctx->stage_begin(DRV_STAGE_LIGHT_CACHES);
  st = drv_impl_gather(ctx, sharded);   // renderer.cpp:556
  if (st == DRV_OK) st = drv_impl_specular_light(ctx);
  ctx->stage_end(DRV_STAGE_LIGHT_CACHES);
  if (st != DRV_OK) return st;
  st = drv_impl_prepare_specular(ctx);  // renderer.cpp:557-558
  if (st != DRV_OK) return st;
  // ... and all peers' stores have landed before anybody applies
  if (sharded && (st = drv_impl_peer_barrier(ctx)) != DRV_OK) return st;
  if (flags & (DRV_FRAME_APPLY_OWN_ROWS | DRV_FRAME_GATHER_IMAGE)) {
    const uint32_t H = ctx->cfg.backbuffer_height, band = (H + ctx->shard_world - 1) / ctx->shard_world;
    const uint32_t y0 = ctx->shard_rank * band, y1 = y0 + band < H ? y0 + band : H;
    if (flags & DRV_FRAME_GATHER_IMAGE) {
      // this rank's band goes straight into rank 0's target (NVLink stores from the apply kernel); the closing
      // barrier orders them before anything rank 0 enqueues after the frame
      void* target = ctx->shard_rank == 0 ? ctx->hdr16 : ctx->peer_hdr[0]; // validated by drv_draw_frame
      st = drv_impl_apply_rows(ctx, target, format, y0 < H ? y0 : H, y1, true);
      if (st != DRV_OK) return st;
      return drv_impl_peer_barrier(ctx);
    }
    return drv_impl_apply_rows(ctx, hdr_out, format, y0 < H ? y0 : H, y1, true);
  }
  return drv_impl_apply(ctx, hdr_out, format); // renderer.cpp:570
}

extern "C" drv_status drv_bind_scene(drv_ctx* ctx, const float* tri_pos, uint32_t num_tris, const float world[16], float adaption) {
  NEED_CTX();
  MUTATES();
  if ((num_tris && !tri_pos) || !world) return ctx->fail(DRV_ERR_INVALID, "drv_bind_scene: null argument");
  ctx->scene_tris = tri_pos;
  ctx->scene_num_tris = num_tris;
  memcpy(ctx->scene_world, world, sizeof(ctx->scene_world));
  ctx->scene_adaption = adaption;
  return DRV_OK;
}

extern "C" drv_status drv_draw_frame(drv_ctx* ctx, void* hdr_out, uint32_t format, uint32_t flags) {
  NEED_CTX();
  if (!hdr_out && !(flags & DRV_FRAME_GATHER_IMAGE)) return ctx->fail(DRV_ERR_INVALID, "drv_draw_frame: null output");
  // checked before anything is enqueued: a rank that bails out later would leave its peers in a barrier, and a
  // missing mapping would turn into stores through a null-based pointer
  if (ctx->shard_world > 1 && ctx->peers_open && !drv_peers_complete(ctx))
    return ctx->fail(DRV_ERR_NOT_BOUND, "drv_draw_frame: sharded context, but not every peer's entries buffer has been "
                                        "imported (drv_import_peer_entries for each rank != own)");
  if (flags & DRV_FRAME_GATHER_IMAGE) {
    const bool sharded = drv_peers_complete(ctx);
    if (!sharded || !(ctx->shard_rank == 0 ? ctx->hdr16 : ctx->peer_hdr[0]) || format != DRV_HDR_RGBA16F_WRITE)
      return ctx->fail(DRV_ERR_NOT_BOUND, "drv_draw_frame: DRV_FRAME_GATHER_IMAGE needs a sharded context, rank 0's target "
                                          "(drv_export_hdr_ipc / drv_import_peer_hdr) and DRV_HDR_RGBA16F_WRITE");
  }
  if ((flags & DRV_FRAME_VOXELIZE) && ctx->cfg.indirect_shadow && ctx->scene_num_tris && !ctx->scene_tris)
    return ctx->fail(DRV_ERR_NOT_BOUND, "drv_draw_frame: DRV_FRAME_VOXELIZE needs drv_bind_scene");
  if (!ctx->side) {
    DRV_CUDA(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
    DRV_CUDA(cudaStreamCreateWithFlags(&ctx->side2, cudaStreamNonBlocking));
    DRV_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    DRV_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    DRV_CUDA(cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming));
  }
  const bool want_graph = (flags & DRV_FRAME_GRAPH) && !ctx->timers;
  if (!want_graph) {
    drv_status st = frame_body(ctx, hdr_out, format, flags);
    if (st == DRV_OK) ctx->warm_gen = ctx->shape_gen;
    return st;
  }
  const bool valid = ctx->frame_graph && ctx->graph_gen == ctx->state_gen && ctx->graph_out == hdr_out &&
                     ctx->graph_format == format && ctx->graph_flags == flags;
  if (!valid) {
    if (ctx->warm_gen != ctx->shape_gen) {
      // first frame of a new SHAPE runs eagerly: it sizes the scratch buffers (no allocation may happen while
      // a stream is being captured) and reports binding errors the ordinary way
      if (ctx->frame_graph) { cudaGraphExecDestroy(ctx->frame_graph); ctx->frame_graph = nullptr; }
      drv_status st = frame_body(ctx, hdr_out, format, flags);
      if (st == DRV_OK) ctx->warm_gen = ctx->shape_gen;
      return st;
    }
    const uint64_t launches0 = ctx->launches;
    DRV_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    drv_status st = frame_body(ctx, hdr_out, format, flags);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    ctx->graph_launches = ctx->launches - launches0;
    ctx->launches = launches0; // nothing has run yet
    if (st != DRV_OK) { if (graph) cudaGraphDestroy(graph); return st; }
    if (e != cudaSuccess || !graph) {
      cudaGetLastError();
      return ctx->fail(DRV_ERR_CUDA, std::string("drv_draw_frame: capture failed: ") + cudaGetErrorString(e));
    }
    // same topology, new kernel arguments (uniform blocks of a moving camera, another output pointer): patch the
    // instantiated graph in place; anything else: instantiate afresh
    bool updated = false;
    if (ctx->frame_graph && ctx->graph_flags == flags && ctx->graph_format == format) {
      cudaGraphExecUpdateResultInfo info;
      if (cudaGraphExecUpdate(ctx->frame_graph, graph, &info) == cudaSuccess) { updated = true; ctx->graph_updates++; }
      else cudaGetLastError();
    }
    if (!updated) {
      if (ctx->frame_graph) { cudaGraphExecDestroy(ctx->frame_graph); ctx->frame_graph = nullptr; }
      e = cudaGraphInstantiate(&ctx->frame_graph, graph, 0);
      if (e != cudaSuccess) {
        cudaGraphDestroy(graph);
        ctx->frame_graph = nullptr;
        return ctx->fail(DRV_ERR_CUDA, std::string("drv_draw_frame: instantiate failed: ") + cudaGetErrorString(e));
      }
      ctx->graph_instantiations++;
    }
    cudaGraphDestroy(graph);
    ctx->graph_gen = ctx->state_gen;
    ctx->graph_out = hdr_out;
    ctx->graph_format = format;
    ctx->graph_flags = flags;
  }
  DRV_CUDA(cudaGraphLaunch(ctx->frame_graph, ctx->stream));
  ctx->launches += ctx->graph_launches;
  return DRV_OK;
}

extern "C" drv_status drv_graph_stats(drv_ctx* ctx, uint64_t* instantiations, uint64_t* updates) {
  NEED_CTX();
  if (instantiations) *instantiations = ctx->graph_instantiations;
  if (updates) *updates = ctx->graph_updates;
  return DRV_OK;
}

extern "C" drv_status drv_get_buffers(drv_ctx* ctx, drv_buffers* out) {
  NEED_CTX();
  if (!out) return ctx->fail(DRV_ERR_INVALID, "drv_get_buffers: null argument");
  memset(out, 0, sizeof(*out));
  out->entries = ctx->entries;
  out->entry_stride = ctx->entry_stride;
  out->max_cache_count = ctx->cfg.max_cache_count;
  out->counter = ctx->counter;
  out->cav_atlas = ctx->atlas;
  out->cav_width = ctx->cfg.cav_cascades * ctx->cfg.cav_resolution;
  out->cav_height = out->cav_depth = ctx->cfg.cav_resolution;
  out->voxel_chain = ctx->voxel_chain;
  out->voxel_target = ctx->voxel_target;
  out->voxel_resolution = ctx->cfg.voxel_resolution;
  out->voxel_levels = ctx->voxel_levels;
  out->voxel_chain_bytes = ctx->voxel_chain_bytes;
  for (uint32_t l = 0; l < ctx->cfg.max_lights; ++l) {
    out->vpls[l] = ctx->lights[l].vpls;
    out->shadow_blocks[l] = ctx->lights[l].blocks;
    out->rsm_flux_mips[l] = ctx->lights[l].flux_mips;
    out->rsm_normal_mips[l] = ctx->lights[l].normal_mips;
    out->rsm_depth_mips[l] = ctx->lights[l].depth_mips;
    out->rsm_flux0[l] = ctx->lights[l].flux0;
    out->rsm_normal0[l] = ctx->lights[l].normal0;
    out->rsm_depth0[l] = ctx->lights[l].depth0;
  }
  out->hdr16 = ctx->hdr16;
  out->specular_mips = ctx->spec_mips;
  out->specular_total_size = ctx->spec_total;
  out->specular_levels = ctx->spec_levels;
  return DRV_OK;
}

extern "C" uint64_t drv_rsm_level_offset(uint32_t res, uint32_t level) { return rsm_level_offset_texels(res, level); }
extern "C" uint64_t drv_voxel_level_offset(uint32_t res, uint32_t level) { return voxel_level_offset_bytes(res, level); }

extern "C" drv_status drv_active_cache_count(drv_ctx* ctx, uint32_t* count, uint32_t* overflow, uint32_t* oob) {
  NEED_CTX();
  drv_cache_counter c;
  uint32_t stats[2];
  DRV_CUDA(cudaMemcpyAsync(&c, ctx->counter, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
  DRV_CUDA(cudaMemcpyAsync(stats, ctx->stats, sizeof(stats), cudaMemcpyDeviceToHost, ctx->stream));
  DRV_CUDA(cudaStreamSynchronize(ctx->stream));
  if (count) *count = (uint32_t)c.TotalLightCacheCount;
  if (overflow) *overflow = stats[0];
  if (oob) *oob = stats[1];
  if (ctx->shard_world > 1 && ctx->peers_open) { // a cross-GPU barrier of the frame gave up: the entries are incomplete
    drv_status ps = drv_peer_status(ctx, nullptr, nullptr);
    if (ps != DRV_OK) return ps;
  }
  return stats[0] ? DRV_ERR_CAPACITY : DRV_OK;
}

extern "C" drv_status drv_peer_status(drv_ctx* ctx, uint32_t* timed_out_epoch, uint32_t* missing_rank) {
  NEED_CTX();
  uint32_t w[12] = {0};
  DRV_CUDA(cudaMemcpyAsync(w, ctx->sync_flags, sizeof(w), cudaMemcpyDeviceToHost, ctx->stream));
  DRV_CUDA(cudaStreamSynchronize(ctx->stream));
  if (timed_out_epoch) *timed_out_epoch = w[8];
  if (missing_rank) *missing_rank = w[10];
  if (w[8] != 0u)
    return ctx->fail(DRV_ERR_PEER, "cross-GPU barrier " + std::to_string(w[8]) + " timed out waiting for rank " +
                                       std::to_string(w[10]) + ": the frame's entries / image are incomplete (drv_peer_reset)");
  return DRV_OK;
}

extern "C" drv_status drv_peer_reset(drv_ctx* ctx) {
  NEED_CTX();
  DRV_CUDA(cudaMemsetAsync(ctx->sync_flags, 0, kSyncBytes, ctx->stream));
  DRV_CUDA(cudaStreamSynchronize(ctx->stream));
  return DRV_OK;
}

extern "C" drv_status drv_live_vpl_counts(drv_ctx* ctx, uint32_t* counts) {
  NEED_CTX();
  if (!counts) return ctx->fail(DRV_ERR_INVALID, "drv_live_vpl_counts: null argument");
  DRV_CUDA(cudaMemcpyAsync(counts, ctx->live_counts, DRV_MAX_LIGHTS * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  DRV_CUDA(cudaStreamSynchronize(ctx->stream));
  for (uint32_t l = ctx->num_lights; l < DRV_MAX_LIGHTS; ++l) counts[l] = 0;
  return DRV_OK;
}

extern "C" drv_status drv_set_synthetic_entries(drv_ctx* ctx, const float* pos, uint32_t n) {
  NEED_CTX();
  MUTATES();
  if (n && !pos) return ctx->fail(DRV_ERR_INVALID, "drv_set_synthetic_entries: null positions");
  return drv_impl_set_synthetic_entries(ctx, pos, n);
}

extern "C" drv_status drv_set_vpls(drv_ctx* ctx, uint32_t light, const drv_vpl* vpls, uint32_t n) {
  NEED_CTX();
  MUTATES();
  if (light >= ctx->cfg.max_lights || !vpls) return ctx->fail(DRV_ERR_INVALID, "drv_set_vpls: bad argument");
  if ((size_t)n > (size_t)ctx->cfg.max_rsm_resolution * ctx->cfg.max_rsm_resolution)
    return ctx->fail(DRV_ERR_CAPACITY, "drv_set_vpls: more VPLs than max_rsm_resolution^2");
  LightState& S = ctx->lights[light];
  DRV_CUDA(cudaMemcpyAsync(S.vpls, vpls, (size_t)n * sizeof(drv_vpl), cudaMemcpyDefault, ctx->stream));
  S.num_vpls = n;
  S.vpls_external = true;
  return DRV_OK;
}

extern "C" void drv_shard_range(uint32_t count, uint32_t rank, uint32_t world, uint32_t* begin, uint32_t* end) {
  if (world == 0) world = 1;
  uint32_t groups = (count + 63u) / 64u;
  uint32_t g0 = (uint32_t)(((unsigned long long)groups * rank) / world);
  uint32_t g1 = (uint32_t)(((unsigned long long)groups * (rank + 1)) / world);
  uint32_t b = g0 * 64u < count ? g0 * 64u : count;
  uint32_t e = g1 * 64u < count ? g1 * 64u : count;
  if (begin) *begin = b;
  if (end) *end = e;
}

extern "C" drv_status drv_set_shard(drv_ctx* ctx, uint32_t rank, uint32_t world) {
  NEED_CTX();
  MUTATES();
  if (world == 0 || rank >= world || world > 8) return ctx->fail(DRV_ERR_INVALID, "drv_set_shard: need rank < world <= 8");
  ctx->shard_rank = rank;
  ctx->shard_world = world;
  return DRV_OK;
}

extern "C" drv_status drv_set_shard_interleave(drv_ctx* ctx, uint32_t enable) {
  NEED_CTX();
  MUTATES();
  ctx->shard_interleave = enable != 0;
  return DRV_OK;
}

extern "C" void drv_shard_entry(uint32_t local, uint32_t rank, uint32_t world, uint32_t* entry) {
  if (world == 0) world = 1;
  if (entry) *entry = (((local >> 6) * world + rank) << 6) | (local & 63u);
}

extern "C" drv_status drv_export_entries_ipc(drv_ctx* ctx, uint8_t handle[DRV_IPC_HANDLE_BYTES]) {
  NEED_CTX();
  static_assert(sizeof(cudaIpcMemHandle_t) == DRV_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  DRV_CUDA(cudaIpcGetMemHandle(&h, ctx->entries));
  memcpy(handle, &h, sizeof(h));
  return DRV_OK;
}

extern "C" drv_status drv_import_peer_entries(drv_ctx* ctx, uint32_t peer_rank, const uint8_t handle[DRV_IPC_HANDLE_BYTES]) {
  NEED_CTX();
  MUTATES();
  if (peer_rank >= 8 || peer_rank == ctx->shard_rank) return ctx->fail(DRV_ERR_INVALID, "drv_import_peer_entries: bad peer rank");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return ctx->fail(DRV_ERR_PEER, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
  ctx->peer_entries[peer_rank] = p;
  ctx->peers_open = true;
  return DRV_OK;
}

extern "C" drv_status drv_export_hdr_ipc(drv_ctx* ctx, uint8_t handle[DRV_IPC_HANDLE_BYTES]) {
  NEED_CTX();
  MUTATES();
  const size_t bytes = (size_t)ctx->cfg.backbuffer_width * ctx->cfg.backbuffer_height * 8;
  if (!ctx->hdr16) {
    DRV_CUDA(cudaMalloc(&ctx->hdr16, bytes));
    DRV_CUDA(cudaMemsetAsync(ctx->hdr16, 0, bytes, ctx->stream));
    DRV_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  cudaIpcMemHandle_t h;
  DRV_CUDA(cudaIpcGetMemHandle(&h, ctx->hdr16));
  memcpy(handle, &h, sizeof(h));
  return DRV_OK;
}

extern "C" drv_status drv_import_peer_hdr(drv_ctx* ctx, uint32_t peer_rank, const uint8_t handle[DRV_IPC_HANDLE_BYTES]) {
  NEED_CTX();
  MUTATES();
  if (peer_rank >= 8 || peer_rank == ctx->shard_rank) return ctx->fail(DRV_ERR_INVALID, "drv_import_peer_hdr: bad peer rank");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return ctx->fail(DRV_ERR_PEER, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
  ctx->peer_hdr[peer_rank] = p;
  return DRV_OK;
}

extern "C" drv_status drv_peer_barrier(drv_ctx* ctx) {
  NEED_CTX();
  if (ctx->shard_world <= 1) return DRV_OK;
  if (!ctx->peers_open) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_peer_barrier: peers not imported");
  for (uint32_t r = 0; r < ctx->shard_world; ++r)
    if (r != ctx->shard_rank && !ctx->peer_entries[r]) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_peer_barrier: a peer is missing");
  return drv_impl_peer_barrier(ctx);
}

static const char* kStageNames[DRV_STAGE_COUNT] = {"VoxelizeScene", "VoxelBlendMipMap", "AllocateCaches", "LightCaches",
                                                   "ApplyCaches",   "PrepareRSM",       "GatherKernel",     "ConeKernel"};
extern "C" const char* drv_stage_name(drv_stage s) { return (s >= 0 && s < DRV_STAGE_COUNT) ? kStageNames[s] : "?"; }

extern "C" drv_status drv_enable_stage_timers(drv_ctx* ctx, int enable) {
  NEED_CTX();
  MUTATES();
  ctx->timers = enable != 0;
  for (auto& v : ctx->ev_valid) v = false;
  return DRV_OK;
}

extern "C" drv_status drv_stage_ms(drv_ctx* ctx, drv_stage s, float* ms) {
  NEED_CTX();
  if (s < 0 || s >= DRV_STAGE_COUNT || !ms) return ctx->fail(DRV_ERR_INVALID, "drv_stage_ms: bad argument");
  if (!ctx->ev_valid[s]) { *ms = 0.0f; return ctx->fail(DRV_ERR_NOT_BOUND, "drv_stage_ms: stage not recorded"); }
  DRV_CUDA(cudaEventSynchronize(ctx->ev_end[s]));
  DRV_CUDA(cudaEventElapsedTime(ms, ctx->ev_begin[s], ctx->ev_end[s]));
  return DRV_OK;
}

extern "C" uint64_t drv_kernel_launches(const drv_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- host-buffer convenience (end-to-end measurement) --------------------------------
extern "C" drv_status drv_upload_gbuffer(drv_ctx* ctx, const float* depth, const int16_t* normal, const uint8_t* diffuse,
                                         uint32_t w, uint32_t h) {
  NEED_CTX();
  MUTATES();
  if (!depth || !normal || !diffuse || w != ctx->cfg.backbuffer_width || h != ctx->cfg.backbuffer_height)
    return ctx->fail(DRV_ERR_INVALID, "drv_upload_gbuffer: bad argument");
  const size_t px = (size_t)w * h;
  if (!ctx->st_depth) {
    DRV_CUDA(dmalloc(&ctx->st_depth, px * 4));
    DRV_CUDA(dmalloc(&ctx->st_normal, px * 4));
    DRV_CUDA(dmalloc(&ctx->st_diffuse, px * 4));
  }
  DRV_CUDA(cudaMemcpyAsync(ctx->st_depth, depth, px * 4, cudaMemcpyHostToDevice, ctx->stream));
  DRV_CUDA(cudaMemcpyAsync(ctx->st_normal, normal, px * 4, cudaMemcpyHostToDevice, ctx->stream));
  DRV_CUDA(cudaMemcpyAsync(ctx->st_diffuse, diffuse, px * 4, cudaMemcpyHostToDevice, ctx->stream));
  return drv_bind_gbuffer(ctx, ctx->st_depth, ctx->st_normal, ctx->st_diffuse, w, h);
}

extern "C" drv_status drv_upload_rsm(drv_ctx* ctx, uint32_t light, const uint16_t* flux, const int16_t* normal,
                                     const uint16_t* depth, uint32_t res) {
  NEED_CTX();
  MUTATES();
  if (light >= ctx->cfg.max_lights || !flux || !normal || !depth || !is_pow2(res) || res > ctx->cfg.max_rsm_resolution)
    return ctx->fail(DRV_ERR_INVALID, "drv_upload_rsm: bad argument");
  LightState& S = ctx->lights[light];
  const size_t tx = (size_t)ctx->cfg.max_rsm_resolution * ctx->cfg.max_rsm_resolution;
  if (!S.st_flux) {
    DRV_CUDA(dmalloc(&S.st_flux, tx * 8));
    DRV_CUDA(dmalloc(&S.st_normal, tx * 4));
    DRV_CUDA(dmalloc(&S.st_depth, tx * 4));
  }
  const size_t n = (size_t)res * res;
  DRV_CUDA(cudaMemcpyAsync(S.st_flux, flux, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  DRV_CUDA(cudaMemcpyAsync(S.st_normal, normal, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  DRV_CUDA(cudaMemcpyAsync(S.st_depth, depth, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  return drv_bind_rsm(ctx, light, S.st_flux, S.st_normal, S.st_depth, res);
}

extern "C" drv_status drv_draw_to_host(drv_ctx* ctx, void* hdr_host) {
  NEED_CTX();
  if (!hdr_host) return ctx->fail(DRV_ERR_INVALID, "drv_draw_to_host: null output");
  const size_t bytes = (size_t)ctx->cfg.backbuffer_width * ctx->cfg.backbuffer_height * 8;
  if (!ctx->hdr16) DRV_CUDA(cudaMalloc(&ctx->hdr16, bytes));
  DRV_CUDA(cudaMemsetAsync(ctx->hdr16, 0, bytes, ctx->stream)); // glClear(GL_COLOR_BUFFER_BIT), renderer.cpp:562
  drv_status st = drv_draw(ctx, ctx->hdr16, DRV_HDR_RGBA16F_ADD);
  if (st != DRV_OK) return st;
  DRV_CUDA(cudaMemcpyAsync(hdr_host, ctx->hdr16, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  DRV_CUDA(cudaStreamSynchronize(ctx->stream));
  return DRV_OK;
}

// Diagnostics: when the pieces of the last drv_draw_host_frame finished, in ms after its first copy was queued.
extern "C" drv_status drv_debug_host_frame_timeline(drv_ctx* ctx, float* out, uint32_t capacity, uint32_t* bands) {
  NEED_CTX();
  const uint32_t nb = ctx->host_frame_bands;
  if (!out || !bands || !nb) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_debug_host_frame_timeline: no host frame drawn yet");
  if (!ctx->host_timeline)
    return ctx->fail(DRV_ERR_NOT_BOUND, "drv_debug_host_frame_timeline: enable the stage timers before the first drv_draw_host_frame");
  if (capacity < 3 + 3 * nb) return ctx->fail(DRV_ERR_INVALID, "drv_debug_host_frame_timeline: capacity < 3 + 3 * bands");
  DRV_CUDA(cudaStreamSynchronize(ctx->copy_out));
  DRV_CUDA(cudaStreamSynchronize(ctx->stream));
  const uint32_t last_light = ctx->num_lights ? ctx->num_lights - 1 : 0;
  DRV_CUDA(cudaEventElapsedTime(out + 0, ctx->ev_frame_start, ctx->ev_rsm[last_light]));
  DRV_CUDA(cudaEventElapsedTime(out + 1, ctx->ev_frame_start, ctx->ev_depth));
  DRV_CUDA(cudaEventElapsedTime(out + 2, ctx->ev_frame_start, ctx->ev_lit));
  for (uint32_t b = 0; b < nb; ++b) {
    DRV_CUDA(cudaEventElapsedTime(out + 3 + 3 * b, ctx->ev_frame_start, ctx->ev_band_in[b]));
    DRV_CUDA(cudaEventElapsedTime(out + 4 + 3 * b, ctx->ev_frame_start, ctx->ev_band_done[b]));
    DRV_CUDA(cudaEventElapsedTime(out + 5 + 3 * b, ctx->ev_frame_start, ctx->ev_band_out[b]));
  }
  *bands = nb;
  return DRV_OK;
}

// ---- pipelined end-to-end frame -------------------------------------------------------
extern "C" drv_status drv_draw_host_frame(drv_ctx* ctx, const drv_host_frame* f) {
  NEED_CTX();
  if (!f || !f->depth || !f->normal_rg16i || !f->diffuse_srgb8x || !f->hdr_out || f->num_lights > ctx->cfg.max_lights)
    return ctx->fail(DRV_ERR_INVALID, "drv_draw_host_frame: bad argument");
  const uint32_t W = ctx->cfg.backbuffer_width, H = ctx->cfg.backbuffer_height;
  const size_t px = (size_t)W * H;
  if (!ctx->copy_in) {
    DRV_CUDA(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    DRV_CUDA(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    // the events that order the streams also give the frame's timeline (drv_debug_host_frame_timeline) when the
    // stage timers are on at this point; otherwise they are created without time stamps (cheaper to record)
    ctx->host_timeline = ctx->timers;
    const unsigned ef = ctx->host_timeline ? cudaEventDefault : cudaEventDisableTiming;
    for (auto& e : ctx->ev_rsm) DRV_CUDA(cudaEventCreateWithFlags(&e, ef));
    for (auto& e : ctx->ev_band_in) DRV_CUDA(cudaEventCreateWithFlags(&e, ef));
    for (auto& e : ctx->ev_band_done) DRV_CUDA(cudaEventCreateWithFlags(&e, ef));
    for (auto& e : ctx->ev_band_out) DRV_CUDA(cudaEventCreateWithFlags(&e, ef));
    DRV_CUDA(cudaEventCreateWithFlags(&ctx->ev_depth, ef));
    DRV_CUDA(cudaEventCreateWithFlags(&ctx->ev_lit, ef));
    DRV_CUDA(cudaEventCreateWithFlags(&ctx->ev_frame_start, ef));
  }
  if (!ctx->st_depth) {
    DRV_CUDA(dmalloc(&ctx->st_depth, px * 4));
    DRV_CUDA(dmalloc(&ctx->st_normal, px * 4));
    DRV_CUDA(dmalloc(&ctx->st_diffuse, px * 4));
  }
  if (!ctx->hdr16) DRV_CUDA(cudaMalloc(&ctx->hdr16, px * 8));
  const size_t rsm_cap = (size_t)ctx->cfg.max_rsm_resolution * ctx->cfg.max_rsm_resolution;
  for (uint32_t l = 0; l < f->num_lights; ++l) {
    LightState& S = ctx->lights[l];
    const uint32_t res = f->rsm_resolution[l];
    if (!f->rsm_flux_rgbx16f[l] || !f->rsm_normal_rg16i[l] || !f->rsm_depthlinsq_rg16f[l] || !is_pow2(res) ||
        res > ctx->cfg.max_rsm_resolution)
      return ctx->fail(DRV_ERR_INVALID, "drv_draw_host_frame: bad RSM");
    if (!S.st_flux) {
      DRV_CUDA(dmalloc(&S.st_flux, rsm_cap * 8));
      DRV_CUDA(dmalloc(&S.st_normal, rsm_cap * 4));
      DRV_CUDA(dmalloc(&S.st_depth, rsm_cap * 4));
    }
  }
  // band boundaries (whole 8-row apply blocks). Every band costs ~8 us of copy / launch latency and its copy-out
  // competes with the copy-in of the next one; measured at 1080p (tools/e2e_probe.py): 1 band 1.15 ms, 2: 1.05,
  // 4: 1.03, 8: 1.06, 16: 1.11, 32: 1.28 — the default is 4
  uint32_t band_y[34];
  uint32_t bands = f->bands ? f->bands : 4;
  if (bands > 32) bands = 32;
  {
    const uint32_t rows_per_band = (((H + bands - 1) / bands) + 7) & ~7u;
    bands = (H + rows_per_band - 1) / rows_per_band;
    for (uint32_t b = 0; b <= bands; ++b) band_y[b] = std::min(H, b * rows_per_band);
  }

  // the copy stream starts after everything already queued on the context's stream (previous users of the staging images)
  DRV_CUDA(cudaEventRecord(ctx->ev_frame_start, ctx->stream));
  DRV_CUDA(cudaStreamWaitEvent(ctx->copy_in, ctx->ev_frame_start, 0));
  // H2D order = the order in which the stages need their inputs: RSMs, depth, then normal + albedo bands
  for (uint32_t l = 0; l < f->num_lights; ++l) {
    LightState& S = ctx->lights[l];
    const size_t n = (size_t)f->rsm_resolution[l] * f->rsm_resolution[l];
    DRV_CUDA(cudaMemcpyAsync(S.st_flux, f->rsm_flux_rgbx16f[l], n * 8, cudaMemcpyHostToDevice, ctx->copy_in));
    DRV_CUDA(cudaMemcpyAsync(S.st_normal, f->rsm_normal_rg16i[l], n * 4, cudaMemcpyHostToDevice, ctx->copy_in));
    DRV_CUDA(cudaMemcpyAsync(S.st_depth, f->rsm_depthlinsq_rg16f[l], n * 4, cudaMemcpyHostToDevice, ctx->copy_in));
    DRV_CUDA(cudaEventRecord(ctx->ev_rsm[l], ctx->copy_in));
  }
  DRV_CUDA(cudaMemcpyAsync(ctx->st_depth, f->depth, px * 4, cudaMemcpyHostToDevice, ctx->copy_in));
  DRV_CUDA(cudaEventRecord(ctx->ev_depth, ctx->copy_in));
  for (uint32_t b = 0; b < bands; ++b) {
    const size_t y0 = band_y[b], y1 = band_y[b + 1];
    const size_t off = y0 * W * 4, bytes = (y1 - y0) * W * 4;
    DRV_CUDA(cudaMemcpyAsync((uint8_t*)ctx->st_normal + off, (const uint8_t*)f->normal_rg16i + off, bytes,
                             cudaMemcpyHostToDevice, ctx->copy_in));
    DRV_CUDA(cudaMemcpyAsync(ctx->st_diffuse + off, f->diffuse_srgb8x + off, bytes, cudaMemcpyHostToDevice, ctx->copy_in));
    DRV_CUDA(cudaEventRecord(ctx->ev_band_in[b], ctx->copy_in));
  }
  // compute, on the context's stream
  // glClear(GL_COLOR_BUFFER_BIT) (renderer.cpp:562) is fused into the apply pass: DRV_HDR_RGBA16F_WRITE
  drv_status st = drv_set_light_count(ctx, f->num_lights);
  if (st != DRV_OK) return st;
  for (uint32_t l = 0; l < f->num_lights; ++l) {
    LightState& S = ctx->lights[l];
    DRV_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_rsm[l], 0));
    st = drv_bind_rsm(ctx, l, S.st_flux, S.st_normal, S.st_depth, f->rsm_resolution[l]);
    if (st == DRV_OK) st = drv_impl_prepare_rsm(ctx, l, true);
    if (st != DRV_OK) return st;
  }
  DRV_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_depth, 0));
  st = drv_bind_gbuffer(ctx, ctx->st_depth, ctx->st_normal, ctx->st_diffuse, W, H);
  if (st == DRV_OK) st = drv_impl_allocate(ctx);
  if (st == DRV_OK) st = drv_light_caches(ctx);
  if (st != DRV_OK) return st;
  if (ctx->host_timeline) DRV_CUDA(cudaEventRecord(ctx->ev_lit, ctx->stream));
  ctx->host_frame_bands = bands;
  ctx->stage_begin(DRV_STAGE_APPLY_CACHES);
  for (uint32_t b = 0; b < bands; ++b) {
    const uint32_t y0 = band_y[b], y1 = band_y[b + 1];
    DRV_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_band_in[b], 0));
    st = drv_impl_apply_rows(ctx, ctx->hdr16, DRV_HDR_RGBA16F_WRITE, y0, y1, false);
    if (st != DRV_OK) return st;
    DRV_CUDA(cudaEventRecord(ctx->ev_band_done[b], ctx->stream));
    DRV_CUDA(cudaStreamWaitEvent(ctx->copy_out, ctx->ev_band_done[b], 0));
    const size_t off = (size_t)y0 * W * 8, bytes = (size_t)(y1 - y0) * W * 8;
    DRV_CUDA(cudaMemcpyAsync((uint8_t*)f->hdr_out + off, (const uint8_t*)ctx->hdr16 + off, bytes, cudaMemcpyDeviceToHost,
                             ctx->copy_out));
    if (ctx->host_timeline) DRV_CUDA(cudaEventRecord(ctx->ev_band_out[b], ctx->copy_out));
  }
  ctx->stage_end(DRV_STAGE_APPLY_CACHES);
  DRV_CUDA(cudaStreamSynchronize(ctx->copy_out));
  DRV_CUDA(cudaStreamSynchronize(ctx->stream));
  return DRV_OK;
}
