/*
 * ref_main.cpp — C API of oracle/_ref/libdrv_ref.so: picks the shader variant the reference's host code would have
 * compiled (its #define option switches: INDDIFFUSE_VIA_SH1/2, ADDRESSVOL_CASCADE_TRANSITIONS, INDIRECT_SHADOW;
 * renderer.cpp:159-246) and forwards to the translation unit glsl2cpp.py generated from the UNMODIFIED shader text.
 * TEST INFRASTRUCTURE ONLY: the library exists to pin oracle/ (tests/test_oracle_vs_ref.py).
 */
#include <cstdint>

#include "../../include/drv_gi.h"

#define DECL_ALLOC(v) extern "C" int ref_allocate_##v(const drv_constant*, const drv_per_frame*, const drv_volume_info*, const float*, uint32_t*, void*, uint32_t, drv_cache_counter*, int);
DECL_ALLOC(gather_sh1_n) DECL_ALLOC(gather_sh1_t) DECL_ALLOC(gather_sh2_n) DECL_ALLOC(gather_sh2_t)
#define DECL_LIGHT(v) extern "C" void ref_light_##v(const drv_constant*, const drv_per_frame*, const drv_volume_info*, const drv_spot_light*, uint32_t, const uint16_t* const*, const int16_t* const*, const uint16_t* const* const*, const uint32_t*, const uint8_t*, uint32_t, void*, uint32_t, float* const*, uint32_t*, int);
DECL_LIGHT(light_sh1_n) DECL_LIGHT(light_sh1_s) DECL_LIGHT(light_sh2_n) DECL_LIGHT(light_sh2_s) DECL_LIGHT(light_sh1_n_spec) DECL_LIGHT(light_sh2_s_spec)
#define DECL_APPLY(v) extern "C" void ref_apply_##v(const drv_constant*, const drv_per_frame*, const drv_volume_info*, const float*, const int16_t*, const uint8_t*, const uint32_t*, const void*, uint32_t, float*, const uint8_t*, const uint32_t*, int);
DECL_APPLY(apply_sh1_n) DECL_APPLY(apply_sh1_t) DECL_APPLY(apply_sh2_n) DECL_APPLY(apply_sh2_t) DECL_APPLY(apply_sh1_t_spec) DECL_APPLY(apply_sh2_n_spec)
extern "C" void ref_prepare_prepare(drv_cache_counter*);
extern "C" void ref_voxel_blend_voxelblend(uint8_t*, const uint8_t*, uint32_t, float);
extern "C" void ref_voxel_mips_voxelmipmap(uint8_t*, uint32_t);
extern "C" void ref_rsm_downsample_downsample(const uint16_t*, const int16_t*, const uint16_t*, uint32_t, uint16_t*, int16_t*, uint16_t*);

/* shader/cacheGather.comp + cachePrepareLighting.comp (Renderer::AllocateCaches, renderer.cpp:951-992). `entries`
 * must hold one slot per CAV cell: the shader has no capacity check. Indices come from an atomic counter, i.e. in
 * execution order — compare as a set keyed by cell. */
extern "C" int ref_allocate_caches(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi, int transitions,
                                   int sh_order, const float* depth, uint32_t* atlas, void* entries, uint32_t entry_capacity,
                                   drv_cache_counter* counter, int threads) {
  int n;
  if (sh_order == 2) n = transitions ? ref_allocate_gather_sh2_t(cb, pf, vi, depth, atlas, entries, entry_capacity, counter, threads)
                                     : ref_allocate_gather_sh2_n(cb, pf, vi, depth, atlas, entries, entry_capacity, counter, threads);
  else n = transitions ? ref_allocate_gather_sh1_t(cb, pf, vi, depth, atlas, entries, entry_capacity, counter, threads)
                       : ref_allocate_gather_sh1_n(cb, pf, vi, depth, atlas, entries, entry_capacity, counter, threads);
  ref_prepare_prepare(counter);
  return n;
}

/* shader/cacheLightingRSM.comp, one dispatch per light (Renderer::LightCachesRSM, renderer.cpp:899-933). */
extern "C" void ref_light_caches(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi,
                                 const drv_spot_light* lights, uint32_t num_lights, const uint16_t* const* flux_read,
                                 const int16_t* const* normal_read, const uint16_t* const* const* depth_levels,
                                 const uint32_t* num_depth_levels, const uint8_t* voxel_chain, uint32_t voxel_res,
                                 void* entries, uint32_t count, int sh_order, int indirect_shadow, float* const* vpl_tap,
                                 int threads) {
#define GO(v) ref_light_##v(cb, pf, vi, lights, num_lights, flux_read, normal_read, depth_levels, num_depth_levels, voxel_chain, voxel_res, entries, count, vpl_tap, nullptr, threads)
  if (sh_order == 2) { if (indirect_shadow) GO(light_sh2_s); else GO(light_sh2_n); }
  else { if (indirect_shadow) GO(light_sh1_s); else GO(light_sh1_n); }
#undef GO
}

/* shader/cacheApply.frag (Renderer::ApplyCaches, renderer.cpp:1047-1079). out_rgba = (colour, 1) or zeros where the
 * fragment is discarded. */
extern "C" void ref_apply_caches(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi, int transitions,
                                 int sh_order, const float* depth, const int16_t* normal, const uint8_t* diffuse,
                                 const uint32_t* atlas, const void* entries, uint32_t entry_count, float* out_rgba, int threads) {
#define GO(v) ref_apply_##v(cb, pf, vi, depth, normal, diffuse, atlas, entries, entry_count, out_rgba, nullptr, nullptr, threads)
  if (sh_order == 2) { if (transitions) GO(apply_sh2_t); else GO(apply_sh2_n); }
  else { if (transitions) GO(apply_sh1_t); else GO(apply_sh1_n); }
#undef GO
}

extern "C" void ref_voxel_blend(uint8_t* volume, const uint8_t* target, uint32_t res, float adaption) {
  ref_voxel_blend_voxelblend(volume, target, res, adaption);
}
extern "C" void ref_voxel_mips(uint8_t* chain, uint32_t res) { ref_voxel_mips_voxelmipmap(chain, res); }
extern "C" void ref_rsm_downsample(const uint16_t* flux_src, const int16_t* normal_src, const uint16_t* depth_src, uint32_t res,
                                   uint16_t* flux_dst, int16_t* normal_dst, uint16_t* depth_dst) {
  ref_rsm_downsample_downsample(flux_src, normal_src, depth_src, res, flux_dst, normal_dst, depth_dst);
}
extern "C" void ref_cone_trace_ao_ao(const drv_constant*, const drv_per_frame*, const drv_volume_info*, const uint8_t*, uint32_t, const float*, const int16_t*, uint32_t, uint32_t, float*, int);
extern "C" void ref_fill_rsm_fillrsm(const drv_spot_light*, const float*, const float*, const float*, const uint8_t*, uint32_t, uint16_t*, int16_t*, uint16_t*);
extern "C" void ref_tonemap_tonemap(const float*, uint32_t, float, float, float*);
/* the rows next to the path: shader/ambientocclusion.frag, fillrsm.frag, tonemapping.frag */
extern "C" void ref_cone_trace_ao(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi, const uint8_t* chain,
                                  uint32_t res, const float* depth, const int16_t* normal, uint32_t W, uint32_t H, float* out, int threads) {
  ref_cone_trace_ao_ao(cb, pf, vi, chain, res, depth, normal, W, H, out, threads);
}
extern "C" void ref_fill_rsm(const drv_spot_light* light, const float* position, const float* normal, const float* basecolor,
                             const uint8_t* coverage, uint32_t res, uint16_t* flux, int16_t* normal_out, uint16_t* depth) {
  ref_fill_rsm_fillrsm(light, position, normal, basecolor, coverage, res, flux, normal_out, depth);
}
extern "C" void ref_tonemap(const float* hdr, uint32_t n, float exposure, float drago_divider, float* out) {
  ref_tonemap_tonemap(hdr, n, exposure, drago_divider, out);
}
/* ---- INDIRECT_SPECULAR (+ DIRECT_SPECULAR_MAP_WRITE, SPECULARENVMAP_PERCACHESIZE = 16) ---- */
extern "C" void ref_specular_mips_specmip(const drv_constant*, uint32_t, uint32_t*);
extern "C" void ref_specular_fill_holes_specfill(const drv_constant*, uint32_t, uint32_t, uint32_t*);
/* variant 0: SH1 unshadowed; variant 1: SH2 + INDIRECT_SHADOW. specular_atlas: SpecularEnvmapTotalSize^2 R11G11B10F. */
extern "C" int ref_light_caches_specular(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi,
                                         const drv_spot_light* lights, uint32_t num_lights, const uint16_t* const* flux_read,
                                         const int16_t* const* normal_read, const uint16_t* const* const* depth_levels,
                                         const uint32_t* num_depth_levels, const uint8_t* voxel_chain, uint32_t voxel_res,
                                         void* entries, uint32_t count, int sh_order, int indirect_shadow,
                                         uint32_t* specular_atlas) {
#define GO(v) ref_light_##v(cb, pf, vi, lights, num_lights, flux_read, normal_read, depth_levels, num_depth_levels, voxel_chain, voxel_res, entries, count, nullptr, specular_atlas, 1)
  if (sh_order == 1 && !indirect_shadow) { GO(light_sh1_n_spec); return 0; }
  if (sh_order == 2 && indirect_shadow) { GO(light_sh2_s_spec); return 0; }
#undef GO
  return -1; /* variant not built */
}
extern "C" void ref_specular_mips(const drv_constant* cb, uint32_t cache_count, uint32_t* mips) { ref_specular_mips_specmip(cb, cache_count, mips); }
extern "C" void ref_specular_fill_holes(const drv_constant* cb, uint32_t cache_count, uint32_t max_level, uint32_t* mips) {
  ref_specular_fill_holes_specfill(cb, cache_count, max_level, mips);
}
/* variant: SH1 + transitions, or SH2 without. */
extern "C" int ref_apply_caches_specular(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi, int transitions,
                                         int sh_order, const float* depth, const int16_t* normal, const uint8_t* diffuse,
                                         const uint8_t* roughness_metallic, const uint32_t* atlas, const void* entries,
                                         uint32_t entry_count, const uint32_t* specular_mips, float* out_rgba, int threads) {
#define GO(v) ref_apply_##v(cb, pf, vi, depth, normal, diffuse, atlas, entries, entry_count, out_rgba, roughness_metallic, specular_mips, threads)
  if (sh_order == 1 && transitions) { GO(apply_sh1_t_spec); return 0; }
  if (sh_order == 2 && !transitions) { GO(apply_sh2_n_spec); return 0; }
#undef GO
  return -1;
}
extern "C" const char* ref_source(void) { return "glsl2cpp translation of /root/reference/DynamicRadianceVolume/shader"; }
