/*
 * oracle.h — C API of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A scalar C++ restatement of the GLSL programs on DynamicRadianceVolume's
 * indirect-lighting path. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; libdrv_gi never
 * links or calls it.
 *
 * PARITY PINNED TO THE REFERENCE'S SOURCE TEXT: the reference ships no golden
 * vectors and its host is Win32/OpenGL, but its shaders are C-like text. A
 * mechanical GLSL -> C++ rewrite (oracle/ref/glsl2cpp.py + glsl_compat.h: work
 * groups as fibers, shared memory, barriers, atomics, software samplers)
 * compiles the UNMODIFIED shader files into oracle/_ref/libdrv_ref.so, and
 * tests/test_oracle_vs_ref.py requires this restatement to equal it bit for
 * bit (allocation set, entry positions, VPL list, SH coefficients with and
 * without cone-traced shadows, applied image, voxel blend / mips, RSM mips).
 * Not covered by that pin: the voxeliser (fixed-function rasterisation in the
 * reference; the oracle defines the covered set) and the host-side
 * uniform packers (float64 restatements in tests/test_host_math.py).
 *
 * Arithmetic policy (what GLSL leaves open is fixed here, DESIGN.md "Parity
 * policy"): IEEE-754 binary32, round-to-nearest-even, every * and + rounded
 * separately (built with -ffp-contract=off), dot products summed left to
 * right, normalize(v) = v * (1/sqrt(dot(v,v))), mix(a,b,t) = a*(1-t) + b*t,
 * float->int conversion truncates and saturates (NaN -> 0).
 *
 * Citations are relative to /root/reference/DynamicRadianceVolume/.
 */
#ifndef DRV_ORACLE_H
#define DRV_ORACLE_H

#include "../include/drv_gi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Number of worker threads used when `threads` <= 0 (hardware_concurrency). */
int orc_default_threads(void);

/* shader/cacheGather.comp:93-164 + cachePrepareLighting.comp:8-14.
 * atlas: (nCasc*res) x res x res u32, x fastest. entries: `max_caches` slots
 * of `entry_stride` bytes; Position written, SH zeroed. Indices are assigned
 * in ascending linear-cell-id order (one of the orders the reference's atomic
 * counter can produce). Returns the number of caches. */
int orc_allocate_caches(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi,
                        int transitions, const float* depth, uint32_t* atlas, void* entries,
                        uint32_t entry_stride, uint32_t max_caches, drv_cache_counter* counter,
                        uint32_t* overflow, uint32_t* oob_corners, int threads);

/* The same trigger logic, returning only the sorted linear cell ids that get
 * a cache (the parity object of SURVEY D.1). `ids` may be NULL to count. */
int orc_allocated_cell_ids(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi,
                           int transitions, const float* depth, int32_t* ids, uint32_t max_ids,
                           int threads);

/* shader/downsamplersm.frag:15-33: one mip step. src is res x res, dst is
 * (res/2) x (res/2). flux 4 halfs/texel, normal 2 int16, depth 2 halfs. */
void orc_rsm_downsample(const uint16_t* flux_src, const int16_t* normal_src, const uint16_t* depth_src,
                        uint32_t res, uint16_t* flux_dst, int16_t* normal_dst, uint16_t* depth_dst);

/* cacheLightingRSM.comp:137-163: the VPL list of one light in Morton order
 * from the three RSM images at the READ level (RSMReadResolution^2 texels). */
void orc_generate_vpls(const drv_spot_light* light, const uint16_t* flux_rgbx16f,
                       const int16_t* normal_rg16i, const uint16_t* depthlinsq_rg16f, drv_vpl* out);

/* cacheLightingRSM.comp:168-192: the cache-independent half of the indirect
 * shadow sample for every block. depth_lod is the depthLinSq image at mip
 * IndirectShadowComputationLod relative to the read level
 * (resolution RSMReadResolution >> lod). Writes RSMReadResolution^2 /
 * SampleInterval records. */
void orc_shadow_blocks(const drv_spot_light* light, const uint16_t* depth_lod, drv_shadow_block* out);

/* cacheLightingRSM.comp:195-230 for one cache position and one block. */
float orc_cone_trace(const drv_volume_info* vi, const uint8_t* voxel_chain, uint32_t voxel_res,
                     const float cache_pos[3], const drv_shadow_block* block);

/* D.0 sampler: trilinear + mip-linear, clamp to edge; p in [0,1]^3. */
float orc_sample_voxel(const uint8_t* voxel_chain, uint32_t voxel_res, const float p[3], float lod);

/* cacheLightingRSM.comp:83-374 for entries [first, first+count): for each
 * light in order, zeroed accumulators, the Morton-ordered VPL loop with the
 * optional cone-traced shadowing, then `entry.SH += acc`.
 * accumulate_fp64 != 0 keeps the accumulators in double (error attribution
 * only; the parity oracle is the float one). */
void orc_light_caches(const drv_constant* cb, const drv_volume_info* vi, const drv_spot_light* lights,
                      uint32_t num_lights, const drv_vpl* const* vpls,
                      const drv_shadow_block* const* blocks, const uint8_t* voxel_chain,
                      void* entries, uint32_t entry_stride, uint32_t first, uint32_t count,
                      int sh_order, int indirect_shadow, int accumulate_fp64, int threads);

/* cacheApply.frag:120-195 + lightcache.glsl:137-183. out_rgba: width*height
 * float4 = (rgb before the additive blend, 1) or zeros for discarded pixels. */
void orc_apply_caches(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi,
                      int transitions, int sh_order, const float* depth, const int16_t* normal_rg16i,
                      const uint8_t* diffuse_srgb8x, const uint32_t* atlas, const void* entries,
                      uint32_t entry_stride, uint32_t max_caches, float* out_rgba, int threads);

/* voxelize.vert:15-23, voxelize.geom:19-112, voxelize.frag:21-58 as a
 * closed-form coverage test. Sets target voxels to 255; does
 * not clear. `world` row-major. */
void orc_voxelize(const drv_volume_info* vi, uint32_t res, const float* tri_pos, uint32_t num_tris,
                  const float world[16], uint8_t* target);
/* voxelblend.comp:8-19 on UNORM8 (adaption = k/255). */
void orc_voxel_blend(uint8_t* volume, const uint8_t* target, uint32_t res, float adaption);
/* voxelmipmap.comp:8-13 + voxelization.cpp:155-172: fills levels 1.. of the
 * chain from level 0. */
void orc_voxel_mips(uint8_t* chain, uint32_t res);

/* ---- rows SURVEY.md 8(f) ranks next to the hot path (adjacent.cpp) ---- */
/* f1, shader/fillrsm.frag:32-61: RSM level 0 (flux RGBX16F, packed normal RG16I, depthLinSq RG16F) from the
 * rasteriser's per-fragment attributes. coverage NULL = every texel covered. */
void orc_fill_rsm(const drv_spot_light* light, const float* position_xyz, const float* normal_xyz,
                  const float* basecolor_rgb, const uint8_t* coverage, uint32_t res, uint16_t* flux_rgbx16f,
                  int16_t* normal_rg16i, uint16_t* depthlinsq_rg16f);
/* f2, shader/ambientocclusion.frag:25-89: one float per pixel, discarded pixels untouched. */
void orc_cone_trace_ao(const drv_per_frame* pf, const drv_volume_info* vi, const uint8_t* voxel_chain,
                       uint32_t voxel_res, const float* depth, const int16_t* normal_rg16i, uint32_t width,
                       uint32_t height, float* out, int threads);
/* f3, shader/tonemapping.frag:21-31. */
void orc_tonemap(const float* hdr_rgba, uint32_t n, float exposure, float drago_divider, float* out_rgb);

/* f4, INDIRECT_SPECULAR + DIRECT_SPECULAR_MAP_WRITE (specular.cpp): cacheLightingRSM.comp with the per-cache
 * hemispherical environment maps (SH + R11F_G11F_B10F atlas of SpecularEnvmapTotalSize^2 texels, cleared first),
 * Renderer::PrepareSpecularEnvmaps (mip chain, hole filling; `mips` = every level, level 0 first), and
 * cacheApply.frag with the specular term. */
void orc_light_caches_specular(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi,
                               const drv_spot_light* lights, uint32_t num_lights, const drv_vpl* const* vpls,
                               const drv_shadow_block* const* blocks, const uint8_t* voxel_chain, void* entries,
                               uint32_t entry_stride, uint32_t count, int sh_order, int indirect_shadow, uint32_t* atlas);
void orc_specular_mips(const drv_constant* cb, uint32_t count, uint32_t* mips);
void orc_specular_fill_holes(const drv_constant* cb, uint32_t count, uint32_t max_level, uint32_t* mips);
void orc_apply_caches_specular(const drv_constant* cb, const drv_per_frame* pf, const drv_volume_info* vi, int transitions,
                               int sh_order, const float* depth, const int16_t* normal_rg16i, const uint8_t* diffuse_srgb8x,
                               const uint8_t* roughness_metallic_rg8, const uint32_t* atlas, const void* entries,
                               uint32_t entry_stride, uint32_t max_caches, const uint32_t* specular_mips, float* out_rgba,
                               int threads);

/* Exact helpers shared with tests. */
float    orc_half_to_float(uint16_t h);
uint16_t orc_float_to_half(float f);
void     orc_pack_normal16i(const float n[3], int16_t out[2]);   /* utils.glsl:62-89 + clamp */
void     orc_unpack_normal16i(const int16_t in[2], float n[3]);  /* utils.glsl:44-58 */
float    orc_srgb8_to_linear(uint8_t v);
uint32_t orc_morton_decode_x(uint32_t k);                        /* cacheLightingRSM.comp:46-62 */
uint32_t orc_morton_decode_y(uint32_t k);

#ifdef __cplusplus
}
#endif
#endif
