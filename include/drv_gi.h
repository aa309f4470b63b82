/*
 * drv_gi.h — C-ABI of libdrv_gi: the B200-native indirect-lighting path of
 * DynamicRadianceVolume (allocate -> VPL generation -> voxelise+mips ->
 * cache x VPL SH gather with cone-traced visibility -> per-pixel apply).
 *
 * The reference has no FFI for this path; it sits behind the private stage
 * methods of `class Renderer` (rendering/renderer.hpp:148-216), its public
 * setters (renderer.hpp:63-141) and `class Voxelization`
 * (rendering/voxelization.hpp:21-59). Every entry point below names the
 * reference method / shader it replaces.  All citations are relative to
 * /root/reference/DynamicRadianceVolume/.
 *
 * Conventions
 *  - plain C, POD structs, raw pointers + sizes; no C++/torch types.
 *  - every function returns drv_status (0 = ok, <0 = error) and never throws.
 *  - one caller thread per context; all device work is enqueued on the stream
 *    given at create time, in call order; no implicit host synchronisation
 *    except where a function says so.
 *  - the context owns every device buffer it creates (the reference's
 *    Renderer owns all GL objects, renderer.hpp:227-352); inputs passed as
 *    device pointers are BORROWED and must stay valid until the stage calls
 *    that read them have completed.
 *  - images are row-major, origin lower-left (GL), pixel centre (x+.5,y+.5).
 */
#ifndef DRV_GI_H
#define DRV_GI_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRV_MAX_CASCADES 4 /* MAX_NUM_ADDRESS_VOLUME_CASCADES, shader/globalubos.glsl:46; Renderer::s_maxNumCAVCascades renderer.hpp:146 */
#define DRV_MAX_LIGHTS 16  /* maxExpectedLights, renderer.cpp:82 */
#define DRV_LIGHTING_THREADS_PER_GROUP 64 /* shader/lightcache.glsl:95 */

typedef enum drv_status {
  DRV_OK = 0,
  DRV_ERR_INVALID = -1,   /* bad argument / configuration */
  DRV_ERR_CUDA = -2,      /* a CUDA runtime call failed (see drv_last_error) */
  DRV_ERR_CAPACITY = -3,  /* more caches requested than max_cache_count */
  DRV_ERR_NOT_BOUND = -4, /* a required input (g-buffer, RSM, uniform block) was never set */
  DRV_ERR_NO_DEVICE = -5, /* no CUDA device: there is no CPU fallback */
  DRV_ERR_PEER = -6       /* peer (NVLink) mapping failed */
} drv_status;

/* ------------------------------------------------------------------------
 * Uniform blocks — byte-exact std140 images of shader/globalubos.glsl.
 * ---------------------------------------------------------------------- */

/* `Constant`, globalubos.glsl:2-30, filled by Renderer::UpdateConstantUBO
 * (renderer.cpp:290-322). 80 bytes. */
typedef struct drv_constant {
  float ShCosLobeFactor0;         /* @0  sqrt(pi)/2 */
  float ShCosLobeFactor1;         /* @4  +sqrt(pi/3)           (renderer.cpp:297) */
  float ShCosLobeFactor2n2_p1_n1; /* @8  -sqrt(15 pi)/8  (sic, renderer.cpp:298) */
  float ShCosLobeFactor20;        /* @12 sqrt(5 pi)/16 */
  float ShCosLobeFactor2p2;       /* @16 sqrt(15 pi)/16 */
  float ShEvaFactor0;             /* @20 1/(2 sqrt(pi)) */
  float ShEvaFactor1;             /* @24 sqrt(3)/(2 sqrt(pi)) */
  float ShEvaFactor2n2_p1_n1;     /* @28 sqrt(15/(4 pi)) */
  float ShEvaFactor20;            /* @32 sqrt(5/(16 pi)) */
  float ShEvaFactor2p2;           /* @36 sqrt(15/(16 pi)) */
  int32_t BackbufferResolution[2];/* @40 */
  int32_t VoxelResolution;        /* @48 */
  int32_t AddressVolumeResolution;/* @52 */
  int32_t NumAddressVolumeCascades;/* @56 */
  uint32_t MaxNumLightCaches;     /* @60 (never read by the shaders) */
  int32_t SpecularEnvmapTotalSize;            /* @64 read only with drv_config.indirect_specular (drv_pack_specular) */
  int32_t SpecularEnvmapPerCacheSize_Texel;   /* @68 */
  float SpecularEnvmapPerCacheSize_Texcoord;  /* @72 */
  int32_t SpecularEnvmapNumCachesPerDimension;/* @76 */
} drv_constant;

/* `PerFrame`, globalubos.glsl:33-43, Renderer::UpdatePerFrameUBO
 * (renderer.cpp:324-344). Matrices are the raw row-major ei::Mat4x4 bytes;
 * GLSL `v * M` on those bytes equals ei `M * v`. 288 bytes. */
typedef struct drv_per_frame {
  float Projection[16];            /* @0 */
  float ViewProjection[16];        /* @64 */
  float InverseView[16];           /* @128 */
  float InverseViewProjection[16]; /* @192 */
  float CameraPosition[3];         /* @256 */
  float _pad0;
  float CameraDirection[3];        /* @272 */
  float PassedTime;                /* @284 */
} drv_per_frame;

/* `CAVCascade`, globalubos.glsl:48-62. 64 bytes. */
typedef struct drv_cav_cascade {
  float Min[3];
  float WorldVoxelSize;
  float Max[3];
  float _padding0;
  float DecisionMin[3];
  float _padding1;
  float DecisionMax[3];
  float _padding2;
} drv_cav_cascade;

/* `VolumeInfo`, globalubos.glsl:65-79, Renderer::UpdateVolumeUBO
 * (renderer.cpp:346-431). 288 bytes. */
typedef struct drv_volume_info {
  float VolumeWorldMin[3];   /* @0 */
  float VoxelSizeInWorld;    /* @12 */
  float VolumeWorldMax[3];   /* @16 */
  float CAVTransitionZoneSize; /* @28 */
  drv_cav_cascade AddressVolumeCascades[DRV_MAX_CASCADES]; /* @32 */
} drv_volume_info;

/* `SpotLight`, globalubos.glsl:88-112, Renderer::PrepareLights
 * (renderer.cpp:664-725). 224 bytes. */
typedef struct drv_spot_light {
  float LightIntensity[3];     /* @0 */
  float ShadowNormalOffset;    /* @12 */
  float ShadowBias;            /* @16 */
  float _pad0[3];
  float LightPosition[3];      /* @32 */
  float _pad1;
  float LightDirection[3];     /* @48 */
  float LightCosHalfAngle;     /* @60 */
  float LightViewProjection[16];        /* @64 */
  float InverseLightViewProjection[16]; /* @128 */
  int32_t RSMRenderResolution; /* @192 */
  int32_t RSMReadResolution;   /* @196 */
  float ValAreaFactor;         /* @200 */
  float IndirectShadowComputationLod;            /* @204 */
  float IndirectShadowComputationBlockSize;      /* @208 */
  int32_t IndirectShadowComputationSampleInterval; /* @212 */
  float IndirectShadowComputationSuperValWidth;  /* @216 */
  float IndirectShadowSamplingOffset;            /* @220 */
} drv_spot_light;

/* ------------------------------------------------------------------------
 * Buffers — std430 images of shader/lightcache.glsl.
 * ---------------------------------------------------------------------- */

/* LightCacheEntry with INDDIFFUSE_VIA_SH1, lightcache.glsl:33-47. 64 bytes.
 * Each vec3 is the RGB of one SH coefficient. */
typedef struct drv_cache_entry_sh1 {
  float Position[3]; float _padding0;
  float SH1neg1[3];  float SH00_r;
  float SH10[3];     float SH00_g;
  float SH1pos1[3];  float SH00_b;
} drv_cache_entry_sh1;

/* LightCacheEntry with INDDIFFUSE_VIA_SH2, lightcache.glsl:33-57. 128 bytes. */
typedef struct drv_cache_entry_sh2 {
  float Position[3]; float _padding0;
  float SH1neg1[3];  float SH00_r;
  float SH10[3];     float SH00_g;
  float SH1pos1[3];  float SH00_b;
  float SH2neg2[3];  float SH20_r;
  float SH2neg1[3];  float SH20_g;
  float SH2pos1[3];  float SH20_b;
  float SH2pos2[3];  float _padding1;
} drv_cache_entry_sh2;

/* LightCacheCounter, lightcache.glsl:83-90; doubles as the indirect-dispatch
 * argument buffer written by cachePrepareLighting.comp:8-14. 16 bytes. */
typedef struct drv_cache_counter {
  uint32_t NumCacheLightingThreadGroupsX; /* (count+63)/64 */
  uint32_t NumCacheLightingThreadGroupsY; /* 1 */
  uint32_t NumCacheLightingThreadGroupsZ; /* 1 */
  int32_t TotalLightCacheCount;
} drv_cache_counter;

/* One virtual area light = one RSM texel at the read resolution, the
 * `LightInfo` of cacheLightingRSM.comp:34-41 materialised once per light per
 * frame instead of once per 64-cache group (:137-163). Morton order. 48 bytes
 * (three 128-bit loads). */
typedef struct drv_vpl {
  float Position[3]; float DiscArea;
  float Normal[3];   float _pad0;
  float Flux[3];     float _pad1;
} drv_vpl;

/* Per indirect-shadow block (every SampleInterval VPLs): the cache-independent
 * half of cacheLightingRSM.comp:171-192. 16 bytes. */
typedef struct drv_shadow_block {
  float AverageValPos[3];
  float DistToSphereRad;
} drv_shadow_block;

/* ------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------- */

typedef struct drv_ctx drv_ctx;

/* ≙ Renderer ctor defaults (renderer.cpp:36-51,86-90) + the setters
 * SetMaxCacheCount / SetCAVCascades / SetVoxelVolumeResultion /
 * SetIndirectDiffuseMode / SetIndirectShadow / SetCAVCascadeTransitionSize /
 * OnScreenResize (renderer.hpp:63-141). Changing any of these in the
 * reference reallocates buffers / recompiles shader variants; here it means
 * drv_destroy + drv_create. */
typedef struct drv_config {
  uint32_t max_cache_count;   /* SetMaxCacheCount; entries buffer = max*128 B always (renderer.cpp:266-269) */
  uint32_t cav_cascades;      /* SetCAVCascades(numCascades, resolutionPerCascade), 1..4 */
  uint32_t cav_resolution;
  uint32_t voxel_resolution;  /* SetVoxelVolumeResultion; power of two */
  uint32_t sh_order;          /* 1 = INDDIFFUSE_VIA_SH1, 2 = INDDIFFUSE_VIA_SH2 */
  uint32_t indirect_shadow;   /* 0/1 = INDIRECT_SHADOW */
  uint32_t cascade_transitions; /* 0/1 = ADDRESSVOL_CASCADE_TRANSITIONS (m_CAVCascadeTransitionSize > 0, renderer.cpp:212) */
  uint32_t backbuffer_width;  /* OnScreenResize */
  uint32_t backbuffer_height;
  uint32_t max_lights;        /* <= DRV_MAX_LIGHTS */
  uint32_t max_rsm_resolution;/* largest RSMRenderResolution that will be bound */
  int32_t  device;            /* CUDA device ordinal */
  void*    stream;            /* cudaStream_t; NULL = context creates its own */
  uint32_t gather_variant;    /* 0 = default. Tuning switches measured in profiles/ (DESIGN.md 4.1, 4.5):
                                 bits 0..7   pair-kernel variant (0 = 35: warp-split kernel, one 8-warp CTA per SM, per-VPL
                                             scalars as 32-bit broadcast operands of the packed instructions; 31 the same
                                             with the duplicated shared-memory record of round 2's first half; 36..40 other
                                             shapes of 35; 30 the cooperative stream-K kernel; 1 TMA staging, 3/4/12 scalar or
                                             single-pair maths, 6 three CTAs/SM, 7/8 64-/32-thread CTAs, 9/10/11 VPL
                                             loop unrolled 2/8/4x, 20..28 / 32..34 other warp-split shapes)
                                 bits 8..11  apply: resident blocks/SM the kernel is compiled for (4, 6; default 5);
                                             7 = SH1 with red / green in packed FFMA2 (bit-identical, measured equal)
                                 bits 12..15 apply: rows per thread (1, 2, 8; default 4)
                                 bit 16      variant 30: two-kernel gather + finalize instead of the cooperative launch
                                 bit 17      software-pipelined cone march instead of the plain loop
                                 bit 18      per-CTA time stamps of the pair kernel (drv_debug_gather_trace)
                                 bit 19      cone pass compiled for 8 instead of 10 resident CTAs per SM (64 registers) */
  uint32_t indirect_specular; /* 0/1 = INDIRECT_SPECULAR with DIRECT_SPECULAR_MAP_WRITE (SetIndirectSpecular):
                                 per-cache hemispherical environment maps in an R11F_G11F_B10F atlas */
  uint32_t specular_per_cache_size;  /* SetPerCacheSpecularEnvMapSize: power of two, 2..16; 0 = 16 (renderer.cpp:43) */
  uint32_t specular_fill_holes_level;/* SetSpecularEnvMapHoleFillLevel: 0..log2(per cache size) (renderer.cpp:44) */
} drv_config;

drv_status drv_create(const drv_config* cfg, drv_ctx** out);
void       drv_destroy(drv_ctx* ctx);
/* Human-readable text of the last error on this context (or of the last
 * failed drv_create when ctx == NULL). */
const char* drv_last_error(const drv_ctx* ctx);
/* Version / build string: "libdrv_gi <ver> sm_100a". */
const char* drv_version(void);

/* ≙ Update*UBO / PrepareLights: the caller packs the block (include/
 * drv_math.h has the packers that restate renderer.cpp:290-431,664-725) and
 * hands over the bytes. Copied (stream-ordered) into device constant storage. */
drv_status drv_set_constant(drv_ctx* ctx, const drv_constant* block);
drv_status drv_set_per_frame(drv_ctx* ctx, const drv_per_frame* block);
drv_status drv_set_volume_info(drv_ctx* ctx, const drv_volume_info* block);
drv_status drv_set_light_count(drv_ctx* ctx, uint32_t num_lights);
drv_status drv_set_spot_light(drv_ctx* ctx, uint32_t light, const drv_spot_light* block);

/* ≙ Renderer::BindGBuffer (renderer.cpp:727-738), formats renderer.cpp:468-471:
 * depth D32F reversed-Z; normal RG16I (PackNormal16I, utils.glsl:83-89);
 * diffuse sRGB8 stored as 4 bytes/pixel (R,G,B,x). Device pointers. */
drv_status drv_bind_gbuffer(drv_ctx* ctx, const float* depth, const int16_t* normal_rg16i,
                            const uint8_t* diffuse_srgb8x, uint32_t width, uint32_t height);

/* ≙ binding a ShadowMap's three RSM textures (renderer.cpp:926-928; formats
 * renderer.cpp:1288-1291): level 0 at `resolution` = RSMRenderResolution.
 * flux RGB16F stored as 4 halfs/texel (r,g,b,x); normal RG16I; depthLinSq
 * RG16F = (dist, dist^2) (fillrsm.frag:48). Device pointers. */
/* The roughness / metallic plane of the G-buffer (RG8, renderer.cpp:469; texture unit 1): read by the apply pass
 * only with drv_config.indirect_specular. Same resolution as drv_bind_gbuffer. */
drv_status drv_bind_gbuffer_material(drv_ctx* ctx, const uint8_t* roughness_metallic_rg8);
drv_status drv_bind_rsm(drv_ctx* ctx, uint32_t light, const uint16_t* flux_rgbx16f,
                        const int16_t* normal_rg16i, const uint16_t* depthlinsq_rg16f,
                        uint32_t resolution);

/* ≙ Renderer::ShadowMap::PrepareRSM (renderer.cpp:1300-1339) running
 * downsamplersm.frag:15-33 for levels 1..log2(res)-1 into context-owned
 * storage. Must follow drv_bind_rsm each time level 0 changes. */
drv_status drv_prepare_rsm(drv_ctx* ctx, uint32_t light);

/* ≙ Voxelization::VoxelizeScene (voxelization.cpp:90-176): clear target,
 * conservative voxelisation of `num_tris` triangles (9 floats each, object
 * space, transformed by the row-major `world` matrix), blend with
 * `adaption` (= floor(dt*rate*255)/255; the call is a no-op for 0, as in
 * voxelization.cpp:100), then the mip chain. Device pointer. May be called
 * several times per frame with `clear_target` = 0 to add further entities
 * before the blend: pass DRV_VOXELIZE_* flags. */
#define DRV_VOXELIZE_CLEAR 1u  /* clear the target volume first */
#define DRV_VOXELIZE_FINISH 2u /* run blend + mips after rasterising */
drv_status drv_voxelize(drv_ctx* ctx, const float* tri_pos, uint32_t num_tris,
                        const float world[16], float adaption, uint32_t flags);

/* Installs a ready-made level-0 voxel volume (voxel_resolution^3 bytes, x fastest; host or device pointer)
 * into the persistent volume, bypassing rasterisation and the temporal blend, then rebuilds the mip chain
 * (voxelmipmap.comp) and the cone tracer's gather-ready records. For volumes produced elsewhere and for tests. */
drv_status drv_set_voxel_volume(drv_ctx* ctx, const uint8_t* level0);

/* ≙ Renderer::AllocateCaches (renderer.cpp:951-992): cacheGather.comp +
 * cachePrepareLighting.comp. Clears counter + atlas, marks, scans, compacts. */
drv_status drv_allocate_caches(drv_ctx* ctx);

/* ≙ Renderer::LightCachesRSM (renderer.cpp:899-933): for every light, VPL
 * generation + cacheLightingRSM.comp; results accumulate into the entries.
 * Scratch memory, grown on demand by the first call that needs it (that call synchronises the stream once and is
 * therefore never part of a recorded frame graph — drv_draw_frame runs one eager frame first): the partial sums of
 * the pair kernel (CTAs x 2 x coefficients x tile entries floats: 2.7 MB for SH1, 4.1 MB for SH2 on 148 SMs), and with indirect_shadow the visibility table, one float per
 * (shadow block, cache): total_blocks x min(262144, max_cache_count rounded up to 512) x 4 B, capped at 6 GiB by
 * halving the cache chunk (the frame then walks the entries in chunks). BASELINE configs[2]: 1024 x 65536 x 4 B =
 * 256 MB; configs[3]: 4096 x 262144 x 4 B = 4 GiB. */
drv_status drv_light_caches(drv_ctx* ctx);

/* ≙ Renderer::ApplyCaches (renderer.cpp:1047-1079): cacheApply.frag. */
#define DRV_HDR_RGBA16F_ADD 0u   /* reference behaviour: additive blend into RGBA16F (renderer.cpp:119,480) */
#define DRV_HDR_RGBA32F_WRITE 1u /* parity readback: overwrite float4 (rgb, 1); discarded pixels get 0 */
#define DRV_HDR_RGBA16F_WRITE 2u /* glClear(0,0,0,0) + additive blend fused: overwrite RGBA16F with (rgb, 0);
                                    discarded pixels get 0 (renderer.cpp:562 + :1053 in one pass) */
/* ≙ Renderer::PrepareSpecularEnvmaps (renderer.cpp:994-1045): mip chain of the environment-map atlas
 * (specularenvmap_mipmap.frag) and the hole-filling push-down (specularenvmap_fillholes.frag). drv_draw /
 * drv_draw_frame call it between lighting and apply when drv_config.indirect_specular is set. */
drv_status drv_prepare_specular_envmaps(drv_ctx* ctx);
drv_status drv_apply_caches(drv_ctx* ctx, void* hdr_out, uint32_t format);

/* The same for the pixel rows [y_begin, y_end) only — sort-first sharding of the apply pass over GPUs, and the
 * banded host pipeline of drv_draw_host_frame. */
drv_status drv_apply_caches_rows(drv_ctx* ctx, void* hdr_out, uint32_t format, uint32_t y_begin, uint32_t y_end);

/* The DYN_RADIANCE_VOLUME case of Renderer::Draw (renderer.cpp:539-570)
 * minus the producers: allocate -> light -> apply, one call. */
drv_status drv_draw(drv_ctx* ctx, void* hdr_out, uint32_t format);

/* One whole frame of the path, scheduled for the GPU rather than in the reference's serial GL order:
 * the light side (RSM mip chains + VPL generation, ShadowMap::PrepareRSM / cacheLightingRSM.comp:137-163) runs
 * on a second stream concurrently with the camera side (AllocateCaches), the two join before the gather, then
 * apply. Results are identical to drv_prepare_rsm (every light) + drv_draw.
 *   DRV_FRAME_PREPARE_RSM  rebuild the RSM mip chain of every bound light, down to the last level this frame
 *                          reads (drv_prepare_rsm builds the whole chain); else the mips are used as they stand
 *   DRV_FRAME_GRAPH        record the frame into a CUDA graph and replay it while nothing that feeds a kernel
 *                          argument changes (uniform blocks, bindings, shard, hdr_out, format, flags); any
 *                          drv_set_* / drv_bind_* / drv_upload_* call makes the next frame re-record.
 *                          Ignored while stage timers are enabled.
 * With drv_set_shard + imported peers the frame is the sharded one: allocate (replicated) -> peer barrier ->
 * gather of the own entry range, finished entries stored to every peer -> peer barrier -> apply. */
#define DRV_FRAME_PREPARE_RSM 1u
#define DRV_FRAME_GRAPH 2u
#define DRV_FRAME_VOXELIZE 16u      /* start the frame with VoxelizeScene (renderer.cpp:546: clear + raster + blend + mips)
                                       of the geometry bound with drv_bind_scene, on its own stream, concurrently with
                                       the light side and the allocation; joined before the gather */
#define DRV_FRAME_APPLY_OWN_ROWS 4u /* sharded runs: apply only this rank's band of rows, [rank*ceil(H/world), ...) */
#define DRV_FRAME_GATHER_IMAGE 8u   /* sharded runs (implies APPLY_OWN_ROWS; format DRV_HDR_RGBA16F_WRITE): every rank
                                       stores its band straight into rank 0's context-owned RGBA16F target over NVLink
                                       (drv_export_hdr_ipc / drv_import_peer_hdr) and a closing peer barrier makes the
                                       image complete on rank 0 — the frame contains no collective. hdr_out is ignored;
                                       the image is drv_buffers.hdr16 of rank 0. */
drv_status drv_draw_frame(drv_ctx* ctx, void* hdr_out, uint32_t format, uint32_t flags);
/* How often drv_draw_frame(DRV_FRAME_GRAPH) had to instantiate its frame graph, and how often it patched the
 * instantiated graph in place (cudaGraphExecUpdate) because only kernel arguments had changed — a moving camera
 * (drv_set_per_frame / drv_set_volume_info every frame) must show up as updates, not instantiations. */
drv_status drv_graph_stats(drv_ctx* ctx, uint64_t* instantiations, uint64_t* updates);
/* Geometry for DRV_FRAME_VOXELIZE: the arguments of drv_voxelize (device pointer to num_tris * 9 floats, borrowed
 * until replaced; world matrix and adaption factor are copied). */
drv_status drv_bind_scene(drv_ctx* ctx, const float* tri_pos, uint32_t num_tris, const float world[16], float adaption);

/* VPLs the gather actually streams, per light: drv_light_caches drops VPLs whose flux is zero in all channels
 * (RSM texels that saw no surface; they add exactly zero to every cache) and keeps the order of the rest.
 * counts[DRV_MAX_LIGHTS]; synchronises. */
drv_status drv_live_vpl_counts(drv_ctx* ctx, uint32_t* counts);

/* ------------------------------------------------------------------------
 * Rows next to the hot path, built to the same parity bar.
 * ------------------------------------------------------------------------ */
/* f1 ≙ shader/fillrsm.frag:32-61 (the FillRSM pass of Renderer::DrawShadowMaps, renderer.cpp:785-800) minus the
 * rasteriser: level 0 of light `light`'s RSM — flux RGBX16F, packed normal RG16I, depthLinSq RG16F — from the
 * per-fragment attributes of the light's view (device pointers, resolution^2 texels, row-major): world position,
 * shading normal (unnormalised), base colour as sampled (linear RGB). coverage (nullable): 0 = no fragment, the
 * texel keeps the clear value. The result is bound as the light's RSM (as drv_bind_rsm would);
 * drv_set_spot_light must have been called. */
drv_status drv_fill_rsm(drv_ctx* ctx, uint32_t light, const float* position_xyz, const float* normal_xyz,
                        const float* basecolor_rgb, const uint8_t* coverage, uint32_t resolution);

/* f2 ≙ Renderer::ConeTraceAO (renderer.cpp:936-949) + shader/ambientocclusion.frag:25-89: six voxel cones per
 * pixel through the voxel chain of the bound G-buffer; one float per pixel, discarded pixels (depth < 1e-6) are
 * left untouched. Needs a context created with indirect_shadow (the cone tracer's record chain). */
drv_status drv_cone_trace_ao(drv_ctx* ctx, float* ao_out);

/* f3 ≙ shader/tonemapping.frag:21-31 with Exposure and DragoDivider = log2(l_max + 1) (renderer.cpp:1225-1227):
 * RGBA16F HDR target -> float4 (rgb, 1). */
drv_status drv_tonemap(drv_ctx* ctx, const void* hdr_rgba16f, float exposure, float l_max, float* ldr_rgba32f);
/* ≙ WritePfm (rendering/hdrimage.cpp:6-32): host RGBA float image -> PFM file (pure host code). */
drv_status drv_write_pfm(const char* path, const float* rgba, uint32_t width, uint32_t height);
/* ≙ Renderer::SaveToPFM (renderer.cpp:1229-1235): read the RGBA16F HDR target back and write it. Synchronises. */
drv_status drv_save_to_pfm(drv_ctx* ctx, const void* hdr_rgba16f, const char* path);

/* Device pointers for parity readback / interop. Valid until drv_destroy. */
typedef struct drv_buffers {
  void*     entries;        /* LightCacheBuffer, stride entry_stride */
  uint32_t  entry_stride;   /* 64 (SH1) or 128 (SH2) */
  uint32_t  max_cache_count;
  drv_cache_counter* counter; /* LightCacheCounter */
  uint32_t* cav_atlas;      /* (nCasc*res) x res x res R32UI, x fastest (renderer.cpp:1179) */
  uint32_t  cav_width, cav_height, cav_depth;
  uint8_t*  voxel_chain;    /* R8 volume, level 0 then each mip, x fastest per level */
  uint8_t*  voxel_target;   /* R8 target volume (level 0 only) */
  uint32_t  voxel_resolution, voxel_levels;
  uint64_t  voxel_chain_bytes;
  drv_vpl*  vpls[DRV_MAX_LIGHTS];            /* Morton order, RSMReadResolution^2 each */
  drv_shadow_block* shadow_blocks[DRV_MAX_LIGHTS];
  /* RSM mip chains owned by the context (levels >= 1); level l at offset
   * rsm_level_offset(l) texels — see drv_rsm_level_offset(). */
  uint16_t* rsm_flux_mips[DRV_MAX_LIGHTS];
  int16_t*  rsm_normal_mips[DRV_MAX_LIGHTS];
  uint16_t* rsm_depth_mips[DRV_MAX_LIGHTS];
  void*     hdr16;          /* context-owned RGBA16F target (drv_draw_to_host / DRV_FRAME_GATHER_IMAGE); NULL until used */
  /* level 0 of every light's RSM as currently bound (drv_bind_rsm / drv_upload_rsm / drv_fill_rsm) */
  const uint16_t* rsm_flux0[DRV_MAX_LIGHTS];
  const int16_t*  rsm_normal0[DRV_MAX_LIGHTS];
  const uint16_t* rsm_depth0[DRV_MAX_LIGHTS];
  /* indirect specular: every level of the environment-map atlas, level 0 first, R11F_G11F_B10F texels
   * (include/drv_r11g11b10.h); level l is (specular_total_size >> l)^2; NULL unless drv_config.indirect_specular */
  uint32_t* specular_mips;
  uint32_t  specular_total_size, specular_levels;
} drv_buffers;
drv_status drv_get_buffers(drv_ctx* ctx, drv_buffers* out);
/* Texel offset of mip level `level` (>=1) inside a context-owned RSM mip
 * buffer for a level-0 resolution of `resolution`. */
uint64_t drv_rsm_level_offset(uint32_t resolution, uint32_t level);
/* Byte offset of mip `level` inside voxel_chain. */
uint64_t drv_voxel_level_offset(uint32_t resolution, uint32_t level);

/* ≙ Renderer::GetLightCacheActiveCount (renderer.cpp:1162-1165) after
 * SetReadLightCacheCount(true): synchronises the stream and reads the
 * counter. `overflow` (nullable) = caches dropped because of max_cache_count;
 * `oob_corners` (nullable) = corner cells skipped by the SURVEY B.3 policy. */
drv_status drv_active_cache_count(drv_ctx* ctx, uint32_t* count, uint32_t* overflow,
                                  uint32_t* oob_corners);

/* Synthetic inputs for the gather sweep (BASELINE config 5): install `n`
 * cache positions (float4 each, w ignored; SH zeroed) as the entry list, and
 * a ready-made VPL list for a light, bypassing allocation / VPL generation. */
drv_status drv_set_synthetic_entries(drv_ctx* ctx, const float* positions_xyzw, uint32_t n);
drv_status drv_set_vpls(drv_ctx* ctx, uint32_t light, const drv_vpl* vpls, uint32_t n);

/* Multi-GPU: this context lights only its share of the cache entries
 * (SURVEY 8e). Allocation stays replicated (deterministic => identical on
 * every rank). Entries are split into `world` contiguous ranges on 64-entry
 * boundaries of the cell-ordered list, so a shard is a run of
 * (cascade, brick) keys. rank/world = 0/1 restores single-GPU behaviour. */
drv_status drv_set_shard(drv_ctx* ctx, uint32_t rank, uint32_t world);
/* Entry range [begin,end) of `rank` for `count` active entries. Pure host maths. */
void drv_shard_range(uint32_t count, uint32_t rank, uint32_t world, uint32_t* begin, uint32_t* end);
/* Interleaved sharding (for the fused NVLink exchange, where no rank needs a contiguous range): the 64-entry groups of
 * the cell-ordered list are dealt round-robin, group g to rank g % world, so that regions whose cone marches are long
 * (or short) are spread over all GPUs instead of landing on one. drv_shard_entry maps the local index of a rank's
 * entry to the entry. */
drv_status drv_set_shard_interleave(drv_ctx* ctx, uint32_t enable);
void drv_shard_entry(uint32_t local, uint32_t rank, uint32_t world, uint32_t* entry);
/* Peer exchange over NVLink: every rank exports an IPC handle of its entries
 * buffer, imports the others', and drv_light_caches then stores each finished
 * entry to all peers from inside the gather epilogue (fused all-gather). */
#define DRV_IPC_HANDLE_BYTES 64
drv_status drv_export_entries_ipc(drv_ctx* ctx, uint8_t handle[DRV_IPC_HANDLE_BYTES]);
drv_status drv_import_peer_entries(drv_ctx* ctx, uint32_t peer_rank, const uint8_t handle[DRV_IPC_HANDLE_BYTES]);

/* Fused image gather: rank 0 exports its context-owned RGBA16F target, every other rank maps it. */
drv_status drv_export_hdr_ipc(drv_ctx* ctx, uint8_t handle[DRV_IPC_HANDLE_BYTES]);
drv_status drv_import_peer_hdr(drv_ctx* ctx, uint32_t peer_rank, const uint8_t handle[DRV_IPC_HANDLE_BYTES]);

/* Stream-ordered barrier across the ranks of a sharded run, through flags in NVLink peer memory (no host
 * round trip, no NCCL call): work enqueued after it on this context's stream starts only when every rank's
 * work enqueued before its own drv_peer_barrier has finished. Every rank must call it the same number of
 * times. Requires drv_set_shard + drv_import_peer_entries for all peers (all contexts must share
 * max_cache_count). A no-op for world == 1. */
drv_status drv_peer_barrier(drv_ctx* ctx);
/* Health of the cross-GPU barriers: a barrier gives up after ~4 s (a peer died or left the frame early), later
 * barriers of this context then no longer wait, and this call — like drv_active_cache_count on a sharded context —
 * returns DRV_ERR_PEER with the epoch of the barrier that gave up and the rank it was waiting for. */
drv_status drv_peer_status(drv_ctx* ctx, uint32_t* timed_out_epoch, uint32_t* missing_rank);
/* Re-arms the barriers after an error: zeroes this context's epoch counter, flags and time-out marker. Call it on
 * EVERY rank, with a host-side barrier (torch.distributed / MPI) before and after, while no frame is in flight. */
drv_status drv_peer_reset(drv_ctx* ctx);

/* ≙ FrameProfiler (frameprofiler.hpp:138-148): CUDA-event stage timers with
 * the reference's scope names. Enabled timers record events around each
 * stage; drv_stage_ms synchronises on the stage's end event. */
typedef enum drv_stage {
  DRV_STAGE_VOXELIZE_SCENE = 0,   /* "VoxelizeScene"   voxelization.cpp:103 */
  DRV_STAGE_VOXEL_BLEND_MIPMAP,   /* "VoxelBlendMipMap" voxelization.cpp:145 */
  DRV_STAGE_ALLOCATE_CACHES,      /* "AllocateCaches"  renderer.cpp:953 */
  DRV_STAGE_LIGHT_CACHES,         /* "LightCaches"     renderer.cpp:901 */
  DRV_STAGE_APPLY_CACHES,         /* "ApplyCaches"     renderer.cpp:1049 */
  DRV_STAGE_PREPARE_RSM,          /* RSM mip chain (ShadowMap::PrepareRSM) */
  DRV_STAGE_GATHER_KERNEL,        /* the cache x VPL kernel(s) alone, inside LightCaches (cone pass + pair pass when shadowed) */
  DRV_STAGE_CONE_KERNEL,          /* the cone pass (visibility table) alone, inside GatherKernel */
  DRV_STAGE_COUNT
} drv_stage;
drv_status drv_enable_stage_timers(drv_ctx* ctx, int enable);
drv_status drv_stage_ms(drv_ctx* ctx, drv_stage stage, float* ms);
const char* drv_stage_name(drv_stage stage);
/* Number of kernels this context has launched since creation. */
uint64_t drv_kernel_launches(const drv_ctx* ctx);

/* Host-buffer convenience for the end-to-end measurement: stream-ordered
 * H2D copies of a frame's inputs into context-owned device images, and D2H of
 * the HDR result. Host pointers should be pinned. */
drv_status drv_upload_gbuffer(drv_ctx* ctx, const float* depth, const int16_t* normal_rg16i,
                              const uint8_t* diffuse_srgb8x, uint32_t width, uint32_t height);
drv_status drv_upload_rsm(drv_ctx* ctx, uint32_t light, const uint16_t* flux_rgbx16f,
                          const int16_t* normal_rg16i, const uint16_t* depthlinsq_rg16f,
                          uint32_t resolution);
/* Runs drv_draw into the context-owned RGBA16F target (cleared first) and
 * copies it to `hdr_host` (width*height*8 bytes). Synchronises. */
drv_status drv_draw_to_host(drv_ctx* ctx, void* hdr_host);

/* The same end-to-end frame, pipelined: RSM level 0 and depth are copied first, the mip chain / allocation /
 * lighting run while the normal + albedo images stream in band by band, every band is applied as soon as it has
 * arrived and its RGBA16F rows start their way back while the next band is still being applied — H2D, compute
 * and D2H overlap inside ONE frame (no cross-frame pipelining). Host pointers must be pinned for the overlap to
 * happen. Uniform blocks must have been set; RSM mips are rebuilt; with indirect shadows the voxel volume is
 * used as it stands (call drv_voxelize before). Synchronises. */
typedef struct drv_host_frame {
  const float* depth;            /* W*H float32 */
  const int16_t* normal_rg16i;   /* W*H*2 int16 */
  const uint8_t* diffuse_srgb8x; /* W*H*4 bytes */
  uint32_t num_lights;
  const uint16_t* rsm_flux_rgbx16f[DRV_MAX_LIGHTS];
  const int16_t* rsm_normal_rg16i[DRV_MAX_LIGHTS];
  const uint16_t* rsm_depthlinsq_rg16f[DRV_MAX_LIGHTS];
  uint32_t rsm_resolution[DRV_MAX_LIGHTS];
  void* hdr_out;                 /* W*H*8 bytes RGBA16F: the cleared target plus the indirect light (clear fused into the apply pass) */
  uint32_t bands;                /* equal bands of rows, <= 32; 0 = default (4) */
} drv_host_frame;
drv_status drv_draw_host_frame(drv_ctx* ctx, const drv_host_frame* frame);

/* ------------------------------------------------------------------------
 * Host-side packers (pure CPU, usable without a device): C entry points of
 * the C++ packers in include/drv_math.h, which restate
 * Renderer::UpdateConstantUBO / UpdatePerFrameUBO / UpdateVolumeUBO /
 * PrepareLights (renderer.cpp:290-431, 664-725) and the ei maths they use.
 * ---------------------------------------------------------------------- */
typedef struct drv_camera_desc { /* camera/camera.hpp:15-39 */
  float position[3];
  float direction[3];
  float up[3];
  float hfov_degrees;
  float aspect_ratio;
  float near_plane;
  float far_plane;
} drv_camera_desc;

typedef struct drv_light_desc { /* scene/light.hpp:8-55, scene/scene.cpp:6-7 */
  float intensity[3];
  float position[3];
  float direction[3];
  float half_angle;
  uint32_t rsm_resolution;
  uint32_t rsm_read_lod;
  float normal_offset_shadow_bias;
  float shadow_bias;
  uint32_t indirect_shadow_lod;
  float near_plane;
  float far_plane;
} drv_light_desc;

void drv_pack_constant(drv_constant* out, int32_t width, int32_t height, int32_t voxel_resolution,
                       int32_t cav_resolution, int32_t cav_cascades, uint32_t max_caches);
/* The specular fields of an already packed Constant block (renderer.cpp:253, 316-319). */
void drv_pack_specular(drv_constant* inout, uint32_t max_caches, uint32_t per_cache_size);
void drv_pack_per_frame(drv_per_frame* out, const drv_camera_desc* camera, float passed_time);
void drv_pack_volume_info(drv_volume_info* out, const drv_camera_desc* camera, const float scene_min[3],
                          const float scene_max[3], int32_t voxel_resolution, int32_t cav_resolution,
                          int32_t cav_cascades, const float* cascade_world_size, float transition_zone_size);
void drv_pack_spot_light(drv_spot_light* out, const drv_light_desc* light);

/* Micro-benchmarks that give the FP32 / MUFU / shared-memory / L2 roofline
 * denominators on the device in use (SURVEY 8d asks for a measured FP32
 * peak). `which`: see drv_microbench_name. Result in the unit it names. */
drv_status drv_microbench(int32_t device, uint32_t which, double* result);
const char* drv_microbench_name(uint32_t which);
uint32_t    drv_microbench_count(void);

/* Diagnostics (drv_config.gather_variant bit 18): %globaltimer stamps in ns of the last cache x VPL kernel launch,
 * at 4 points of every CTA (entry, prologue done / pair loop done, exit), then %smid and clock64 at the last
 * three: 8 words per CTA — how the fixed costs and the balance of that kernel are measured
 * (tools/gather_trace.py). `out` holds 8 * capacity_ctas words. */
drv_status drv_debug_gather_trace(drv_ctx* ctx, uint64_t* out, uint32_t capacity_ctas, uint32_t* num_ctas);
/* Diagnostics: voxel samples (cone steps) the cone pass has taken since the last call (and resets the count) — the
 * unit of that kernel's roofline record in bench.py. */
drv_status drv_debug_cone_steps(drv_ctx* ctx, uint64_t* steps);
/* Diagnostics: timeline of the last drv_draw_host_frame, in ms after its first copy was queued (waits for the
 * frame): out[0] = RSMs on the device, out[1] = depth on the device, out[2] = caches lit, then for every band b
 * out[3+3b] = its normals / albedo on the device, out[4+3b] = band applied, out[5+3b] = band back on the host.
 * capacity >= 3 + 3 * bands (bands <= 32). The time stamps are taken only if drv_enable_stage_timers(ctx, 1) was
 * called before the context's first drv_draw_host_frame. */
drv_status drv_debug_host_frame_timeline(drv_ctx* ctx, float* out, uint32_t capacity, uint32_t* bands);

#ifdef __cplusplus
}
#endif
#endif /* DRV_GI_H */
