/*
 * scenes.h — procedural synthetic inputs for tests and bench:
 * box-built scenes ray-cast analytically into a G-buffer and reflective
 * shadow maps in the REFERENCE ENCODINGS (depth D32F reversed-Z, normals
 * RG16I via PackNormal16I, diffuse sRGB8, RSM flux/depth IEEE half), the
 * matching triangle list for the voxeliser, and the seeded entry / VPL
 * generators of the gather sweep. Deterministic; no files, no network.
 *
 * Independent of oracle/ (it only produces inputs), C API for ctypes.
 * Citations are relative to /root/reference/DynamicRadianceVolume/.
 */
#ifndef DRV_SCENES_H
#define DRV_SCENES_H

#include "../include/drv_gi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct scn_scene scn_scene;

/* "cornell": 5-wall box [-2.5,2.5]x[0,5]x[-2.5,2.5] open towards +z with two
 *            blocks (BASELINE config 1).
 * "atrium":  closed Sponza-like hall 10 x 7 x 12 m with two column rows,
 *            lintels, gallery slabs and ceiling beams (configs 2-4).
 * `scale` multiplies all coordinates (config 4 uses a larger hall). */
scn_scene* scn_create(const char* name, float scale);
void scn_destroy(scn_scene* s);
uint32_t scn_num_boxes(const scn_scene* s);
/* Scene::GetBoundingBox equivalent (scene/scene.hpp:28-41). */
void scn_bounding_box(const scn_scene* s, float bmin[3], float bmax[3]);
/* 12 triangles per box, 9 floats each, world space. Returns triangle count;
 * `out` may be NULL to query. */
uint32_t scn_triangles(const scn_scene* s, float* out, uint32_t max_tris);

/* G-buffer as DrawSceneToGBuffer would leave it (fillgbuffer.frag:22-42,
 * formats renderer.cpp:468-471): one primary ray per pixel centre. Pixels
 * that hit nothing keep depth 0 (the clear value, renderer.cpp:113-117). */
void scn_render_gbuffer(const scn_scene* s, const drv_per_frame* pf, uint32_t width, uint32_t height,
                        float* depth, int16_t* normal_rg16i, uint8_t* diffuse_srgb8x, int threads);

/* RSM level 0 as DrawShadowMaps + fillrsm.frag:32-61 would leave it at
 * `light->RSMRenderResolution`: flux = albedo * I * spotFalloff *
 * pixelSteradian / pi (4 halfs, x = 0), normal RG16I, depthLinSq =
 * (dist, dist^2) halfs. Texels that hit nothing get zeros. */
void scn_render_rsm(const scn_scene* s, const drv_spot_light* light, uint16_t* flux_rgbx16f,
                    int16_t* normal_rg16i, uint16_t* depthlinsq_rg16f, int threads);

/* Gather sweep inputs (BASELINE config 5 / SURVEY 8d C5). RNG: Wang hash of
 * (seed + element index) then 32-bit xorshift per draw, as
 * utilities/random.cpp:5-22. */
void scn_sweep_entries(uint32_t seed, uint32_t n, float* positions_xyzw);
void scn_sweep_vpls(uint32_t seed, uint32_t n, float val_area_factor, drv_vpl* out);

#ifdef __cplusplus
}
#endif
#endif
