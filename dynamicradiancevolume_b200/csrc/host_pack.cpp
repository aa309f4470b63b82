// host_pack.cpp — C entry points of the packers in include/drv_math.h
// (≙ Renderer::UpdateConstantUBO / UpdatePerFrameUBO / UpdateVolumeUBO /
// PrepareLights, rendering/renderer.cpp:290-431, 664-725). Pure host code.
#include "../../include/drv_math.h"

extern "C" void drv_pack_constant(drv_constant* out, int32_t width, int32_t height, int32_t voxel_resolution,
                                  int32_t cav_resolution, int32_t cav_cascades, uint32_t max_caches) {
  drv::packConstant(out, width, height, voxel_resolution, cav_resolution, cav_cascades, max_caches);
}

static drv::Camera to_camera(const drv_camera_desc* c) {
  drv::Camera cam;
  cam.position = drv::Vec3(c->position);
  cam.direction = drv::Vec3(c->direction);
  cam.up = drv::Vec3(c->up);
  cam.hfovDegrees = c->hfov_degrees;
  cam.aspectRatio = c->aspect_ratio;
  cam.nearPlane = c->near_plane;
  cam.farPlane = c->far_plane;
  return cam;
}

extern "C" void drv_pack_specular(drv_constant* inout, uint32_t max_caches, uint32_t per_cache_size) {
  drv::packSpecular(inout, max_caches, per_cache_size);
}

extern "C" void drv_pack_per_frame(drv_per_frame* out, const drv_camera_desc* camera, float passed_time) {
  drv::packPerFrame(out, to_camera(camera), passed_time);
}

extern "C" void drv_pack_volume_info(drv_volume_info* out, const drv_camera_desc* camera, const float scene_min[3],
                                     const float scene_max[3], int32_t voxel_resolution, int32_t cav_resolution,
                                     int32_t cav_cascades, const float* cascade_world_size, float transition_zone_size) {
  drv::packVolumeInfo(out, to_camera(camera), drv::Vec3(scene_min), drv::Vec3(scene_max), voxel_resolution,
                      cav_resolution, cav_cascades, cascade_world_size, transition_zone_size);
}

extern "C" void drv_pack_spot_light(drv_spot_light* out, const drv_light_desc* l) {
  drv::Light light;
  light.intensity = drv::Vec3(l->intensity);
  light.position = drv::Vec3(l->position);
  light.direction = drv::Vec3(l->direction);
  light.halfAngle = l->half_angle;
  light.rsmResolution = l->rsm_resolution;
  light.rsmReadLod = l->rsm_read_lod;
  light.normalOffsetShadowBias = l->normal_offset_shadow_bias;
  light.shadowBias = l->shadow_bias;
  light.indirectShadowComputationLod = l->indirect_shadow_lod;
  light.nearPlane = l->near_plane;
  light.farPlane = l->far_plane;
  drv::packSpotLight(out, light);
}
