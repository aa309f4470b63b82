// drv_renderer.hpp — C++ host side above the C-ABI: `drv::Renderer`, a mirror of the reference's
// `class Renderer` (rendering/renderer.hpp:36-216) for the indirect-lighting path.
//
// Same method names, argument meaning, defaults and call order as the reference, so code (and tests) written
// against `Renderer` read the same here:
//   setters / getters                     renderer.hpp:63-141   (SetMaxCacheCount, SetCAVCascades, SetIndirectDiffuseMode ...)
//   UpdateConstantUBO / PerFrame / Volume renderer.cpp:290-431  -> drv_pack_* + drv_set_*
//   PrepareLights                         renderer.cpp:664-725  -> drv_pack_spot_light + drv_set_spot_light
//   VoxelizeScene                         voxelization.cpp:90-176 (adaption = floor(dt*rate*255)/255, remainder carried)
//   AllocateCaches / LightCachesRSM / PrepareSpecularEnvmaps / ApplyCaches   renderer.cpp:951-992 / 899-933 / 994-1045 / 1047-1079
//   Draw                                  renderer.cpp:501-645  (modes DYN_RADIANCE_VOLUME and AMBIENTOCCLUSION)
//   ConeTraceAO, SaveToPFM, SetExposure / SetTonemapLMax         renderer.cpp:936-949, 1229-1235, 1218-1227
// What the reference rasterises itself (G-buffer, RSM level 0) is handed in as device pointers: BindGBuffer,
// BindShadowMap. Everything below the method bodies is include/drv_gi.h; header-only, C++17. The only CUDA runtime
// calls are the allocation / clear of the HDR target and the stream the mirror owns — no kernels, no CPU fallback.
//
// Errors: the reference's methods are void and log (GL errors only in _DEBUG). The mirror keeps the signatures and
// records the first failing drv_status since the last Draw began (GetLastStatus / GetLastError); later calls still
// go to the library, which validates every call on its own.
#pragma once

#include <cuda_runtime_api.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "drv_gi.h"
#include "drv_math.h"

namespace drv {

// Camera (camera/camera.hpp:15-39) and Light (scene/light.hpp:8-55) are the structs of include/drv_math.h, which also
// holds the packers as header code; the mirror goes through their C entry points (drv_pack_*) like any other host.
inline drv_camera_desc ToDesc(const Camera& c) {
  drv_camera_desc d;
  store3(d.position, c.position);
  store3(d.direction, c.direction);
  store3(d.up, c.up);
  d.hfov_degrees = c.hfovDegrees;
  d.aspect_ratio = c.aspectRatio;
  d.near_plane = c.nearPlane;
  d.far_plane = c.farPlane;
  return d;
}
inline drv_light_desc ToDesc(const Light& l) {
  drv_light_desc d;
  store3(d.intensity, l.intensity);
  store3(d.position, l.position);
  store3(d.direction, l.direction);
  d.half_angle = l.halfAngle;
  d.rsm_resolution = l.rsmResolution;
  d.rsm_read_lod = l.rsmReadLod;
  d.normal_offset_shadow_bias = l.normalOffsetShadowBias;
  d.shadow_bias = l.shadowBias;
  d.indirect_shadow_lod = l.indirectShadowComputationLod;
  d.near_plane = l.nearPlane;
  d.far_plane = l.farPlane;
  return d;
}

// The slice of SceneEntity / Model (scene/sceneentity.hpp, scene/model.hpp) the voxeliser consumes.
struct SceneEntity {
  const float* devicePositions = nullptr;  // numTriangles * 9 floats, object space, device memory
  uint32_t numTriangles = 0;
  float world[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // ComputeWorldMatrix(), row-major
};

// The slice of class Scene (scene/scene.hpp:28-41) the path consumes.
class Scene {
 public:
  std::vector<Light>& GetLights() { return m_lights; }
  const std::vector<Light>& GetLights() const { return m_lights; }
  std::vector<SceneEntity>& GetEntities() { return m_entities; }
  const std::vector<SceneEntity>& GetEntities() const { return m_entities; }
  void SetBoundingBox(const float mn[3], const float mx[3]) {
    std::memcpy(m_bboxMin, mn, 12);
    std::memcpy(m_bboxMax, mx, 12);
  }
  const float* GetBoundingBoxMin() const { return m_bboxMin; }
  const float* GetBoundingBoxMax() const { return m_bboxMax; }

 private:
  std::vector<Light> m_lights;
  std::vector<SceneEntity> m_entities;
  float m_bboxMin[3] = {0, 0, 0}, m_bboxMax[3] = {1, 1, 1};
};

class Renderer {
 public:
  enum class Mode {  // renderer.hpp:49-60; the modes this library implements are marked
    RSM_BRUTEFORCE = 0,
    DYN_RADIANCE_VOLUME = 1,  // implemented
    DYN_RADIANCE_VOLUME_DEBUG = 2,
    GBUFFER_DEBUG = 3,
    DIRECTONLY,
    VOXELVIS = 5,
    AMBIENTOCCLUSION  // implemented
  };
  enum class IndirectDiffuseMode { SH1, SH2 };
  static const unsigned int s_maxNumCAVCascades = DRV_MAX_CASCADES;  // renderer.hpp:146

  // `stream`: the cudaStream_t every stage is enqueued on (the reference issues everything on the GL context's
  // queue); nullptr = the mirror creates and owns one.
  Renderer(const std::shared_ptr<const Scene>& scene, unsigned int width, unsigned int height, int device = 0,
           void* stream = nullptr)
      : m_scene(scene), m_width(width), m_height(height), m_device(device), m_stream(stream) {
    // constructor order of renderer.cpp:84-90: voxelisation at 128^3, 16384 caches, 3 cascades of 32^3
    SetMaxCacheCount(16384);
    SetCAVCascades(3, 32);
  }
  ~Renderer() {
    ReleaseContext();
    if (m_hdr) cudaFree(m_hdr);
    if (m_ownStream) cudaStreamDestroy(static_cast<cudaStream_t>(m_stream));
  }
  Renderer(const Renderer&) = delete;
  Renderer& operator=(const Renderer&) = delete;

  // ---- Draw (renderer.cpp:501-645) minus rasterisation, direct lighting and the tonemap to the back buffer
  void Draw(const Camera& camera, bool detachViewFromCameraUpdate, float timeSinceLastFrame) {
    m_status = DRV_OK;
    m_passedTime += timeSinceLastFrame;
    UpdatePerFrameUBO(camera);
    if (!detachViewFromCameraUpdate) UpdateVolumeUBO(camera);
    PrepareLights();
    if (m_overlappedFrame && !detachViewFromCameraUpdate && m_scene->GetEntities().size() <= 1 &&
        (m_mode == Mode::DYN_RADIANCE_VOLUME || m_mode == Mode::DYN_RADIANCE_VOLUME_DEBUG)) {
      DrawOverlapped(timeSinceLastFrame);
      return;
    }
    switch (m_mode) {
      case Mode::DYN_RADIANCE_VOLUME_DEBUG:
      case Mode::DYN_RADIANCE_VOLUME:
        if (m_indirectShadow) VoxelizeScene(timeSinceLastFrame);
        if (!detachViewFromCameraUpdate) {
          AllocateCaches();
          LightCachesRSM();
          if (m_indirectSpecular) PrepareSpecularEnvmaps();
        }
        ClearHDRBackbuffer();  // glClear(GL_COLOR_BUFFER_BIT), renderer.cpp:562
        ApplyCaches();
        break;
      case Mode::AMBIENTOCCLUSION:  // renderer.cpp:631-642
        VoxelizeScene(timeSinceLastFrame);
        ConeTraceAO();
        break;
      default:
        Fail(DRV_ERR_INVALID, "Renderer::Draw: this mode is outside the indirect-lighting path");
        break;
    }
  }

  // Not in the reference: the same frame issued as ONE drv_draw_frame — voxelisation, RSM mips + VPLs and the
  // allocation on three streams, joined before the gather, the clear fused into the apply pass, recorded into a CUDA
  // graph that a moving camera only patches. Bit-identical to the serial order above; this is the path bench.py times.
  // Takes effect for scenes with at most one entity (drv_bind_scene holds one triangle list).
  void SetOverlappedFrame(bool enable) { m_overlappedFrame = enable; }
  bool GetOverlappedFrame() const { return m_overlappedFrame; }

  void SaveToPFM(const std::string& filename) {  // renderer.cpp:1229-1235
    if (!Context() || !m_hdr) return;
    Check(drv_save_to_pfm(m_ctx, m_hdr, filename.c_str()));
  }

  void SetMode(Mode mode) { m_mode = mode; }
  Mode GetMode() const { return m_mode; }

  // ---- settings that select a shader variant or reallocate in the reference: the context is rebuilt lazily
  IndirectDiffuseMode GetIndirectDiffuseMode() const { return m_indirectDiffuseMode; }
  void SetIndirectDiffuseMode(IndirectDiffuseMode mode) { m_indirectDiffuseMode = mode; ReleaseContext(); }
  void SetIndirectShadow(bool active) { m_indirectShadow = active; ReleaseContext(); }
  bool GetIndirectShadow() const { return m_indirectShadow; }
  void SetIndirectSpecular(bool active) { m_indirectSpecular = active; ReleaseContext(); }
  bool GetIndirectSpecular() const { return m_indirectSpecular; }

  void SetVoxelVolumeResultion(unsigned int resolution) { m_voxelResolution = resolution; ReleaseContext(); }
  unsigned int GetVoxelVolumeResultion() const { return m_voxelResolution; }
  void SetVoxelVolumeAdaptionRate(float adaptionRate) { m_adaptionRate = adaptionRate; }
  float GetVoxelVolumeAdaptionRate() const { return m_adaptionRate; }

  void SetPerCacheSpecularEnvMapSize(unsigned int specularEnvmapPerCacheSize) {  // renderer.cpp:453-464
    if (specularEnvmapPerCacheSize == 0 || (specularEnvmapPerCacheSize & (specularEnvmapPerCacheSize - 1)) != 0) {
      Fail(DRV_ERR_INVALID, "Per cache specular envmap size needs to be a power of two!");
      return;
    }
    m_specularEnvmapPerCacheSize = specularEnvmapPerCacheSize;
    m_specularEnvmapMaxFillHolesLevel = std::min(m_specularEnvmapMaxFillHolesLevel, Log2(m_specularEnvmapPerCacheSize));
    ReleaseContext();
  }
  unsigned int GetPerCacheSpecularEnvMapSize() const { return m_specularEnvmapPerCacheSize; }
  void SetSpecularEnvMapHoleFillLevel(unsigned int holeFillLevel) {  // renderer.hpp:99
    m_specularEnvmapMaxFillHolesLevel = std::min(holeFillLevel, Log2(m_specularEnvmapPerCacheSize));
    ReleaseContext();
  }
  unsigned int GetSpecularEnvMapHoleFillLevel() const { return m_specularEnvmapMaxFillHolesLevel; }
  // Only the reference's default (direct write) is built; the shared-exponent register path is out of scope.
  void SetSpecularEnvMapDirectWrite(bool directWrite) {
    if (!directWrite) Fail(DRV_ERR_INVALID, "only direct specular map write (the reference's default) is implemented");
  }
  bool GetSpecularEnvMapDirectWrite() const { return true; }

  void SetMaxCacheCount(unsigned int maxCacheCount) { m_maxNumLightCaches = maxCacheCount; ReleaseContext(); }
  unsigned int GetMaxCacheCount() const { return m_maxNumLightCaches; }

  void OnScreenResize(unsigned int width, unsigned int height) {  // renderer.cpp:433-499
    m_width = width;
    m_height = height;
    if (m_hdr) { cudaFree(m_hdr); m_hdr = nullptr; }
    m_gbDepth = nullptr;  // the G-buffer textures are recreated: bind the new ones
    ReleaseContext();
  }

  void SetScene(const std::shared_ptr<const Scene>& scene) { m_scene = scene; ReleaseContext(); }
  const std::shared_ptr<const Scene>& GetScene() const { return m_scene; }

  void SetReadLightCacheCount(bool trackLightCacheCreationStats) {  // renderer.cpp:1145-1153
    m_readLightCacheCount = trackLightCacheCreationStats;
    m_lastNumLightCaches = 0;
  }
  bool GetReadLightCacheCount() const { return m_readLightCacheCount; }
  unsigned int GetLightCacheActiveCount() const { return m_lastNumLightCaches; }

  // ---- address volume (renderer.cpp:1166-1216)
  unsigned int GetCAVCascadeCount() const { return static_cast<unsigned int>(m_CAVCascadeWorldSize.size()); }
  unsigned int GetCAVResolution() const { return m_cavResolution; }
  float GetCAVCascadeWorldSize(unsigned int cascade) const {
    return cascade >= m_CAVCascadeWorldSize.size() ? std::numeric_limits<float>::quiet_NaN() : m_CAVCascadeWorldSize[cascade];
  }
  void SetCAVCascades(unsigned int numCascades, unsigned int resolutionPerCascade) {
    if (numCascades == 0 || resolutionPerCascade == 0 || numCascades > s_maxNumCAVCascades) {
      Fail(DRV_ERR_INVALID, "Invalid address volume cascade settings!");
      return;
    }
    // existing sizes are kept; the first is 4, every new cascade doubles its predecessor
    const size_t had = m_CAVCascadeWorldSize.size();
    m_CAVCascadeWorldSize.resize(numCascades);
    if (had == 0) m_CAVCascadeWorldSize[0] = 4.0f;
    for (size_t i = std::max<size_t>(1, had); i < m_CAVCascadeWorldSize.size(); ++i)
      m_CAVCascadeWorldSize[i] = m_CAVCascadeWorldSize[i - 1] * 2.0f;
    m_cavResolution = resolutionPerCascade;
    ReleaseContext();
  }
  void SetCAVCascadeWorldSize(unsigned int cascade, float cascadeWorldSize) {
    if (cascade >= m_CAVCascadeWorldSize.size() || !(cascadeWorldSize > 0.0f)) {
      Fail(DRV_ERR_INVALID, "Given address volume cascade does not exist / size not positive");
      return;
    }
    m_CAVCascadeWorldSize[cascade] = cascadeWorldSize;
  }
  float GetCAVCascadeTransitionSize() const { return m_CAVCascadeTransitionSize; }
  void SetCAVCascadeTransitionSize(float transitionZoneSize) {  // a value of zero means off
    const bool variantChanges = (m_CAVCascadeTransitionSize > 0) != (transitionZoneSize > 0);
    m_CAVCascadeTransitionSize = transitionZoneSize;
    if (variantChanges) ReleaseContext();
  }

  float GetExposure() const { return m_tonemapExposure; }
  void SetExposure(float exposure) { m_tonemapExposure = exposure; }
  float GetTonemapLMax() const { return m_tonemapLMax; }
  void SetTonemapLMax(float tonemapLMax) { m_tonemapLMax = tonemapLMax; }

  // ---- inputs the reference rasterises itself (DrawSceneToGBuffer / DrawShadowMaps): device pointers, borrowed
  // renderer.cpp:727-738; formats :468-471. roughnessMetallic (RG8) is read only with indirect specular.
  void BindGBuffer(const float* depth, const int16_t* normalRG16I, const uint8_t* diffuseSRGB8X,
                   const uint8_t* roughnessMetallicRG8 = nullptr) {
    m_gbDepth = depth;
    m_gbNormal = normalRG16I;
    m_gbDiffuse = diffuseSRGB8X;
    m_gbRoughMetal = roughnessMetallicRG8;
    if (m_ctx) BindInputs();
  }
  // level 0 of a light's RSM at Light::rsmResolution (formats renderer.cpp:1288-1291)
  void BindShadowMap(unsigned int lightIndex, const uint16_t* fluxRGBX16F, const int16_t* normalRG16I,
                     const uint16_t* depthLinSqRG16F) {
    m_rsms[lightIndex] = Rsm{fluxRGBX16F, normalRG16I, depthLinSqRG16F};
    if (m_ctx) BindInputs();
  }

  // ---- the stage methods (private in the reference, renderer.hpp:148-216; public here so that tests can drive
  // and read back single stages)
  void UpdateConstantUBO() {  // renderer.cpp:290-322
    drv_pack_constant(&m_constant, (int32_t)m_width, (int32_t)m_height, (int32_t)m_voxelResolution, (int32_t)m_cavResolution,
                      (int32_t)m_CAVCascadeWorldSize.size(), m_maxNumLightCaches);
    if (m_indirectSpecular) drv_pack_specular(&m_constant, m_maxNumLightCaches, m_specularEnvmapPerCacheSize);
    if (m_ctx) Check(drv_set_constant(m_ctx, &m_constant));
  }
  void UpdatePerFrameUBO(const Camera& camera) {  // renderer.cpp:324-344
    const drv_camera_desc d = ToDesc(camera);
    drv_pack_per_frame(&m_perFrame, &d, m_passedTime);
    if (Context()) Check(drv_set_per_frame(m_ctx, &m_perFrame));
  }
  void UpdateVolumeUBO(const Camera& camera) {  // renderer.cpp:346-431
    const drv_camera_desc d = ToDesc(camera);
    drv_pack_volume_info(&m_volumeInfo, &d, m_scene->GetBoundingBoxMin(), m_scene->GetBoundingBoxMax(), (int32_t)m_voxelResolution,
                         (int32_t)m_cavResolution, (int32_t)m_CAVCascadeWorldSize.size(), m_CAVCascadeWorldSize.data(),
                         m_CAVCascadeTransitionSize);
    if (Context()) Check(drv_set_volume_info(m_ctx, &m_volumeInfo));
  }
  void PrepareLights() {  // renderer.cpp:664-725
    if (!Context()) return;
    // The blocks are packed every frame like the reference does, but handed to the library only when their bytes
    // changed: a light block can change launch shapes (RSM read resolution, shadow sample interval), so every
    // drv_set_spot_light makes drv_draw_frame run one eager frame before it records its graph again.
    const std::vector<Light>& lights = m_scene->GetLights();
    if (m_uploadedLightCount != (int)lights.size()) {
      Check(drv_set_light_count(m_ctx, (uint32_t)lights.size()));
      m_uploadedLightCount = (int)lights.size();
      m_uploadedSpotLights.clear();
    }
    m_spotLights.resize(lights.size());
    m_uploadedSpotLights.resize(lights.size());
    for (size_t i = 0; i < lights.size(); ++i) {
      const drv_light_desc d = ToDesc(lights[i]);
      drv_pack_spot_light(&m_spotLights[i], &d);
      if (!m_uploadedSpotLights[i].valid || std::memcmp(&m_uploadedSpotLights[i].block, &m_spotLights[i], sizeof(drv_spot_light)) != 0) {
        Check(drv_set_spot_light(m_ctx, (uint32_t)i, &m_spotLights[i]));
        m_uploadedSpotLights[i].block = m_spotLights[i];
        m_uploadedSpotLights[i].valid = true;
      }
    }
  }

  // voxelization.cpp:90-100: what this frame blends, k/255 with the fractional part carried to the next frame
  // (the reference moves m_lastUpdateTime back by it). Returns 0 when nothing is to be done this frame.
  float ConsumeVoxelAdaption(float timeSinceLastBlend) {
    m_voxelAdaptionPending += (double)timeSinceLastBlend * m_adaptionRate * 255.0;
    const double k = std::floor(m_voxelAdaptionPending);
    m_voxelAdaptionPending -= k;
    if (k <= 0.0) return 0.0f;
    return (float)(std::min(k, 255.0) / 255.0);
  }
  void VoxelizeScene(float timeSinceLastBlend) {  // voxelization.cpp:90-176
    const float adaption = ConsumeVoxelAdaption(timeSinceLastBlend);
    if (adaption <= 0.0f || !Context()) return;
    const std::vector<SceneEntity>& ents = m_scene->GetEntities();
    const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (ents.empty()) Check(drv_voxelize(m_ctx, nullptr, 0, identity, adaption, DRV_VOXELIZE_CLEAR | DRV_VOXELIZE_FINISH));
    for (size_t i = 0; i < ents.size(); ++i) {
      const uint32_t flags = (i == 0 ? DRV_VOXELIZE_CLEAR : 0u) | (i + 1 == ents.size() ? DRV_VOXELIZE_FINISH : 0u);
      Check(drv_voxelize(m_ctx, ents[i].devicePositions, ents[i].numTriangles, ents[i].world, adaption, flags));
    }
  }
  void AllocateCaches() {  // renderer.cpp:951-992
    if (!Context()) return;
    if (m_readLightCacheCount) {  // read BEFORE the clear: the count of the previous frame (renderer.cpp:960-966)
      uint32_t n = 0;
      const drv_status st = drv_active_cache_count(m_ctx, &n, nullptr, nullptr);
      if (st == DRV_OK || st == DRV_ERR_CAPACITY) m_lastNumLightCaches = n;
      else Check(st);
    }
    Check(drv_allocate_caches(m_ctx));
  }
  void LightCachesRSM() { if (Context()) Check(drv_light_caches(m_ctx)); }                   // renderer.cpp:899-933
  void PrepareSpecularEnvmaps() { if (Context()) Check(drv_prepare_specular_envmaps(m_ctx)); } // renderer.cpp:994-1045
  // renderer.cpp:1047-1079: additive blend into the HDR back buffer. `target` / `format` let a test read the
  // radiance back unblended (DRV_HDR_RGBA32F_WRITE into a float4 image).
  void ApplyCaches(void* target = nullptr, uint32_t format = DRV_HDR_RGBA16F_ADD) {
    if (!Context()) return;
    if (!target) target = HDRBackbuffer();
    if (target) Check(drv_apply_caches(m_ctx, target, format));
  }
  void ConeTraceAO(float* aoTarget = nullptr) {  // renderer.cpp:936-949; one float per pixel (device memory)
    if (!Context()) return;
    if (!aoTarget) {
      if (!m_ao && cudaMalloc(&m_aoStorage, (size_t)m_width * m_height * sizeof(float)) == cudaSuccess) m_ao = (float*)m_aoStorage;
      if (m_ao) cudaMemsetAsync(m_ao, 0, (size_t)m_width * m_height * sizeof(float), (cudaStream_t)m_stream);
      aoTarget = m_ao;
    }
    if (aoTarget) Check(drv_cone_trace_ao(m_ctx, aoTarget));
  }
  // the tonemap pass at the end of Draw (tonemapping.frag): HDR back buffer -> float4 (device memory)
  void Tonemap(float* ldrRGBA32F) {
    if (Context() && m_hdr) Check(drv_tonemap(m_ctx, m_hdr, m_tonemapExposure, m_tonemapLMax, ldrRGBA32F));
  }

  // ---- what replaces the GL objects
  void* HDRBackbuffer() {  // m_HDRBackbufferTexture: RGBA16F, width x height, device memory
    if (!m_hdr && Stream()) {
      if (cudaMalloc(&m_hdr, (size_t)m_width * m_height * 8) != cudaSuccess) { m_hdr = nullptr; Fail(DRV_ERR_CUDA, "cudaMalloc of the HDR target failed"); }
      else cudaMemsetAsync(m_hdr, 0, (size_t)m_width * m_height * 8, (cudaStream_t)m_stream);
    }
    return m_hdr;
  }
  const float* AOTarget() const { return m_ao; }
  drv_ctx* Context() {  // created on first use from the current settings (≙ AllocateCacheData + the shader variants)
    if (m_ctx) return m_ctx;
    if (!Stream()) return nullptr;
    drv_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    const std::vector<Light>& lights = m_scene->GetLights();
    uint32_t maxRsm = 16;
    for (const Light& l : lights) maxRsm = std::max(maxRsm, NextPowerOfTwo(l.rsmResolution));
    cfg.max_cache_count = m_maxNumLightCaches;
    cfg.cav_cascades = (uint32_t)m_CAVCascadeWorldSize.size();
    cfg.cav_resolution = m_cavResolution;
    cfg.voxel_resolution = m_voxelResolution;
    cfg.sh_order = m_indirectDiffuseMode == IndirectDiffuseMode::SH2 ? 2u : 1u;
    cfg.indirect_shadow = m_indirectShadow ? 1u : 0u;
    cfg.cascade_transitions = m_CAVCascadeTransitionSize > 0.0f ? 1u : 0u;
    cfg.backbuffer_width = m_width;
    cfg.backbuffer_height = m_height;
    cfg.max_lights = std::max<uint32_t>(1u, (uint32_t)lights.size());
    cfg.max_rsm_resolution = maxRsm;
    cfg.device = m_device;
    cfg.stream = m_stream;
    cfg.indirect_specular = m_indirectSpecular ? 1u : 0u;
    cfg.specular_per_cache_size = m_specularEnvmapPerCacheSize;
    cfg.specular_fill_holes_level = m_specularEnvmapMaxFillHolesLevel;
    const drv_status st = drv_create(&cfg, &m_ctx);
    if (st != DRV_OK) {
      m_ctx = nullptr;
      const char* text = drv_last_error(nullptr);
      Fail(st, text ? text : "drv_create failed");
      return nullptr;
    }
    UpdateConstantUBO();
    BindInputs();
    return m_ctx;
  }
  void* Stream() {
    if (!m_stream) {
      cudaStream_t s = nullptr;
      if (cudaSetDevice(m_device) != cudaSuccess || cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
        Fail(DRV_ERR_NO_DEVICE, "no CUDA device: there is no CPU fallback");
        return nullptr;
      }
      m_stream = s;
      m_ownStream = true;
    }
    return m_stream;
  }
  void Finish() { if (m_stream) cudaStreamSynchronize((cudaStream_t)m_stream); }  // ≙ glFinish

  const drv_constant& GetConstantBlock() const { return m_constant; }
  const drv_per_frame& GetPerFrameBlock() const { return m_perFrame; }
  const drv_volume_info& GetVolumeInfoBlock() const { return m_volumeInfo; }
  const std::vector<drv_spot_light>& GetSpotLightBlocks() const { return m_spotLights; }

  drv_status GetLastStatus() const { return m_status; }
  const std::string& GetLastError() const { return m_error; }

 private:
  struct Rsm { const uint16_t* flux; const int16_t* normal; const uint16_t* depth; };

  static unsigned int Log2(unsigned int v) { unsigned int l = 0; while (v >>= 1) ++l; return l; }
  static uint32_t NextPowerOfTwo(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

  void DrawOverlapped(float timeSinceLastFrame) {
    if (!Context() || !HDRBackbuffer()) return;
    uint32_t flags = DRV_FRAME_PREPARE_RSM | DRV_FRAME_GRAPH;
    if (m_indirectShadow) {
      const float adaption = ConsumeVoxelAdaption(timeSinceLastFrame);
      if (adaption > 0.0f) {
        const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        const std::vector<SceneEntity>& ents = m_scene->GetEntities();
        const float* tris = ents.empty() ? nullptr : ents[0].devicePositions;
        const uint32_t n = ents.empty() ? 0u : ents[0].numTriangles;
        const float* world = ents.empty() ? identity : ents[0].world;
        // a changed binding makes the next frame re-record its graph: only rebind what changed
        if (tris != m_boundTris || n != m_boundNumTris || adaption != m_boundAdaption || std::memcmp(world, m_boundWorld, 64) != 0) {
          Check(drv_bind_scene(m_ctx, tris, n, world, adaption));
          m_boundTris = tris; m_boundNumTris = n; m_boundAdaption = adaption;
          std::memcpy(m_boundWorld, world, 64);
        }
        flags |= DRV_FRAME_VOXELIZE;
      }
    }
    if (m_readLightCacheCount) {  // the count of the previous frame, as in AllocateCaches
      uint32_t count = 0;
      const drv_status st = drv_active_cache_count(m_ctx, &count, nullptr, nullptr);
      if (st == DRV_OK || st == DRV_ERR_CAPACITY) m_lastNumLightCaches = count;
      else Check(st);
    }
    Check(drv_draw_frame(m_ctx, m_hdr, DRV_HDR_RGBA16F_WRITE, flags));
  }
  void ClearHDRBackbuffer() {
    if (HDRBackbuffer()) cudaMemsetAsync(m_hdr, 0, (size_t)m_width * m_height * 8, (cudaStream_t)m_stream);
  }
  void BindInputs() {
    if (m_gbDepth) {
      Check(drv_bind_gbuffer(m_ctx, m_gbDepth, m_gbNormal, m_gbDiffuse, m_width, m_height));
      if (m_gbRoughMetal && m_indirectSpecular) Check(drv_bind_gbuffer_material(m_ctx, m_gbRoughMetal));
    }
    const std::vector<Light>& lights = m_scene->GetLights();
    for (const auto& kv : m_rsms) {
      if (kv.first >= lights.size()) continue;
      Check(drv_bind_rsm(m_ctx, kv.first, kv.second.flux, kv.second.normal, kv.second.depth, lights[kv.first].rsmResolution));
      Check(drv_prepare_rsm(m_ctx, kv.first));  // ShadowMap::PrepareRSM, renderer.cpp:1300-1339
    }
  }
  void ReleaseContext() {
    if (m_ctx) drv_destroy(m_ctx);
    m_ctx = nullptr;
    m_boundTris = nullptr; m_boundNumTris = 0; m_boundAdaption = -1.0f;  // a new context has no scene bound ...
    m_uploadedLightCount = -1;                                          // ... and no lights
    m_uploadedSpotLights.clear();
    if (m_aoStorage) { cudaFree(m_aoStorage); m_aoStorage = nullptr; m_ao = nullptr; }
  }
  void Check(drv_status st) {
    if (st != DRV_OK && m_status == DRV_OK) {
      const char* text = m_ctx ? drv_last_error(m_ctx) : nullptr;
      Fail(st, text ? text : "libdrv_gi call failed");
    }
  }
  void Fail(drv_status st, const char* text) {
    if (m_status == DRV_OK) { m_status = st; m_error = text; }
  }

  std::shared_ptr<const Scene> m_scene;
  unsigned int m_width, m_height;
  int m_device;
  void* m_stream;
  bool m_ownStream = false;
  drv_ctx* m_ctx = nullptr;

  // constructor defaults of renderer.cpp:36-51
  bool m_readLightCacheCount = false;
  unsigned int m_lastNumLightCaches = 0;
  float m_tonemapExposure = 1.0f;
  float m_tonemapLMax = 1.2f;
  Mode m_mode = Mode::DYN_RADIANCE_VOLUME;
  IndirectDiffuseMode m_indirectDiffuseMode = IndirectDiffuseMode::SH1;
  unsigned int m_specularEnvmapPerCacheSize = 16;
  unsigned int m_specularEnvmapMaxFillHolesLevel = 0;
  float m_CAVCascadeTransitionSize = 2.0f;
  bool m_indirectShadow = true;
  bool m_indirectSpecular = false;
  float m_passedTime = 0.0f;
  unsigned int m_maxNumLightCaches = 16384;
  std::vector<float> m_CAVCascadeWorldSize;
  unsigned int m_cavResolution = 0;
  unsigned int m_voxelResolution = 128;  // Voxelization(128), renderer.cpp:84
  float m_adaptionRate = 10.0f;          // Voxelization ctor, voxelization.cpp:25
  double m_voxelAdaptionPending = 0.0;

  drv_constant m_constant{};
  drv_per_frame m_perFrame{};
  drv_volume_info m_volumeInfo{};
  std::vector<drv_spot_light> m_spotLights;

  const float* m_gbDepth = nullptr;
  const int16_t* m_gbNormal = nullptr;
  const uint8_t* m_gbDiffuse = nullptr;
  const uint8_t* m_gbRoughMetal = nullptr;
  std::map<unsigned int, Rsm> m_rsms;
  void* m_hdr = nullptr;
  void* m_aoStorage = nullptr;
  float* m_ao = nullptr;

  struct UploadedLight { drv_spot_light block; bool valid = false; };
  int m_uploadedLightCount = -1;
  std::vector<UploadedLight> m_uploadedSpotLights;
  bool m_overlappedFrame = false;
  const float* m_boundTris = nullptr;
  uint32_t m_boundNumTris = 0;
  float m_boundAdaption = -1.0f;
  float m_boundWorld[16] = {0};

  drv_status m_status = DRV_OK;
  std::string m_error;
};

}  // namespace drv
