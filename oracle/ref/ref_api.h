/*
 * ref_api.h — glue between the C-ABI PODs of this repo (include/drv_gi.h) and the generated shader structs of
 * oracle/ref/glsl2cpp.py. TEST INFRASTRUCTURE ONLY. The reference reflects uniform offsets at run time
 * (renderer.cpp:60-83, `layout(shared)`), i.e. it sets uniforms BY NAME; so does this header.
 */
#ifndef DRV_REF_API_H
#define DRV_REF_API_H

#include "../../include/drv_gi.h"
#include "glsl_compat.h"

#define REF_CAT2(a, b) a##_##b
#define REF_CAT(a, b) REF_CAT2(a, b)
#define REF_FN(name) REF_CAT(ref_##name, REF_VARIANT)

namespace glsl {

inline vec3 v3(const float* p) { return vec3(p[0], p[1], p[2]); }
inline mat4 m4(const float* p) { /* raw bytes: column c = floats 4c..4c+3 */
  mat4 m;
  for (int c = 0; c < 4; ++c) m[c] = vec4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
  return m;
}

/* globalubos.glsl:2-30 */
template <class S>
inline void set_constant(const drv_constant* cb) {
  S::ShCosLobeFactor0 = cb->ShCosLobeFactor0;
  S::ShCosLobeFactor1 = cb->ShCosLobeFactor1;
  S::ShCosLobeFactor2n2_p1_n1 = cb->ShCosLobeFactor2n2_p1_n1;
  S::ShCosLobeFactor20 = cb->ShCosLobeFactor20;
  S::ShCosLobeFactor2p2 = cb->ShCosLobeFactor2p2;
  S::ShEvaFactor0 = cb->ShEvaFactor0;
  S::ShEvaFactor1 = cb->ShEvaFactor1;
  S::ShEvaFactor2n2_p1_n1 = cb->ShEvaFactor2n2_p1_n1;
  S::ShEvaFactor20 = cb->ShEvaFactor20;
  S::ShEvaFactor2p2 = cb->ShEvaFactor2p2;
  S::BackbufferResolution = ivec2(cb->BackbufferResolution[0], cb->BackbufferResolution[1]);
  S::VoxelResolution = cb->VoxelResolution;
  S::AddressVolumeResolution = cb->AddressVolumeResolution;
  S::NumAddressVolumeCascades = cb->NumAddressVolumeCascades;
  S::MaxNumLightCaches = cb->MaxNumLightCaches;
  S::SpecularEnvmapTotalSize = cb->SpecularEnvmapTotalSize;
  S::SpecularEnvmapPerCacheSize_Texel = cb->SpecularEnvmapPerCacheSize_Texel;
  S::SpecularEnvmapPerCacheSize_Texcoord = cb->SpecularEnvmapPerCacheSize_Texcoord;
  S::SpecularEnvmapNumCachesPerDimension = cb->SpecularEnvmapNumCachesPerDimension;
}
/* globalubos.glsl:33-43 */
template <class S>
inline void set_per_frame(const drv_per_frame* pf) {
  S::Projection = m4(pf->Projection);
  S::ViewProjection = m4(pf->ViewProjection);
  S::InverseView = m4(pf->InverseView);
  S::InverseViewProjection = m4(pf->InverseViewProjection);
  S::CameraPosition = v3(pf->CameraPosition);
  S::CameraDirection = v3(pf->CameraDirection);
  S::PassedTime = pf->PassedTime;
}
/* globalubos.glsl:46-79 */
template <class S>
inline void set_volume_info(const drv_volume_info* vi) {
  S::VolumeWorldMin = v3(vi->VolumeWorldMin);
  S::VoxelSizeInWorld = vi->VoxelSizeInWorld;
  S::VolumeWorldMax = v3(vi->VolumeWorldMax);
  S::CAVTransitionZoneSize = vi->CAVTransitionZoneSize;
  for (int c = 0; c < DRV_MAX_CASCADES; ++c) {
    const drv_cav_cascade& k = vi->AddressVolumeCascades[c];
    S::AddressVolumeCascades[c].Min = v3(k.Min);
    S::AddressVolumeCascades[c].WorldVoxelSize = k.WorldVoxelSize;
    S::AddressVolumeCascades[c].Max = v3(k.Max);
    S::AddressVolumeCascades[c].DecisionMin = v3(k.DecisionMin);
    S::AddressVolumeCascades[c].DecisionMax = v3(k.DecisionMax);
  }
}
/* globalubos.glsl:88-112 */
template <class S>
inline void set_spot_light(const drv_spot_light* L) {
  S::LightIntensity = v3(L->LightIntensity);
  S::ShadowNormalOffset = L->ShadowNormalOffset;
  S::ShadowBias = L->ShadowBias;
  S::LightPosition = v3(L->LightPosition);
  S::LightDirection = v3(L->LightDirection);
  S::LightCosHalfAngle = L->LightCosHalfAngle;
  S::LightViewProjection = m4(L->LightViewProjection);
  S::InverseLightViewProjection = m4(L->InverseLightViewProjection);
  S::RSMRenderResolution = L->RSMRenderResolution;
  S::RSMReadResolution = L->RSMReadResolution;
  S::ValAreaFactor = L->ValAreaFactor;
  S::IndirectShadowComputationLod = L->IndirectShadowComputationLod;
  S::IndirectShadowComputationBlockSize = L->IndirectShadowComputationBlockSize;
  S::IndirectShadowComputationSampleInterval = L->IndirectShadowComputationSampleInterval;
  S::IndirectShadowComputationSuperValWidth = L->IndirectShadowComputationSuperValWidth;
  S::IndirectShadowSamplingOffset = L->IndirectShadowSamplingOffset;
}

} // namespace glsl
#endif
