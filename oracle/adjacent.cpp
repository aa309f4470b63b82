/*
 * adjacent.cpp — CPU oracle for the rows SURVEY.md 8(f) ranks next to the hot path (TEST INFRASTRUCTURE ONLY,
 * see oracle.h): the RSM producer's flux model (f1), voxel cone-traced ambient occlusion (f2) and the tonemap
 * that follows the apply pass (f3). Same arithmetic policy as the rest of the oracle.
 * Citations are relative to /root/reference/DynamicRadianceVolume/.
 */
#include "oracle.h"
#include "glsl_scalar.h"

using namespace orc;

/* f1 — shader/fillrsm.frag:32-61 with ComputeSpotFalloff (shader/lightingfunctions.glsl:3-10). The rasteriser's
 * per-fragment attributes arrive as arrays: world position, the (normal-mapped, unnormalised) shading normal and
 * the base colour the sampler returned (linear). coverage == 0 (or NULL = all covered): no fragment, the texel
 * keeps the clear value 0 of renderer.cpp:793. */
extern "C" void orc_fill_rsm(const drv_spot_light* L, const float* position_xyz, const float* normal_xyz,
                             const float* basecolor_rgb, const uint8_t* coverage, uint32_t res,
                             uint16_t* flux_rgbx16f, int16_t* normal_rg16i, uint16_t* depthlinsq_rg16f) {
  const float PI = GLSL_PI, PI_2 = 6.28318530717958f; /* utils.glsl:1-2 */
  const vec3 lightPos = V3(L->LightPosition), lightDir = V3(L->LightDirection);
  const float R = (float)L->RSMRenderResolution;
  for (size_t t = 0; t < (size_t)res * res; ++t) {
    for (int c = 0; c < 4; ++c) flux_rgbx16f[t * 4 + c] = 0;
    normal_rg16i[t * 2] = normal_rg16i[t * 2 + 1] = 0;
    depthlinsq_rg16f[t * 2] = depthlinsq_rg16f[t * 2 + 1] = 0;
    if (coverage && !coverage[t]) continue;
    vec3 toLight = lightPos - V3(position_xyz + t * 3);                       /* :38 */
    float distToLight = length(toLight);                                     /* :39 */
    toLight = toLight / distToLight;                                         /* :40 */
    float cosToLight = saturate(dot(-toLight, lightDir));                    /* :42 */
    float totalSpotSteradian = PI_2 * (1.0f - L->LightCosHalfAngle);         /* :44 */
    float pixelSteradian = totalSpotSteradian * cosToLight / R / R;          /* :45 */
    float spotFalloff = saturate(cosToLight - L->LightCosHalfAngle) / (1.0f - L->LightCosHalfAngle); /* lightingfunctions.glsl:9 */
    float k = spotFalloff * pixelSteradian / PI;                             /* :53 */
    for (int c = 0; c < 3; ++c)
      flux_rgbx16f[t * 4 + c] = float_to_half(basecolor_rgb[t * 3 + c] * L->LightIntensity[c] * k);
    depthlinsq_rg16f[t * 2] = float_to_half(distToLight);                    /* :54 */
    depthlinsq_rg16f[t * 2 + 1] = float_to_half(distToLight * distToLight);
    int16_t ox, oy;
    pack_normal16i(normalize(V3(normal_xyz + t * 3)), ox, oy);               /* :68 */
    normal_rg16i[t * 2] = ox;
    normal_rg16i[t * 2 + 1] = oy;
  }
}

/* f2 — shader/ambientocclusion.frag:25-89 (Renderer::ConeTraceAO, renderer.cpp:936-949): six cones around the
 * normal through the voxel chain. out: one float per pixel; discarded pixels (:31-32) are left untouched. */
extern "C" void orc_cone_trace_ao(const drv_per_frame* pf, const drv_volume_info* vi, const uint8_t* voxel_chain,
                                  uint32_t voxel_res, const float* depth, const int16_t* normal_rg16i, uint32_t W,
                                  uint32_t H, float* out, int threads) {
  const float PI = GLSL_PI;
  const float dirs[6][4] = {                                                 /* :39-47 */
      {0.0f, 1.0f, 0.0f, PI / 4.0f},
      {0.0f, 0.5f, 0.866025f, 3.0f * PI / 20.0f},
      {0.823639f, 0.5f, 0.267617f, 3.0f * PI / 20.0f},
      {0.509037f, 0.5f, -0.700629f, 3.0f * PI / 20.0f},
      {-0.509037f, 0.5f, -0.700629f, 3.0f * PI / 20.0f},
      {-0.823639f, 0.5f, 0.267617f, 3.0f * PI / 20.0f},
  };
  const float distToSphereRad = 0.5f; /* sin(PI / 3 * 0.5), :48-49: 0.4999999999999993 rounds to 0.5f */
  const float volumeSize = (float)voxel_res;
  parallel_for((int64_t)H, threads, [&](int64_t y0, int64_t y1, int) {
    for (int64_t y = y0; y < y1; ++y)
      for (uint32_t x = 0; x < W; ++x) {
        const size_t t = (size_t)y * W + x;
        const float d = depth[t];
        if (d < 0.000001f) continue;                                         /* :31-32 */
        const float tx = ((float)x + 0.5f) / (float)W, ty = ((float)y + 0.5f) / (float)H;
        const float clip[4] = {tx * 2.0f - 1.0f, ty * 2.0f - 1.0f, d, 1.0f};
        float w4[4];
        mul_row_major(pf->InverseViewProjection, clip, w4);                  /* :33 */
        const vec3 worldPosition = V3(w4[0] / w4[3], w4[1] / w4[3], w4[2] / w4[3]);
        const vec3 n = unpack_normal16i(normal_rg16i[t * 2], normal_rg16i[t * 2 + 1]); /* :35 */
        vec3 U = cross(n, V3(0.0f, 1.0f, 0.0f));                             /* CreateONB, :16-23 */
        if (std::fabs(U.x) < 0.0001f && std::fabs(U.y) < 0.0001f && std::fabs(U.z) < 0.0001f) U = cross(n, V3(1.0f, 0.0f, 0.0f));
        U = normalize(U);
        const vec3 V = cross(n, U);
        const vec3 startWorld = worldPosition + n * vi->VoxelSizeInWorld * 1.6f; /* :57 */
        const vec3 startVoxel = (startWorld - V3(vi->VolumeWorldMin)) / (vi->VoxelSizeInWorld * volumeSize); /* :58 */
        float total = 0.0f;
        for (int k = 0; k < 6; ++k) {
          const vec3 dirWorld = (V * dirs[k][0] + n * dirs[k][1]) + U * dirs[k][2]; /* :64 */
          const vec3 dirVoxel = dirWorld / volumeSize;                       /* :65 */
          vec3 p = startVoxel;
          float stepSize = 1.0f, dist = 0.0f, coneWeight = 0.0f;
          for (int s = 0; s < 16 && coneWeight < 0.99f && saturate(p.x) == p.x && saturate(p.y) == p.y && saturate(p.z) == p.z; ++s) { /* :73-74 */
            p = p + dirVoxel * stepSize;
            dist += stepSize;
            const float radius = dist * distToSphereRad;
            const float pp[3] = {p.x, p.y, p.z};
            const float occ = orc_sample_voxel(voxel_chain, voxel_res, pp, std::log2(radius)); /* :82 */
            coneWeight += (1.0f - coneWeight) * occ;
            stepSize = radius * 2.0f;
          }
          total += coneWeight * dirs[k][3] / 6.0f;                           /* :87 */
        }
        out[t] = saturate(1.0f - total);                                     /* :90 */
      }
  });
}

/* f3 — shader/tonemapping.frag:21-31: Drago operator on the exposed colour; drago_divider = log2(LMax + 1)
 * (renderer.cpp:1225-1227). hdr: n RGBA float texels; out: n RGB floats. */
extern "C" void orc_tonemap(const float* hdr_rgba, uint32_t n, float exposure, float drago_divider, float* out_rgb) {
  for (uint32_t i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c) out_rgb[i * 3 + c] = std::log2(hdr_rgba[i * 4 + c] * exposure + 1.0f) / drago_divider;
}
