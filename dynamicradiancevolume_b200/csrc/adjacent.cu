// adjacent.cu — the rows SURVEY.md 8(f) ranks next to the hot path, to the same parity bar:
//   f1  drv_fill_rsm        the RSM producer's flux model (shader/fillrsm.frag:32-61): the step just before the path
//   f2  drv_cone_trace_ao   voxel cone-traced ambient occlusion (shader/ambientocclusion.frag:25-89;
//                           Renderer::ConeTraceAO, renderer.cpp:936-949): a second consumer of the record chain
//   f3  drv_tonemap         the Drago tonemap after the apply pass (shader/tonemapping.frag:21-31) and
//       drv_save_to_pfm     Renderer::SaveToPFM (renderer.cpp:1229-1235) / WritePfm (rendering/hdrimage.cpp:6-32)
#include "ctx.h"
#include "device_math.cuh"
#include "voxel_sample.cuh"

#include <cstdio>
#include <vector>

using namespace drvk;

namespace {

// ---- f1 ----------------------------------------------------------------------------------------------------
// One thread per RSM texel. Everything is decision maths (separately rounded IEEE operations in the shader's
// order) so that the half-float outputs are bit-identical to the oracle's; the normal goes through atan2f and may
// differ by one int16 code.
__global__ void __launch_bounds__(256) fill_rsm_kernel(drv_spot_light L, const float* __restrict__ pos,
                                                       const float* __restrict__ nrm, const float* __restrict__ base,
                                                       const uint8_t* __restrict__ coverage, uint32_t texels,
                                                       uint2* __restrict__ flux, int* __restrict__ normal,
                                                       uint32_t* __restrict__ depth) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= texels) return;
  if (coverage && !coverage[t]) { // no fragment: the clear value of renderer.cpp:793
    flux[t] = make_uint2(0u, 0u);
    normal[t] = 0;
    depth[t] = 0u;
    return;
  }
  const float PI = DRV_GLSL_PI, PI_2 = 6.28318530717958f; // utils.glsl:1-2
  const float R = (float)L.RSMRenderResolution;
  float tx = ex_sub(L.LightPosition[0], pos[t * 3]), ty = ex_sub(L.LightPosition[1], pos[t * 3 + 1]),
        tz = ex_sub(L.LightPosition[2], pos[t * 3 + 2]);                                      // :38
  const float dist = ex_sqrt(ex_dot3(tx, ty, tz, tx, ty, tz));                                 // :39
  tx = ex_div(tx, dist); ty = ex_div(ty, dist); tz = ex_div(tz, dist);                         // :40
  const float cosToLight = saturatef(ex_dot3(-tx, -ty, -tz, L.LightDirection[0], L.LightDirection[1], L.LightDirection[2])); // :42
  const float totalSpotSteradian = ex_mul(PI_2, ex_sub(1.0f, L.LightCosHalfAngle));            // :44
  const float pixelSteradian = ex_div(ex_div(ex_mul(totalSpotSteradian, cosToLight), R), R);   // :45
  const float spotFalloff = ex_div(saturatef(ex_sub(cosToLight, L.LightCosHalfAngle)), ex_sub(1.0f, L.LightCosHalfAngle));
  const float k = ex_div(ex_mul(spotFalloff, pixelSteradian), PI);                             // :53
  const uint32_t fr = float_to_half_bits(ex_mul(ex_mul(base[t * 3], L.LightIntensity[0]), k));
  const uint32_t fg = float_to_half_bits(ex_mul(ex_mul(base[t * 3 + 1], L.LightIntensity[1]), k));
  const uint32_t fb = float_to_half_bits(ex_mul(ex_mul(base[t * 3 + 2], L.LightIntensity[2]), k));
  flux[t] = make_uint2(fr | (fg << 16), fb);
  depth[t] = (uint32_t)float_to_half_bits(dist) | ((uint32_t)float_to_half_bits(ex_mul(dist, dist)) << 16); // :54
  const float nx = nrm[t * 3], ny = nrm[t * 3 + 1], nz = nrm[t * 3 + 2];
  const float inv = ex_rsqrt(ex_dot3(nx, ny, nz, nx, ny, nz));                                 // :68 normalize
  int ox, oy;
  pack_normal16i(ex_mul(nx, inv), ex_mul(ny, inv), ex_mul(nz, inv), ox, oy);
  normal[t] = (int)(((uint32_t)ox & 0xffffu) | ((uint32_t)oy << 16));
}

// ---- f2 ----------------------------------------------------------------------------------------------------
struct AoParams {
  int W, H;
  float ivp[16];
  float vmin[3];
  float voxel_size;
  const uint2* rec;
  uint32_t rec_offset[16];
  int vres, vlevels;
};

// One thread per pixel, six cones of at most 16 steps each. Positions and the two loop conditions (inside the
// volume, coneWeight < 0.99) are evaluated on separately rounded operations in the shader's order so that the trip
// counts follow the oracle; sampling and accumulation are continuous maths.
__global__ void __launch_bounds__(256) cone_trace_ao_kernel(const __grid_constant__ AoParams p, const float* __restrict__ depth,
                                                            const int* __restrict__ normal, const float* __restrict__ ndc_xy,
                                                            float* __restrict__ out) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= p.W || y >= p.H) return;
  const uint32_t t = (uint32_t)y * p.W + x;
  const float d = __ldg(depth + t);
  if (d < 0.000001f) return;                                                                   // :31-32 discard
  const F3 wp = ex_unproject(p.ivp, __ldg(ndc_xy + x), __ldg(ndc_xy + p.W + y), d);             // :33-34
  const int pn = __ldg(normal + t);
  const F3 n = unpack_normal16i((int)(short)(pn & 0xffff), (int)(short)((uint32_t)pn >> 16));  // :35
  // CreateONB, :16-23
  float ux = ex_sub(ex_mul(n.y, 0.0f), ex_mul(n.z, 1.0f)), uy = ex_sub(ex_mul(n.z, 0.0f), ex_mul(n.x, 0.0f)),
        uz = ex_sub(ex_mul(n.x, 1.0f), ex_mul(n.y, 0.0f));
  if (fabsf(ux) < 0.0001f && fabsf(uy) < 0.0001f && fabsf(uz) < 0.0001f) {
    ux = ex_sub(ex_mul(n.y, 0.0f), ex_mul(n.z, 0.0f)); uy = ex_sub(ex_mul(n.z, 1.0f), ex_mul(n.x, 0.0f));
    uz = ex_sub(ex_mul(n.x, 0.0f), ex_mul(n.y, 1.0f));
  }
  const float ui = ex_rsqrt(ex_dot3(ux, uy, uz, ux, uy, uz));
  ux = ex_mul(ux, ui); uy = ex_mul(uy, ui); uz = ex_mul(uz, ui);
  const float vx = ex_sub(ex_mul(n.y, uz), ex_mul(n.z, uy)), vy = ex_sub(ex_mul(n.z, ux), ex_mul(n.x, uz)),
              vz = ex_sub(ex_mul(n.x, uy), ex_mul(n.y, ux));
  const float volumeSize = (float)p.vres;
  const float denom = ex_mul(p.voxel_size, volumeSize);
  const float sx = ex_div(ex_sub(ex_add(wp.x, ex_mul(ex_mul(n.x, p.voxel_size), 1.6f)), p.vmin[0]), denom);  // :57-58
  const float sy = ex_div(ex_sub(ex_add(wp.y, ex_mul(ex_mul(n.y, p.voxel_size), 1.6f)), p.vmin[1]), denom);
  const float sz = ex_div(ex_sub(ex_add(wp.z, ex_mul(ex_mul(n.z, p.voxel_size), 1.6f)), p.vmin[2]), denom);
  VoxelVol V;
  V.rec = p.rec; V.rec_offset = p.rec_offset; V.res = p.vres; V.levels = p.vlevels; V.voxel_size = p.voxel_size;
  V.vmin[0] = p.vmin[0]; V.vmin[1] = p.vmin[1]; V.vmin[2] = p.vmin[2];
  const float PI = DRV_GLSL_PI;
  const float dirs[6][4] = {                                                                   // :39-47
      {0.0f, 1.0f, 0.0f, PI / 4.0f},
      {0.0f, 0.5f, 0.866025f, 3.0f * PI / 20.0f},
      {0.823639f, 0.5f, 0.267617f, 3.0f * PI / 20.0f},
      {0.509037f, 0.5f, -0.700629f, 3.0f * PI / 20.0f},
      {-0.509037f, 0.5f, -0.700629f, 3.0f * PI / 20.0f},
      {-0.823639f, 0.5f, 0.267617f, 3.0f * PI / 20.0f},
  };
  float total = 0.0f;
#pragma unroll 1
  for (int k = 0; k < 6; ++k) {
    // dirInWorld = S.x * V + S.y * n + S.z * U (:64), / volumeSize (:65)
    const float dx = ex_div(ex_add(ex_add(ex_mul(vx, dirs[k][0]), ex_mul(n.x, dirs[k][1])), ex_mul(ux, dirs[k][2])), volumeSize);
    const float dy = ex_div(ex_add(ex_add(ex_mul(vy, dirs[k][0]), ex_mul(n.y, dirs[k][1])), ex_mul(uy, dirs[k][2])), volumeSize);
    const float dz = ex_div(ex_add(ex_add(ex_mul(vz, dirs[k][0]), ex_mul(n.z, dirs[k][1])), ex_mul(uz, dirs[k][2])), volumeSize);
    float px = sx, py = sy, pz = sz, stepSize = 1.0f, dist = 0.0f, coneWeight = 0.0f;
#pragma unroll 1
    for (int s = 0; s < 16 && coneWeight < 0.99f && saturatef(px) == px && saturatef(py) == py && saturatef(pz) == pz; ++s) { // :73-74
      px = ex_add(px, ex_mul(dx, stepSize)); py = ex_add(py, ex_mul(dy, stepSize)); pz = ex_add(pz, ex_mul(dz, stepSize));
      dist = ex_add(dist, stepSize);
      const float radius = ex_mul(dist, 0.5f);                                                 // sin(PI / 3 * 0.5) = 0.5f
      const float occ = sample_voxel_records(V, px, py, pz, __log2f(radius));                   // :82
      coneWeight = fmaf(1.0f - coneWeight, occ, coneWeight);
      stepSize = ex_mul(radius, 2.0f);
    }
    total += coneWeight * dirs[k][3] / 6.0f;                                                   // :87
  }
  out[t] = saturatef(1.0f - total);                                                            // :90
}

// ---- f3 ----------------------------------------------------------------------------------------------------
// Two pixels per thread: one 16-byte load, two 16-byte stores (a streaming pass: 8 B in, 16 B out per pixel).
__device__ __forceinline__ float4 drago(uint2 v, float exposure, float inv_divider) {
  const float2 rg = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
  const float2 ba = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  // Drago: log2(exposedColor + 1) / DragoDivider, tonemapping.frag:21-24, 29-31
  return make_float4(log2f(fmaf(rg.x, exposure, 1.0f)) * inv_divider, log2f(fmaf(rg.y, exposure, 1.0f)) * inv_divider,
                     log2f(fmaf(ba.x, exposure, 1.0f)) * inv_divider, 1.0f);
}
__global__ void __launch_bounds__(256) tonemap_kernel(const uint2* __restrict__ hdr16, uint32_t n, float exposure,
                                                      float divider, float4* __restrict__ out) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * 2u;
  const float inv_divider = 1.0f / divider;
  if (i + 1 < n) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(hdr16 + i));
    __stcs(out + i, drago(make_uint2(v.x, v.y), exposure, inv_divider));
    __stcs(out + i + 1, drago(make_uint2(v.z, v.w), exposure, inv_divider));
  } else if (i < n) {
    out[i] = drago(__ldg(hdr16 + i), exposure, inv_divider);
  }
}

} // namespace

extern "C" drv_status drv_fill_rsm(drv_ctx* ctx, uint32_t light, const float* position_xyz, const float* normal_xyz,
                                   const float* basecolor_rgb, const uint8_t* coverage, uint32_t resolution) {
  if (!ctx) return DRV_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (light >= ctx->cfg.max_lights || !position_xyz || !normal_xyz || !basecolor_rgb)
    return ctx->fail(DRV_ERR_INVALID, "drv_fill_rsm: bad argument");
  LightState& S = ctx->lights[light];
  if (!S.block_set) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_fill_rsm: SpotLight block not set");
  if (resolution == 0 || (resolution & (resolution - 1)) || resolution > ctx->cfg.max_rsm_resolution ||
      (uint32_t)S.block.RSMRenderResolution != resolution)
    return ctx->fail(DRV_ERR_INVALID, "drv_fill_rsm: resolution must equal SpotLight.RSMRenderResolution (a power of two <= max_rsm_resolution)");
  const size_t cap = (size_t)ctx->cfg.max_rsm_resolution * ctx->cfg.max_rsm_resolution;
  if (!S.st_flux) {
    DRV_CUDA(cudaMalloc(&S.st_flux, cap * 8));
    DRV_CUDA(cudaMalloc(&S.st_normal, cap * 4));
    DRV_CUDA(cudaMalloc(&S.st_depth, cap * 4));
  }
  const uint32_t texels = resolution * resolution;
  fill_rsm_kernel<<<(texels + 255) / 256, 256, 0, ctx->stream>>>(S.block, position_xyz, normal_xyz, basecolor_rgb, coverage,
                                                                texels, (uint2*)S.st_flux, (int*)S.st_normal,
                                                                (uint32_t*)S.st_depth);
  DRV_LAUNCH_CHECK();
  return drv_bind_rsm(ctx, light, S.st_flux, S.st_normal, S.st_depth, resolution);
}

extern "C" drv_status drv_cone_trace_ao(drv_ctx* ctx, float* ao_out) {
  if (!ctx) return DRV_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (!ao_out) return ctx->fail(DRV_ERR_INVALID, "drv_cone_trace_ao: null output");
  if (!ctx->have_per_frame || !ctx->have_volume) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_cone_trace_ao: uniform blocks not set");
  if (!ctx->gb_depth || !ctx->gb_normal) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_cone_trace_ao: g-buffer not bound");
  if (!ctx->voxel_records) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_cone_trace_ao: the context was created without indirect_shadow (no voxel records)");
  AoParams p;
  memset(&p, 0, sizeof(p));
  p.W = (int)ctx->gb_w; p.H = (int)ctx->gb_h;
  memcpy(p.ivp, ctx->per_frame.InverseViewProjection, sizeof(p.ivp));
  memcpy(p.vmin, ctx->volume.VolumeWorldMin, 12);
  p.voxel_size = ctx->volume.VoxelSizeInWorld;
  p.rec = ctx->voxel_records;
  for (int l = 0; l < 16; ++l) p.rec_offset[l] = ctx->voxel_record_offset[l];
  p.vres = (int)ctx->cfg.voxel_resolution;
  p.vlevels = (int)ctx->voxel_levels;
  dim3 block(32, 8), grid((p.W + 31) / 32, (p.H + 7) / 8);
  cone_trace_ao_kernel<<<grid, block, 0, ctx->stream>>>(p, ctx->gb_depth, (const int*)ctx->gb_normal, ctx->ndc_xy, ao_out);
  DRV_LAUNCH_CHECK();
  return DRV_OK;
}

extern "C" drv_status drv_tonemap(drv_ctx* ctx, const void* hdr_rgba16f, float exposure, float l_max, float* ldr_rgba32f) {
  if (!ctx) return DRV_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (!hdr_rgba16f || !ldr_rgba32f) return ctx->fail(DRV_ERR_INVALID, "drv_tonemap: null argument");
  const uint32_t n = ctx->cfg.backbuffer_width * ctx->cfg.backbuffer_height;
  const float divider = log2f(l_max + 1.0f); // renderer.cpp:1226
  tonemap_kernel<<<((n + 1) / 2 + 255) / 256, 256, 0, ctx->stream>>>((const uint2*)hdr_rgba16f, n, exposure, divider, (float4*)ldr_rgba32f);
  DRV_LAUNCH_CHECK();
  return DRV_OK;
}

// WritePfm, rendering/hdrimage.cpp:6-32: "PF\n", "<w> <h>\n", "-1.000000\n", then the RGB floats of the RGBA image in
// memory order. Pure host code.
extern "C" drv_status drv_write_pfm(const char* path, const float* rgba, uint32_t width, uint32_t height) {
  if (!path || !rgba) return DRV_ERR_INVALID;
  FILE* f = fopen(path, "wb");
  if (!f) return DRV_ERR_INVALID;
  fwrite("PF\n", 1, 3, f);
  fprintf(f, "%u %u\n", width, height);
  fwrite("-1.000000\n", 1, 10, f);
  for (size_t i = 0; i < (size_t)width * height; ++i) fwrite(rgba + i * 4, sizeof(float), 3, f);
  const bool ok = !ferror(f);
  fclose(f);
  return ok ? DRV_OK : DRV_ERR_INVALID;
}

// Renderer::SaveToPFM, renderer.cpp:1229-1235: read the HDR target back as RGBA float and write it.
extern "C" drv_status drv_save_to_pfm(drv_ctx* ctx, const void* hdr_rgba16f, const char* path) {
  if (!ctx) return DRV_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (!hdr_rgba16f || !path) return ctx->fail(DRV_ERR_INVALID, "drv_save_to_pfm: null argument");
  const uint32_t W = ctx->cfg.backbuffer_width, H = ctx->cfg.backbuffer_height;
  std::vector<uint16_t> half((size_t)W * H * 4);
  DRV_CUDA(cudaMemcpyAsync(half.data(), hdr_rgba16f, half.size() * 2, cudaMemcpyDeviceToHost, ctx->stream));
  DRV_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<float> rgba(half.size());
  for (size_t i = 0; i < half.size(); ++i) rgba[i] = __half2float(__ushort_as_half(half[i]));
  if (drv_write_pfm(path, rgba.data(), W, H) != DRV_OK) return ctx->fail(DRV_ERR_INVALID, std::string("drv_save_to_pfm: cannot write ") + path);
  return DRV_OK;
}
