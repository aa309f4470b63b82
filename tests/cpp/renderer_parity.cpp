// renderer_parity.cpp — C++ test of the host-side mirror `drv::Renderer` (include/drv_renderer.hpp), written the way
// a test against the reference's `class Renderer` (rendering/renderer.hpp:36-216) would read: construct a Renderer
// over a Scene, call the reference's setters, hand in the rasterised inputs, Draw(camera, detach, dt), read back.
//
//   renderer_parity --host   no GPU needed: constructor defaults, setter semantics, the voxel adaption carry and the
//                            uniform blocks the mirror packs (run by `pytest -m "not gpu"`)
//   renderer_parity          on a B200: whole frames through the mirror against the CPU oracle (oracle/oracle.h,
//                            TEST INFRASTRUCTURE — only tests may link it) with the north-star gate
//                            |a-b| <= 1e-5 + 1e-3 max(|a|,|b|); allocation bit-exact (run by `pytest -m gpu`)
// Inputs come from scenes/ (procedural, deterministic). Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/drv_renderer.hpp"
#include "../../oracle/oracle.h"
#include "../../scenes/scenes.h"

namespace {

int g_failures = 0;
#define EXPECT(cond, ...)                                  \
  do {                                                     \
    if (!(cond)) {                                         \
      ++g_failures;                                        \
      std::printf("FAIL %s:%d  %s  ", __FILE__, __LINE__, #cond); \
      std::printf(__VA_ARGS__);                            \
      std::printf("\n");                                   \
    }                                                      \
  } while (0)

struct Gate {
  double worst = 0.0;  // max |err| / tol
  void add(double a, double b, double rtol, double atol) {
    const double err = std::fabs(a - b), tol = atol + rtol * std::fmax(std::fabs(a), std::fabs(b));
    if (err / tol > worst || std::isnan(err)) worst = std::isnan(err) ? 1e30 : err / tol;
  }
};

template <typename T>
struct DeviceArray {
  T* ptr = nullptr;
  size_t n = 0;
  explicit DeviceArray(const std::vector<T>& host) : n(host.size()) {
    if (cudaMalloc(reinterpret_cast<void**>(&ptr), n * sizeof(T)) != cudaSuccess) { std::printf("cudaMalloc failed\n"); std::exit(2); }
    cudaMemcpy(ptr, host.data(), n * sizeof(T), cudaMemcpyHostToDevice);
  }
  explicit DeviceArray(size_t count) : n(count) {
    if (cudaMalloc(reinterpret_cast<void**>(&ptr), n * sizeof(T)) != cudaSuccess) { std::printf("cudaMalloc failed\n"); std::exit(2); }
    cudaMemset(ptr, 0, n * sizeof(T));
  }
  ~DeviceArray() { cudaFree(ptr); }
  std::vector<T> download() const {
    std::vector<T> h(n);
    cudaMemcpy(h.data(), ptr, n * sizeof(T), cudaMemcpyDeviceToHost);
    return h;
  }
  DeviceArray(const DeviceArray&) = delete;
  DeviceArray& operator=(const DeviceArray&) = delete;
};

// ------------------------------------------------------------------------------------------------------------
// The Cornell workload of BASELINE configs[0]: camera and light of application.cpp:51-52, 86-91.
struct Cornell {
  unsigned width = 512, height = 512;
  unsigned rsmResolution = 64, rsmReadLod = 0;
  unsigned cavResolution = 32;
  float cascadeSize = 16.0f;
  unsigned voxelResolution = 64;
  unsigned maxCaches = 16384;
  drv::Camera camera;
  std::shared_ptr<drv::Scene> scene = std::make_shared<drv::Scene>();
  std::vector<float> triangles;

  scn_scene* geo = nullptr;
  Cornell() {
    geo = scn_create("cornell", 1.0f);
    float mn[3], mx[3];
    scn_bounding_box(geo, mn, mx);
    scene->SetBoundingBox(mn, mx);
    const uint32_t nt = scn_triangles(geo, nullptr, 0);
    triangles.resize((size_t)nt * 9);
    scn_triangles(geo, triangles.data(), nt);
    camera.aspectRatio = (float)width / (float)height;
    drv::Light l;
    l.intensity = drv::Vec3(100.0f, 100.0f, 100.0f);
    l.position = drv::Vec3(0.0f, 1.7f, 3.3f);
    l.direction = drv::Vec3(0.0f, 0.0f, -1.0f);
    l.halfAngle = 30.0f * 3.14159265358979f / 180.0f;
    l.rsmResolution = rsmResolution;
    l.rsmReadLod = rsmReadLod;
    l.indirectShadowComputationLod = 2;
    scene->GetLights().push_back(l);
  }
  ~Cornell() { scn_destroy(geo); }
};

void configure(drv::Renderer& r, const Cornell& wl, bool sh2, bool shadow) {
  r.SetCAVCascades(1, wl.cavResolution);
  r.SetCAVCascadeWorldSize(0, wl.cascadeSize);
  r.SetCAVCascadeTransitionSize(0.0f);
  r.SetIndirectDiffuseMode(sh2 ? drv::Renderer::IndirectDiffuseMode::SH2 : drv::Renderer::IndirectDiffuseMode::SH1);
  r.SetIndirectShadow(shadow);
  r.SetVoxelVolumeResultion(wl.voxelResolution);
  r.SetVoxelVolumeAdaptionRate(1.0f);  // dt = 1 s: floor(1 * 1 * 255) / 255 = 1 -> a converged volume in one frame
  r.SetMaxCacheCount(wl.maxCaches);
}

// ------------------------------------------------------------------------------------------------------------
int host_selftest() {
  Cornell wl;
  drv::Renderer r(wl.scene, 1920, 1080);
  // constructor defaults, renderer.cpp:36-51, 84-90
  EXPECT(r.GetMode() == drv::Renderer::Mode::DYN_RADIANCE_VOLUME, "mode");
  EXPECT(r.GetIndirectDiffuseMode() == drv::Renderer::IndirectDiffuseMode::SH1, "diffuse mode");
  EXPECT(r.GetIndirectShadow() && !r.GetIndirectSpecular(), "shadow / specular defaults");
  EXPECT(r.GetMaxCacheCount() == 16384u, "%u", r.GetMaxCacheCount());
  EXPECT(r.GetCAVCascadeCount() == 3u && r.GetCAVResolution() == 32u, "cascades");
  EXPECT(r.GetCAVCascadeWorldSize(0) == 4.0f && r.GetCAVCascadeWorldSize(1) == 8.0f && r.GetCAVCascadeWorldSize(2) == 16.0f, "sizes");
  EXPECT(std::isnan(r.GetCAVCascadeWorldSize(3)), "out-of-range cascade is NaN (renderer.hpp:127)");
  EXPECT(r.GetCAVCascadeTransitionSize() == 2.0f, "transition");
  EXPECT(r.GetVoxelVolumeResultion() == 128u && r.GetVoxelVolumeAdaptionRate() == 10.0f, "voxelisation defaults");
  EXPECT(r.GetPerCacheSpecularEnvMapSize() == 16u && r.GetSpecularEnvMapHoleFillLevel() == 0u && r.GetSpecularEnvMapDirectWrite(), "specular defaults");
  EXPECT(r.GetExposure() == 1.0f && r.GetTonemapLMax() == 1.2f, "tonemap defaults");
  EXPECT(!r.GetReadLightCacheCount() && r.GetLightCacheActiveCount() == 0u, "cache count tracking");

  // SetCAVCascades keeps existing sizes and doubles into new cascades (renderer.cpp:1174-1193)
  r.SetCAVCascadeWorldSize(1, 10.0f);
  r.SetCAVCascades(4, 64);
  EXPECT(r.GetCAVCascadeWorldSize(1) == 10.0f && r.GetCAVCascadeWorldSize(2) == 16.0f && r.GetCAVCascadeWorldSize(3) == 32.0f, "grow");
  r.SetCAVCascades(2, 64);
  r.SetCAVCascades(3, 64);
  EXPECT(r.GetCAVCascadeWorldSize(2) == 20.0f, "a re-added cascade doubles its predecessor: %f", r.GetCAVCascadeWorldSize(2));
  r.SetCAVCascades(5, 32);
  EXPECT(r.GetLastStatus() == DRV_ERR_INVALID && r.GetCAVCascadeCount() == 3u, "more than s_maxNumCAVCascades is refused");

  // hole-fill level is clamped to log2(per-cache size), both ways round (renderer.hpp:99, renderer.cpp:458)
  r.SetSpecularEnvMapHoleFillLevel(9);
  EXPECT(r.GetSpecularEnvMapHoleFillLevel() == 4u, "%u", r.GetSpecularEnvMapHoleFillLevel());
  r.SetPerCacheSpecularEnvMapSize(8);
  EXPECT(r.GetSpecularEnvMapHoleFillLevel() == 3u && r.GetPerCacheSpecularEnvMapSize() == 8u, "clamp on resize");

  // voxel adaption: floor(dt * rate * 255) / 255 with the remainder carried (voxelization.cpp:90-100)
  r.SetVoxelVolumeAdaptionRate(10.0f);
  EXPECT(r.ConsumeVoxelAdaption(0.0001f) == 0.0f, "0.255 of a step: nothing this frame");
  EXPECT(std::fabs(r.ConsumeVoxelAdaption(0.0003f) - 1.0f / 255.0f) < 1e-9f, "0.255 + 0.765 = 1.02 -> one step");
  EXPECT(r.ConsumeVoxelAdaption(0.00038f) == 0.0f, "0.02 + 0.969 = 0.989 -> none");
  EXPECT(r.ConsumeVoxelAdaption(1.0f) == 1.0f, "a long frame saturates at 255 / 255");

  // the Constant block the mirror packs (renderer.cpp:290-322), against the formulas in double
  r.UpdateConstantUBO();
  const drv_constant& c = r.GetConstantBlock();
  const double pi = 3.14159265358979323846;
  EXPECT(std::fabs(c.ShCosLobeFactor0 - std::sqrt(pi) / 2.0) < 1e-6, "g0");
  EXPECT(std::fabs(c.ShCosLobeFactor1 - std::sqrt(pi / 3.0)) < 1e-6, "g1");
  EXPECT(std::fabs(c.ShEvaFactor0 - 1.0 / (2.0 * std::sqrt(pi))) < 1e-6, "f0");
  EXPECT(std::fabs(c.ShEvaFactor1 - std::sqrt(3.0) / (2.0 * std::sqrt(pi))) < 1e-6, "f1");
  EXPECT(c.BackbufferResolution[0] == 1920 && c.BackbufferResolution[1] == 1080, "resolution");
  EXPECT(c.AddressVolumeResolution == 64 && c.NumAddressVolumeCascades == 3 && c.VoxelResolution == 128, "volume fields");
  EXPECT(c.MaxNumLightCaches == 16384u, "max caches");
  // the header packers of include/drv_math.h (what a C++ host may call directly) and their C entry points agree byte for byte
  {
    drv::Camera cam;
    cam.aspectRatio = 1920.0f / 1080.0f;
    drv_per_frame a, b;
    std::memset(&a, 0, sizeof(a)); std::memset(&b, 0, sizeof(b));
    drv::packPerFrame(&a, cam, 1.5f);
    const drv_camera_desc cd = drv::ToDesc(cam);
    drv_pack_per_frame(&b, &cd, 1.5f);
    EXPECT(std::memcmp(&a, &b, sizeof(a)) == 0, "packPerFrame != drv_pack_per_frame");
    drv_spot_light la, lb;
    std::memset(&la, 0, sizeof(la)); std::memset(&lb, 0, sizeof(lb));
    const drv::Light& light = wl.scene->GetLights()[0];
    drv::packSpotLight(&la, light);
    const drv_light_desc ld = drv::ToDesc(light);
    drv_pack_spot_light(&lb, &ld);
    EXPECT(std::memcmp(&la, &lb, sizeof(la)) == 0, "packSpotLight != drv_pack_spot_light");
    EXPECT(la.RSMReadResolution == 64 && la.IndirectShadowComputationSampleInterval == 16, "light block fields");
  }
  static_assert(sizeof(drv_constant) == 80 && sizeof(drv_per_frame) == 288 && sizeof(drv_volume_info) == 288 &&
                    sizeof(drv_spot_light) == 224, "std140 block sizes (SURVEY A.1)");
  std::printf(g_failures ? "HOST FAILED (%d)\n" : "HOST OK\n", g_failures);
  return g_failures ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------------------
// One Cornell frame through drv::Renderer against the oracle run stage by stage on the same inputs.
void frame_parity(bool sh2, bool shadow) {
  std::printf("== Cornell 512x512, 4096 VPLs, 1 x 32^3, %s, %s\n", sh2 ? "SH2" : "SH1", shadow ? "cone-traced shadows" : "unshadowed");
  Cornell wl;
  const int threads = orc_default_threads();
  drv::Renderer r(wl.scene, wl.width, wl.height);
  configure(r, wl, sh2, shadow);
  EXPECT(r.GetLastStatus() == DRV_OK, "%s", r.GetLastError().c_str());

  // uniform blocks first: the scene is "rasterised" (ray-cast by scenes/) with the mirror's own blocks
  r.UpdatePerFrameUBO(wl.camera);
  r.UpdateVolumeUBO(wl.camera);
  r.PrepareLights();
  EXPECT(r.GetLastStatus() == DRV_OK, "%s", r.GetLastError().c_str());
  if (r.GetLastStatus() != DRV_OK) return;
  const drv_constant cb = r.GetConstantBlock();
  const drv_per_frame pf = r.GetPerFrameBlock();
  const drv_volume_info vi = r.GetVolumeInfoBlock();
  const drv_spot_light sl = r.GetSpotLightBlocks()[0];

  const size_t px = (size_t)wl.width * wl.height;
  std::vector<float> depth(px);
  std::vector<int16_t> normal(px * 2);
  std::vector<uint8_t> diffuse(px * 4);
  scn_render_gbuffer(wl.geo, &pf, wl.width, wl.height, depth.data(), normal.data(), diffuse.data(), threads);
  const size_t rt = (size_t)wl.rsmResolution * wl.rsmResolution;
  std::vector<uint16_t> flux(rt * 4), rdepth(rt * 2);
  std::vector<int16_t> rnormal(rt * 2);
  scn_render_rsm(wl.geo, &sl, flux.data(), rnormal.data(), rdepth.data(), threads);

  DeviceArray<float> d_depth(depth);
  DeviceArray<int16_t> d_normal(normal);
  DeviceArray<uint8_t> d_diffuse(diffuse);
  DeviceArray<uint16_t> d_flux(flux), d_rdepth(rdepth);
  DeviceArray<int16_t> d_rnormal(rnormal);
  DeviceArray<float> d_tris(wl.triangles);
  drv::SceneEntity ent;
  ent.devicePositions = d_tris.ptr;
  ent.numTriangles = (uint32_t)(wl.triangles.size() / 9);
  wl.scene->GetEntities().assign(1, ent);

  r.BindGBuffer(d_depth.ptr, d_normal.ptr, d_diffuse.ptr);
  r.BindShadowMap(0, d_flux.ptr, d_rnormal.ptr, d_rdepth.ptr);
  r.SetReadLightCacheCount(true);
  r.Draw(wl.camera, false, 1.0f);
  r.Finish();
  EXPECT(r.GetLastStatus() == DRV_OK, "Draw: %s", r.GetLastError().c_str());
  if (r.GetLastStatus() != DRV_OK) return;
  EXPECT(drv_kernel_launches(r.Context()) > 0, "no kernel was launched");

  // ---- oracle, the reference's stage order (Renderer::Draw, renderer.cpp:539-594)
  const uint32_t stride = sh2 ? 128u : 64u;
  std::vector<drv_vpl> vpls(rt);
  orc_generate_vpls(&sl, flux.data(), rnormal.data(), rdepth.data(), vpls.data());  // read LOD 0: level 0 is the read level
  std::vector<drv_shadow_block> blocks;
  std::vector<uint8_t> chain;
  if (shadow) {
    // depthLinSq at mip IndirectShadowComputationLod (downsamplersm.frag), then the block records
    std::vector<uint16_t> f0 = flux, d0 = rdepth;
    std::vector<int16_t> n0 = rnormal;
    uint32_t res = wl.rsmResolution;
    for (unsigned l = 0; l < (unsigned)sl.IndirectShadowComputationLod; ++l) {
      const size_t h = (size_t)(res / 2) * (res / 2);
      std::vector<uint16_t> f1(h * 4), d1(h * 2);
      std::vector<int16_t> n1(h * 2);
      orc_rsm_downsample(f0.data(), n0.data(), d0.data(), res, f1.data(), n1.data(), d1.data());
      f0.swap(f1); n0.swap(n1); d0.swap(d1);
      res /= 2;
    }
    blocks.resize(rt / (size_t)sl.IndirectShadowComputationSampleInterval);
    orc_shadow_blocks(&sl, d0.data(), blocks.data());
    // VoxelizeScene: clear, rasterise, blend with adaption 1, mips
    const uint32_t vr = wl.voxelResolution;
    const size_t v0 = (size_t)vr * vr * vr;
    size_t chain_bytes = 0;
    for (uint32_t s = vr; s >= 1; s /= 2) chain_bytes += (size_t)s * s * s;
    std::vector<uint8_t> target(v0, 0);
    const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    orc_voxelize(&vi, vr, wl.triangles.data(), (uint32_t)(wl.triangles.size() / 9), identity, target.data());
    chain.assign(chain_bytes, 0);
    orc_voxel_blend(chain.data(), target.data(), vr, 1.0f);
    orc_voxel_mips(chain.data(), vr);
    // the GPU volume is bit-exact
    drv_buffers b;
    drv_get_buffers(r.Context(), &b);
    EXPECT(b.voxel_chain_bytes == chain_bytes, "%llu vs %zu", (unsigned long long)b.voxel_chain_bytes, chain_bytes);
    std::vector<uint8_t> gpu_chain(chain_bytes);
    cudaMemcpy(gpu_chain.data(), b.voxel_chain, chain_bytes, cudaMemcpyDeviceToHost);
    EXPECT(gpu_chain == chain, "voxel chain differs from the oracle");
  }
  const uint32_t R = (uint32_t)cb.AddressVolumeResolution, C = (uint32_t)cb.NumAddressVolumeCascades;
  std::vector<uint32_t> atlas((size_t)R * R * R * C, 0);
  std::vector<uint8_t> entries((size_t)wl.maxCaches * stride, 0);
  drv_cache_counter counter;
  uint32_t overflow = 0, oob = 0;
  const int count = orc_allocate_caches(&cb, &pf, &vi, 0, depth.data(), atlas.data(), entries.data(), stride, wl.maxCaches,
                                        &counter, &overflow, &oob, threads);
  const drv_vpl* vl[1] = {vpls.data()};
  const drv_shadow_block* bl[1] = {shadow ? blocks.data() : nullptr};
  orc_light_caches(&cb, &vi, &sl, 1, vl, bl, shadow ? chain.data() : nullptr, entries.data(), stride, 0, (uint32_t)count,
                   sh2 ? 2 : 1, shadow ? 1 : 0, 0, threads);
  std::vector<float> image(px * 4, 0.0f);
  orc_apply_caches(&cb, &pf, &vi, 0, sh2 ? 2 : 1, depth.data(), normal.data(), diffuse.data(), atlas.data(), entries.data(),
                   stride, wl.maxCaches, image.data(), threads);
  EXPECT(count > 100 && overflow == 0, "oracle allocated %d caches", count);

  // ---- cache entries: count, positions bit-exact, SH within the gate
  uint32_t n = 0, ov = 0;
  EXPECT(drv_active_cache_count(r.Context(), &n, &ov, nullptr) == DRV_OK, "%s", drv_last_error(r.Context()));
  EXPECT((int)n == count && ov == 0, "GPU %u caches, oracle %d", n, count);
  drv_buffers b;
  drv_get_buffers(r.Context(), &b);
  EXPECT(b.entry_stride == stride, "stride %u", b.entry_stride);
  std::vector<uint8_t> gpu_entries((size_t)n * stride);
  cudaMemcpy(gpu_entries.data(), b.entries, gpu_entries.size(), cudaMemcpyDeviceToHost);
  Gate sh;
  size_t bad_pos = 0;
  const uint32_t words = stride / 4;
  for (uint32_t i = 0; i < n && (int)i < count; ++i) {
    const float* g = reinterpret_cast<const float*>(gpu_entries.data() + (size_t)i * stride);
    const float* o = reinterpret_cast<const float*>(entries.data() + (size_t)i * stride);
    if (std::memcmp(g, o, 16) != 0) ++bad_pos;
    for (uint32_t w = 4; w < words; ++w) sh.add(g[w], o[w], 1e-3, 1e-5);
  }
  EXPECT(bad_pos == 0, "%zu entry positions differ (allocation must be bit-exact)", bad_pos);
  EXPECT(sh.worst <= 1.0, "SH: worst |err|/tol = %.3f", sh.worst);

  // ---- radiance: the apply pass once more into a float4 image (tight gate), and Draw's RGBA16F back buffer
  DeviceArray<float> d_out32(px * 4);
  r.ApplyCaches(d_out32.ptr, DRV_HDR_RGBA32F_WRITE);
  r.Finish();
  const std::vector<float> out32 = d_out32.download();
  Gate rad;
  float peak = 0.0f;
  size_t alpha_bad = 0;
  for (size_t i = 0; i < px; ++i) {
    for (int c = 0; c < 3; ++c) { rad.add(out32[i * 4 + c], image[i * 4 + c], 1e-3, 1e-5); peak = std::fmax(peak, image[i * 4 + c]); }
    if (out32[i * 4 + 3] != image[i * 4 + 3]) ++alpha_bad;
  }
  EXPECT(alpha_bad == 0, "%zu pixels: discarded / shaded disagree", alpha_bad);
  EXPECT(rad.worst <= 1.0, "radiance: worst |err|/tol = %.3f", rad.worst);
  EXPECT(peak > 1e-3f, "the frame is not dark (%g)", peak);
  std::vector<uint16_t> hdr(px * 4);
  cudaMemcpy(hdr.data(), r.HDRBackbuffer(), hdr.size() * 2, cudaMemcpyDeviceToHost);
  Gate rad16;  // RGBA16F target: half a unit in the last place of a half on top of the gate
  for (size_t i = 0; i < px; ++i)
    for (int c = 0; c < 3; ++c) rad16.add(orc_half_to_float(hdr[i * 4 + c]), image[i * 4 + c], 2e-3, 1e-4);
  EXPECT(rad16.worst <= 1.0, "RGBA16F back buffer: worst |err|/tol = %.3f", rad16.worst);

  // ---- GetLightCacheActiveCount is one frame late (renderer.cpp:960-966): 0 after the first Draw, then the count
  EXPECT(r.GetLightCacheActiveCount() == 0u, "%u", r.GetLightCacheActiveCount());
  r.Draw(wl.camera, false, 0.0f);
  r.Finish();
  EXPECT(r.GetLastStatus() == DRV_OK, "second Draw: %s", r.GetLastError().c_str());
  EXPECT((int)r.GetLightCacheActiveCount() == count, "%u vs %d", r.GetLightCacheActiveCount(), count);

  // ---- detachViewFromCameraUpdate: caches stay where they are, only the view changes (renderer.hpp:43-45)
  drv::Camera moved = wl.camera;
  moved.position.x += 0.25f;
  const uint64_t launches = drv_kernel_launches(r.Context());
  r.Draw(moved, true, 0.0f);
  r.Finish();
  uint32_t n2 = 0;
  drv_active_cache_count(r.Context(), &n2, nullptr, nullptr);
  EXPECT(n2 == n, "detached view re-allocated caches (%u -> %u)", n, n2);
  EXPECT(drv_kernel_launches(r.Context()) > launches, "the apply pass still runs");

  {  // the overlapped frame (one drv_draw_frame: three streams, fused clear, CUDA graph) is bit-identical to the serial order
    r.Draw(wl.camera, false, 0.0f);  // serial reference of this state (voxel volume converged, no blend this frame)
    r.Finish();
    std::vector<uint16_t> serial(px * 4), fast(px * 4);
    cudaMemcpy(serial.data(), r.HDRBackbuffer(), serial.size() * 2, cudaMemcpyDeviceToHost);
    r.SetOverlappedFrame(true);
    for (int rep = 0; rep < 3; ++rep) {  // eager, record + replay, replay
      cudaMemset(r.HDRBackbuffer(), 0x3c, px * 8);
      r.Draw(wl.camera, false, 0.0f);
      r.Finish();
      EXPECT(r.GetLastStatus() == DRV_OK, "overlapped Draw %d: %s", rep, r.GetLastError().c_str());
      cudaMemcpy(fast.data(), r.HDRBackbuffer(), fast.size() * 2, cudaMemcpyDeviceToHost);
      EXPECT(fast == serial, "overlapped frame %d differs from the serial order", rep);
    }
    uint64_t inst = 0, upd = 0;
    drv_graph_stats(r.Context(), &inst, &upd);
    // informational: with unchanged light blocks the mirror uploads them once, so the second overlapped frame records
    // the graph and the third replays it (a light that moves every frame keeps the frame eager — still overlapped)
    r.SetOverlappedFrame(false);
    std::printf("   overlapped frame == serial frame (graph instantiations %llu, updates %llu)\n", (unsigned long long)inst, (unsigned long long)upd);
  }

  if (shadow) {  // output mode AMBIENT_OCCLUSION through the same mirror (renderer.cpp:631-642)
    r.SetMode(drv::Renderer::Mode::AMBIENTOCCLUSION);
    r.Draw(wl.camera, false, 0.0f);
    r.Finish();
    EXPECT(r.GetLastStatus() == DRV_OK, "AO Draw: %s", r.GetLastError().c_str());
    std::vector<float> ao(px), ao_ref(px, 0.0f);
    if (r.AOTarget()) cudaMemcpy(ao.data(), r.AOTarget(), px * 4, cudaMemcpyDeviceToHost);
    orc_cone_trace_ao(&pf, &vi, chain.data(), wl.voxelResolution, depth.data(), normal.data(), wl.width, wl.height, ao_ref.data(), threads);
    double worst = 0.0;
    for (size_t i = 0; i < px; ++i) worst = std::fmax(worst, std::fabs((double)ao[i] - ao_ref[i]));
    EXPECT(worst <= 2e-3, "AO: worst abs error %.2e", worst);
    std::printf("   AO worst abs err %.2e\n", worst);
  }
  std::printf("   %u caches, worst |err|/tol: SH %.3f  radiance %.3f  RGBA16F back buffer %.3f\n", n, sh.worst, rad.worst, rad16.worst);
}

}  // namespace

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--host") return host_selftest();
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    std::printf("no CUDA device: the product path has no CPU fallback\n");
    return 3;
  }
  std::printf("%s\n", drv_version());
  frame_parity(false, false);
  frame_parity(true, true);
  std::printf(g_failures ? "PARITY FAILED (%d)\n" : "PARITY OK\n", g_failures);
  return g_failures ? 1 : 0;
}
