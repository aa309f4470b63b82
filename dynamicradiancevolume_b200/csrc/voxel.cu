// voxel.cu — stage 3: scene voxelisation, temporal blend and the mip chain
// (≙ Voxelization::VoxelizeScene, rendering/voxelization.cpp:90-176;
// shader/voxelize.vert:15-23, voxelize.geom:19-112, voxelize.frag:21-58,
// voxelblend.comp:8-19, voxelmipmap.comp:8-13).
//
// The reference leans on the hardware rasteriser (one draw per entity into a
// res x res viewport, conservative dilation in the geometry shader, image
// stores in the fragment shader). Here one warp owns one triangle: lane 0..31
// set up the same dilated edge planes, then the warp sweeps the triangle's
// dilated bounding box 32 pixels at a time and stores the covered voxels.
// All coverage / depth maths is DECISION maths (device_math.cuh) so the voxel
// set equals the oracle's bit for bit.
//
// Layout: the R8 volume and all its mips live in ONE contiguous buffer
// (level 0, then level 1, ...; x fastest). At 128^3 the whole chain is
// 2,396,745 bytes — a sliver of B200's 126 MB L2, where it stays resident for
// the cone tracer (stage 4) once touched.
#include "ctx.h"
#include "device_math.cuh"

using namespace drvk;

namespace {

struct VoxParams {
  float vmin[3], vmax[3];
  float world[16];
  int res;
};

struct Plane { float x, y, z; };
__device__ __forceinline__ Plane cross_h(float ax, float ay, float az, float bx, float by, float bz) {
  Plane p = {ex_sub(ex_mul(ay, bz), ex_mul(az, by)), ex_sub(ex_mul(az, bx), ex_mul(ax, bz)),
             ex_sub(ex_mul(ax, by), ex_mul(ay, bx))};
  return p;
}

__device__ __forceinline__ void set_voxel(uint8_t* __restrict__ vol, int res, int side, int px, int py, int pz) {
  int x, y, z; // UnswizzlePos, voxelize.frag:17-20
  if (side == 0) { x = pz; y = py; z = px; }
  else if (side == 1) { x = px; y = pz; z = py; }
  else { x = px; y = py; z = pz; }
  if (x < 0 || y < 0 || z < 0 || x >= res || y >= res || z >= res) return;
  vol[(size_t)x + (size_t)res * ((size_t)y + (size_t)res * z)] = 255;
}

// Per-triangle raster set-up (voxelize.vert + voxelize.geom), written to shared memory by one lane.
struct TriSetup {
  Plane pl[3];
  float aabb[4];
  float vx0, vy0, vz0, gx, gy, maxChange;
  int side, x0, y0, w, total;
};

constexpr int kTrisPerBlock = 8;

__device__ __forceinline__ void setup_triangle(const VoxParams& P, const float* __restrict__ tp, TriSetup& T) {
  const int res = P.res;
  const float fres = (float)res;
  T.total = 0;
  float cx[3], cy[3], cz[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { // voxelize.vert:20-22
    float vx = __ldg(tp + i * 3), vy = __ldg(tp + i * 3 + 1), vz = __ldg(tp + i * 3 + 2);
    float wx = ex_dot4(P.world + 0, vx, vy, vz, 1.0f);
    float wy = ex_dot4(P.world + 4, vx, vy, vz, 1.0f);
    float wz = ex_dot4(P.world + 8, vx, vy, vz, 1.0f);
    cx[i] = ex_sub(ex_mul(ex_div(ex_sub(wx, P.vmin[0]), ex_sub(P.vmax[0], P.vmin[0])), 2.0f), 1.0f);
    cy[i] = ex_sub(ex_mul(ex_div(ex_sub(wy, P.vmin[1]), ex_sub(P.vmax[1], P.vmin[1])), 2.0f), 1.0f);
    cz[i] = ex_sub(ex_mul(ex_div(ex_sub(wz, P.vmin[2]), ex_sub(P.vmax[2], P.vmin[2])), 2.0f), 1.0f);
  }
  // voxelize.geom:21-26 — dominant axis of the face normal
  Plane nr = cross_h(ex_sub(cx[1], cx[0]), ex_sub(cy[1], cy[0]), ex_sub(cz[1], cz[0]), ex_sub(cx[2], cx[0]),
                     ex_sub(cy[2], cy[0]), ex_sub(cz[2], cz[0]));
  float inv = ex_rsqrt(ex_dot3(nr.x, nr.y, nr.z, nr.x, nr.y, nr.z));
  float an[3] = {fabsf(ex_mul(nr.x, inv)), fabsf(ex_mul(nr.y, inv)), fabsf(ex_mul(nr.z, inv))};
  int side = an[0] > an[1] ? 0 : 1;
  side = (side == 0 ? an[0] : an[1]) > an[2] ? side : 2;
  T.side = side;
  float rx[3], ry[3], rz[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { // geom:31-53
    if (side == 0) { rx[i] = cz[i]; ry[i] = cy[i]; rz[i] = cx[i]; }
    else if (side == 1) { rx[i] = cx[i]; ry[i] = cz[i]; rz[i] = cy[i]; }
    else { rx[i] = cx[i]; ry[i] = cy[i]; rz[i] = cz[i]; }
  }
  const float h = ex_div(1.0f, fres); // geom:56
  float aabb[4] = {ex_sub(fminf(fminf(rx[0], rx[1]), rx[2]), h), ex_sub(fminf(fminf(ry[0], ry[1]), ry[2]), h),
                   ex_add(fmaxf(fmaxf(rx[0], rx[1]), rx[2]), h), ex_add(fmaxf(fmaxf(ry[0], ry[1]), ry[2]), h)};
#pragma unroll
  for (int i = 0; i < 4; ++i) T.aabb[i] = aabb[i] = ex_mul(ex_add(ex_mul(aabb[i], 0.5f), 0.5f), fres); // geom:61
  float ax = ex_sub(rx[0], rx[2]), ay = ex_sub(ry[0], ry[2]);
  float bx = ex_sub(rx[1], rx[0]), by = ex_sub(ry[1], ry[0]);
  Plane pl[3];
  pl[0] = cross_h(ax, ay, 0.0f, rx[2], ry[2], 1.0f); // geom:64-69
  pl[1] = cross_h(bx, by, 0.0f, rx[0], ry[0], 1.0f);
  pl[2] = cross_h(ex_sub(rx[2], rx[1]), ex_sub(ry[2], ry[1]), 0.0f, rx[1], ry[1], 1.0f);
  float wnd = ex_sub(ex_mul(ax, by), ex_mul(bx, ay)); // geom:72
  float winding = wnd > 0.0f ? 1.0f : (wnd < 0.0f ? -1.0f : 0.0f);
  if (winding == 0.0f) return;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    pl[i].x = ex_mul(pl[i].x, winding); pl[i].y = ex_mul(pl[i].y, winding); pl[i].z = ex_mul(pl[i].z, winding);
    pl[i].z = ex_sub(pl[i].z, ex_add(ex_mul(h, fabsf(pl[i].x)), ex_mul(h, fabsf(pl[i].y)))); // geom:78-80
    T.pl[i] = pl[i];
  }
  float vx[3], vy[3], vz[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { // geom:96-106
    const Plane& p0 = pl[i];
    const Plane& p1 = pl[(i + 1) % 3];
    Plane c = cross_h(p0.x, p0.y, p0.z, p1.x, p1.y, p1.z);
    vx[i] = ex_mul(ex_add(ex_mul(ex_div(c.x, c.z), 0.5f), 0.5f), fres);
    vy[i] = ex_mul(ex_add(ex_mul(ex_div(c.y, c.z), 0.5f), 0.5f), fres);
    vz[i] = ex_mul(ex_add(ex_mul(rz[i], 0.5f), 0.5f), fres);
  }
  float e1x = ex_sub(vx[1], vx[0]), e1y = ex_sub(vy[1], vy[0]), e1z = ex_sub(vz[1], vz[0]);
  float e2x = ex_sub(vx[2], vx[0]), e2y = ex_sub(vy[2], vy[0]), e2z = ex_sub(vz[2], vz[0]);
  float det = ex_sub(ex_mul(e1x, e2y), ex_mul(e2x, e1y));
  if (det == 0.0f || det != det) return;
  T.gx = ex_div(ex_sub(ex_mul(e1z, e2y), ex_mul(e2z, e1y)), det); // dFdx, frag:39
  T.gy = ex_div(ex_sub(ex_mul(e1x, e2z), ex_mul(e2x, e1z)), det); // dFdy, frag:40
  T.maxChange = ex_mul(ex_sqrt(ex_add(ex_mul(T.gx, T.gx), ex_mul(T.gy, T.gy))), 1.414f); // frag:41
  T.vx0 = vx[0]; T.vy0 = vy[0]; T.vz0 = vz[0];
  int x0 = max(0, ex_trunc(floorf(ex_sub(aabb[0], 0.5f))));
  int y0 = max(0, ex_trunc(floorf(ex_sub(aabb[1], 0.5f))));
  int x1 = min(res - 1, ex_trunc(floorf(aabb[2])));
  int y1 = min(res - 1, ex_trunc(floorf(aabb[3])));
  if (x1 < x0 || y1 < y0) return;
  T.x0 = x0; T.y0 = y0; T.w = x1 - x0 + 1;
  T.total = T.w * (y1 - y0 + 1);
}

// A block rasterises kTrisPerBlock triangles: eight lanes set them up, then all 256 threads sweep the
// concatenated bounding-box pixel lists, so one wall-sized triangle does not serialise on a single warp.
__global__ void __launch_bounds__(256) voxelize_kernel(VoxParams P, const float* __restrict__ tris, uint32_t num_tris,
                                                       uint8_t* __restrict__ vol) {
  __shared__ TriSetup s_tri[kTrisPerBlock];
  __shared__ int s_start[kTrisPerBlock + 1];
  const uint32_t tri0 = blockIdx.x * kTrisPerBlock;
  if (threadIdx.x < kTrisPerBlock) {
    s_tri[threadIdx.x].total = 0;
    if (tri0 + threadIdx.x < num_tris) setup_triangle(P, tris + (size_t)(tri0 + threadIdx.x) * 9, s_tri[threadIdx.x]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < kTrisPerBlock; ++i) { s_start[i] = acc; acc += s_tri[i].total; }
    s_start[kTrisPerBlock] = acc;
  }
  __syncthreads();
  const int res = P.res;
  const float fres = (float)res;
  const int total = s_start[kTrisPerBlock];
  // gridDim.y blocks share one triangle group: block y takes every gridDim.y-th stripe of 256 pixels
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.y) {
    int t = 0;
#pragma unroll
    for (int k = 1; k < kTrisPerBlock; ++k) t += (i >= s_start[k]) ? 1 : 0;
    const TriSetup& T = s_tri[t];
    const int j = i - s_start[t];
    int py = T.y0 + j / T.w, px = T.x0 + j % T.w;
    float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
    if (fx < T.aabb[0] || fy < T.aabb[1] || fx > T.aabb[2] || fy > T.aabb[3]) continue; // frag:24-27
    float ccx = ex_sub(ex_mul(ex_div(fx, fres), 2.0f), 1.0f), ccy = ex_sub(ex_mul(ex_div(fy, fres), 2.0f), 1.0f);
    bool inside = true;
#pragma unroll
    for (int e = 0; e < 3; ++e)
      if (ex_add(ex_add(ex_mul(T.pl[e].x, ccx), ex_mul(T.pl[e].y, ccy)), T.pl[e].z) > 0.0f) inside = false;
    if (!inside) continue;
    float zv = ex_add(T.vz0, ex_add(ex_mul(T.gx, ex_sub(fx, T.vx0)), ex_mul(T.gy, ex_sub(fy, T.vy0)))); // frag:33
    if (zv < 0.0f || zv > fres) continue;
    int zi = ex_trunc(zv);                                                                          // frag:34
    set_voxel(vol, res, T.side, px, py, zi);
    if (zi != ex_trunc(ex_sub(zv, T.maxChange))) set_voxel(vol, res, T.side, px, py, zi - 1);       // frag:46-51
    if (zi != ex_trunc(ex_add(zv, T.maxChange))) set_voxel(vol, res, T.side, px, py, zi + 1);       // frag:52-57
  }
}

// voxelblend.comp:16 on UNORM8, 16 voxels per thread.
__global__ void voxel_blend_kernel(uint4* __restrict__ vol, const uint4* __restrict__ target, size_t n16, int k) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n16) return;
  uint4 o = vol[i], t = __ldg(target + i);
  uint32_t ow[4] = {o.x, o.y, o.z, o.w}, tw[4] = {t.x, t.y, t.z, t.w}, r[4];
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    uint32_t out = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int ov = (ow[w] >> (8 * b)) & 0xff, tv = (tw[w] >> (8 * b)) & 0xff;
      int s = (tv > ov) - (tv < ov);
      out |= (uint32_t)clampi(ov + s * k, 0, 255) << (8 * b);
    }
    r[w] = out;
  }
  vol[i] = make_uint4(r[0], r[1], r[2], r[3]);
}

// voxelmipmap.comp:11-12, several levels per launch (the reference dispatches one pass per level,
// voxelization.cpp:161-171). A block owns a T^3 tile of the source level (T = min(16, source resolution))
// and reduces it through shared memory: 16^3 -> 8^3 -> 4^3 -> 2^3 -> 1, storing each level. Every level is
// rounded to UNORM8 ((sum + 4) >> 3) before it feeds the next one, exactly as separate passes would.
struct VoxMipArgs {
  const uint8_t* src;
  uint8_t* dst[4];
  int src_res;
  int levels; // <= 4
};

__global__ void __launch_bounds__(256) voxel_mip_chain_kernel(VoxMipArgs A) {
  __shared__ __align__(16) uint8_t s_a[16 * 16 * 16];
  __shared__ uint8_t s_b[8 * 8 * 8];
  const int T = min(A.src_res, 16);
  const int bx = blockIdx.x * T, by = blockIdx.y * T, bz = blockIdx.z * T;
  // stage the tile (rows of T bytes; T is 16 or a smaller power of two)
  if (T == 16) { // a row of the tile is one aligned 16-byte word (every level starts 16-byte aligned: res^3 bytes each)
    const int y = threadIdx.x & 15, z = threadIdx.x >> 4; // 256 threads = 16 x 16 rows
    reinterpret_cast<uint4*>(s_a)[threadIdx.x] = *reinterpret_cast<const uint4*>(
        A.src + (size_t)bx + (size_t)A.src_res * ((size_t)(by + y) + (size_t)A.src_res * (bz + z)));
  } else {
    for (int i = threadIdx.x; i < T * T * T; i += blockDim.x) {
      const int x = i % T, y = (i / T) % T, z = i / (T * T);
      s_a[i] = A.src[(size_t)(bx + x) + (size_t)A.src_res * ((size_t)(by + y) + (size_t)A.src_res * (bz + z))];
    }
  }
  __syncthreads();
  uint8_t* cur = s_a;
  uint8_t* nxt = s_b;
  int span = T, res = A.src_res;
  for (int l = 0; l < A.levels; ++l) {
    const int hs = span >> 1, h = res >> 1;
    const int ox = bx >> (l + 1), oy = by >> (l + 1), oz = bz >> (l + 1);
    for (int i = threadIdx.x; i < hs * hs * hs; i += blockDim.x) {
      const int x = i % hs, y = (i / hs) % hs, z = i / (hs * hs);
      int sum = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        sum += cur[(2 * x + (k & 1)) + span * ((2 * y + ((k >> 1) & 1)) + span * (2 * z + (k >> 2)))];
      const uint8_t v = (uint8_t)((sum + 4) >> 3);
      nxt[i] = v;
      A.dst[l][(size_t)(ox + x) + (size_t)h * ((size_t)(oy + y) + (size_t)h * (oz + z))] = v;
    }
    __syncthreads();
    uint8_t* t = cur; cur = nxt; nxt = t; // 8^3 fits s_b; 4^3 and below fit either buffer
    span = hs;
    res = h;
  }
}

// Gather-ready records for the cone tracer (stage 4): for every level and every lower-corner coordinate
// (x,y,z) in [-1, r-1]^3 the eight clamp-to-edge texels of the trilinear footprint, packed
// .x = a000 a100 a010 a110, .y = a001 a101 a011 a111 (low byte first). One thread per record, all levels in
// one launch; `offsets[l]` = first record of level l.
struct RecordLevels {
  uint32_t offset[16];
  uint32_t chain_offset[16];
  int levels, res;
};

// Block = 4 record rows (y) x ZB slabs (z) of one level, 64 threads per row. The record grid of a level is padded:
// D = r + 1 + 2 P records per axis (P = kVoxelRecordPad), record x <-> lower corner x - 1 - P, which reads texels
// clamp(x - 1 - P) and clamp(x - P) per axis (clamp to [0, r - 1]). Per source row one aligned 32-bit word (texels
// 4m-4..4m-1; P is a multiple of 4) and the byte before it cover the four records 4m..4m+3, and the ZB slabs share
// their ZB + 1 source planes — 2 (ZB + 1) x 2 loads, all in flight together, per 4 ZB records (the kernel is
// latency-bound: a record costs 8 dependent-free byte loads and one store, and 8 000 small blocks ran in ~60
// waves). The records pass through shared memory so that the 8-byte stores of a warp are consecutive.
// blockIdx.y = slab group summed over the levels (RecordRows), blockIdx.x * 4 + threadIdx.y = y; levels with r < 4
// take the byte path.
struct RecordRows {
  uint32_t z_offset[17]; // first blockIdx.y of level l (level l has ceil(D / ZB) of them), [levels] = total
};

template <int ZB>
__global__ void __launch_bounds__(256) voxel_records_kernel(RecordLevels L, RecordRows R, const uint8_t* __restrict__ chain,
                                                            uint2* __restrict__ records) {
  extern __shared__ uint2 s_out[]; // ZB x 4 rows of (level-0 row length + 3) records
  constexpr int P = (int)kVoxelRecordPad;
  static_assert(P % 4 == 0, "the word loads of the record build need a pad that is a multiple of 4");
  int l = 0;
#pragma unroll 1
  while (l + 1 < L.levels && blockIdx.y >= R.z_offset[l + 1]) ++l;
  const int r = L.res >> l, D = r + 1 + 2 * P;
  const int yr = (int)(blockIdx.x * blockDim.y + threadIdx.y), zr0 = (int)(blockIdx.y - R.z_offset[l]) * ZB;
  if ((int)(blockIdx.x * blockDim.y) >= D) return; // block-uniform
  const bool live = yr < D;
  const uint8_t* lvl = chain + L.chain_offset[l];
  auto clampt = [&](int t) { return min(max(t, 0), r - 1); };
  const int ys[2] = {clampt(yr - 1 - P), clampt(yr - P)};
  const int stride = L.res + 1 + 2 * P + 3;
  if (live && r < 4) {
    for (int zz = 0; zz < ZB && zr0 + zz < D; ++zz) {
      const int zr = zr0 + zz, zs[2] = {clampt(zr - 1 - P), clampt(zr - P)};
      for (int x = (int)threadIdx.x; x < D; x += blockDim.x) {
        const int x0 = clampt(x - 1 - P), x1 = clampt(x - P);
        auto T = [&](int xx, int k) -> uint32_t { return lvl[(size_t)xx + (size_t)r * ((size_t)ys[k & 1] + (size_t)r * zs[k >> 1])]; };
        uint2 o;
        o.x = T(x0, 0) | (T(x1, 0) << 8) | (T(x0, 1) << 16) | (T(x1, 1) << 24);
        o.y = T(x0, 2) | (T(x1, 2) << 8) | (T(x0, 3) << 16) | (T(x1, 3) << 24);
        s_out[(size_t)(zz * 4 + threadIdx.y) * stride + x] = o;
      }
    }
  } else if (live) {
    for (int m = (int)threadIdx.x; 4 * m < D; m += blockDim.x) {
      // per source plane a (z = zr0 - 1 - P + a, clamped) and source row (y0 | y1): bytes [0] = texel t0 - 1,
      // [1..4] = texels t0..t0+3 with t0 = 4m - P (all clamped to [0, r-1]; t0 is a multiple of 4)
      const int t0 = 4 * m - P;
      unsigned long long v[ZB + 1][2];
#pragma unroll
      for (int a = 0; a <= ZB; ++a) {
        const int z = clampt(zr0 - 1 - P + a);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint8_t* row = lvl + (size_t)r * ((size_t)ys[k] + (size_t)r * z);
          const uint32_t prev = row[clampt(t0 - 1)];
          const uint32_t word = (t0 >= 0 && t0 < r) ? *reinterpret_cast<const uint32_t*>(row + t0)
                                                    : (uint32_t)row[t0 < 0 ? 0 : r - 1] * 0x01010101u;
          v[a][k] = ((unsigned long long)word << 8) | prev;
        }
      }
#pragma unroll
      for (int zz = 0; zz < ZB; ++zz) {
        uint2* row_out = s_out + (size_t)(zz * 4 + threadIdx.y) * stride;
#pragma unroll
        for (int i = 0; i < 4; ++i) { // (up to three slots past the row end are written: the rows are padded)
          uint2 o;
          o.x = (uint32_t)((v[zz][0] >> (8 * i)) & 0xffffu) | ((uint32_t)((v[zz][1] >> (8 * i)) & 0xffffu) << 16);
          o.y = (uint32_t)((v[zz + 1][0] >> (8 * i)) & 0xffffu) | ((uint32_t)((v[zz + 1][1] >> (8 * i)) & 0xffffu) << 16);
          row_out[4 * m + i] = o;
        }
      }
    }
  }
  __syncthreads();
  if (live) {
    for (int zz = 0; zz < ZB && zr0 + zz < D; ++zz) {
      const uint2* row_out = s_out + (size_t)(zz * 4 + threadIdx.y) * stride;
      uint2* out = records + L.offset[l] + ((size_t)(zr0 + zz) * D + yr) * D;
      for (int x = (int)threadIdx.x; x < D; x += blockDim.x) out[x] = row_out[x];
    }
  }
}

} // namespace

// voxelization.cpp:161-171 (mip chain) + the gather-ready records.
static drv_status drv_impl_voxel_mips_and_records(drv_ctx* ctx) {
  const int res = (int)ctx->cfg.voxel_resolution;
  int sres = res;
  uint8_t* src = ctx->voxel_chain;
  while (sres > 1) { // up to four levels per launch
    VoxMipArgs A;
    memset(&A, 0, sizeof(A));
    A.src = src;
    A.src_res = sres;
    int r = sres, n = 0;
    uint8_t* level = src;
    while (n < 4 && r > 1) {
      level += (size_t)r * r * r;
      r >>= 1;
      A.dst[n++] = level;
    }
    A.levels = n;
    const int tiles = sres > 16 ? sres / 16 : 1;
    voxel_mip_chain_kernel<<<dim3(tiles, tiles, tiles), 256, 0, ctx->stream>>>(A);
    DRV_LAUNCH_CHECK();
    src = level;
    sres = r;
  }
  if (ctx->voxel_records) {
    RecordLevels L;
    memset(&L, 0, sizeof(L));
    L.levels = (int)ctx->voxel_levels;
    L.res = res;
    for (uint32_t l = 0; l < ctx->voxel_levels; ++l) {
      L.offset[l] = ctx->voxel_record_offset[l];
      L.chain_offset[l] = (uint32_t)voxel_level_offset_bytes((uint32_t)res, l);
    }
    RecordRows R;
    memset(&R, 0, sizeof(R));
    const uint32_t d0 = (uint32_t)res + 1u + 2u * kVoxelRecordPad; // records per axis of level 0
    // four slabs per block while 16 rows of records fit the default 48 KB of dynamic shared memory (res <= 256)
    const uint32_t zb = (size_t)16 * ((size_t)d0 + 3) * sizeof(uint2) <= 48 * 1024 ? 4u : 1u;
    uint32_t groups = 0;
    for (uint32_t l = 0; l < ctx->voxel_levels; ++l) {
      R.z_offset[l] = groups;
      groups += (((uint32_t)res >> l) + 1u + 2u * kVoxelRecordPad + zb - 1u) / zb;
    }
    R.z_offset[ctx->voxel_levels] = groups;
    // 64 x 4 threads: 64 groups of four records cover a row of up to 256 records (longer rows loop), 4 rows per block
    const dim3 grid((d0 + 3u) / 4u, groups), block(64, 4);
    const size_t smem = (size_t)zb * 4 * ((size_t)d0 + 3) * sizeof(uint2);
    if (zb == 4) voxel_records_kernel<4><<<grid, block, smem, ctx->stream>>>(L, R, ctx->voxel_chain, ctx->voxel_records);
    else voxel_records_kernel<1><<<grid, block, smem, ctx->stream>>>(L, R, ctx->voxel_chain, ctx->voxel_records);
    DRV_LAUNCH_CHECK();

  }
  return DRV_OK;
}

drv_status drv_impl_set_voxel_volume(drv_ctx* ctx, const uint8_t* level0) {
  const size_t vox = (size_t)ctx->cfg.voxel_resolution * ctx->cfg.voxel_resolution * ctx->cfg.voxel_resolution;
  DRV_CUDA(cudaMemcpyAsync(ctx->voxel_chain, level0, vox, cudaMemcpyDefault, ctx->stream));
  ctx->stage_begin(DRV_STAGE_VOXEL_BLEND_MIPMAP);
  drv_status st = drv_impl_voxel_mips_and_records(ctx);
  ctx->stage_end(DRV_STAGE_VOXEL_BLEND_MIPMAP);
  return st;
}

drv_status drv_impl_voxelize(drv_ctx* ctx, const float* tris, uint32_t n, const float* world, float adaption,
                             uint32_t flags) {
  if (!ctx->have_volume) return ctx->fail(DRV_ERR_NOT_BOUND, "drv_voxelize: VolumeInfo not set");
  const int res = (int)ctx->cfg.voxel_resolution;
  const size_t vox = (size_t)res * res * res;
  // adaptionThisFrameFloor == 0 skips everything (voxelization.cpp:100)
  const int k = (int)floorf(adaption * 255.0f + 0.5f);
  if (k <= 0) return DRV_OK;
  ctx->stage_begin(DRV_STAGE_VOXELIZE_SCENE);
  if (flags & DRV_VOXELIZE_CLEAR) DRV_CUDA(cudaMemsetAsync(ctx->voxel_target, 0, vox, ctx->stream)); // :105
  if (n > 0) {
    VoxParams P;
    memcpy(P.vmin, ctx->volume.VolumeWorldMin, 12);
    memcpy(P.vmax, ctx->volume.VolumeWorldMax, 12);
    memcpy(P.world, world, 64);
    P.res = res;
    const uint32_t groups = (n + kTrisPerBlock - 1) / kTrisPerBlock;
    // few, large triangles (architectural scenes): spread each group's pixels over several blocks
    const uint32_t slices = groups >= 4096 ? 1 : (groups >= 512 ? 4 : 16);
    voxelize_kernel<<<dim3(groups, slices), 256, 0, ctx->stream>>>(P, tris, n, ctx->voxel_target);
    DRV_LAUNCH_CHECK();
  }
  ctx->stage_end(DRV_STAGE_VOXELIZE_SCENE);
  if (!(flags & DRV_VOXELIZE_FINISH)) return DRV_OK;
  ctx->stage_begin(DRV_STAGE_VOXEL_BLEND_MIPMAP);
  size_t n16 = vox / 16;
  voxel_blend_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, ctx->stream>>>((uint4*)ctx->voxel_chain,
                                                                           (const uint4*)ctx->voxel_target, n16, k);
  DRV_LAUNCH_CHECK();
  drv_status st = drv_impl_voxel_mips_and_records(ctx);
  if (st != DRV_OK) return st;
  ctx->stage_end(DRV_STAGE_VOXEL_BLEND_MIPMAP);
  return DRV_OK;
}
