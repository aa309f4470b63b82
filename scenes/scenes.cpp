/* scenes.cpp — see scenes.h. Synthetic-input generator; not on the product path. */
#include "scenes.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Box {
  double lo[3], hi[3];
  float albedo[3];
};

struct D3 { double x, y, z; };
inline D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline D3 normalize(D3 a) { double l = std::sqrt(dot(a, a)); return {a.x / l, a.y / l, a.z / l}; }

struct Hit { double t; int box; int axis; int sign; };

const double kPi = 3.14159265358979323846;

} // namespace

struct scn_scene {
  std::vector<Box> boxes;
  double bbMin[3], bbMax[3];
};

namespace {

void add_box(scn_scene* s, double x0, double y0, double z0, double x1, double y1, double z1, float r, float g, float b) {
  Box bx;
  bx.lo[0] = std::min(x0, x1); bx.lo[1] = std::min(y0, y1); bx.lo[2] = std::min(z0, z1);
  bx.hi[0] = std::max(x0, x1); bx.hi[1] = std::max(y0, y1); bx.hi[2] = std::max(z0, z1);
  bx.albedo[0] = r; bx.albedo[1] = g; bx.albedo[2] = b;
  s->boxes.push_back(bx);
}

void finish(scn_scene* s, double scale) {
  for (int a = 0; a < 3; ++a) { s->bbMin[a] = 1e30; s->bbMax[a] = -1e30; }
  for (auto& b : s->boxes)
    for (int a = 0; a < 3; ++a) {
      b.lo[a] *= scale; b.hi[a] *= scale;
      s->bbMin[a] = std::min(s->bbMin[a], b.lo[a]);
      s->bbMax[a] = std::max(s->bbMax[a], b.hi[a]);
    }
}

void build_cornell(scn_scene* s) {
  const double t = 0.25; /* wall thickness, outside the unit interior */
  add_box(s, -2.5 - t, -t, -2.5 - t, 2.5 + t, 0.0, 2.5, 0.73f, 0.73f, 0.73f);        /* floor */
  add_box(s, -2.5 - t, 5.0, -2.5 - t, 2.5 + t, 5.0 + t, 2.5, 0.73f, 0.73f, 0.73f);   /* ceiling */
  add_box(s, -2.5 - t, 0.0, -2.5 - t, 2.5 + t, 5.0, -2.5, 0.73f, 0.73f, 0.73f);      /* back */
  add_box(s, -2.5 - t, 0.0, -2.5, -2.5, 5.0, 2.5, 0.65f, 0.05f, 0.05f);              /* left, red */
  add_box(s, 2.5, 0.0, -2.5, 2.5 + t, 5.0, 2.5, 0.12f, 0.45f, 0.15f);                /* right, green */
  add_box(s, -1.7, 0.0, -1.6, -0.2, 3.0, -0.3, 0.73f, 0.73f, 0.73f);                 /* tall block */
  add_box(s, 0.4, 0.0, 0.1, 1.9, 1.5, 1.6, 0.73f, 0.73f, 0.73f);                     /* short block */
}

void build_atrium(scn_scene* s) {
  const double t = 0.5;
  const double X = 5.0, Y = 7.0, Z0 = -5.0, Z1 = 7.0;
  add_box(s, -X - t, -t, Z0 - t, X + t, 0.0, Z1 + t, 0.58f, 0.54f, 0.48f);   /* floor */
  add_box(s, -X - t, Y, Z0 - t, X + t, Y + t, Z1 + t, 0.70f, 0.68f, 0.62f);  /* ceiling */
  add_box(s, -X - t, 0.0, Z0 - t, -X, Y, Z1 + t, 0.62f, 0.35f, 0.28f);       /* left wall */
  add_box(s, X, 0.0, Z0 - t, X + t, Y, Z1 + t, 0.30f, 0.42f, 0.58f);         /* right wall */
  add_box(s, -X, 0.0, Z0 - t, X, Y, Z0, 0.66f, 0.62f, 0.52f);                /* far end wall */
  add_box(s, -X, 0.0, Z1, X, Y, Z1 + t, 0.66f, 0.62f, 0.52f);                /* near end wall */
  const double colX = 2.8, colW = 0.35, colH = 4.4;
  const double zs[5] = {-3.6, -1.2, 1.2, 3.6, 6.0};
  for (int side = -1; side <= 1; side += 2) {
    for (int i = 0; i < 5; ++i) { /* columns with a base and a capital */
      double cx = side * colX, cz = zs[i];
      add_box(s, cx - colW, 0.0, cz - colW, cx + colW, colH, cz + colW, 0.72f, 0.70f, 0.64f);
      add_box(s, cx - colW - 0.12, 0.0, cz - colW - 0.12, cx + colW + 0.12, 0.3, cz + colW + 0.12, 0.60f, 0.58f, 0.55f);
      add_box(s, cx - colW - 0.12, colH - 0.3, cz - colW - 0.12, cx + colW + 0.12, colH, cz + colW + 0.12, 0.60f, 0.58f, 0.55f);
      /* arch from the column to the wall */
      add_box(s, side * (colX + colW), colH, cz - 0.25, side * X, colH + 0.5, cz + 0.25, 0.68f, 0.64f, 0.56f);
    }
    /* lintel along the column row */
    add_box(s, side * colX - 0.3, colH, Z0, side * colX + 0.3, colH + 0.6, Z1, 0.68f, 0.64f, 0.56f);
    /* gallery slab between the row and the wall */
    add_box(s, side * (colX - 0.3), colH + 0.6, Z0, side * X, colH + 0.85, Z1, 0.55f, 0.50f, 0.44f);
    /* gallery parapet */
    add_box(s, side * (colX - 0.3), colH + 0.85, Z0, side * (colX - 0.1), colH + 1.6, Z1, 0.50f, 0.46f, 0.40f);
  }
  for (int i = 0; i < 4; ++i) { /* ceiling beams */
    double cz = -3.0 + 3.0 * i;
    add_box(s, -X, Y - 0.45, cz - 0.2, X, Y, cz + 0.2, 0.45f, 0.32f, 0.22f);
  }
  /* a few props on the floor */
  add_box(s, -0.8, 0.0, -2.6, 0.8, 0.9, -1.6, 0.25f, 0.50f, 0.30f);
  add_box(s, 1.2, 0.0, 0.4, 1.9, 1.4, 1.1, 0.70f, 0.20f, 0.18f);
  add_box(s, -2.0, 0.0, 2.4, -1.3, 0.6, 3.6, 0.22f, 0.30f, 0.62f);
}

bool trace(const scn_scene* s, D3 o, D3 d, Hit& best) {
  best.t = 1e300; best.box = -1;
  const double inv[3] = {1.0 / d.x, 1.0 / d.y, 1.0 / d.z};
  const double oo[3] = {o.x, o.y, o.z};
  for (size_t bi = 0; bi < s->boxes.size(); ++bi) {
    const Box& b = s->boxes[bi];
    double tn = -1e300, tf = 1e300;
    int an = -1, sn = 0;
    bool miss = false;
    for (int a = 0; a < 3; ++a) {
      double t0 = (b.lo[a] - oo[a]) * inv[a], t1 = (b.hi[a] - oo[a]) * inv[a];
      int sg = -1;
      if (t0 > t1) { std::swap(t0, t1); sg = 1; }
      if (t0 != t0 || t1 != t1) { /* ray parallel to the slab and on its plane */
        if (oo[a] < b.lo[a] || oo[a] > b.hi[a]) { miss = true; break; }
        continue;
      }
      if (t0 > tn) { tn = t0; an = a; sn = sg; }
      if (t1 < tf) tf = t1;
      if (tn > tf) { miss = true; break; }
    }
    if (miss || tf < 1e-9 || tn < 1e-9) continue; /* origin is never inside a solid */
    if (tn < best.t) { best.t = tn; best.box = (int)bi; best.axis = an; best.sign = sn; }
  }
  return best.box >= 0;
}

inline void mulRM(const float* m, const double v[4], double out[4]) {
  for (int j = 0; j < 4; ++j)
    out[j] = (double)m[j * 4] * v[0] + (double)m[j * 4 + 1] * v[1] + (double)m[j * 4 + 2] * v[2] + (double)m[j * 4 + 3] * v[3];
}

inline uint16_t to_half(float f) {
  _Float16 h = (_Float16)f; /* round-to-nearest-even */
  uint16_t u;
  std::memcpy(&u, &h, 2);
  return u;
}

inline int sat_trunc(double v) {
  if (v != v) return 0;
  if (v > 32767.0) return 32767;
  if (v < -32768.0) return -32768;
  return (int)v;
}

/* utils.glsl:62-89 on float inputs; int16 clamp as the render target does. */
inline void pack_normal(const float n[3], int16_t* out) {
  const float PI = 3.14159265358979f;
  float px = (n[0] == 0.0f) ? ((n[1] > 0.0f ? 1.0f : (n[1] < 0.0f ? -1.0f : 0.0f)) * PI / 2) : std::atan2(n[1], n[0]);
  out[0] = (int16_t)sat_trunc(px * (32768.0f / PI));
  out[1] = (int16_t)sat_trunc(n[2] * 32768.0f);
}

inline uint8_t linear_to_srgb8(float c) {
  double l = std::min(1.0, std::max(0.0, (double)c));
  double e = (l <= 0.0031308) ? l * 12.92 : 1.055 * std::pow(l, 1.0 / 2.4) - 0.055;
  return (uint8_t)std::floor(e * 255.0 + 0.5);
}

void parallel_rows(int rows, int threads, const std::function<void(int, int)>& fn) {
  if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
  threads = std::min(threads, rows);
  if (threads <= 1) { fn(0, rows); return; }
  std::vector<std::thread> pool;
  int chunk = (rows + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    int b = t * chunk, e = std::min(rows, b + chunk);
    if (b < e) pool.emplace_back(fn, b, e);
  }
  for (auto& th : pool) th.join();
}

/* utilities/random.cpp:5-22 */
inline uint32_t wang_hash(uint32_t seed) {
  seed = (seed ^ 61u) ^ (seed >> 16);
  seed *= 9u;
  seed = seed ^ (seed >> 4);
  seed *= 0x27d4eb2du;
  seed = seed ^ (seed >> 15);
  return seed;
}
inline uint32_t xorshift(uint32_t s) {
  s ^= (s << 13);
  s ^= (s >> 17);
  s ^= (s << 5);
  return s;
}
struct Rng {
  uint32_t s;
  explicit Rng(uint32_t seed) : s(wang_hash(seed)) { if (s == 0) s = 0x9E3779B9u; }
  float next() { s = xorshift(s); return (float)(s >> 8) * (1.0f / 16777216.0f); } /* [0,1) */
};

} // namespace

extern "C" scn_scene* scn_create(const char* name, float scale) {
  scn_scene* s = new scn_scene();
  std::string n(name ? name : "");
  if (n == "cornell") build_cornell(s);
  else if (n == "atrium") build_atrium(s);
  else { delete s; return nullptr; }
  finish(s, scale > 0.0f ? scale : 1.0);
  return s;
}
extern "C" void scn_destroy(scn_scene* s) { delete s; }
extern "C" uint32_t scn_num_boxes(const scn_scene* s) { return (uint32_t)s->boxes.size(); }
extern "C" void scn_bounding_box(const scn_scene* s, float bmin[3], float bmax[3]) {
  for (int a = 0; a < 3; ++a) { bmin[a] = (float)s->bbMin[a]; bmax[a] = (float)s->bbMax[a]; }
}

extern "C" uint32_t scn_triangles(const scn_scene* s, float* out, uint32_t max_tris) {
  uint32_t n = (uint32_t)s->boxes.size() * 12;
  if (!out) return n;
  static const int faces[6][4] = {{0, 2, 6, 4}, {1, 5, 7, 3}, {0, 4, 5, 1}, {2, 3, 7, 6}, {0, 1, 3, 2}, {4, 6, 7, 5}};
  uint32_t w = 0;
  for (const Box& b : s->boxes) {
    float c[8][3];
    for (int i = 0; i < 8; ++i) {
      c[i][0] = (float)((i & 1) ? b.hi[0] : b.lo[0]);
      c[i][1] = (float)((i & 2) ? b.hi[1] : b.lo[1]);
      c[i][2] = (float)((i & 4) ? b.hi[2] : b.lo[2]);
    }
    for (int f = 0; f < 6; ++f) {
      const int tri[2][3] = {{faces[f][0], faces[f][1], faces[f][2]}, {faces[f][0], faces[f][2], faces[f][3]}};
      for (int k = 0; k < 2; ++k) {
        if (w >= max_tris) return w;
        for (int v = 0; v < 3; ++v) std::memcpy(out + (size_t)w * 9 + v * 3, c[tri[k][v]], 12);
        ++w;
      }
    }
  }
  return w;
}

extern "C" void scn_render_gbuffer(const scn_scene* s, const drv_per_frame* pf, uint32_t width, uint32_t height,
                                   float* depth, int16_t* normal, uint8_t* diffuse, int threads) {
  D3 cam = {pf->CameraPosition[0], pf->CameraPosition[1], pf->CameraPosition[2]};
  parallel_rows((int)height, threads, [&](int y0, int y1) {
    for (int y = y0; y < y1; ++y)
      for (uint32_t x = 0; x < width; ++x) {
        size_t t = (size_t)y * width + x;
        depth[t] = 0.0f;
        normal[t * 2] = normal[t * 2 + 1] = 0;
        diffuse[t * 4] = diffuse[t * 4 + 1] = diffuse[t * 4 + 2] = 0; diffuse[t * 4 + 3] = 255;
        double ndc[4] = {((double)x + 0.5) / width * 2.0 - 1.0, ((double)y + 0.5) / height * 2.0 - 1.0, 0.5, 1.0};
        double w4[4];
        mulRM(pf->InverseViewProjection, ndc, w4);
        D3 p = {w4[0] / w4[3], w4[1] / w4[3], w4[2] / w4[3]};
        D3 d = normalize(p - cam);
        Hit h;
        if (!trace(s, cam, d, h)) continue;
        D3 P = cam + d * h.t;
        double pw[4] = {P.x, P.y, P.z, 1.0}, clip[4];
        mulRM(pf->ViewProjection, pw, clip);
        depth[t] = (float)(clip[2] / clip[3]);
        float n[3] = {0, 0, 0};
        n[h.axis] = (float)h.sign;
        pack_normal(n, normal + t * 2);
        const Box& b = s->boxes[h.box];
        for (int c = 0; c < 3; ++c) diffuse[t * 4 + c] = linear_to_srgb8(b.albedo[c]);
      }
  });
}

extern "C" void scn_render_rsm(const scn_scene* s, const drv_spot_light* L, uint16_t* flux, int16_t* normal,
                               uint16_t* depthlinsq, int threads) {
  const int R = L->RSMRenderResolution;
  D3 lp = {L->LightPosition[0], L->LightPosition[1], L->LightPosition[2]};
  D3 ld = {L->LightDirection[0], L->LightDirection[1], L->LightDirection[2]};
  const double cosHalf = L->LightCosHalfAngle;
  parallel_rows(R, threads, [&](int y0, int y1) {
    for (int y = y0; y < y1; ++y)
      for (int x = 0; x < R; ++x) {
        size_t t = (size_t)y * R + x;
        for (int c = 0; c < 4; ++c) flux[t * 4 + c] = 0;
        normal[t * 2] = normal[t * 2 + 1] = 0;
        depthlinsq[t * 2] = depthlinsq[t * 2 + 1] = 0;
        double clip[4] = {((double)x + 0.5) / R * 2.0 - 1.0, ((double)y + 0.5) / R * 2.0 - 1.0, 0.5, 1.0};
        double w4[4];
        mulRM(L->InverseLightViewProjection, clip, w4);
        D3 p = {w4[0] / w4[3], w4[1] / w4[3], w4[2] / w4[3]};
        D3 d = normalize(p - lp);
        Hit h;
        if (!trace(s, lp, d, h)) continue;
        /* fillrsm.frag:32-48 */
        double distToLight = h.t;
        double cosToLight = std::min(1.0, std::max(0.0, dot(d, ld)));
        double totalSpotSteradian = 2.0 * kPi * (1.0 - cosHalf);
        double pixelSteradian = totalSpotSteradian * cosToLight / R / R;
        double spotFalloff = std::min(1.0, std::max(0.0, cosToLight - cosHalf)) / (1.0 - cosHalf);
        const Box& b = s->boxes[h.box];
        for (int c = 0; c < 3; ++c)
          flux[t * 4 + c] = to_half((float)(b.albedo[c] * L->LightIntensity[c] * (spotFalloff * pixelSteradian / kPi)));
        depthlinsq[t * 2] = to_half((float)distToLight);
        depthlinsq[t * 2 + 1] = to_half((float)(distToLight * distToLight));
        float n[3] = {0, 0, 0};
        n[h.axis] = (float)h.sign;
        pack_normal(n, normal + t * 2);
      }
  });
}

extern "C" void scn_sweep_entries(uint32_t seed, uint32_t n, float* out) {
  for (uint32_t i = 0; i < n; ++i) {
    Rng r(seed + i);
    out[i * 4 + 0] = r.next() * 16.0f - 8.0f;
    out[i * 4 + 1] = r.next() * 16.0f - 8.0f;
    out[i * 4 + 2] = r.next() * 16.0f - 8.0f;
    out[i * 4 + 3] = 0.0f;
  }
}

extern "C" void scn_sweep_vpls(uint32_t seed, uint32_t n, float val_area_factor, drv_vpl* out) {
  for (uint32_t i = 0; i < n; ++i) {
    Rng r(seed + i);
    drv_vpl v;
    std::memset(&v, 0, sizeof(v));
    int face = std::min(5, (int)(r.next() * 6.0f));
    int axis = face >> 1;
    float sgn = (face & 1) ? 1.0f : -1.0f;
    float a = r.next() * 16.0f - 8.0f, b = r.next() * 16.0f - 8.0f;
    float pos[3];
    pos[axis] = sgn * 8.0f;
    pos[(axis + 1) % 3] = a;
    pos[(axis + 2) % 3] = b;
    /* inward face normal jittered within 30 degrees */
    float cosT = 1.0f - r.next() * (1.0f - 0.8660254f);
    float sinT = std::sqrt(std::max(0.0f, 1.0f - cosT * cosT));
    float phi = r.next() * 6.2831853f;
    float nrm[3];
    nrm[axis] = -sgn * cosT;
    nrm[(axis + 1) % 3] = sinT * std::cos(phi);
    nrm[(axis + 2) % 3] = sinT * std::sin(phi);
    float len = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    for (int c = 0; c < 3; ++c) { v.Position[c] = pos[c]; v.Normal[c] = nrm[c] / len; v.Flux[c] = r.next() * 1e-3f; }
    float d = 1.0f + r.next() * 15.0f;
    v.DiscArea = d * d * val_area_factor;
    out[i] = v;
  }
}
