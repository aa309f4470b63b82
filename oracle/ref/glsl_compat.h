/*
 * glsl_compat.h — a small GLSL 4.50 execution environment in C++17, so that the
 * reference's UNMODIFIED shader text (shader/*.comp, *.frag, *.glsl under
 * /root/reference) can be compiled by g++ after the mechanical rewrite of
 * oracle/ref/glsl2cpp.py and run on the CPU. TEST INFRASTRUCTURE ONLY: it pins
 * oracle/ (the hand-written restatement) to the reference's own source text;
 * nothing under dynamicradiancevolume_b200/ may include or link it.
 *
 * What a GL driver would decide is decided here once, with the arithmetic policy
 * of oracle/oracle.h: IEEE binary32, every * and + rounded separately (build with
 * -ffp-contract=off), dot products summed left to right, normalize(v) =
 * v * (1/sqrt(dot(v,v))), mix(a,b,t) = a*(1-t) + b*t, float->int truncates and
 * saturates; texture filtering follows the OpenGL 4.5 spec section 8.14 formulas
 * (SURVEY.md D.0); UNORM8 stores round to nearest.
 *
 * Work-group semantics: every invocation of a compute work group is a fiber
 * (ucontext); barrier() switches to the next fiber, so shared memory and
 * barriers behave as on the GPU without touching the shader text.
 */
#ifndef DRV_GLSL_COMPAT_H
#define DRV_GLSL_COMPAT_H

#include <ucontext.h>

#include "../../include/drv_r11g11b10.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <type_traits>
#include <vector>

namespace glsl {

typedef unsigned int uint;

/* ---------------------------------------------------------------- scalar helpers */
inline int f2i(float f) { /* float -> int: truncate, saturate, NaN -> 0 */
  if (f != f) return 0;
  if (f >= 2147483648.0f) return INT_MAX;
  if (f <= -2147483648.0f) return INT_MIN;
  return static_cast<int>(f);
}
inline uint f2u(float f) {
  if (!(f > 0.0f)) return 0u;
  if (f >= 4294967296.0f) return 0xFFFFFFFFu;
  return static_cast<uint>(f);
}
template <class T, class U>
inline T conv(U v) {
  if constexpr (std::is_floating_point<U>::value && std::is_same<T, int>::value) return f2i((float)v);
  else if constexpr (std::is_floating_point<U>::value && std::is_same<T, uint>::value) return f2u((float)v);
  else return static_cast<T>(v);
}
/* GLSL's implicit conversions: int -> uint, int -> float, uint -> float. */
template <class U, class T>
struct implicit_ok : std::integral_constant<bool, (std::is_same<U, int>::value && (std::is_same<T, uint>::value || std::is_same<T, float>::value)) ||
                                                      (std::is_same<U, uint>::value && std::is_same<T, float>::value)> {};
template <class S>
using arith = typename std::enable_if<std::is_arithmetic<S>::value, int>::type;

template <class T> struct tvec2;
template <class T> struct tvec3;
template <class T> struct tvec4;
template <class T, int N> struct vec_of;
template <class T> struct vec_of<T, 1> { typedef T type; };
template <class T> struct vec_of<T, 2> { typedef tvec2<T> type; };
template <class T> struct vec_of<T, 3> { typedef tvec3<T> type; };
template <class T> struct vec_of<T, 4> { typedef tvec4<T> type; };

/* Operators are non-template friends defined in the class (so GLSL's implicit conversions apply to the other
 * operand); an integer vector combined with a floating scalar promotes to the float vector, as in GLSL. */
#define GLSL_VEC_COMMON(V, N)                                                                                     \
  T& operator[](int i) { return (&x)[i]; }                                                                        \
  const T& operator[](int i) const { return (&x)[i]; }                                                            \
  template <int... I> typename vec_of<T, sizeof...(I)>::type swz() const {                                        \
    return typename vec_of<T, sizeof...(I)>::type((*this)[I]...);                                                 \
  }                                                                                                               \
  template <int... I> void set_swz(const typename vec_of<T, sizeof...(I)>::type& v) {                             \
    const int idx[] = {I...};                                                                                     \
    for (int k = 0; k < (int)sizeof...(I); ++k) (*this)[idx[k]] = v[k];                                           \
  }                                                                                                               \
  friend V operator+(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + b[i]; return r; }   \
  friend V operator-(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - b[i]; return r; }   \
  friend V operator*(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * b[i]; return r; }   \
  friend V operator/(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / b[i]; return r; }   \
  friend V operator+(const V& a, T s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + s; return r; }             \
  friend V operator-(const V& a, T s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - s; return r; }             \
  friend V operator*(const V& a, T s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * s; return r; }             \
  friend V operator/(const V& a, T s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / s; return r; }             \
  friend V operator+(T s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s + a[i]; return r; }             \
  friend V operator-(T s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s - a[i]; return r; }             \
  friend V operator*(T s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s * a[i]; return r; }             \
  friend V operator/(T s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s / a[i]; return r; }             \
  friend V operator-(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = -a[i]; return r; }                     \
  template <class S, typename std::enable_if<std::is_floating_point<S>::value && std::is_integral<T>::value, int>::type = 0> \
  friend typename vec_of<float, N>::type operator*(const V& a, S s) { return typename vec_of<float, N>::type(a) * (float)s; } \
  template <class S, typename std::enable_if<std::is_floating_point<S>::value && std::is_integral<T>::value, int>::type = 0> \
  friend typename vec_of<float, N>::type operator*(S s, const V& a) { return (float)s * typename vec_of<float, N>::type(a); } \
  template <class S, typename std::enable_if<std::is_floating_point<S>::value && std::is_integral<T>::value, int>::type = 0> \
  friend typename vec_of<float, N>::type operator/(const V& a, S s) { return typename vec_of<float, N>::type(a) / (float)s; } \
  template <class S, typename std::enable_if<std::is_floating_point<S>::value && std::is_integral<T>::value, int>::type = 0> \
  friend typename vec_of<float, N>::type operator+(const V& a, S s) { return typename vec_of<float, N>::type(a) + (float)s; } \
  template <class S, typename std::enable_if<std::is_floating_point<S>::value && std::is_integral<T>::value, int>::type = 0> \
  friend typename vec_of<float, N>::type operator-(const V& a, S s) { return typename vec_of<float, N>::type(a) - (float)s; } \
  V& operator+=(const V& b) { *this = *this + b; return *this; }                                                  \
  V& operator-=(const V& b) { *this = *this - b; return *this; }                                                  \
  V& operator*=(const V& b) { *this = *this * b; return *this; }                                                  \
  V& operator/=(const V& b) { *this = *this / b; return *this; }                                                  \
  V& operator+=(T s) { *this = *this + s; return *this; }                                                         \
  V& operator-=(T s) { *this = *this - s; return *this; }                                                         \
  V& operator*=(T s) { *this = *this * s; return *this; }                                                         \
  V& operator/=(T s) { *this = *this / s; return *this; }                                                         \
  friend bool operator==(const V& a, const V& b) { for (int i = 0; i < N; ++i) if (!(a[i] == b[i])) return false; return true; } \
  friend bool operator!=(const V& a, const V& b) { return !(a == b); }                                            \
  friend typename vec_of<bool, N>::type lessThan(const V& a, const V& b) { typename vec_of<bool, N>::type r; for (int i = 0; i < N; ++i) r[i] = a[i] < b[i]; return r; } \
  friend typename vec_of<bool, N>::type lessThanEqual(const V& a, const V& b) { typename vec_of<bool, N>::type r; for (int i = 0; i < N; ++i) r[i] = a[i] <= b[i]; return r; } \
  friend typename vec_of<bool, N>::type greaterThan(const V& a, const V& b) { typename vec_of<bool, N>::type r; for (int i = 0; i < N; ++i) r[i] = a[i] > b[i]; return r; } \
  friend typename vec_of<bool, N>::type greaterThanEqual(const V& a, const V& b) { typename vec_of<bool, N>::type r; for (int i = 0; i < N; ++i) r[i] = a[i] >= b[i]; return r; } \
  friend typename vec_of<bool, N>::type equal(const V& a, const V& b) { typename vec_of<bool, N>::type r; for (int i = 0; i < N; ++i) r[i] = a[i] == b[i]; return r; } \
  friend typename vec_of<bool, N>::type notEqual(const V& a, const V& b) { typename vec_of<bool, N>::type r; for (int i = 0; i < N; ++i) r[i] = a[i] != b[i]; return r; } \
  /* integer-only operators: instantiated lazily, so float vectors never see them */                              \
  template <class Q = T, typename std::enable_if<std::is_integral<Q>::value && !std::is_same<Q, bool>::value, int>::type = 0> \
  V& operator&=(Q s) { for (int i = 0; i < N; ++i) (*this)[i] &= s; return *this; }                               \
  template <class Q = T, typename std::enable_if<std::is_integral<Q>::value && !std::is_same<Q, bool>::value, int>::type = 0> \
  V& operator|=(const typename vec_of<Q, N>::type& b) { for (int i = 0; i < N; ++i) (*this)[i] |= b[i]; return *this; } \
  template <class Q = T, typename std::enable_if<std::is_integral<Q>::value && !std::is_same<Q, bool>::value, int>::type = 0> \
  friend V operator>>(const V& a, int s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] >> s; return r; }         \
  template <class Q = T, typename std::enable_if<std::is_integral<Q>::value && !std::is_same<Q, bool>::value, int>::type = 0> \
  friend V operator<<(const V& a, int s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] << s; return r; }         \
  template <class Q = T, typename std::enable_if<std::is_integral<Q>::value && !std::is_same<Q, bool>::value, int>::type = 0> \
  friend V operator%(const V& a, Q s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] % s; return r; }

template <class T>
struct tvec2 {
  union { struct { T x, y; }; struct { T r, g; }; struct { T s, t; }; };
  tvec2() : x(T()), y(T()) {}
  template <class A, arith<A> = 0> explicit tvec2(A v) : x(conv<T>(v)), y(conv<T>(v)) {}
  template <class A, class B, arith<A> = 0, arith<B> = 0> tvec2(A a, B b) : x(conv<T>(a)), y(conv<T>(b)) {}
  template <class U, typename std::enable_if<implicit_ok<U, T>::value, int>::type = 0>
  tvec2(const tvec2<U>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)) {}
  template <class U, typename std::enable_if<!implicit_ok<U, T>::value && !std::is_same<U, T>::value, int>::type = 0>
  explicit tvec2(const tvec2<U>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)) {}
  template <class U> explicit tvec2(const tvec3<U>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)) {}
  template <class U> explicit tvec2(const tvec4<U>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)) {}
  GLSL_VEC_COMMON(tvec2, 2)
};
template <class T>
struct tvec3 {
  union { struct { T x, y, z; }; struct { T r, g, b; }; struct { T s, t, p; }; };
  tvec3() : x(T()), y(T()), z(T()) {}
  template <class A, arith<A> = 0> explicit tvec3(A v) : x(conv<T>(v)), y(conv<T>(v)), z(conv<T>(v)) {}
  template <class A, class B, class C, arith<A> = 0, arith<B> = 0, arith<C> = 0>
  tvec3(A a, B b, C c) : x(conv<T>(a)), y(conv<T>(b)), z(conv<T>(c)) {}
  template <class U, class C, arith<C> = 0> tvec3(const tvec2<U>& a, C c) : x(conv<T>(a.x)), y(conv<T>(a.y)), z(conv<T>(c)) {}
  template <class A, class U, arith<A> = 0> tvec3(A a, const tvec2<U>& b) : x(conv<T>(a)), y(conv<T>(b.x)), z(conv<T>(b.y)) {}
  template <class U, typename std::enable_if<implicit_ok<U, T>::value, int>::type = 0>
  tvec3(const tvec3<U>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)), z(conv<T>(o.z)) {}
  template <class U, typename std::enable_if<!implicit_ok<U, T>::value && !std::is_same<U, T>::value, int>::type = 0>
  explicit tvec3(const tvec3<U>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)), z(conv<T>(o.z)) {}
  template <class U> explicit tvec3(const tvec4<U>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)), z(conv<T>(o.z)) {}
  GLSL_VEC_COMMON(tvec3, 3)
};
template <class T>
struct tvec4 {
  union { struct { T x, y, z, w; }; struct { T r, g, b, a; }; struct { T s, t, p, q; }; };
  tvec4() : x(T()), y(T()), z(T()), w(T()) {}
  template <class A, arith<A> = 0> explicit tvec4(A v) : x(conv<T>(v)), y(conv<T>(v)), z(conv<T>(v)), w(conv<T>(v)) {}
  template <class A, class B, class C, class D, arith<A> = 0, arith<B> = 0, arith<C> = 0, arith<D> = 0>
  tvec4(A a, B b, C c, D d) : x(conv<T>(a)), y(conv<T>(b)), z(conv<T>(c)), w(conv<T>(d)) {}
  template <class U, class C, class D, arith<C> = 0, arith<D> = 0>
  tvec4(const tvec2<U>& a, C c, D d) : x(conv<T>(a.x)), y(conv<T>(a.y)), z(conv<T>(c)), w(conv<T>(d)) {}
  template <class U, class V2> tvec4(const tvec2<U>& a, const tvec2<V2>& b) : x(conv<T>(a.x)), y(conv<T>(a.y)), z(conv<T>(b.x)), w(conv<T>(b.y)) {}
  template <class U, class D, arith<D> = 0> tvec4(const tvec3<U>& a, D d) : x(conv<T>(a.x)), y(conv<T>(a.y)), z(conv<T>(a.z)), w(conv<T>(d)) {}
  template <class A, class U, arith<A> = 0> tvec4(A a, const tvec3<U>& b) : x(conv<T>(a)), y(conv<T>(b.x)), z(conv<T>(b.y)), w(conv<T>(b.z)) {}
  template <class U, typename std::enable_if<implicit_ok<U, T>::value, int>::type = 0>
  tvec4(const tvec4<U>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)), z(conv<T>(o.z)), w(conv<T>(o.w)) {}
  template <class U, typename std::enable_if<!implicit_ok<U, T>::value && !std::is_same<U, T>::value, int>::type = 0>
  explicit tvec4(const tvec4<U>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)), z(conv<T>(o.z)), w(conv<T>(o.w)) {}
  GLSL_VEC_COMMON(tvec4, 4)
};

typedef tvec2<float> vec2; typedef tvec3<float> vec3; typedef tvec4<float> vec4;
typedef tvec2<int> ivec2;  typedef tvec3<int> ivec3;  typedef tvec4<int> ivec4;
typedef tvec2<uint> uvec2; typedef tvec3<uint> uvec3; typedef tvec4<uint> uvec4;
typedef tvec2<bool> bvec2; typedef tvec3<bool> bvec3; typedef tvec4<bool> bvec4;

inline bool any(const bvec2& v) { return v.x || v.y; }
inline bool any(const bvec3& v) { return v.x || v.y || v.z; }
inline bool any(const bvec4& v) { return v.x || v.y || v.z || v.w; }
inline bool all(const bvec2& v) { return v.x && v.y; }
inline bool all(const bvec3& v) { return v.x && v.y && v.z; }
inline bool all(const bvec4& v) { return v.x && v.y && v.z && v.w; }
/* NVIDIA's compiler accepts && between boolean vectors (lightcache.glsl:115-116) as the component-wise AND */
inline bvec2 operator&&(const bvec2& a, const bvec2& b) { return bvec2(a.x && b.x, a.y && b.y); }
inline bvec3 operator&&(const bvec3& a, const bvec3& b) { return bvec3(a.x && b.x, a.y && b.y, a.z && b.z); }

/* ---------------------------------------------------------------- built-in functions (scalar) */
inline float sqrt(float x) { return ::sqrtf(x); }
inline float inversesqrt(float x) { return 1.0f / ::sqrtf(x); }
inline float abs(float x) { return ::fabsf(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float floor(float x) { return ::floorf(x); }
inline float ceil(float x) { return ::ceilf(x); }
inline float fract(float x) { return x - ::floorf(x); }
inline float sin(float x) { return ::sinf(x); }
inline float cos(float x) { return ::cosf(x); }
inline float tan(float x) { return ::tanf(x); }
inline float acos(float x) { return ::acosf(x); }
inline float atan(float y, float x) { return ::atan2f(y, x); }
inline float atan(float x) { return ::atanf(x); }
inline float pow(float a, float b) { return ::powf(a, b); }
inline float exp2(float x) { return ::exp2f(x); }
inline float log2(float x) { return ::log2f(x); }
inline float exp(float x) { return ::expf(x); }
inline float log(float x) { return ::logf(x); }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float min(float a, float b) { return ::fminf(a, b); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline uint max(uint a, uint b) { return a > b ? a : b; }
inline float clamp(float x, float lo, float hi) { return ::fminf(::fmaxf(x, lo), hi); }
inline int clamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float mod(float x, float y) { return x - y * ::floorf(x / y); }

/* component-wise lifts */
#define GLSL_LIFT1(F)                                                                              \
  inline vec2 F(const vec2& a) { return vec2(F(a.x), F(a.y)); }                                    \
  inline vec3 F(const vec3& a) { return vec3(F(a.x), F(a.y), F(a.z)); }                            \
  inline vec4 F(const vec4& a) { return vec4(F(a.x), F(a.y), F(a.z), F(a.w)); }
GLSL_LIFT1(sqrt) GLSL_LIFT1(inversesqrt) GLSL_LIFT1(abs) GLSL_LIFT1(floor) GLSL_LIFT1(ceil) GLSL_LIFT1(fract)
GLSL_LIFT1(sign) GLSL_LIFT1(exp2) GLSL_LIFT1(log2) GLSL_LIFT1(sin) GLSL_LIFT1(cos)
#define GLSL_LIFT2(F, V2, V3, V4, S)                                                               \
  inline V2 F(const V2& a, const V2& b) { return V2(F(a.x, b.x), F(a.y, b.y)); }                   \
  inline V3 F(const V3& a, const V3& b) { return V3(F(a.x, b.x), F(a.y, b.y), F(a.z, b.z)); }      \
  inline V4 F(const V4& a, const V4& b) { return V4(F(a.x, b.x), F(a.y, b.y), F(a.z, b.z), F(a.w, b.w)); } \
  inline V2 F(const V2& a, S b) { return V2(F(a.x, b), F(a.y, b)); }                               \
  inline V3 F(const V3& a, S b) { return V3(F(a.x, b), F(a.y, b), F(a.z, b)); }                    \
  inline V4 F(const V4& a, S b) { return V4(F(a.x, b), F(a.y, b), F(a.z, b), F(a.w, b)); }
GLSL_LIFT2(min, vec2, vec3, vec4, float) GLSL_LIFT2(max, vec2, vec3, vec4, float)
GLSL_LIFT2(min, ivec2, ivec3, ivec4, int) GLSL_LIFT2(max, ivec2, ivec3, ivec4, int)
GLSL_LIFT2(pow, vec2, vec3, vec4, float) GLSL_LIFT2(mod, vec2, vec3, vec4, float)
#define GLSL_CLAMP(V, S)                                                                           \
  inline V clamp(const V& x, const V& lo, const V& hi) { return min(max(x, lo), hi); }             \
  inline V clamp(const V& x, S lo, S hi) { return min(max(x, lo), hi); }
GLSL_CLAMP(vec2, float) GLSL_CLAMP(vec3, float) GLSL_CLAMP(vec4, float)
GLSL_CLAMP(ivec2, int) GLSL_CLAMP(ivec3, int) GLSL_CLAMP(ivec4, int)
#define GLSL_MIX(V)                                                                                \
  inline V mix(const V& a, const V& b, float t) { return a * (1.0f - t) + b * t; }                 \
  inline V mix(const V& a, const V& b, const V& t) { return a * (V(1.0f) - t) + b * t; }
GLSL_MIX(vec2) GLSL_MIX(vec3) GLSL_MIX(vec4)

inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline float length(const vec2& a) { return ::sqrtf(dot(a, a)); }
inline float length(const vec3& a) { return ::sqrtf(dot(a, a)); }
inline float length(const vec4& a) { return ::sqrtf(dot(a, a)); }
inline float length(float a) { return ::fabsf(a); }
inline vec2 normalize(const vec2& a) { return a * inversesqrt(dot(a, a)); }
inline vec3 normalize(const vec3& a) { return a * inversesqrt(dot(a, a)); }
inline vec4 normalize(const vec4& a) { return a * inversesqrt(dot(a, a)); }
inline vec3 cross(const vec3& a, const vec3& b) {
  return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

inline uint packUnorm4x8(const vec4& v) { /* GLSL 4.50 section 8.4: round(clamp(c, 0, 1) * 255) */
  uint r = 0;
  for (int i = 0; i < 4; ++i) r |= ((uint)::lrintf(clamp(v[i], 0.0f, 1.0f) * 255.0f) & 0xFFu) << (8 * i);
  return r;
}
inline vec4 unpackUnorm4x8(uint p) {
  return vec4((float)(p & 0xFFu) / 255.0f, (float)((p >> 8) & 0xFFu) / 255.0f, (float)((p >> 16) & 0xFFu) / 255.0f,
              (float)((p >> 24) & 0xFFu) / 255.0f);
}

/* ---------------------------------------------------------------- matrices (column-major, as GLSL) */
struct mat3 {
  vec3 c[3];
  mat3() {}
  mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
  vec3& operator[](int i) { return c[i]; }
  const vec3& operator[](int i) const { return c[i]; }
};
struct mat4 {
  vec4 c[4];
  vec4& operator[](int i) { return c[i]; }
  const vec4& operator[](int i) const { return c[i]; }
};
/* v * M: component j = dot(v, column j) */
inline vec3 operator*(const vec3& v, const mat3& m) { return vec3(dot(v, m[0]), dot(v, m[1]), dot(v, m[2])); }
inline vec4 operator*(const vec4& v, const mat4& m) { return vec4(dot(v, m[0]), dot(v, m[1]), dot(v, m[2]), dot(v, m[3])); }
inline vec3 operator*(const mat3& m, const vec3& v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z; }
inline vec4 operator*(const mat4& m, const vec4& v) { return ((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3] * v.w; }

/* ---------------------------------------------------------------- value formats */
inline float half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else {
      int e = -1;
      do { e++; man <<= 1; } while ((man & 0x400u) == 0);
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
    }
  } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
  else bits = sign | ((exp + 112) << 23) | (man << 13);
  float f;
  std::memcpy(&f, &bits, 4);
  return f;
}
inline uint16_t float_to_half(float f) { /* round to nearest even */
  uint32_t x;
  std::memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u, absx = x & 0x7fffffffu;
  if (absx >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((absx > 0x7f800000u) ? 0x200u : 0u));
  if (absx >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);
  if (absx < 0x33000001u) return (uint16_t)sign;
  int e = (int)(absx >> 23) - 127;
  uint32_t man = (absx & 0x7fffffu) | 0x800000u;
  int shift = e < -14 ? 13 + (-14 - e) : 13;
  uint32_t hexp = e < -14 ? 0u : (uint32_t)(e + 15);
  uint32_t q = man >> shift, rem = man & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (q & 1u))) q++;
  uint32_t out = hexp == 0 ? q : ((hexp << 10) + (q - 0x400u));
  return (uint16_t)(sign | out);
}
inline float srgb8_to_linear(uint8_t v) { /* the sRGB EOTF of the GL spec, evaluated in double, rounded once */
  double c = (double)v / 255.0;
  return (float)((c <= 0.04045) ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
}
inline uint8_t unorm8(float v) { /* store conversion: clamp, scale, round to nearest */
  float c = clamp(v, 0.0f, 1.0f) * 255.0f;
  return (uint8_t)::lrintf(c);
}

/* ---------------------------------------------------------------- textures, samplers, images */
enum Format { F_R32F, F_RG16F, F_RGBA16F, F_RG16I, F_SRGB8_A8, F_RG8, F_R8, F_R32UI, F_RGBA32F, F_R11G11B10F };
struct Level { const void* data; int w, h, d; };
struct Texture {
  Format fmt = F_R32F;
  int levels = 0;
  Level lv[16] = {};
  bool linear = false;     /* min/mag filter: linear (else nearest) */
  bool mip_linear = false; /* mip filter: linear (else nearest) */
  void set_level(int l, const void* data, int w, int h, int d = 1) {
    lv[l].data = data; lv[l].w = w; lv[l].h = h; lv[l].d = d;
    if (l + 1 > levels) levels = l + 1;
  }
};
inline vec4 fetch_f(const Texture& t, int l, int x, int y, int z) { /* in-range texel of a float-valued format */
  const Level& L = t.lv[l];
  size_t i = (size_t)x + (size_t)L.w * ((size_t)y + (size_t)L.h * (size_t)z);
  switch (t.fmt) {
    case F_R32F: return vec4(((const float*)L.data)[i], 0.0f, 0.0f, 1.0f);
    case F_RGBA32F: { const float* p = (const float*)L.data + i * 4; return vec4(p[0], p[1], p[2], p[3]); }
    case F_RG16F: { const uint16_t* p = (const uint16_t*)L.data + i * 2; return vec4(half_to_float(p[0]), half_to_float(p[1]), 0.0f, 1.0f); }
    case F_RGBA16F: { const uint16_t* p = (const uint16_t*)L.data + i * 4;
                      return vec4(half_to_float(p[0]), half_to_float(p[1]), half_to_float(p[2]), half_to_float(p[3])); }
    case F_SRGB8_A8: { const uint8_t* p = (const uint8_t*)L.data + i * 4;
                       return vec4(srgb8_to_linear(p[0]), srgb8_to_linear(p[1]), srgb8_to_linear(p[2]), (float)p[3] / 255.0f); }
    case F_RG8: { const uint8_t* p = (const uint8_t*)L.data + i * 2; return vec4((float)p[0] / 255.0f, (float)p[1] / 255.0f, 0.0f, 1.0f); }
    case F_R8: return vec4((float)((const uint8_t*)L.data)[i] / 255.0f, 0.0f, 0.0f, 1.0f);
    case F_R11G11B10F: { float r, g, b; drv_unpack_r11g11b10(((const uint32_t*)L.data)[i], &r, &g, &b); return vec4(r, g, b, 1.0f); }
    default: return vec4(0.0f);
  }
}
inline ivec4 fetch_i(const Texture& t, int l, int x, int y, int z) {
  const Level& L = t.lv[l];
  size_t i = (size_t)x + (size_t)L.w * ((size_t)y + (size_t)L.h * (size_t)z);
  const int16_t* p = (const int16_t*)L.data + i * 2; /* F_RG16I */
  return ivec4((int)p[0], (int)p[1], 0, 1);
}
inline uvec4 fetch_u(const Texture& t, int l, int x, int y, int z) {
  const Level& L = t.lv[l];
  size_t i = (size_t)x + (size_t)L.w * ((size_t)y + (size_t)L.h * (size_t)z);
  return uvec4(((const uint32_t*)L.data)[i], 0u, 0u, 1u); /* F_R32UI */
}
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct sampler2D { const Texture* t = nullptr; };
struct isampler2D { const Texture* t = nullptr; };
struct sampler3D { const Texture* t = nullptr; };
struct usampler3D { const Texture* t = nullptr; };

/* one level, 2-D, clamp to edge (spec 8.14.2-3) */
inline vec4 sample_level_2d(const Texture& t, int l, const vec2& uv) {
  const Level& L = t.lv[l];
  if (!t.linear) {
    int x = clampi(f2i(::floorf(uv.x * (float)L.w)), 0, L.w - 1), y = clampi(f2i(::floorf(uv.y * (float)L.h)), 0, L.h - 1);
    return fetch_f(t, l, x, y, 0);
  }
  float fx = uv.x * (float)L.w - 0.5f, fy = uv.y * (float)L.h - 0.5f;
  float flx = ::floorf(fx), fly = ::floorf(fy), tx = fx - flx, ty = fy - fly;
  int x0 = f2i(flx), y0 = f2i(fly);
  int x1 = clampi(x0 + 1, 0, L.w - 1), y1 = clampi(y0 + 1, 0, L.h - 1);
  x0 = clampi(x0, 0, L.w - 1); y0 = clampi(y0, 0, L.h - 1);
  vec4 a = mix(fetch_f(t, l, x0, y0, 0), fetch_f(t, l, x1, y0, 0), tx);
  vec4 b = mix(fetch_f(t, l, x0, y1, 0), fetch_f(t, l, x1, y1, 0), tx);
  return mix(a, b, ty);
}
inline vec4 sample_level_3d(const Texture& t, int l, const vec3& p) {
  const Level& L = t.lv[l];
  if (!t.linear) {
    int x = clampi(f2i(::floorf(p.x * (float)L.w)), 0, L.w - 1), y = clampi(f2i(::floorf(p.y * (float)L.h)), 0, L.h - 1),
        z = clampi(f2i(::floorf(p.z * (float)L.d)), 0, L.d - 1);
    return fetch_f(t, l, x, y, z);
  }
  float fx = p.x * (float)L.w - 0.5f, fy = p.y * (float)L.h - 0.5f, fz = p.z * (float)L.d - 0.5f;
  float flx = ::floorf(fx), fly = ::floorf(fy), flz = ::floorf(fz), tx = fx - flx, ty = fy - fly, tz = fz - flz;
  int x0 = f2i(flx), y0 = f2i(fly), z0 = f2i(flz);
  int x1 = clampi(x0 + 1, 0, L.w - 1), y1 = clampi(y0 + 1, 0, L.h - 1), z1 = clampi(z0 + 1, 0, L.d - 1);
  x0 = clampi(x0, 0, L.w - 1); y0 = clampi(y0, 0, L.h - 1); z0 = clampi(z0, 0, L.d - 1);
  vec4 c00 = mix(fetch_f(t, l, x0, y0, z0), fetch_f(t, l, x1, y0, z0), tx);
  vec4 c10 = mix(fetch_f(t, l, x0, y1, z0), fetch_f(t, l, x1, y1, z0), tx);
  vec4 c01 = mix(fetch_f(t, l, x0, y0, z1), fetch_f(t, l, x1, y0, z1), tx);
  vec4 c11 = mix(fetch_f(t, l, x0, y1, z1), fetch_f(t, l, x1, y1, z1), tx);
  return mix(mix(c00, c10, ty), mix(c01, c11, ty), tz);
}
/* level-of-detail selection for an explicit lod: clamp to the chain, then mip filter */
template <class F>
inline vec4 with_lod(const Texture& t, float lod, F level_fn) {
  float maxLod = (float)(t.levels - 1);
  if (!(lod > 0.0f)) lod = 0.0f; /* also NaN / -inf */
  if (lod > maxLod) lod = maxLod;
  if (!t.mip_linear) return level_fn(f2i(::floorf(lod + 0.5f)));
  float fl = ::floorf(lod), f = lod - fl;
  int l0 = (int)fl, l1 = std::min(l0 + 1, t.levels - 1);
  vec4 a = level_fn(l0);
  if (f == 0.0f) return a;
  return mix(a, level_fn(l1), f);
}
inline vec4 textureLod(const sampler2D& s, const vec2& uv, float lod) {
  return with_lod(*s.t, lod, [&](int l) { return sample_level_2d(*s.t, l, uv); });
}
inline vec4 texture(const sampler2D& s, const vec2& uv) { return textureLod(s, uv, 0.0f); } /* screen-aligned, no mips */
inline vec4 textureLod(const sampler3D& s, const vec3& p, float lod) {
  return with_lod(*s.t, lod, [&](int l) { return sample_level_3d(*s.t, l, p); });
}
inline ivec4 textureLod(const isampler2D& s, const vec2& uv, float lod) { /* integer textures: nearest only */
  const Level& L = s.t->lv[0];
  int x = clampi(f2i(::floorf(uv.x * (float)L.w)), 0, L.w - 1), y = clampi(f2i(::floorf(uv.y * (float)L.h)), 0, L.h - 1);
  return fetch_i(*s.t, 0, x, y, 0);
}
inline ivec4 texture(const isampler2D& s, const vec2& uv) { return textureLod(s, uv, 0.0f); }
inline vec4 texelFetch(const sampler2D& s, const ivec2& p, int l) {
  const Level& L = s.t->lv[l];
  if (p.x < 0 || p.y < 0 || p.x >= L.w || p.y >= L.h) return vec4(0.0f);
  return fetch_f(*s.t, l, p.x, p.y, 0);
}
inline vec4 texelFetch(const sampler3D& s, const ivec3& p, int l) {
  const Level& L = s.t->lv[l];
  if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= L.w || p.y >= L.h || p.z >= L.d) return vec4(0.0f);
  return fetch_f(*s.t, l, p.x, p.y, p.z);
}
inline uvec4 texelFetch(const usampler3D& s, const ivec3& p, int l) { /* out of range: zeros (robust access) */
  const Level& L = s.t->lv[l];
  if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= L.w || p.y >= L.h || p.z >= L.d) return uvec4(0u);
  return fetch_u(*s.t, l, p.x, p.y, p.z);
}
inline ivec3 textureSize(const sampler3D& s, int l) { return ivec3(s.t->lv[l].w, s.t->lv[l].h, s.t->lv[l].d); }
inline ivec2 textureSize(const sampler2D& s, int l) { return ivec2(s.t->lv[l].w, s.t->lv[l].h); }
/* textureGather (spec 8.14.5... "Texture Gather"): the 2x2 footprint of linear filtering, clamp to edge, in the
 * order (i0,j1) (i1,j1) (i1,j0) (i0,j0) */
struct GatherTaps { int x0, x1, y0, y1; };
inline GatherTaps gather_taps(const Level& L, const vec2& uv) {
  int x0 = f2i(::floorf(uv.x * (float)L.w - 0.5f)), y0 = f2i(::floorf(uv.y * (float)L.h - 0.5f));
  GatherTaps g = {clampi(x0, 0, L.w - 1), clampi(x0 + 1, 0, L.w - 1), clampi(y0, 0, L.h - 1), clampi(y0 + 1, 0, L.h - 1)};
  return g;
}
inline vec4 textureGather(const sampler2D& s, const vec2& uv, int comp) {
  GatherTaps g = gather_taps(s.t->lv[0], uv);
  return vec4(fetch_f(*s.t, 0, g.x0, g.y1, 0)[comp], fetch_f(*s.t, 0, g.x1, g.y1, 0)[comp], fetch_f(*s.t, 0, g.x1, g.y0, 0)[comp],
              fetch_f(*s.t, 0, g.x0, g.y0, 0)[comp]);
}
inline ivec4 textureGather(const isampler2D& s, const vec2& uv, int comp) {
  GatherTaps g = gather_taps(s.t->lv[0], uv);
  return ivec4(fetch_i(*s.t, 0, g.x0, g.y1, 0)[comp], fetch_i(*s.t, 0, g.x1, g.y1, 0)[comp], fetch_i(*s.t, 0, g.x1, g.y0, 0)[comp],
               fetch_i(*s.t, 0, g.x0, g.y0, 0)[comp]);
}

/* images: one level, load/store/atomics; out-of-range accesses are dropped / return 0 */
struct image3D { void* data = nullptr; int w = 0, h = 0, d = 0; Format fmt = F_R8; };
struct uimage3D { uint32_t* data = nullptr; int w = 0, h = 0, d = 0; };
struct image2D { uint32_t* data = nullptr; int w = 0, h = 0; }; /* r11f_g11f_b10f texels (include/drv_r11g11b10.h) */
inline ivec2 imageSize(const image2D& im) { return ivec2(im.w, im.h); }
inline vec4 imageLoad(const image2D& im, const ivec2& p) {
  if (p.x < 0 || p.y < 0 || p.x >= im.w || p.y >= im.h) return vec4(0.0f);
  float r, g, b;
  drv_unpack_r11g11b10(im.data[(size_t)p.x + (size_t)im.w * (size_t)p.y], &r, &g, &b);
  return vec4(r, g, b, 1.0f);
}
inline void imageStore(const image2D& im, const ivec2& p, const vec4& v) {
  if (p.x < 0 || p.y < 0 || p.x >= im.w || p.y >= im.h) return;
  im.data[(size_t)p.x + (size_t)im.w * (size_t)p.y] = drv_pack_r11g11b10(v.x, v.y, v.z);
}
inline bool in_range(int w, int h, int d, const ivec3& p) { return p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < w && p.y < h && p.z < d; }
inline vec4 imageLoad(const image3D& im, const ivec3& p) {
  if (!in_range(im.w, im.h, im.d, p)) return vec4(0.0f);
  return vec4((float)((const uint8_t*)im.data)[(size_t)p.x + (size_t)im.w * ((size_t)p.y + (size_t)im.h * (size_t)p.z)] / 255.0f, 0.0f, 0.0f, 1.0f);
}
inline void imageStore(const image3D& im, const ivec3& p, const vec4& v) {
  if (!in_range(im.w, im.h, im.d, p)) return;
  ((uint8_t*)im.data)[(size_t)p.x + (size_t)im.w * ((size_t)p.y + (size_t)im.h * (size_t)p.z)] = unorm8(v.x);
}
inline uint imageAtomicCompSwap(const uimage3D& im, const ivec3& p, uint compare, uint value) {
  if (!in_range(im.w, im.h, im.d, p)) return 0xFFFFFFFEu; /* dropped: never equals `compare` == 0 (SURVEY B.3 policy) */
  uint32_t* a = im.data + ((size_t)p.x + (size_t)im.w * ((size_t)p.y + (size_t)im.h * (size_t)p.z));
  uint32_t expected = compare;
  __atomic_compare_exchange_n(a, &expected, value, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return expected;
}
inline void imageStore(const uimage3D& im, const ivec3& p, const uvec4& v) {
  if (!in_range(im.w, im.h, im.d, p)) return;
  __atomic_store_n(im.data + ((size_t)p.x + (size_t)im.w * ((size_t)p.y + (size_t)im.h * (size_t)p.z)), v.x, __ATOMIC_SEQ_CST);
}
template <class T, class V>
inline T atomicAdd(T& mem, V v) { return __atomic_fetch_add(&mem, (T)v, __ATOMIC_SEQ_CST); }
inline void memoryBarrierImage() {}
inline void memoryBarrier() {}

/* shader storage block with an unsized array: out-of-range reads return zeros and writes are dropped (robust
 * buffer access — what cacheApply.frag:91-100 relies on for an unallocated neighbour) */
template <class T>
struct ssbo {
  T* data = nullptr;
  size_t count = 0;
  void bind(void* p, size_t n) { data = (T*)p; count = n; }
  T& operator[](size_t i) const {
    if (i < count) return data[i];
    static thread_local T dummy;
    dummy = T();
    return dummy;
  }
};

/* ---------------------------------------------------------------- work groups as fibers */
struct FiberGroup {
  ucontext_t main_ctx;
  std::vector<ucontext_t> ctx;
  std::vector<char*> stacks;
  std::vector<char> done;
  std::function<void(int)> body;
  int current = -1;
  static constexpr size_t kStack = 256 * 1024;
  ~FiberGroup() { for (char* s : stacks) std::free(s); }
};
inline FiberGroup*& current_group() { static thread_local FiberGroup* g = nullptr; return g; }
inline void fiber_trampoline() {
  FiberGroup* g = current_group();
  int i = g->current;
  g->body(i);
  g->done[i] = 1;
  swapcontext(&g->ctx[i], &g->main_ctx);
}
/* GLSL barrier(): park this invocation until every other live invocation of the group has arrived */
inline void barrier() {
  FiberGroup* g = current_group();
  swapcontext(&g->ctx[g->current], &g->main_ctx);
}
/* Runs `n` invocations of one work group. `body(i)` is invocation i; `on_barrier(k)` (optional) is called
 * after every invocation has reached its k-th barrier — the harness uses it to read shared memory. */
inline void run_group(FiberGroup& g, int n, const std::function<void(int)>& body, const std::function<void(int)>* on_barrier = nullptr) {
  if ((int)g.ctx.size() < n) {
    g.ctx.resize(n);
    while ((int)g.stacks.size() < n) g.stacks.push_back((char*)std::malloc(FiberGroup::kStack));
  }
  g.done.assign(n, 0);
  g.body = body;
  FiberGroup* prev = current_group();
  current_group() = &g;
  for (int i = 0; i < n; ++i) {
    getcontext(&g.ctx[i]);
    g.ctx[i].uc_stack.ss_sp = g.stacks[i];
    g.ctx[i].uc_stack.ss_size = FiberGroup::kStack;
    g.ctx[i].uc_link = &g.main_ctx;
    makecontext(&g.ctx[i], fiber_trampoline, 0);
  }
  int barrier_index = 0;
  for (;;) {
    int live = 0;
    for (int i = 0; i < n; ++i) {
      if (g.done[i]) continue;
      g.current = i;
      swapcontext(&g.main_ctx, &g.ctx[i]);
      if (!g.done[i]) ++live;
    }
    if (live == 0) break;
    if (on_barrier) (*on_barrier)(barrier_index);
    ++barrier_index;
  }
  current_group() = prev;
}

/* Built-in variables of a compute / fragment invocation; the generated shader struct derives from this. */
struct Invocation {
  uvec3 gl_GlobalInvocationID, gl_LocalInvocationID, gl_WorkGroupID, gl_NumWorkGroups;
  uint gl_LocalInvocationIndex = 0;
  vec4 gl_FragCoord;
  bool gl_Discarded = false;
};

inline int hw_threads() {
  unsigned n = std::thread::hardware_concurrency();
  return n == 0 ? 1 : (int)n;
}
/* static partition of [0, n) over OS threads */
inline void parallel_for(int64_t n, int threads, const std::function<void(int64_t, int64_t)>& fn) {
  if (threads <= 0) threads = hw_threads();
  if (threads > n) threads = (int)std::max<int64_t>(n, 1);
  if (threads <= 1) { fn(0, n); return; }
  std::vector<std::thread> pool;
  int64_t chunk = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    int64_t b = t * chunk, e = std::min<int64_t>(n, b + chunk);
    if (b >= e) break;
    pool.emplace_back(fn, b, e);
  }
  for (auto& th : pool) th.join();
}

/* Dispatch of a compute shader type S with local size (lx, ly, lz) over (gx, gy, gz) groups. Every invocation is
 * its own S object (plain globals of the shader are per-invocation members; uniforms, buffers and `shared`
 * variables are static members). Groups run in parallel on OS threads when `threads` != 1. */
template <class S>
inline void dispatch(int gx, int gy, int gz, int threads = 0,
                     const std::function<void(int group, int barrier_index)>* tap = nullptr) {
  const int lx = S::gl_WorkGroupSize_x, ly = S::gl_WorkGroupSize_y, lz = S::gl_WorkGroupSize_z, n = lx * ly * lz;
  parallel_for((int64_t)gx * gy * gz, threads, [&](int64_t b, int64_t e) {
    FiberGroup fg;
    std::vector<S> inv(n);
    for (int64_t grp = b; grp < e; ++grp) {
      const int wx = (int)(grp % gx), wy = (int)((grp / gx) % gy), wz = (int)(grp / ((int64_t)gx * gy));
      for (int i = 0; i < n; ++i) {
        inv[i] = S();
        const int ix = i % lx, iy = (i / lx) % ly, iz = i / (lx * ly);
        inv[i].gl_LocalInvocationID = uvec3(ix, iy, iz);
        inv[i].gl_WorkGroupID = uvec3(wx, wy, wz);
        inv[i].gl_NumWorkGroups = uvec3(gx, gy, gz);
        inv[i].gl_GlobalInvocationID = uvec3(wx * lx + ix, wy * ly + iy, wz * lz + iz);
        inv[i].gl_LocalInvocationIndex = (uint)i;
      }
      std::function<void(int)> body = [&](int i) { inv[i].main(); };
      if (tap) {
        std::function<void(int)> cb = [&](int k) { (*tap)((int)grp, k); };
        run_group(fg, n, body, &cb);
      } else {
        run_group(fg, n, body, nullptr);
      }
    }
  });
}

} // namespace glsl
#endif
