// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Minimal OpenCL-C 1.x shim so that the *verbatim* reference kernel
// (/root/reference/Source/kernels.cl) can be compiled by g++ as C++17 after
// the mechanical token rewrites done by oracle/build_ref.py.  Only what that
// one file uses is provided.  All arithmetic is plain IEEE-754 binary32 in
// source order (build with -O2 -ffp-contract=off, never -ffast-math):
//   native_recip(x)   = 1.0f / x
//   native_divide(a,b)= a / b
//   native_sqrt(x)    = sqrtf(x)
//   normalize(v)      = v * (1.0f / sqrtf(dot(v,v)))
//   dot(a,b)          = (a.x*b.x + a.y*b.y) + a.z*b.z
//   min(x,y)          = y < x ? y : x      (OpenCL 1.2 spec 6.12.4)
//   max(x,y)          = x < y ? y : x
// The C restatement (oracle/cornell_oracle.c) and the CUDA strict-IEEE kernel
// (uob_raytracer_b200/csrc) follow exactly these definitions so that all three
// are bit-identical.
#pragma once
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace refcl {

typedef unsigned int uint;

struct alignas(16) float3 {
  float x, y, z, _pad;
};
struct float4;
struct alignas(16) uint3 {
  uint x, y, z, _pad;
};

static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z, 0.0f}; }
static inline float3 make_float3(float s) { return float3{s, s, s, 0.0f}; }

struct alignas(16) float4 {
  float x, y, z, w;
  float3 xyz() const { return make_float3(x, y, z); }
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

// (uint3)(a, b, c): each scalar is converted with the usual C conversion
// (float -> uint truncates toward zero; values stay < 2^32 for every config).
template <class A, class B, class C>
static inline uint3 make_uint3(A a, B b, C c) {
  return uint3{(uint)a, (uint)b, (uint)c, 0u};
}

// ---- float3 arithmetic (component-wise; scalars widen to float first) ------
static inline float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float3 operator/(float3 a, float3 b) { return make_float3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline float3 operator-(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
template <class S> static inline float3 operator*(S s, float3 a) { float f = (float)s; return make_float3(f * a.x, f * a.y, f * a.z); }
template <class S> static inline float3 operator*(float3 a, S s) { float f = (float)s; return make_float3(a.x * f, a.y * f, a.z * f); }
template <class S> static inline float3 operator/(float3 a, S s) { float f = (float)s; return make_float3(a.x / f, a.y / f, a.z / f); }
template <class S> static inline float3 operator-(float3 a, S s) { float f = (float)s; return make_float3(a.x - f, a.y - f, a.z - f); }
template <class S> static inline float3 operator+(float3 a, S s) { float f = (float)s; return make_float3(a.x + f, a.y + f, a.z + f); }
static inline float3 &operator+=(float3 &a, float3 b) { a = a + b; return a; }
template <class S> static inline float3 &operator*=(float3 &a, S s) { a = a * s; return a; }

static inline float dot(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float3 normalize(float3 v) {
  const float inv = 1.0f / sqrtf(dot(v, v));
  return make_float3(v.x * inv, v.y * inv, v.z * inv);
}

// Scalar float overloads: without these, ::sqrt/::fabs would promote to double.
static inline float sqrt(float x) { return sqrtf(x); }
static inline float fabs(float x) { return fabsf(x); }
static inline float min(float x, float y) { return y < x ? y : x; }
static inline float max(float x, float y) { return x < y ? y : x; }
static inline float3 min(float3 a, float s) { return make_float3(min(a.x, s), min(a.y, s), min(a.z, s)); }
static inline float3 max(float3 a, float s) { return make_float3(max(a.x, s), max(a.y, s), max(a.z, s)); }

static inline float native_recip(float x) { return 1.0f / x; }
static inline float native_divide(float a, float b) { return a / b; }
static inline float native_sqrt(float x) { return sqrtf(x); }

// ---- uint3 (xorshift) ------------------------------------------------------
static inline uint3 operator<<(uint3 a, int s) { return uint3{a.x << s, a.y << s, a.z << s, 0u}; }
static inline uint3 operator>>(uint3 a, int s) { return uint3{a.x >> s, a.y >> s, a.z >> s, 0u}; }
static inline uint3 &operator^=(uint3 &a, uint3 b) { a.x ^= b.x; a.y ^= b.y; a.z ^= b.z; return a; }

static inline float3 convert_float3(uint3 v) { return make_float3((float)v.x, (float)v.y, (float)v.z); }
static inline uint3 convert_uint3(float3 v) { return uint3{(uint)v.x, (uint)v.y, (uint)v.z, 0u}; }

// ---- address spaces, work-item functions, async copies ---------------------
#define constant static const
#define global
#define local
#define kernel
#ifndef MAXFLOAT
#define MAXFLOAT FLT_MAX
#endif

typedef int event_t;
static thread_local int shim_global_id[2];
static float shim_screen_w = 1024.0f;  // SCREEN_WIDTH / SCREEN_HEIGHT become run-time
static float shim_screen_h = 1024.0f;
static inline int get_global_id(int dim) { return shim_global_id[dim]; }
template <class T>
static inline event_t async_work_group_copy(T *dst, const T *src, size_t n, event_t) {
  if (dst != src) memcpy((void *)dst, (const void *)src, n * sizeof(T));
  return 0;
}
static inline void wait_group_events(int, event_t *) {}

}  // namespace refcl
