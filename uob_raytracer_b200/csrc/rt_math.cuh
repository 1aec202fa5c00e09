// rt_math.cuh — the two arithmetic policies of the render path.
//
//   float   : the fast path.  Plain operators; nvcc contracts a*b+c into FFMA,
//             reciprocals / rsqrt / sqrt are single MUFU ops.
//   sfloat  : RT_FLAG_STRICT_IEEE.  Every operation is an explicitly rounded
//             intrinsic (__fmul_rn, __fadd_rn, ...) which the compiler never
//             fuses, so a kernel instantiated with sfloat evaluates exactly the
//             operation sequence of the reference kernel under IEEE-754 rules
//             (Source/kernels.cl; conventions of oracle/cl_shim.h) and is
//             bit-identical to the CPU oracle.
//
// All device code is written once, templated on the number type T.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rt {

struct sfloat {
  float v;
  __device__ __forceinline__ sfloat() {}
  __device__ __forceinline__ sfloat(float x) : v(x) {}
};

__device__ __forceinline__ sfloat operator+(sfloat a, sfloat b) { return sfloat(__fadd_rn(a.v, b.v)); }
__device__ __forceinline__ sfloat operator-(sfloat a, sfloat b) { return sfloat(__fsub_rn(a.v, b.v)); }
__device__ __forceinline__ sfloat operator*(sfloat a, sfloat b) { return sfloat(__fmul_rn(a.v, b.v)); }
__device__ __forceinline__ sfloat operator-(sfloat a) { return sfloat(-a.v); }
__device__ __forceinline__ bool operator<(sfloat a, sfloat b) { return a.v < b.v; }
__device__ __forceinline__ bool operator<=(sfloat a, sfloat b) { return a.v <= b.v; }
__device__ __forceinline__ bool operator>(sfloat a, sfloat b) { return a.v > b.v; }
__device__ __forceinline__ bool operator>=(sfloat a, sfloat b) { return a.v >= b.v; }
__device__ __forceinline__ bool operator==(sfloat a, sfloat b) { return a.v == b.v; }

template <class T> struct is_strict { static constexpr bool value = false; };
template <> struct is_strict<sfloat> { static constexpr bool value = true; };

__device__ __forceinline__ float raw(float x) { return x; }
__device__ __forceinline__ float raw(sfloat x) { return x.v; }

// ---- approximate single-instruction ops (fast path only) --------------------
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// native_recip / native_divide / native_sqrt / sqrt of kernels.cl.
// Fast path: IEEE-rounded (used only on the closest-hit path, ~10 % of the
// work, where knife-edge decisions at the default camera are rounding
// sensitive — SURVEY.md §7); the shadow path uses division-free tests.
#ifdef RT_STRICT_NOINLINE  // A/B switch (code size): one out-of-line copy of each correctly rounded sequence per kernel
static __device__ __noinline__ float rt_frcp_rn(float x) { return __frcp_rn(x); }
static __device__ __noinline__ float rt_fdiv_rn(float a, float b) { return __fdiv_rn(a, b); }
static __device__ __noinline__ float rt_fsqrt_rn(float x) { return __fsqrt_rn(x); }
#else
#define rt_frcp_rn __frcp_rn
#define rt_fdiv_rn __fdiv_rn
#define rt_fsqrt_rn __fsqrt_rn
#endif
__device__ __forceinline__ float rcp_(float x) { return __frcp_rn(x); }
__device__ __forceinline__ sfloat rcp_(sfloat x) { return sfloat(rt_frcp_rn(x.v)); }  // rcp.rn: correctly rounded == 1.0f/x
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ sfloat div_(sfloat a, sfloat b) { return sfloat(rt_fdiv_rn(a.v, b.v)); }
__device__ __forceinline__ float sqrt_(float x) { return __fsqrt_rn(x); }
__device__ __forceinline__ sfloat sqrt_(sfloat x) { return sfloat(rt_fsqrt_rn(x.v)); }
__device__ __forceinline__ float abs_(float x) { return fabsf(x); }
__device__ __forceinline__ sfloat abs_(sfloat x) { return sfloat(fabsf(x.v)); }
// OpenCL min/max semantics (oracle/cl_shim.h): min(x,y) = y<x?y:x, max(x,y) = x<y?y:x
template <class T> __device__ __forceinline__ T cl_min(T x, T y) { return (y < x) ? y : x; }
template <class T> __device__ __forceinline__ T cl_max(T x, T y) { return (x < y) ? y : x; }

template <class T> struct V3 {
  T x, y, z;
  __device__ __forceinline__ V3() {}
  __device__ __forceinline__ V3(T a, T b, T c) : x(a), y(b), z(c) {}
};
template <class T> __device__ __forceinline__ V3<T> operator+(V3<T> a, V3<T> b) { return V3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> __device__ __forceinline__ V3<T> operator-(V3<T> a, V3<T> b) { return V3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class T> __device__ __forceinline__ V3<T> operator*(V3<T> a, V3<T> b) { return V3<T>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <class T> __device__ __forceinline__ V3<T> operator-(V3<T> a) { return V3<T>(-a.x, -a.y, -a.z); }
template <class T> __device__ __forceinline__ V3<T> scale(T s, V3<T> a) { return V3<T>(s * a.x, s * a.y, s * a.z); }
// dot = (x*x' + y*y') + z*z'   (oracle/cl_shim.h)
template <class T> __device__ __forceinline__ T dot(V3<T> a, V3<T> b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

// normalize(v) = v * (1/sqrt(dot(v,v)))
__device__ __forceinline__ V3<sfloat> normalize(V3<sfloat> v) {
  const sfloat inv = rcp_(sqrt_(dot(v, v)));
  return V3<sfloat>(v.x * inv, v.y * inv, v.z * inv);
}
__device__ __forceinline__ V3<float> normalize(V3<float> v) {
  const float inv = rcp_(sqrt_(dot(v, v)));
  return V3<float>(v.x * inv, v.y * inv, v.z * inv);
}

template <class T> __device__ __forceinline__ V3<T> xyz(float4 a) { return V3<T>(T(a.x), T(a.y), T(a.z)); }

// kernels.cl:42-47
__device__ __forceinline__ uint32_t xorshift32(uint32_t s) {
  s ^= s << 13;
  s ^= s >> 17;
  s ^= s << 5;
  return s;
}

// kernels.cl:49-52 : range*float(v)/float(UINT_MAX) - range/2, (float)UINT_MAX == 2^32
template <class T> __device__ __forceinline__ T crush1(uint32_t v, float range) {
  if constexpr (is_strict<T>::value) {
    return div_(T(range) * T(__uint2float_rn(v)), T(4294967296.0f)) - T(range / 2.f);
  } else {
    // division by 2^32 is an exact scaling, so this is the same value with one rounding less
    return fmaf(__uint2float_rn(v), range * 2.3283064365386963e-10f, -(range * 0.5f));  // (spelled out: every copy rounds alike)
  }
}

}  // namespace rt
