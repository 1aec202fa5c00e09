// rt_brute.cuh — the generic render path: device equivalents of the helpers of
// Source/kernels.cl, written once for both arithmetic policies (rt_math.cuh) and
// for both ways of finding candidate triangles (a Tracer: brute force over a
// small scene in shared memory here, BVH traversal in rt_bvh.cuh).
//
// What differs from the reference's loop structure, and why it does not change
// results:
//   * per-triangle constants (e1 = v1-v0, e2 = v2-v0 and the three cofactors of
//     the first row of det[., e1, e2]) are precomputed at upload with the same
//     single-rounded operations the kernel would perform (rt_api.cu);
//   * the any-hit loop (in_shadow, kernels.cl:243-311) runs triangle-outer /
//     shadow-sample-inner: everything that depends only on (start, triangle) —
//     b, det[b,e1,e2] and the cofactors of det[-d,b,e2], det[-d,e1,b] — is
//     computed once per triangle instead of once per sample.  in_shadow returns
//     a boolean OR over occluders, so the order of evaluation is immaterial;
//   * closest hit keeps (t, u, v, index) of the winner and forms the hit point
//     after the loop; with a BVH the triangles arrive in arbitrary order, so the
//     reference's "strict <, ascending index" rule (lowest index wins an exact tie)
//     is applied explicitly;
//   * the fast policy decides t >= 0, |t d|^2 < r^2, u >= 0, v >= 0, u+v <= 1
//     without dividing (the sign of det A is carried instead).
#pragma once
#include "rt_math.cuh"
#include "rt_types.h"

namespace rt {

// Analytic spheres baked into the reference kernel (kernels.cl:7-10): centre.xyz,
// radius^2 in .w; colour.w is the material (-1 glass, 0 mirror).
#define RT_SPHERES 2
static __constant__ float4 c_sphere_center_r2[RT_SPHERES] = {{0.3f, 0.1f, -0.5f, 0.075f}, {-0.4f, 0.8f, -0.5f, 0.05f}};
// host-side copy (tile classification of mixed launches, rt_api.cu)
static const float kSphereCenterR2[RT_SPHERES][4] = {{0.3f, 0.1f, -0.5f, 0.075f}, {-0.4f, 0.8f, -0.5f, 0.05f}};
static __constant__ float4 c_sphere_color[RT_SPHERES] = {{0.0f, 0.0f, 0.0f, -1.0f}, {0.0f, 0.0f, 0.0f, 0.0f}};

#define RT_GLASS 1.52f
#define RT_AIR 1.0f
#define RT_BIAS 0.0001f
#define RT_INDIRECT 0.5f
#define RT_LIGHT_COLOR 16.0f
#define RT_LIGHT_SPREAD 0.05f
#define RT_PI_F 3.14159265358979323846f
#define RT_MAXFLOAT 3.402823466e+38f

// Scene as the brute-force kernels read it.  Closest-hit arrays hold all n
// triangles in upload order; the shadow arrays hold only shadow casters
// (material != -1, kernels.cl:247), also in upload order.
//   ta[i] = (v0.xyz, c0)   tb[i] = (e1.xyz, c1)   tc[i] = (e2.xyz, c2)
//   with c0 = e1.y*e2.z - e1.z*e2.y, c1 = e1.x*e2.z - e1.z*e2.x, c2 = e1.x*e2.y - e1.y*e2.x
//   so that det[m, e1, e2] = (m.x*c0 - m.y*c1) + m.z*c2   (kernels.cl:31-35)
struct SceneView {
  const float4 *ta, *tb, *tc;  // closest-hit triangles
  const float4 *tn, *tcol;     // normals, colours (w = material)
  const float4 *sa, *sb, *sc;  // shadow casters
  const float4 *tnd;           // fast kernel only: (N.xyz, v0.N) per triangle, N = e1 x e2 (rt_fast.cuh: closest_hit_bounce)
  int n, n_sh;
};

template <class T> struct HitRec {
  int id;  // -1 none, -2 sphere, >=0 triangle (kernels.cl:28)
  V3<T> point, normal;
  float4 color;
};

// Running state of a closest-hit search over triangles.
template <class T> struct ClosestState {
  T t, u, v;
  int id;     // index in the caller's numbering (upload order), -1 = none
  int slot;   // where the caller finds the winner's data (== id for brute force)
  __device__ __forceinline__ void reset() {
    t = T(RT_MAXFLOAT);
    u = T(0.0f);
    v = T(0.0f);
    id = -1;
    slot = -1;
  }
};

// One ray/triangle test of the closest-hit search (kernels.cl:102-128).  ORDERED: the caller visits
// triangles in ascending index, so `t < current` alone implements the tie rule.  PRECISE: the fast
// policy also uses the reference's cofactor expressions (with FMA) instead of the cheaper triple
// products — for the sub-pixel triangles of a mesh, where u and v are differences of nearly equal
// numbers and the cheaper form misplaces visibly more edge hits (measured: 21 vs 5 pixels of a 96x96
// frame over a 5,120-triangle mesh).
template <class T, bool ORDERED, bool PRECISE = false>
__device__ __forceinline__ void closest_tri_test(float4 A, float4 Bq, float4 C, V3<T> start, V3<T> nd, int id, int slot,
                                                 ClosestState<T> &cs) {
  const V3<T> v0 = xyz<T>(A), e1 = xyz<T>(Bq), e2 = xyz<T>(C);
  const T c0 = T(A.w), c1 = T(Bq.w), c2 = T(C.w);
  const V3<T> b = start - v0;
  if constexpr (is_strict<T>::value || PRECISE) {
    const T detA = (nd.x * c0 - nd.y * c1) + nd.z * c2;
    const T inv = rcp_(detA);
    const T t = ((b.x * c0 - b.y * c1) + b.z * c2) * inv;
    const T u = ((nd.x * (b.y * e2.z - b.z * e2.y) - nd.y * (b.x * e2.z - b.z * e2.x)) + nd.z * (b.x * e2.y - b.y * e2.x)) * inv;
    const T v = ((nd.x * (e1.y * b.z - e1.z * b.y) - nd.y * (e1.x * b.z - e1.z * b.x)) + nd.z * (e1.x * b.y - e1.y * b.x)) * inv;
    const bool closer = ORDERED ? (t < cs.t) : ((t < cs.t) || (t == cs.t && id < cs.id));
    if (closer && u >= T(0.0f) && v >= T(0.0f) && (u + v) <= T(1.0f) && t >= T(0.0f)) {
      cs.id = id;
      cs.slot = slot;
      cs.u = u;
      cs.v = v;
      cs.t = t;
    }
  } else {
    // Fast policy: the same quantities through scalar triple products, deciding the inside test on
    // numerators and dividing only for a triangle that passes it.  With d = -nd, q = b x d:
    //   det A = -d.N = -dn,  det[b,e1,e2] = b.N,  det[-d,b,e2] = e2.q,  det[-d,e1,b] = -e1.q
    //   t = -(b.N)/dn,  u = -(e2.q)/dn,  v = (e1.q)/dn
    const V3<T> d = -nd;
    const float dn = (d.x * c0 - d.y * c1) + d.z * c2;
    const float bn = (b.x * c0 - b.y * c1) + b.z * c2;
    const V3<T> q(b.y * d.z - b.z * d.y, b.z * d.x - b.x * d.z, b.x * d.y - b.y * d.x);
    const float eu = -dot(e2, q), ev = dot(e1, q);
    const unsigned sb = __float_as_uint(dn) & 0x80000000u;
    const float us = __uint_as_float(__float_as_uint(eu) ^ sb), vs = __uint_as_float(__float_as_uint(ev) ^ sb),
                ts = __uint_as_float(__float_as_uint(-bn) ^ sb), adn = fabsf(dn);
    if ((us >= 0.0f) & (vs >= 0.0f) & ((us + vs) <= adn) & (ts >= 0.0f)) {
      const float inv = __frcp_rn(adn);
      const float t = ts * inv;
      const bool closer = ORDERED ? (t < cs.t) : ((t < cs.t) || (t == cs.t && id < cs.id));
      if (closer) {
        cs.id = id;
        cs.slot = slot;
        cs.u = us * inv;
        cs.v = vs * inv;
        cs.t = t;
      }
    }
  }
}

// (v0 + u*e1) + v*e2 of the winning triangle (kernels.cl:124)
template <class T> __device__ __forceinline__ V3<T> hit_point(float4 A, float4 Bq, float4 C, T u, T v) {
  return (xyz<T>(A) + scale(u, xyz<T>(Bq))) + scale(v, xyz<T>(C));
}

// Sphere part of the closest-hit search (kernels.cl:132-163), continuing from current_t.
template <class T>
__device__ __forceinline__ void closest_spheres(V3<T> start, V3<T> dir, T current_t, HitRec<T> &hit) {
#pragma unroll
  for (int i = 0; i < RT_SPHERES; i++) {
    const float4 cr = c_sphere_center_r2[i];
    const V3<T> ctr = xyz<T>(cr);
    const V3<T> L = start - ctr;
    const T a = dot(dir, dir);
    const T b = T(2.0f) * dot(dir, L);
    const T c = dot(L, L) - T(cr.w);
    const T disc = b * b - T(4.0f) * a * c;
    if (disc < T(0.0f)) continue;
    const T sq = sqrt_(disc);
    const T q = (b > T(0.0f)) ? T(-0.5f) * (b + sq) : T(-0.5f) * (b - sq);
    const T x0 = div_(q, a);
    const T x1 = div_(c, q);
    const T x_min = cl_min(x0, x1);
    const T x_max = cl_max(x0, x1);
    T x;
    if (x_min >= T(0.0f) && x_min < current_t) x = x_min;
    else if (x_max >= T(0.0f) && x_max < current_t) x = x_max;
    else continue;
    hit.id = -2;
    hit.point = start + scale(x, dir);
    hit.normal = normalize(hit.point - ctr);
    hit.color = c_sphere_color[i];
    current_t = x;
  }
}

// |d_k|^2 of the CH shadow directions (fast policy only)
// (RT_RAYS_DD_CACHE=1 keeps |d_k|^2 in CH more registers; recomputed at its two uses it is three operations, spelled out so
// that every inlined copy rounds alike — the BVH kernel carries the rays through a traversal with a 64-entry stack and is
// register-bound at 80; see rt_fast.cuh: RT_DD_CACHE for the same change in the brute-force kernel)
#ifndef RT_RAYS_DD_CACHE
#define RT_RAYS_DD_CACHE 0
#endif
template <class T, int CH> struct ShadowRays {
  V3<T> d[CH];
#if RT_RAYS_DD_CACHE
  T dd[CH];
#endif
  __device__ __forceinline__ void finish() {
#if RT_RAYS_DD_CACHE
    if constexpr (!is_strict<T>::value) {
#pragma unroll
      for (int k = 0; k < CH; k++) dd[k] = fmaf(raw(d[k].z), raw(d[k].z), fmaf(raw(d[k].x), raw(d[k].x), __fmul_rn(raw(d[k].y), raw(d[k].y))));
    }
#endif
  }
  __device__ __forceinline__ float dd_of(int k) const {
#if RT_RAYS_DD_CACHE
    return raw(dd[k]);
#else
    return fmaf(raw(d[k].z), raw(d[k].z), fmaf(raw(d[k].x), raw(d[k].x), __fmul_rn(raw(d[k].y), raw(d[k].y))));
#endif
  }
};

// One (origin, triangle) pair of in_shadow (kernels.cl:246-276) against CH rays; ORs hits into occ.
template <class T, int CH>
__device__ __forceinline__ void shadow_pair(float4 A, float4 Bq, float4 C, V3<T> start, const ShadowRays<T, CH> &rays, T radius_sq,
                                            unsigned &occ, unsigned alive = 0xffffffffu) {
  const V3<T> v0 = xyz<T>(A), e1 = xyz<T>(Bq), e2 = xyz<T>(C);
  const T c0 = T(A.w), c1 = T(Bq.w), c2 = T(C.w);
  const V3<T> b = start - v0;
  const T detA0 = (b.x * c0 - b.y * c1) + b.z * c2;
  // cofactors of det[-d, b, e2] and det[-d, e1, b]: independent of the sample
  const T U0 = b.y * e2.z - b.z * e2.y, U1 = b.x * e2.z - b.z * e2.x, U2 = b.x * e2.y - b.y * e2.x;
  const T V0 = e1.y * b.z - e1.z * b.y, V1 = e1.x * b.z - e1.z * b.x, V2 = e1.x * b.y - e1.y * b.x;
  if constexpr (is_strict<T>::value) {
#pragma unroll
    for (int k = 0; k < CH; k++) {
      if (((occ | ~alive) >> k) & 1u) continue;
      const V3<T> nd = -rays.d[k];
      const T detA = (nd.x * c0 - nd.y * c1) + nd.z * c2;
      const T inv = rcp_(detA);
      const T t = detA0 * inv;
      const V3<T> dv = scale(t, rays.d[k]);
      const T dist = dv.x * dv.x + dv.y * dv.y + dv.z * dv.z;
      if (t >= T(0.0f) && dist < radius_sq) {
        const T u = ((nd.x * U0 - nd.y * U1) + nd.z * U2) * inv;
        const T v = ((nd.x * V0 - nd.y * V1) + nd.z * V2) * inv;
        if (u >= T(0.0f) && v >= T(0.0f) && (u + v) <= T(1.0f)) occ |= 1u << k;
      }
    }
  } else {
    const T num2 = detA0 * detA0;
#pragma unroll
    for (int k = 0; k < CH; k++) {
      // den = det[-d, e1, e2];  t = detA0/den
      const T den = -((rays.d[k].x * c0 - rays.d[k].y * c1) + rays.d[k].z * c2);
      // t >= 0  and  t^2 |d|^2 < r^2, without dividing (den == 0 fails the second test)
      const bool s1 = (detA0 * den >= 0.0f) && (num2 * rays.dd_of(k) < radius_sq * (den * den)) && ((alive >> k) & 1u);
      if (s1) {
        const T sg = (den < 0.0f) ? -1.0f : 1.0f;
        const T D1 = -((rays.d[k].x * U0 - rays.d[k].y * U1) + rays.d[k].z * U2) * sg;  // u * |den|
        const T D2 = -((rays.d[k].x * V0 - rays.d[k].y * V1) + rays.d[k].z * V2) * sg;  // v * |den|
        if (D1 >= 0.0f && D2 >= 0.0f && (D1 + D2) <= den * sg) occ |= 1u << k;
      }
    }
  }
}

// jitter components lie in [-0.025, 0.025] (crush, kernels.cl:49-52): |j| <= 0.025*sqrt(3) = 0.0433013
constexpr float kJitterMax = 0.0445f;  // inflated by 2.7 %
constexpr float kSlack = 1.002f;

__device__ __forceinline__ float xor_sign(float v, unsigned signbit) { return __uint_as_float(__float_as_uint(v) ^ signbit); }

// Fast-policy (origin, triangle) pair with the two conservative culls of rt_fast.cuh (plane, edges) in front of
// the division-free per-sample test, for callers that hold the CH ray directions in registers.
// bound = (jmax|N|, jmax|e1|, jmax|e2|, -); r = un-jittered direction; kk >= |r|/|d_k| for every k; inv_r2 = 1/|r|^2.
template <int CH>
__device__ __forceinline__ void shadow_pair_culled(float4 A, float4 Bq, float4 C, float4 bound, V3<float> start, V3<float> r,
                                                   const ShadowRays<float, CH> &rays, float kk, float inv_r2, unsigned &occ,
                                                   unsigned alive = 0xffffffffu) {
  const V3<float> b(start.x - A.x, start.y - A.y, start.z - A.z);
  const float c0 = A.w, c1 = Bq.w, c2 = C.w;
  const float num = (b.x * c0 - b.y * c1) + b.z * c2;  // det[b,e1,e2] = b.N
  const float rN = (r.x * c0 - r.y * c1) + r.z * c2;
  const float w = xor_sign(rN, __float_as_uint(num) & 0x80000000u);
  if (fabsf(num) >= (bound.x - w) * kk) return;  // plane cull: no sample passes stage 1
  const V3<float> e1(Bq.x, Bq.y, Bq.z), e2(C.x, C.y, C.z);
  const V3<float> U(b.y * e2.z - b.z * e2.y, b.z * e2.x - b.x * e2.z, b.x * e2.y - b.y * e2.x);  // b x e2
  const V3<float> V(e1.y * b.z - e1.z * b.y, e1.z * b.x - e1.x * b.z, e1.x * b.y - e1.y * b.x);  // e1 x b
  const float arN = fabsf(rN);
  if (arN > bound.x) {  // every sample has sign(dn) = sign(rN): edge cull on the un-jittered ray
    const float lb = sqrt_approx(dot(b, b));
    const float mU = lb * bound.z, mV = lb * bound.y;
    const unsigned sg = __float_as_uint(rN) & 0x80000000u;
    const float eu = xor_sign(dot(r, U), sg), ev = xor_sign(dot(r, V), sg);
    if ((eu < -mU) | (ev < -mV) | ((eu + ev) - (mU + mV) > arN + bound.x)) return;
  }
  const float q1 = num * num * inv_r2;  // t^2 |d|^2 < r^2  <=>  q1 |d|^2 < dn^2
  const unsigned numb = __float_as_uint(num);
#pragma unroll
  for (int k = 0; k < CH; k++) {
    const V3<float> d = rays.d[k];
    const float dn = (d.x * c0 - d.y * c1) + d.z * c2;  // det A = -dn; t = -num/dn; u = E1/dn; v = E2/dn
    const float E1 = dot(d, U), E2 = dot(d, V);
    const unsigned dnb = __float_as_uint(dn);
    const unsigned sx = ((__float_as_uint(E1) ^ dnb) | (__float_as_uint(E2) ^ dnb)) | ~(numb ^ dnb);
    const bool hit = ((int)sx >= 0) & (fabsf(E1 + E2) <= fabsf(dn)) & (q1 * rays.dd_of(k) < dn * dn) & (((alive >> k) & 1u) != 0u);
    occ |= hit ? (1u << k) : 0u;
  }
}

// Sphere part of in_shadow (kernels.cl:278-307)
template <class T, int CH>
__device__ __forceinline__ void shadow_spheres(V3<T> start, const ShadowRays<T, CH> &rays, T radius_sq, unsigned &occ) {
#pragma unroll
  for (int i = 0; i < RT_SPHERES; i++) {
    if (c_sphere_color[i].w == -1.0f) continue;  // glass casts no shadow (kernels.cl:279)
    const float4 cr = c_sphere_center_r2[i];
    const V3<T> L = start - xyz<T>(cr);
    const T c = dot(L, L) - T(cr.w);
#pragma unroll
    for (int k = 0; k < CH; k++) {
      if ((occ >> k) & 1u) continue;
      const T a = dot(rays.d[k], rays.d[k]);
      const T b = T(2.0f) * dot(rays.d[k], L);
      const T disc = b * b - T(4.0f) * a * c;
      if (disc < T(0.0f)) continue;
      T x0, x1;
      if constexpr (is_strict<T>::value) {
        const T sq = sqrt_(disc);
        const T q = (b > T(0.0f)) ? T(-0.5f) * (b + sq) : T(-0.5f) * (b - sq);
        x0 = div_(q, a);
        x1 = div_(c, q);
      } else {
        const float sq = sqrt_approx(disc);
        const float q = (b > 0.0f) ? -0.5f * (b + sq) : -0.5f * (b - sq);
        x0 = q * rcp_approx(a);
        x1 = c * rcp_approx(q);
      }
      const T x_min = cl_min(x0, x1);
      const T x_max = cl_max(x0, x1);
      const V3<T> min_dir = scale(x_min, rays.d[k]);
      const V3<T> max_dir = scale(x_max, rays.d[k]);
      const T min_dist = dot(min_dir, min_dir);
      const T max_dist = dot(max_dir, max_dir);
      if ((x_min >= T(0.0f) && min_dist < radius_sq) || (x_max >= T(0.0f) && max_dist < radius_sq)) occ |= 1u << k;
    }
  }
}

// ---------------------------------------------------------------------------
// Tracer: brute force over the shared-memory scene.
//   closest(start, dir, hit)           kernels.cl:92-166 / :168-241
//   shadow(start, rays, r, radius_sq)  kernels.cl:243-311 for CH rays sharing an origin -> occlusion mask
// ---------------------------------------------------------------------------
template <class T> struct BruteTracer {
  static constexpr bool kSkipUnlit = false;  // the reference's plain loops: trace every shadow ray
  SceneView sc;
  __device__ __forceinline__ BruteTracer<sfloat> strict() const {  // same scene, reference arithmetic
    BruteTracer<sfloat> t;
    t.sc = sc;
    return t;
  }
  // primary ray of the fast policy: reference arithmetic (see shade_pixel)
  __device__ __forceinline__ void primary_strict(V3<sfloat> cam, V3<sfloat> dir, HitRec<sfloat> &hit) const { strict().closest(cam, dir, hit); }

  __device__ __forceinline__ void closest(V3<T> start, V3<T> dir, HitRec<T> &hit) const {
    ClosestState<T> cs;
    cs.reset();
    const V3<T> nd = -dir;
    for (int i = 0; i < sc.n; i++) closest_tri_test<T, true>(sc.ta[i], sc.tb[i], sc.tc[i], start, nd, i, i, cs);
    if (cs.id >= 0) {
      hit.id = cs.id;
      hit.point = hit_point<T>(sc.ta[cs.id], sc.tb[cs.id], sc.tc[cs.id], cs.u, cs.v);
      hit.normal = xyz<T>(sc.tn[cs.id]);
      hit.color = sc.tcol[cs.id];
    }
    closest_spheres<T>(start, dir, cs.t, hit);
  }

  template <int CH>
  __device__ __forceinline__ unsigned shadow(V3<T> start, const ShadowRays<T, CH> &rays, V3<T> r, T radius_sq) const {
    constexpr unsigned FULL = (CH >= 32) ? 0xffffffffu : ((1u << CH) - 1u);
    unsigned occ = 0u;
    for (int i = 0; i < sc.n_sh; i++) {
      shadow_pair<T, CH>(sc.sa[i], sc.sb[i], sc.sc[i], start, rays, radius_sq, occ);
      if (occ == FULL) return occ;
    }
    shadow_spheres<T, CH>(start, rays, radius_sq, occ);
    return occ;
  }
};

// Compatibility wrapper used by the fast kernel for its bounce rays.
template <class T> __device__ __forceinline__ void closest_hit(const SceneView &sc, V3<T> start, V3<T> dir, HitRec<T> &hit) {
  BruteTracer<T> tr;
  tr.sc = sc;
  tr.closest(start, dir, hit);
}

// ---------------------------------------------------------------------------
// direct_light: kernels.cl:313-340.  The jitter sequence depends on the pixel
// id only (same S jitters for every AA sample and bounce of a pixel).
// ---------------------------------------------------------------------------
template <class T, int CH, class Tracer>
__device__ __forceinline__ V3<T> direct_light(const Tracer &tr, V3<T> point, V3<T> normal, V3<T> light_pos, int S, int global_id) {
  // (uint3)(global_id, global_id*91.0f, global_id*19.0f) then one xorshift (kernels.cl:319)
  uint32_t rx = xorshift32((uint32_t)global_id);
  uint32_t ry = xorshift32(__float2uint_rz(__fmul_rn(__int2float_rn(global_id), 91.0f)));
  uint32_t rz = xorshift32(__float2uint_rz(__fmul_rn(__int2float_rn(global_id), 19.0f)));
  const V3<T> dir = light_pos - point;
  const V3<T> start = point + scale(T(RT_BIAS), dir);
  const T radius_sq = (dir.x * dir.x + dir.y * dir.y) + dir.z * dir.z;
  const T lam = T(RT_LIGHT_COLOR) * cl_max(dot(dir, normal), T(0.0f));
  const T den = T(4.0f) * T(RT_PI_F) * radius_sq;
  if constexpr (Tracer::kSkipUnlit) {
    // a point that faces away from the light: every term mask*lam/den is exactly +0 whatever the shadow rays
    // find (kernels.cl:334-336), so they need not be traced
    if (raw(lam) == 0.0f && raw(den) > 0.0f && raw(den) < 3.0e38f) return V3<T>(T(0.0f), T(0.0f), T(0.0f));
  }
  T total = T(0.0f);
  [[maybe_unused]] int lit = 0;
  for (int s0 = 0; s0 < S; s0 += CH) {
    ShadowRays<T, CH> rays;
#pragma unroll
    for (int k = 0; k < CH; k++) {
      rx = xorshift32(rx);
      ry = xorshift32(ry);
      rz = xorshift32(rz);
      rays.d[k] = dir + V3<T>(crush1<T>(rx, RT_LIGHT_SPREAD), crush1<T>(ry, RT_LIGHT_SPREAD), crush1<T>(rz, RT_LIGHT_SPREAD));
    }
    rays.finish();
    unsigned occ = tr.template shadow<CH>(start, rays, dir, radius_sq);
    if (s0 + CH > S) occ |= ~0u << (S - s0);  // ragged last chunk: ignore the padding samples
    if constexpr (is_strict<T>::value) {
#pragma unroll
      for (int k = 0; k < CH; k++) {
        if (s0 + k < S) {
          const T mask = ((occ >> k) & 1u) ? T(0.0f) : T(1.0f);
          total = total + div_(mask * lam, den);
        }
      }
    } else {
      lit += CH - __popc(occ & ((CH >= 32) ? 0xffffffffu : ((1u << CH) - 1u)));
    }
  }
  if constexpr (is_strict<T>::value) {
    const T r = div_(total, T(__int2float_rn(S)));
    return V3<T>(r, r, r);
  } else {
    const float r = (float)lit * lam * rcp_approx(den * (float)S);
    return V3<T>(r, r, r);
  }
}

// kernels.cl:54-65
template <class T>
__device__ __forceinline__ void reflect_ray(V3<T> dir, V3<T> normal, V3<T> point, V3<T> &o_start, V3<T> &o_dir, float &o_medium) {
  const T dn = dot(dir, normal);
  const V3<T> r = dir - scale(T(2.0f), scale(dn, normal));
  o_start = point + scale(T(RT_BIAS), r);
  o_medium = RT_AIR;
  o_dir = normalize(r);
}

// kernels.cl:67-88 (TIR branch is dead code: sqrt(<0) is NaN, NaN<0 is false; the
// NaN ray then hits nothing and the sample is black)
template <class T>
__device__ __forceinline__ void refract_ray(V3<T> dir, V3<T> normal, V3<T> point, float medium, V3<T> &o_start, V3<T> &o_dir,
                                            float &o_medium) {
  const bool air = (medium == RT_AIR);
  const T n1 = air ? T(RT_AIR) : T(RT_GLASS);
  const T n2 = air ? T(RT_GLASS) : T(RT_AIR);
  T c1 = dot(normal, dir);
  if (c1 < T(0.0f)) normal = scale(T(-1.0f), normal);
  c1 = abs_(c1);
  const T n = div_(n1, n2);
  const T c2 = sqrt_(T(1.0f) - (n * n) * (T(1.0f) - (c1 * c1)));
  const V3<T> r = scale(n, dir) + scale(n * c1 - c2, -normal);
  o_start = point + scale(T(RT_BIAS), r);
  o_medium = raw(n2);
  o_dir = normalize(r);
}

// kernels.cl:37-40
template <class T> __device__ __forceinline__ uint32_t pack_argb(V3<T> c) {
  const uint32_t r = __float2uint_rz(raw(cl_min(cl_max(T(255.0f) * c.x, T(0.f)), T(255.f))));
  const uint32_t g = __float2uint_rz(raw(cl_min(cl_max(T(255.0f) * c.y, T(0.f)), T(255.f))));
  const uint32_t b = __float2uint_rz(raw(cl_min(cl_max(T(255.0f) * c.z, T(0.f)), T(255.f))));
  return (255u << 24) + (r << 16) + (g << 8) + b;
}

__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ sfloat add_rn(sfloat a, sfloat b) { return a + b; }
__device__ __forceinline__ sfloat mul_rn(sfloat a, sfloat b) { return a * b; }

// One ray of a pixel of `draw` (kernels.cl:368-428): sub-pixel sample (dx, dy) of pixel (x, y).  The direct light of a
// diffuse hit and the tail of secondary_light (kernels.cl:342-365) share one call site: loop { miss -> black; diffuse ->
// shade (x0.9 after a bounce) and stop; mirror/glass -> bounce, up to B times }.  False: the ray adds nothing to the pixel.
template <class T, int CH, class Tracer>
__device__ __forceinline__ bool shade_sample(const Tracer &tr, const FrameParams &p, int x, int y, int dx, int dy, int global_id, V3<T> &contrib,
                                             unsigned &n_shadow_calls, unsigned &n_bounce) {
  const T SW = T(__int2float_rn(p.W)), SH = T(__int2float_rn(p.H));
  const int A = p.A;
  const T fA = T(__int2float_rn(A));
  const V3<T> light(T(p.light[0]), T(p.light[1]), T(p.light[2]));
  V3<T> dir;
  HitRec<T> hit;
  hit.id = -1;
  hit.color = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
  if constexpr (is_strict<T>::value) {
    const V3<T> base(T(__int2float_rn(x * A)) - div_(SW * fA, T(2.0f)), T(__int2float_rn(y * A)) - div_(SH * fA, T(2.0f)), T(p.focal));
    const V3<T> r0(T(p.rot[0]), T(p.rot[1]), T(p.rot[2])), r1(T(p.rot[3]), T(p.rot[4]), T(p.rot[5])), r2(T(p.rot[6]), T(p.rot[7]), T(p.rot[8]));
    const V3<T> cam(T(p.cam[0]), T(p.cam[1]), T(p.cam[2]));
    const V3<T> d0 = base + V3<T>(T(__int2float_rn(dx)), T(__int2float_rn(dy)), T(0.0f));
    dir = normalize(V3<T>(dot(r0, d0), dot(r1, d0), dot(r2, d0)));
    tr.closest(cam, dir, hit);
  } else {
    // Fast policy: the PRIMARY ray and its hit are still evaluated with the reference's exact sequence
    // (a small part of the work), so that primary visibility is bit-identical to the reference — the
    // default camera puts box edges exactly on pixel boundaries.
    typedef sfloat S;
    const S SWs(__int2float_rn(p.W)), SHs(__int2float_rn(p.H)), fAs(__int2float_rn(A));
    const V3<S> bs(S(__int2float_rn(x * A)) - div_(SWs * fAs, S(2.0f)), S(__int2float_rn(y * A)) - div_(SHs * fAs, S(2.0f)), S(p.focal));
    const V3<S> d0 = bs + V3<S>(S(__int2float_rn(dx)), S(__int2float_rn(dy)), S(0.0f));
    const V3<S> s0(S(p.rot[0]), S(p.rot[1]), S(p.rot[2])), s1(S(p.rot[3]), S(p.rot[4]), S(p.rot[5])), s2(S(p.rot[6]), S(p.rot[7]), S(p.rot[8]));
    const V3<S> ds = normalize(V3<S>(dot(s0, d0), dot(s1, d0), dot(s2, d0)));
    HitRec<S> hs;
    hs.id = -1;
    hs.color = hit.color;
    hs.point = V3<S>(S(0.0f), S(0.0f), S(0.0f));
    hs.normal = hs.point;
    tr.primary_strict(V3<S>(S(p.cam[0]), S(p.cam[1]), S(p.cam[2])), ds, hs);
    dir = V3<T>(ds.x.v, ds.y.v, ds.z.v);
    hit.id = hs.id;
    hit.point = V3<T>(hs.point.x.v, hs.point.y.v, hs.point.z.v);
    hit.normal = V3<T>(hs.normal.x.v, hs.normal.y.v, hs.normal.z.v);
    hit.color = hs.color;
  }
  float medium = RT_AIR;
  bool bounced = false;
  int bounce = 0;
  while (hit.id != -1) {
    if (hit.color.w > 0.0f) {
      const V3<T> dl = direct_light<T, CH, Tracer>(tr, hit.point, hit.normal, light, p.S, global_id);
      n_shadow_calls++;
      // (single-rounded operations under either policy: the one-ray-per-lane and the looped form of a pixel must agree)
      const V3<T> lightv(add_rn(T(RT_INDIRECT), dl.x), add_rn(T(RT_INDIRECT), dl.y), add_rn(T(RT_INDIRECT), dl.z));
      const V3<T> col = xyz<T>(hit.color);
      // primary: colour*(indirect + direct) (kernels.cl:422); after a bounce: 0.9*light*colour (:355)
      contrib = bounced ? V3<T>(mul_rn(mul_rn(T(0.9f), lightv.x), col.x), mul_rn(mul_rn(T(0.9f), lightv.y), col.y), mul_rn(mul_rn(T(0.9f), lightv.z), col.z))
                        : V3<T>(mul_rn(col.x, lightv.x), mul_rn(col.y, lightv.y), mul_rn(col.z, lightv.z));
      return true;
    }
    if (bounce >= p.B) break;
    bounce++;
    V3<T> start, ndir;
    if (hit.color.w == 0.0f) reflect_ray<T>(dir, hit.normal, hit.point, start, ndir, medium);
    else refract_ray<T>(dir, hit.normal, hit.point, medium, start, ndir, medium);
    dir = ndir;
    hit.id = -1;
    hit.color.w = 1.0f;
    tr.closest(start, dir, hit);
    n_bounce++;
    bounced = true;
  }
  return false;
}

// const int global_id = y*SCREEN_WIDTH + x  in float arithmetic (kernels.cl:380)
template <class T> __device__ __forceinline__ int pixel_global_id(const FrameParams &p, int x, int y) {
  return __float2int_rz(raw(T(__int2float_rn(y)) * T(__int2float_rn(p.W)) + T(__int2float_rn(x))));
}

// One pixel of `draw` (kernels.cl:368-428): its A*A rays one after the other, summed in the reference's order.
template <class T, int CH, class Tracer>
__device__ __forceinline__ uint32_t shade_pixel(const Tracer &tr, const FrameParams &p, int x, int y) {
  const int A = p.A;
  const int global_id = pixel_global_id<T>(p, x, y);
  V3<T> total(T(0.0f), T(0.0f), T(0.0f));
  unsigned n_shadow_calls = 0u, n_bounce = 0u;  // ray statistics (only reported with RT_FLAG_COUNT_RAYS)
#pragma unroll 1
  for (int dy = 0; dy < A; dy++) {
#pragma unroll 1
    for (int dx = 0; dx < A; dx++) {
      V3<T> c;
      if (shade_sample<T, CH, Tracer>(tr, p, x, y, dx, dy, global_id, c, n_shadow_calls, n_bounce))
        total = V3<T>(add_rn(total.x, c.x), add_rn(total.y, c.y), add_rn(total.z, c.z));
    }
  }
  if (p.ray_counters) {
    atomicAdd(p.ray_counters + 0, (unsigned long long)(A * A));
    atomicAdd(p.ray_counters + 1, (unsigned long long)n_shadow_calls * (unsigned long long)p.S);
    atomicAdd(p.ray_counters + 2, (unsigned long long)n_bounce);
  }
  const T fa = T(__int2float_rn(A * A));
  return pack_argb<T>(V3<T>(div_(total.x, fa), div_(total.y, fa), div_(total.z, fa)));
}

// The same pixel by FOUR lanes (lane & 3 = which of the 2x2 rays; requires A == 2): each lane traces one ray, the four
// contributions are added in the reference's order by every lane of the quad, so all four return the pixel.  A long ray
// chain (a mesh behind a BVH: hundreds of dependent node fetches per ray) ends four times sooner, and the launch with it.
// Called by all 32 lanes of a warp; `active` = this lane's pixel lies inside the launch.
template <class T, int CH, class Tracer>
__device__ __forceinline__ uint32_t shade_pixel_quad(const Tracer &tr, const FrameParams &p, int x, int y, bool active) {
  const int lane = threadIdx.x & 31, q = lane & 3;
  V3<T> c(T(0.0f), T(0.0f), T(0.0f));
  unsigned n_shadow_calls = 0u, n_bounce = 0u;
  bool has = false;
  if (active) has = shade_sample<T, CH, Tracer>(tr, p, x, y, q & 1, q >> 1, pixel_global_id<T>(p, x, y), c, n_shadow_calls, n_bounce);
  V3<T> total(T(0.0f), T(0.0f), T(0.0f));
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int src = (lane & ~3) | k;
    const bool hk = __shfl_sync(0xffffffffu, has ? 1 : 0, src) != 0;
    const V3<T> ck(T(__shfl_sync(0xffffffffu, raw(c.x), src)), T(__shfl_sync(0xffffffffu, raw(c.y), src)), T(__shfl_sync(0xffffffffu, raw(c.z), src)));
    if (hk) total = V3<T>(add_rn(total.x, ck.x), add_rn(total.y, ck.y), add_rn(total.z, ck.z));
  }
  if (p.ray_counters && active) {
    atomicAdd(p.ray_counters + 0, 1ull);
    atomicAdd(p.ray_counters + 1, (unsigned long long)n_shadow_calls * (unsigned long long)p.S);
    atomicAdd(p.ray_counters + 2, (unsigned long long)n_bounce);
  }
  const T fa = T(4.0f);
  return pack_argb<T>(V3<T>(div_(total.x, fa), div_(total.y, fa), div_(total.z, fa)));
}

}  // namespace rt
