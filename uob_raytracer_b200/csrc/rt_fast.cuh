// rt_fast.cuh — the fast arithmetic policy of the brute-force path, restructured
// around what is shared between rays.  Same algorithm and same results as the
// generic code in rt_brute.cuh up to floating-point rounding (within the
// north-star tolerance; the strict kernel is the bit-exact anchor):
//
//  * Primary rays all start at the camera, so b = cam - v0, det[b,e1,e2] and the
//    two cross products U = b x e2, V = e1 x b are per-frame constants of each
//    triangle (computed once per block into shared memory).  A closest-hit test
//    is then three dot products with the ray direction:
//        dn = d.N        (det A = -dn,   N = e1 x e2)
//        E1 = d.U, E2 = d.V              (u = E1/dn, v = E2/dn, t = -det[b,e1,e2]/dn)
//    and the inside test u>=0, v>=0, u+v<=1, t>=0 is decided on sign-corrected
//    numerators, dividing only for the few candidates that pass.
//    A batch of RB rays of one pixel is tested per triangle load.
//
//  * Shadow rays of one shading point share their origin.  Per (point, triangle)
//    a conservative plane test first decides whether ANY jittered ray can pass
//    the reference's first stage (0 <= t and |t d|^2 < |r|^2): with
//    num = b.N, rN = r.N, |j.N| <= m := jmax|N| and |r|/|d_s| <= k,
//        stage 1 passes for sample s  =>  |num| < (m - sign(num) rN) k.
//    Pairs failing that bound are skipped for all S samples (about 70 % of them
//    in the Cornell box: walls never lie between a surface point and the light).
//    The bound is inflated (kJitterMax, kSlack) far beyond rounding error, so a
//    skipped pair is one the reference rejects for every sample.
//
//  * The mirror sphere's shadow test is skipped per point when the whole cone of
//    jittered rays misses the sphere (same kind of conservative bound).
#pragma once
#include "rt_brute.cuh"

namespace rt {

// jitter components lie in [-0.025, 0.025] (crush, kernels.cl:49-52): |j| <= 0.025*sqrt(3) = 0.0433013
constexpr float kJitterMax = 0.0445f;  // inflated by 2.7 %
constexpr float kSlack = 1.002f;

struct FastScene {
  SceneView g;                  // generic arrays (bounce rays, hit attributes)
  const float4 *pa, *pb, *pc;   // primary-ray constants: (c0,c1,c2,-det[b,e1,e2]), (U,0), (V,0)
  const float4 *sd;             // shadow casters: (m = kJitterMax*|N|, 0, 0, 0)
};

__device__ __forceinline__ float xor_sign(float v, unsigned signbit) { return __uint_as_float(__float_as_uint(v) ^ signbit); }

// Per-block prologue: per-triangle constants of rays starting at `cam`, evaluated with the
// reference's own (strict) operation sequence, so the confirm step below reproduces the
// reference's hit decisions bit for bit.
//   pa = (c0, c1, c2, det[b,e1,e2])
//   pb = (U0, U1, U2, g)    cofactors of det[-d, b, e2] (kernels.cl:31-35 sign convention)
//   pc = (V0, V1, V2, 3g)   cofactors of det[-d, e1, b]
// g bounds the rounding error of the filter's dot products: 2e-6 * max L1 norm (|d_i| <= 1).
__device__ __forceinline__ void primary_constants(const SceneView &g, float4 *pa, float4 *pb, float4 *pc, V3<float> cam, int i) {
  typedef sfloat S;
  const float4 A = g.ta[i], Bq = g.tb[i], C = g.tc[i];
  const V3<S> b(S(cam.x) - S(A.x), S(cam.y) - S(A.y), S(cam.z) - S(A.z));
  const V3<S> e1 = xyz<S>(Bq), e2 = xyz<S>(C);
  const S detA0 = (b.x * S(A.w) - b.y * S(Bq.w)) + b.z * S(C.w);
  const S U0 = b.y * e2.z - b.z * e2.y, U1 = b.x * e2.z - b.z * e2.x, U2 = b.x * e2.y - b.y * e2.x;
  const S V0 = e1.y * b.z - e1.z * b.y, V1 = e1.x * b.z - e1.z * b.x, V2 = e1.x * b.y - e1.y * b.x;
  const float l1 = fmaxf(fmaxf(fabsf(U0.v) + fabsf(U1.v) + fabsf(U2.v), fabsf(V0.v) + fabsf(V1.v) + fabsf(V2.v)),
                         fabsf(A.w) + fabsf(Bq.w) + fabsf(C.w));
  const float tol = 2e-6f * l1;
  pa[i] = make_float4(A.w, Bq.w, C.w, detA0.v);
  pb[i] = make_float4(U0.v, U1.v, U2.v, tol);
  pc[i] = make_float4(V0.v, V1.v, V2.v, 3.0f * tol);
}

// Closest triangle for RB rays from the camera (kernels.cl:100-129).  The fast arithmetic only
// FILTERS: a (ray, triangle) pair that passes the inside test with the tolerance g is confirmed
// with the reference's exact operation sequence (strict IEEE), so hit / miss decisions, the
// winning triangle and t, u, v are bit-identical to the reference — the default camera puts box
// edges exactly on pixel boundaries, where any other rounding flips pixels.
// best = -1 on a miss; bt = t, bu/bv = barycentrics (all strict values).
template <int RB>
__device__ __forceinline__ void primary_triangles(const FastScene &sc, const V3<float> (&d)[RB], int (&best)[RB], float (&bt)[RB],
                                                  float (&bu)[RB], float (&bv)[RB]) {
#pragma unroll
  for (int k = 0; k < RB; k++) {
    best[k] = -1;
    bt[k] = 3.402823466e+38f;
    bu[k] = 0.0f;
    bv[k] = 0.0f;
  }
  for (int i = 0; i < sc.g.n; i++) {
    const float4 PA = sc.pa[i], PB = sc.pb[i], PC = sc.pc[i];
#pragma unroll
    for (int k = 0; k < RB; k++) {
      // dn = -det A,  E1 = -det[-d,b,e2],  E2 = -det[-d,e1,b]
      const float dn = (d[k].x * PA.x - d[k].y * PA.y) + d[k].z * PA.z;
      const float E1 = (d[k].x * PB.x - d[k].y * PB.y) + d[k].z * PB.z;
      const float E2 = (d[k].x * PC.x - d[k].y * PC.y) + d[k].z * PC.z;
      const unsigned sb = __float_as_uint(dn) & 0x80000000u;
      const float e1 = xor_sign(E1, sb), e2 = xor_sign(E2, sb), t0 = xor_sign(PA.w, sb), adn = fabsf(dn);
      // u >= 0, v >= 0, u + v <= 1 (each widened by the tolerance) and t = det[b,e1,e2]/det A >= 0
      if ((e1 >= -PB.w) & (e2 >= -PB.w) & ((e1 + e2) - adn <= PC.w) & (t0 <= 0.0f)) {
        typedef sfloat S;
        const V3<S> nd(S(-d[k].x), S(-d[k].y), S(-d[k].z));
        const S detA = (nd.x * S(PA.x) - nd.y * S(PA.y)) + nd.z * S(PA.z);
        const S inv = rcp_(detA);
        const S t = S(PA.w) * inv;
        const S u = ((nd.x * S(PB.x) - nd.y * S(PB.y)) + nd.z * S(PB.z)) * inv;
        const S v = ((nd.x * S(PC.x) - nd.y * S(PC.y)) + nd.z * S(PC.z)) * inv;
        if (t < S(bt[k]) && u >= S(0.0f) && v >= S(0.0f) && (u + v) <= S(1.0f) && t >= S(0.0f)) {
          bt[k] = t.v;
          best[k] = i;
          bu[k] = u.v;
          bv[k] = v.v;
        }
      }
    }
  }
}

// Shading-point-shared state of the S shadow rays of a pixel.
template <int CH> struct Jitters {
  float x[CH], y[CH], z[CH];
};

// kernels.cl:319,331: the S jitters of a pixel (they depend on the pixel id only)
template <int CH>
__device__ __forceinline__ void make_jitters(int global_id, uint32_t &rx, uint32_t &ry, uint32_t &rz, Jitters<CH> &j) {
#pragma unroll
  for (int k = 0; k < CH; k++) {
    rx = xorshift32(rx);
    ry = xorshift32(ry);
    rz = xorshift32(rz);
    j.x[k] = crush1<float>(rx, RT_LIGHT_SPREAD);
    j.y[k] = crush1<float>(ry, RT_LIGHT_SPREAD);
    j.z[k] = crush1<float>(rz, RT_LIGHT_SPREAD);
  }
}

__device__ __forceinline__ void seed_rng(int global_id, uint32_t &rx, uint32_t &ry, uint32_t &rz) {
  rx = xorshift32((uint32_t)global_id);
  ry = xorshift32(__float2uint_rz(__fmul_rn(__int2float_rn(global_id), 91.0f)));
  rz = xorshift32(__float2uint_rz(__fmul_rn(__int2float_rn(global_id), 19.0f)));
}

// Number of UNOCCLUDED samples among the CH shadow rays start + t (r + j_k)  (in_shadow, kernels.cl:243-311).
// The ray directions d_k = r + j_k are never materialised: every dot product d_k.X is evaluated as
// r.X + j_k.X with r.X computed once per (point, triangle).
template <int CH>
__device__ __forceinline__ int shadow_lit_count(const FastScene &sc, V3<float> start, V3<float> r, float radius_sq,
                                                const Jitters<CH> &j, unsigned valid_mask) {
  constexpr unsigned FULL = (CH >= 32) ? 0xffffffffu : ((1u << CH) - 1u);
  unsigned occ = ~valid_mask & FULL;  // padding samples of a ragged chunk count as occluded (ignored by the caller)
  float dd[CH];                       // |d_k|^2
#pragma unroll
  for (int k = 0; k < CH; k++) {
    const float dx = r.x + j.x[k], dy = r.y + j.y[k], dz = r.z + j.z[k];
    dd[k] = (dx * dx + dy * dy) + dz * dz;
  }
  // |r| / |d_s| <= R / (R - jmax); no bound (k huge) when the light is closer than 2 jmax
  const float R = sqrt_approx(radius_sq);
  const float kk = (R > 2.0f * kJitterMax) ? kSlack * R * rcp_approx(R - kJitterMax) : 1e30f;

  const int n_sh = sc.g.n_sh;
  for (int i = 0; i < n_sh; i++) {
    const float4 A = sc.g.sa[i], Bq = sc.g.sb[i], C = sc.g.sc[i];
    const V3<float> b(start.x - A.x, start.y - A.y, start.z - A.z);
    const float c0 = A.w, c1 = Bq.w, c2 = C.w;
    const float num = (b.x * c0 - b.y * c1) + b.z * c2;  // det[b,e1,e2] = b.N
    const float rN = (r.x * c0 - r.y * c1) + r.z * c2;   // r.N
    const float m = sc.sd[i].x;
    const float w = xor_sign(rN, __float_as_uint(num) & 0x80000000u);
    if (fabsf(num) >= (m - w) * kk) continue;  // no sample can pass stage 1 (see header)
    const V3<float> e1(Bq.x, Bq.y, Bq.z), e2(C.x, C.y, C.z);
    const V3<float> U(b.y * e2.z - b.z * e2.y, b.z * e2.x - b.x * e2.z, b.x * e2.y - b.y * e2.x);  // b x e2
    const V3<float> V(e1.y * b.z - e1.z * b.y, e1.z * b.x - e1.x * b.z, e1.x * b.y - e1.y * b.x);  // e1 x b
    const float rU = dot(r, U), rV = dot(r, V);
    const float num2 = num * num;
#pragma unroll
    for (int k = 0; k < CH; k++) {
      // det A = -dn;  t = -num/dn;  u = E1/dn;  v = E2/dn
      const float dn = fmaf(j.x[k], c0, fmaf(-j.y[k], c1, fmaf(j.z[k], c2, rN)));
      const float E1 = fmaf(j.x[k], U.x, fmaf(j.y[k], U.y, fmaf(j.z[k], U.z, rU)));
      const float E2 = fmaf(j.x[k], V.x, fmaf(j.y[k], V.y, fmaf(j.z[k], V.z, rV)));
      const unsigned sb = __float_as_uint(dn) & 0x80000000u;
      const float e1s = xor_sign(E1, sb), e2s = xor_sign(E2, sb), ns = xor_sign(num, sb);
      const bool hit = (ns <= 0.0f) & (num2 * dd[k] < radius_sq * (dn * dn)) & (e1s >= 0.0f) & (e2s >= 0.0f) & ((e1s + e2s) <= fabsf(dn));
      occ |= hit ? (1u << k) : 0u;
    }
    if (occ == FULL) return 0;
  }
#pragma unroll
  for (int i = 0; i < RT_SPHERES; i++) {
    if (c_sphere_color[i].w == -1.0f) continue;  // glass casts no shadow (kernels.cl:279)
    const float4 cr = c_sphere_center_r2[i];
    const V3<float> L(start.x - cr.x, start.y - cr.y, start.z - cr.z);
    const float LL = dot(L, L);
    const float c = LL - cr.w;
    // Cone cull: the distance from the centre to the line (start, d_s) is at least
    // perp_c - |L| * |d_s/|d_s| - r/|r||  >=  perp_c - |L| jmax/(R - jmax); no real root if that exceeds the radius.
    const float Lr = dot(L, r);
    const float perp2 = LL - Lr * Lr * rcp_approx(radius_sq);
    const float lim = kSlack * sqrt_approx(cr.w) + sqrt_approx(LL) * (kk * kJitterMax * rcp_approx(R));
    if (perp2 > lim * lim) continue;
#pragma unroll
    for (int k = 0; k < CH; k++) {
      if ((occ >> k) & 1u) continue;
      const float a = dd[k];
      const float b = 2.0f * fmaf(j.x[k], L.x, fmaf(j.y[k], L.y, fmaf(j.z[k], L.z, Lr)));
      const float disc = b * b - 4.0f * a * c;
      if (disc < 0.0f) continue;
      const float sq = sqrt_approx(disc);
      const float q = (b > 0.0f) ? -0.5f * (b + sq) : -0.5f * (b - sq);
      const float x0 = q * rcp_approx(a);
      const float x1 = c * rcp_approx(q);
      const float x_min = fminf(x0, x1), x_max = fmaxf(x0, x1);
      // |x d|^2 < r^2
      if ((x_min >= 0.0f && x_min * x_min * a < radius_sq) || (x_max >= 0.0f && x_max * x_max * a < radius_sq)) occ |= 1u << k;
    }
  }
  return CH - __popc(occ);
}

// direct_light (kernels.cl:313-340) for one shading point; jit = the pixel's jitters when S == CH
// (precomputed once per pixel), otherwise they are regenerated chunk by chunk.
template <int CH>
__device__ __forceinline__ float direct_light_fast(const FastScene &sc, V3<float> point, V3<float> normal, V3<float> light_pos, int S,
                                                   int global_id, const Jitters<CH> &jit) {
  const V3<float> r = light_pos - point;
  const V3<float> start = point + scale(RT_BIAS, r);
  const float radius_sq = (r.x * r.x + r.y * r.y) + r.z * r.z;
  const float lam = RT_LIGHT_COLOR * fmaxf(dot(r, normal), 0.0f);
  int lit = 0;
  if (S == CH) {
    lit = shadow_lit_count<CH>(sc, start, r, radius_sq, jit, 0xffffffffu);
  } else {
    uint32_t rx, ry, rz;
    seed_rng(global_id, rx, ry, rz);
    for (int s0 = 0; s0 < S; s0 += CH) {
      Jitters<CH> jj;
      make_jitters<CH>(global_id, rx, ry, rz, jj);
      const unsigned valid = (s0 + CH > S) ? ((1u << (S - s0)) - 1u) : 0xffffffffu;
      const int pad = (s0 + CH > S) ? (s0 + CH - S) : 0;
      lit += shadow_lit_count<CH>(sc, start, r, radius_sq, jj, valid);
      (void)pad;
    }
  }
  return (float)lit * lam * rcp_approx(4.0f * RT_PI_F * radius_sq * (float)S);
}

// secondary_light (kernels.cl:342-365): bounce rays use the generic closest-hit search.
template <int CH>
__device__ __forceinline__ V3<float> secondary_light_fast(const FastScene &sc, V3<float> dir, HitRec<float> hit, V3<float> light_pos, int S,
                                                          int B, int global_id, const Jitters<CH> &jit) {
  float medium = RT_AIR;
  for (int b = 0; b < B && hit.color.w <= 0.0f; b++) {
    V3<float> start, ndir;
    if (hit.color.w == 0.0f) reflect_ray<float>(dir, hit.normal, hit.point, start, ndir, medium);
    else refract_ray<float>(dir, hit.normal, hit.point, medium, start, ndir, medium);
    dir = ndir;
    hit.id = -1;
    hit.color.w = 1.0f;
    closest_hit<float>(sc.g, start, dir, hit);
    if (hit.id != -1 && hit.color.w > 0.0f) {
      const float l = 0.9f * (RT_INDIRECT + direct_light_fast<CH>(sc, hit.point, hit.normal, light_pos, S, global_id, jit));
      return V3<float>(l * hit.color.x, l * hit.color.y, l * hit.color.z);
    }
  }
  return V3<float>(0.0f, 0.0f, 0.0f);
}

}  // namespace rt
