// rt_fast.cuh — the fast arithmetic policy of the brute-force path, restructured
// around what is shared between rays.  Same algorithm and same results as the
// generic code in rt_brute.cuh up to floating-point rounding of the SHADING
// (primary visibility is bit-identical, see below).  The culls are shared by
// RT_FLAG_STRICT_IEEE (shadow_occ_strict / direct_light_strict at the end of this
// file): there they only choose which tests are evaluated, the evaluation itself is
// the reference's operation sequence and the frame stays bit-identical.
//
//  * Primary rays all start at the camera, so b = cam - v0, det[b,e1,e2] and the
//    cofactors of det[-d,b,e2], det[-d,e1,b] are per-frame constants of each
//    triangle (computed once per block into shared memory, with the reference's
//    own operation sequence).  A closest-hit test is three dot products with the
//    ray direction:  dn = -det A,  E1 = -det[-d,b,e2],  E2 = -det[-d,e1,b]
//    (u = E1/dn, v = E2/dn, t = -det[b,e1,e2]/dn).  The fast arithmetic only
//    FILTERS (inside test on sign-corrected numerators, widened by a rounding
//    tolerance); every survivor is confirmed with the reference's exact strict-IEEE
//    sequence, so hit/miss, the winning triangle and (t,u,v) are bit-identical to
//    the reference.  This matters: the default camera puts box edges exactly on
//    pixel boundaries.
//
//  * Block-level triangle binning: the three edge functions and dn are affine in
//    the pixel coordinates, so a triangle whose edge function is negative at the
//    four corner rays of the block's pixel tile fails the inside test for every
//    ray of the block.  Such triangles are dropped once per block (ballot
//    compaction into an index list that keeps the upload order, which the
//    lowest-index-wins tie rule of kernels.cl:120 depends on).
//
//  * Shadow rays of one shading point share their origin.  Two conservative
//    culls per (point, triangle) decide whether ANY of the S jittered rays can hit:
//      plane:  with num = b.N, rN = r.N, |j.N| <= m := jmax|N|, |r|/|d_s| <= k,
//              stage 1 (0 <= t, |t d|^2 < |r|^2) passes for some s
//              =>  |num| < (m - sign(num) rN) k;
//      edges:  with U = b x e2, V = e1 x b:  |j.U| <= jmax|b||e2|, |j.V| <= jmax|b||e1|;
//              if the un-jittered ray misses an edge by more than that bound, all do.
//    The bounds are inflated (kJitterMax, kSlack) far beyond rounding error, so a
//    skipped pair is one the reference rejects for every sample.  Survivors run
//    the division-free per-sample test.
//
//  * The mirror sphere's shadow test is skipped per point when the whole cone of
//    jittered rays misses the sphere.
//
//  * Warp-level caster cull (box_may_be_shadowed_by): one caster list per warp from the bounding
//    box of the warp's diffuse primary hits — plane bound over the box, and a beam cull against
//    the cone frustum that contains every shadow ray from the box to the jittered light.
#pragma once
#include <type_traits>

#include "rt_brute.cuh"

namespace rt {

// kJitterMax / kSlack (the inflated bound on |j| and the slack factor) live in rt_brute.cuh.

// Shared-memory view of the fast kernel.
//   prim[3i+0] = (c0, c1, c2, det[b,e1,e2])     b = cam - v0
//   prim[3i+1] = (U0, U1, U2, g)                cofactors of det[-d,b,e2] (kernels.cl:31-35 convention)
//   prim[3i+2] = (V0, V1, V2, 3g)               cofactors of det[-d,e1,b]
//   aff[3i+0]  = (gA.x, gA.y, f gA.z, det[b,e1,e2])   dn(vx, vy) = vx gA.x + vy gA.y + f gA.z   (see primary_affine)
//   aff[3i+1]  = (gU.x, gU.y, f gU.z, tau)            E1(vx, vy)
//   aff[3i+2]  = (gV.x, gV.y, f gV.z, 0)              E2(vx, vy)
//   shad[4k+0] = (v0.xyz, jmax|N|)   shad[4k+1] = (c0, c1, c2, 0)
//   shad[4k+2] = (e1.xyz, jmax|e1|)  shad[4k+3] = (e2.xyz, jmax|e2|)     k-th shadow caster
struct FastScene {
  SceneView g;          // generic SoA arrays (bounce rays, hit attributes)
  const float4 *prim;   // 3 per triangle
  const float4 *aff;    // 3 per triangle: the same three determinants as affine functions of the sub-pixel coordinates
  const float4 *shad;   // 4 per shadow caster
  const int *plist;     // triangles surviving the block's binning, ascending
  int n_prim;           // entries in plist
  const int *clist;     // shadow casters to test for the current shading point (warp list or all)
  int n_clist;
};

constexpr int kWarpListMax = 64;  // per-warp caster lists are kept for scenes with at most this many casters

// Monotonic float -> unsigned map, so that warp-wide min/max can use the integer REDUX unit.
__device__ __forceinline__ unsigned ord_of_float(float f) {
  const unsigned u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float float_of_ord(unsigned o) {
  return __uint_as_float(o ^ ((o >> 31) ? 0x80000000u : 0xffffffffu));
}

// Warp-level version of the plane cull of shadow_lit_count: can caster `c` be hit by a shadow ray
// from ANY point of the box [lo,hi] (the bounding box of the warp's 8x4 shading points)?
// num(P) = (P + bias (L-P) - v0).N and rN(P) = (L-P).N are affine in P, so over the box they stay within
// +-hN of their centre values, hN = |N|.half-extent; |r|/|d_s| <= k(R) is largest at the box point
// nearest to the light.  False only if the per-point cull would reject the pair at every point.
__device__ __forceinline__ bool box_may_be_shadowed_by(const float4 *shad, const float4 *sbound, int c, V3<float> lo, V3<float> hi,
                                                       V3<float> light) {
  const float4 Q0 = shad[4 * c], Q1 = shad[4 * c + 1];
  const V3<float> ctr((lo.x + hi.x) * 0.5f, (lo.y + hi.y) * 0.5f, (lo.z + hi.z) * 0.5f);
  const V3<float> half((hi.x - lo.x) * 0.5f, (hi.y - lo.y) * 0.5f, (hi.z - lo.z) * 0.5f);
  const V3<float> r(light.x - ctr.x, light.y - ctr.y, light.z - ctr.z);
  const V3<float> b(ctr.x + RT_BIAS * r.x - Q0.x, ctr.y + RT_BIAS * r.y - Q0.y, ctr.z + RT_BIAS * r.z - Q0.z);
  const float num = (b.x * Q1.x - b.y * Q1.y) + b.z * Q1.z;
  const float rN = (r.x * Q1.x - r.y * Q1.y) + r.z * Q1.z;
  const float hN = 1.0001f * (fabsf(Q1.x) * half.x + fabsf(Q1.y) * half.y + fabsf(Q1.z) * half.z) + 1e-7f * (fabsf(num) + fabsf(rN));
  // nearest distance from the light to the box
  const float dx = fmaxf(fmaxf(lo.x - light.x, light.x - hi.x), 0.0f), dy = fmaxf(fmaxf(lo.y - light.y, light.y - hi.y), 0.0f),
              dz = fmaxf(fmaxf(lo.z - light.z, light.z - hi.z), 0.0f);
  const float rmin = sqrtf(dx * dx + dy * dy + dz * dz) * 0.9999f;
  if (!(rmin > 2.0f * kJitterMax)) return true;
  {
    // Beam cull: every shadow ray from a point of the box runs inside the convex hull of ball(box centre,
    // box half-diagonal) and ball(light, 2 jmax) — a cone frustum around the axis centre -> light.  A caster
    // whose bounding sphere lies outside that hull cannot be hit from anywhere in the box.
    const float4 bs = sbound[c];
    const float rb = sqrtf(dot(half, half)) * 1.0001f + 1e-6f, rl = 2.0f * kJitterMax;
    const float aa = dot(r, r);
    const V3<float> xc(bs.x - ctr.x, bs.y - ctr.y, bs.z - ctr.z);
    const float s = fminf(fmaxf(dot(xc, r) / aa, 0.0f), 1.0f);
    const V3<float> off(xc.x - s * r.x, xc.y - s * r.y, xc.z - s * r.z);
    const float dist = sqrtf(dot(off, off));
    const float sin_a = fabsf(rb - rl) * rsqrtf(aa);
    if (sin_a < 0.999f) {
      const float cos_a = sqrtf(1.0f - sin_a * sin_a);
      if ((dist - (rb + (rl - rb) * s)) * cos_a > bs.w + 1e-5f) return false;
    }
  }
  const float kk = kSlack * rmin / (rmin - kJitterMax);
  if (num - hN > 0.0f) {        // every point is on the + side of the plane
    const float rhs = Q0.w - (rN - hN);
    return !(rhs <= 0.0f || (num - hN) >= rhs * kk);
  }
  if (num + hN < 0.0f) {        // every point is on the - side
    const float rhs = Q0.w + (rN + hN);
    return !(rhs <= 0.0f || -(num + hN) >= rhs * kk);
  }
  return true;
}

// Per-triangle constants of rays starting at `cam`, with the reference's (strict) operation
// sequence, so the confirm step reproduces the reference's decisions bit for bit.
// g bounds the rounding error of the filter's dot products: 2e-6 * max L1 norm (|d_i| <= 1).
__device__ __forceinline__ void primary_constants(float4 A, float4 Bq, float4 C, float4 *prim, V3<float> cam, int i) {
  typedef sfloat S;
  const V3<S> b(S(cam.x) - S(A.x), S(cam.y) - S(A.y), S(cam.z) - S(A.z));
  const V3<S> e1 = xyz<S>(Bq), e2 = xyz<S>(C);
  const S detA0 = (b.x * S(A.w) - b.y * S(Bq.w)) + b.z * S(C.w);
  const S U0 = b.y * e2.z - b.z * e2.y, U1 = b.x * e2.z - b.z * e2.x, U2 = b.x * e2.y - b.y * e2.x;
  const S V0 = e1.y * b.z - e1.z * b.y, V1 = e1.x * b.z - e1.z * b.x, V2 = e1.x * b.y - e1.y * b.x;
  const float l1 = fmaxf(fmaxf(fabsf(U0.v) + fabsf(U1.v) + fabsf(U2.v), fabsf(V0.v) + fabsf(V1.v) + fabsf(V2.v)),
                         fabsf(A.w) + fabsf(Bq.w) + fabsf(C.w));
  const float tol = 2e-6f * l1;
  prim[3 * i + 0] = make_float4(A.w, Bq.w, C.w, detA0.v);
  prim[3 * i + 1] = make_float4(U0.v, U1.v, U2.v, tol);
  prim[3 * i + 2] = make_float4(V0.v, V1.v, V2.v, 3.0f * tol);
}
__device__ __forceinline__ void primary_constants(const SceneView &g, float4 *prim, V3<float> cam, int i) {
  primary_constants(g.ta[i], g.tb[i], g.tc[i], prim, cam, i);
}

// Can triangle i be hit by any ray of a tile whose four corner rays (un-normalised) are dc[0..3]?
// dmax >= |dc| of every corner.  Conservative: false only if every ray of the tile fails the
// filter of primary_triangles.
__device__ __forceinline__ bool tile_may_hit(const float4 *prim, int i, const V3<float> (&dc)[4], float dmax) {
  const float4 PA = prim[3 * i], PB = prim[3 * i + 1], PC = prim[3 * i + 2];
  // a triangle without area (cofactors exactly zero): det A = +-0 and det[b,e1,e2] = 0 for every ray, so t = 0/0 is NaN
  // and every comparison of kernels.cl:120 fails — it is never hit
  if (PA.x == 0.0f && PA.y == 0.0f && PA.z == 0.0f) return false;
  float dn[4], e1[4], e2[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    dn[c] = (dc[c].x * PA.x - dc[c].y * PA.y) + dc[c].z * PA.z;
    e1[c] = (dc[c].x * PB.x - dc[c].y * PB.y) + dc[c].z * PB.z;
    e2[c] = (dc[c].x * PC.x - dc[c].y * PC.y) + dc[c].z * PC.z;
  }
  const bool pos = (dn[0] > 0.0f) & (dn[1] > 0.0f) & (dn[2] > 0.0f) & (dn[3] > 0.0f);
  const bool neg = (dn[0] < 0.0f) & (dn[1] < 0.0f) & (dn[2] < 0.0f) & (dn[3] < 0.0f);
  if (!(pos | neg)) return true;  // the plane's horizon crosses the tile: no conclusion
  const unsigned sb = neg ? 0x80000000u : 0u;
  if (xor_sign(PA.w, sb) > 0.0f) return false;  // t = det[b,e1,e2]/det A < 0 for the whole tile
  const float tol = PB.w * dmax;
  bool out1 = true, out2 = true, out3 = true;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const float a = xor_sign(e1[c], sb), b = xor_sign(e2[c], sb);
    out1 &= a < -tol;
    out2 &= b < -tol;
    out3 &= (a + b) - fabsf(dn[c]) > 3.0f * tol;
  }
  return !(out1 | out2 | out3);
}

// The primary ray through sub-pixel (vx, vy) has the un-normalised direction d0 = R (vx, vy, f) (kernels.cl:384-400), and
// the three determinants of a triangle's test are linear in the direction: dn = d0.a with a = (c0, -c1, c2), hence
// dn = (vx, vy, f).(R^T a) — an AFFINE function of (vx, vy) whose three coefficients are per-frame constants of the
// triangle; likewise E1 and E2.  Scaling the direction scales all three alike, so every sign and every ratio (u = E1/dn,
// v = E2/dn) is that of the reference's normalised ray.  tau bounds the rounding of these evaluations AND of the
// reference's own strict evaluation: 4e-6 x (largest L1 norm of the cofactor vectors) x (longest d0 of the tile).
__device__ __forceinline__ void primary_affine(const float4 *prim, float4 *aff, int i, const float *rot, float focal, float dmax) {
  const float4 PA = prim[3 * i], PB = prim[3 * i + 1], PC = prim[3 * i + 2];
  auto g = [&](float x, float y, float z, float w) {  // R^T (x, -y, z), third component times f
    y = -y;
    return make_float4(rot[0] * x + rot[3] * y + rot[6] * z, rot[1] * x + rot[4] * y + rot[7] * z,
                       (rot[2] * x + rot[5] * y + rot[8] * z) * focal, w);
  };
  aff[3 * i + 0] = g(PA.x, PA.y, PA.z, PA.w);
  aff[3 * i + 1] = g(PB.x, PB.y, PB.z, PB.w * 2.0f * dmax);
  aff[3 * i + 2] = g(PC.x, PC.y, PC.z, 0.0f);
}

// Lazy-exact primary visibility (fast policy).  Classifies every binned triangle for the ray through (vx, vy) as surely
// missed, surely hit, or too close to call: a margin of tau on every inside test, a relative 1e-5 between the distances
// of two hits.  True: the decisions are beyond doubt — the reference's strict arithmetic takes the same ones — and
// best / bu / bv hold the hit (-1 = miss; u, v from approximate division, which only moves the hit point by rounding).
// False: the caller repeats the ray with the reference's exact sequence (primary_triangles).  The default camera puts
// box edges exactly on pixel boundaries: those rays are the ones that come back false.
__device__ __forceinline__ bool primary_fast(const FastScene &sc, float vx, float vy, int &best, float &bu, float &bv) {
  best = -1;
  bu = 0.0f;
  bv = 0.0f;
  float bt = 0.0f;
  bool sure = true;
  for (int l = 0; l < sc.n_prim; l++) {
    const int i = sc.plist[l];
    const float4 *q = sc.aff + 3 * i;
    const float4 QA = q[0], QB = q[1], QC = q[2];
    const float dn = fmaf(vx, QA.x, fmaf(vy, QA.y, QA.z));
    const float E1 = fmaf(vx, QB.x, fmaf(vy, QB.y, QB.z));
    const float E2 = fmaf(vx, QC.x, fmaf(vy, QC.y, QC.z));
    const unsigned sb = __float_as_uint(dn) & 0x80000000u;
    const float e1 = xor_sign(E1, sb), e2 = xor_sign(E2, sb), ts = xor_sign(QA.w, sb), adn = fabsf(dn), tau = QB.w;
    const float slack = adn - (e1 + e2), tau3 = 3.0f * tau;  // (1 - u - v) |dn|
    const bool dn_ok = adn > tau;                             // the sign of dn is beyond doubt
    const bool out = (e1 < -tau) | (e2 < -tau) | (slack < -tau3) | (ts > 0.0f);  // u < 0, v < 0, u + v > 1 or t < 0
    if (dn_ok & out) continue;
    const bool in = (e1 > tau) & (e2 > tau) & (slack > tau3) & (ts < 0.0f);
    if (!(dn_ok & in)) {
      sure = false;
      continue;
    }
    const float inv = rcp_approx(adn);
    const float t = -ts * inv;  // in units of |d0|: the same scale for every triangle of this ray
    if (best >= 0 && fabsf(t - bt) <= 1e-5f * (t + bt)) sure = false;  // two hits at (nearly) the same distance: the tie rule decides
    if (best < 0 || t < bt) {
      bt = t;
      best = i;
      bu = e1 * inv;
      bv = e2 * inv;
    }
  }
  return sure;
}

// (u, v) of the ray through (vx, vy) on triangle i, by the affine forms and approximate division — the values primary_fast
// returns for a hit, for a winner that the exact path had to decide.
__device__ __forceinline__ void primary_uv(const FastScene &sc, int i, float vx, float vy, float &bu, float &bv) {
  const float4 *q = sc.aff + 3 * i;
  const float4 QA = q[0], QB = q[1], QC = q[2];
  const float dn = fmaf(vx, QA.x, fmaf(vy, QA.y, QA.z));
  const float E1 = fmaf(vx, QB.x, fmaf(vy, QB.y, QB.z));
  const float E2 = fmaf(vx, QC.x, fmaf(vy, QC.y, QC.z));
  const unsigned sb = __float_as_uint(dn) & 0x80000000u;
  const float inv = rcp_approx(fabsf(dn));
  bu = xor_sign(E1, sb) * inv;
  bv = xor_sign(E2, sb) * inv;
}

// Does the ray cam + x d0 surely miss both spheres?  (disc/4 = (d0.L)^2 - (d0.d0)(L.L - r^2) < 0 with a relative margin;
// kernels.cl:132-163.)  Anything else — a hit, a graze, the camera inside a sphere — goes to the exact sequence.
__device__ __forceinline__ bool spheres_surely_missed(V3<float> d0, V3<float> cam) {
  bool miss = true;
  const float a = dot(d0, d0);
#pragma unroll
  for (int i = 0; i < RT_SPHERES; i++) {
    const float4 cr = c_sphere_center_r2[i];
    const V3<float> L(cam.x - cr.x, cam.y - cr.y, cam.z - cr.z);
    const float hb = dot(d0, L), h2 = hb * hb, ac = a * (dot(L, L) - cr.w);
    miss &= (h2 - ac) < -1e-5f * (h2 + fabsf(ac));
  }
  return miss;
}

// Closest triangle for one ray from the camera (kernels.cl:100-129); see the header.
// best = -1 on a miss; bt = t, bu/bv = barycentrics (all strict values).
__device__ __forceinline__ void primary_triangles(const FastScene &sc, V3<float> d, int &best, float &bt, float &bu, float &bv) {
  best = -1;
  bt = 3.402823466e+38f;
  bu = 0.0f;
  bv = 0.0f;
  for (int l = 0; l < sc.n_prim; l++) {
    const int i = sc.plist[l];
    const float4 *q = sc.prim + 3 * i;
    const float4 PA = q[0], PB = q[1], PC = q[2];
    const float dn = (d.x * PA.x - d.y * PA.y) + d.z * PA.z;
    const float E1 = (d.x * PB.x - d.y * PB.y) + d.z * PB.z;
    const float E2 = (d.x * PC.x - d.y * PC.y) + d.z * PC.z;
    const unsigned sb = __float_as_uint(dn) & 0x80000000u;
    const float e1 = xor_sign(E1, sb), e2 = xor_sign(E2, sb), t0 = xor_sign(PA.w, sb), adn = fabsf(dn);
    // u >= 0, v >= 0, u + v <= 1 (each widened by the tolerance) and t = det[b,e1,e2]/det A >= 0
    if ((e1 >= -PB.w) & (e2 >= -PB.w) & ((e1 + e2) - adn <= PC.w) & (t0 <= 0.0f)) {
      typedef sfloat S;
      const V3<S> nd(S(-d.x), S(-d.y), S(-d.z));
      const S detA = (nd.x * S(PA.x) - nd.y * S(PA.y)) + nd.z * S(PA.z);
      const S inv = rcp_(detA);
      const S t = S(PA.w) * inv;
      const S u = ((nd.x * S(PB.x) - nd.y * S(PB.y)) + nd.z * S(PB.z)) * inv;
      const S v = ((nd.x * S(PC.x) - nd.y * S(PC.y)) + nd.z * S(PC.z)) * inv;
      if (t < S(bt) && u >= S(0.0f) && v >= S(0.0f) && (u + v) <= S(1.0f) && t >= S(0.0f)) {
        bt = t.v;
        best = i;
        bu = u.v;
        bv = v.v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Closest hit of a bounce ray (kernels.cl:168-241), in two steps: the triangle search fills a ClosestState, then
// finish_closest turns the winner into a hit record and continues with the two spheres.  The search comes in a
// lane-parallel form (every lane scans the triangles for its own ray) and a warp-cooperative form (the lanes of a warp
// share out the triangles of ONE ray at a time) — same per-triangle arithmetic, same winner, hence the same bits.
// ---------------------------------------------------------------------------------------------

// Sphere part of a bounce ray's closest hit and the mirror / glass bounce itself, fast policy: evaluated in the strict
// arithmetic (every operation separately rounded).  They are a few dozen operations on one ray in a hundred, and they are
// inlined into several search loops — with plain float operators each copy would be free to contract differently, and the
// frame would depend on which copy a ray went through (lane mapping, partition over GPUs).
__device__ __forceinline__ V3<sfloat> to_strict(V3<float> v) { return V3<sfloat>(sfloat(v.x), sfloat(v.y), sfloat(v.z)); }
__device__ __forceinline__ V3<float> to_fast(V3<sfloat> v) { return V3<float>(v.x.v, v.y.v, v.z.v); }
__device__ __forceinline__ void closest_spheres_fixed(V3<float> start, V3<float> dir, float current_t, HitRec<float> &hit) {
  HitRec<sfloat> hs;
  hs.id = -1;
  hs.color = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
  hs.point = V3<sfloat>(sfloat(0.0f), sfloat(0.0f), sfloat(0.0f));
  hs.normal = hs.point;
  closest_spheres<sfloat>(to_strict(start), to_strict(dir), sfloat(current_t), hs);
  if (hs.id == -2) {
    hit.id = -2;
    hit.point = to_fast(hs.point);
    hit.normal = to_fast(hs.normal);
    hit.color = hs.color;
  }
}
__device__ __forceinline__ void bounce_ray_fixed(bool mirror, V3<float> dir, V3<float> normal, V3<float> point, float &medium, V3<float> &start,
                                                 V3<float> &ndir) {
  V3<sfloat> s, d;
  if (mirror) reflect_ray<sfloat>(to_strict(dir), to_strict(normal), to_strict(point), s, d, medium);
  else refract_ray<sfloat>(to_strict(dir), to_strict(normal), to_strict(point), medium, s, d, medium);
  start = to_fast(s);
  ndir = to_fast(d);
}

// One candidate of the fast policy: (t, u, v) of triangle i for the ray (start, dir), or false.  The inside test runs on
// sign-corrected triple products, t = ts / |dn|.  Two plane tests that cost one 16-byte load and six FMAs come first:
//   dn = d.N, bn = (o - v0).N = o.N - v0.N:   t = -bn/dn < 0 (the plane lies behind the ray)  -> no candidate
//   |bn| > t_limit |dn| (the plane is crossed beyond t_limit, with a relative margin far above rounding) -> no candidate
// (t_limit = the current best of a sequential scan, which would reject the candidate anyway; RT_MAXFLOAT = no limit).
__device__ __forceinline__ bool bounce_candidate(const SceneView &sc, int i, V3<float> start, V3<float> dir, float t_limit, float &t, float &u,
                                                 float &v) {
  // every operation spelled out (fmaf / __fmul_rn / __fsub_rn): this function is inlined into the lane-parallel and the
  // warp-cooperative search, and the two copies must round alike — the frame may not depend on which one a ray took
  const float4 P = sc.tnd[i];
  const float dn = fmaf(dir.x, P.x, fmaf(dir.y, P.y, __fmul_rn(dir.z, P.z)));
  const float bn = fmaf(start.x, P.x, fmaf(start.y, P.y, fmaf(start.z, P.z, -P.w)));
  const unsigned sb = __float_as_uint(dn) & 0x80000000u;
  const float ts = xor_sign(-bn, sb), adn = fabsf(dn);
  if (!(ts >= 0.0f) || ts > __fmul_rn(__fmul_rn(t_limit, adn), 1.0001f)) return false;  // (NaN rays fail the first test, as they fail the full one)
  const float4 A = sc.ta[i], Bq = sc.tb[i], C = sc.tc[i];
  const float bx = __fsub_rn(start.x, A.x), by = __fsub_rn(start.y, A.y), bz = __fsub_rn(start.z, A.z);
  const float qx = fmaf(by, dir.z, -__fmul_rn(bz, dir.y)), qy = fmaf(bz, dir.x, -__fmul_rn(bx, dir.z)), qz = fmaf(bx, dir.y, -__fmul_rn(by, dir.x));
  const float eu = fmaf(C.x, qx, fmaf(C.y, qy, __fmul_rn(C.z, qz))), ev = fmaf(Bq.x, qx, fmaf(Bq.y, qy, __fmul_rn(Bq.z, qz)));
  const float us = xor_sign(-eu, sb), vs = xor_sign(ev, sb);
  if (!((us >= 0.0f) & (vs >= 0.0f) & (__fadd_rn(us, vs) <= adn))) return false;
  const float inv = __frcp_rn(adn);
  t = __fmul_rn(ts, inv);
  u = __fmul_rn(us, inv);
  v = __fmul_rn(vs, inv);
  return true;
}

// (v0 + u e1) + v e2 of the winning triangle (kernels.cl:124), fast policy, spelled out for the same reason
__device__ __forceinline__ V3<float> bounce_hit_point(const SceneView &sc, int i, float u, float v) {
  const float4 A = sc.ta[i], Bq = sc.tb[i], C = sc.tc[i];
  return V3<float>(fmaf(v, C.x, fmaf(u, Bq.x, A.x)), fmaf(v, C.y, fmaf(u, Bq.y, A.y)), fmaf(v, C.z, fmaf(u, Bq.z, A.z)));
}

// Lane-parallel search.  Fast policy: a bounce ray inside the box faces about half the planes and, once it has a hit, most
// of the rest lie behind it; the rays of a warp leave neighbouring points of a sphere, so they mostly agree and only the
// survivors of the plane tests load the vertices.  Strict policy: the reference's loop.
__device__ __forceinline__ void closest_triangles(const SceneView &sc, V3<float> start, V3<float> dir, ClosestState<float> &cs) {
  for (int i = 0; i < sc.n; i++) {
    float t, u, v;
    if (bounce_candidate(sc, i, start, dir, cs.t, t, u, v) && t < cs.t) {  // strict '<', ascending index: lowest index wins a tie
      cs.id = i;
      cs.t = t;
      cs.u = u;
      cs.v = v;
    }
  }
}
__device__ __forceinline__ void closest_triangles(const SceneView &sc, V3<sfloat> start, V3<sfloat> dir, ClosestState<sfloat> &cs) {
  const V3<sfloat> nd = -dir;
  for (int i = 0; i < sc.n; i++) closest_tri_test<sfloat, true>(sc.ta[i], sc.tb[i], sc.tc[i], start, nd, i, i, cs);
}

// Warp-cooperative search — ray compaction across bounces: after the first bounces only a few lanes of a warp still hold
// a live mirror / glass ray (`rays`: their ballot), and a lane scanning all triangles alone is the serial chain a small
// launch ends up waiting for.  So the live rays are taken one at a time: the owner broadcasts origin and direction
// (shuffles), every lane of the warp tests ITS share of the triangles with the same arithmetic as the lane-parallel
// search, and two integer REDUX reductions pick the winner — smallest t (a non-negative float orders like its bits),
// then smallest triangle index among equal t, which is the reference's "strict <, ascending index" rule (kernels.cl:120).
// (u, v, t) travel back to the owner by shuffle.  All lanes of warp_mask must call this together.
template <class T>
__device__ __forceinline__ void closest_triangles_coop(const SceneView &sc, unsigned warp_mask, unsigned rays, V3<T> start, V3<T> dir,
                                                       ClosestState<T> &cs) {
  const int lane = threadIdx.x & 31;
  const int n_act = __popc(warp_mask), rank = __popc(warp_mask & ((1u << lane) - 1u));
  while (rays) {
    const int s = __ffs(rays) - 1;
    rays &= rays - 1u;
    const V3<T> o(T(__shfl_sync(warp_mask, raw(start.x), s)), T(__shfl_sync(warp_mask, raw(start.y), s)), T(__shfl_sync(warp_mask, raw(start.z), s)));
    const V3<T> d(T(__shfl_sync(warp_mask, raw(dir.x), s)), T(__shfl_sync(warp_mask, raw(dir.y), s)), T(__shfl_sync(warp_mask, raw(dir.z), s)));
    ClosestState<T> my;
    my.reset();
    if constexpr (is_strict<T>::value) {
      const V3<T> nd = -d;
      for (int i = rank; i < sc.n; i += n_act) closest_tri_test<T, true>(sc.ta[i], sc.tb[i], sc.tc[i], o, nd, i, i, my);
    } else {
      for (int i = rank; i < sc.n; i += n_act) {
        float t, u, v;
        if (bounce_candidate(sc, i, o, d, RT_MAXFLOAT, t, u, v) && t < my.t) {
          my.id = i;
          my.t = t;
          my.u = u;
          my.v = v;
        }
      }
    }
    // a candidate has 0 <= t < MAXFLOAT (never NaN); -0 orders with +0
    const unsigned key = (my.id < 0) ? 0xffffffffu : ((raw(my.t) == 0.0f) ? 0u : __float_as_uint(raw(my.t)));
    const unsigned kmin = __reduce_min_sync(warp_mask, key);
    if (kmin == 0xffffffffu) continue;  // no triangle hit (warp-uniform)
    const unsigned idk = (key == kmin) ? (unsigned)my.id : 0xffffffffu;
    const unsigned imin = __reduce_min_sync(warp_mask, idk);
    const int w = __ffs(__ballot_sync(warp_mask, idk == imin)) - 1;
    const float wt = __shfl_sync(warp_mask, raw(my.t), w), wu = __shfl_sync(warp_mask, raw(my.u), w), wv = __shfl_sync(warp_mask, raw(my.v), w);
    if (lane == s) {
      cs.id = (int)imin;
      cs.slot = (int)imin;
      cs.t = T(wt);
      cs.u = T(wu);
      cs.v = T(wv);
    }
  }
}

// The winner's hit record, then the two spheres (kernels.cl:124-127, :132-163)
template <class T>
__device__ __forceinline__ void finish_closest(const SceneView &sc, V3<T> start, V3<T> dir, const ClosestState<T> &cs, HitRec<T> &hit) {
  if (cs.id >= 0) {
    hit.id = cs.id;
    if constexpr (is_strict<T>::value) hit.point = hit_point<T>(sc.ta[cs.id], sc.tb[cs.id], sc.tc[cs.id], cs.u, cs.v);
    else hit.point = bounce_hit_point(sc, cs.id, cs.u, cs.v);
    hit.normal = xyz<T>(sc.tn[cs.id]);
    hit.color = sc.tcol[cs.id];
  }
  if constexpr (is_strict<T>::value) closest_spheres<T>(start, dir, cs.t, hit);
  else closest_spheres_fixed(start, dir, cs.t, hit);
}

// Closest hit of a bounce ray, fast policy (kernels.cl:168-241), lane-parallel, in one piece (the form the full-size
// launches use: it compiles to fewer live registers than closest_triangles + finish_closest).  Same decisions as closest_tri_test<float> — the inside
// test on sign-corrected triple products, t = ts / |dn| compared with strict '<' in ascending index order — behind two
// plane tests that cost one 16-byte load and six FMAs per triangle:  dn = d.N, bn = (o - v0).N = o.N - v0.N,
//   t = -bn/dn < 0        (the plane lies behind the ray)              -> skip
//   |bn| > best_t |dn|    (the plane is crossed beyond the current hit, with a relative margin far above rounding) -> skip
// A bounce ray inside the box faces about half the planes and, once it has a hit, most of the rest lie behind it; the
// rays of a warp leave neighbouring points of a sphere, so they mostly agree.  Only survivors load the vertices.
__device__ __forceinline__ void closest_hit_bounce(const SceneView &sc, V3<float> start, V3<float> dir, HitRec<float> &hit) {
  int best = -1;
  float bt = RT_MAXFLOAT, bu = 0.0f, bv = 0.0f;
  for (int i = 0; i < sc.n; i++) {
    float t, u, v;
    if (bounce_candidate(sc, i, start, dir, bt, t, u, v) && t < bt) {  // strict '<', ascending index: lowest index wins a tie
      best = i;
      bt = t;
      bu = u;
      bv = v;
    }
  }
  if (best >= 0) {
    hit.id = best;
    hit.point = bounce_hit_point(sc, best, bu, bv);
    hit.normal = xyz<float>(sc.tn[best]);
    hit.color = sc.tcol[best];
  }
  closest_spheres_fixed(start, dir, bt, hit);
}

// The S jitters of a pixel: they depend on the pixel id only (kernels.cl:319,331).
template <int CH> struct Jitters {  // in registers
  static constexpr bool kPacked = false;
  static constexpr bool kIndexable = false;  // a loop over the samples must be fully unrolled (no dynamic register index)
  float x[CH], y[CH], z[CH];
  __device__ __forceinline__ float jx(int k) const { return x[k]; }
  __device__ __forceinline__ float jy(int k) const { return y[k]; }
  __device__ __forceinline__ float jz(int k) const { return z[k]; }
};
// The same in shared memory, one column per thread with a stride of one block so that the threads of a warp read
// consecutive words: the jitters are only touched once per shading point (|d_k|^2) and by the few (point, triangle) pairs
// that survive the culls, so they need not occupy 3*CH registers.  Even CH: samples are stored in PAIRS, component-major —
// float2 slot (c*CH/2 + k/2)*STRIDE + thread holds (j_c[k], j_c[k+1]) — so one LDS.64 feeds one packed fma.rn.f32x2
// (two samples per issue slot, rt_fast.cuh: shadow_lit_count).  Odd CH: one float per slot, (3k + c)*STRIDE + thread.
template <int CH, int STRIDE> struct JittersShared {
#ifdef RT_F32X2  // A/B switch (measured: no gain on B200 — the per-sample loop is 4 % of the instructions — and the thirteen
                 // duplicated operand pairs cost registers the kernel does not have)
  static constexpr bool kPacked = (CH % 2) == 0;
#else
  static constexpr bool kPacked = false;
#endif
  static constexpr bool kIndexable = true;
  float *p;  // this thread's column: base + thread (odd CH) or base + 2*thread (even CH)
  __device__ __forceinline__ static float *column(float *base, int thread) { return base + (kPacked ? 2 * thread : thread); }
  __device__ __forceinline__ int at(int c, int k) const { return kPacked ? 2 * STRIDE * (c * (CH / 2) + (k >> 1)) + (k & 1) : (3 * k + c) * STRIDE; }
  __device__ __forceinline__ float jx(int k) const { return p[at(0, k)]; }
  __device__ __forceinline__ float jy(int k) const { return p[at(1, k)]; }
  __device__ __forceinline__ float jz(int k) const { return p[at(2, k)]; }
  // samples 2*k2 and 2*k2 + 1 of one component (kPacked only)
  __device__ __forceinline__ float2 x2(int k2) const { return *reinterpret_cast<const float2 *>(p + 2 * STRIDE * (0 * (CH / 2) + k2)); }
  __device__ __forceinline__ float2 y2(int k2) const { return *reinterpret_cast<const float2 *>(p + 2 * STRIDE * (1 * (CH / 2) + k2)); }
  __device__ __forceinline__ float2 z2(int k2) const { return *reinterpret_cast<const float2 *>(p + 2 * STRIDE * (2 * (CH / 2) + k2)); }
};

__device__ __forceinline__ void seed_rng(int global_id, uint32_t &rx, uint32_t &ry, uint32_t &rz) {
  rx = xorshift32((uint32_t)global_id);
  ry = xorshift32(__float2uint_rz(__fmul_rn(__int2float_rn(global_id), 91.0f)));
  rz = xorshift32(__float2uint_rz(__fmul_rn(__int2float_rn(global_id), 19.0f)));
}

template <int CH, bool STRICT = false>
__device__ __forceinline__ void make_jitters(uint32_t &rx, uint32_t &ry, uint32_t &rz, Jitters<CH> &j) {
  typedef typename std::conditional<STRICT, sfloat, float>::type T;  // STRICT: crush's own operation sequence
#pragma unroll
  for (int k = 0; k < CH; k++) {
    rx = xorshift32(rx);
    ry = xorshift32(ry);
    rz = xorshift32(rz);
    j.x[k] = raw(crush1<T>(rx, RT_LIGHT_SPREAD));
    j.y[k] = raw(crush1<T>(ry, RT_LIGHT_SPREAD));
    j.z[k] = raw(crush1<T>(rz, RT_LIGHT_SPREAD));
  }
}

// The pixel's CH jitters generated straight into its shared-memory column (no register array in between: the compiler
// parked that one in local memory).  Same sequence as make_jitters.
template <int CH, int STRIDE, bool STRICT>
__device__ __forceinline__ void make_jitters_shared(int global_id, const JittersShared<CH, STRIDE> &jit) {
  typedef typename std::conditional<STRICT, sfloat, float>::type T;
  uint32_t rx, ry, rz;
  seed_rng(global_id, rx, ry, rz);
  if constexpr (JittersShared<CH, STRIDE>::kPacked) {
#pragma unroll 1
    for (int k2 = 0; k2 < CH / 2; k2++) {
      float2 vx, vy, vz;
      rx = xorshift32(rx);
      ry = xorshift32(ry);
      rz = xorshift32(rz);
      vx.x = raw(crush1<T>(rx, RT_LIGHT_SPREAD));
      vy.x = raw(crush1<T>(ry, RT_LIGHT_SPREAD));
      vz.x = raw(crush1<T>(rz, RT_LIGHT_SPREAD));
      rx = xorshift32(rx);
      ry = xorshift32(ry);
      rz = xorshift32(rz);
      vx.y = raw(crush1<T>(rx, RT_LIGHT_SPREAD));
      vy.y = raw(crush1<T>(ry, RT_LIGHT_SPREAD));
      vz.y = raw(crush1<T>(rz, RT_LIGHT_SPREAD));
      *reinterpret_cast<float2 *>(jit.p + 2 * STRIDE * (0 * (CH / 2) + k2)) = vx;
      *reinterpret_cast<float2 *>(jit.p + 2 * STRIDE * (1 * (CH / 2) + k2)) = vy;
      *reinterpret_cast<float2 *>(jit.p + 2 * STRIDE * (2 * (CH / 2) + k2)) = vz;
    }
  } else {
    // (two samples per trip: fully unrolled, the compiler computes all 3 CH values first and parks them in local memory)
#pragma unroll 2
    for (int k = 0; k < CH; k++) {
      rx = xorshift32(rx);
      ry = xorshift32(ry);
      rz = xorshift32(rz);
      jit.p[jit.at(0, k)] = raw(crush1<T>(rx, RT_LIGHT_SPREAD));
      jit.p[jit.at(1, k)] = raw(crush1<T>(ry, RT_LIGHT_SPREAD));
      jit.p[jit.at(2, k)] = raw(crush1<T>(rz, RT_LIGHT_SPREAD));
    }
  }
}

// Number of UNOCCLUDED samples among the CH shadow rays start + t (r + j_k)  (in_shadow, kernels.cl:243-311).
// The directions d_k = r + j_k are never materialised: d_k.X = r.X + j_k.X with r.X once per (point, triangle).
// Code size.  The dominant kernel is ~5 000 SASS instructions (79 KB) against a 32 KB instruction cache per SM, and its
// warps run in different phases: ncu shows ~10 % of the stall samples as "no instruction".  Rare paths therefore stay
// rolled: the mirror sphere's per-sample shadow test below was 550 instructions (11 % of the kernel) fully unrolled and
// is executed for the few points whose ray cone touches the sphere — rolled: cfg2 -1.2 %, cfg3 -2.3 % (same frames).
// -DRT_SPHERE_SHADOW_ROLLED=0 restores the unrolled form.
#ifndef RT_SPHERE_SHADOW_ROLLED
#define RT_SPHERE_SHADOW_ROLLED 1
#endif
// Registers.  The kernel is compiled for 80 registers (three blocks per SM) and the shadow loop is where it needs most.
// Until round 2 it cached |d_k|^2 of the CH samples of a shading point in CH registers, filled on first use; recomputing
// the value where it is needed (three adds, one multiply, two FMAs on jitters that are in shared memory anyway — the same
// operations, hence the same bits) frees them: cfg2 0.2002 -> 0.1916 ms, cfg3 2.302 -> 2.111 ms, HEAD 0.205 -> 0.188 ms.
// -DRT_DD_CACHE=1 restores the cache.
#ifndef RT_DD_CACHE
#define RT_DD_CACHE 0
#endif
#ifndef RT_SAMPLE_UNROLL  // A/B switch: unroll factor of the per-sample triangle test (0 = fully)
#define RT_SAMPLE_UNROLL 0
#endif
constexpr int kSampleUnroll = RT_SAMPLE_UNROLL;
constexpr bool kSphereShadowRolled = RT_SPHERE_SHADOW_ROLLED;
template <int CH, class J>
__device__ __forceinline__ int shadow_lit_count(const FastScene &sc, V3<float> start, V3<float> r, float radius_sq,
                                                const J &j, unsigned valid_mask) {
  constexpr unsigned FULL = (CH >= 32) ? 0xffffffffu : ((1u << CH) - 1u);
  unsigned occ = ~valid_mask & FULL;  // padding samples of a ragged chunk count as occluded (ignored by the caller)
  // |d_k|^2, filled on first use: most points never reach a per-sample test.  kPacked: two samples per register pair,
  // evaluated with the packed FP32 instructions of sm_100 (fma.rn.f32x2: one issue slot, two samples).
  constexpr bool P = J::kPacked;
  float dd[P ? 1 : CH];
  float2 dd2[P ? CH / 2 : 1];
  bool have_dd = false;
  auto need_dd = [&]() {
#if !RT_DD_CACHE
    if constexpr (!P) return;
#endif
    if (have_dd) return;
    have_dd = true;
    if constexpr (P) {
      const float2 rx2 = make_float2(r.x, r.x), ry2 = make_float2(r.y, r.y), rz2 = make_float2(r.z, r.z);
#pragma unroll
      for (int k2 = 0; k2 < CH / 2; k2++) {
        const float2 ddx = __fadd2_rn(rx2, j.x2(k2)), ddy = __fadd2_rn(ry2, j.y2(k2)), ddz = __fadd2_rn(rz2, j.z2(k2));
        dd2[k2] = __ffma2_rn(ddz, ddz, __ffma2_rn(ddx, ddx, __fmul2_rn(ddy, ddy)));
      }
    } else {
#pragma unroll
      for (int k = 0; k < CH; k++) {
        const float ddx = r.x + j.jx(k), ddy = r.y + j.jy(k), ddz = r.z + j.jz(k);
        dd[k] = fmaf(ddz, ddz, fmaf(ddx, ddx, __fmul_rn(ddy, ddy)));
      }
    }
  };
  auto dd_of = [&](int k) -> float {
    if constexpr (P) return (k & 1) ? dd2[k >> 1].y : dd2[k >> 1].x;
#if !RT_DD_CACHE  // |d_k|^2 recomputed where it is used (same operations)
    else {
      const float ddx = r.x + j.jx(k), ddy = r.y + j.jy(k), ddz = r.z + j.jz(k);
      return fmaf(ddz, ddz, fmaf(ddx, ddx, __fmul_rn(ddy, ddy)));
    }
#else
    else return dd[k];
#endif
  };
  // |r| / |d_s| <= R / (R - jmax); no bound (k huge) when the light is closer than 2 jmax
  const float R = sqrt_approx(radius_sq);
  const float inv_r2 = rcp_approx(radius_sq);
  const float kk = (R > 2.0f * kJitterMax) ? kSlack * R * rcp_approx(R - kJitterMax) : 1e30f;

  for (int l = 0; l < sc.n_clist; l++) {
    const float4 *q = sc.shad + 4 * sc.clist[l];
    const float4 Q0 = q[0], Q1 = q[1];
    const V3<float> b(start.x - Q0.x, start.y - Q0.y, start.z - Q0.z);
    const float c0 = Q1.x, c1 = Q1.y, c2 = Q1.z;
    // (everything a per-sample decision or the pixel value depends on is spelled out — fmaf / __fmul_rn — so that every
    // instantiation of this function rounds alike: the frame may not depend on which kernel a tile went through)
    const float num = fmaf(b.z, c2, fmaf(b.x, c0, -__fmul_rn(b.y, c1)));  // det[b,e1,e2] = b.N
    const float rN = fmaf(r.z, c2, fmaf(r.x, c0, -__fmul_rn(r.y, c1)));   // r.N   (det A of sample k = -(rN + j_k.N))
    const float w = xor_sign(rN, __float_as_uint(num) & 0x80000000u);
    if (fabsf(num) >= (Q0.w - w) * kk) continue;  // plane cull: no sample passes stage 1
    const float4 Q2 = q[2], Q3 = q[3];
    const V3<float> e1(Q2.x, Q2.y, Q2.z), e2(Q3.x, Q3.y, Q3.z);
    const V3<float> U(fmaf(b.y, e2.z, -__fmul_rn(b.z, e2.y)), fmaf(b.z, e2.x, -__fmul_rn(b.x, e2.z)), fmaf(b.x, e2.y, -__fmul_rn(b.y, e2.x)));  // b x e2
    const V3<float> V(fmaf(e1.y, b.z, -__fmul_rn(e1.z, b.y)), fmaf(e1.z, b.x, -__fmul_rn(e1.x, b.z)), fmaf(e1.x, b.y, -__fmul_rn(e1.y, b.x)));  // e1 x b
    const float rU = fmaf(r.z, U.z, fmaf(r.y, U.y, __fmul_rn(r.x, U.x))), rV = fmaf(r.z, V.z, fmaf(r.y, V.y, __fmul_rn(r.x, V.x)));
    const float arN = fabsf(rN);
    if (arN > Q0.w) {
      // every sample has sign(dn) = sign(rN): edge cull on the un-jittered ray
      const float lb = sqrt_approx(dot(b, b));
      const float mU = lb * Q3.w, mV = lb * Q2.w;
      const unsigned sg = __float_as_uint(rN) & 0x80000000u;
      const float eu = xor_sign(rU, sg), ev = xor_sign(rV, sg);
      if ((eu < -mU) | (ev < -mV) | ((eu + ev) - (mU + mV) > arN + Q0.w)) continue;
    }
    const float q1 = __fmul_rn(__fmul_rn(num, num), inv_r2);  // t^2 |d|^2 < r^2  <=>  q1 |d|^2 < dn^2
    const unsigned numb = __float_as_uint(num);
    need_dd();
    // det A = -dn;  t = -num/dn;  u = E1/dn;  v = E2/dn
    // sign bit of sx clear  <=>  sign(E1) == sign(dn) && sign(E2) == sign(dn) && sign(num) != sign(dn)
    //                       <=>  u >= 0 && v >= 0 && t >= 0   (exact zeros aside)
    if constexpr (P) {
      const float2 c0p = make_float2(c0, c0), c1n = make_float2(-c1, -c1), c2p = make_float2(c2, c2), rNp = make_float2(rN, rN);
      const float2 Ux = make_float2(U.x, U.x), Uy = make_float2(U.y, U.y), Uz = make_float2(U.z, U.z), rUp = make_float2(rU, rU);
      const float2 Vx = make_float2(V.x, V.x), Vy = make_float2(V.y, V.y), Vz = make_float2(V.z, V.z), rVp = make_float2(rV, rV);
      const float2 q1p = make_float2(q1, q1);
#pragma unroll
      for (int k2 = 0; k2 < CH / 2; k2++) {
        const float2 jx = j.x2(k2), jy = j.y2(k2), jz = j.z2(k2);
        const float2 dn = __ffma2_rn(jx, c0p, __ffma2_rn(jy, c1n, __ffma2_rn(jz, c2p, rNp)));
        const float2 E1 = __ffma2_rn(jx, Ux, __ffma2_rn(jy, Uy, __ffma2_rn(jz, Uz, rUp)));
        const float2 E2 = __ffma2_rn(jx, Vx, __ffma2_rn(jy, Vy, __ffma2_rn(jz, Vz, rVp)));
        const float2 es = __fadd2_rn(E1, E2), lhs = __fmul2_rn(q1p, dd2[k2]), rhs = __fmul2_rn(dn, dn);
        {
          const unsigned dnb = __float_as_uint(dn.x);
          const unsigned sx = ((__float_as_uint(E1.x) ^ dnb) | (__float_as_uint(E2.x) ^ dnb)) | ~(numb ^ dnb);
          const bool hit = ((int)sx >= 0) & (fabsf(es.x) <= fabsf(dn.x)) & (lhs.x < rhs.x);
          occ |= hit ? (1u << (2 * k2)) : 0u;
        }
        {
          const unsigned dnb = __float_as_uint(dn.y);
          const unsigned sx = ((__float_as_uint(E1.y) ^ dnb) | (__float_as_uint(E2.y) ^ dnb)) | ~(numb ^ dnb);
          const bool hit = ((int)sx >= 0) & (fabsf(es.y) <= fabsf(dn.y)) & (lhs.y < rhs.y);
          occ |= hit ? (2u << (2 * k2)) : 0u;
        }
      }
    } else {
#pragma unroll((kSampleUnroll > 0 && J::kIndexable) ? kSampleUnroll : CH)
      for (int k = 0; k < CH; k++) {
        const float jx = j.jx(k), jy = j.jy(k), jz = j.jz(k);
        const float dn = fmaf(jx, c0, fmaf(-jy, c1, fmaf(jz, c2, rN)));
        const float E1 = fmaf(jx, U.x, fmaf(jy, U.y, fmaf(jz, U.z, rU)));
        const float E2 = fmaf(jx, V.x, fmaf(jy, V.y, fmaf(jz, V.z, rV)));
        const unsigned dnb = __float_as_uint(dn);
        const unsigned sx = ((__float_as_uint(E1) ^ dnb) | (__float_as_uint(E2) ^ dnb)) | ~(numb ^ dnb);
        const bool hit = ((int)sx >= 0) & (fabsf(__fadd_rn(E1, E2)) <= fabsf(dn)) & (__fmul_rn(q1, dd_of(k)) < __fmul_rn(dn, dn));
        occ |= hit ? (1u << k) : 0u;
      }
    }
    if (occ == FULL) return 0;
  }
#pragma unroll
  for (int i = 0; i < RT_SPHERES; i++) {
    if (c_sphere_color[i].w == -1.0f) continue;  // glass casts no shadow (kernels.cl:279)
    const float4 cr = c_sphere_center_r2[i];
    const V3<float> L(start.x - cr.x, start.y - cr.y, start.z - cr.z);
    const float LL = fmaf(L.z, L.z, fmaf(L.y, L.y, __fmul_rn(L.x, L.x)));
    const float c = __fsub_rn(LL, cr.w);
    // Cone cull: the distance from the centre to the line (start, d_s) is at least
    // perp_c - |L| |d_s/|d_s| - r/|r||  >=  perp_c - |L| jmax/(R - jmax); no real root if that exceeds the radius.
    const float Lr = fmaf(L.z, r.z, fmaf(L.y, r.y, __fmul_rn(L.x, r.x)));
    const float perp2 = LL - Lr * Lr * inv_r2;
    const float lim = kSlack * sqrt_approx(cr.w) + sqrt_approx(LL) * (kk * kJitterMax * rcp_approx(R));
    if (perp2 > lim * lim) continue;
    // (jitters in shared memory: a rolled loop, |d_k|^2 recomputed from the jitters with the same operations — see above)
    constexpr bool kRolled = kSphereShadowRolled && J::kIndexable && !P;
    if constexpr (!kRolled) need_dd();
#pragma unroll(kRolled ? 1 : CH)
    for (int k = 0; k < CH; k++) {
      if ((occ >> k) & 1u) continue;
      float a;
      if constexpr (kRolled) {
        const float ddx = r.x + j.jx(k), ddy = r.y + j.jy(k), ddz = r.z + j.jz(k);
        a = fmaf(ddz, ddz, fmaf(ddx, ddx, __fmul_rn(ddy, ddy)));
      } else {
        a = dd_of(k);
      }
      const float b = __fmul_rn(2.0f, fmaf(j.jx(k), L.x, fmaf(j.jy(k), L.y, fmaf(j.jz(k), L.z, Lr))));
      const float disc = fmaf(b, b, -__fmul_rn(__fmul_rn(4.0f, a), c));
      if (disc < 0.0f) continue;
      const float sq = sqrt_approx(disc);
      const float qq = (b > 0.0f) ? __fmul_rn(-0.5f, __fadd_rn(b, sq)) : __fmul_rn(-0.5f, __fsub_rn(b, sq));
      const float x0 = __fmul_rn(qq, rcp_approx(a));
      const float x1 = __fmul_rn(c, rcp_approx(qq));
      const float x_min = fminf(x0, x1), x_max = fmaxf(x0, x1);
      // |x d|^2 < r^2
      if ((x_min >= 0.0f && __fmul_rn(__fmul_rn(x_min, x_min), a) < radius_sq) || (x_max >= 0.0f && __fmul_rn(__fmul_rn(x_max, x_max), a) < radius_sq)) occ |= 1u << k;
    }
  }
  return CH - __popc(occ);
}

// direct_light (kernels.cl:313-340) for one shading point.  SINGLE (S == CH): jit holds the pixel's
// S jitters, generated once per pixel.  Otherwise they are regenerated chunk by chunk from the seed.
template <int CH, bool SINGLE, class J>
__device__ __forceinline__ float direct_light_fast(const FastScene &sc, V3<float> point, V3<float> normal, V3<float> light_pos, int S,
                                                   int global_id, const J &jit) {
  const V3<float> r = light_pos - point;
  const V3<float> start(fmaf(RT_BIAS, r.x, point.x), fmaf(RT_BIAS, r.y, point.y), fmaf(RT_BIAS, r.z, point.z));
  const float radius_sq = fmaf(r.z, r.z, fmaf(r.y, r.y, __fmul_rn(r.x, r.x)));  // spelled out: see shadow_lit_count
  const float lam = __fmul_rn(RT_LIGHT_COLOR, fmaxf(fmaf(r.z, normal.z, fmaf(r.y, normal.y, __fmul_rn(r.x, normal.x))), 0.0f));
  // facing away from the light: the result is lit*0/den = 0 whatever the shadow rays find
  if (lam == 0.0f && radius_sq > 0.0f && radius_sq < 1.0e37f) return 0.0f;
  int lit = 0;
  if constexpr (SINGLE) {
    lit = shadow_lit_count<CH, J>(sc, start, r, radius_sq, jit, 0xffffffffu);
  } else {
    uint32_t rx, ry, rz;
    seed_rng(global_id, rx, ry, rz);
#pragma unroll 1
    for (int s0 = 0; s0 < S; s0 += CH) {
      Jitters<CH> jj;
      make_jitters<CH>(rx, ry, rz, jj);
      const unsigned valid = (s0 + CH > S) ? ((1u << (S - s0)) - 1u) : 0xffffffffu;
      lit += shadow_lit_count<CH, Jitters<CH>>(sc, start, r, radius_sq, jj, valid);
    }
  }
  return __fmul_rn(__fmul_rn((float)lit, lam), rcp_approx(__fmul_rn(__fmul_rn(4.0f * RT_PI_F, radius_sq), (float)S)));
}

// ---------------------------------------------------------------------------------------------
// RT_FLAG_STRICT_IEEE over the same culls.  The culls are conservative — they only skip (point,
// triangle) pairs in which no sample can pass the reference's tests, with margins (kJitterMax,
// kSlack) orders of magnitude above float rounding — so evaluating the surviving pairs with the
// reference's exact operation sequence gives the same occlusion mask, hence the same frame bit for
// bit, as the un-culled strict kernel (tests compare the two and both with the reference frames).
// ---------------------------------------------------------------------------------------------
template <int CH, class J>
__device__ __forceinline__ unsigned shadow_occ_strict(const FastScene &sc, V3<sfloat> start_s, V3<sfloat> r_s, sfloat radius_sq_s,
                                                      const J &j, unsigned valid_mask) {
  typedef sfloat T;
  constexpr unsigned FULL = (CH >= 32) ? 0xffffffffu : ((1u << CH) - 1u);
  unsigned occ = ~valid_mask & FULL;  // padding samples of a ragged chunk: never evaluated, ignored by the caller
  const V3<float> start(start_s.x.v, start_s.y.v, start_s.z.v), r(r_s.x.v, r_s.y.v, r_s.z.v);
  const float radius_sq = radius_sq_s.v;
  const float R = sqrt_approx(radius_sq);
  const float inv_r2 = rcp_approx(radius_sq);
  const float kk = (R > 2.0f * kJitterMax) ? kSlack * R * rcp_approx(R - kJitterMax) : 1e30f;
  for (int l = 0; l < sc.n_clist; l++) {
    const float4 *q = sc.shad + 4 * sc.clist[l];
    const float4 Q0 = q[0], Q1 = q[1];
    {  // the two culls of shadow_lit_count, in fast float arithmetic
      const V3<float> b(start.x - Q0.x, start.y - Q0.y, start.z - Q0.z);
      const float c0 = Q1.x, c1 = Q1.y, c2 = Q1.z;
      const float num = (b.x * c0 - b.y * c1) + b.z * c2;
      const float rN = (r.x * c0 - r.y * c1) + r.z * c2;
      const float w = xor_sign(rN, __float_as_uint(num) & 0x80000000u);
      // |num| at rounding level: its sign (which selects the side of the cull) is not trustworthy — evaluate
      const bool num_ok = fabsf(num) > 1e-5f * (fabsf(b.x * c0) + fabsf(b.y * c1) + fabsf(b.z * c2));
      if (num_ok && fabsf(num) >= (Q0.w - w) * kk) continue;
      const float arN = fabsf(rN);
      if (arN > Q0.w) {
        const float4 Q2 = q[2], Q3 = q[3];
        const V3<float> e1(Q2.x, Q2.y, Q2.z), e2(Q3.x, Q3.y, Q3.z);
        const V3<float> U(b.y * e2.z - b.z * e2.y, b.z * e2.x - b.x * e2.z, b.x * e2.y - b.y * e2.x);
        const V3<float> V(e1.y * b.z - e1.z * b.y, e1.z * b.x - e1.x * b.z, e1.x * b.y - e1.y * b.x);
        const float lb = sqrt_approx(dot(b, b));
        const float mU = lb * Q3.w, mV = lb * Q2.w;
        const unsigned sg = __float_as_uint(rN) & 0x80000000u;
        const float eu = xor_sign(dot(r, U), sg), ev = xor_sign(dot(r, V), sg);
        if ((eu < -mU) | (ev < -mV) | ((eu + ev) - (mU + mV) > arN + Q0.w)) continue;
      }
    }
    // in_shadow's per-triangle sequence (kernels.cl:246-276), as shadow_pair<sfloat>
    const float4 Q2 = q[2], Q3 = q[3];
    const V3<T> v0(T(Q0.x), T(Q0.y), T(Q0.z)), e1(T(Q2.x), T(Q2.y), T(Q2.z)), e2(T(Q3.x), T(Q3.y), T(Q3.z));
    const T c0 = T(Q1.x), c1 = T(Q1.y), c2 = T(Q1.z);
    const V3<T> b = start_s - v0;
    const T detA0 = (b.x * c0 - b.y * c1) + b.z * c2;
    const T U0 = b.y * e2.z - b.z * e2.y, U1 = b.x * e2.z - b.z * e2.x, U2 = b.x * e2.y - b.y * e2.x;
    const T V0 = e1.y * b.z - e1.z * b.y, V1 = e1.x * b.z - e1.z * b.x, V2 = e1.x * b.y - e1.y * b.x;
#pragma unroll
    for (int k = 0; k < CH; k++) {
      if ((occ >> k) & 1u) continue;
      const V3<T> d = r_s + V3<T>(T(j.jx(k)), T(j.jy(k)), T(j.jz(k)));
      const V3<T> nd = -d;
      const T detA = (nd.x * c0 - nd.y * c1) + nd.z * c2;
      const T inv = rcp_(detA);
      const T t = detA0 * inv;
      const V3<T> dv = scale(t, d);
      const T dist = dv.x * dv.x + dv.y * dv.y + dv.z * dv.z;
      if (t >= T(0.0f) && dist < radius_sq_s) {
        const T u = ((nd.x * U0 - nd.y * U1) + nd.z * U2) * inv;
        const T v = ((nd.x * V0 - nd.y * V1) + nd.z * V2) * inv;
        if (u >= T(0.0f) && v >= T(0.0f) && (u + v) <= T(1.0f)) occ |= 1u << k;
      }
    }
    if (occ == FULL) return occ;
  }
#pragma unroll
  for (int i = 0; i < RT_SPHERES; i++) {
    if (c_sphere_color[i].w == -1.0f) continue;  // glass casts no shadow (kernels.cl:279)
    const float4 cr = c_sphere_center_r2[i];
    {  // cone cull of shadow_lit_count
      const V3<float> L(start.x - cr.x, start.y - cr.y, start.z - cr.z);
      const float LL = dot(L, L);
      const float Lr = dot(L, r);
      const float perp2 = LL - Lr * Lr * inv_r2;
      const float lim = kSlack * sqrt_approx(cr.w) + sqrt_approx(LL) * (kk * kJitterMax * rcp_approx(R));
      if (perp2 > lim * lim) continue;
    }
    const V3<T> L = start_s - xyz<T>(cr);
    const T c = dot(L, L) - T(cr.w);
#pragma unroll((kSphereShadowRolled && J::kIndexable) ? 1 : CH)
    for (int k = 0; k < CH; k++) {  // kernels.cl:278-307, as shadow_spheres<sfloat>
      if ((occ >> k) & 1u) continue;
      const V3<T> d = r_s + V3<T>(T(j.jx(k)), T(j.jy(k)), T(j.jz(k)));
      const T a = dot(d, d);
      const T b = T(2.0f) * dot(d, L);
      const T disc = b * b - T(4.0f) * a * c;
      if (disc < T(0.0f)) continue;
      const T sq = sqrt_(disc);
      const T qq = (b > T(0.0f)) ? T(-0.5f) * (b + sq) : T(-0.5f) * (b - sq);
      const T x0 = div_(qq, a);
      const T x1 = div_(c, qq);
      const T x_min = cl_min(x0, x1);
      const T x_max = cl_max(x0, x1);
      const V3<T> min_dir = scale(x_min, d);
      const V3<T> max_dir = scale(x_max, d);
      const T min_dist = dot(min_dir, min_dir);
      const T max_dist = dot(max_dir, max_dir);
      if ((x_min >= T(0.0f) && min_dist < radius_sq_s) || (x_max >= T(0.0f) && max_dist < radius_sq_s)) occ |= 1u << k;
    }
  }
  return occ;
}

// direct_light (kernels.cl:313-340) in the reference's operation sequence (cf. direct_light<sfloat> in rt_brute.cuh);
// the jitters come from the pixel's shared-memory column (SINGLE) or are regenerated per chunk.
template <int CH, bool SINGLE, class J>
__device__ __forceinline__ sfloat direct_light_strict(const FastScene &sc, V3<sfloat> point, V3<sfloat> normal, V3<sfloat> light_pos, int S,
                                                      int global_id, const J &jit) {
  typedef sfloat T;
  const V3<T> dir = light_pos - point;
  const V3<T> start = point + scale(T(RT_BIAS), dir);
  const T radius_sq = (dir.x * dir.x + dir.y * dir.y) + dir.z * dir.z;
  const T lam = T(RT_LIGHT_COLOR) * cl_max(dot(dir, normal), T(0.0f));
  const T den = T(4.0f) * T(RT_PI_F) * radius_sq;
  // facing away from the light: every term mask*lam/den is exactly +0, and so is the sum and total/S
  if (lam.v == 0.0f && den.v > 0.0f && den.v < 3.0e38f) return T(0.0f);
  // mask * lam / den for mask = 1 and mask = 0 (kernels.cl:334-336): the two possible per-sample terms
  const T term_lit = div_(T(1.0f) * lam, den), term_occ = div_(T(0.0f) * lam, den);
  T total = T(0.0f);
  if constexpr (SINGLE) {
    const unsigned occ = shadow_occ_strict<CH, J>(sc, start, dir, radius_sq, jit, 0xffffffffu);
#pragma unroll
    for (int k = 0; k < CH; k++) total = total + (((occ >> k) & 1u) ? term_occ : term_lit);
  } else {
    uint32_t rx, ry, rz;
    seed_rng(global_id, rx, ry, rz);
#pragma unroll 1
    for (int s0 = 0; s0 < S; s0 += CH) {
      Jitters<CH> jj;
      make_jitters<CH, true>(rx, ry, rz, jj);
      const unsigned valid = (s0 + CH > S) ? ((1u << (S - s0)) - 1u) : 0xffffffffu;
      const unsigned occ = shadow_occ_strict<CH, Jitters<CH>>(sc, start, dir, radius_sq, jj, valid);
#pragma unroll
      for (int k = 0; k < CH; k++)
        if (s0 + k < S) total = total + (((occ >> k) & 1u) ? term_occ : term_lit);
    }
  }
  return div_(total, T(__int2float_rn(S)));
}

}  // namespace rt
