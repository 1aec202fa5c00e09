// rt_bvh.cuh — BVH tracer for large meshes (Loader.cpp scenes: ~1 M triangles in the box).
//
// The reference has no acceleration structure: every ray tests every triangle
// (kernels.cl:100, :174, :246), which is ~10^14 tests per 1080p frame at 1 M triangles.
// Here a GPU-built LBVH (rt_bvh.cu) only PRUNES: a leaf hands its triangles to exactly the
// per-triangle tests of the brute-force path (rt_brute.cuh), boxes are padded well beyond
// rounding error, and the reference's tie rule (lowest upload index wins an equal t) is applied
// explicitly, so a BVH frame equals the brute-force frame — bit for bit under RT_FLAG_STRICT_IEEE.
//
// Layout (all in HBM, float4 = one 16-byte load):
//   tri_a/b/c[s]   (v0,c0) (e1,c1) (e2,c2) of the triangle at sorted slot s   (same as SceneView)
//   tri_n/col[s]   normal, colour+material;  tri_id[s] = upload index
//   slots [0, n_bvh) are in Morton order and covered by the tree; slots [n_bvh, n) hold the few
//   "big" triangles (room walls) that would bloat the boxes — they are tested linearly.
//   node j = 4 float4: child0 box min.xyz max.x | child0 max.yz child1 min.xy | child1 min.z max.xyz |
//            (ref0, ref1, -, -) as ints.  ref >= 0: internal node; ref < 0: leaf, ~ref = first<<3 | (count-1).
//
// Shadow rays: the S jittered rays of a shading point share their origin and nearly their
// direction, so they traverse the tree as ONE packet with a per-node mask of live rays: node
// fetches are shared, a leaf runs the same (origin, triangle) x CH-rays test as the brute-force
// path on the rays that reach it.
#pragma once
#include "rt_fast.cuh"

namespace rt {

#ifndef RT_BVH_LEAF
#define RT_BVH_LEAF 4
#endif
constexpr int kBvhLeafMax = RT_BVH_LEAF;  // triangles per leaf (<= 8 by the ref encoding)
constexpr int kBvhStack = 64;

struct BvhView {
  const float4 *tri_a, *tri_b, *tri_c, *tri_n, *tri_col;
  const int *tri_id;
  const float4 *nodes;
  const float4 *big_bound;  // per big triangle (slot - n_bvh): (jmax|N|, jmax|e1|, jmax|e2|, -) for the fast policy's culls
  int n, n_bvh;      // all triangles / those covered by the tree
  int root;          // root reference (may be a leaf)
  int all_casters;   // no triangle has material -1
};

__device__ __forceinline__ float safe_rcp_dir(float d) {
  const float a = fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d;
  return 1.0f / a;
}

// Slab test of the ray o + t*d, t in [0, tmax], against [lo, hi]; returns entry distance or -1.
__device__ __forceinline__ float box_entry(float lox, float loy, float loz, float hix, float hiy, float hiz, float ox, float oy, float oz,
                                           float ix, float iy, float iz, float tmax) {
  const float tx0 = (lox - ox) * ix, tx1 = (hix - ox) * ix;
  const float ty0 = (loy - oy) * iy, ty1 = (hiy - oy) * iy;
  const float tz0 = (loz - oz) * iz, tz1 = (hiz - oz) * iz;
  const float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), 0.0f));
  const float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tmax));
  return (tn <= tf) ? tn : -1.0f;
}

// Shared-memory helper of the fast policy: the big triangles (room walls) seen from the camera, with the
// per-frame constants and per-tile binning of the brute-force fast path (rt_fast.cuh), so that a primary ray
// filters them with three dot products each and confirms survivors strictly.  Slot numbers are relative to
// bv.n_bvh.  n_big < 0: not available (too many big triangles) -> plain strict loop.
struct BvhBigPrimary {
  const float4 *prim;
  const int *plist;
  int n_list;
  int n_big;
};

template <class T> struct BvhTracer {
  static constexpr bool kSkipUnlit = true;  // direct_light: no shadow rays for points facing away from the light
  BvhView bv;
  BvhBigPrimary big{nullptr, nullptr, 0, -1};
  __device__ __forceinline__ BvhTracer<sfloat> strict() const {  // same tree, reference arithmetic
    BvhTracer<sfloat> t;
    t.bv = bv;
    return t;
  }

  // primary ray of the fast policy: reference arithmetic; the big triangles through the camera-constant filter
  __device__ void primary_strict(V3<sfloat> cam, V3<sfloat> dir, HitRec<sfloat> &hit) const {
    BvhTracer<sfloat> t;
    t.bv = bv;
    if (big.n_big < 0) {
      t.closest(cam, dir, hit);
      return;
    }
    FastScene fs;
    fs.prim = big.prim;
    fs.plist = big.plist;
    fs.n_prim = big.n_list;
    int bi;
    float bt, bu, bv_;
    primary_triangles(fs, V3<float>(dir.x.v, dir.y.v, dir.z.v), bi, bt, bu, bv_);  // strict t,u,v of the closest big triangle
    ClosestState<sfloat> cs;
    cs.reset();
    if (bi >= 0) {
      cs.t = sfloat(bt);
      cs.u = sfloat(bu);
      cs.v = sfloat(bv_);
      cs.slot = bv.n_bvh + bi;
      cs.id = bv.tri_id[cs.slot];
    }
    t.closest_from(cam, dir, hit, cs, false);
  }

  // kernels.cl:92-166 / :168-241
  __device__ void closest(V3<T> start, V3<T> dir, HitRec<T> &hit) const {
    ClosestState<T> cs;
    cs.reset();
    closest_from(start, dir, hit, cs, true);
  }

  // continue a closest-hit search from state cs (test_big: the linear list has not been searched yet)
  __device__ void closest_from(V3<T> start, V3<T> dir, HitRec<T> &hit, ClosestState<T> cs, bool test_big) const {
    const V3<T> nd = -dir;
    // the big triangles, linearly
    if (test_big)
      for (int s = bv.n_bvh; s < bv.n; s++)
        closest_tri_test<T, false>(bv.tri_a[s], bv.tri_b[s], bv.tri_c[s], start, nd, bv.tri_id[s], s, cs);
    if (bv.n_bvh > 0) {
      const float ox = raw(start.x), oy = raw(start.y), oz = raw(start.z);
      const float ix = safe_rcp_dir(raw(dir.x)), iy = safe_rcp_dir(raw(dir.y)), iz = safe_rcp_dir(raw(dir.z));
      // "while-while" traversal: a lane walks down internal nodes until it holds a leaf (or runs out of nodes); the lanes
      // of the warp re-converge behind that inner loop and test their leaves TOGETHER.  (With one loop that does either a
      // node step or a leaf step per trip, the two kinds of step alternate lane by lane and each runs with a handful of
      // the 32 lanes: ncu showed 4-9 active threads per instruction in the leaf tests.)
      int stack[kBvhStack];
      int sp = 0;
      int ref = bv.root;
      bool done = false;
      while (!done) {
        while (ref >= 0) {
          const float4 *nd4 = bv.nodes + 4 * (size_t)ref;
          const float4 n0 = __ldg(nd4), n1 = __ldg(nd4 + 1), n2 = __ldg(nd4 + 2), n3 = __ldg(nd4 + 3);
          const float tmax = raw(cs.t);
          const float e0 = box_entry(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, ox, oy, oz, ix, iy, iz, tmax);
          const float e1 = box_entry(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, ox, oy, oz, ix, iy, iz, tmax);
          const int r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
          if (e0 >= 0.0f && e1 >= 0.0f) {
            const bool first0 = e0 <= e1;
            if (sp < kBvhStack) stack[sp++] = first0 ? r1 : r0;
            ref = first0 ? r0 : r1;
          } else if (e0 >= 0.0f) {
            ref = r0;
          } else if (e1 >= 0.0f) {
            ref = r1;
          } else if (sp > 0) {
            ref = stack[--sp];
          } else {
            done = true;
            break;
          }
        }
        if (done) break;
        {
          const int first = (~ref) >> 3, cnt = ((~ref) & 7) + 1;
          for (int k = 0; k < cnt; k++) {
            const int s = first + k;
            closest_tri_test<T, false, true>(__ldg(bv.tri_a + s), __ldg(bv.tri_b + s), __ldg(bv.tri_c + s), start, nd, __ldg(bv.tri_id + s), s, cs);
          }
        }
        if (sp == 0) break;
        ref = stack[--sp];
      }
    }
    if (cs.id >= 0) {
      const int s = cs.slot;
      hit.id = cs.id;
      hit.point = hit_point<T>(bv.tri_a[s], bv.tri_b[s], bv.tri_c[s], cs.u, cs.v);
      hit.normal = xyz<T>(bv.tri_n[s]);
      hit.color = bv.tri_col[s];
    }
    closest_spheres<T>(start, dir, cs.t, hit);
  }

  __device__ __forceinline__ bool casts_shadow(int s) const { return bv.all_casters || __ldg(&bv.tri_col[s].w) != -1.0f; }

  // kernels.cl:243-311 for CH rays start + t (r + j_k): occlusion mask
  template <int CH>
  __device__ unsigned shadow(V3<T> start, const ShadowRays<T, CH> &rays, V3<T> r, T radius_sq) const {
    constexpr unsigned FULL = (CH >= 32) ? 0xffffffffu : ((1u << CH) - 1u);
    unsigned occ = 0u;
    if constexpr (is_strict<T>::value) {
      for (int s = bv.n_bvh; s < bv.n; s++) {
        if (!casts_shadow(s)) continue;
        shadow_pair<T, CH>(bv.tri_a[s], bv.tri_b[s], bv.tri_c[s], start, rays, radius_sq, occ);
        if (occ == FULL) return occ;
      }
    } else {
      // the big triangles are the room: almost all of them are culled per shading point (rt_fast.cuh)
      const float Rb = sqrt_approx(radius_sq), inv_r2 = rcp_approx(radius_sq);
      const float kk = (Rb > 2.0f * kJitterMax) ? kSlack * Rb * rcp_approx(Rb - kJitterMax) : 1e30f;
      for (int s = bv.n_bvh; s < bv.n; s++) {
        if (!casts_shadow(s)) continue;
        shadow_pair_culled<CH>(bv.tri_a[s], bv.tri_b[s], bv.tri_c[s], bv.big_bound[s - bv.n_bvh], start, r, rays, kk, inv_r2, occ);
        if (occ == FULL) return occ;
      }
    }
    if (bv.n_bvh > 0) {
      // Packet traversal: the CH rays share every node fetch and every box test; a leaf tests the rays that are not
      // occluded yet.  A node is dropped when the packet as a whole cannot touch its (padded) box.
      const float ox = raw(start.x), oy = raw(start.y), oz = raw(start.z);
      const float R = sqrtf(raw(radius_sq));
      // The CH rays share their origin and differ by a jitter of a few per cent of |r|: ONE interval-arithmetic slab test
      // per box decides for the whole packet.  Per axis the inverse directions span [lo, hi] (same sign, else the axis
      // puts no constraint); every ray's entry / exit parameter on that axis lies between the smallest and the largest of
      // the four products (box face - origin) x {lo, hi}, so  max_axes(smallest) <= min_axes(largest, t_max)  holds for
      // every ray that touches the box.  Conservative (a few more nodes are visited), a third of the arithmetic of CH
      // separate tests; the leaves test every live ray exactly as before.
      // (An axis on which the direction component changes sign within the packet — typical for the two axes across the
      // beam — bounds no parameter, but the POSITION o + t d does: with t in [0, t_far] it stays within
      // [o + t_far d_min, o + t_far d_max], which has to overlap the box.)
      float ilo[3], ihi[3], dlo[3], dhi[3], tmax_all = 0.0f;
      bool axis_ok[3] = {true, true, true};
      {
        const float big = 3.0e38f;
        dlo[0] = dlo[1] = dlo[2] = big;
        dhi[0] = dhi[1] = dhi[2] = -big;
#pragma unroll
        for (int k = 0; k < CH; k++) {
          const float dx = raw(rays.d[k].x), dy = raw(rays.d[k].y), dz = raw(rays.d[k].z);
          dlo[0] = fminf(dlo[0], dx);
          dhi[0] = fmaxf(dhi[0], dx);
          dlo[1] = fminf(dlo[1], dy);
          dhi[1] = fmaxf(dhi[1], dy);
          dlo[2] = fminf(dlo[2], dz);
          dhi[2] = fmaxf(dhi[2], dz);
          tmax_all = fmaxf(tmax_all, 1.0001f * R * rsqrtf(dx * dx + dy * dy + dz * dz));
        }
#pragma unroll
        for (int a = 0; a < 3; a++) {
          // the component keeps its sign and stays away from zero: its inverse spans [1/d_max, 1/d_min]
          axis_ok[a] = (dlo[a] > 1e-6f) || (dhi[a] < -1e-6f);
          ilo[a] = axis_ok[a] ? 1.0f / dhi[a] : 0.0f;
          ihi[a] = axis_ok[a] ? 1.0f / dlo[a] : 0.0f;
        }
      }
      auto packet_touches = [&](float lox, float loy, float loz, float hix, float hiy, float hiz) -> bool {
        float tn = 0.0f, tf = tmax_all;
        const float o3[3] = {ox, oy, oz}, lo3[3] = {lox, loy, loz}, hi3[3] = {hix, hiy, hiz};
#pragma unroll
        for (int a = 0; a < 3; a++) {
          if (!axis_ok[a]) continue;
          const float l = lo3[a] - o3[a], h = hi3[a] - o3[a];
          const float p1 = l * ilo[a], p2 = l * ihi[a], p3 = h * ilo[a], p4 = h * ihi[a];
          tn = fmaxf(tn, fminf(fminf(p1, p2), fminf(p3, p4)));
          tf = fminf(tf, fmaxf(fmaxf(p1, p2), fmaxf(p3, p4)));
        }
        if (!(tn <= tf)) return false;
        bool ok = true;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          if (axis_ok[a]) continue;
          // positions reached on this axis for t in [0, tf]: between o + tf min(d_min, 0) and o + tf max(d_max, 0)
          const float pmin = o3[a] + tf * fminf(dlo[a], 0.0f), pmax = o3[a] + tf * fmaxf(dhi[a], 0.0f);
          const float pad = 1e-5f * (fabsf(pmin) + fabsf(pmax)) + 1e-7f;
          ok &= (pmin - pad <= hi3[a]) & (pmax + pad >= lo3[a]);
        }
        return ok;
      };
      // while-while, as in closest_from: lanes walk down internal nodes until they hold a leaf, then test their leaves together
      int stack[kBvhStack];
      int sp = 0;
      int ref = bv.root;
      bool done = false;
      while (!done) {
        while (ref >= 0) {  // descend / pop until `ref` is a leaf that the packet touches
          const float4 *nd4 = bv.nodes + 4 * (size_t)ref;
          const float4 n0 = __ldg(nd4), n1 = __ldg(nd4 + 1), n2 = __ldg(nd4 + 2), n3 = __ldg(nd4 + 3);
          const bool m0 = packet_touches(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y);
          const bool m1 = packet_touches(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w);
          const int r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
          if (m0 && m1) {
            if (sp < kBvhStack) stack[sp++] = r1;
            ref = r0;
          } else if (m0) {
            ref = r0;
          } else if (m1) {
            ref = r1;
          } else if (sp > 0) {
            ref = stack[--sp];
          } else {
            done = true;
            break;
          }
        }
        if (done) break;
        {
          const int first = (~ref) >> 3, cnt = ((~ref) & 7) + 1;
          for (int k = 0; k < cnt; k++) {
            const int s = first + k;
            if (!casts_shadow(s)) continue;
            shadow_pair<T, CH>(__ldg(bv.tri_a + s), __ldg(bv.tri_b + s), __ldg(bv.tri_c + s), start, rays, radius_sq, occ, FULL & ~occ);
          }
          if (occ == FULL) return occ;
        }
        if (sp == 0) break;
        ref = stack[--sp];
      }
    }
    shadow_spheres<T, CH>(start, rays, radius_sq, occ);
    return occ;
  }
};

}  // namespace rt
