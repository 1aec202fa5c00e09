// rt_draw_fast.cu — the fast `draw` kernel (see rt_fast.cuh for the method).  Compiled once per
// shadow chunk size: -DRT_FAST_CH=n gives launch_fast_ch<n>.
#include "rt_launch.cuh"

#ifndef RT_FAST_CH
#define RT_FAST_CH 8
#endif

namespace rt {

// CH shadow samples per chunk, SINGLE = (S == CH).
template <int CH, bool SINGLE>
__global__ void __launch_bounds__(kThreads, RT_MINBLOCKS) draw_fast_kernel(const __grid_constant__ FrameParams p,
                                                                           const float4 *__restrict__ scene, int n, int n_sh) {
  extern __shared__ float4 smem[];
  __shared__ int s_warp_count[kThreads / 32];
  __shared__ int s_base;
  __shared__ int s_spheres_visible;

  // ---- stage the scene: generic arrays [0,5n) and the shadow records (global offset 5n+3n_sh) ----
  float4 *const gen = smem;
  float4 *const prim = smem + 5 * n;
  float4 *const shad = prim + 3 * n;
  int *const plist = reinterpret_cast<int *>(shad + 4 * n_sh);
  // per-thread jitter columns (SINGLE only), after the triangle list
  float *const jit_base = reinterpret_cast<float *>(smem + scene_smem_float4(n, n_sh));
  for (int i = threadIdx.x; i < 5 * n; i += kThreads) gen[i] = scene[i];
  for (int i = threadIdx.x; i < 4 * n_sh; i += kThreads) shad[i] = scene[5 * n + 3 * n_sh + i];
  if (threadIdx.x == 0) s_base = 0;
  FastScene sc;
  sc.g.ta = gen;
  sc.g.tb = gen + n;
  sc.g.tc = gen + 2 * n;
  sc.g.tn = gen + 3 * n;
  sc.g.tcol = gen + 4 * n;
  sc.g.sa = sc.g.sb = sc.g.sc = nullptr;  // the SoA shadow arrays belong to the generic kernel
  sc.g.n = n;
  sc.g.n_sh = n_sh;
  sc.prim = prim;
  sc.shad = shad;
  sc.plist = plist;
  const V3<float> cam(p.cam[0], p.cam[1], p.cam[2]), light(p.light[0], p.light[1], p.light[2]);
  const int A = p.A, S = p.S;
  const float SW = (float)p.W, SH = (float)p.H, fA = (float)A;
  int x, y, tile_x, tile_y;
  const bool in_frame = pixel_of_thread(p, x, y, tile_x, tile_y);
  __syncthreads();

  // ---- per-triangle camera constants + binning of the triangles against this block's tile ----
  {
    // corner rays of the tile in virtual (sub-pixel) coordinates, un-normalised (kernels.cl:384-400)
    const float vx0 = (float)(tile_x * A) - SW * fA * 0.5f;
    const float vy0 = (float)(tile_y * A) - SH * fA * 0.5f;
    const float vx1 = vx0 + (float)(kTileW * A - 1), vy1 = vy0 + (float)(kTileH * A - 1);
    V3<float> dc[4];
    float dmax = 0.0f;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const float vx = (c & 1) ? vx1 : vx0, vy = (c & 2) ? vy1 : vy0;
      dc[c] = V3<float>(p.rot[0] * vx + p.rot[1] * vy + p.rot[2] * p.focal, p.rot[3] * vx + p.rot[4] * vy + p.rot[5] * p.focal,
                        p.rot[6] * vx + p.rot[7] * vy + p.rot[8] * p.focal);
      dmax = fmaxf(dmax, sqrtf(dot(dc[c], dc[c])));
    }
    dmax *= 1.001f;
    if (threadIdx.x == 0) {
      // Can any primary ray of the tile hit a sphere?  The tile's rays lie in a cone of half-angle th_t
      // around the centre ray; a sphere is seen from the camera inside a cone of half-angle th_s around
      // the direction to its centre.  No hit if the axes are further apart than th_t + th_s.
      const V3<float> dm((dc[0].x + dc[3].x) * 0.5f, (dc[0].y + dc[3].y) * 0.5f, (dc[0].z + dc[3].z) * 0.5f);
      const float im = rsqrtf(dot(dm, dm));
      float cos_t = 1.0f;
#pragma unroll
      for (int c = 0; c < 4; c++) cos_t = fminf(cos_t, dot(dm, dc[c]) * im * rsqrtf(dot(dc[c], dc[c])));
      cos_t = fminf(cos_t - 1e-5f, 1.0f);
      const float sin_t = sqrtf(fmaxf(1.0f - cos_t * cos_t, 0.0f));
      int vis = 0;
#pragma unroll
      for (int i = 0; i < RT_SPHERES; i++) {
        const float4 cr = c_sphere_center_r2[i];
        const V3<float> L(cr.x - cam.x, cr.y - cam.y, cr.z - cam.z);
        const float l2 = dot(L, L);
        bool may = true;
        if (l2 > cr.w * 1.0001f) {
          const float sin_s = fminf(sqrtf(cr.w / l2) * 1.0001f, 1.0f), cos_s = sqrtf(fmaxf(1.0f - sin_s * sin_s, 0.0f));
          const float cos_sum = (cos_t > 0.0f) ? cos_t * cos_s - sin_t * sin_s : -1.0f;  // th_t + th_s (no cull for huge tiles)
          may = dot(dm, L) * im * rsqrtf(l2) >= cos_sum - 1e-4f;
        }
        vis |= may ? 1 : 0;
      }
      s_spheres_visible = vis;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int base = 0; base < n; base += kThreads) {
      const int i = base + threadIdx.x;
      bool keep = false;
      if (i < n) {
        primary_constants(sc.g, prim, cam, i);
        keep = tile_may_hit(prim, i, dc, dmax);
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_warp_count[warp] = __popc(ballot);
      __syncthreads();
      int offset = s_base, total = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; w++) {
        const int c = s_warp_count[w];
        offset += (w < warp) ? c : 0;
        total += c;
      }
      if (keep) plist[offset + __popc(ballot & ((1u << lane) - 1u))] = i;
      __syncthreads();
      if (threadIdx.x == 0) s_base += total;
    }
    __syncthreads();
    sc.n_prim = s_base;
  }
  const bool spheres_visible = s_spheres_visible != 0;

  if (!in_frame) return;

  const int global_id = __float2int_rz(__fadd_rn(__fmul_rn((float)y, SW), (float)x));  // kernels.cl:380, float arithmetic
  // SINGLE: the pixel's S jitters, generated on the first shading point (44 % of the 1080p frame
  // lies outside the box and never shades) and kept in shared memory
  JittersShared<CH, kThreads> jit;
  jit.p = jit_base + threadIdx.x;
  bool have_jit = false;
  // Primary ray directions with the reference's operation sequence (kernels.cl:384-405): a handful
  // of operations per ray, and it makes the primary hits bit-identical to the reference.
  typedef sfloat SF;
  const V3<SF> base(SF((float)(x * A)) - div_(SF(SW) * SF(fA), SF(2.0f)), SF((float)(y * A)) - div_(SF(SH) * SF(fA), SF(2.0f)), SF(p.focal));
  const V3<SF> r0(SF(p.rot[0]), SF(p.rot[1]), SF(p.rot[2])), r1(SF(p.rot[3]), SF(p.rot[4]), SF(p.rot[5])),
      r2(SF(p.rot[6]), SF(p.rot[7]), SF(p.rot[8]));
  const V3<SF> cam_s(SF(cam.x), SF(cam.y), SF(cam.z));
  V3<float> total(0.0f, 0.0f, 0.0f);
  const int rays = A * A;
  // One primary ray at a time, in the reference's order dy*A + dx (kernels.cl:393-397); the
  // block's binned triangle list is short, so nothing is gained by batching rays.
#pragma unroll 1
  for (int ray_dy = 0; ray_dy < A; ray_dy++) {
#pragma unroll 1
    for (int ray_dx = 0; ray_dx < A; ray_dx++) {
      const V3<SF> d0 = base + V3<SF>(SF((float)ray_dx), SF((float)ray_dy), SF(0.0f));
      const V3<SF> dns = normalize(V3<SF>(dot(r0, d0), dot(r1, d0), dot(r2, d0)));
      V3<float> dir(dns.x.v, dns.y.v, dns.z.v);
      int bi;
      float bt, bu, bv;
      primary_triangles(sc, dir, bi, bt, bu, bv);
      HitRec<float> hit;
      {
        HitRec<SF> hs;
        hs.id = -1;
        hs.color = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
        hs.point = V3<SF>(SF(0.0f), SF(0.0f), SF(0.0f));
        hs.normal = hs.point;
        if (bi >= 0) {
          const V3<SF> v0 = xyz<SF>(sc.g.ta[bi]), e1 = xyz<SF>(sc.g.tb[bi]), e2 = xyz<SF>(sc.g.tc[bi]);
          hs.id = bi;
          hs.point = (v0 + scale(SF(bu), e1)) + scale(SF(bv), e2);  // kernels.cl:124
          hs.normal = xyz<SF>(sc.g.tn[bi]);
          hs.color = sc.g.tcol[bi];
        }
        // the two spheres, strict as well (skipped when no ray of the block's tile can reach one)
        if (spheres_visible) closest_spheres<SF>(cam_s, dns, SF(bt), hs);
        hit.id = hs.id;
        hit.point = V3<float>(hs.point.x.v, hs.point.y.v, hs.point.z.v);
        hit.normal = V3<float>(hs.normal.x.v, hs.normal.y.v, hs.normal.z.v);
        hit.color = hs.color;
      }
      // direct light of a diffuse hit, or secondary_light's bounce loop (kernels.cl:342-365) ending in
      // the same shading — a single call site for both
      float medium = RT_AIR, gain = 1.0f;
      int bounce = 0;
      while (hit.id != -1) {
        if (hit.color.w > 0.0f) {
          if (SINGLE && !have_jit) {
            uint32_t rx, ry, rz;
            seed_rng(global_id, rx, ry, rz);
            Jitters<CH> jr;
            make_jitters<CH>(rx, ry, rz, jr);
            jit.store(jr);
            have_jit = true;
          }
          const float fl = gain * (RT_INDIRECT + direct_light_fast<CH, SINGLE, JittersShared<CH, kThreads>>(sc, hit.point, hit.normal, light, S, global_id, jit));
          total = V3<float>(total.x + hit.color.x * fl, total.y + hit.color.y * fl, total.z + hit.color.z * fl);
          break;
        }
        if (bounce >= p.B) break;
        bounce++;
        V3<float> start, ndir;
        if (hit.color.w == 0.0f) reflect_ray<float>(dir, hit.normal, hit.point, start, ndir, medium);
        else refract_ray<float>(dir, hit.normal, hit.point, medium, start, ndir, medium);
        dir = ndir;
        hit.id = -1;
        hit.color.w = 1.0f;
        closest_hit<float>(sc.g, start, dir, hit);
        gain = 0.9f;
      }
    }
  }
  const float ia = 1.0f / (float)(A * A);
  p.out[(size_t)y * p.W + x] = pack_argb<float>(V3<float>(total.x * ia, total.y * ia, total.z * ia));
}

#define RT_CAT2(a, b) a##b
#define RT_CAT(a, b) RT_CAT2(a, b)

cudaError_t RT_CAT(launch_fast_ch, RT_FAST_CH)(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream) {
  constexpr int CH = RT_FAST_CH;
  ctx->launch_extra_smem = sizeof(float) * 3 * CH * kThreads;  // jitter columns
  if (fp.S == CH) return launch_kernel(draw_fast_kernel<CH, true>, ctx, fp, stream);
  return launch_kernel(draw_fast_kernel<CH, false>, ctx, fp, stream);
}

}  // namespace rt
