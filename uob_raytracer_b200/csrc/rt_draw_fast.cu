// rt_draw_fast.cu — the fast `draw` kernel (see rt_fast.cuh for the method).  Compiled once per
// shadow chunk size: -DRT_FAST_CH=n gives launch_fast_ch<n>.
#include <stdio.h>

#include "rt_launch.cuh"

// A/B switches of the round-2 prologue work
#ifndef RT_HOST_CONSTS
#define RT_HOST_CONSTS 1
#endif
#ifndef RT_SPH_RECT
#define RT_SPH_RECT 1
#endif
#ifndef RT_UV_AFFINE
#define RT_UV_AFFINE 1
#endif

#ifndef RT_FAST_CH
#define RT_FAST_CH 8
#endif
// A/B switch, off.  Block prologue of a small scene (n <= 32, the Cornell box) with its latencies side by side instead of in a
// row: 1 = the binning warp reads its triangle from global memory (L1 hits behind the staging loads) while the other warps
// finish staging, and ONE barrier publishes both (instead of stage -> barrier -> one warp bins, seven wait -> barrier);
// 2 = in addition the staging loads are issued before the launch-order table entry is waited for, i.e. before the
// visible-rectangle exit.  ncu attributes 17 % of the stall samples to that chain, but both forms measured SLOWER on the
// B200: cfg2 0.2061 -> 0.2057 (1) / 0.2082 (2) ms, cfg3 2.342 -> 2.369 / 2.375 ms — the other resident blocks already hide
// the chain, and the empty tiles (44 % of a 16:9 frame) pay for staging they do not need.
#ifndef RT_PROLOGUE_OVERLAP
#define RT_PROLOGUE_OVERLAP 0
#endif
#ifndef RT_STAGE_UNROLL  // the staging loops run once or twice per thread: unrolled four times they were 365 instructions
#define RT_STAGE_UNROLL 1
#endif
constexpr int kStageUnroll = RT_STAGE_UNROLL;

namespace rt {

#ifdef RT_NO_LAZY  // A/B switch: every primary ray through the reference's exact sequence
constexpr bool kLazyPrimary = false;
#else
constexpr bool kLazyPrimary = true;
#endif

#ifndef RT_COOP_RAYS  // a warp with at most this many live bounce rays searches them cooperatively (rt_fast.cuh)
#define RT_COOP_RAYS 8
#endif

// secondary_light's bounce loop (kernels.cl:342-365) in warp-synchronous form, for the launches that are too small to
// fill the GPU (a share of a frame on several GPUs: they end with the longest bounce chain of a single ray).  Called by
// ALL lanes of warp_mask; follows every lane's mirror / glass hit to the diffuse surface it ends on (id = -1: nothing, or
// out of bounces — black).  Every round the lanes holding a mirror / glass hit form their next ray and a ballot counts
// them: few live rays are searched cooperatively by the whole warp — ray compaction across bounces, rt_fast.cuh:
// closest_triangles_coop — many by their own lanes side by side.  Same arithmetic and same winner either way.
template <class T>
__device__ __forceinline__ void resolve_bounces_coop(const FrameParams &p, const SceneView &g, unsigned warp_mask, HitRec<T> &hit, V3<T> dir) {
  float medium = RT_AIR;
  int bounce = 0;
  for (;;) {
    const bool specular = hit.id != -1 && hit.color.w <= 0.0f;
    const bool want = specular && bounce < p.B;
    if (specular && !want) hit.id = -1;  // out of bounces: black (kernels.cl:364)
    V3<T> start(T(0.0f), T(0.0f), T(0.0f));
    if (want) {
      bounce++;
      V3<T> ndir;
      if constexpr (is_strict<T>::value) {
        if (hit.color.w == 0.0f) reflect_ray<T>(dir, hit.normal, hit.point, start, ndir, medium);
        else refract_ray<T>(dir, hit.normal, hit.point, medium, start, ndir, medium);
      } else {
        bounce_ray_fixed(hit.color.w == 0.0f, dir, hit.normal, hit.point, medium, start, ndir);
      }
      dir = ndir;
      hit.id = -1;
      hit.color.w = 1.0f;
    }
    const unsigned live = __ballot_sync(warp_mask, want);
    if (live == 0u) break;
    ClosestState<T> cs;
    cs.reset();
    if (__popc(live) <= RT_COOP_RAYS) closest_triangles_coop<T>(g, warp_mask, live, start, dir, cs);
    else if (want) closest_triangles(g, start, dir, cs);
    if (want) finish_closest<T>(g, start, dir, cs, hit);
  }
}

// CH shadow samples per chunk, SINGLE = (S == CH).  STRICT = RT_FLAG_STRICT_IEEE: same binning, culls and
// caster lists, but every test that survives them — and all shading arithmetic — runs the reference's exact
// operation sequence, so the frame is bit-identical to the reference's (and to draw_brute_kernel<sfloat>).
//
// SPLIT: four lanes per pixel, each tracing every fourth ray of the pixel's A*A (A = 2 or 4), 8x8-pixel tiles.  Same
// frame; used when a launch is too small to fill the GPU (a 1/4 or 1/8 share of a 1080p frame): the launch then ends
// with its slowest pixel, a glass-sphere pixel whose A*A rays x several bounces form one serial chain — split four
// ways.  Per-ray contributions are parked and summed in the reference's ray order, so STRICT stays bit-identical.
template <int CH, bool SINGLE, bool STRICT, bool SPLIT, bool COOP>
__device__ __forceinline__ void draw_fast_body(const FrameParams &p, const float4 *__restrict__ scene, const float4 *__restrict__ fconst, int n, int n_sh,
                                               int bx, int by) {
  extern __shared__ float4 smem[];
  __shared__ int s_warp_count[kThreads / 32];
  __shared__ int s_base;
#if !RT_SPH_RECT
  __shared__ int s_spheres_visible;
#endif
  __shared__ int s_wlist[kThreads / 32][kWarpListMax];  // per-warp shadow caster lists

  if (p.gate_flag) {
    // Frame gate (rt_gate_next_frame): the frame buffer may belong to another GPU that is still reading the previous
    // frame.  One thread per block checks the device-local copy of the flag, and only polls the owner's memory over
    // NVLink while that copy is behind; the barrier orders every store of the block after the acquire.
    if (threadIdx.x == 0) {
      uint32_t seen;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.gate_seen) : "memory");
      if ((int32_t)(seen - p.gate_value) < 0) {
        const long long t0 = clock64();
        for (;;) {
          uint32_t v;
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.gate_flag) : "memory");
          if ((int32_t)(v - p.gate_value) >= 0) break;
          if (clock64() - t0 > 4000000000ll) {  // ~2 s: give up instead of hanging the GPU (reported by rt_synchronize)
            atomicExch(p.gate_status, 1);
            break;
          }
          __nanosleep(100);
        }
        atomicMax(p.gate_seen, p.gate_value);
      }
    }
    __syncthreads();
  }
  constexpr int kTW = SPLIT ? kSplitTileW : kTileW, kTH = SPLIT ? kSplitTileH : kTileH;
#ifdef RT_HOST_CONSTS_FAST
  constexpr bool kHC = RT_HOST_CONSTS;
#else
  constexpr bool kHC = STRICT && RT_HOST_CONSTS;
#endif
  const bool small_scene = n <= 32;  // one warp bins a small scene on its own: no block-wide hand-shakes
  const bool overlap = RT_PROLOGUE_OVERLAP && small_scene;
  int x, y, tile_x, tile_y;
  const bool in_frame = pixel_of_thread<SPLIT>(p, bx, by, x, y, tile_x, tile_y);
  // the tile lies outside the projection of the scene's bounding box (at 16:9 the bands beside the Cornell box, 44 % of
  // the frame): A*A black samples per pixel (kernels.cl:404-425), without binning anything
  const bool outside_visible_rect = tile_x >= p.vis_x1 || tile_x + kTW <= p.vis_x0 || tile_y >= p.vis_y1 || tile_y + kTH <= p.vis_y0;
  if (!(overlap && RT_PROLOGUE_OVERLAP >= 2) && outside_visible_rect) {  // before anything is staged
    if (in_frame) frame_of_row(p, y)[(size_t)y * p.W + x] = 0xff000000u;
    return;
  }

  // ---- stage the scene: generic arrays [0,5n) and the shadow records (global offset 5n+3n_sh) ----
  float4 *const gen = smem;
  float4 *const prim = smem + 5 * n;
  float4 *const aff = prim + 3 * n;
  float4 *const shad = aff + 3 * n;
  float4 *const sbound = shad + 4 * n_sh;  // bounding sphere of every shadow caster
  float4 *const tnd = sbound + n_sh;       // plane record of every triangle (bounce rays of the fast policy)
  int *const plist = reinterpret_cast<int *>(tnd + n);
  int *const full_list = plist + n;  // 0, 1, ..., n_sh-1: "test every caster"
  // per-thread columns after the lists: parked primary hits (4 x 7 words), then the jitters (SINGLE only)
  float *const rec_base = reinterpret_cast<float *>(smem + scene_smem_float4(n, n_sh));
  float *const jit_base = rec_base + 4 * 7 * kThreads;
#pragma unroll kStageUnroll
  for (int i = threadIdx.x; i < 5 * n; i += kThreads) gen[i] = scene[i];
  // Per-frame triangle constants: the bit-exact kernels take them from the host (rt_api.cu: prepare_frame — the same
  // single-rounded operations, done once per frame instead of once per block: 0.32 -> 0.30 ms on cfg2); the fast kernels
  // compute them in the block prologue as before — measured: the extra pointer and loads cost them 64 bytes more stack
  // (they are register-bound at 80 registers, three blocks per SM) and 5 % (cfg2) to 17 % (cfg3) of their speed.
  if constexpr (kHC) {
#pragma unroll kStageUnroll
    for (int i = threadIdx.x; i < 6 * n; i += kThreads) prim[i] = fconst[i];
  }
#pragma unroll kStageUnroll
  for (int i = threadIdx.x; i < 5 * n_sh + n; i += kThreads) shad[i] = scene[5 * n + 3 * n_sh + i];  // records, bounding spheres, plane records
#pragma unroll kStageUnroll
  for (int i = threadIdx.x; i < n_sh; i += kThreads) full_list[i] = i;
  FastScene sc;
  sc.g.ta = gen;
  sc.g.tb = gen + n;
  sc.g.tc = gen + 2 * n;
  sc.g.tn = gen + 3 * n;
  sc.g.tcol = gen + 4 * n;
  sc.g.sa = sc.g.sb = sc.g.sc = nullptr;  // the SoA shadow arrays belong to the generic kernel
  sc.g.tnd = tnd;
  sc.g.n = n;
  sc.g.n_sh = n_sh;
  sc.prim = prim;
  sc.aff = aff;
  sc.shad = shad;
  sc.plist = plist;
  const V3<float> cam(p.cam[0], p.cam[1], p.cam[2]), light(p.light[0], p.light[1], p.light[2]);
  const int A = p.A, S = p.S;
  const float SW = (float)p.W, SH = (float)p.H, fA = (float)A;
  if (overlap) {
    if (RT_PROLOGUE_OVERLAP >= 2 && outside_visible_rect) {  // block-uniform
      if (in_frame) frame_of_row(p, y)[(size_t)y * p.W + x] = 0xff000000u;
      return;
    }
  } else {
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
  }

  // ---- per-triangle camera constants + binning of the triangles against this block's tile ----
  if (!small_scene || threadIdx.x < 32) {
    // corner rays of the tile in virtual (sub-pixel) coordinates, un-normalised (kernels.cl:384-400)
    const float vx0 = (float)(tile_x * A) - SW * fA * 0.5f;
    const float vy0 = (float)(tile_y * A) - SH * fA * 0.5f;
    const float vx1 = vx0 + (float)(kTW * A - 1), vy1 = vy0 + (float)(kTH * A - 1);
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
    if constexpr (kHC) dmax = p.dmax;  // the host's constants carry the frame-wide tolerance
#if !RT_SPH_RECT
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
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (small_scene) {
      bool keep = false;
      if (lane < n) {
        if constexpr (!kHC) {
          if (overlap) primary_constants(scene[lane], scene[n + lane], scene[2 * n + lane], prim, cam, lane);  // (L1 hits: the staging loads just fetched these lines; this lane's own prim[] writes need no barrier to be read back)
          else primary_constants(sc.g, prim, cam, lane);
          if constexpr (!STRICT) primary_affine(prim, aff, lane, p.rot, p.focal, dmax);
          keep = tile_may_hit(prim, lane, dc, dmax);
        } else {
          keep = tile_may_hit(overlap ? fconst : prim, lane, dc, dmax);  // overlap: the staged copy is not published yet
        }
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, keep);
      if (keep) plist[__popc(ballot & ((1u << lane) - 1u))] = lane;
      if (lane == 0) s_base = __popc(ballot);
    } else
    for (int base = 0; base < n; base += kThreads) {
      const int i = base + threadIdx.x;
      bool keep = false;
      if (i < n) {
        if constexpr (!kHC) {
          primary_constants(sc.g, prim, cam, i);
          if constexpr (!STRICT) primary_affine(prim, aff, i, p.rot, p.focal, dmax);
        }
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
  }
  __syncthreads();
  sc.n_prim = s_base;
#if RT_SPH_RECT
  // Can a primary ray of the tile reach a sphere?  The host projected the spheres (rt_api.cu: sphere_pixel_rect): a tile
  // outside both pixel rectangles (which carry two pixels of margin) cannot.
  bool spheres_visible = false;
#pragma unroll
  for (int i = 0; i < RT_SPHERES; i++)
    spheres_visible |= tile_x < p.sph_px[i][2] && tile_x + kTW > p.sph_px[i][0] && tile_y < p.sph_px[i][3] && tile_y + kTH > p.sph_px[i][1];
#else
  const bool spheres_visible = s_spheres_visible != 0;
#endif
#ifdef RT_DEBUG_PROLOGUE_ONLY  // experiment: cost of staging + binning alone
  if (in_frame) frame_of_row(p, y)[(size_t)y * p.W + x] = 0xff000000u | (unsigned)sc.n_prim;
  return;
#endif
  if (sc.n_prim == 0 && !spheres_visible) {
    // nothing can be hit from this tile (at 1080p 44 % of the frame lies beside the box): every ray misses, the
    // pixel is the average of A*A black samples (kernels.cl:404-425)
    if (in_frame) frame_of_row(p, y)[(size_t)y * p.W + x] = 0xff000000u;
    return;
  }

  const unsigned warp_mask = __ballot_sync(0xffffffffu, in_frame);  // lanes that stay for the warp collectives below
  if (!in_frame) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int global_id = __float2int_rz(__fadd_rn(__fmul_rn((float)y, SW), (float)x));  // kernels.cl:380, float arithmetic
  // SINGLE: the pixel's S jitters, generated on the first shading point (44 % of the 1080p frame
  // lies outside the box and never shades) and kept in shared memory
  JittersShared<CH, kThreads> jit;
  jit.p = JittersShared<CH, kThreads>::column(jit_base, threadIdx.x);
  bool have_jit = false;
  // Primary ray directions with the reference's operation sequence (kernels.cl:384-405): a handful
  // of operations per ray, and it makes the primary hits bit-identical to the reference.
  typedef sfloat SF;
  const V3<SF> base(SF((float)(x * A)) - div_(SF(SW) * SF(fA), SF(2.0f)), SF((float)(y * A)) - div_(SF(SH) * SF(fA), SF(2.0f)), SF(p.focal));
  const V3<SF> r0(SF(p.rot[0]), SF(p.rot[1]), SF(p.rot[2])), r1(SF(p.rot[3]), SF(p.rot[4]), SF(p.rot[5])),
      r2(SF(p.rot[6]), SF(p.rot[7]), SF(p.rot[8]));
  const V3<SF> cam_s(SF(cam.x), SF(cam.y), SF(cam.z));
  V3<float> total(0.0f, 0.0f, 0.0f);
  V3<SF> total_s(SF(0.0f), SF(0.0f), SF(0.0f));  // STRICT accumulates in the reference's order and arithmetic
  const V3<SF> light_s(SF(light.x), SF(light.y), SF(light.z));
  const int split_q = SPLIT ? (int)(threadIdx.x & 3u) : 0;  // SPLIT: this lane traces rays split_q, split_q + 4, ...
  const int rays = SPLIT ? (A * A) >> 2 : A * A;
  // Rays of a pixel are processed in groups of kGroup, in the reference's order dy*A + dx
  // (kernels.cl:393-397).  Phase 1 finds the primary hit of each ray of the group and parks it in this
  // thread's shared-memory column; phase 2 builds ONE shadow-caster list for the warp from the bounding
  // box of all its diffuse hits of the group; phase 3 shades the parked hits.
  constexpr int kGroup = 4, kRec = 7;  // record: id, point.xyz, normal.xyz
  float *const rec = rec_base + threadIdx.x;
  int ray_dx = 0, ray_dy = 0;
#pragma unroll 1
  for (int g0 = 0; g0 < rays; g0 += kGroup) {
    const int group = min(kGroup, rays - g0);
    const float big = 3.0e38f;
    V3<float> blo(big, big, big), bhi(-big, -big, -big);
    int first_dx = ray_dx, first_dy = ray_dy;
#pragma unroll 1
    for (int k = 0; k < group; k++) {
      int cur_dx = ray_dx, cur_dy = ray_dy;
      if constexpr (SPLIT) {
        const int idx = split_q + 4 * (g0 + k);
        cur_dy = idx / A;
        cur_dx = idx - cur_dy * A;
      } else if (++ray_dx == A) {
        ray_dx = 0;
        ray_dy++;
      }
      HitRec<SF> hs;
      hs.id = -1;
      hs.color = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
      hs.point = V3<SF>(SF(0.0f), SF(0.0f), SF(0.0f));
      hs.normal = hs.point;
      bool exact = true;
      if constexpr (!STRICT && kLazyPrimary) {
        // Lazy-exact (rt_fast.cuh: primary_fast): two FMAs per determinant on the sub-pixel coordinates decide every ray
        // that is not within rounding of a triangle edge, of a tie between two triangles, or of a sphere; no direction,
        // no normalisation, no division by the reference's rules.  The few rays in doubt take the exact path below.
        const float vx = base.x.v + (float)cur_dx, vy = base.y.v + (float)cur_dy;  // exact: small integers (or halves)
        int bi;
        float bu, bv;
        exact = !primary_fast(sc, vx, vy, bi, bu, bv);
        if (!exact && spheres_visible) {
          const V3<float> w(vx, vy, p.focal);
          const V3<float> d0(p.rot[0] * w.x + p.rot[1] * w.y + p.rot[2] * w.z, p.rot[3] * w.x + p.rot[4] * w.y + p.rot[5] * w.z,
                             p.rot[6] * w.x + p.rot[7] * w.y + p.rot[8] * w.z);
          exact = !spheres_surely_missed(d0, cam);
        }
        if (!exact && bi >= 0) {
          const float4 A0 = sc.g.ta[bi], B0 = sc.g.tb[bi], C0 = sc.g.tc[bi], N0 = sc.g.tn[bi];
          hs.id = bi;
          hs.point = V3<SF>(SF(fmaf(bv, C0.x, fmaf(bu, B0.x, A0.x))), SF(fmaf(bv, C0.y, fmaf(bu, B0.y, A0.y))), SF(fmaf(bv, C0.z, fmaf(bu, B0.z, A0.z))));
          hs.normal = xyz<SF>(N0);
          hs.color = sc.g.tcol[bi];
        }
      }
      if (exact) {
        // the reference's own operation sequence (kernels.cl:384-405, :100-163): direction, triangles, spheres
        const V3<SF> d0 = base + V3<SF>(SF((float)cur_dx), SF((float)cur_dy), SF(0.0f));
        const V3<SF> dns = normalize(V3<SF>(dot(r0, d0), dot(r1, d0), dot(r2, d0)));
        int bi;
        float bt, bu, bv;
        primary_triangles(sc, V3<float>(dns.x.v, dns.y.v, dns.z.v), bi, bt, bu, bv);
        if (bi >= 0) {
          const V3<SF> v0 = xyz<SF>(sc.g.ta[bi]), e1 = xyz<SF>(sc.g.tb[bi]), e2 = xyz<SF>(sc.g.tc[bi]);
          hs.id = bi;
          hs.point = (v0 + scale(SF(bu), e1)) + scale(SF(bv), e2);  // kernels.cl:124
          hs.normal = xyz<SF>(sc.g.tn[bi]);
          hs.color = sc.g.tcol[bi];
#if RT_UV_AFFINE
          if constexpr (!STRICT && kLazyPrimary) {
            // Which rays take this exact path depends on the binned list, hence on the tile shape — so a triangle hit
            // takes its hit point from the same affine (u, v) as the rays that were never in doubt: every lane mapping and
            // every partition of the frame then produces the same bits.
            const float vx = base.x.v + (float)cur_dx, vy = base.y.v + (float)cur_dy;
            primary_uv(sc, bi, vx, vy, bu, bv);
            const float4 A0 = sc.g.ta[bi], B0 = sc.g.tb[bi], C0 = sc.g.tc[bi];
            hs.point = V3<SF>(SF(fmaf(bv, C0.x, fmaf(bu, B0.x, A0.x))), SF(fmaf(bv, C0.y, fmaf(bu, B0.y, A0.y))), SF(fmaf(bv, C0.z, fmaf(bu, B0.z, A0.z))));
          }
#endif
        }
        // the two spheres, strict as well (skipped when no ray of the block's tile can reach one)
        if (spheres_visible) closest_spheres<SF>(cam_s, dns, SF(bt), hs);
      }
      // sphere i is parked as id -2-i so that phase 3 can recover its colour (kernels.cl:28 uses -2 for both)
      int id = hs.id;
      if (id == -2) id = (hs.color.w == c_sphere_color[0].w) ? -2 : -3;
      float *q = rec + k * kRec * kThreads;
      q[0] = __int_as_float(id);
      q[1 * kThreads] = hs.point.x.v;
      q[2 * kThreads] = hs.point.y.v;
      q[3 * kThreads] = hs.point.z.v;
      q[4 * kThreads] = hs.normal.x.v;
      q[5 * kThreads] = hs.normal.y.v;
      q[6 * kThreads] = hs.normal.z.v;
      if (hs.id != -1 && hs.color.w > 0.0f) {
        blo = V3<float>(fminf(blo.x, hs.point.x.v), fminf(blo.y, hs.point.y.v), fminf(blo.z, hs.point.z.v));
        bhi = V3<float>(fmaxf(bhi.x, hs.point.x.v), fmaxf(bhi.y, hs.point.y.v), fmaxf(bhi.z, hs.point.z.v));
      }
    }
    // Phase 2 — warp-level caster cull: the diffuse primary hits of the warp's 8x4 pixels lie close
    // together; casters that cannot shadow any point of their bounding box (walls, far faces) are dropped
    // for the whole warp.  All lanes of warp_mask are converged here (uniform loops).
    const int *warp_list = full_list;
    int n_warp_list = n_sh;
    if (n_sh <= kWarpListMax) {
      if (__ballot_sync(warp_mask, blo.x <= bhi.x) != 0u) {
        V3<float> lo, hi;
        lo.x = float_of_ord(__reduce_min_sync(warp_mask, ord_of_float(blo.x)));
        lo.y = float_of_ord(__reduce_min_sync(warp_mask, ord_of_float(blo.y)));
        lo.z = float_of_ord(__reduce_min_sync(warp_mask, ord_of_float(blo.z)));
        hi.x = float_of_ord(__reduce_max_sync(warp_mask, ord_of_float(bhi.x)));
        hi.y = float_of_ord(__reduce_max_sync(warp_mask, ord_of_float(bhi.y)));
        hi.z = float_of_ord(__reduce_max_sync(warp_mask, ord_of_float(bhi.z)));
        // the casters are dealt out to the lanes that are present (edge warps have fewer than 32)
        const int n_act = __popc(warp_mask), rank = __popc(warp_mask & ((1u << lane) - 1u));
        int count = 0;
        for (int c0 = 0; c0 < n_sh; c0 += n_act) {
          const int c = c0 + rank;
          const bool keep = c < n_sh && box_may_be_shadowed_by(shad, sbound, c, lo, hi, light);
          const unsigned bal = __ballot_sync(warp_mask, keep);
          if (keep) s_wlist[warp][count + __popc(bal & ((1u << lane) - 1u))] = c;
          count += __popc(bal);
        }
        __syncwarp(warp_mask);
        warp_list = s_wlist[warp];
        n_warp_list = count;
      }
    }
    // Phase 3 — shade the parked hits
#pragma unroll 1
    for (int k = 0; k < group; k++) {
      float *q = rec + k * kRec * kThreads;
      HitRec<float> hit;
      hit.id = __float_as_int(q[0]);
      hit.point = V3<float>(q[1 * kThreads], q[2 * kThreads], q[3 * kThreads]);
      hit.normal = V3<float>(q[4 * kThreads], q[5 * kThreads], q[6 * kThreads]);
      hit.color = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
      if (hit.id >= 0) hit.color = sc.g.tcol[hit.id];
      else if (hit.id <= -2) hit.color = c_sphere_color[-2 - hit.id];
      V3<SF> dir_s(SF(0.0f), SF(0.0f), SF(0.0f));
      if (hit.id != -1 && hit.color.w <= 0.0f) {
        // mirror / glass: the bounce needs the ray direction again (same strict sequence as phase 1)
        const int idx = SPLIT ? split_q + 4 * (g0 + k) : first_dy * A + first_dx + k;
        const int ddy = idx / A, ddx = idx - ddy * A;
        const V3<SF> d0 = base + V3<SF>(SF((float)ddx), SF((float)ddy), SF(0.0f));
        dir_s = normalize(V3<SF>(dot(r0, d0), dot(r1, d0), dot(r2, d0)));
      }
      sc.clist = warp_list;
      sc.n_clist = n_warp_list;
      // direct light of a diffuse hit, or secondary_light's bounce loop (kernels.cl:342-365) ending in
      // the same shading — a single call site for both
      float medium = RT_AIR;
      int bounce = 0;
      if constexpr (SPLIT) {  // this ray's contribution, parked in its record slot (black unless shaded below)
        q[1 * kThreads] = 0.0f;
        q[2 * kThreads] = 0.0f;
        q[3 * kThreads] = 0.0f;
      }
      if constexpr (STRICT) {
        HitRec<SF> hs;
        hs.id = hit.id;
        hs.point = V3<SF>(SF(hit.point.x), SF(hit.point.y), SF(hit.point.z));
        hs.normal = V3<SF>(SF(hit.normal.x), SF(hit.normal.y), SF(hit.normal.z));
        hs.color = hit.color;
        bool bounced = false;
        if constexpr (COOP) {
          bounced = hs.id != -1 && hs.color.w <= 0.0f;
          if (__ballot_sync(warp_mask, bounced) != 0u) resolve_bounces_coop<SF>(p, sc.g, warp_mask, hs, dir_s);
          if (bounced) {
            sc.clist = full_list;  // the warp list was built for the primary hits only
            sc.n_clist = n_sh;
          }
        }
        while (hs.id != -1) {
          if (hs.color.w > 0.0f) {
            if (SINGLE && !have_jit) {
              make_jitters_shared<CH, kThreads, true>(global_id, jit);
              have_jit = true;
            }
            const SF dl = direct_light_strict<CH, SINGLE, JittersShared<CH, kThreads>>(sc, hs.point, hs.normal, light_s, S, global_id, jit);
            const V3<SF> lightv(SF(RT_INDIRECT) + dl, SF(RT_INDIRECT) + dl, SF(RT_INDIRECT) + dl);
            // primary: colour*(indirect + direct) (kernels.cl:422); after a bounce: 0.9*light*colour (:355)
            const V3<SF> contrib = bounced ? scale(SF(0.9f), lightv) * xyz<SF>(hs.color) : xyz<SF>(hs.color) * lightv;
            if constexpr (SPLIT) {
              q[1 * kThreads] = contrib.x.v;
              q[2 * kThreads] = contrib.y.v;
              q[3 * kThreads] = contrib.z.v;
            } else {
              total_s = total_s + contrib;
            }
            break;
          }
          if (bounce >= p.B) break;
          bounce++;
          V3<SF> start, ndir;
          if (hs.color.w == 0.0f) reflect_ray<SF>(dir_s, hs.normal, hs.point, start, ndir, medium);
          else refract_ray<SF>(dir_s, hs.normal, hs.point, medium, start, ndir, medium);
          dir_s = ndir;
          hs.id = -1;
          hs.color.w = 1.0f;
          closest_hit<SF>(sc.g, start, dir_s, hs);
          bounced = true;
          sc.clist = full_list;  // the warp list was built for the primary hits only
          sc.n_clist = n_sh;
        }
      } else {
      V3<float> dir(dir_s.x.v, dir_s.y.v, dir_s.z.v);
      float gain = 1.0f;
      if constexpr (COOP) {
        const bool specular = hit.id != -1 && hit.color.w <= 0.0f;
        if (__ballot_sync(warp_mask, specular) != 0u) resolve_bounces_coop<float>(p, sc.g, warp_mask, hit, dir);
        if (specular) {
          gain = 0.9f;
          sc.clist = full_list;  // the warp list was built for the primary hits only
          sc.n_clist = n_sh;
        }
      }
      while (hit.id != -1) {
        if (hit.color.w > 0.0f) {
          if (SINGLE && !have_jit) {
            make_jitters_shared<CH, kThreads, false>(global_id, jit);
            have_jit = true;
          }
          const float fl = __fmul_rn(gain, __fadd_rn(RT_INDIRECT, direct_light_fast<CH, SINGLE, JittersShared<CH, kThreads>>(sc, hit.point, hit.normal, light, S, global_id, jit)));
          // product and sum rounded separately: the split mapping parks the product and adds it later, and every
          // lane mapping (hence every multi-GPU partition) has to produce the same bits
          const V3<float> contrib(__fmul_rn(hit.color.x, fl), __fmul_rn(hit.color.y, fl), __fmul_rn(hit.color.z, fl));
          if constexpr (SPLIT) {
            q[1 * kThreads] = contrib.x;
            q[2 * kThreads] = contrib.y;
            q[3 * kThreads] = contrib.z;
          } else {
            total = V3<float>(__fadd_rn(total.x, contrib.x), __fadd_rn(total.y, contrib.y), __fadd_rn(total.z, contrib.z));
          }
          break;
        }
        if (bounce >= p.B) break;
        bounce++;
        V3<float> start, ndir;
        bounce_ray_fixed(hit.color.w == 0.0f, dir, hit.normal, hit.point, medium, start, ndir);
        dir = ndir;
        hit.id = -1;
        hit.color.w = 1.0f;
        closest_hit_bounce(sc.g, start, dir, hit);
        gain = 0.9f;
        sc.clist = full_list;  // the warp list was built for the primary hits only
        sc.n_clist = n_sh;
      }
      }
    }
  }
  if constexpr (SPLIT) {
    // the pixel's first lane sums the parked contributions in the reference's ray order (kernels.cl:415-425)
    __syncwarp(warp_mask);
    if (split_q != 0) return;
    for (int j = 0; j < rays; j++)
      for (int qq = 0; qq < 4; qq++) {
        const float *c = rec + qq + j * kRec * kThreads;
        const V3<float> v(c[1 * kThreads], c[2 * kThreads], c[3 * kThreads]);
        if constexpr (STRICT) total_s = total_s + V3<SF>(SF(v.x), SF(v.y), SF(v.z));
        else total = V3<float>(__fadd_rn(total.x, v.x), __fadd_rn(total.y, v.y), __fadd_rn(total.z, v.z));
      }
  }
  if constexpr (STRICT) {
    const SF fa = SF(__int2float_rn(A * A));
    frame_of_row(p, y)[(size_t)y * p.W + x] = pack_argb<SF>(V3<SF>(div_(total_s.x, fa), div_(total_s.y, fa), div_(total_s.z, fa)));
    return;
  }
  const float ia = 1.0f / (float)(A * A);
  frame_of_row(p, y)[(size_t)y * p.W + x] = pack_argb<float>(V3<float>(__fmul_rn(total.x, ia), __fmul_rn(total.y, ia), __fmul_rn(total.z, ia)));
}

// rt_signal_after_frame: every block makes its stores visible system-wide before it counts itself done; the block that
// sees the count complete has all of them behind it, tells the owner of the frame and re-arms the counter.
__device__ __forceinline__ void block_done(const FrameParams &p) {
  if (!p.signal_flag) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(p.work_counter + 1, 1u) == gridDim.x - 1) {
      p.work_counter[1] = 0u;
      __threadfence_system();
      asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(p.signal_flag) : "memory");
    }
  }
}

template <int CH, bool SINGLE, bool STRICT, bool SPLIT>
__global__ void __launch_bounds__(kThreads, STRICT ? RT_STRICT_MINBLOCKS : RT_MINBLOCKS)
    draw_fast_kernel(const __grid_constant__ FrameParams p, const float4 *__restrict__ scene, const float4 *__restrict__ fconst, int n, int n_sh) {
  int bx, by;
  tile_of_block<SPLIT, false>(p, (int)blockIdx.x, bx, by);
  draw_fast_body<CH, SINGLE, STRICT, SPLIT, SPLIT>(p, scene, fconst, n, n_sh, bx, by);
}

// Mixed launch for shares of a frame that cannot fill the GPU: tiles inside the screen rectangle of a sphere (mirror / glass
// bounce chains, the pixels a small launch ends up waiting for) are rendered as four 8x8 sub-tiles with four lanes per pixel,
// by the first n_split blocks of the grid; every other tile by an ordinary block.  The rectangles come with the launch
// arguments (rt_api.cu: sphere_rects) — nothing is built, cached or copied per camera; they only steer performance: either
// mapping renders any tile correctly.
template <int CH, bool SINGLE, bool STRICT>
__global__ void __launch_bounds__(kThreads, STRICT ? RT_STRICT_MINBLOCKS : RT_MINBLOCKS)
    draw_fast_mixed_kernel(const __grid_constant__ FrameParams p, const float4 *__restrict__ scene, const float4 *__restrict__ fconst, int n, int n_sh) {
  int bx, by;
  if ((int)blockIdx.x < p.n_split) {
    if (tile_of_block<true, true>(p, (int)blockIdx.x, bx, by)) draw_fast_body<CH, SINGLE, STRICT, true, true>(p, scene, fconst, n, n_sh, bx, by);
  } else {
    if (tile_of_block<false, true>(p, (int)blockIdx.x - p.n_split, bx, by)) draw_fast_body<CH, SINGLE, STRICT, false, false>(p, scene, fconst, n, n_sh, bx, by);  // no sphere in sight of these tiles
  }
#ifdef RT_TAIL_SIGNAL  // A/B switch, off: see launch_fast_ch
  block_done(p);
#endif
}

#define RT_CAT2(a, b) a##b
#define RT_CAT(a, b) RT_CAT2(a, b)

// One block per tile; `blocks` of them.
template <class K>
static cudaError_t launch_tiles(K kern, rt_ctx *ctx, FrameParams &fp, int blocks, cudaStream_t stream, size_t extra_smem, const char *name) {
  const size_t smem = brute_smem_bytes(ctx->n, ctx->n_sh) + extra_smem;
  if (smem > 32 * 1024) {  // dynamic + the kernels' static shared memory (kDrawStaticSmem) must stay under the 48 KB default
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  ctx->last_kernel = name;
  if (blocks <= 0) {
    if (fp.signal_flag) return launch_peer_add(fp.signal_flag, stream);  // nothing to draw: the delivery still has to be reported
    return cudaSuccess;
  }
  fp.work_counter = ctx->d_work + 2 * (ctx->work_seq++ % rt_ctx::kWorkSlots);
  kern<<<blocks, kThreads, smem, stream>>>(fp, ctx->d_scene, ctx->d_fconst, ctx->n, ctx->n_sh);
  ctx->launches++;
  return cudaGetLastError();
}

cudaError_t RT_CAT(launch_fast_ch, RT_FAST_CH)(rt_ctx *ctx, const FrameParams &fp_in, cudaStream_t stream) {
  constexpr int CH = RT_FAST_CH;
  const size_t extra = sizeof(float) * (3 * CH + 4 * 7) * kThreads;  // jitter columns + parked primary hits (== fast_extra_smem(S))
  const bool strict = (ctx->cfg.flags & RT_FLAG_STRICT_IEEE) != 0, single = fp_in.S == CH;
  FrameParams fp = fp_in;
  SplitMode mode = split_mode(ctx, fp);
  int n_sub = 0;
  if (mode == kSplitHeavy && (n_sub = sphere_rects(fp)) == 0) mode = kSplitNone;  // no sphere in sight: nothing to split
  if (mode != kSplitHeavy) fp.n_rect = 0;
  // this rank's share of the tiles (multi-GPU block interleave)
  const int tw = mode == kSplitAll ? kSplitTileW : kTileW, th = mode == kSplitAll ? kSplitTileH : kTileH;
  fp.grid_x = (fp.W + tw - 1) / tw;
  fp.n_blocks = fp.grid_x * ((fp.rows + th - 1) / th);
  fp.blk_stride = ctx->cfg.block_stride > 1 ? ctx->cfg.block_stride : 1;
  fp.blk_phase = ctx->cfg.block_stride > 1 ? ctx->cfg.block_phase : 0;
  fp.tile_order = tile_order_for(ctx, fp.row0, fp.rows, fp.grid_x, fp.n_blocks, tw, th);
  const int my_tiles = fp.n_blocks > fp.blk_phase ? (fp.n_blocks - fp.blk_phase + fp.blk_stride - 1) / fp.blk_stride : 0;
  const int my_split = n_sub > fp.blk_phase ? (n_sub - fp.blk_phase + fp.blk_stride - 1) / fp.blk_stride : 0;
  fp.n_split = my_split;
  fp.n_split_items = my_split;
  fp.n_items = my_tiles + my_split;
  char name[96];
  snprintf(name, sizeof name, "draw_fast_kernel<%d,%s,%s,%s>", CH, single ? "true" : "false", strict ? "true" : "false",
           mode == kSplitHeavy ? "mixed" : (mode == kSplitAll ? "split" : "plain"));
  // rt_signal_after_frame: a one-thread kernel behind the draw kernel reports the delivery.  (Folding it into the draw
  // kernel — last block by an atomic count, -DRT_TAIL_SIGNAL — was measured slower: the exit path through a block-wide
  // barrier costs the register-bound kernels 48 bytes more stack, 3-7 % of a small launch, more than the 2 us launch saves.)
  uint32_t *signal_behind = nullptr;
#ifdef RT_TAIL_SIGNAL
  if (mode != kSplitHeavy && fp.signal_flag) {
#else
  if (fp.signal_flag) {
#endif
    signal_behind = fp.signal_flag;
    fp.signal_flag = nullptr;
  }
  auto finish = [&](cudaError_t e) {
    if (e == cudaSuccess && signal_behind) {
      e = launch_peer_add(signal_behind, stream);
      ctx->launches++;
    }
    return e;
  };
#define RT_LAUNCH(K_) finish(launch_tiles(K_, ctx, fp, fp.n_items, stream, extra, name))
  if (mode == kSplitHeavy) {
    if (strict) return single ? RT_LAUNCH((draw_fast_mixed_kernel<CH, true, true>)) : RT_LAUNCH((draw_fast_mixed_kernel<CH, false, true>));
    return single ? RT_LAUNCH((draw_fast_mixed_kernel<CH, true, false>)) : RT_LAUNCH((draw_fast_mixed_kernel<CH, false, false>));
  }
  if (mode == kSplitAll) {
    if (strict) return single ? RT_LAUNCH((draw_fast_kernel<CH, true, true, true>)) : RT_LAUNCH((draw_fast_kernel<CH, false, true, true>));
    return single ? RT_LAUNCH((draw_fast_kernel<CH, true, false, true>)) : RT_LAUNCH((draw_fast_kernel<CH, false, false, true>));
  }
  if (strict) return single ? RT_LAUNCH((draw_fast_kernel<CH, true, true, false>)) : RT_LAUNCH((draw_fast_kernel<CH, false, true, false>));
  return single ? RT_LAUNCH((draw_fast_kernel<CH, true, false, false>)) : RT_LAUNCH((draw_fast_kernel<CH, false, false, false>));
#undef RT_LAUNCH
}

}  // namespace rt
