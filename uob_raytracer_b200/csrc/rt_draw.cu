// rt_draw.cu — the `draw` kernel (Source/kernels.cl:368-428) for sm_100a,
// brute-force variant: whole scene staged in shared memory, as the reference
// does with async_work_group_copy into __local (kernels.cl:374-376).
#include "rt_fast.cuh"
#include "rt_internal.h"

namespace rt {

// Block = 256 threads = 8 warps; a warp covers an 8x4 pixel tile (coherent rays,
// and each row of the tile is one full 32-byte sector of the ARGB frame), a
// block covers 16x16 pixels.
#ifndef RT_THREADS
#define RT_THREADS 256
#endif
#ifndef RT_MINBLOCKS
#define RT_MINBLOCKS 2
#endif
constexpr int kThreads = RT_THREADS, kTileW = 16, kTileH = kThreads / 16;

__device__ __forceinline__ SceneView stage_scene(float4 *smem, const float4 *__restrict__ scene, int n, int n_sh) {
  const int total = 5 * n + 4 * n_sh;
  for (int i = threadIdx.x; i < total; i += kThreads) smem[i] = scene[i];
  SceneView sc;
  sc.ta = smem;
  sc.tb = smem + n;
  sc.tc = smem + 2 * n;
  sc.tn = smem + 3 * n;
  sc.tcol = smem + 4 * n;
  sc.sa = smem + 5 * n;
  sc.sb = sc.sa + n_sh;
  sc.sc = sc.sb + n_sh;
  sc.sd = sc.sc + n_sh;
  sc.n = n;
  sc.n_sh = n_sh;
  return sc;
}

__device__ __forceinline__ bool pixel_of_thread(const FrameParams &p, int &x, int &y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  x = blockIdx.x * kTileW + (warp & 1) * 8 + (lane & 7);
  y = p.row0 + blockIdx.y * kTileH + (warp >> 1) * 4 + (lane >> 3);
  return x < p.W && y < p.row0 + p.rows;
}

// Generic kernel: one thread per pixel, reference loop structure (rt_brute.cuh).  Instantiated
// with sfloat it is the RT_FLAG_STRICT_IEEE path.
template <class T, int CH>
__global__ void __launch_bounds__(kThreads) draw_brute_kernel(const __grid_constant__ FrameParams p, const float4 *__restrict__ scene,
                                                              int n, int n_sh) {
  extern __shared__ float4 smem[];
  const SceneView sc = stage_scene(smem, scene, n, n_sh);
  __syncthreads();
  int x, y;
  if (!pixel_of_thread(p, x, y)) return;
  p.out[(size_t)y * p.W + x] = shade_pixel<T, CH>(sc, p, x, y);
}

// Fast kernel (rt_fast.cuh): CH shadow samples and RB primary rays of a pixel per triangle load.
template <int CH, int RB>
__global__ void __launch_bounds__(kThreads, RT_MINBLOCKS) draw_fast_kernel(const __grid_constant__ FrameParams p, const float4 *__restrict__ scene,
                                                             int n, int n_sh) {
  extern __shared__ float4 smem[];
  FastScene sc;
  sc.g = stage_scene(smem, scene, n, n_sh);
  float4 *pa = smem + 5 * n + 4 * n_sh, *pb = pa + n, *pc = pb + n;
  sc.pa = pa;
  sc.pb = pb;
  sc.pc = pc;
  sc.sd = sc.g.sd;
  const V3<float> cam(p.cam[0], p.cam[1], p.cam[2]), light(p.light[0], p.light[1], p.light[2]);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kThreads) primary_constants(sc.g, pa, pb, pc, cam, i);
  __syncthreads();
  int x, y;
  if (!pixel_of_thread(p, x, y)) return;

  const int A = p.A, S = p.S;
  const float SW = (float)p.W, SH = (float)p.H, fA = (float)A;
  const int global_id = __float2int_rz(__fadd_rn(__fmul_rn((float)y, SW), (float)x));  // kernels.cl:380, float arithmetic
  Jitters<CH> jit;
  if (S == CH) {
    uint32_t rx, ry, rz;
    seed_rng(global_id, rx, ry, rz);
    make_jitters<CH>(global_id, rx, ry, rz, jit);
  }
  // Primary ray directions with the reference's operation sequence (kernels.cl:384-405): a handful
  // of operations per ray, and it makes the primary hits bit-identical to the reference.
  typedef sfloat SF;
  const V3<SF> base(SF((float)(x * A)) - div_(SF(SW) * SF(fA), SF(2.0f)), SF((float)(y * A)) - div_(SF(SH) * SF(fA), SF(2.0f)), SF(p.focal));
  const V3<SF> r0(SF(p.rot[0]), SF(p.rot[1]), SF(p.rot[2])), r1(SF(p.rot[3]), SF(p.rot[4]), SF(p.rot[5])),
      r2(SF(p.rot[6]), SF(p.rot[7]), SF(p.rot[8]));
  V3<float> total(0.0f, 0.0f, 0.0f);
  const int rays = A * A;
  for (int r0i = 0; r0i < rays; r0i += RB) {
    V3<float> dir[RB];
#pragma unroll
    for (int k = 0; k < RB; k++) {
      const int idx = r0i + k;  // ray index dy*A + dx (kernels.cl:393-397)
      const int dy = idx / A, dx = idx - dy * A;
      const V3<SF> d0 = base + V3<SF>(SF((float)dx), SF((float)dy), SF(0.0f));
      const V3<SF> dn = normalize(V3<SF>(dot(r0, d0), dot(r1, d0), dot(r2, d0)));
      dir[k] = V3<float>(dn.x.v, dn.y.v, dn.z.v);
    }
    int best[RB];
    float bt[RB], bu[RB], bv[RB];
    primary_triangles<RB>(sc, dir, best, bt, bu, bv);
#pragma unroll
    for (int k = 0; k < RB; k++) {
      HitRec<float> hit;
      hit.id = -1;
      hit.color = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
      if (best[k] >= 0) {
        const V3<SF> v0 = xyz<SF>(sc.g.ta[best[k]]), e1 = xyz<SF>(sc.g.tb[best[k]]), e2 = xyz<SF>(sc.g.tc[best[k]]);
        const V3<SF> pt = (v0 + scale(SF(bu[k]), e1)) + scale(SF(bv[k]), e2);  // kernels.cl:124
        hit.id = best[k];
        hit.point = V3<float>(pt.x.v, pt.y.v, pt.z.v);
        hit.normal = xyz<float>(sc.g.tn[best[k]]);
        hit.color = sc.g.tcol[best[k]];
      }
      {  // spheres, strict as well (two per ray)
        HitRec<SF> hs;
        hs.id = hit.id;
        hs.point = V3<SF>(SF(hit.point.x), SF(hit.point.y), SF(hit.point.z));
        hs.normal = V3<SF>(SF(hit.normal.x), SF(hit.normal.y), SF(hit.normal.z));
        hs.color = hit.color;
        closest_spheres<SF>(V3<SF>(SF(cam.x), SF(cam.y), SF(cam.z)), V3<SF>(SF(dir[k].x), SF(dir[k].y), SF(dir[k].z)), SF(bt[k]), hs);
        hit.id = hs.id;
        hit.point = V3<float>(hs.point.x.v, hs.point.y.v, hs.point.z.v);
        hit.normal = V3<float>(hs.normal.x.v, hs.normal.y.v, hs.normal.z.v);
        hit.color = hs.color;
      }
      if (hit.id == -1) continue;
      if (hit.color.w <= 0.0f) {
        total = total + secondary_light_fast<CH>(sc, dir[k], hit, light, S, p.B, global_id, jit);
      } else {
        const float fl = RT_INDIRECT + direct_light_fast<CH>(sc, hit.point, hit.normal, light, S, global_id, jit);
        total = V3<float>(total.x + hit.color.x * fl, total.y + hit.color.y * fl, total.z + hit.color.z * fl);
      }
    }
  }
  const float ia = 1.0f / (float)rays;
  p.out[(size_t)y * p.W + x] = pack_argb<float>(V3<float>(total.x * ia, total.y * ia, total.z * ia));
}

size_t brute_smem_bytes(int n, int n_sh) { return sizeof(float4) * (size_t)(8 * n + 4 * n_sh); }
size_t brute_smem_limit() { return 200 * 1024; }

template <class K>
static cudaError_t launch_kernel(K kern, rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream) {
  const size_t smem = brute_smem_bytes(ctx->n, ctx->n_sh);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((fp.W + kTileW - 1) / kTileW, (fp.rows + kTileH - 1) / kTileH);
  kern<<<grid, kThreads, smem, stream>>>(fp, ctx->d_scene, ctx->n, ctx->n_sh);
  ctx->launches++;
  return cudaGetLastError();
}

// Shadow samples are processed CH at a time (fully unrolled): the largest chunk that divides S, so
// that no padding samples are traced.  Primary rays: 4 per triangle load when aa*aa % 4 == 0.
template <int CH>
static cudaError_t launch_fast(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream) {
  if ((fp.A * fp.A) % 4 == 0) return launch_kernel(draw_fast_kernel<CH, 4>, ctx, fp, stream);
  return launch_kernel(draw_fast_kernel<CH, 1>, ctx, fp, stream);
}

#define RT_DISPATCH_CH(CALL)            \
  do {                                  \
    const int S_ = fp.S;                \
    if (S_ % 10 == 0) return CALL(10);  \
    if (S_ % 8 == 0) return CALL(8);    \
    if (S_ % 5 == 0) return CALL(5);    \
    if (S_ % 4 == 0) return CALL(4);    \
    if (S_ % 2 == 0) return CALL(2);    \
    return CALL(1);                     \
  } while (0)

cudaError_t launch_draw_brute(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream) {
  if (ctx->cfg.flags & RT_FLAG_STRICT_IEEE) {
#define RT_STRICT(CH) launch_kernel(draw_brute_kernel<sfloat, CH>, ctx, fp, stream)
    RT_DISPATCH_CH(RT_STRICT);
  }
#define RT_FAST(CH) launch_fast<CH>(ctx, fp, stream)
  RT_DISPATCH_CH(RT_FAST);
}

}  // namespace rt
