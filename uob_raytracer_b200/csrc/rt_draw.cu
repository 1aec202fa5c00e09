// rt_draw.cu — the `draw` kernel (Source/kernels.cl:368-428) for sm_100a,
// brute-force variant: whole scene staged in shared memory, as the reference
// does with async_work_group_copy into __local (kernels.cl:374-376).  This file holds the generic
// kernel (the reference's plain loops: RT_FLAG_REFERENCE_LOOPS and RT_FLAG_COUNT_RAYS) and the dispatch;
// the tuned kernels — fast and bit-exact — are in rt_draw_fast.cu.
#include "rt_launch.cuh"

namespace rt {

// Block = 256 threads = 8 warps; a warp covers an 8x4 pixel tile (coherent rays,
// and each row of the tile is one full 32-byte sector of the ARGB frame), a
// block covers 16x16 pixels.
// Generic kernel: one thread per pixel, reference loop structure (rt_brute.cuh).  Instantiated
// with sfloat it is the un-culled anchor of the RT_FLAG_STRICT_IEEE path (RT_FLAG_REFERENCE_LOOPS).
template <class T, int CH>
__global__ void __launch_bounds__(kThreads) draw_brute_kernel(const __grid_constant__ FrameParams p, const float4 *__restrict__ scene,
                                                              int n, int n_sh) {
  extern __shared__ float4 smem[];
  BruteTracer<T> tr;
  tr.sc = stage_scene(smem, scene, n, n_sh);
  __syncthreads();
  int x, y, tx, ty;
  if (!pixel_of_thread(p, x, y, tx, ty)) return;
  p.out[(size_t)y * p.W + x] = shade_pixel<T, CH, BruteTracer<T>>(tr, p, x, y);
}

// strict kernel: 5n + 3n_sh float4; fast kernel: 5n (generic) + 3n (primary constants) + 3n (their affine form) + 5n_sh (shadow records and
// bounds) + n (plane records of the bounce rays) float4 + 8n ints (one binned triangle list per warp) + n_sh ints (identity caster list)
size_t brute_smem_bytes(int n, int n_sh) { return sizeof(float4) * (size_t)scene_smem_float4(n, n_sh); }
// what launch_fast_ch<CH> adds for the chunk size RT_DISPATCH_CH picks for S shadow samples
size_t fast_extra_smem(int S) {
  const int ch = S % 10 == 0 ? 10 : S % 8 == 0 ? 8 : S % 5 == 0 ? 5 : S % 4 == 0 ? 4 : S % 2 == 0 ? 2 : 1;
  return sizeof(float) * (size_t)(3 * ch + 4 * 7) * kFastThreads;
}

// Shadow samples are processed CH at a time (fully unrolled): the largest chunk that divides S, so
// that no padding samples are traced.  The fast kernels are compiled one translation unit per CH
// (rt_draw_fast.cu, -DRT_FAST_CH=n).
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
  // strict: the culled kernel computes the same frame; RT_FLAG_REFERENCE_LOOPS keeps the plain loop structure
  if ((ctx->cfg.flags & RT_FLAG_STRICT_IEEE) && ((ctx->cfg.flags & RT_FLAG_REFERENCE_LOOPS) || fp.ray_counters)) {
#define RT_STRICT(CH) launch_kernel(draw_brute_kernel<sfloat, CH>, ctx, fp, stream, 0, "draw_brute_kernel<sfloat," #CH ">")
    RT_DISPATCH_CH(RT_STRICT);
  }
  if (fp.ray_counters) {  // RT_FLAG_COUNT_RAYS: the generic kernel carries the counters
#define RT_COUNTING(CH) launch_kernel(draw_brute_kernel<float, CH>, ctx, fp, stream, 0, "draw_brute_kernel<float," #CH ">")
    RT_DISPATCH_CH(RT_COUNTING);
  }
#define RT_FAST(CH) launch_fast_ch##CH(ctx, fp, stream)
  RT_DISPATCH_CH(RT_FAST);
}

}  // namespace rt
