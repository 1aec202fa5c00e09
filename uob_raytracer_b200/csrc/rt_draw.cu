// rt_draw.cu — the `draw` kernel (Source/kernels.cl:368-428) for sm_100a,
// brute-force variant: whole scene staged in shared memory, as the reference
// does with async_work_group_copy into __local (kernels.cl:374-376).
#include "rt_brute.cuh"
#include "rt_internal.h"

namespace rt {

// Block = 256 threads = 8 warps; a warp covers an 8x4 pixel tile (coherent rays,
// and each row of the tile is one full 32-byte sector of the ARGB frame), a
// block covers 16x16 pixels.
constexpr int kTileW = 16, kTileH = 16, kThreads = 256;

template <class T, int CH>
__global__ void __launch_bounds__(kThreads) draw_brute_kernel(const __grid_constant__ FrameParams p, const float4 *__restrict__ scene,
                                                              int n, int n_sh) {
  extern __shared__ float4 smem[];
  const int total = 5 * n + 3 * n_sh;
  for (int i = threadIdx.x; i < total; i += kThreads) smem[i] = scene[i];
  __syncthreads();
  SceneView sc;
  sc.ta = smem;
  sc.tb = smem + n;
  sc.tc = smem + 2 * n;
  sc.tn = smem + 3 * n;
  sc.tcol = smem + 4 * n;
  sc.sa = smem + 5 * n;
  sc.sb = sc.sa + n_sh;
  sc.sc = sc.sb + n_sh;
  sc.n = n;
  sc.n_sh = n_sh;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x = blockIdx.x * kTileW + (warp & 1) * 8 + (lane & 7);
  const int y = p.row0 + blockIdx.y * kTileH + (warp >> 1) * 4 + (lane >> 3);
  if (x >= p.W || y >= p.row0 + p.rows) return;
  p.out[(size_t)y * p.W + x] = shade_pixel<T, CH>(sc, p, x, y);
}

size_t brute_smem_bytes(int n, int n_sh) { return sizeof(float4) * (size_t)(5 * n + 3 * n_sh); }
size_t brute_smem_limit() { return 200 * 1024; }

template <class T, int CH>
static cudaError_t launch_t(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream) {
  const size_t smem = brute_smem_bytes(ctx->n, ctx->n_sh);
  auto kern = draw_brute_kernel<T, CH>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((fp.W + kTileW - 1) / kTileW, (fp.rows + kTileH - 1) / kTileH);
  kern<<<grid, kThreads, smem, stream>>>(fp, ctx->d_scene, ctx->n, ctx->n_sh);
  ctx->launches++;
  return cudaGetLastError();
}

template <class T>
static cudaError_t launch_ch(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream) {
  // shadow samples are processed CH at a time (fully unrolled); pick the largest
  // chunk that divides S so no padding samples are traced
  const int S = fp.S;
  if (S % 10 == 0) return launch_t<T, 10>(ctx, fp, stream);
  if (S % 8 == 0) return launch_t<T, 8>(ctx, fp, stream);
  if (S % 5 == 0) return launch_t<T, 5>(ctx, fp, stream);
  if (S % 4 == 0) return launch_t<T, 4>(ctx, fp, stream);
  if (S % 2 == 0) return launch_t<T, 2>(ctx, fp, stream);
  return launch_t<T, 1>(ctx, fp, stream);
}

cudaError_t launch_draw_brute(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream) {
  if (ctx->cfg.flags & RT_FLAG_STRICT_IEEE) return launch_ch<sfloat>(ctx, fp, stream);
  return launch_ch<float>(ctx, fp, stream);
}

}  // namespace rt
