// rt_launch.cuh — block geometry and launch helper shared by the draw kernels.
#pragma once
#include "rt_fast.cuh"
#include "rt_internal.h"

namespace rt {

// Block = kThreads threads; a warp covers an 8x4 pixel tile (coherent rays, and each row of the
// tile is one full 32-byte sector of the ARGB frame); a block covers 16 x (kThreads/16) pixels.
#ifndef RT_THREADS
#define RT_THREADS 256
#endif
#ifndef RT_MINBLOCKS
#define RT_MINBLOCKS 3
#endif
constexpr int kThreads = RT_THREADS, kTileW = 16, kTileH = kThreads / 16;

__device__ __forceinline__ SceneView stage_scene(float4 *smem, const float4 *__restrict__ scene, int n, int n_sh) {
  const int total = 5 * n + 3 * n_sh;
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
  sc.n = n;
  sc.n_sh = n_sh;
  return sc;
}

// Pixel tile of this block (top-left corner) and pixel of this thread; false if outside the frame rows.
__device__ __forceinline__ bool pixel_of_thread(const FrameParams &p, int &x, int &y, int &tile_x, int &tile_y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int gb = blockIdx.x * p.blk_stride + p.blk_phase;
  if (p.tile_order) gb = p.tile_order[gb];
  const int by = gb / p.grid_x, bx = gb - by * p.grid_x;
  tile_x = bx * kTileW;
  tile_y = p.row0 + by * kTileH;
  x = tile_x + (warp & 1) * 8 + (lane & 7);
  y = tile_y + (warp >> 1) * 4 + (lane >> 3);
  return x < p.W && y < p.row0 + p.rows;
}


// rt_api.cu: device table of the launch order for a (row0, rows) range, built on first use
const int *tile_order_for(rt_ctx *ctx, int row0, int rows, int grid_x, int n_blocks);

// float4 slots of the scene part of the fast kernel's shared memory (see brute_smem_bytes)
__host__ __device__ inline int scene_smem_float4(int n, int n_sh) { return 8 * n + 5 * n_sh + (n + n_sh + 3) / 4 + 1; }

template <class K>
inline cudaError_t launch_kernel(K kern, rt_ctx *ctx, const FrameParams &fp_in, cudaStream_t stream) {
  FrameParams fp = fp_in;
  fp.grid_x = (fp.W + kTileW - 1) / kTileW;
  fp.n_blocks = fp.grid_x * ((fp.rows + kTileH - 1) / kTileH);
  fp.blk_stride = ctx->cfg.block_stride > 1 ? ctx->cfg.block_stride : 1;
  fp.blk_phase = ctx->cfg.block_stride > 1 ? ctx->cfg.block_phase : 0;
  fp.tile_order = tile_order_for(ctx, fp.row0, fp.rows, fp.grid_x, fp.n_blocks);
  const int my_blocks = (fp.n_blocks - fp.blk_phase + fp.blk_stride - 1) / fp.blk_stride;
  const size_t smem = brute_smem_bytes(ctx->n, ctx->n_sh) + ctx->launch_extra_smem;
  ctx->launch_extra_smem = 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (my_blocks <= 0) return cudaSuccess;
  kern<<<my_blocks, kThreads, smem, stream>>>(fp, ctx->d_scene, ctx->n, ctx->n_sh);
  ctx->launches++;
  return cudaGetLastError();
}


}  // namespace rt
