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
#ifndef RT_STRICT_MINBLOCKS  // the bit-exact kernels: 127 registers at 2 blocks per SM
#define RT_STRICT_MINBLOCKS 2
#endif
constexpr int kThreads = RT_THREADS, kTileW = 16, kTileH = kThreads / 16;
// SPLIT launches (four lanes per pixel, see rt_draw_fast.cu): a block covers 8 x 8 pixels, a warp 4 x 2
constexpr int kSplitTileW = 8, kSplitTileH = kThreads / 4 / 8;

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
// block = index of this block within its list (blockIdx.x, or blockIdx.x - n_split in a mixed launch); order / grid_x =
// that list's launch-order table and grid width.
template <bool SPLIT = false>
__device__ __forceinline__ bool pixel_of_thread(const FrameParams &p, int block, const int *order, int grid_x, int &x, int &y,
                                                int &tile_x, int &tile_y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int gb = block * p.blk_stride + p.blk_phase;
  if (order) gb = order[gb];
  const int by = gb / grid_x, bx = gb - by * grid_x;
  if constexpr (SPLIT) {
    const int pix = lane >> 2;  // 8 pixels per warp, 4 lanes each
    tile_x = bx * kSplitTileW;
    tile_y = p.row0 + by * kSplitTileH;
    x = tile_x + (warp & 1) * 4 + (pix & 3);
    y = tile_y + (warp >> 1) * 2 + (pix >> 2);
  } else {
    tile_x = bx * kTileW;
    tile_y = p.row0 + by * kTileH;
    x = tile_x + (warp & 1) * 8 + (lane & 7);
    y = tile_y + (warp >> 1) * 4 + (lane >> 3);
  }
  return x < p.W && y < p.row0 + p.rows;
}


__device__ __forceinline__ bool pixel_of_thread(const FrameParams &p, int &x, int &y, int &tile_x, int &tile_y) {
  return pixel_of_thread<false>(p, (int)blockIdx.x, p.tile_order, p.grid_x, x, y, tile_x, tile_y);
}

// rt_api.cu: device table of the launch order for a (row0, rows) range, built on first use
const int *tile_order_for(rt_ctx *ctx, int row0, int rows, int grid_x, int n_blocks, int tile_w, int tile_h);
// rt_api.cu: fills fp.vis_* (pixel rectangle outside which every primary ray misses the scene)
void visible_rect(const rt_ctx *ctx, FrameParams &fp);
// rt_api.cu: how should this launch map lanes to pixels? (flags, AA grid, size of the launch)
enum SplitMode { kSplitNone = 0, kSplitAll = 1, kSplitHeavy = 2 };
SplitMode split_mode(const rt_ctx *ctx, const FrameParams &fp);
// rt_api.cu: launch-order tables of a mixed launch for this camera: ordinary 16x16 tiles that cannot see a sphere, and the
// 8x8 sub-tiles of those that can (both centre-out).  Cached until the camera changes.  False: tables unavailable.
bool mixed_tables_for(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream, const int **light, int *n_light, const int **split,
                      int *n_split);

// float4 slots of the scene part of the fast kernel's shared memory (see brute_smem_bytes)
__host__ __device__ inline int scene_smem_float4(int n, int n_sh) { return 8 * n + 5 * n_sh + (n + n_sh + 3) / 4 + 1; }

template <class K>
inline cudaError_t launch_kernel(K kern, rt_ctx *ctx, const FrameParams &fp_in, cudaStream_t stream, int tile_w = kTileW,
                                 int tile_h = kTileH) {
  FrameParams fp = fp_in;
  fp.grid_x = (fp.W + tile_w - 1) / tile_w;
  fp.n_blocks = fp.grid_x * ((fp.rows + tile_h - 1) / tile_h);
  fp.blk_stride = ctx->cfg.block_stride > 1 ? ctx->cfg.block_stride : 1;
  fp.blk_phase = ctx->cfg.block_stride > 1 ? ctx->cfg.block_phase : 0;
  fp.tile_order = tile_order_for(ctx, fp.row0, fp.rows, fp.grid_x, fp.n_blocks, tile_w, tile_h);
  fp.n_split = 0;
  fp.split_grid_x = 1;
  fp.split_order = nullptr;
  const int my_blocks = (fp.n_blocks - fp.blk_phase + fp.blk_stride - 1) / fp.blk_stride;
  const size_t smem = brute_smem_bytes(ctx->n, ctx->n_sh) + ctx->launch_extra_smem;
  ctx->launch_extra_smem = 0;
  if (smem > 32 * 1024) {  // dynamic + the kernels' static shared memory (up to 4.2 KB) must stay under the 48 KB default
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (my_blocks <= 0) return cudaSuccess;
  kern<<<my_blocks, kThreads, smem, stream>>>(fp, ctx->d_scene, ctx->n, ctx->n_sh);
  ctx->launches++;
  return cudaGetLastError();
}

// Mixed launch: this rank's share of the split sub-tiles first (they are the expensive ones), then of the ordinary tiles.
template <class K>
inline cudaError_t launch_kernel_mixed(K kern, rt_ctx *ctx, const FrameParams &fp_in, cudaStream_t stream, const int *light, int n_light,
                                       const int *split, int n_split) {
  FrameParams fp = fp_in;
  fp.grid_x = (fp.W + kTileW - 1) / kTileW;
  fp.n_blocks = n_light;
  fp.blk_stride = ctx->cfg.block_stride > 1 ? ctx->cfg.block_stride : 1;
  fp.blk_phase = ctx->cfg.block_stride > 1 ? ctx->cfg.block_phase : 0;
  fp.tile_order = light;
  fp.split_order = split;
  fp.split_grid_x = (fp.W + kSplitTileW - 1) / kSplitTileW;
  const int my_light = n_light > fp.blk_phase ? (n_light - fp.blk_phase + fp.blk_stride - 1) / fp.blk_stride : 0;
  const int my_split = n_split > fp.blk_phase ? (n_split - fp.blk_phase + fp.blk_stride - 1) / fp.blk_stride : 0;
  fp.n_split = my_split;
  const size_t smem = brute_smem_bytes(ctx->n, ctx->n_sh) + ctx->launch_extra_smem;
  ctx->launch_extra_smem = 0;
  if (smem > 32 * 1024) {  // dynamic + the kernels' static shared memory (up to 4.2 KB) must stay under the 48 KB default
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (my_light + my_split <= 0) return cudaSuccess;
  kern<<<my_light + my_split, kThreads, smem, stream>>>(fp, ctx->d_scene, ctx->n, ctx->n_sh);
  ctx->launches++;
  return cudaGetLastError();
}

}  // namespace rt
