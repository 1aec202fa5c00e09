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
#ifndef RT_FAST_THREADS  // block size of the tuned kernels (rt_draw_fast.cu): their warps are independent workers
#define RT_FAST_THREADS 256
#endif
constexpr int kFastThreads = RT_FAST_THREADS;
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
  sc.tnd = nullptr;
  sc.n = n;
  sc.n_sh = n_sh;
  return sc;
}

// Where pixel row y of this launch is stored (FrameParams::outs: strips of rows dealt out over several frame buffers).
__device__ __forceinline__ uint32_t *frame_of_row(const FrameParams &p, int y) {
  return p.n_out > 1 ? p.outs[((unsigned)y / (unsigned)p.strip_rows) % (unsigned)p.n_out] : p.out;
}

// Which tile does this block render?  block = index of the block within its list (blockIdx.x, or blockIdx.x - n_split for
// the ordinary blocks of a mixed launch); the list is dealt over the ranks of a multi-GPU interleave (blk_stride / blk_phase)
// and, for ordinary blocks, permuted by the launch-order table.  (bx, by) = tile coordinates on the launch's tile grid
// (16x16 tiles, or 8x8 sub-tiles when SPLIT).  False: nothing to do for this block.
template <bool SPLIT, bool MIXED>
__device__ __forceinline__ bool tile_of_block(const FrameParams &p, int block, int &bx, int &by) {
  int gb = block * p.blk_stride + p.blk_phase;
  if constexpr (SPLIT && MIXED) {
    // sub-tiles of the sphere rectangles, row-major inside each rectangle, glass sphere first
    const int r = (p.n_rect > 1 && gb >= p.rect_first[1]) ? 1 : 0;
    const int local = gb - p.rect_first[r];
    const int sw = 2 * (p.rect[r][2] - p.rect[r][0]);
    const int sy = local / sw, sx = local - sy * sw;
    bx = 2 * p.rect[r][0] + sx;
    by = 2 * p.rect[r][1] + sy;
    if (r == 1 && (bx >> 1) >= p.rect[0][0] && (bx >> 1) < p.rect[0][2] && (by >> 1) >= p.rect[0][1] && (by >> 1) < p.rect[0][3])
      return false;  // rendered as part of rectangle 0
    return true;
  } else {
    if (p.tile_order) {  // launch-order table: tile coordinates packed as by << 16 | bx (no division per block)
      const int t = p.tile_order[gb];
      by = t >> 16;
      bx = t & 0xffff;
    } else {
      by = gb / p.grid_x;
      bx = gb - by * p.grid_x;
    }
    if constexpr (MIXED) {  // tiles inside a sphere rectangle belong to the split blocks
      for (int r = 0; r < p.n_rect; r++)
        if (bx >= p.rect[r][0] && bx < p.rect[r][2] && by >= p.rect[r][1] && by < p.rect[r][3]) return false;
    }
    return true;
  }
}

// Pixel tile of this block (top-left corner) and pixel of this thread; false if outside the frame rows.
template <bool SPLIT = false>
__device__ __forceinline__ bool pixel_of_thread(const FrameParams &p, int bx, int by, int &x, int &y, int &tile_x, int &tile_y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  int bx, by;
  tile_of_block<false, false>(p, (int)blockIdx.x, bx, by);
  return pixel_of_thread<false>(p, bx, by, x, y, tile_x, tile_y);
}

// rt_api.cu: device table of the launch order for a (row0, rows) range, built on first use
const int *tile_order_for(rt_ctx *ctx, int row0, int rows, int grid_x, int n_blocks, int tile_w, int tile_h);
// rt_api.cu: fills fp.vis_* (pixel rectangle outside which every primary ray misses the scene)
void visible_rect(const rt_ctx *ctx, FrameParams &fp);
// rt_api.cu: how should this launch map lanes to pixels? (flags, AA grid, size of the launch)
enum SplitMode { kSplitNone = 0, kSplitAll = 1, kSplitHeavy = 2 };
SplitMode split_mode(const rt_ctx *ctx, const FrameParams &fp);
// rt_api.cu: per-frame host work of the tuned kernels — triangle constants for this camera into ctx->d_fconst (copied on
// `stream` when the camera changed), fp.dmax, fp.sph_px.  Call once per frame, in front of the launch(es).
cudaError_t prepare_frame(rt_ctx *ctx, FrameParams &fp, cudaStream_t stream);
// rt_api.cu: the sphere rectangles of a mixed launch for this camera and row range (fills fp.n_rect, fp.rect, fp.rect_first);
// returns the number of 8x8 sub-tiles they hold.  Pure host arithmetic per launch: nothing is cached, copied or synchronised.
int sphere_rects(FrameParams &fp);
// rt_api.cu: BVH scenes — the one rectangle of tiles in which the mesh under the tree can be seen (same fields, same return value)
int mesh_rect(const rt_ctx *ctx, FrameParams &fp);

// float4 slots of the scene part of the fast kernel's shared memory (see brute_smem_bytes)
__host__ __device__ inline int scene_smem_float4(int n, int n_sh) { return 12 * n + 5 * n_sh + ((kFastThreads / 32) * n + n_sh + 3) / 4 + 1; }

// extra_smem: dynamic shared memory beyond the scene (fast kernels: fast_extra_smem); name: what rt_last_kernel_name reports
template <class K>
inline cudaError_t launch_kernel(K kern, rt_ctx *ctx, const FrameParams &fp_in, cudaStream_t stream, size_t extra_smem, const char *name,
                                 int tile_w = kTileW, int tile_h = kTileH) {
  FrameParams fp = fp_in;
  fp.grid_x = (fp.W + tile_w - 1) / tile_w;
  fp.n_blocks = fp.grid_x * ((fp.rows + tile_h - 1) / tile_h);
  fp.blk_stride = ctx->cfg.block_stride > 1 ? ctx->cfg.block_stride : 1;
  fp.blk_phase = ctx->cfg.block_stride > 1 ? ctx->cfg.block_phase : 0;
  fp.tile_order = tile_order_for(ctx, fp.row0, fp.rows, fp.grid_x, fp.n_blocks, tile_w, tile_h);
  fp.n_split = 0;
  fp.n_rect = 0;
  const int my_blocks = (fp.n_blocks - fp.blk_phase + fp.blk_stride - 1) / fp.blk_stride;
  const size_t smem = brute_smem_bytes(ctx->n, ctx->n_sh) + extra_smem;
  if (smem > 32 * 1024) {  // dynamic + the kernels' static shared memory (kDrawStaticSmem) must stay under the 48 KB default
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  ctx->last_kernel = name;
  if (my_blocks <= 0) return cudaSuccess;
  kern<<<my_blocks, kThreads, smem, stream>>>(fp, ctx->d_scene, ctx->n, ctx->n_sh);
  ctx->launches++;
  return cudaGetLastError();
}

// Mixed launch: this rank's share of the split sub-tiles first (they are the expensive ones), then of the ordinary tiles.
// fp_in carries the sphere rectangles (sphere_rects); n_sub = sub-tiles they hold.
template <class K>
inline cudaError_t launch_kernel_mixed(K kern, rt_ctx *ctx, const FrameParams &fp_in, cudaStream_t stream, int n_sub, size_t extra_smem,
                                       const char *name) {
  FrameParams fp = fp_in;
  fp.grid_x = (fp.W + kTileW - 1) / kTileW;
  fp.n_blocks = fp.grid_x * ((fp.rows + kTileH - 1) / kTileH);
  fp.blk_stride = ctx->cfg.block_stride > 1 ? ctx->cfg.block_stride : 1;
  fp.blk_phase = ctx->cfg.block_stride > 1 ? ctx->cfg.block_phase : 0;
  fp.tile_order = tile_order_for(ctx, fp.row0, fp.rows, fp.grid_x, fp.n_blocks, kTileW, kTileH);
  const int my_light = fp.n_blocks > fp.blk_phase ? (fp.n_blocks - fp.blk_phase + fp.blk_stride - 1) / fp.blk_stride : 0;
  const int my_split = n_sub > fp.blk_phase ? (n_sub - fp.blk_phase + fp.blk_stride - 1) / fp.blk_stride : 0;
  fp.n_split = my_split;
  const size_t smem = brute_smem_bytes(ctx->n, ctx->n_sh) + extra_smem;
  if (smem > 32 * 1024) {  // dynamic + the kernels' static shared memory (kDrawStaticSmem) must stay under the 48 KB default
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  ctx->last_kernel = name;
  if (my_light + my_split <= 0) return cudaSuccess;
  kern<<<my_light + my_split, kThreads, smem, stream>>>(fp, ctx->d_scene, ctx->n, ctx->n_sh);
  ctx->launches++;
  return cudaGetLastError();
}

}  // namespace rt
