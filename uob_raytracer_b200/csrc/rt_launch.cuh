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

__device__ __forceinline__ bool pixel_of_thread(const FrameParams &p, int &x, int &y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  x = blockIdx.x * kTileW + (warp & 1) * 8 + (lane & 7);
  y = p.row0 + blockIdx.y * kTileH + (warp >> 1) * 4 + (lane >> 3);
  return x < p.W && y < p.row0 + p.rows;
}


// float4 slots of the scene part of the fast kernel's shared memory (see brute_smem_bytes)
__host__ __device__ inline int scene_smem_float4(int n, int n_sh) { return 8 * n + 4 * n_sh + (n + 3) / 4 + 1; }

template <class K>
inline cudaError_t launch_kernel(K kern, rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream) {
  const size_t smem = brute_smem_bytes(ctx->n, ctx->n_sh) + ctx->launch_extra_smem;
  ctx->launch_extra_smem = 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((fp.W + kTileW - 1) / kTileW, (fp.rows + kTileH - 1) / kTileH);
  kern<<<grid, kThreads, smem, stream>>>(fp, ctx->d_scene, ctx->n, ctx->n_sh);
  ctx->launches++;
  return cudaGetLastError();
}


}  // namespace rt
