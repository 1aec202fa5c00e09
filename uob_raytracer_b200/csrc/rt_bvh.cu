// rt_bvh.cu — GPU LBVH build (Morton codes -> radix sort -> Karras 2012 hierarchy -> bottom-up
// refit -> 64-byte two-child nodes) and the draw kernel that traverses it (rt_bvh.cuh).
#include <cub/device/device_radix_sort.cuh>

#include <vector>

#include <stdio.h>

#include "rt_bvh.cuh"
#include "rt_launch.cuh"

namespace rt {

// ---------------------------------------------------------------------------------------------
// build kernels
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ unsigned long long expand_bits21(unsigned long long v) {  // 21 bits -> every third bit
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}

// Per triangle (upload order): sort key = [big flag | Morton code of the box centre, `bits` per axis | upload index, idx_bits].
// Big triangles (box diagonal > big_diag) get bit 62 so that they sort behind everything the tree covers; the index in
// the low bits makes keys unique.  The Morton grid spans the box of the small triangles and is as fine as 62 bits allow
// (13 bits per axis at 1.3 M triangles: cells well below the size of a triangle, so neighbours in the order are
// neighbours in space down to the leaves).
__global__ void bvh_keys_kernel(const float4 *__restrict__ verts, int n, float3 lo, float3 inv_ext, float big_diag2, int bits, int idx_bits,
                                unsigned long long *__restrict__ keys, int *__restrict__ vals, int *__restrict__ n_small) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = verts[3 * (size_t)i], b = verts[3 * (size_t)i + 1], c = verts[3 * (size_t)i + 2];
  const float minx = fminf(a.x, fminf(b.x, c.x)), maxx = fmaxf(a.x, fmaxf(b.x, c.x));
  const float miny = fminf(a.y, fminf(b.y, c.y)), maxy = fmaxf(a.y, fmaxf(b.y, c.y));
  const float minz = fminf(a.z, fminf(b.z, c.z)), maxz = fmaxf(a.z, fmaxf(b.z, c.z));
  const float dx = maxx - minx, dy = maxy - miny, dz = maxz - minz;
  unsigned long long key;
  if (dx * dx + dy * dy + dz * dz > big_diag2) {
    key = (1ull << 62) | (unsigned)i;
  } else {
    const float cx = ((minx + maxx) * 0.5f - lo.x) * inv_ext.x, cy = ((miny + maxy) * 0.5f - lo.y) * inv_ext.y,
                cz = ((minz + maxz) * 0.5f - lo.z) * inv_ext.z;
    const float cells = (float)(1u << bits), top = cells - 1.0f;
    const unsigned long long mx = (unsigned long long)fminf(fmaxf(cx * cells, 0.0f), top), my = (unsigned long long)fminf(fmaxf(cy * cells, 0.0f), top),
                             mz = (unsigned long long)fminf(fmaxf(cz * cells, 0.0f), top);
    const unsigned long long morton = (expand_bits21(mx) << 2) | (expand_bits21(my) << 1) | expand_bits21(mz);
    key = (morton << idx_bits) | (unsigned long long)(unsigned)i;
    atomicAdd(n_small, 1);
  }
  keys[i] = key;
  vals[i] = i;
}

// Sorted slot s <- upload index order[s]: the per-triangle constants of the brute-force path
// (same single-rounded operations as rt_api.cu does on the host for small scenes) and the leaf box.
__global__ void bvh_gather_kernel(const float4 *__restrict__ verts, const float4 *__restrict__ normals, const float4 *__restrict__ colors,
                                  const int *__restrict__ order, int n, float4 *__restrict__ tri_a, float4 *__restrict__ tri_b,
                                  float4 *__restrict__ tri_c, float4 *__restrict__ tri_n, float4 *__restrict__ tri_col,
                                  int *__restrict__ tri_id, float4 *__restrict__ leaf_lo, float4 *__restrict__ leaf_hi,
                                  int *__restrict__ non_casters) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int i = order[s];
  const float4 v0 = verts[3 * (size_t)i], v1 = verts[3 * (size_t)i + 1], v2 = verts[3 * (size_t)i + 2];
  const float e1x = __fsub_rn(v1.x, v0.x), e1y = __fsub_rn(v1.y, v0.y), e1z = __fsub_rn(v1.z, v0.z);
  const float e2x = __fsub_rn(v2.x, v0.x), e2y = __fsub_rn(v2.y, v0.y), e2z = __fsub_rn(v2.z, v0.z);
  const float c0 = __fsub_rn(__fmul_rn(e1y, e2z), __fmul_rn(e1z, e2y));
  const float c1 = __fsub_rn(__fmul_rn(e1x, e2z), __fmul_rn(e1z, e2x));
  const float c2 = __fsub_rn(__fmul_rn(e1x, e2y), __fmul_rn(e1y, e2x));
  tri_a[s] = make_float4(v0.x, v0.y, v0.z, c0);
  tri_b[s] = make_float4(e1x, e1y, e1z, c1);
  tri_c[s] = make_float4(e2x, e2y, e2z, c2);
  const float4 nn = normals[i];
  tri_n[s] = make_float4(nn.x, nn.y, nn.z, 0.0f);
  const float4 col = colors[i];
  tri_col[s] = col;
  tri_id[s] = i;
  if (col.w == -1.0f) atomicAdd(non_casters, 1);
  leaf_lo[s] = make_float4(fminf(v0.x, fminf(v1.x, v2.x)), fminf(v0.y, fminf(v1.y, v2.y)), fminf(v0.z, fminf(v1.z, v2.z)), 0.0f);
  leaf_hi[s] = make_float4(fmaxf(v0.x, fmaxf(v1.x, v2.x)), fmaxf(v0.y, fmaxf(v1.y, v2.y)), fmaxf(v0.z, fmaxf(v1.z, v2.z)), 0.0f);
}

// Bounds used by the fast policy's per-(point, triangle) culls for the triangles kept out of the tree.
__global__ void bvh_big_bounds_kernel(const float4 *__restrict__ tri_a, const float4 *__restrict__ tri_b, const float4 *__restrict__ tri_c,
                                      int n_bvh, int n, float4 *__restrict__ out) {
  const int s = n_bvh + blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const float4 A = tri_a[s], B = tri_b[s], C = tri_c[s];
  out[s - n_bvh] = make_float4(kJitterMax * sqrtf(A.w * A.w + B.w * B.w + C.w * C.w), kJitterMax * sqrtf(B.x * B.x + B.y * B.y + B.z * B.z),
                               kJitterMax * sqrtf(C.x * C.x + C.y * C.y + C.z * C.z), 0.0f);
}

__device__ __forceinline__ int key_delta(const unsigned long long *keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  return __clzll((long long)(keys[i] ^ keys[j]));
}

// Karras 2012: internal node i of n-1.  Children are stored as refs: >= 0 internal node, < 0 ~leaf slot.
__global__ void bvh_hierarchy_kernel(const unsigned long long *__restrict__ keys, int n, int *__restrict__ left, int *__restrict__ right,
                                     int *__restrict__ first, int *__restrict__ last, int *__restrict__ parent_node,
                                     int *__restrict__ parent_leaf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = (key_delta(keys, n, i, i + 1) - key_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = key_delta(keys, n, i, i - d);
  int lmax = 2;
  while (key_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (key_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = key_delta(keys, n, i, j);
  int s = 0, t = l;
  do {
    t = (t + 1) / 2;
    if (key_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  if (lo == gamma) {
    left[i] = ~gamma;
    parent_leaf[gamma] = i;
  } else {
    left[i] = gamma;
    parent_node[gamma] = i;
  }
  if (hi == gamma + 1) {
    right[i] = ~(gamma + 1);
    parent_leaf[gamma + 1] = i;
  } else {
    right[i] = gamma + 1;
    parent_node[gamma + 1] = i;
  }
  first[i] = lo;
  last[i] = hi;
  if (i == 0) parent_node[0] = -1;
}

// Bottom-up boxes: one thread per leaf; the second thread to reach a node merges its children.
__global__ void bvh_refit_kernel(int n, const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ parent_node,
                                 const int *__restrict__ parent_leaf, const float4 *__restrict__ leaf_lo, const float4 *__restrict__ leaf_hi,
                                 float4 *node_lo, float4 *node_hi, int *__restrict__ visits) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int cur = parent_leaf[s];
  while (cur >= 0) {
    __threadfence();
    if (atomicAdd(&visits[cur], 1) == 0) return;  // first arrival: the sibling subtree is not done yet
    __threadfence();
    const int l = left[cur], r = right[cur];
    const volatile float4 *llo = (l < 0) ? leaf_lo + ~l : node_lo + l, *lhi = (l < 0) ? leaf_hi + ~l : node_hi + l;
    const volatile float4 *rlo = (r < 0) ? leaf_lo + ~r : node_lo + r, *rhi = (r < 0) ? leaf_hi + ~r : node_hi + r;
    node_lo[cur] = make_float4(fminf(llo->x, rlo->x), fminf(llo->y, rlo->y), fminf(llo->z, rlo->z), 0.0f);
    node_hi[cur] = make_float4(fmaxf(lhi->x, rhi->x), fmaxf(lhi->y, rhi->y), fmaxf(lhi->z, rhi->z), 0.0f);
    cur = parent_node[cur];
  }
}

// Final 64-byte nodes with padded child boxes; subtrees of <= kBvhLeafMax triangles become leaves.
__global__ void bvh_emit_kernel(int n, const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ first,
                                const int *__restrict__ last, const float4 *__restrict__ leaf_lo, const float4 *__restrict__ leaf_hi,
                                const float4 *__restrict__ node_lo, const float4 *__restrict__ node_hi, float pad, float4 *__restrict__ nodes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int ref[2];
  float4 lo[2], hi[2];
  const int child[2] = {left[i], right[i]};
#pragma unroll
  for (int c = 0; c < 2; c++) {
    const int k = child[c];
    if (k < 0) {
      ref[c] = ~((~k) << 3);
      lo[c] = leaf_lo[~k];
      hi[c] = leaf_hi[~k];
    } else {
      const int cnt = last[k] - first[k] + 1;
      ref[c] = (cnt <= kBvhLeafMax) ? ~((first[k] << 3) | (cnt - 1)) : k;
      lo[c] = node_lo[k];
      hi[c] = node_hi[k];
    }
    // pad: absolute + relative to the coordinates, far above float rounding of the slab test
    const float px = pad + 1e-6f * fmaxf(fabsf(lo[c].x), fabsf(hi[c].x)), py = pad + 1e-6f * fmaxf(fabsf(lo[c].y), fabsf(hi[c].y)),
                pz = pad + 1e-6f * fmaxf(fabsf(lo[c].z), fabsf(hi[c].z));
    lo[c].x -= px; lo[c].y -= py; lo[c].z -= pz;
    hi[c].x += px; hi[c].y += py; hi[c].z += pz;
  }
  float4 *o = nodes + 4 * (size_t)i;
  o[0] = make_float4(lo[0].x, lo[0].y, lo[0].z, hi[0].x);
  o[1] = make_float4(hi[0].y, hi[0].z, lo[1].x, lo[1].y);
  o[2] = make_float4(lo[1].z, hi[1].x, hi[1].y, hi[1].z);
  o[3] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), 0.0f, 0.0f);
}

// ---------------------------------------------------------------------------------------------
// draw kernel
// ---------------------------------------------------------------------------------------------

constexpr int kBigPrimaryMax = 32;  // big triangles handled by the camera-constant filter (one warp bins them)

// QUAD blocks: 8x8-pixel sub-tiles with four lanes per pixel (one ray of the 2x2 each), the tiles the mesh can be seen in;
// the other blocks: 16x16 tiles, one lane per pixel.  Same frame either way (rt_brute.cuh: shade_pixel_quad).
template <class T, int CH, bool QUAD>
__device__ __forceinline__ void draw_bvh_body(const FrameParams &p, const BvhView &bv, int bx, int by, float4 *s_prim, int *s_plist, int *s_nlist) {
  int x, y, tx, ty;
  const bool in_frame = pixel_of_thread<QUAD>(p, bx, by, x, y, tx, ty);
  constexpr int TW = QUAD ? kSplitTileW : kTileW, TH = QUAD ? kSplitTileH : kTileH;
  BvhTracer<T> tr;
  tr.bv = bv;
  const int n_big = bv.n - bv.n_bvh;
  if constexpr (!is_strict<T>::value) {
    if (n_big > 0 && n_big <= kBigPrimaryMax) {  // uniform
      if (threadIdx.x < 32) {
        // per-frame camera constants of the big triangles and their binning against this block's tile
        // (same as the prologue of draw_fast_kernel)
        const int lane = threadIdx.x, A = p.A;
        const float SW = (float)p.W, SH = (float)p.H, fA = (float)A;
        const float vx0 = (float)(tx * A) - SW * fA * 0.5f, vy0 = (float)(ty * A) - SH * fA * 0.5f;
        const float vx1 = vx0 + (float)(TW * A - 1), vy1 = vy0 + (float)(TH * A - 1);
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
        bool keep = false;
        if (lane < n_big) {
          SceneView g;
          g.ta = bv.tri_a + bv.n_bvh;
          g.tb = bv.tri_b + bv.n_bvh;
          g.tc = bv.tri_c + bv.n_bvh;
          primary_constants(g, s_prim, V3<float>(p.cam[0], p.cam[1], p.cam[2]), lane);
          keep = tile_may_hit(s_prim, lane, dc, dmax);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (keep) s_plist[__popc(ballot & ((1u << lane) - 1u))] = lane;
        if (lane == 0) *s_nlist = __popc(ballot);
      }
      __syncthreads();
      tr.big.prim = s_prim;
      tr.big.plist = s_plist;
      tr.big.n_list = *s_nlist;
      tr.big.n_big = n_big;
    }
  }
  if constexpr (QUAD) {
    const uint32_t px = shade_pixel_quad<T, CH, BvhTracer<T>>(tr, p, x, y, in_frame);
    if (in_frame && (threadIdx.x & 3) == 0) p.out[(size_t)y * p.W + x] = px;
  } else {
    if (!in_frame) return;
    p.out[(size_t)y * p.W + x] = shade_pixel<T, CH, BvhTracer<T>>(tr, p, x, y);
  }
}

#ifndef RT_BVH_MINBLOCKS  // resident blocks per SM the BVH kernel is compiled for; cfg4 fast / strict: 2 -> 1.48 / 2.73 ms, 3 -> 1.42 / 2.45, 4 -> 1.53 / 2.54
#define RT_BVH_MINBLOCKS 3
#endif
template <class T, int CH>
__global__ void __launch_bounds__(kThreads, RT_BVH_MINBLOCKS) draw_bvh_kernel(const __grid_constant__ FrameParams p, const __grid_constant__ BvhView bv) {
  __shared__ float4 s_prim[3 * kBigPrimaryMax];
  __shared__ int s_plist[kBigPrimaryMax];
  __shared__ int s_nlist;
  int bx, by;
  if ((int)blockIdx.x < p.n_split) {  // the mesh's tiles first: they hold the longest rays of the frame
    if (tile_of_block<true, true>(p, (int)blockIdx.x, bx, by)) draw_bvh_body<T, CH, true>(p, bv, bx, by, s_prim, s_plist, &s_nlist);
  } else {
    if (tile_of_block<false, true>(p, (int)blockIdx.x - p.n_split, bx, by)) draw_bvh_body<T, CH, false>(p, bv, bx, by, s_prim, s_plist, &s_nlist);
  }
}

template <class T, int CH>
static cudaError_t launch_bvh_t(rt_ctx *ctx, const FrameParams &fp_in, cudaStream_t stream) {
  FrameParams fp = fp_in;
  fp.grid_x = (fp.W + kTileW - 1) / kTileW;
  fp.n_blocks = fp.grid_x * ((fp.rows + kTileH - 1) / kTileH);
  fp.blk_stride = ctx->cfg.block_stride > 1 ? ctx->cfg.block_stride : 1;
  fp.blk_phase = ctx->cfg.block_stride > 1 ? ctx->cfg.block_phase : 0;
  fp.tile_order = tile_order_for(ctx, fp.row0, fp.rows, fp.grid_x, fp.n_blocks, kTileW, kTileH);
  // the tiles in which the mesh can be seen: four lanes per pixel (2x2 rays only), at the head of the launch
  int n_sub = 0;
  fp.n_rect = 0;
  fp.rect_first[0] = fp.rect_first[1] = 0;
  if (fp.A == 2 && !(ctx->cfg.flags & RT_FLAG_NO_SPLIT)) n_sub = mesh_rect(ctx, fp);
  const int my_tiles = fp.n_blocks > fp.blk_phase ? (fp.n_blocks - fp.blk_phase + fp.blk_stride - 1) / fp.blk_stride : 0;
  const int my_split = n_sub > fp.blk_phase ? (n_sub - fp.blk_phase + fp.blk_stride - 1) / fp.blk_stride : 0;
  fp.n_split = my_split;
  const int my_blocks = my_tiles + my_split;
  char name[64];
  snprintf(name, sizeof name, "draw_bvh_kernel<%s,%d>%s", is_strict<T>::value ? "sfloat" : "float", CH, my_split ? "+quad" : "");
  ctx->last_kernel = name;
  if (my_blocks <= 0) return cudaSuccess;
  draw_bvh_kernel<T, CH><<<my_blocks, kThreads, 0, stream>>>(fp, *static_cast<const BvhView *>(ctx->bvh_view));
  ctx->launches++;
  return cudaGetLastError();
}

cudaError_t launch_draw_bvh(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream) {
  const bool strict = (ctx->cfg.flags & RT_FLAG_STRICT_IEEE) != 0;
  const int S = fp.S;
#define RT_BVH(CH) (strict ? launch_bvh_t<sfloat, CH>(ctx, fp, stream) : launch_bvh_t<float, CH>(ctx, fp, stream))
  if (S % 10 == 0) return RT_BVH(10);
  if (S % 8 == 0) return RT_BVH(8);
  if (S % 4 == 0) return RT_BVH(4);
  if (S % 2 == 0) return RT_BVH(2);
  return RT_BVH(1);
#undef RT_BVH
}

// ---------------------------------------------------------------------------------------------
// host side of the build
// ---------------------------------------------------------------------------------------------

void bvh_free(rt_ctx *ctx) {
  for (void *p : ctx->bvh_allocs) cudaFree(p);
  ctx->bvh_allocs.clear();
  delete static_cast<BvhView *>(ctx->bvh_view);
  ctx->bvh_view = nullptr;
}

#define BVH_CHECK(call)              \
  do {                               \
    cudaError_t e_ = (call);         \
    if (e_ != cudaSuccess) return e_; \
  } while (0)

template <class U> static cudaError_t dev_alloc(rt_ctx *ctx, U **p, size_t count, bool keep) {
  void *q = nullptr;
  cudaError_t e = cudaMalloc(&q, sizeof(U) * (count ? count : 1));
  if (e != cudaSuccess) return e;
  *p = static_cast<U *>(q);
  if (keep) ctx->bvh_allocs.push_back(q);
  return cudaSuccess;
}

// verts/normals/colors: HOST arrays of the C ABI (3n, n, n float4).
cudaError_t bvh_build(rt_ctx *ctx, const float *verts, const float *normals, const float *colors, int n) {
  bvh_free(ctx);
  cudaStream_t st = ctx->stream;
  // scene bounds on the host (one pass over the vertices; everything else happens on the device)
  float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (size_t v = 0; v < 3 * (size_t)n; v++)
    for (int c = 0; c < 3; c++) {
      const float x = verts[4 * v + c];
      lo[c] = x < lo[c] ? x : lo[c];
      hi[c] = x > hi[c] ? x : hi[c];
    }
  float ext[3], diag2 = 0.0f;
  for (int c = 0; c < 3; c++) {
    ext[c] = hi[c] - lo[c];
    if (!(ext[c] > 0.0f)) ext[c] = 1.0f;
    diag2 += ext[c] * ext[c];
  }
  const float big_diag2 = 0.35f * 0.35f * diag2;  // triangles longer than 35 % of the scene diagonal stay out of the tree
  // bounds of what the tree will cover (same criterion as bvh_keys_kernel): where on the screen the long rays are
  for (int c = 0; c < 3; c++) {
    ctx->mesh_lo[c] = 3.4e38f;
    ctx->mesh_hi[c] = -3.4e38f;
  }
  for (size_t t = 0; t < (size_t)n; t++) {
    float tlo[3], thi[3], d2 = 0.0f;
    for (int c = 0; c < 3; c++) {
      const float a = verts[12 * t + c], b = verts[12 * t + 4 + c], cc = verts[12 * t + 8 + c];
      tlo[c] = fminf(a, fminf(b, cc));
      thi[c] = fmaxf(a, fmaxf(b, cc));
      d2 += (thi[c] - tlo[c]) * (thi[c] - tlo[c]);
    }
    if (d2 > big_diag2) continue;
    for (int c = 0; c < 3; c++) {
      ctx->mesh_lo[c] = fminf(ctx->mesh_lo[c], tlo[c]);
      ctx->mesh_hi[c] = fmaxf(ctx->mesh_hi[c], thi[c]);
    }
  }
  const float pad = 1e-5f * sqrtf(diag2);

  float4 *d_verts = nullptr, *d_normals = nullptr, *d_colors = nullptr;
  unsigned long long *d_keys = nullptr, *d_keys2 = nullptr;
  int *d_vals = nullptr, *d_vals2 = nullptr, *d_counters = nullptr;
  std::vector<void *> temps;
  auto cleanup = [&]() {
    for (void *p : temps) cudaFree(p);
  };
#define TMP_ALLOC(ptr, count)                                  \
  do {                                                         \
    cudaError_t e2_ = dev_alloc(ctx, &ptr, (count), false);    \
    if (e2_ != cudaSuccess) { cleanup(); return e2_; }         \
    temps.push_back(ptr);                                      \
  } while (0)
#define BVH_TRY(call)                                          \
  do {                                                         \
    cudaError_t e2_ = (call);                                  \
    if (e2_ != cudaSuccess) { cleanup(); return e2_; }         \
  } while (0)

  TMP_ALLOC(d_verts, 3 * (size_t)n);
  TMP_ALLOC(d_normals, (size_t)n);
  TMP_ALLOC(d_colors, (size_t)n);
  TMP_ALLOC(d_keys, (size_t)n);
  TMP_ALLOC(d_keys2, (size_t)n);
  TMP_ALLOC(d_vals, (size_t)n);
  TMP_ALLOC(d_vals2, (size_t)n);
  TMP_ALLOC(d_counters, 2);
  BVH_TRY(cudaMemcpyAsync(d_verts, verts, sizeof(float4) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
  BVH_TRY(cudaMemcpyAsync(d_normals, normals, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
  BVH_TRY(cudaMemcpyAsync(d_colors, colors, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
  BVH_TRY(cudaMemsetAsync(d_counters, 0, 2 * sizeof(int), st));

  const int tpb = 256, blocks = (n + tpb - 1) / tpb;
  int idx_bits = 1;
  while (idx_bits < 31 && (1ll << idx_bits) < (long long)n) idx_bits++;
  const int morton_bits = std::min(21, (62 - idx_bits) / 3);
  float mlo[3], minv[3];  // Morton grid: the box of the small triangles (the scene box if there is none)
  for (int c = 0; c < 3; c++) {
    const bool have = ctx->mesh_lo[c] <= ctx->mesh_hi[c];
    const float l = have ? ctx->mesh_lo[c] : lo[c], e = have ? ctx->mesh_hi[c] - ctx->mesh_lo[c] : ext[c];
    mlo[c] = l;
    minv[c] = 1.0f / (e > 0.0f ? e : 1.0f);
  }
  bvh_keys_kernel<<<blocks, tpb, 0, st>>>(d_verts, n, make_float3(mlo[0], mlo[1], mlo[2]), make_float3(minv[0], minv[1], minv[2]), big_diag2,
                                          morton_bits, idx_bits, d_keys, d_vals, d_counters);
  BVH_TRY(cudaGetLastError());
  size_t sort_bytes = 0;
  BVH_TRY(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, d_keys, d_keys2, d_vals, d_vals2, n, 0, 64, st));
  unsigned char *d_sort = nullptr;
  TMP_ALLOC(d_sort, sort_bytes);
  BVH_TRY(cub::DeviceRadixSort::SortPairs(d_sort, sort_bytes, d_keys, d_keys2, d_vals, d_vals2, n, 0, 64, st));

  BvhView *bv = new BvhView();
  ctx->bvh_view = bv;
  float4 *tri_a, *tri_b, *tri_c, *tri_n, *tri_col, *leaf_lo, *leaf_hi;
  int *tri_id;
  BVH_TRY(dev_alloc(ctx, &tri_a, n, true));
  BVH_TRY(dev_alloc(ctx, &tri_b, n, true));
  BVH_TRY(dev_alloc(ctx, &tri_c, n, true));
  BVH_TRY(dev_alloc(ctx, &tri_n, n, true));
  BVH_TRY(dev_alloc(ctx, &tri_col, n, true));
  BVH_TRY(dev_alloc(ctx, &tri_id, n, true));
  TMP_ALLOC(leaf_lo, (size_t)n);
  TMP_ALLOC(leaf_hi, (size_t)n);
  bvh_gather_kernel<<<blocks, tpb, 0, st>>>(d_verts, d_normals, d_colors, d_vals2, n, tri_a, tri_b, tri_c, tri_n, tri_col, tri_id, leaf_lo,
                                            leaf_hi, d_counters + 1);
  BVH_TRY(cudaGetLastError());
  int counters[2] = {0, 0};
  BVH_TRY(cudaMemcpyAsync(counters, d_counters, sizeof counters, cudaMemcpyDeviceToHost, st));
  BVH_TRY(cudaStreamSynchronize(st));
  int n_bvh = counters[0];
  if (n_bvh < 2) n_bvh = 0;  // nothing worth a tree: everything is tested linearly (slots keep the sorted order)

  float4 *nodes = nullptr;
  if (n_bvh >= 2) {
    int *left, *right, *first, *last, *parent_node, *parent_leaf, *visits;
    float4 *node_lo, *node_hi;
    TMP_ALLOC(left, (size_t)n_bvh);
    TMP_ALLOC(right, (size_t)n_bvh);
    TMP_ALLOC(first, (size_t)n_bvh);
    TMP_ALLOC(last, (size_t)n_bvh);
    TMP_ALLOC(parent_node, (size_t)n_bvh);
    TMP_ALLOC(parent_leaf, (size_t)n_bvh);
    TMP_ALLOC(visits, (size_t)n_bvh);
    TMP_ALLOC(node_lo, (size_t)n_bvh);
    TMP_ALLOC(node_hi, (size_t)n_bvh);
    BVH_TRY(dev_alloc(ctx, &nodes, 4 * (size_t)(n_bvh - 1), true));
    BVH_TRY(cudaMemsetAsync(visits, 0, sizeof(int) * (size_t)n_bvh, st));
    const int b2 = (n_bvh + tpb - 1) / tpb;
    bvh_hierarchy_kernel<<<b2, tpb, 0, st>>>(d_keys2, n_bvh, left, right, first, last, parent_node, parent_leaf);
    BVH_TRY(cudaGetLastError());
    bvh_refit_kernel<<<b2, tpb, 0, st>>>(n_bvh, left, right, parent_node, parent_leaf, leaf_lo, leaf_hi, node_lo, node_hi, visits);
    BVH_TRY(cudaGetLastError());
    bvh_emit_kernel<<<b2, tpb, 0, st>>>(n_bvh, left, right, first, last, leaf_lo, leaf_hi, node_lo, node_hi, pad, nodes);
    BVH_TRY(cudaGetLastError());
  }
  float4 *big_bound = nullptr;
  BVH_TRY(dev_alloc(ctx, &big_bound, (size_t)(n - n_bvh), true));
  if (n > n_bvh) {
    bvh_big_bounds_kernel<<<(n - n_bvh + tpb - 1) / tpb, tpb, 0, st>>>(tri_a, tri_b, tri_c, n_bvh, n, big_bound);
    BVH_TRY(cudaGetLastError());
  }
  BVH_TRY(cudaStreamSynchronize(st));
  cleanup();
  bv->big_bound = big_bound;
  bv->tri_a = tri_a;
  bv->tri_b = tri_b;
  bv->tri_c = tri_c;
  bv->tri_n = tri_n;
  bv->tri_col = tri_col;
  bv->tri_id = tri_id;
  bv->nodes = nodes;
  bv->n = n;
  bv->n_bvh = n_bvh;
  bv->root = (n_bvh >= 2 && n_bvh <= kBvhLeafMax) ? ~((0 << 3) | (n_bvh - 1)) : 0;
  bv->all_casters = counters[1] == 0;
  ctx->launches += 6;
  return cudaSuccess;
#undef TMP_ALLOC
#undef BVH_TRY
}

}  // namespace rt
