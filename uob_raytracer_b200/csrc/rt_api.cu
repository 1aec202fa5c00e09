// rt_api.cu — implementation of the C ABI in include/uob_rt.h.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <utility>

#include "rt_internal.h"
#include "rt_launch.cuh"
#include "rt_types.h"

static thread_local std::string g_create_err;

#define RT_CUDA(ctx, call, what)                                                                    \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) {                                                                        \
      char buf_[512];                                                                               \
      snprintf(buf_, sizeof buf_, "CUDA error during '%s': %s (%d)", what, cudaGetErrorString(e_), (int)e_); \
      (ctx)->err = buf_;                                                                            \
      return RT_ERR_CUDA;                                                                           \
    }                                                                                               \
  } while (0)

namespace rt {

// Blocks start in launch order.  The expensive tiles of a frame (spheres, blocks, penumbrae) sit
// around the image centre and the cheapest (background outside the box) at its edges; in row-major
// order the last rows start when the frame is almost done and their long blocks run alone.
// Starting tiles centre-out removes that tail (60 us of a 355 us 1080p frame).
// A launch-order table sorted by distance keeps mirror-image tiles next to each other, always in the same order (left
// before right, top before bottom), so an N-way interleave would hand one rank the same side of every ring — and the
// sides of the box do not cost the same.  Shuffling inside short windows keeps the overall order (expensive first)
// and decorrelates the deal from the geometry.  Deterministic: every rank builds the same table.
static void shuffle_windows(std::vector<int> &order, size_t begin, size_t end, size_t window = 32) {
  for (size_t w0 = begin; w0 < end; w0 += window) {
    const size_t w1 = std::min(end, w0 + window);
    uint32_t s = (uint32_t)(w0 * 2654435761u) ^ 0x9e3779b9u;
    for (size_t i = w1 - 1; i > w0; i--) {
      s ^= s << 13;
      s ^= s >> 17;
      s ^= s << 5;
      std::swap(order[i], order[w0 + s % (uint32_t)(i - w0 + 1)]);
    }
  }
}

// Four lanes per pixel pay off when the launch cannot fill the GPU for long: its duration is then set by its slowest
// blocks, the tiles whose pixels run mirror / glass bounce chains.  kSplitHeavy (mixed launch) splits only those tiles.
// RT_FLAG_SPLIT_PIXELS / RT_FLAG_SPLIT_HEAVY / RT_FLAG_NO_SPLIT force a mode; default: by the size of the launch.
// Measured on B200 (scripts/gpu_split.py; 1/N block-interleaved shares of cfg2, HEAD and cfg3): the mixed launch wins
// below ~10 k pixels per SM (1080p on 2+ GPUs: 129 -> 112 us at 1/2, 87 -> 56 us at 1/8; the 1024^2 HEAD frame on one
// GPU: 221 -> 214 us; 4K on 8 GPUs: 427 -> 329 us) and loses a few per cent above (4K on 2 GPUs).
constexpr double kSplitBelowPixelsPerSm = 10000.0;
SplitMode split_mode(const rt_ctx *ctx, const FrameParams &fp) {
  const int R = fp.A * fp.A;
  if (R != 4 && R != 16) return kSplitNone;
  if (ctx->cfg.flags & RT_FLAG_NO_SPLIT) return kSplitNone;
  if (ctx->cfg.flags & RT_FLAG_SPLIT_PIXELS) return kSplitAll;
  if (ctx->cfg.flags & RT_FLAG_SPLIT_HEAVY) return kSplitHeavy;
  const int stride = ctx->cfg.block_stride > 1 ? ctx->cfg.block_stride : 1;
  const double pixels_per_sm = (double)fp.W * fp.rows / stride / (ctx->sm_count > 0 ? ctx->sm_count : 148);
  return pixels_per_sm < kSplitBelowPixelsPerSm ? kSplitHeavy : kSplitNone;
}

// Pixel rectangle that contains every primary ray able to hit the scene: the eight corners of the scene's bounding box
// seen from the camera.  A primary ray is cam + t R (vx, vy, f) with R the rotation matrix (rows r0, r1, r2), so a
// point P lies on the ray through (vx, vy) = f (w.x, w.y) / w.z with w = R^-1 (P - cam).  The hull of the projected
// corners contains the projection of the box as long as all corners are in front of the camera; otherwise (or if R is
// singular) the rectangle is the whole frame.  Two pixels of margin absorb the rounding of this float arithmetic.
static bool inverse_rotation(const FrameParams &fp, double inv[9]) {
  const float *m = fp.rot;
  const double det = (double)m[0] * (m[4] * m[8] - m[5] * m[7]) - (double)m[1] * (m[3] * m[8] - m[5] * m[6]) +
                     (double)m[2] * (m[3] * m[7] - m[4] * m[6]);
  if (!(fabs(det) > 1e-6) || !(fp.focal > 0.0f)) return false;
  inv[0] = (m[4] * m[8] - m[5] * m[7]) / det;
  inv[1] = (m[2] * m[7] - m[1] * m[8]) / det;
  inv[2] = (m[1] * m[5] - m[2] * m[4]) / det;
  inv[3] = (m[5] * m[6] - m[3] * m[8]) / det;
  inv[4] = (m[0] * m[8] - m[2] * m[6]) / det;
  inv[5] = (m[2] * m[3] - m[0] * m[5]) / det;
  inv[6] = (m[3] * m[7] - m[4] * m[6]) / det;
  inv[7] = (m[1] * m[6] - m[0] * m[7]) / det;
  inv[8] = (m[0] * m[4] - m[1] * m[3]) / det;
  return true;
}

// pixel position (fractional) of the world point cam + P_rel seen through the camera; false if beside / behind it
static bool project_point(const FrameParams &fp, const double inv[9], const double P[3], double *px, double *py, double *depth) {
  const double wx = inv[0] * P[0] + inv[1] * P[1] + inv[2] * P[2], wy = inv[3] * P[0] + inv[4] * P[1] + inv[5] * P[2],
               wz = inv[6] * P[0] + inv[7] * P[1] + inv[8] * P[2];
  if (!(wz > 1e-3)) return false;
  const double vx = fp.focal * wx / wz, vy = fp.focal * wy / wz;  // virtual (sub-pixel) coordinates, kernels.cl:384-400
  *px = (vx + 0.5 * fp.W * fp.A) / fp.A;
  *py = (vy + 0.5 * fp.H * fp.A) / fp.A;
  if (depth) *depth = wz;
  return true;
}

static void visible_rect_of(const float lo[3], const float hi[3], FrameParams &fp);
void visible_rect(const rt_ctx *ctx, FrameParams &fp) { visible_rect_of(ctx->scene_lo, ctx->scene_hi, fp); }

static void visible_rect_of(const float scene_lo[3], const float scene_hi[3], FrameParams &fp) {
  fp.vis_x0 = 0;
  fp.vis_y0 = 0;
  fp.vis_x1 = fp.W;
  fp.vis_y1 = fp.H;
  double inv[9];
  if (!inverse_rotation(fp, inv)) return;
  double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
  for (int c = 0; c < 8; c++) {
    const double P[3] = {(c & 1 ? scene_hi[0] : scene_lo[0]) - (double)fp.cam[0], (c & 2 ? scene_hi[1] : scene_lo[1]) - (double)fp.cam[1],
                         (c & 4 ? scene_hi[2] : scene_lo[2]) - (double)fp.cam[2]};
    if (!(fabs(P[0]) < 1e30 && fabs(P[1]) < 1e30 && fabs(P[2]) < 1e30)) return;
    double px, py;
    if (!project_point(fp, inv, P, &px, &py, nullptr)) return;  // a corner beside or behind the camera: no bound
    x0 = fmin(x0, px);
    x1 = fmax(x1, px);
    y0 = fmin(y0, py);
    y1 = fmax(y1, py);
  }
  if (!(x0 <= x1 && y0 <= y1)) return;
  // pixel x covers sub-pixel rays [x*A, x*A + A - 1] / A: a ray at fractional pixel position q belongs to pixel floor(q)
  const double lo_x = floor(x0) - 2.0, lo_y = floor(y0) - 2.0, hi_x = floor(x1) + 3.0, hi_y = floor(y1) + 3.0;
  fp.vis_x0 = (int)fmax(0.0, fmin((double)fp.W, lo_x));
  fp.vis_y0 = (int)fmax(0.0, fmin((double)fp.H, lo_y));
  fp.vis_x1 = (int)fmax(0.0, fmin((double)fp.W, hi_x));
  fp.vis_y1 = (int)fmax(0.0, fmin((double)fp.H, hi_y));
}

// Mixed launches: the region rendered with four lanes per pixel is the screen rectangle of each sphere, rounded out to
// whole 16x16 tiles of the launch's grid.  In camera space (w = R^-1 (P - cam); the ray through sub-pixel (vx, vy) is
// (vx, vy, f)) the silhouette of a sphere with centre c and radius r spans, along x, the slopes x/z = k of the two planes
// through the camera and the y axis that touch it:  k = (cx cz -+ r sqrt(cx^2 + cz^2 - r^2)) / (cz^2 - r^2)  (likewise y),
// valid while the sphere lies wholly in front of the camera (cz > r); otherwise the whole launch is one rectangle.
// The rectangles only steer performance — either lane mapping renders any tile correctly — so nothing here needs to be
// conservative; two pixels of margin keep the silhouette inside anyway.  Pure arithmetic, redone for every launch: a
// moving camera costs nothing (no tables, no copies, no synchronisation).
// Pixel rectangle {x0, y0, x1, y1} (half-open, two pixels of margin, clipped to the frame) of sphere i for this camera.
// 0: no primary ray can reach the sphere (behind the camera or outside the frame); 1: bounded; 2: not bounded (the camera
// plane cuts the sphere, the camera sits inside it, or the rotation matrix cannot be inverted) — rect = the whole frame.
static int sphere_pixel_rect(const FrameParams &fp, const double inv[9], bool have_inv, int i, int rect[4]) {
  rect[0] = 0;
  rect[1] = 0;
  rect[2] = fp.W;
  rect[3] = fp.H;
  if (!have_inv) return 2;
  const double P[3] = {(double)rt::kSphereCenterR2[i][0] - fp.cam[0], (double)rt::kSphereCenterR2[i][1] - fp.cam[1],
                       (double)rt::kSphereCenterR2[i][2] - fp.cam[2]};
  const double cx = inv[0] * P[0] + inv[1] * P[1] + inv[2] * P[2], cy = inv[3] * P[0] + inv[4] * P[1] + inv[5] * P[2],
               cz = inv[6] * P[0] + inv[7] * P[1] + inv[8] * P[2];
  const double r = sqrt((double)rt::kSphereCenterR2[i][3]) * 1.001;
  if (!(cz == cz) || !(cx == cx) || !(cy == cy)) return 2;
  if (cz <= -r) {  // wholly behind the camera: no primary ray (t >= 0) reaches it
    rect[2] = rect[0];
    rect[3] = rect[1];
    return 0;
  }
  if (!(cz > r * 1.001)) return 2;
  const double den = cz * cz - r * r;
  const double sx = r * sqrt(cx * cx + den), sy = r * sqrt(cy * cy + den);
  const double kx0 = (cx * cz - sx) / den, kx1 = (cx * cz + sx) / den, ky0 = (cy * cz - sy) / den, ky1 = (cy * cz + sy) / den;
  // sub-pixel coordinate v = f k; pixel = (v + W A / 2) / A
  const double px0 = (fp.focal * kx0 + 0.5 * fp.W * fp.A) / fp.A, px1 = (fp.focal * kx1 + 0.5 * fp.W * fp.A) / fp.A;
  const double py0 = (fp.focal * ky0 + 0.5 * fp.H * fp.A) / fp.A, py1 = (fp.focal * ky1 + 0.5 * fp.H * fp.A) / fp.A;
  if (!(fabs(px0) < 1e9 && fabs(px1) < 1e9 && fabs(py0) < 1e9 && fabs(py1) < 1e9)) return 2;
  const double x0 = fmax(0.0, floor(px0) - 2.0), x1 = fmin((double)fp.W, floor(px1) + 3.0);
  const double y0 = fmax(0.0, floor(py0) - 2.0), y1 = fmin((double)fp.H, floor(py1) + 3.0);
  if (!(x0 < x1 && y0 < y1)) {
    rect[2] = rect[0];
    rect[3] = rect[1];
    return 0;
  }
  rect[0] = (int)x0;
  rect[1] = (int)y0;
  rect[2] = (int)x1;
  rect[3] = (int)y1;
  return 1;
}

int sphere_rects(FrameParams &fp) {
  const int gx = (fp.W + kTileW - 1) / kTileW, gy = (fp.rows + kTileH - 1) / kTileH;
  fp.n_rect = 0;
  fp.rect_first[0] = fp.rect_first[1] = 0;
  double inv[9];
  const bool have_inv = inverse_rotation(fp, inv);
  int total = 0;
  for (int i = 0; i < RT_SPHERES; i++) {
    int px[4];
    const int kind = sphere_pixel_rect(fp, inv, have_inv, i, px);
    if (kind == 0) continue;
    if (kind == 2) {  // no bound: the whole launch is one rectangle
      fp.n_rect = 1;
      fp.rect[0][0] = 0;
      fp.rect[0][1] = 0;
      fp.rect[0][2] = gx;
      fp.rect[0][3] = gy;
      fp.rect_first[0] = 0;
      return 4 * gx * gy;
    }
    // clip to this launch's rows and round out to whole tiles of its grid
    const int y0 = std::max(px[1], fp.row0), y1 = std::min(px[3], fp.row0 + fp.rows);
    if (y0 >= y1) continue;
    int *q = fp.rect[fp.n_rect];
    q[0] = px[0] / kTileW;
    q[1] = (y0 - fp.row0) / kTileH;
    q[2] = std::min(gx, (px[2] + kTileW - 1) / kTileW);
    q[3] = std::min(gy, (y1 - fp.row0 + kTileH - 1) / kTileH);
    fp.rect_first[fp.n_rect] = total;
    total += 4 * (q[2] - q[0]) * (q[3] - q[1]);
    fp.n_rect++;
  }
  return total;
}

int mesh_rect(const rt_ctx *ctx, FrameParams &fp) {
  const int gx = (fp.W + kTileW - 1) / kTileW, gy = (fp.rows + kTileH - 1) / kTileH;
  fp.n_rect = 0;
  fp.rect_first[0] = fp.rect_first[1] = 0;
  if (!(ctx->mesh_lo[0] <= ctx->mesh_hi[0])) return 0;  // the tree covers nothing
  FrameParams q = fp;
  visible_rect_of(ctx->mesh_lo, ctx->mesh_hi, q);  // whole frame when a corner of the box lies beside or behind the camera
  const int y0 = std::max(q.vis_y0, fp.row0), y1 = std::min(q.vis_y1, fp.row0 + fp.rows);
  if (q.vis_x0 >= q.vis_x1 || y0 >= y1) return 0;
  int *r = fp.rect[0];
  r[0] = q.vis_x0 / kTileW;
  r[1] = (y0 - fp.row0) / kTileH;
  r[2] = std::min(gx, (q.vis_x1 + kTileW - 1) / kTileW);
  r[3] = std::min(gy, (y1 - fp.row0 + kTileH - 1) / kTileH);
  fp.n_rect = 1;
  return 4 * (r[2] - r[0]) * (r[3] - r[1]);
}

// Per-frame host work of the tuned kernels.  The constants of the exact primary test — b = cam - v0, det[b,e1,e2] and the
// cofactors of det[-d,b,e2], det[-d,e1,b] (rt_fast.cuh) — depend on the camera position only; they are computed here
// with the reference's single-rounded operation sequence (this file is compiled without FMA contraction; IEEE binary32 on
// the host is the same arithmetic as the __fmul_rn / __fadd_rn the kernels would use), followed by their affine form for
// this rotation matrix and focal length.  2 KB for the Cornell box; recomputed and copied only when the camera moves.
cudaError_t prepare_frame(rt_ctx *ctx, FrameParams &fp, cudaStream_t stream) {
  // longest un-normalised primary ray: a corner ray of the frame
  {
    const float hx = 0.5f * (float)fp.W * (float)fp.A, hy = 0.5f * (float)fp.H * (float)fp.A;
    float dmax = 0.0f;
    for (int c = 0; c < 4; c++) {
      const float vx = (c & 1) ? hx : -hx, vy = (c & 2) ? hy : -hy;
      float d2 = 0.0f;
      for (int r = 0; r < 3; r++) {
        const float d = fp.rot[3 * r] * vx + fp.rot[3 * r + 1] * vy + fp.rot[3 * r + 2] * fp.focal;
        d2 += d * d;
      }
      dmax = fmaxf(dmax, sqrtf(d2));
    }
    fp.dmax = dmax * 1.001f;
  }
  double inv[9];
  const bool have_inv = inverse_rotation(fp, inv);
  for (int i = 0; i < RT_SPHERES; i++) sphere_pixel_rect(fp, inv, have_inv, i, fp.sph_px[i]);
  if (ctx->use_bvh || ctx->n == 0) return cudaSuccess;
  float key[14] = {fp.rot[0], fp.rot[1], fp.rot[2], fp.rot[3], fp.rot[4], fp.rot[5], fp.rot[6], fp.rot[7], fp.rot[8],
                   fp.cam[0], fp.cam[1], fp.cam[2], fp.focal, fp.dmax};
  if (ctx->fconst_valid && memcmp(key, ctx->fconst_key, sizeof key) == 0) return cudaSuccess;
  const int n = ctx->n;
  const float4 *ta = ctx->h_tri.data(), *tb = ta + n, *tc = tb + n;
  ctx->h_fconst.resize(6 * (size_t)n);
  float4 *prim = ctx->h_fconst.data(), *aff = prim + 3 * (size_t)n;
  const float *R = fp.rot;
  for (int i = 0; i < n; i++) {
    const float4 A = ta[i], Bq = tb[i], C = tc[i];
    // rt_fast.cuh: primary_constants, operation for operation
    const float bx = fp.cam[0] - A.x, by = fp.cam[1] - A.y, bz = fp.cam[2] - A.z;
    const float p0 = bx * A.w, p1 = by * Bq.w, p2 = bz * C.w;
    const float detA0 = (p0 - p1) + p2;
    const float u0a = by * C.z, u0b = bz * C.y, u1a = bx * C.z, u1b = bz * C.x, u2a = bx * C.y, u2b = by * C.x;
    const float U0 = u0a - u0b, U1 = u1a - u1b, U2 = u2a - u2b;
    const float v0a = Bq.y * bz, v0b = Bq.z * by, v1a = Bq.x * bz, v1b = Bq.z * bx, v2a = Bq.x * by, v2b = Bq.y * bx;
    const float V0 = v0a - v0b, V1 = v1a - v1b, V2 = v2a - v2b;
    const float l1 = fmaxf(fmaxf(fabsf(U0) + fabsf(U1) + fabsf(U2), fabsf(V0) + fabsf(V1) + fabsf(V2)), fabsf(A.w) + fabsf(Bq.w) + fabsf(C.w));
    const float tol = 2e-6f * l1;
    prim[3 * i + 0] = make_float4(A.w, Bq.w, C.w, detA0);
    prim[3 * i + 1] = make_float4(U0, U1, U2, tol);
    prim[3 * i + 2] = make_float4(V0, V1, V2, 3.0f * tol);
    // rt_fast.cuh: primary_affine — R^T (x, -y, z), third component times f; tau = 2 tol dmax
    auto g = [&](float x, float y, float z, float w) {
      y = -y;
      return make_float4(R[0] * x + R[3] * y + R[6] * z, R[1] * x + R[4] * y + R[7] * z, (R[2] * x + R[5] * y + R[8] * z) * fp.focal, w);
    };
    aff[3 * i + 0] = g(A.w, Bq.w, C.w, detA0);
    aff[3 * i + 1] = g(U0, U1, U2, tol * 2.0f * fp.dmax);
    aff[3 * i + 2] = g(V0, V1, V2, 0.0f);
  }
  // the buffer may still be read by launches of the previous camera on another stream
  if (ctx->fconst_valid && ctx->fconst_stream != stream) {
    cudaError_t e = cudaStreamSynchronize(ctx->fconst_stream);
    if (e != cudaSuccess) return e;
  }
  // pageable source: staged before the call returns, and ordered before the launch on the stream
  cudaError_t e = cudaMemcpyAsync(ctx->d_fconst, ctx->h_fconst.data(), sizeof(float4) * ctx->h_fconst.size(), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  memcpy(ctx->fconst_key, key, sizeof key);
  ctx->fconst_valid = true;
  ctx->fconst_stream = stream;
  return cudaSuccess;
}

// centre-out launch order of the tile grid of rows [row0, row0 + rows) (pure host code)
static std::vector<int> build_tile_order(int W, int H, int row0, int grid_x, int n_blocks, int tile_w, int tile_h) {
  std::vector<std::pair<float, int>> key((size_t)n_blocks);
  const float cx = 0.5f * (float)W, cy = 0.5f * (float)H;
  for (int b = 0; b < n_blocks; b++) {
    const int by = b / grid_x, bx = b - by * grid_x;
    const float x = (float)(bx * tile_w + tile_w / 2) - cx, y = (float)(row0 + by * tile_h + tile_h / 2) - cy;
    key[(size_t)b] = std::make_pair(x * x + y * y, b);
  }
  std::sort(key.begin(), key.end());
  std::vector<int> order((size_t)n_blocks);
  for (int b = 0; b < n_blocks; b++) order[(size_t)b] = key[(size_t)b].second;
  shuffle_windows(order, 0, order.size());
  return order;
}

const int *tile_order_for(rt_ctx *ctx, int row0, int rows, int grid_x, int n_blocks, int tile_w, int tile_h) {
  for (const auto &t : ctx->tile_orders)
    if (t.row0 == row0 && t.rows == rows && t.tile_h == tile_h && t.tile_w == tile_w) return t.d_order;
  std::vector<int> order = build_tile_order(ctx->cfg.width, ctx->cfg.height, row0, grid_x, n_blocks, tile_w, tile_h);
  if (grid_x > 0xffff || (n_blocks + grid_x - 1) / std::max(grid_x, 1) > 0x7fff) return nullptr;  // does not pack: row-major order
  for (int &t : order) t = ((t / grid_x) << 16) | (t % grid_x);  // what the kernels read: by << 16 | bx
  int *d = nullptr;
  if (cudaMalloc(&d, sizeof(int) * (size_t)(n_blocks ? n_blocks : 1)) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;  // fall back to row-major order
  }
  if (cudaMemcpy(d, order.data(), sizeof(int) * (size_t)n_blocks, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(d);
    return nullptr;
  }
  if (ctx->tile_orders.size() >= 64) {
    // a caller cycling through many row ranges: drop the old tables once nothing in flight can still read them
    cudaDeviceSynchronize();
    for (auto &t : ctx->tile_orders) cudaFree(t.d_order);
    ctx->tile_orders.clear();
  }
  ctx->tile_orders.push_back(rt_ctx::TileOrder{row0, rows, tile_w, tile_h, d});
  return d;
}

}  // namespace rt

extern "C" {

const char *rt_version(void) { return "uob_rt 0.1 (sm_100a)"; }

void rt_default_config(rt_config *cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof *cfg);
  cfg->width = 1024;   // skeleton.cpp:32-33
  cfg->height = 1024;
  cfg->aa = 2;              // kernels.cl:12-14
  cfg->shadow_samples = 10; // kernels.cl:316
  cfg->max_bounces = 10;    // kernels.cl:343
  cfg->device = 0;
  cfg->row0 = 0;
  cfg->rows = 0;
  cfg->flags = 0;
}

const char *rt_last_error(const rt_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

static int free_ctx(rt_ctx *ctx) {
  if (!ctx) return RT_OK;
  cudaSetDevice(ctx->cfg.device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream && ctx->own_stream != ctx->stream) cudaStreamSynchronize(ctx->own_stream);
  if (ctx->d_frame) cudaFree(ctx->d_frame);
  if (ctx->d_scene) cudaFree(ctx->d_scene);
  if (ctx->d_fconst) cudaFree(ctx->d_fconst);
  if (ctx->d_work) cudaFree(ctx->d_work);
  if (ctx->d_ray_counters) cudaFree(ctx->d_ray_counters);
  if (ctx->d_wait_status) cudaFree(ctx->d_wait_status);
  if (ctx->d_gate_seen) cudaFree(ctx->d_gate_seen);
  for (auto &t : ctx->tile_orders) cudaFree(t.d_order);
  rt::bvh_free(ctx);
  for (int b = 0; b < rt_ctx::kBands; b++) {
    if (ctx->band_stream[b]) cudaStreamDestroy(ctx->band_stream[b]);
    if (ctx->band_done[b]) cudaEventDestroy(ctx->band_done[b]);
  }
  if (ctx->band_start) cudaEventDestroy(ctx->band_start);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
  }
  for (int sl = 0; sl < 2; sl++) {
    if (ctx->slot_kernel_done[sl]) cudaEventDestroy(ctx->slot_kernel_done[sl]);
    if (ctx->slot_copy_done[sl]) cudaEventDestroy(ctx->slot_copy_done[sl]);
  }
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return RT_OK;
}

rt_ctx *rt_create(const rt_config *cfg) {
  g_create_err.clear();
  if (!cfg) {
    g_create_err = "rt_create: cfg is NULL";
    return nullptr;
  }
  if (cfg->width <= 0 || cfg->height <= 0 || cfg->aa < 1 || cfg->aa > 16 || cfg->shadow_samples < 1 || cfg->max_bounces < 0) {
    g_create_err = "rt_create: width/height must be > 0, 1 <= aa <= 16, shadow_samples >= 1, max_bounces >= 0";
    return nullptr;
  }
  // pixel ids are computed in float and converted to int (kernels.cl:380); keep them representable
  if ((long long)cfg->width * cfg->height >= (1ll << 31)) {
    g_create_err = "rt_create: width*height must be < 2^31";
    return nullptr;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("rt_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
    return nullptr;
  }
  if (cfg->device < 0 || cfg->device >= ndev) {
    char buf[128];
    snprintf(buf, sizeof buf, "device index set to %d but only %d devices available", cfg->device, ndev);  // cf. skeleton.cpp:560-565
    g_create_err = buf;
    return nullptr;
  }
  rt_ctx *ctx = new rt_ctx();
  ctx->cfg = *cfg;
  ctx->row0 = cfg->rows > 0 ? cfg->row0 : 0;
  ctx->rows = cfg->rows > 0 ? cfg->rows : cfg->height;
  if (cfg->block_stride > 1 && (cfg->block_phase < 0 || cfg->block_phase >= cfg->block_stride)) {
    g_create_err = "rt_create: block_phase must be in [0, block_stride)";
    delete ctx;
    return nullptr;
  }
  if (ctx->row0 < 0 || ctx->row0 + ctx->rows > cfg->height) {
    g_create_err = "rt_create: row tile outside the frame";
    delete ctx;
    return nullptr;
  }
  auto fail = [&](const char *what, cudaError_t err) -> rt_ctx * {
    g_create_err = std::string("CUDA error during '") + what + "': " + cudaGetErrorString(err);
    free_ctx(ctx);
    return nullptr;
  };
  if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return fail("selecting device", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) return fail("querying device", e);
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("creating stream", e);
  ctx->stream = ctx->own_stream;
  if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return fail("creating event", e);
  if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return fail("creating event", e);
  // earlier bands get the higher stream priority, so that the bands finish roughly in order and each
  // read-back starts while the later bands still render
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  for (int b = 0; b < rt_ctx::kBands; b++) {
    int prio = prio_greatest + b;
    if (prio > prio_least) prio = prio_least;
    if ((e = cudaStreamCreateWithPriority(&ctx->band_stream[b], cudaStreamNonBlocking, prio)) != cudaSuccess) return fail("creating stream", e);
    if ((e = cudaEventCreateWithFlags(&ctx->band_done[b], cudaEventDisableTiming)) != cudaSuccess) return fail("creating event", e);
  }
  if ((e = cudaEventCreateWithFlags(&ctx->band_start, cudaEventDisableTiming)) != cudaSuccess) return fail("creating event", e);
  // the frame, followed by RT_PEER_FLAGS hand-over flags (same allocation, so one IPC handle maps both)
  // two frame slots, each followed by RT_PEER_FLAGS hand-over flags — ONE allocation, so one IPC handle maps everything a
  // peer GPU needs: slot s starts at rt_frame_slot_words() * s words.  Slot 0 is "the" frame; slot 1 is the second frame
  // in flight of rt_render_begin and the other half of a double-buffered multi-GPU hand-over.
  const size_t frame_words = 2 * ((size_t)cfg->width * cfg->height + RT_PEER_FLAGS);
  if ((e = cudaMalloc(&ctx->d_frame, sizeof(uint32_t) * frame_words)) != cudaSuccess) return fail("creating screen buffer", e);
  ctx->d_frame_alt = ctx->d_frame + frame_words / 2;
  if ((e = cudaMalloc(&ctx->d_gate_seen, 2 * sizeof(uint32_t))) != cudaSuccess) return fail("creating gate flag copy", e);
  if ((e = cudaMemsetAsync(ctx->d_gate_seen, 0, 2 * sizeof(uint32_t), ctx->stream)) != cudaSuccess) return fail("clearing gate flag copy", e);
  if ((e = cudaMalloc(&ctx->d_work, 2 * sizeof(unsigned) * rt_ctx::kWorkSlots)) != cudaSuccess) return fail("creating work counters", e);
  if ((e = cudaMemsetAsync(ctx->d_work, 0, 2 * sizeof(unsigned) * rt_ctx::kWorkSlots, ctx->stream)) != cudaSuccess) return fail("clearing work counters", e);
  if ((e = cudaMalloc(&ctx->d_wait_status, sizeof(int))) != cudaSuccess) return fail("creating wait status", e);
  if ((e = cudaMemsetAsync(ctx->d_wait_status, 0, sizeof(int), ctx->stream)) != cudaSuccess) return fail("clearing wait status", e);
  if (cfg->flags & RT_FLAG_COUNT_RAYS) {
    if ((e = cudaMalloc(&ctx->d_ray_counters, 3 * sizeof(unsigned long long))) != cudaSuccess) return fail("creating ray counters", e);
    if ((e = cudaMemsetAsync(ctx->d_ray_counters, 0, 3 * sizeof(unsigned long long), ctx->stream)) != cudaSuccess)
      return fail("clearing ray counters", e);
  }
  if ((e = cudaMemsetAsync(ctx->d_frame, 0, sizeof(uint32_t) * frame_words, ctx->stream)) != cudaSuccess)
    return fail("clearing screen buffer", e);
  return ctx;
}

void rt_destroy(rt_ctx *ctx) { free_ctx(ctx); }

int rt_upload_scene(rt_ctx *ctx, const float *verts, const float *normals, const float *colors, int n) {
  if (!ctx) return RT_ERR_INVALID;
  if (!verts || !normals || !colors || n < 0) {
    ctx->err = "rt_upload_scene: NULL buffer or negative triangle count";
    return RT_ERR_INVALID;
  }
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  // Bounding box of everything a primary ray can hit: the triangles and the two spheres baked into the kernel
  for (int c = 0; c < 3; c++) {
    ctx->scene_lo[c] = 3.0e38f;
    ctx->scene_hi[c] = -3.0e38f;
  }
  for (int i = 0; i < RT_SPHERES; i++) {
    const float r = sqrtf(rt::kSphereCenterR2[i][3]) * 1.0001f;
    for (int c = 0; c < 3; c++) {
      ctx->scene_lo[c] = fminf(ctx->scene_lo[c], rt::kSphereCenterR2[i][c] - r);
      ctx->scene_hi[c] = fmaxf(ctx->scene_hi[c], rt::kSphereCenterR2[i][c] + r);
    }
  }
  for (size_t i = 0; i < 3 * (size_t)n; i++)
    for (int c = 0; c < 3; c++) {
      const float v = verts[4 * i + c];
      if (!(v == v)) {  // NaN vertex: no usable bound
        ctx->scene_lo[c] = -3.0e38f;
        ctx->scene_hi[c] = 3.0e38f;
      } else {
        ctx->scene_lo[c] = fminf(ctx->scene_lo[c], v);
        ctx->scene_hi[c] = fmaxf(ctx->scene_hi[c], v);
      }
    }
  // Shadow casters: everything except material == -1 (kernels.cl:247)
  int n_sh = 0;
  for (int i = 0; i < n; i++) n_sh += (colors[4 * i + 3] != -1.0f);
  // The brute-force path keeps the whole scene in shared memory: the scene arrays, what the fast kernels add per launch
  // for this context's shadow-sample count (parked hits + jitter columns) and the kernels' static shared memory must fit
  // the device's opt-in limit per block (227 KB on sm_100) — otherwise the scene goes through the BVH.
  const size_t smem = rt::brute_smem_bytes(n, n_sh) + rt::fast_extra_smem(ctx->cfg.shadow_samples) + rt::kDrawStaticSmem;
  const bool brute_ok = smem <= ctx->smem_optin;
  bool use_bvh = !brute_ok;
  if (ctx->cfg.flags & RT_FLAG_FORCE_BVH) use_bvh = true;
  if ((ctx->cfg.flags & RT_FLAG_FORCE_BRUTE) && !brute_ok) {
    ctx->err = "rt_upload_scene: scene too large for the brute-force (shared-memory) path";
    return RT_ERR_INVALID;
  }
  if (use_bvh) {
    // Large meshes: GPU-built LBVH over float4 SoA buffers in HBM (rt_bvh.cu)
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "waiting before scene replacement");
    RT_CUDA(ctx, rt::bvh_build(ctx, verts, normals, colors, n), "building the BVH");
    ctx->n = n;
    ctx->n_sh = n_sh;
    ctx->use_bvh = true;
    ctx->have_scene = true;
    return RT_OK;
  }
  rt::bvh_free(ctx);
  // Per-triangle constants, computed with the single-rounded operations the
  // reference kernel performs per ray (kernels.cl:102-104 and the cofactors of
  // det, :31-35).  volatile keeps the host compiler from contracting a*b-c*d.
  std::vector<float4> h(6 * (size_t)n + 8 * (size_t)n_sh);
  float4 *ta = h.data(), *tb = ta + n, *tc = tb + n, *tn = tc + n, *tcol = tn + n;
  float4 *sa = tcol + n, *sb = sa + n_sh, *sc = sb + n_sh, *rec = sc + n_sh, *bnd = rec + 4 * (size_t)n_sh, *tnd = bnd + n_sh;
  int k = 0;
  for (int i = 0; i < n; i++) {
    const float *v0 = verts + 12 * (size_t)i, *v1 = v0 + 4, *v2 = v0 + 8;
    volatile float e1x = v1[0] - v0[0], e1y = v1[1] - v0[1], e1z = v1[2] - v0[2];
    volatile float e2x = v2[0] - v0[0], e2y = v2[1] - v0[1], e2z = v2[2] - v0[2];
    volatile float p0 = e1y * e2z, p1 = e1z * e2y, p2 = e1x * e2z, p3 = e1z * e2x, p4 = e1x * e2y, p5 = e1y * e2x;
    volatile float c0 = p0 - p1, c1 = p2 - p3, c2 = p4 - p5;
    ta[i] = make_float4(v0[0], v0[1], v0[2], c0);
    tb[i] = make_float4(e1x, e1y, e1z, c1);
    tc[i] = make_float4(e2x, e2y, e2z, c2);
    tn[i] = make_float4(normals[4 * i], normals[4 * i + 1], normals[4 * i + 2], 0.0f);
    tcol[i] = make_float4(colors[4 * i], colors[4 * i + 1], colors[4 * i + 2], colors[4 * i + 3]);
    // plane record of the fast policy's bounce rays (rt_fast.cuh: closest_hit_bounce): N = e1 x e2 = (c0, -c1, c2) and
    // v0.N, so that (o - v0).N = o.N - v0.N and the plane tests of a triangle cost one 16-byte load
    tnd[i] = make_float4(c0, -c1, c2, v0[0] * c0 - v0[1] * c1 + v0[2] * c2);
    if (colors[4 * i + 3] != -1.0f) {
      sa[k] = ta[i];
      sb[k] = tb[i];
      sc[k] = tc[i];
      // fast-path record (rt_fast.cuh): bounds on |j.N|, |j|*|e1|, |j|*|e2| over all jitters j of the
      // area light; 0.0445 = kJitterMax
      const float kj = 0.0445f;
      rec[4 * k + 0] = make_float4(v0[0], v0[1], v0[2], kj * sqrtf(c0 * c0 + c1 * c1 + c2 * c2));
      rec[4 * k + 1] = make_float4(c0, c1, c2, 0.0f);
      rec[4 * k + 2] = make_float4(e1x, e1y, e1z, kj * sqrtf(e1x * e1x + e1y * e1y + e1z * e1z));
      rec[4 * k + 3] = make_float4(e2x, e2y, e2z, kj * sqrtf(e2x * e2x + e2y * e2y + e2z * e2z));
      {  // bounding sphere (centroid, farthest vertex) for the warp-level beam cull
        const float cx = (v0[0] + v1[0] + v2[0]) / 3.0f, cy = (v0[1] + v1[1] + v2[1]) / 3.0f, cz = (v0[2] + v1[2] + v2[2]) / 3.0f;
        float r2 = 0.0f;
        const float *vs[3] = {v0, v1, v2};
        for (const float *v : vs) {
          const float dx = v[0] - cx, dy = v[1] - cy, dz = v[2] - cz, d2 = dx * dx + dy * dy + dz * dz;
          r2 = d2 > r2 ? d2 : r2;
        }
        bnd[k] = make_float4(cx, cy, cz, sqrtf(r2) * 1.0001f);
      }
      k++;
    }
  }
  if (ctx->d_scene) {
    RT_CUDA(ctx, cudaDeviceSynchronize(), "waiting before scene replacement");
    RT_CUDA(ctx, cudaFree(ctx->d_scene), "releasing triangle buffer");
    ctx->d_scene = nullptr;
  }
  if (ctx->d_fconst) {
    RT_CUDA(ctx, cudaFree(ctx->d_fconst), "releasing per-frame constants");
    ctx->d_fconst = nullptr;
  }
  ctx->fconst_valid = false;
  ctx->h_tri.assign(h.begin(), h.begin() + 3 * (size_t)n);  // ta | tb | tc: what prepare_frame computes the constants from
  RT_CUDA(ctx, cudaMalloc(&ctx->d_fconst, sizeof(float4) * (size_t)(n ? 6 * n : 1)), "creating per-frame constants");
  RT_CUDA(ctx, cudaMalloc(&ctx->d_scene, sizeof(float4) * (h.size() ? h.size() : 1)), "creating triangle buffer");
  if (!h.empty())
    RT_CUDA(ctx, cudaMemcpyAsync(ctx->d_scene, h.data(), sizeof(float4) * h.size(), cudaMemcpyHostToDevice, ctx->stream),
            "writing triangle buffer data");
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "writing triangle buffer data");
  ctx->n = n;
  ctx->n_sh = n_sh;
  ctx->use_bvh = use_bvh;
  ctx->have_scene = true;
  return RT_OK;
}

// what the reference passes per frame (skeleton.cpp:149-167) plus the context's run-time constants
static void fill_camera(const rt_ctx *ctx, rt::FrameParams &fp, const float rot12[12], const float cam[4], const float light[4], float focal) {
  memset(&fp, 0, sizeof fp);
  fp.W = ctx->cfg.width;
  fp.H = ctx->cfg.height;
  fp.row0 = ctx->row0;
  fp.rows = ctx->rows;
  fp.A = ctx->cfg.aa;
  fp.S = ctx->cfg.shadow_samples;
  fp.B = ctx->cfg.max_bounces;
  fp.focal = focal;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) fp.rot[3 * r + c] = rot12[4 * r + c];  // float4-strided rows (skeleton.cpp:149-151)
  for (int c = 0; c < 3; c++) {
    fp.cam[c] = cam[c];
    fp.light[c] = light[c];
  }
}

static int render_impl(rt_ctx *ctx, const float rot12[12], const float cam[4], const float light[4], float focal,
                       uint32_t *dev_argb, cudaStream_t stream, int band_row0 = -1, int band_rows = 0) {
  if (!ctx) return RT_ERR_INVALID;
  if (!ctx->have_scene) {
    ctx->err = "rt_render: no scene uploaded";
    return RT_ERR_NO_SCENE;
  }
  if (!rot12 || !cam || !light) {
    ctx->err = "rt_render: NULL argument";
    return RT_ERR_INVALID;
  }
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  rt::FrameParams fp;
  fill_camera(ctx, fp, rot12, cam, light, focal);
  fp.row0 = band_row0 >= 0 ? band_row0 : ctx->row0;
  fp.rows = band_row0 >= 0 ? band_rows : ctx->rows;
  rt::visible_rect(ctx, fp);
  // per-frame constants for this camera (a band finds them in place: rt_render prepared them in front of the bands)
  RT_CUDA(ctx, rt::prepare_frame(ctx, fp, stream), "writing per-frame constants");
  // frame gate: the tuned brute-force kernels poll the flag themselves; the other kernels get a wait kernel in front
  fp.gate_flag = nullptr;
  fp.gate_value = 0;
  fp.gate_seen = ctx->d_gate_seen;
  fp.gate_status = ctx->d_wait_status;
  if (ctx->gate_flag) {
    const bool in_kernel = !ctx->use_bvh && !ctx->d_ray_counters &&
                           !((ctx->cfg.flags & RT_FLAG_STRICT_IEEE) && (ctx->cfg.flags & RT_FLAG_REFERENCE_LOOPS));
    if (in_kernel) {
      // device-local copies of the last value seen, one per flag for up to two flags (the two slots of a
      // double-buffered hand-over alternate between two "consumed" flags)
      int e = ctx->gate_flag == ctx->gate_flag_cached[0] ? 0 : (ctx->gate_flag == ctx->gate_flag_cached[1] ? 1 : -1);
      if (e < 0) {
        e = (int)(ctx->gate_victim++ & 1u);
        RT_CUDA(ctx, cudaMemsetAsync(ctx->d_gate_seen + e, 0, sizeof(uint32_t), stream), "clearing gate flag copy");
        ctx->gate_flag_cached[e] = ctx->gate_flag;
      }
      fp.gate_seen = ctx->d_gate_seen + e;
      fp.gate_flag = ctx->gate_flag;
      fp.gate_value = ctx->gate_value;
    } else {
      RT_CUDA(ctx, rt::launch_peer_wait(ctx->gate_flag, 1, ctx->gate_value, ctx->d_wait_status, stream), "enqueueing frame gate");
      ctx->launches++;
    }
    ctx->peer_waits = true;
    ctx->gate_flag = nullptr;
  }
  // rt_signal_after_frame: the tuned kernels add to the counter themselves once the last block is done; the other kernels
  // get a one-thread kernel behind them
  uint32_t *signal_after = ctx->signal_flag;
  ctx->signal_flag = nullptr;
  fp.signal_flag = nullptr;
  if (signal_after && !ctx->use_bvh && !ctx->d_ray_counters &&
      !((ctx->cfg.flags & RT_FLAG_STRICT_IEEE) && (ctx->cfg.flags & RT_FLAG_REFERENCE_LOOPS))) {
    fp.signal_flag = signal_after;
    signal_after = nullptr;
  }
  fp.out = dev_argb ? dev_argb : ctx->d_frame;
  fp.n_out = 0;
  fp.strip_rows = 1;
  if (ctx->n_strip_targets > 1) {
    if (ctx->use_bvh || ctx->d_ray_counters || ((ctx->cfg.flags & RT_FLAG_STRICT_IEEE) && (ctx->cfg.flags & RT_FLAG_REFERENCE_LOOPS))) {
      ctx->err = "rt_set_strip_targets: only the tuned brute-force kernels deal rows out over several frame buffers";
      return RT_ERR_INVALID;
    }
    fp.n_out = ctx->n_strip_targets;
    fp.strip_rows = ctx->strip_rows;
    for (int i = 0; i < 8; i++) fp.outs[i] = ctx->strip_targets[i];
  }
  fp.ray_counters = ctx->d_ray_counters;
  if (ctx->d_ray_counters && band_row0 < 0)
    RT_CUDA(ctx, cudaMemsetAsync(ctx->d_ray_counters, 0, 3 * sizeof(unsigned long long), stream), "clearing ray counters");
  const bool whole = band_row0 < 0;
  if (whole) RT_CUDA(ctx, cudaEventRecord(ctx->ev0, stream), "recording start event");
  RT_CUDA(ctx, ctx->use_bvh ? rt::launch_draw_bvh(ctx, fp, stream) : rt::launch_draw_brute(ctx, fp, stream), "enqueueing draw kernel");
  if (signal_after) {
    RT_CUDA(ctx, rt::launch_peer_add(signal_after, stream), "enqueueing peer signal");
    ctx->launches++;
  }
  if (whole) {
    RT_CUDA(ctx, cudaEventRecord(ctx->ev1, stream), "recording stop event");
    ctx->timed = true;
  }
  return RT_OK;
}

int rt_render_device(rt_ctx *ctx, const float rot12[12], const float cam[4], const float light[4], float focal,
                     uint32_t *dev_argb, void *stream) {
  if (!ctx) return RT_ERR_INVALID;
  return render_impl(ctx, rot12, cam, light, focal, dev_argb, stream ? (cudaStream_t)stream : ctx->stream);
}

int rt_render(rt_ctx *ctx, const float rot12[12], const float cam[4], const float light[4], float focal, uint32_t *host_argb) {
  if (!ctx) return RT_ERR_INVALID;
  if (!host_argb) {
    ctx->err = "rt_render: host_argb is NULL";
    return RT_ERR_INVALID;
  }
  const int W = ctx->cfg.width;
  // Band height: a multiple of the block height.  Small tiles and block-interleaved contexts (whose frame is completed
  // by other GPUs) render in one piece.  The exposed part of the read-back is the last band's copy, so more bands hide
  // more of it — until the bands are such small launches that the kernels lose more than the copy gains.  Measured on
  // B200 (scripts/gpu_e2e_variants.sh, ms per frame for 2 / 3 / 4 / 5 / 6 / 8 bands): 1080p cfg2 0.303 / 0.276 / 0.267 /
  // 0.260 / 0.261 / 0.254; 4K cfg3 - / - / 2.302 / 2.277 / 2.345 / 2.35.
#ifdef RT_BANDS
  const int n_bands = RT_BANDS < rt_ctx::kBands ? RT_BANDS : rt_ctx::kBands;
#else
  const int n_bands = (size_t)W * ctx->rows <= (size_t)3 << 20 ? 8 : 5;
#endif
  int band_rows = ((ctx->rows + n_bands - 1) / n_bands + 15) / 16 * 16;
  if (ctx->rows < 256 || ctx->cfg.block_stride > 1 || ctx->d_ray_counters) band_rows = ctx->rows;
  if (band_rows >= ctx->rows) {
    int rc = render_impl(ctx, rot12, cam, light, focal, nullptr, ctx->stream);
    if (rc != RT_OK) return rc;
    const size_t off = (size_t)ctx->row0 * W, cnt = (size_t)ctx->rows * W;
    RT_CUDA(ctx, cudaMemcpyAsync(host_argb, ctx->d_frame + off, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost, ctx->stream),
            "reading screen buffer data");
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "reading screen buffer data");
    return RT_OK;
  }
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  if (ctx->have_scene && rot12 && cam && light) {  // (render_impl reports the errors)
    // the per-frame constants go in on the context's stream, in front of the bands that read them
    rt::FrameParams fp;
    fill_camera(ctx, fp, rot12, cam, light, focal);
    RT_CUDA(ctx, rt::prepare_frame(ctx, fp, ctx->stream), "writing per-frame constants");
  }
  RT_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream), "recording start event");
  RT_CUDA(ctx, cudaEventRecord(ctx->band_start, ctx->stream), "recording start event");
  int nb = 0;
  for (int r = 0; r < ctx->rows; r += band_rows, nb++) {
    cudaStream_t bs = ctx->band_stream[nb];
    const int rows = (r + band_rows <= ctx->rows) ? band_rows : ctx->rows - r;
    RT_CUDA(ctx, cudaStreamWaitEvent(bs, ctx->band_start, 0), "ordering a band after the stream");
    int rc = render_impl(ctx, rot12, cam, light, focal, nullptr, bs, ctx->row0 + r, rows);
    if (rc == RT_OK) {
      const size_t off = (size_t)(ctx->row0 + r) * W;
      cudaError_t e = cudaMemcpyAsync(host_argb + (size_t)r * W, ctx->d_frame + off, sizeof(uint32_t) * (size_t)rows * W, cudaMemcpyDeviceToHost, bs);
      if (e == cudaSuccess) e = cudaEventRecord(ctx->band_done[nb], bs);
      if (e != cudaSuccess) {
        ctx->err = std::string("CUDA error during 'reading screen buffer data': ") + cudaGetErrorString(e);
        rc = RT_ERR_CUDA;
      }
    }
    if (rc != RT_OK) {
      // leave nothing in flight that still writes into the caller's buffer: join the bands already started
      const std::string why = ctx->err;
      for (int b = 0; b <= nb && b < rt_ctx::kBands; b++) cudaStreamSynchronize(ctx->band_stream[b]);
      cudaGetLastError();
      ctx->err = why;
      return rc;
    }
  }
  for (int b = 0; b < nb; b++) RT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->band_done[b], 0), "joining the bands");
  RT_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream), "recording stop event");
  ctx->timed = true;
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "reading screen buffer data");
  return RT_OK;
}

void *rt_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void rt_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int rt_render_begin(rt_ctx *ctx, const float rot12[12], const float cam[4], const float light[4], float focal, uint32_t *host_argb) {
  if (!ctx) return RT_ERR_INVALID;
  if (!host_argb) {
    ctx->err = "rt_render_begin: host_argb is NULL";
    return RT_ERR_INVALID;
  }
  if (ctx->frames_begun - ctx->frames_ended >= 2) {
    ctx->err = "rt_render_begin: two frames are already in flight; call rt_render_end first";
    return RT_ERR_INVALID;
  }
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  if (!ctx->copy_stream) {  // first use: copy stream, events
    RT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking), "creating copy stream");
    for (int sl = 0; sl < 2; sl++) {
      RT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->slot_kernel_done[sl], cudaEventDisableTiming), "creating event");
      RT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->slot_copy_done[sl], cudaEventDisableTiming), "creating event");
    }
  }
  const int sl = (int)(ctx->frames_begun & 1);
  uint32_t *dst = sl ? ctx->d_frame_alt : ctx->d_frame;
  // the kernel may overwrite this slot only after its previous read-back (two frames ago) finished
  if (ctx->slot_busy[sl]) RT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->slot_copy_done[sl], 0), "ordering the frame slot");
  int rc = render_impl(ctx, rot12, cam, light, focal, dst, ctx->stream);
  if (rc != RT_OK) return rc;
  RT_CUDA(ctx, cudaEventRecord(ctx->slot_kernel_done[sl], ctx->stream), "recording kernel completion");
  RT_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->slot_kernel_done[sl], 0), "ordering the read-back");
  const size_t off = (size_t)ctx->row0 * ctx->cfg.width, cnt = (size_t)ctx->rows * ctx->cfg.width;
  RT_CUDA(ctx, cudaMemcpyAsync(host_argb, dst + off, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost, ctx->copy_stream),
          "reading screen buffer data");
  RT_CUDA(ctx, cudaEventRecord(ctx->slot_copy_done[sl], ctx->copy_stream), "recording read-back completion");
  ctx->slot_busy[sl] = true;
  ctx->frames_begun++;
  return RT_OK;
}

int rt_render_end(rt_ctx *ctx) {
  if (!ctx) return RT_ERR_INVALID;
  if (ctx->frames_begun == ctx->frames_ended) {
    ctx->err = "rt_render_end: no frame in flight";
    return RT_ERR_INVALID;
  }
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  const int sl = (int)(ctx->frames_ended & 1);
  RT_CUDA(ctx, cudaEventSynchronize(ctx->slot_copy_done[sl]), "reading screen buffer data");
  ctx->frames_ended++;
  return RT_OK;
}

int rt_read_frame(rt_ctx *ctx, uint32_t *host_argb) { return rt_read_frame_slot(ctx, 0, host_argb); }

int rt_read_frame_slot(rt_ctx *ctx, int slot, uint32_t *host_argb) {
  if (!ctx || !host_argb || slot < 0 || slot > 1) return RT_ERR_INVALID;
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  RT_CUDA(ctx, cudaMemcpyAsync(host_argb, ctx->d_frame + rt_frame_slot_words(ctx) * (size_t)slot, sizeof(uint32_t) * (size_t)ctx->cfg.width * ctx->cfg.height,
                               cudaMemcpyDeviceToHost, ctx->stream), "reading screen buffer data");
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "reading screen buffer data");
  return RT_OK;
}

int rt_enable_peer(rt_ctx *ctx, int peer_device) {
  if (!ctx) return RT_ERR_INVALID;
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  if (peer_device == ctx->cfg.device) return RT_OK;
  int can = 0;
  RT_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->cfg.device, peer_device), "querying peer access");
  if (!can) {
    ctx->err = "rt_enable_peer: device cannot access the peer";
    return RT_ERR_CUDA;
  }
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return RT_OK;
  }
  RT_CUDA(ctx, e, "enabling peer access");
  return RT_OK;
}

int rt_ipc_export_frame(rt_ctx *ctx, void *handle64) {
  if (!ctx || !handle64) return RT_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  cudaIpcMemHandle_t h;
  RT_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->d_frame), "exporting the frame buffer");
  memcpy(handle64, &h, sizeof h);
  return RT_OK;
}

int rt_ipc_open_frame(rt_ctx *ctx, const void *handle64, uint32_t **dev_argb) {
  if (!ctx || !handle64 || !dev_argb) return RT_ERR_INVALID;
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof h);
  void *p = nullptr;
  RT_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "mapping the peer frame buffer");
  *dev_argb = static_cast<uint32_t *>(p);
  return RT_OK;
}

int rt_ipc_close_frame(rt_ctx *ctx, uint32_t *dev_argb) {
  if (!ctx || !dev_argb) return RT_ERR_INVALID;
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "waiting for the stream");
  RT_CUDA(ctx, cudaIpcCloseMemHandle(dev_argb), "unmapping the peer frame buffer");
  return RT_OK;
}

int rt_get_ray_counts(rt_ctx *ctx, uint64_t counts[3]) {
  if (!ctx || !counts) return RT_ERR_INVALID;
  if (!ctx->d_ray_counters) {
    ctx->err = "rt_get_ray_counts: the context was not created with RT_FLAG_COUNT_RAYS";
    return RT_ERR_INVALID;
  }
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  unsigned long long h[3];
  RT_CUDA(ctx, cudaMemcpyAsync(h, ctx->d_ray_counters, sizeof h, cudaMemcpyDeviceToHost, ctx->stream), "reading ray counters");
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "reading ray counters");
  for (int i = 0; i < 3; i++) counts[i] = h[i];
  return RT_OK;
}

int rt_set_stream(rt_ctx *ctx, void *stream) {
  if (!ctx) return RT_ERR_INVALID;
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "waiting for the stream");
  ctx->stream = stream ? static_cast<cudaStream_t>(stream) : ctx->own_stream;
  return RT_OK;
}

int rt_synchronize(rt_ctx *ctx) {
  if (!ctx) return RT_ERR_INVALID;
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "waiting for the stream");
  if (ctx->peer_waits) {  // did a rt_peer_wait give up?
    int status = 0;
    RT_CUDA(ctx, cudaMemcpy(&status, ctx->d_wait_status, sizeof status, cudaMemcpyDeviceToHost), "reading wait status");
    if (status) {
      ctx->err = "rt_peer_wait: timed out waiting for a peer GPU's flag";
      return RT_ERR_CUDA;
    }
  }
  return RT_OK;
}

uint32_t *rt_peer_flags(rt_ctx *ctx) { return ctx ? ctx->d_frame + (size_t)ctx->cfg.width * ctx->cfg.height : nullptr; }

int rt_peer_signal(rt_ctx *ctx, uint32_t *dev_flag, uint32_t value, void *stream) {
  if (!ctx || !dev_flag) return RT_ERR_INVALID;
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  RT_CUDA(ctx, rt::launch_peer_signal(dev_flag, value, stream ? (cudaStream_t)stream : ctx->stream), "enqueueing peer signal");
  ctx->launches++;
  return RT_OK;
}

int rt_peer_wait(rt_ctx *ctx, const uint32_t *dev_flags, int n, uint32_t value, void *stream) {
  if (!ctx || !dev_flags || n < 1 || n > 32) return RT_ERR_INVALID;
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  RT_CUDA(ctx, rt::launch_peer_wait(dev_flags, n, value, ctx->d_wait_status, stream ? (cudaStream_t)stream : ctx->stream),
          "enqueueing peer wait");
  ctx->launches++;
  ctx->peer_waits = true;
  return RT_OK;
}

static void debug_frame_params(const rt_config *cfg, const float rot12[12], const float cam[4], float focal, rt::FrameParams &fp) {
  memset(&fp, 0, sizeof fp);
  fp.W = cfg->width;
  fp.H = cfg->height;
  fp.row0 = cfg->row0;
  fp.rows = cfg->rows > 0 ? cfg->rows : cfg->height - cfg->row0;
  fp.A = cfg->aa;
  fp.S = cfg->shadow_samples;
  fp.B = cfg->max_bounces;
  fp.focal = focal;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) fp.rot[3 * r + c] = rot12[4 * r + c];
  for (int c = 0; c < 3; c++) fp.cam[c] = cam[c];
}

int rt_debug_visible_rect(const rt_config *cfg, const float lo[3], const float hi[3], const float rot12[12], const float cam[4], float focal,
                          int rect[4]) {
  if (!cfg || !lo || !hi || !rot12 || !cam || !rect) return RT_ERR_INVALID;
  rt::FrameParams fp;
  debug_frame_params(cfg, rot12, cam, focal, fp);
  rt::visible_rect_of(lo, hi, fp);
  rect[0] = fp.vis_x0;
  rect[1] = fp.vis_y0;
  rect[2] = fp.vis_x1;
  rect[3] = fp.vis_y1;
  return RT_OK;
}

int rt_debug_tile_lists(const rt_config *cfg, const float rot12[12], const float cam[4], float focal, int *tiles, int capacity, int *n_light,
                        int *n_split) {
  if (!cfg || !rot12 || !cam || !n_light || !n_split || capacity < 0 || (capacity > 0 && !tiles)) return RT_ERR_INVALID;
  rt::FrameParams fp;
  debug_frame_params(cfg, rot12, cam, focal, fp);
  // what a mixed launch over these rows renders, block by block (rt_launch.cuh: tile_of_block)
  const int gx = (fp.W + rt::kTileW - 1) / rt::kTileW, gy = (fp.rows + rt::kTileH - 1) / rt::kTileH, sgx = (fp.W + rt::kSplitTileW - 1) / rt::kSplitTileW;
  const int n_sub = rt::sphere_rects(fp);
  auto in_rect = [&](int r, int bx, int by) { return bx >= fp.rect[r][0] && bx < fp.rect[r][2] && by >= fp.rect[r][1] && by < fp.rect[r][3]; };
  std::vector<int> light, split;
  for (int t : rt::build_tile_order(fp.W, fp.H, fp.row0, gx, gx * gy, rt::kTileW, rt::kTileH)) {
    const int by = t / gx, bx = t - by * gx;
    bool split_it = false;
    for (int r = 0; r < fp.n_rect; r++) split_it |= in_rect(r, bx, by);
    if (!split_it) light.push_back(t);
  }
  for (int gb = 0; gb < n_sub; gb++) {
    const int r = (fp.n_rect > 1 && gb >= fp.rect_first[1]) ? 1 : 0;
    const int local = gb - fp.rect_first[r], sw = 2 * (fp.rect[r][2] - fp.rect[r][0]);
    const int sy = local / sw, sx = local - sy * sw, bx = 2 * fp.rect[r][0] + sx, by = 2 * fp.rect[r][1] + sy;
    if (r == 1 && in_rect(0, bx >> 1, by >> 1)) continue;
    if (bx * rt::kSplitTileW >= fp.W || by * rt::kSplitTileH >= fp.rows) continue;  // the block finds no pixel of the frame and exits
    split.push_back(by * sgx + bx);
  }
  *n_light = (int)light.size();
  *n_split = (int)split.size();
  if ((size_t)capacity < light.size() + split.size()) return RT_ERR_INVALID;
  std::copy(light.begin(), light.end(), tiles);
  std::copy(split.begin(), split.end(), tiles + light.size());
  return RT_OK;
}

int rt_gate_next_frame(rt_ctx *ctx, const uint32_t *dev_flag, uint32_t value) {
  if (!ctx || !dev_flag) return RT_ERR_INVALID;
  ctx->gate_flag = dev_flag;
  ctx->gate_value = value;
  return RT_OK;
}

uint32_t *rt_device_frame(rt_ctx *ctx) { return ctx ? ctx->d_frame : nullptr; }

size_t rt_frame_slot_words(const rt_ctx *ctx) { return ctx ? (size_t)ctx->cfg.width * ctx->cfg.height + RT_PEER_FLAGS : 0; }

int rt_set_strip_targets(rt_ctx *ctx, uint32_t *const *dev_frames, int n, int strip_rows) {
  if (!ctx || n < 0 || n > 8 || (n > 1 && (!dev_frames || strip_rows < 16 || strip_rows % 16 != 0))) return RT_ERR_INVALID;
  ctx->n_strip_targets = n > 1 ? n : 0;
  ctx->strip_rows = strip_rows;
  for (int i = 0; i < 8; i++) ctx->strip_targets[i] = (n > 1 && i < n) ? dev_frames[i] : nullptr;
  return RT_OK;
}

// Strips phase, phase + n, ... of frame slot `slot` (rows [s*strip_rows, (s+1)*strip_rows) of strip s) into the same rows of
// host_argb: one 2-D copy for the full strips, one more for a ragged last strip.  Asynchronous on `stream`.
int rt_read_strips(rt_ctx *ctx, int slot, int strip_rows, int n, int phase, uint32_t *host_argb, void *stream) {
  if (!ctx || !host_argb || slot < 0 || slot > 1 || n < 1 || phase < 0 || phase >= n || strip_rows < 1) return RT_ERR_INVALID;
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  const size_t W = (size_t)ctx->cfg.width, H = (size_t)ctx->cfg.height;
  const uint32_t *src = ctx->d_frame + rt_frame_slot_words(ctx) * (size_t)slot;
  const size_t strip_words = (size_t)strip_rows * W, pitch = sizeof(uint32_t) * strip_words * (size_t)n;
  const size_t first = (size_t)phase * strip_words;         // word offset of this rank's first strip
  const size_t total_strips = (H + strip_rows - 1) / strip_rows;
  if ((size_t)phase >= total_strips) return RT_OK;
  const size_t mine = (total_strips - phase + n - 1) / n;   // strips phase, phase+n, ... < total_strips
  const size_t last_strip = (size_t)phase + (mine - 1) * n; // index of my last strip
  const size_t last_rows = std::min((size_t)strip_rows, H - last_strip * strip_rows);
  const size_t full = last_rows == (size_t)strip_rows ? mine : mine - 1;
  if (full > 0)
    RT_CUDA(ctx, cudaMemcpy2DAsync(host_argb + first, pitch, src + first, pitch, sizeof(uint32_t) * strip_words, full, cudaMemcpyDeviceToHost, st),
            "reading screen buffer strips");
  if (full < mine) {
    const size_t off = last_strip * strip_words;
    RT_CUDA(ctx, cudaMemcpyAsync(host_argb + off, src + off, sizeof(uint32_t) * last_rows * W, cudaMemcpyDeviceToHost, st), "reading screen buffer strips");
  }
  return RT_OK;
}

// One parallel-egress frame in one call (what a host loop does per frame on every rank, without six trips through the
// binding): deal the rows of frame slot `slot` out over dev_frames (rt_set_strip_targets), draw, report the delivery to the
// other owners, wait until the other ranks have delivered into this rank's strips, copy them into host_argb.  Asynchronous
// on the context's stream; the caller synchronises (rt_synchronize) and then meets the other ranks.
int rt_render_strips(rt_ctx *ctx, const float rot12[12], const float cam[4], const float light[4], float focal, uint32_t *const *dev_frames,
                     int n, int rank, int strip_rows, int slot, uint32_t deliveries_expected, uint32_t *host_argb) {
  if (!ctx || !dev_frames || n < 2 || n > 8 || rank < 0 || rank >= n || slot < 0 || slot > 1 || !host_argb) return RT_ERR_INVALID;
  const size_t slot_words = rt_frame_slot_words(ctx), frame_words = (size_t)ctx->cfg.width * ctx->cfg.height;
  uint32_t *frames[8], *counters[8];
  int nc = 0;
  for (int k = 0; k < n; k++) {
    frames[k] = dev_frames[k] + slot_words * (size_t)slot;
    if (k != rank) counters[nc++] = frames[k] + frame_words;  // word 0 behind the pixels: deliveries counted
  }
  int rc = rt_set_strip_targets(ctx, frames, n, strip_rows);
  if (rc != RT_OK) return rc;
  rc = rt_render_device(ctx, rot12, cam, light, focal, frames[rank], nullptr);
  if (rc == RT_OK) rc = rt_peer_add(ctx, counters, nc, nullptr);
  if (rc == RT_OK) rc = rt_stream_wait_geq(ctx, frames[rank] + frame_words, deliveries_expected, nullptr);
  if (rc == RT_OK) rc = rt_read_strips(ctx, slot, strip_rows, n, rank, host_argb, nullptr);
  return rc;
}

int rt_host_register(void *p, size_t bytes) {
  if (!p || !bytes) return RT_ERR_INVALID;
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return RT_OK;
  }
  return e == cudaSuccess ? RT_OK : RT_ERR_CUDA;
}

int rt_host_unregister(void *p) {
  if (!p) return RT_ERR_INVALID;
  cudaHostUnregister(p);
  cudaGetLastError();
  return RT_OK;
}

int rt_peer_add(rt_ctx *ctx, uint32_t *const *dev_counters, int n, void *stream) {
  if (!ctx || !dev_counters || n < 1 || n > 8) return RT_ERR_INVALID;
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  RT_CUDA(ctx, rt::launch_peer_add_many(dev_counters, n, stream ? (cudaStream_t)stream : ctx->stream), "enqueueing peer signal");
  ctx->launches++;
  return RT_OK;
}

int rt_signal_after_frame(rt_ctx *ctx, uint32_t *dev_counter) {
  if (!ctx || !dev_counter) return RT_ERR_INVALID;
  ctx->signal_flag = dev_counter;
  return RT_OK;
}

// Stream-ordered wait / write on a 32-bit word without a kernel: the driver's stream memory operations
// (cuStreamWaitValue32 / cuStreamWriteValue32), resolved through the runtime so that the library does not link libcuda.
typedef int (*stream_memop_fn)(void *stream, unsigned long long addr, unsigned value, unsigned flags);
static stream_memop_fn driver_memop(const char *name) {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return (stream_memop_fn)fn;
}

int rt_stream_wait_geq(rt_ctx *ctx, const uint32_t *dev_word, uint32_t value, void *stream) {
  if (!ctx || !dev_word) return RT_ERR_INVALID;
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  static stream_memop_fn wait32 = driver_memop("cuStreamWaitValue32");
  if (wait32 && !(ctx->cfg.flags & RT_FLAG_NO_STREAM_MEMOPS)) {
    const int rc = wait32(st, (unsigned long long)(uintptr_t)dev_word, value, 0x0 /* CU_STREAM_WAIT_VALUE_GEQ */);
    if (rc == 0) return RT_OK;
  }
  RT_CUDA(ctx, rt::launch_peer_wait(dev_word, 1, value, ctx->d_wait_status, st), "enqueueing peer wait");
  ctx->launches++;
  ctx->peer_waits = true;
  return RT_OK;
}

int rt_stream_write(rt_ctx *ctx, uint32_t *dev_word, uint32_t value, void *stream) {
  if (!ctx || !dev_word) return RT_ERR_INVALID;
  RT_CUDA(ctx, cudaSetDevice(ctx->cfg.device), "selecting device");
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  static stream_memop_fn write32 = driver_memop("cuStreamWriteValue32");
  if (write32 && !(ctx->cfg.flags & RT_FLAG_NO_STREAM_MEMOPS)) {
    const int rc = write32(st, (unsigned long long)(uintptr_t)dev_word, value, 0x0 /* CU_STREAM_WRITE_VALUE_DEFAULT: memory barrier in front */);
    if (rc == 0) return RT_OK;
  }
  RT_CUDA(ctx, rt::launch_peer_signal(dev_word, value, st), "enqueueing peer signal");
  ctx->launches++;
  return RT_OK;
}

float rt_last_kernel_ms(rt_ctx *ctx) {
  if (!ctx || !ctx->timed) return -1.0f;
  if (cudaEventSynchronize(ctx->ev1) != cudaSuccess) return -1.0f;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess) return -1.0f;
  return ms;
}

uint64_t rt_kernel_launches(const rt_ctx *ctx) { return ctx ? ctx->launches : 0; }

const char *rt_scene_mode(const rt_ctx *ctx) { return (ctx && ctx->use_bvh) ? "bvh" : "brute"; }

const char *rt_last_kernel_name(const rt_ctx *ctx) { return ctx ? ctx->last_kernel.c_str() : ""; }

}  // extern "C"
