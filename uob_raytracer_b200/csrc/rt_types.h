// rt_types.h — plain structs shared by the C-ABI layer and the kernels.
#pragma once
#include <stdint.h>

namespace rt {

// Per-launch arguments of `draw`: what the reference passes as kernel args 4-8
// (kernels.cl:368-371; skeleton.cpp:160-167) plus the run-time versions of its
// compile-time constants and the row tile of this launch.
struct FrameParams {
  int W, H;        // whole frame
  int row0, rows;  // rows rendered by this launch
  // 16 x kTileH pixel blocks of those rows are numbered row-major; this launch renders the blocks
  // gb = b*blk_stride + blk_phase (b = blockIdx.x): 1/0 = all of them, N/g = rank g of an N-way interleave
  int blk_stride, blk_phase, grid_x, n_blocks;
  // optional launch order: tile_order[k] = row-major block number of the k-th block to start
  // (expensive centre tiles first, so that no long-running block is left for the tail)
  const int *tile_order;
  // mixed launches (rt_draw_fast.cu): the first n_split blocks of the grid render 8x8-pixel sub-tiles with four lanes per
  // pixel; the rest are ordinary blocks.  The split region is up to two rectangles of 16x16 tiles (the screen rectangles of
  // the two spheres, rt_api.cu: sphere_rects): rect[r] = {tx0, ty0, tx1, ty1}, half-open, in tiles of this launch's grid
  // (ty counted from row0).  The sub-tiles of rect r are numbered row-major from rect_first[r]; an ordinary block whose
  // tile lies in a rectangle exits at once, and so does a sub-tile of rect 1 that rect 0 already covers.
  int n_split, n_rect;
  int rect[2][4];
  int rect_first[2];
  // persistent launches of the tuned kernels (rt_draw_fast.cu): blocks take tiles from work_counter[0] until n_items are
  // handed out (n_split_items = n_split of them are the four-lanes-per-pixel sub-tiles of a mixed launch and come first).
  // work_counter[1] counts finished blocks: the last one re-arms both words.
  int n_items, n_split_items;
  unsigned *work_counter;
  // per-frame camera data computed on the host (rt_api.cu: prepare_frame): the longest un-normalised primary ray direction
  // of the frame (x 1.001), and the pixel rectangle {x0, y0, x1, y1} (half-open, two pixels of margin) that contains every
  // primary ray able to reach sphere i — empty if none can, the whole frame if the projection is not bounded
  float dmax;
  int sph_px[2][4];
  // pixel rectangle [vis_x0, vis_x1) x [vis_y0, vis_y1) outside which no primary ray can hit anything (projection of the
  // scene's bounding box, rt_api.cu): tiles outside it are black without looking at the scene
  int vis_x0, vis_y0, vis_x1, vis_y1;
  // rt_gate_next_frame: no pixel of this launch is stored before *gate_flag >= gate_value (another GPU's "I have consumed
  // the previous frame" flag).  gate_seen is a device-local copy of the last value observed, gate_status the time-out flag.
  const uint32_t *gate_flag;
  uint32_t gate_value;
  // rt_signal_after_frame: when the last block of this launch is done and every pixel it stored is visible system-wide,
  // *signal_flag += 1 (a counter in the frame owner's memory: "one more rank has delivered this frame")
  uint32_t *signal_flag;
  uint32_t *gate_seen;
  int *gate_status;
  int A, S, B;     // AA edge, shadow samples, max bounces
  float focal;
  float rot[9];    // rows r0, r1, r2 (skeleton.cpp:149-151)
  float cam[3], light[3];
  uint32_t *out;   // whole-frame ARGB buffer (may be a peer pointer)
  // Parallel egress (rt_set_strip_targets): frame rows are dealt out in strips of strip_rows rows over n_out owners — row y
  // belongs to outs[(y / strip_rows) % n_out], the whole-frame buffer of that owner (its own, or a peer's over NVLink) — so
  // that every GPU ends up with the strips it will copy to the host over its own PCIe link.  n_out <= 1: everything to out.
  int n_out, strip_rows;
  uint32_t *outs[8];
  // RT_FLAG_COUNT_RAYS: {primary, shadow, bounce} rays traced, SURVEY.md §8d definition (else NULL)
  unsigned long long *ray_counters;
};


}  // namespace rt
