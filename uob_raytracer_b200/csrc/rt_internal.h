// rt_internal.h — context object behind the C ABI (include/uob_rt.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/uob_rt.h"

namespace rt {
struct FrameParams;
}

struct rt_ctx {
  rt_config cfg{};
  int row0 = 0, rows = 0;  // resolved tile
  cudaStream_t stream = nullptr;      // the stream in use
  cudaStream_t own_stream = nullptr;  // created by rt_create
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // rt_render overlaps the read-back with the kernel: the tile is rendered in row bands (at most kBands) on
  // their own streams and each band is copied to the host as soon as it is done.  How many: rt_api.cu, band_count
  // (-DRT_BANDS=n forces a count, for A/B runs).
  static constexpr int kBands = 8;
  cudaStream_t band_stream[kBands] = {};
  cudaEvent_t band_done[kBands] = {};
  cudaEvent_t band_start = nullptr;
  // rt_render_begin / rt_render_end: two frame slots, a copy stream, per-slot events
  uint32_t *d_frame_alt = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t slot_kernel_done[2] = {}, slot_copy_done[2] = {};
  bool slot_busy[2] = {false, false};
  unsigned long long frames_begun = 0, frames_ended = 0;
  bool timed = false;
  uint32_t *d_frame = nullptr;  // whole frame, W*H
  // Brute-force scene: one float4 buffer, [ta|tb|tc|tn|tcol] x n, [sa|sb|sc] x n_sh (generic kernel),
  // then 4 float4 per shadow caster and n_sh bounding spheres (fast kernel, rt_fast.cuh)
  float4 *d_scene = nullptr;
  int n = 0, n_sh = 0;
  // Per-frame triangle constants of the tuned kernels (6n float4: the exact primary test's constants, then their affine
  // form — rt_fast.cuh), computed on the host whenever the camera changes and copied in stream order in front of the launch
  float4 *d_fconst = nullptr;
  std::vector<float4> h_tri;     // ta | tb | tc of the uploaded scene (host copy the constants are computed from)
  std::vector<float4> h_fconst;
  float fconst_key[14] = {0};
  bool fconst_valid = false;
  cudaStream_t fconst_stream = nullptr;  // the stream the current constants were copied on
  // work counters of the persistent launches: a ring, so that launches in flight on different streams never share one
  static constexpr int kWorkSlots = 64;
  unsigned *d_work = nullptr;
  unsigned long long work_seq = 0;
  bool have_scene = false;
  bool use_bvh = false;
  void *bvh_view = nullptr;         // rt::BvhView (rt_bvh.cu)
  std::vector<void *> bvh_allocs;   // device allocations owned by the BVH
  int sm_count = 0;
  uint64_t launches = 0;
  unsigned long long *d_ray_counters = nullptr;  // RT_FLAG_COUNT_RAYS
  uint32_t *signal_flag = nullptr;               // rt_signal_after_frame: applies to the next draw launch, then cleared
  uint32_t *strip_targets[8] = {};               // rt_set_strip_targets
  int n_strip_targets = 0, strip_rows = 0;
  const uint32_t *gate_flag = nullptr;           // rt_gate_next_frame: applies to the next draw launch, then cleared
  uint32_t gate_value = 0;
  const uint32_t *gate_flag_cached[2] = {nullptr, nullptr};  // the flags d_gate_seen[0..1] mirror
  unsigned gate_victim = 0;
  uint32_t *d_gate_seen = nullptr;
  int *d_wait_status = nullptr;                  // set by a rt_peer_wait kernel that timed out
  bool peer_waits = false;
  std::string last_kernel;                      // name of the draw kernel instantiation launched last (rt_last_kernel_name)
  size_t smem_optin = 0;                        // cudaDevAttrMaxSharedMemoryPerBlockOptin of the device
  // launch-order tables keyed by (row0, rows): centre-out order of the 16-row block grid
  struct TileOrder { int row0, rows, tile_w, tile_h; int *d_order; };
  std::vector<TileOrder> tile_orders;  // set by a launcher that needs shared memory beyond the scene
  float scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};  // bounding box of the triangles and the two spheres
  float mesh_lo[3] = {0, 0, 0}, mesh_hi[3] = {0, 0, 0};    // BVH scenes: bounding box of the triangles the tree covers (rt_bvh.cu)
  std::string err;
};

namespace rt {

// rt_draw.cu
cudaError_t launch_draw_brute(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream);
size_t brute_smem_bytes(int n, int n_sh);
// dynamic shared memory the fast kernels need on top of the scene (parked primary hits + jitter columns) for S shadow samples
size_t fast_extra_smem(int S);
// static shared memory of the draw kernels (per-warp caster lists, counters), rounded up
constexpr size_t kDrawStaticSmem = 6144;  // the mixed kernel carries the statics of both lane mappings (5.2 KB)
// rt_peak.cu (small utility kernels)
cudaError_t launch_peer_signal(uint32_t *flag, uint32_t value, cudaStream_t stream);
cudaError_t launch_peer_add(uint32_t *counter, cudaStream_t stream);
cudaError_t launch_peer_add_many(uint32_t *const *counters, int n, cudaStream_t stream);
cudaError_t launch_peer_wait(const uint32_t *flags, int n, uint32_t value, int *status, cudaStream_t stream);
// rt_bvh.cu
cudaError_t bvh_build(rt_ctx *ctx, const float *verts, const float *normals, const float *colors, int n);
void bvh_free(rt_ctx *ctx);
cudaError_t launch_draw_bvh(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream);
// rt_draw_fast.cu, one translation unit per shadow chunk size
cudaError_t launch_fast_ch1(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream);
cudaError_t launch_fast_ch2(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream);
cudaError_t launch_fast_ch4(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream);
cudaError_t launch_fast_ch5(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream);
cudaError_t launch_fast_ch8(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream);
cudaError_t launch_fast_ch10(rt_ctx *ctx, const FrameParams &fp, cudaStream_t stream);

}  // namespace rt
