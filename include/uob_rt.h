/* uob_rt.h — C ABI of the B200-native render path.
 *
 * This is the seam the reference crosses with OpenCL: `opencl_initialise`
 * (Source/skeleton.cpp:366-497) and `offload_rendering` (:146-182) around the
 * one device kernel `draw` (Source/kernels.cl:368-428).  Every entry point
 * below names the reference code it replaces.  Plain pointers and sizes only;
 * no C++ or torch types.  All functions are thread-compatible per context (one
 * host thread per rt_ctx, like the reference's single in-order queue).
 *
 * Error convention: functions return RT_OK (0) or an RT_ERR_* code and record a
 * message retrievable with rt_last_error(); the reference instead prints
 * "OpenCL error during '<op>' on line N" and exits (skeleton.cpp:499-507) — the
 * C++ host in uob_raytracer_b200/csrc/host keeps that print-and-exit behaviour
 * on top of these codes.  There is NO CPU fallback: without a CUDA device
 * rt_create() fails.
 */
#ifndef UOB_RT_H
#define UOB_RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rt_ctx rt_ctx;

enum {
  RT_OK = 0,
  RT_ERR_INVALID = 1,  /* bad argument */
  RT_ERR_CUDA = 2,     /* a CUDA runtime call failed (message has the op and cudaError) */
  RT_ERR_NO_SCENE = 3, /* rt_render before rt_upload_scene */
  RT_ERR_NO_DEVICE = 4 /* no usable CUDA device */
};

/* rt_config.flags */
enum {
  /* Arithmetic exactly as the CPU oracle: IEEE binary32, reference operation
   * order, no FMA contraction, IEEE 1/x and sqrt.  Frames are bit-identical to
   * oracle/cornell_oracle.c.  Off = the fast path (FMA, MUFU reciprocals,
   * division-free shadow tests), within 1/255 on >= 99.9 % of pixels. */
  RT_FLAG_STRICT_IEEE = 1u << 0,
  RT_FLAG_FORCE_BRUTE = 1u << 1, /* never build a BVH, whatever the triangle count */
  RT_FLAG_FORCE_BVH = 1u << 2,   /* always traverse a BVH, even for 26 triangles */
  /* Count the rays of every frame (rt_get_ray_counts).  Measurement aid: frames are rendered by the
   * generic kernel (same pixels), not the tuned one, so do not time a counting context. */
  RT_FLAG_COUNT_RAYS = 1u << 3,
  /* With RT_FLAG_STRICT_IEEE: the reference's plain loop structure (every ray against every triangle), without the
   * conservative culls the strict path otherwise shares with the fast one.  Same frame bit for bit, several times
   * slower; kept as the anchor the culled strict path is tested against. */
  RT_FLAG_REFERENCE_LOOPS = 1u << 4,
  /* Four lanes per pixel (each traces every fourth ray of the pixel's aa*aa; aa = 2 or 4 only).  Same frame.  The
   * default picks it by itself for launches too small to fill the GPU (a share of a frame on 4 or 8 GPUs), where
   * the launch lasts as long as its slowest pixel; these two flags force it on or off. */
  RT_FLAG_SPLIT_PIXELS = 1u << 5,
  RT_FLAG_NO_SPLIT = 1u << 6,
  /* Four lanes per pixel only for the tiles that can see a sphere (the pixels with mirror / glass bounce chains), ordinary
   * mapping elsewhere, in one launch.  What the default chooses for small launches. */
  RT_FLAG_SPLIT_HEAVY = 1u << 7,
  /* rt_stream_wait_geq / rt_stream_write use one-warp kernels instead of the driver's stream memory operations */
  RT_FLAG_NO_STREAM_MEMOPS = 1u << 8
};

/* Everything that is a compile-time constant in the reference
 * (skeleton.cpp:27-34, kernels.cl:12-17, :316, :343) plus the row tile this
 * context renders. */
typedef struct rt_config {
  int width, height;  /* SCREEN_WIDTH / SCREEN_HEIGHT of the whole frame */
  int aa;             /* rays_x = rays_y (kernels.cl:12-13); aa_rays = aa*aa */
  int shadow_samples; /* light_sources (kernels.cl:316) */
  int max_bounces;    /* bounces (kernels.cl:343) */
  int device;         /* CUDA ordinal; the reference picks via OCL_DEVICE (skeleton.cpp:551) */
  int row0, rows;     /* this context renders frame rows [row0, row0+rows); rows<=0 = whole frame.
                         Pixel ids and ray directions stay frame-global, so N row tiles
                         concatenate to exactly the 1-GPU frame. */
  uint32_t flags;
  /* Finer multi-GPU partition: the 16x16-pixel blocks of rows [row0,row0+rows) are numbered
   * row-major and this context renders blocks block_phase, block_phase+block_stride, ... —
   * rank g of N uses stride N, phase g; N interleaved contexts cover the rows exactly once
   * with near-perfect load balance.  block_stride <= 1 = every block (plain row tile). */
  int block_stride, block_phase;
} rt_config;

/* Defaults of the reference at HEAD: 1024x1024, aa 2, 10 shadow samples, 10 bounces, device 0. */
void rt_default_config(rt_config *cfg);

/* Replaces selectOpenCLDevice + context/queue/program/kernel/buffer creation
 * (skeleton.cpp:374-446).  Returns NULL on failure; rt_last_error(NULL) then
 * holds the reason. */
rt_ctx *rt_create(const rt_config *cfg);

/* Replaces the scene flatten + three blocking clEnqueueWriteBuffer calls
 * (skeleton.cpp:474-496).  verts_xyzw: 3n float4 (v0,v1,v2 per triangle, w
 * ignored); normals_xyzw: n float4; colors_rgbm: n float4 with w = material
 * (>0 diffuse, 0 mirror, <0 glass; -1 casts no shadow).  Triangle order is
 * significant (lowest index wins exact ties, kernels.cl:120).  May be called
 * again to replace the scene. */
int rt_upload_scene(rt_ctx *ctx, const float *verts_xyzw, const float *normals_xyzw,
                    const float *colors_rgbm, int n_triangles);

/* Replaces offload_rendering (skeleton.cpp:146-182): rotation-matrix upload,
 * the four per-frame kernel arguments, the NDRange launch and the BLOCKING
 * read-back.  rot12 = three rows with float4 stride (:149-151); cam/light =
 * the 16 bytes the reference passes as float3 (w ignored, :162-165).
 * host_argb receives rows [row0,row0+rows) only: rows*width uint32 ARGB8888
 * (kernels.cl:37-40), i.e. the whole frame for an untiled context.  The frame
 * is valid on return.  Internally the tile is rendered in four row bands whose
 * read-back overlaps the rendering of the next band; pass pinned memory
 * (cudaHostAlloc / cudaHostRegister) for the overlap and full copy speed. */
int rt_render(rt_ctx *ctx, const float rot12[12], const float cam[4], const float light[4],
              float focal_length, uint32_t *host_argb);

/* Page-locked host memory for frame buffers — what `screen->buffer = new uint32_t[W*H]`
 * (SDLauxiliary.h:105) should become: read-back into pageable memory is several times slower and cannot
 * overlap with rendering.  rt_host_free(NULL) is a no-op. */
void *rt_host_alloc(size_t bytes);
void rt_host_free(void *p);

/* Pipelined form of rt_render for a render loop (the reference's main loop renders frame after
 * frame, skeleton.cpp:117-138): rt_render_begin enqueues the frame and its read-back into host_argb
 * and returns at once; rt_render_end blocks until the OLDEST frame begun and not yet ended is
 * complete in its host buffer.  Up to two frames may be in flight (two device frame buffers): the
 * read-back of frame k overlaps the kernel of frame k+1, so a loop
 *     begin(f0); for k: begin(f[k+1]); end();  ...  end();
 * runs at max(kernel, copy) per frame instead of their sum.  Frames are identical to rt_render's.
 * Each host buffer must stay untouched until its rt_render_end returns. */
int rt_render_begin(rt_ctx *ctx, const float rot12[12], const float cam[4], const float light[4],
                    float focal_length, uint32_t *host_argb);
int rt_render_end(rt_ctx *ctx);

/* Kernel only, asynchronous, no read-back.  dev_argb: device pointer to the
 * WHOLE frame (width*height uint32) — this context's tile is written at row
 * offset row0; NULL = the context's own frame buffer (rt_device_frame).  May
 * be a peer-mapped pointer of another GPU (NVLink store path).  stream: a
 * cudaStream_t, or NULL for the context's stream.  Timing of the last launch:
 * rt_last_kernel_ms (synchronises). */
int rt_render_device(rt_ctx *ctx, const float rot12[12], const float cam[4], const float light[4],
                     float focal_length, uint32_t *dev_argb, void *stream);

/* Make `stream` (a cudaStream_t of the context's device) the context's stream for all later
 * launches, copies and rt_synchronize; NULL restores the context's own stream.  Lets a host that
 * already owns a stream (torch, a render loop with its own queue) keep everything in order. */
int rt_set_stream(rt_ctx *ctx, void *stream);

/* Wait for everything queued on the context's stream. */
int rt_synchronize(rt_ctx *ctx);

/* Device pointer of the context's own whole-frame buffer (width*height uint32). */
uint32_t *rt_device_frame(rt_ctx *ctx);

/* Milliseconds (CUDA events on the launching stream) of the last render launch
 * made by rt_render / rt_render_device; < 0 on error. */
float rt_last_kernel_ms(rt_ctx *ctx);

/* Number of kernels launched by this context so far. */
uint64_t rt_kernel_launches(const rt_ctx *ctx);

/* "brute" or "bvh": which traversal rt_upload_scene selected. */
const char *rt_scene_mode(const rt_ctx *ctx);

/* Name of the draw kernel instantiation the last render call launched, e.g.
 * "draw_fast_kernel<8,true,false,plain>" (shadow chunk, S == chunk, strict arithmetic, lane mapping: plain / split / mixed)
 * or "draw_bvh_kernel<float,8>" — what bench.py prints as roofline.kernel and what the tests pin the benchmarked instantiation with.  "" before the
 * first launch.  Diagnostic; no reference counterpart. */
const char *rt_last_kernel_name(const rt_ctx *ctx);

/* The reference never releases anything (no clRelease* anywhere); this does. */
void rt_destroy(rt_ctx *ctx);

/* Last error message of ctx (or of rt_create when ctx == NULL). Never NULL. */
const char *rt_last_error(const rt_ctx *ctx);

/* Rays traced by the last frame of a RT_FLAG_COUNT_RAYS context: {primary, shadow, bounce}, with
 * the definition of SURVEY.md §8d (one ray = one closest-hit search or one in_shadow call).  Under
 * RT_FLAG_STRICT_IEEE the counts equal the CPU oracle's counters exactly (tests); this is where the
 * ray count of scenes too large for the oracle (1.3 M triangles) comes from. */
int rt_get_ray_counts(rt_ctx *ctx, uint64_t counts[3]);

/* Diagnostic: measured FP32 FFMA throughput of this device in TFLOP/s (dependent-chain-free
 * FFMA microbenchmark, best of 5, CUDA events) — the roofline denominator bench.py reports
 * the render kernel against.  Not part of the reference's surface. */
int rt_measure_fp32_peak(rt_ctx *ctx, float *tflops);

/* Multi-GPU without a collective: every rank writes its pixels straight into ONE rank's frame
 * buffer over NVLink (peer-mapped stores from inside the draw kernel).
 *   single process:  rt_enable_peer(ctx, peer_device) then pass the other context's
 *                    rt_device_frame() as dev_argb of rt_render_device;
 *   one process per GPU:  the owner exports its frame with rt_ipc_export_frame (64-byte CUDA IPC
 *                    handle, to be sent to the peers by any means), the peers map it with
 *                    rt_ipc_open_frame and pass the mapped pointer as dev_argb.
 * Completion of the peers' stores is the caller's job (a stream/event wait in one process, any
 * barrier enqueued after the kernel across processes). */
int rt_enable_peer(rt_ctx *ctx, int peer_device);
int rt_ipc_export_frame(rt_ctx *ctx, void *handle64);
int rt_ipc_open_frame(rt_ctx *ctx, const void *handle64, uint32_t **dev_argb);
int rt_ipc_close_frame(rt_ctx *ctx, uint32_t *dev_argb);

/* Frame hand-over between GPUs without a collective.  Every frame buffer carries RT_PEER_FLAGS 32-bit
 * flags behind its pixels (rt_peer_flags() of the owner; at dev_argb + width*height through a mapping).
 * rt_peer_signal enqueues "all my earlier work on this stream is visible system-wide, then flag = value";
 * rt_peer_wait enqueues "wait until each of the n flags is >= value" (a one-warp kernel spinning with
 * system-scope loads; gives up after ~2 s and records an error readable with rt_last_error after
 * rt_synchronize).  Typical frame f:  peers: wait(owner's consumed flag >= f-1), draw into the owner's
 * frame, signal(done[rank] = f);  owner: draw, wait(done[1..N-1] >= f), read back, signal(consumed = f).
 * Only between DIFFERENT GPUs: a waiting kernel and the kernel it waits for must not share a device. */
#define RT_PEER_FLAGS 64
uint32_t *rt_peer_flags(rt_ctx *ctx);
int rt_peer_signal(rt_ctx *ctx, uint32_t *dev_flag, uint32_t value, void *stream);
int rt_peer_wait(rt_ctx *ctx, const uint32_t *dev_flags, int n, uint32_t value, void *stream);
/* The same wait folded into the next draw launch of this context (rt_render_device): none of its pixels is stored
 * before *dev_flag >= value.  Cheaper than rt_peer_wait in front of the launch — no extra kernel; each block checks
 * a device-local copy of the flag and only the first ones poll the owner's memory.  One launch, then cleared.
 * No reference counterpart (the reference has one device and one in-order queue, skeleton.cpp:388). */
int rt_gate_next_frame(rt_ctx *ctx, const uint32_t *dev_flag, uint32_t value);

/* Double-buffered hand-over.  A context owns TWO frame slots in one allocation (so one IPC handle maps both): slot s
 * starts rt_frame_slot_words(ctx) * s words behind rt_device_frame(ctx) and is followed by its own RT_PEER_FLAGS words.
 * Frame f goes to slot f & 1, so the peers may store frame f + 1 while the owner still reads frame f back.
 *   rt_signal_after_frame: the next draw launch of this context ends with "all my pixels are visible system-wide, then
 *                          *dev_counter += 1" — done by the last block of the draw kernel itself, no extra launch.  The
 *                          owner waits for the count of deliveries: (ranks - 1) per frame that used the slot.
 *   rt_stream_wait_geq:    stream-ordered wait until *dev_word >= value  (driver stream memory operation, no kernel)
 *   rt_stream_write:       stream-ordered *dev_word = value after everything queued before it (e.g. "slot consumed")
 * No reference counterpart (one device, one in-order queue: skeleton.cpp:388). */
size_t rt_frame_slot_words(const rt_ctx *ctx);
int rt_signal_after_frame(rt_ctx *ctx, uint32_t *dev_counter);
int rt_stream_wait_geq(rt_ctx *ctx, const uint32_t *dev_word, uint32_t value, void *stream);
int rt_stream_write(rt_ctx *ctx, uint32_t *dev_word, uint32_t value, void *stream);
int rt_read_frame_slot(rt_ctx *ctx, int slot, uint32_t *host_argb);

/* Parallel egress.  The reference reads the frame back over one link (skeleton.cpp:179); with N GPUs behind N PCIe links
 * the frame can leave the GPUs N times faster if every GPU holds the rows it will copy.  rt_set_strip_targets deals the
 * rows out in strips of strip_rows rows (a multiple of 16): from now on the draw kernels of this context store row y into
 * dev_frames[(y / strip_rows) % n] — the whole-frame buffer of the strip's owner, its own or a peer's (mapped with
 * rt_ipc_open_frame; the stores cross NVLink) — whatever partition of the pixels the context renders.  Every GPU then
 * ends up with strips phase, phase + n, ... of the complete frame and rt_read_strips copies exactly those into the same
 * rows of a host frame (asynchronously; a 2-D copy).  With the host frame in memory that all ranks' processes share and
 * have registered (rt_host_register on a shared mapping), the N copies run side by side.  n <= 1 switches it off.
 * Tuned brute-force kernels only. */
int rt_set_strip_targets(rt_ctx *ctx, uint32_t *const *dev_frames, int n, int strip_rows);
/* "everything queued on the stream so far is visible system-wide, then *dev_counters[i] += 1 for i < n" (n <= 8): one
 * small kernel that reports a delivery to several strip owners at once. */
int rt_peer_add(rt_ctx *ctx, uint32_t *const *dev_counters, int n, void *stream);
int rt_read_strips(rt_ctx *ctx, int slot, int strip_rows, int n, int phase, uint32_t *host_argb, void *stream);
/* One parallel-egress frame in one call, asynchronous on the context's stream: rows of frame slot `slot` dealt out over
 * dev_frames[0..n) (slot 0 bases, e.g. rt_device_frame / rt_ipc_open_frame), draw, report the delivery to the other owners,
 * wait for `deliveries_expected` deliveries into this rank's strips ((n - 1) per frame that has used the slot), copy this
 * rank's strips into host_argb.  Then rt_synchronize and a barrier of the ranks' hosts complete the frame. */
int rt_render_strips(rt_ctx *ctx, const float rot12[12], const float cam[4], const float light[4], float focal_length,
                     uint32_t *const *dev_frames, int n, int rank, int strip_rows, int slot, uint32_t deliveries_expected,
                     uint32_t *host_argb);
int rt_host_register(void *p, size_t bytes);
int rt_host_unregister(void *p);

/* Host-side launch planning, exposed for the CPU test suite (pure functions: no context, no device).  No reference
 * counterpart — the reference launches one work-item per pixel over the whole frame (skeleton.cpp:170-172).
 * rt_debug_visible_rect: the pixel rectangle {x0, y0, x1, y1} (half-open) outside which no primary ray of this camera
 * can hit the box lo..hi; tiles outside it are written black without looking at the scene.
 * rt_debug_tile_lists: what a mixed launch for cfg (width, height, aa, row0, rows) and this camera renders, in launch
 * order: tiles[0..n_light) = row-major numbers of the ordinary 16x16 tiles, tiles[n_light..n_light+n_split) = numbers of
 * the 8x8 sub-tiles (on a grid ceil(width/8) wide) of the tiles inside a sphere's screen rectangle.  capacity = room in tiles[]. */
int rt_debug_visible_rect(const rt_config *cfg, const float lo[3], const float hi[3], const float rot12[12], const float cam[4], float focal,
                          int rect[4]);
int rt_debug_tile_lists(const rt_config *cfg, const float rot12[12], const float cam[4], float focal, int *tiles, int capacity, int *n_light,
                        int *n_split);

/* Blocking read-back of the WHOLE frame buffer of this context (width*height uint32) — what the
 * owner of a peer-written frame calls once the peers are done. */
int rt_read_frame(rt_ctx *ctx, uint32_t *host_argb);

/* Library version string, e.g. "uob_rt 0.1 (sm_100a)". */
const char *rt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* UOB_RT_H */
