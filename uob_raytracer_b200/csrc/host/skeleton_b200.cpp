// skeleton_b200.cpp — the reference's host loop (Source/skeleton.cpp:93-144) on the B200 path.
//
// Same scene, camera and render-call surface as the reference: process globals focal_length,
// camera_position, light_position, yaw, pitch, triangles (skeleton.cpp:61-72); `cuda_initialise`
// stands where `opencl_initialise` (:366-497) stood and `offload_rendering` (:146-182) keeps its
// name, argument meaning and blocking behaviour, calling the C ABI of include/uob_rt.h instead of
// OpenCL.  SDL is replaced by a headless framebuffer dump: the loop runs a fixed number of frames
// (the reference runs until ESC) and writes screenshot.bmp on exit like SDL_SaveImage (:139).
// update() keeps the deterministic part of the reference's update() — the light ping-pong
// (:290-298); mouse and keyboard handling is out of scope.
//
//   skeleton_b200 [--frames N] [--width W --height H] [--aa A] [--shadow S] [--bounces B]
//                 [--obj mesh.obj] [--gpus G] [--strict] [--pipeline] [--out screenshot.bmp] [--quiet]
#include <chrono>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/uob_host.h"
#include "../../../include/uob_rt.h"

// ---- process globals of the reference (skeleton.cpp:27-34, :61-74) -------------------------------
static int SCREEN_WIDTH = 1024;
static int SCREEN_HEIGHT = 1024;
static float focal_length = 2200.0f;
static float camera_position[4] = {0.0f, 0.0f, -3.2f, 1.0f};
static float light_position[4] = {0.0f, -0.5f, -0.7f, 1.0f};
static float yaw = 0.0f, pitch = 0.0f;
static int lor = 1;

struct Triangles {  // the flattened `vector<Triangle> triangles` (skeleton.cpp:72, :474-484)
  std::vector<float> verts, normals, colors;
  int n = 0;
};
static Triangles triangles;

struct screen {  // SDLauxiliary.h:9-16 without the SDL handles
  int width, height;
  uint32_t *buffer;
};

// t_ocl of the reference (skeleton.cpp:36-52) becomes one context per GPU.  With several GPUs the
// 16x16 blocks of the frame are interleaved over them and every GPU's kernel stores its pixels
// straight into GPU 0's frame over NVLink (peer access), so offload_rendering still ends in ONE
// read-back from GPU 0.
struct t_rt {
  std::vector<rt_ctx *> ctx;
};

static void checkError(int err, rt_ctx *ctx, const char *op, const int line) {  // skeleton.cpp:499-507
  if (err != RT_OK) {
    fprintf(stderr, "CUDA error during '%s' on line %d: %d (%s)\n", op, line, err, rt_last_error(ctx));
    fflush(stderr);
    exit(EXIT_FAILURE);
  }
}

static void cuda_initialise(t_rt *rt, int gpus, int aa, int shadow, int bounces, unsigned flags) {
  for (int g = 0; g < gpus; g++) {
    rt_config cfg;
    rt_default_config(&cfg);
    cfg.width = SCREEN_WIDTH;
    cfg.height = SCREEN_HEIGHT;
    cfg.aa = aa;
    cfg.shadow_samples = shadow;
    cfg.max_bounces = bounces;
    cfg.device = g;
    cfg.flags = flags;
    if (gpus > 1) {
      cfg.block_stride = gpus;
      cfg.block_phase = g;
    }
    rt_ctx *c = rt_create(&cfg);
    if (!c) {
      fprintf(stderr, "CUDA error during 'creating context' on line %d: %s\n", __LINE__, rt_last_error(nullptr));
      exit(EXIT_FAILURE);
    }
    if (g > 0) checkError(rt_enable_peer(c, 0), c, "enabling peer access", __LINE__);
    checkError(rt_upload_scene(c, triangles.verts.data(), triangles.normals.data(), triangles.colors.data(), triangles.n), c,
               "writing triangle buffer data", __LINE__);
    rt->ctx.push_back(c);
  }
}

static void offload_rendering(screen *screen, t_rt rt) {  // skeleton.cpp:146-182
  float rot_matrix[12];
  uob_rot_matrix(yaw, pitch, rot_matrix);
  if (rt.ctx.size() == 1) {
    checkError(rt_render(rt.ctx[0], rot_matrix, camera_position, light_position, focal_length, screen->buffer), rt.ctx[0],
               "enqueueing draw kernel / reading screen buffer data", __LINE__);
    return;
  }
  uint32_t *frame0 = rt_device_frame(rt.ctx[0]);
  for (rt_ctx *c : rt.ctx)
    checkError(rt_render_device(c, rot_matrix, camera_position, light_position, focal_length, frame0, nullptr), c,
               "enqueueing draw kernel", __LINE__);
  for (rt_ctx *c : rt.ctx) checkError(rt_synchronize(c), c, "waiting for the draw kernel", __LINE__);
  checkError(rt_read_frame(rt.ctx[0], screen->buffer), rt.ctx[0], "reading screen buffer data", __LINE__);
}

static bool update() {  // skeleton.cpp:282-361, deterministic part
  uob_light_step(&light_position[0], &lor);
  return false;
}

static void append_scene(const char *path_or_null) {
  int n;
  if (path_or_null) {
    // one counting call (capacity 0): INT32_MIN = unreadable file or bad face index, otherwise minus the triangle count
    const int rc = uob_load_obj(path_or_null, nullptr, nullptr, nullptr, 0);
    if (rc == INT32_MIN) {
      fprintf(stderr, "Error: could not read OBJ file: %s\n", path_or_null);
      exit(EXIT_FAILURE);
    }
    n = rc < 0 ? -rc : rc;
  } else {
    n = uob_test_model_count();
  }
  const size_t old = (size_t)triangles.n;
  triangles.verts.resize(12 * (old + n));
  triangles.normals.resize(4 * (old + n));
  triangles.colors.resize(4 * (old + n));
  float *v = triangles.verts.data() + 12 * old, *nn = triangles.normals.data() + 4 * old, *c = triangles.colors.data() + 4 * old;
  const int got = path_or_null ? uob_load_obj(path_or_null, v, nn, c, n) : uob_load_test_model(v, nn, c, n);
  if (got != n) {
    fprintf(stderr, "Error: scene source returned %d triangles, expected %d\n", got, n);
    exit(EXIT_FAILURE);
  }
  triangles.n += n;
}

int main(int argc, char *argv[]) {
  int frames = 10, aa = 2, shadow = 10, bounces = 10, gpus = 1;
  unsigned flags = 0;
  bool quiet = false, size_given = false, pipeline = false;
  const char *obj = nullptr, *out = "screenshot.bmp";
  for (int i = 1; i < argc; i++) {
    auto next = [&](const char *name) -> const char * {
      if (i + 1 >= argc) {
        fprintf(stderr, "missing value for %s\n", name);
        exit(EXIT_FAILURE);
      }
      return argv[++i];
    };
    if (!strcmp(argv[i], "--frames")) frames = atoi(next("--frames"));
    else if (!strcmp(argv[i], "--width")) SCREEN_WIDTH = atoi(next("--width")), size_given = true;
    else if (!strcmp(argv[i], "--height")) SCREEN_HEIGHT = atoi(next("--height")), size_given = true;
    else if (!strcmp(argv[i], "--aa")) aa = atoi(next("--aa")), size_given = true;
    else if (!strcmp(argv[i], "--shadow")) shadow = atoi(next("--shadow"));
    else if (!strcmp(argv[i], "--bounces")) bounces = atoi(next("--bounces"));
    else if (!strcmp(argv[i], "--gpus")) gpus = atoi(next("--gpus"));
    else if (!strcmp(argv[i], "--obj")) obj = next("--obj");
    else if (!strcmp(argv[i], "--out")) out = next("--out");
    else if (!strcmp(argv[i], "--strict")) flags |= RT_FLAG_STRICT_IEEE;
    else if (!strcmp(argv[i], "--quiet")) quiet = true;
    else if (!strcmp(argv[i], "--pipeline")) pipeline = true;
    else {
      fprintf(stderr, "unknown argument %s\n", argv[i]);
      return EXIT_FAILURE;
    }
  }
  if (size_given) focal_length = uob_fitted_focal(aa, SCREEN_HEIGHT);  // keeps the box fitted (2200 at aa 2, 1024)

  t_rt rt;
  // InitializeSDL allocates the frame with new[] (SDLauxiliary.h:105); page-locked memory reads back several times faster
  const size_t frame_bytes = sizeof(uint32_t) * (size_t)SCREEN_WIDTH * SCREEN_HEIGHT;
  screen scr{SCREEN_WIDTH, SCREEN_HEIGHT, static_cast<uint32_t *>(rt_host_alloc(frame_bytes))};
  if (!scr.buffer) {
    fprintf(stderr, "Error: could not allocate the frame buffer (no CUDA device?)\n");
    return EXIT_FAILURE;
  }
  screen *screen = &scr;

  // Load Cornell Box (+ mesh, as the commented-out call site skeleton.cpp:102-103 would)
  append_scene(nullptr);
  if (obj) append_scene(obj);
  printf("Triangles Length size %lu\n", (unsigned long)triangles.n);

  cuda_initialise(&rt, gpus, aa, shadow, bounces, flags);
  printf("Render path: %s (%s), %d GPU(s)\n", rt_version(), rt_scene_mode(rt.ctx[0]), gpus);

  // Draw initial scene
  offload_rendering(screen, rt);

  double total_us = 0.0;
  if (pipeline && gpus == 1 && frames > 0) {
    // Render loop with two frames in flight (rt_render_begin / rt_render_end): frame k is read back while
    // frame k+1 renders.  Same frames, same order; the per-frame time is the loop's throughput.
    uint32_t *second = static_cast<uint32_t *>(rt_host_alloc(frame_bytes));
    if (!second) {
      fprintf(stderr, "Error: could not allocate the second frame buffer\n");
      return EXIT_FAILURE;
    }
    uint32_t *bufs[2] = {screen->buffer, second};
    float rot_matrix[12];
    auto start = std::chrono::high_resolution_clock::now();
    for (int f = 0; f < frames; f++) {
      update();
      uob_rot_matrix(yaw, pitch, rot_matrix);
      checkError(rt_render_begin(rt.ctx[0], rot_matrix, camera_position, light_position, focal_length, bufs[f & 1]), rt.ctx[0],
                 "enqueueing draw kernel", __LINE__);
      if (f > 0) checkError(rt_render_end(rt.ctx[0]), rt.ctx[0], "reading screen buffer data", __LINE__);
    }
    checkError(rt_render_end(rt.ctx[0]), rt.ctx[0], "reading screen buffer data", __LINE__);
    auto stop = std::chrono::high_resolution_clock::now();
    total_us = (double)std::chrono::duration_cast<std::chrono::microseconds>(stop - start).count();
    if ((frames - 1) & 1) memcpy(screen->buffer, second, sizeof(uint32_t) * (size_t)SCREEN_WIDTH * SCREEN_HEIGHT);
    rt_host_free(second);
    frames = -frames;  // skip the blocking loop below
  }
  for (int f = 0; f < frames; f++) {
    update();
    auto start = std::chrono::high_resolution_clock::now();
    offload_rendering(screen, rt);
    auto stop = std::chrono::high_resolution_clock::now();
    const long long us = std::chrono::duration_cast<std::chrono::microseconds>(stop - start).count();
    total_us += (double)us;
    if (!quiet) {
      printf("\nOffloaded GPU Rendertime: %lld micro seconds\n", us);
      printf("Frame Rate: %fFPS\n", 1000000.0f / ((float)us));
    }
  }
  if (frames < 0) frames = -frames;
  if (frames > 0) printf("\n%d frames, mean %.1f micro seconds per frame (%.1f FPS)\n", frames, total_us / frames, 1e6 * frames / total_us);
  if (uob_save_bmp(out, screen->buffer, screen->width, screen->height) != 0) fprintf(stderr, "could not write %s\n", out);
  for (rt_ctx *c : rt.ctx) rt_destroy(c);
  rt_host_free(scr.buffer);  // KillSDL (SDLauxiliary.h:58)
  return 0;
}
