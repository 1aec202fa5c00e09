/* uob_host.h — C ABI of the host-side pieces either side of the render path:
 * the scene sources that feed rt_upload_scene, the per-frame camera/light
 * state of the reference's main loop, and the headless framebuffer dump that
 * replaces SDL.  Implemented GLM-free in uob_raytracer_b200/csrc/host/.
 * Each function cites the reference code it restates; outputs are checked
 * bit-for-bit against the reference's own sources (oracle/_ref/libref_scene.so)
 * in tests/.
 */
#ifndef UOB_HOST_H
#define UOB_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Number of triangles LoadTestModel produces (TestModelH.h:44-219): 26. */
int uob_test_model_count(void);

/* LoadTestModel + the AoS->float4 flatten of skeleton.cpp:474-484.
 * verts: 3*cap float4, normals/colors: cap float4.  Returns the triangle count,
 * or -(needed) if cap is too small. */
int uob_load_test_model(float *verts_xyzw, float *normals_xyzw, float *colors_rgbm, int cap);

/* load_obj (Loader.cpp:11-59) + flatten: `v x y z` and `f a b c` (plain 1-based
 * indices) only; vertices x1.5; every triangle blue (0,0.2,0.4) with material
 * 0.5; normal from the scaled vertices, then v <- -v + (-0.4, 1.15, -0.7) with
 * the normal NOT recomputed.  Two-pass use: call with cap = 0 to get -(count).
 * Returns the count, -(needed) if cap is short, or INT_MIN on I/O or index error. */
int uob_load_obj(const char *path, float *verts_xyzw, float *normals_xyzw, float *colors_rgbm, int cap);

/* Rotation matrix of offload_rendering (skeleton.cpp:149-151): 3 rows, float4 stride. */
void uob_rot_matrix(float yaw, float pitch, float rot12[12]);

/* One step of update()'s light ping-pong (skeleton.cpp:290-298). */
void uob_light_step(float *light_x, int *lor);

/* Reference defaults (skeleton.cpp:61-67): focal 2200, camera (0,0,-3.2,1), light (0,-0.5,-0.7,1). */
void uob_default_camera(float *focal, float cam[4], float light[4]);

/* Focal length that keeps the box fitted to the frame height for other
 * resolutions / AA grids: f = 1100 * aa * height / 1024 (2200 at aa=2, 1024). */
float uob_fitted_focal(int aa, int height);

/* Headless replacements of SDL_SaveImage (SDLauxiliary.h:24-54). ARGB8888 in,
 * 24-bit BMP (bottom-up) or binary PPM out.  Return 0 on success. */
int uob_save_bmp(const char *path, const uint32_t *argb, int width, int height);
int uob_save_ppm(const char *path, const uint32_t *argb, int width, int height);

/* Synthetic mesh for the large-scene config: an icosphere with `subdiv`
 * subdivisions (20*4^subdiv faces) of radius `radius`, radially displaced by a
 * deterministic hash noise of amplitude `noise`, written as plain v/f OBJ.
 * Returns the face count or a negative value on I/O error. */
int uob_write_icosphere_obj(const char *path, int subdiv, float radius, float noise);

#ifdef __cplusplus
}
#endif
#endif
