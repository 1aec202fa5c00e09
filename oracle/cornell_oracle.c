/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement, in plain C99, of the per-pixel render path of
 * harrywaugh/UOB_Raytracer (`kernel draw`, Source/kernels.cl:368-428, and the
 * ten device helpers it calls).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / reference legs of bench.py may load this; the product
 * (uob_raytracer_b200/) never does.
 *
 * PARITY PINNING.  The reference ships no tests, golden images or known-answer
 * vectors for this path, so this restatement is pinned
 * against the reference ITSELF: oracle/build_ref.py compiles the verbatim
 * kernels.cl with g++ (six token rewrites + oracle/cl_shim.h) into
 * oracle/_ref/, and tests/test_oracle_vs_ref.py requires this file to
 * reproduce its frames BIT FOR BIT for every BASELINE config variant;
 * tests/golden/ holds frames and hashes generated that way
 * (tests/golden/make_golden.py).
 *
 * Arithmetic conventions (identical to oracle/cl_shim.h): IEEE-754 binary32,
 * source order, no FMA contraction (build with -O2 -ffp-contract=off),
 * native_recip = 1/x, native_divide = a/b, native_sqrt = sqrtf,
 * normalize(v) = v * (1/sqrtf(dot(v,v))), dot = (x*x' + y*y') + z*z',
 * min(x,y) = y<x?y:x, max(x,y) = x<y?y:x.
 *
 * What is run-time here but compile-time in the reference: W, H
 * (kernels.cl:16-17), the AA grid edge A (:12-14), shadow samples S (:316),
 * max bounces B (:343).
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;

/* kernels.cl:21-29 */
typedef struct {
  v3 start, direction, intersect, intersect_normal;
  v4 intersect_color;
  float medium;
  int intersect_triangle;
} Ray;

typedef struct {
  int W, H;            /* frame size (SCREEN_WIDTH / SCREEN_HEIGHT) */
  int aa;              /* rays_x = rays_y = aa, aa_rays = aa*aa */
  int shadow_samples;  /* light_sources */
  int max_bounces;     /* bounces */
  int y0, y1, row_step;/* rows y0, y0+row_step, ... < y1 are rendered */
  int threads;         /* <=0: all */
  float focal;
} oracle_params;

/* counters[]: 0 primary rays, 1 shadow rays, 2 bounce rays, 3 closest-hit
 * triangle tests (T_c), 4 shadow stage-1 tests (T_s1), 5 shadow stage-2 tests
 * (T_s2), 6 sphere tests (T_sph, closest + shadow), 7 pixels. */
enum { C_PRIMARY, C_SHADOW, C_BOUNCE, C_TC, C_TS1, C_TS2, C_TSPH, C_PIXELS, C_COUNT };

typedef struct {
  const v4 *verts;   /* 3n, w ignored (skeleton.cpp:479-481) */
  const v4 *normals; /* n */
  const v4 *colors;  /* n, w = material */
  int n;
  int S, B;
  uint64_t c[C_COUNT];
} Scene;

/* kernels.cl:3-10, :18-19 */
#define GLASS 1.52f
#define AIR 1.0f
#define SPHERES 2
static const float k_indirect = 0.5f;
static const float k_light_color = 16.0f;
static const float k_bias = 0.0001f;
static const v4 k_sphere_centers[SPHERES] = {{0.3f, 0.1f, -0.5f, 0.0f}, {-0.4f, 0.8f, -0.5f, 0.0f}};
static const v4 k_sphere_colors[SPHERES] = {{0.0f, 0.0f, 0.0f, -1.0f}, {0.0f, 0.0f, 0.0f, 0.0f}};
static const float k_sphere_radius_sqs[SPHERES] = {0.075f, 0.05f};

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 xyz(v4 a) { return V3(a.x, a.y, a.z); }
static inline v3 add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 scale(float s, v3 a) { return V3(s * a.x, s * a.y, s * a.z); }
static inline v3 neg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 normalize3(v3 v) { const float inv = 1.0f / sqrtf(dot3(v, v)); return V3(v.x * inv, v.y * inv, v.z * inv); }
static inline float cl_min(float x, float y) { return y < x ? y : x; }
static inline float cl_max(float x, float y) { return x < y ? y : x; }

/* kernels.cl:31-35 */
static inline float det3(v3 m0, v3 m1, v3 m2) {
  return m0.x * (m1.y * m2.z - m1.z * m2.y) -
         m0.y * (m1.x * m2.z - m1.z * m2.x) +
         m0.z * (m1.x * m2.y - m1.y * m2.x);
}

/* kernels.cl:42-47 — component-wise xorshift32 */
static inline uint32_t xorshift32(uint32_t s) {
  s ^= s << 13;
  s ^= s >> 17;
  s ^= s << 5;
  return s;
}

/* kernels.cl:49-52 — (float)UINT_MAX rounds to 2^32 */
static inline float crush1(uint32_t v, float range) {
  const float fl = range * (float)v / 4294967296.0f;
  return fl - range / 2.f;
}

/* kernels.cl:54-65 */
static Ray reflect_ray(const Ray *ray) {
  Ray r;
  memset(&r, 0, sizeof r); /* reference leaves the rest uninitialised; never read */
  r.intersect_triangle = -1;
  r.intersect_color.w = 1.0f;
  const float dn = dot3(ray->direction, ray->intersect_normal);
  r.direction = sub(ray->direction, scale(2.0f, scale(dn, ray->intersect_normal)));
  r.start = add(ray->intersect, scale(k_bias, r.direction));
  r.medium = AIR;
  r.direction = normalize3(r.direction);
  return r;
}

/* kernels.cl:67-88.  The `c2 < 0` TIR branch is dead code for real inputs:
 * sqrt of a negative is NaN, and NaN < 0 is false. */
static Ray refract_ray(const Ray *ray) {
  Ray r;
  memset(&r, 0, sizeof r);
  v3 normal = ray->intersect_normal;
  const int air = (ray->medium == AIR);
  const float n1 = air ? AIR : GLASS;
  const float n2 = air ? GLASS : AIR;
  float c1 = dot3(normal, ray->direction);
  if (c1 < 0.0f) normal = scale(-1.0f, normal);
  c1 = fabsf(c1);
  const float n = n1 / n2;
  const float c2 = sqrtf(1.0f - (n * n) * (1.0f - (c1 * c1)));
  if (c2 < 0.0f) return reflect_ray(ray);
  r.intersect_triangle = -1;
  r.intersect_color.x = 1.0f; r.intersect_color.y = 0.0f; r.intersect_color.z = 0.0f; r.intersect_color.w = 1.0f;
  r.direction = add(scale(n, ray->direction), scale(n * c1 - c2, neg(normal)));
  r.start = add(ray->intersect, scale(k_bias, r.direction));
  r.medium = n2;
  r.direction = normalize3(r.direction);
  return r;
}

/* kernels.cl:92-166 (one ray of the batch) == :168-241 */
static void closest_hit(Ray *ray, Scene *sc) {
  float current_t = FLT_MAX;
  const v3 nd = neg(ray->direction);
  for (int i = 0; i < sc->n; i++) {
    const v3 v0 = xyz(sc->verts[i * 3]);
    const v3 e1 = sub(xyz(sc->verts[i * 3 + 1]), v0);
    const v3 e2 = sub(xyz(sc->verts[i * 3 + 2]), v0);
    const v3 b = sub(ray->start, v0);
    const float inv = 1.0f / det3(nd, e1, e2);
    const float t = det3(b, e1, e2) * inv;
    const float u = det3(nd, b, e2) * inv;
    const float v = det3(nd, e1, b) * inv;
    if (t < current_t && u >= 0 && v >= 0 && (u + v) <= 1 && t >= 0) {
      ray->intersect_triangle = i;
      ray->intersect = add(add(v0, scale(u, e1)), scale(v, e2));
      ray->intersect_normal = xyz(sc->normals[i]);
      ray->intersect_color = sc->colors[i];
      current_t = t;
    }
  }
  sc->c[C_TC] += (uint64_t)sc->n;
  for (int i = 0; i < SPHERES; i++) {
    sc->c[C_TSPH]++;
    const v3 ctr = xyz(k_sphere_centers[i]);
    const v3 L = sub(ray->start, ctr);
    const float a = dot3(ray->direction, ray->direction);
    const float b = 2.0f * dot3(ray->direction, L);
    const float c = dot3(L, L) - k_sphere_radius_sqs[i];
    const float disc = b * b - 4.0f * a * c;
    if (disc < 0.0f) continue;
    const float q = (b > 0) ? -0.5f * (b + sqrtf(disc)) : -0.5f * (b - sqrtf(disc));
    const float x0 = q / a;
    const float x1 = c / q;
    const float x_min = cl_min(x0, x1);
    const float x_max = cl_max(x0, x1);
    if (x_min >= 0.0f && x_min < current_t) {
      ray->intersect_triangle = -2;
      ray->intersect = add(ray->start, scale(x_min, ray->direction));
      ray->intersect_normal = normalize3(sub(ray->intersect, ctr));
      ray->intersect_color = k_sphere_colors[i];
      current_t = x_min;
    } else if (x_max >= 0.0f && x_max < current_t) {
      ray->intersect_triangle = -2;
      ray->intersect = add(ray->start, scale(x_max, ray->direction));
      ray->intersect_normal = normalize3(sub(ray->intersect, ctr));
      ray->intersect_color = k_sphere_colors[i];
      current_t = x_max;
    }
  }
}

/* kernels.cl:243-311 */
static int in_shadow(v3 start, v3 dir, Scene *sc, float radius_sq) {
  const v3 nd = neg(dir);
  for (int i = 0; i < sc->n; i++) {
    if (sc->colors[i].w == -1.0f) continue;
    sc->c[C_TS1]++;
    const v3 v0 = xyz(sc->verts[i * 3]);
    const v3 e1 = sub(xyz(sc->verts[i * 3 + 1]), v0);
    const v3 e2 = sub(xyz(sc->verts[i * 3 + 2]), v0);
    const v3 b = sub(start, v0);
    const float inv = 1.0f / det3(nd, e1, e2);
    const float t = det3(b, e1, e2) * inv;
    const v3 dv = scale(t, dir);
    const float dist = dv.x * dv.x + dv.y * dv.y + dv.z * dv.z;
    if (t >= 0 && dist < radius_sq) {
      sc->c[C_TS2]++;
      const float u = det3(nd, b, e2) * inv;
      const float v = det3(nd, e1, b) * inv;
      if (u >= 0 && v >= 0 && (u + v) <= 1) return 1;
    }
  }
  for (int i = 0; i < SPHERES; i++) {
    if (k_sphere_colors[i].w == -1.0f) continue;
    sc->c[C_TSPH]++;
    const v3 L = sub(start, xyz(k_sphere_centers[i]));
    const float a = dot3(dir, dir);
    const float b = 2.0f * dot3(dir, L);
    const float c = dot3(L, L) - k_sphere_radius_sqs[i];
    const float disc = b * b - 4.0f * a * c;
    if (disc < 0.0f) continue;
    const float q = (b > 0) ? -0.5f * (b + sqrtf(disc)) : -0.5f * (b - sqrtf(disc));
    const float x0 = q / a;
    const float x1 = c / q;
    const float x_min = cl_min(x0, x1);
    const float x_max = cl_max(x0, x1);
    const v3 min_dir = scale(x_min, dir);
    const v3 max_dir = scale(x_max, dir);
    const float min_dist = dot3(min_dir, min_dir);
    const float max_dist = dot3(max_dir, max_dir);
    if (x_min >= 0.0f && min_dist < radius_sq) return 1;
    else if (x_max >= 0.0f && max_dist < radius_sq) return 1;
  }
  return 0;
}

/* kernels.cl:313-340 */
static v3 direct_light(const Ray *ray, Scene *sc, v3 light_pos, v3 normal, int global_id) {
  const float light_spread = 0.05f;
  v3 total = V3(0.0f, 0.0f, 0.0f);
  /* (uint3)(global_id, global_id*91.0f, global_id*19.0f), then one xorshift */
  uint32_t rx = xorshift32((uint32_t)global_id);
  uint32_t ry = xorshift32((uint32_t)((float)global_id * 91.0f));
  uint32_t rz = xorshift32((uint32_t)((float)global_id * 19.0f));
  const v3 dir = sub(light_pos, ray->intersect);
  const v3 start = add(ray->intersect, scale(k_bias, dir));
  const float radius_sq = dir.x * dir.x + dir.y * dir.y + dir.z * dir.z;
  for (int i = 0; i < sc->S; i++) {
    rx = xorshift32(rx); ry = xorshift32(ry); rz = xorshift32(rz);
    const v3 jit = V3(crush1(rx, light_spread), crush1(ry, light_spread), crush1(rz, light_spread));
    sc->c[C_SHADOW]++;
    const float mask = (float)(!in_shadow(start, add(dir, jit), sc, radius_sq));
    const float lam = k_light_color * cl_max(dot3(dir, normal), 0.0f);
    const float den = 4.0f * ((float)M_PI) * radius_sq;
    const float term = (mask * lam) / den;
    total = add(total, V3(term, term, term));
  }
  const float fs = (float)sc->S;
  return V3(total.x / fs, total.y / fs, total.z / fs);
}

/* kernels.cl:342-365 */
static v3 secondary_light(const Ray *ray, Scene *sc, v3 light_pos, int global_id) {
  Ray pr = *ray;
  for (int b = 0; b < sc->B && pr.intersect_color.w <= 0.0f; b++) {
    pr = (pr.intersect_color.w == 0.0f) ? reflect_ray(&pr) : refract_ray(&pr);
    sc->c[C_BOUNCE]++;
    closest_hit(&pr, sc);
    if (pr.intersect_triangle != -1 && pr.intersect_color.w > 0.0f) {
      const v3 dl = direct_light(&pr, sc, light_pos, pr.intersect_normal, global_id);
      const v3 light = V3(k_indirect + dl.x, k_indirect + dl.y, k_indirect + dl.z);
      return mul(scale(0.9f, light), xyz(pr.intersect_color));
    }
  }
  return V3(0.0f, 0.0f, 0.0f);
}

/* kernels.cl:368-428 for one pixel; returns ARGB (kernels.cl:37-40) */
static uint32_t draw_pixel(int x, int y, const oracle_params *p, Scene *sc, const v4 *rot, v3 camera_pos,
                           v3 light_pos) {
  const float SW = (float)p->W, SH = (float)p->H;
  const int A = p->aa;
  const int global_id = (int)((float)y * SW + (float)x);
  v3 total = V3(0.0f, 0.0f, 0.0f);
  const v3 base_dir = V3((float)(x * A) - (SW * (float)A) / 2.0f, (float)(y * A) - (SH * (float)A) / 2.0f, p->focal);
  const v3 r0 = xyz(rot[0]), r1 = xyz(rot[1]), r2 = xyz(rot[2]);
  for (int dy = 0; dy < A; dy++) {
    for (int dx = 0; dx < A; dx++) { /* ray index dy*A+dx: same order as the reference's batch */
      Ray ray;
      memset(&ray, 0, sizeof ray);
      ray.start = camera_pos;
      const v3 d = add(base_dir, V3((float)dx, (float)dy, 0.0f));
      ray.direction = normalize3(V3(dot3(r0, d), dot3(r1, d), dot3(r2, d)));
      ray.intersect_triangle = -1;
      ray.medium = AIR;
      ray.intersect_color.w = 1.0f;
      sc->c[C_PRIMARY]++;
      closest_hit(&ray, sc);
      if (ray.intersect_triangle != -1) {
        if (ray.intersect_color.w <= 0.0f) {
          total = add(total, secondary_light(&ray, sc, light_pos, global_id));
        } else {
          const v3 fl = direct_light(&ray, sc, light_pos, ray.intersect_normal, global_id);
          total = add(total, mul(xyz(ray.intersect_color), V3(k_indirect + fl.x, k_indirect + fl.y, k_indirect + fl.z)));
        }
      }
    }
  }
  const float fa = (float)(A * A);
  const v3 c = V3(total.x / fa, total.y / fa, total.z / fa);
  const uint32_t r = (uint32_t)cl_min(cl_max(255.0f * c.x, 0.f), 255.f);
  const uint32_t g = (uint32_t)cl_min(cl_max(255.0f * c.y, 0.f), 255.f);
  const uint32_t b = (uint32_t)cl_min(cl_max(255.0f * c.z, 0.f), 255.f);
  sc->c[C_PIXELS]++;
  return (255u << 24) + (r << 16) + (g << 8) + b;
}

typedef struct {
  const oracle_params *p;
  const float *verts, *normals, *colors, *rot12;
  int n, rows, step;
  v3 camera_pos, light_pos;
  uint32_t *out;
  uint64_t *row_rays;
  int *next; /* shared row cursor */
  uint64_t c[C_COUNT];
} Job;

static void *worker(void *arg) {
  Job *j = (Job *)arg;
  Scene sc;
  sc.verts = (const v4 *)j->verts; sc.normals = (const v4 *)j->normals; sc.colors = (const v4 *)j->colors;
  sc.n = j->n; sc.S = j->p->shadow_samples; sc.B = j->p->max_bounces;
  memset(sc.c, 0, sizeof sc.c);
  for (;;) {
    const int r = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
    if (r >= j->rows) break;
    const int y = j->p->y0 + r * j->step;
    const uint64_t before = sc.c[C_PRIMARY] + sc.c[C_SHADOW] + sc.c[C_BOUNCE];
    for (int x = 0; x < j->p->W; x++)
      j->out[(size_t)y * j->p->W + x] = draw_pixel(x, y, j->p, &sc, (const v4 *)j->rot12, j->camera_pos, j->light_pos);
    if (j->row_rays) j->row_rays[y] = sc.c[C_PRIMARY] + sc.c[C_SHADOW] + sc.c[C_BOUNCE] - before;
  }
  memcpy(j->c, sc.c, sizeof sc.c);
  return 0;
}

/* Render rows y0, y0+row_step, ... < y1 into out[W*H] (index y*W+x; other rows
 * untouched).  counters (may be NULL) receives C_COUNT totals; row_rays (may
 * be NULL, length H) receives rays per rendered row.  Pixels are independent,
 * so rows are handed out to p->threads pthreads.  Returns 0. */
int oracle_render(const oracle_params *p, const float *verts, const float *normals, const float *colors, int n,
                  const float *rot12, const float *cam, const float *light, uint32_t *out, uint64_t *counters,
                  uint64_t *row_rays) {
  if (!p || p->W <= 0 || p->H <= 0 || p->aa < 1 || n < 0) return 1;
  const int step = p->row_step > 0 ? p->row_step : 1;
  const int rows = (p->y1 - p->y0 + step - 1) / step;
  int threads = p->threads;
  if (threads <= 0) threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  int next = 0;
  Job jobs[256];
  pthread_t tid[256];
  for (int t = 0; t < threads; t++) {
    Job *j = &jobs[t];
    j->p = p; j->verts = verts; j->normals = normals; j->colors = colors; j->rot12 = rot12;
    j->n = n; j->rows = rows; j->step = step;
    j->camera_pos = V3(cam[0], cam[1], cam[2]);
    j->light_pos = V3(light[0], light[1], light[2]);
    j->out = out; j->row_rays = row_rays; j->next = &next;
    memset(j->c, 0, sizeof j->c);
  }
  for (int t = 1; t < threads; t++) pthread_create(&tid[t], 0, worker, &jobs[t]);
  worker(&jobs[0]);
  for (int t = 1; t < threads; t++) pthread_join(tid[t], 0);
  if (counters) {
    memset(counters, 0, sizeof(uint64_t) * C_COUNT);
    for (int t = 0; t < threads; t++)
      for (int k = 0; k < C_COUNT; k++) counters[k] += jobs[t].c[k];
  }
  return 0;
}

/* ---- per-function entry points for unit-level parity tests ---------------- */

void oracle_xorshift3(const uint32_t *in, uint32_t *out) {
  for (int k = 0; k < 3; k++) out[k] = xorshift32(in[k]);
}

/* seed of direct_light for a pixel id (kernels.cl:319), before the loop */
void oracle_seed(int global_id, uint32_t *out) {
  out[0] = xorshift32((uint32_t)global_id);
  out[1] = xorshift32((uint32_t)((float)global_id * 91.0f));
  out[2] = xorshift32((uint32_t)((float)global_id * 19.0f));
}

void oracle_crush(const uint32_t *v, float range, float *out) {
  for (int k = 0; k < 3; k++) out[k] = crush1(v[k], range);
}

/* frame-global pixel id as the kernel computes it (kernels.cl:380) */
int oracle_global_id(int x, int y, int W) { return (int)((float)y * (float)W + (float)x); }

/* Closest hit of m rays (start/dir: m x 3 floats).  out_id[m]; out_t is not
 * stored by the reference, so the hit point (m x 3), normal (m x 3) and colour
 * (m x 4) are returned instead; rows of a miss are left untouched. */
void oracle_closest_hits(const float *start, const float *dir, int m, const float *verts, const float *normals,
                         const float *colors, int n, int *out_id, float *out_point, float *out_normal,
                         float *out_color) {
  Scene sc;
  sc.verts = (const v4 *)verts; sc.normals = (const v4 *)normals; sc.colors = (const v4 *)colors;
  sc.n = n; sc.S = 0; sc.B = 0;
  memset(sc.c, 0, sizeof sc.c);
  for (int k = 0; k < m; k++) {
    Ray r;
    memset(&r, 0, sizeof r);
    r.start = V3(start[3 * k], start[3 * k + 1], start[3 * k + 2]);
    r.direction = V3(dir[3 * k], dir[3 * k + 1], dir[3 * k + 2]);
    r.intersect_triangle = -1;
    closest_hit(&r, &sc);
    out_id[k] = r.intersect_triangle;
    if (r.intersect_triangle != -1) {
      memcpy(out_point + 3 * k, &r.intersect, 12);
      memcpy(out_normal + 3 * k, &r.intersect_normal, 12);
      memcpy(out_color + 4 * k, &r.intersect_color, 16);
    }
  }
}

void oracle_in_shadow(const float *start, const float *dir, const float *radius_sq, int m, const float *verts,
                      const float *colors, int n, int *out) {
  Scene sc;
  sc.verts = (const v4 *)verts; sc.normals = 0; sc.colors = (const v4 *)colors;
  sc.n = n; sc.S = 0; sc.B = 0;
  memset(sc.c, 0, sizeof sc.c);
  for (int k = 0; k < m; k++)
    out[k] = in_shadow(V3(start[3 * k], start[3 * k + 1], start[3 * k + 2]),
                       V3(dir[3 * k], dir[3 * k + 1], dir[3 * k + 2]), &sc, radius_sq[k]);
}

/* reflect (kind 0) / refract (kind 1): in = direction, normal, hit point, medium;
 * out = start, direction, medium of the new ray */
void oracle_bounce(int kind, const float *dir, const float *normal, const float *point, float medium,
                   float *out_start, float *out_dir, float *out_medium) {
  Ray r;
  memset(&r, 0, sizeof r);
  r.direction = V3(dir[0], dir[1], dir[2]);
  r.intersect_normal = V3(normal[0], normal[1], normal[2]);
  r.intersect = V3(point[0], point[1], point[2]);
  r.medium = medium;
  Ray o = kind ? refract_ray(&r) : reflect_ray(&r);
  memcpy(out_start, &o.start, 12);
  memcpy(out_dir, &o.direction, 12);
  *out_medium = o.medium;
}

/* ---- host-side formulas of the caller (skeleton.cpp) ---------------------- */

/* skeleton.cpp:149-151: three rows, float4 stride */
void oracle_rot_matrix(float yaw, float pitch, float *rot12) {
  const float cy = cosf(yaw), sy = sinf(yaw), cp = cosf(pitch), sp = sinf(pitch);
  const float m[12] = {cy, sp * sy, sy * cp, 0.0f, 0.0f, cp, -sp, 0.0f, -sy, cy * sp, cp * cy, 0.0f};
  memcpy(rot12, m, sizeof m);
}

/* skeleton.cpp:290-298: one step of the light ping-pong.  *lor is the
 * direction flag (starts true), *light_x the light's x (starts 0). */
void oracle_light_step(float *light_x, int *lor) {
  if (*lor) {
    const float diff = -0.5f - *light_x;
    if (diff > -0.001f) *lor = 0;
    *light_x += diff / 20.0f;
  } else {
    const float diff = 0.5f - *light_x;
    if (diff < 0.001f) *lor = 1;
    *light_x += diff / 20.0f;
  }
}
