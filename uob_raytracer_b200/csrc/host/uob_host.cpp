// uob_host.cpp — host-side scene sources, camera/light state and framebuffer
// dump (include/uob_host.h).  GLM-free restatement of the reference's host code;
// float arithmetic is written operation by operation so that the results are
// bit-identical to the reference's GLM expressions (build: -O2 -ffp-contract=off).
#include "../../../include/uob_host.h"

#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <charconv>
#include <functional>
#include <string>
#include <chrono>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct Vec4 {
  float x, y, z, w;
};

// Triangle of TestModelH.h:11-38 (v0, v1, v2, normal, color)
struct Tri {
  Vec4 v0, v1, v2, normal, color;
};

// Triangle::ComputeNormal (TestModelH.h:26-35): normalize(cross(e2, e1)) with
// GLM 0.9.7.2 semantics: cross = (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y),
// dot = (x + y) + z of the products, normalize = v * (1 / sqrt(dot)).
void compute_normal(Tri &t) {
  const float e1x = t.v1.x - t.v0.x, e1y = t.v1.y - t.v0.y, e1z = t.v1.z - t.v0.z;
  const float e2x = t.v2.x - t.v0.x, e2y = t.v2.y - t.v0.y, e2z = t.v2.z - t.v0.z;
  const float cx = e2y * e1z - e1y * e2z;
  const float cy = e2z * e1x - e1z * e2x;
  const float cz = e2x * e1y - e1x * e2y;
  const float d = (cx * cx + cy * cy) + cz * cz;
  const float inv = 1.0f / sqrtf(d);
  t.normal.x = cx * inv;
  t.normal.y = cy * inv;
  t.normal.z = cz * inv;
  t.normal.w = 1.0f;
}

// The eight corners of an axis-extruded quad (floor corners a,b,c,d lifted to
// height h give e,f,g,h) and the triangle lists of TestModelH.h:86-104 (room)
// and :129-147 / :171-189 (blocks), as corner indices a=0 ... h=7.
struct Box {
  float ax, az, bx, bz, cx, cz, dx, dz, height;
};
enum { cA, cB, cC, cD, cE, cF, cG, cH };
const int kRoomTris[10][3] = {{cC, cB, cA}, {cC, cD, cB},   // floor
                              {cA, cE, cC}, {cC, cE, cG},   // left wall
                              {cF, cB, cD}, {cH, cF, cD},   // right wall
                              {cE, cF, cG}, {cF, cH, cG},   // ceiling
                              {cG, cD, cC}, {cG, cH, cD}};  // back wall
const float kRoomColors[5][3] = {{0.25f, 0.25f, 0.25f},  // dark grey floor
                                 {0.25f, 0.0f, 0.25f},   // dark purple
                                 {0.0f, 0.25f, 0.0f},    // dark green
                                 {0.3f, 0.3f, 0.0f},     // dark yellow
                                 {0.75f, 0.75f, 0.75f}}; // white
const int kBlockTris[8][3] = {{cE, cB, cA}, {cE, cF, cB}, {cF, cD, cB}, {cF, cH, cD},
                              {cG, cE, cC}, {cE, cA, cC}, {cG, cF, cE}, {cG, cH, cF}};

Vec4 corner(const Box &b, int k) {
  const float y = (k >= cE) ? b.height : 0.0f;
  switch (k & 3) {
    case 0: return Vec4{b.ax, y, b.az, 1.0f};
    case 1: return Vec4{b.bx, y, b.bz, 1.0f};
    case 2: return Vec4{b.cx, y, b.cz, 1.0f};
    default: return Vec4{b.dx, y, b.dz, 1.0f};
  }
}

void build_test_model(std::vector<Tri> &tris) {
  const float L = 555.0f;  // TestModelH.h:70
  const Box room{L, 0, 0, 0, L, L, 0, L, L};
  const Box shortb{290, 114, 130, 65, 240, 272, 82, 225, 165};  // :116-124
  const Box tallb{423, 247, 265, 296, 472, 406, 314, 456, 330}; // :161-169
  tris.clear();
  tris.reserve(26);
  for (int i = 0; i < 10; i++) {
    const float *c = kRoomColors[i / 2];
    tris.push_back(Tri{corner(room, kRoomTris[i][0]), corner(room, kRoomTris[i][1]), corner(room, kRoomTris[i][2]),
                       Vec4{0, 0, 0, 1}, Vec4{c[0], c[1], c[2], 1.0f}});
  }
  for (int i = 0; i < 8; i++)
    tris.push_back(Tri{corner(shortb, kBlockTris[i][0]), corner(shortb, kBlockTris[i][1]), corner(shortb, kBlockTris[i][2]),
                       Vec4{0, 0, 0, 1}, Vec4{0.6f, 0.0f, 0.0f, 1.0f}});  // red
  for (int i = 0; i < 8; i++)
    tris.push_back(Tri{corner(tallb, kBlockTris[i][0]), corner(tallb, kBlockTris[i][1]), corner(tallb, kBlockTris[i][2]),
                       Vec4{0, 0, 0, 1}, Vec4{0.0f, 0.2f, 0.5f, 1.0f}});  // blue
  // Scale to [-1,1]^3, flip x and y (TestModelH.h:195-218)
  const float s = 2 / L;
  for (Tri &t : tris) {
    Vec4 *vs[3] = {&t.v0, &t.v1, &t.v2};
    for (Vec4 *v : vs) {
      v->x *= s; v->y *= s; v->z *= s;
      v->x -= 1.0f; v->y -= 1.0f; v->z -= 1.0f;
      v->x *= -1.0f;
      v->y *= -1.0f;
      v->w = 1.0f;
    }
    compute_normal(t);
  }
}

// skeleton.cpp:474-484
int flatten(const std::vector<Tri> &tris, float *verts, float *normals, float *colors, int cap) {
  const size_t n = tris.size();
  if (n > (size_t)INT_MAX) return INT_MIN;
  if ((int)n > cap || !verts || !normals || !colors) return -(int)n;
  for (size_t i = 0; i < n; i++) {
    const Tri &t = tris[i];
    const float v[12] = {t.v0.x, t.v0.y, t.v0.z, 0.0f, t.v1.x, t.v1.y, t.v1.z, 0.0f, t.v2.x, t.v2.y, t.v2.z, 0.0f};
    memcpy(verts + 12 * i, v, sizeof v);
    const float nn[4] = {t.normal.x, t.normal.y, t.normal.z, 0.0f};
    memcpy(normals + 4 * i, nn, sizeof nn);
    const float c[4] = {t.color.x, t.color.y, t.color.z, t.color.w};
    memcpy(colors + 4 * i, c, sizeof c);
  }
  return (int)n;
}

// OBJ reader with load_obj's semantics (Loader.cpp:27-56), built for meshes of a million faces.
// The reference tokenises each line with operator>>; a line whose first token is exactly "v" or
// "f" is used, everything else is skipped.  std::from_chars (and strtof/strtol for the spellings it does
// not take: a leading '+', hex floats, out-of-range values) gives the same correctly-rounded values as
// the stream extractors for well-formed numbers.  The file is mapped, cut into one chunk per
// hardware thread at line boundaries, and every chunk is tokenised in parallel into its own vertex
// and face lists; face indices are global (1-based over the whole file, as in the reference, which
// also requires a vertex to precede the faces that use it), so triangles are built in a second
// parallel pass once all vertices are known — straight into the flattened float4 arrays of
// skeleton.cpp:474-484, which is what every caller wants.
struct ObjChunk {
  std::vector<Vec4> vertices;
  std::vector<long> faces;  // 3 per face
  bool bad_face = false;
};

// a tokenised file: per-chunk face lists, all vertices, and where each chunk's vertices / faces start
struct ParsedObj {
  std::vector<ObjChunk> chunks;
  std::vector<size_t> v_before, f_before;
  Vec4 *vertices = nullptr;
  size_t n = 0;  // faces
  void clear() {
    chunks.clear();
    v_before.clear();
    f_before.clear();
    free(vertices);
    vertices = nullptr;
    n = 0;
  }
};

inline bool obj_blank(char c) { return c == ' ' || c == '\t' || c == '\r'; }

// strtof / strtol semantics on [p, eol): skip blanks, convert the longest valid prefix, return where it stopped
// (p itself, with value 0, if nothing converts).  The slow path works on a NUL-terminated copy of the token, so
// nothing ever reads past the mapped file.
template <class T, class Slow> inline const char *obj_number(const char *p, const char *eol, T &out, Slow slow) {
  while (p < eol && obj_blank(*p)) p++;
  T v{};
  const std::from_chars_result r = std::from_chars(p, eol, v);
  if (r.ec == std::errc() && !(r.ptr < eol && (*r.ptr == 'x' || *r.ptr == 'X'))) {
    out = v;
    return r.ptr;
  }
  char tok[64];
  size_t len = 0;
  while (p + len < eol && len + 1 < sizeof tok && !obj_blank(p[len])) {
    tok[len] = p[len];
    len++;
  }
  tok[len] = '\0';
  char *e = tok;
  out = slow(tok, &e);
  return p + (e - tok);
}

void parse_chunk(const char *p, const char *end, ObjChunk &out) {
  const auto slow_f = [](const char *t, char **e) { return strtof(t, e); };
  const auto slow_l = [](const char *t, char **e) { return strtol(t, e, 10); };
  while (p < end) {
    const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
    if (!eol) eol = end;
    const char *q = p;
    while (q < eol && obj_blank(*q)) q++;
    const char *tok = q;
    while (q < eol && !obj_blank(*q)) q++;
    if (q - tok == 1 && *tok == 'v') {
      float x = 0.f, y = 0.f, z = 0.f;
      q = obj_number(q, eol, x, slow_f);
      q = obj_number(q, eol, y, slow_f);
      obj_number(q, eol, z, slow_f);
      out.vertices.push_back(Vec4{1.5f * x, 1.5f * y, 1.5f * z, 1.f});  // Loader.cpp:41
    } else if (q - tok == 1 && *tok == 'f') {
      long a = 0, b = 0, c = 0;
      q = obj_number(q, eol, a, slow_l);
      q = obj_number(q, eol, b, slow_l);
      obj_number(q, eol, c, slow_l);
      out.faces.push_back(a);
      out.faces.push_back(b);
      out.faces.push_back(c);
    }
    p = eol + 1;
  }
}

// UOB_HOST_TRACE=1 prints where load_obj spends its time (stderr)
struct PhaseTimer {
  bool on = getenv("UOB_HOST_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char *what) {
    if (!on) return;
    const auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "uob_load_obj: %-18s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

// the whole file in memory: mapped read-only, or read if it cannot be mapped
struct FileBytes {
  const char *data = nullptr;
  size_t size = 0;
  bool mapped = false;
  std::string owned;
  bool open(const char *path) {
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
      ::close(fd);
      return false;
    }
    size = (size_t)st.st_size;
    if (size == 0) {
      ::close(fd);
      data = "";
      return true;
    }
    void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    if (m != MAP_FAILED) {
      madvise(m, size, MADV_SEQUENTIAL);
      data = (const char *)m;
      mapped = true;
      ::close(fd);
      return true;
    }
    owned.resize(size);
    size_t got = 0;
    while (got < size) {
      const ssize_t k = ::read(fd, &owned[got], size - got);
      if (k <= 0) break;
      got += (size_t)k;
    }
    ::close(fd);
    if (got != size) return false;
    data = owned.data();
    return true;
  }
  ~FileBytes() {
    if (mapped) munmap((void *)data, size);
  }
};

unsigned host_threads(size_t work_bytes) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > 64) nt = 64;
  if (work_bytes < (1u << 20)) nt = 1;
  return nt;
}

template <class F> void run_parallel(unsigned nt, F fn) {
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < nt; t++) pool.emplace_back(fn, t);
  fn(0u);
  for (auto &th : pool) th.join();
}

bool parse_obj(const char *path, ParsedObj &obj) {
  PhaseTimer timer;
  obj.clear();
  FileBytes file;
  if (!file.open(path)) return false;
  timer.lap("map file");
  const char *base = file.data, *end = base + file.size;
  const unsigned nt = host_threads(file.size);
  // chunk boundaries at line starts
  std::vector<const char *> cut(nt + 1, end);
  cut[0] = base;
  for (unsigned t = 1; t < nt; t++) {
    const char *p = base + file.size / nt * t;
    const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
    cut[t] = nl ? nl + 1 : end;
  }
  obj.chunks.resize(nt);
  run_parallel(nt, [&](unsigned t) { parse_chunk(cut[t], cut[t + 1], obj.chunks[t]); });
  timer.lap("tokenise");
  // vertices of the whole file, and for every chunk how many vertices precede it (a face may only
  // use vertices defined before it: Loader.cpp indexes the vector as it grows)
  obj.v_before.assign(nt + 1, 0);
  obj.f_before.assign(nt + 1, 0);
  for (unsigned t = 0; t < nt; t++) {
    obj.v_before[t + 1] = obj.v_before[t] + obj.chunks[t].vertices.size();
    obj.f_before[t + 1] = obj.f_before[t] + obj.chunks[t].faces.size() / 3;
  }
  obj.n = obj.f_before[nt];
  if (obj.n > (size_t)INT_MAX) return false;
  obj.vertices = (Vec4 *)malloc(sizeof(Vec4) * (obj.v_before[nt] ? obj.v_before[nt] : 1));
  if (!obj.vertices) return false;
  run_parallel(nt, [&](unsigned t) {
    std::vector<Vec4> &v = obj.chunks[t].vertices;
    if (!v.empty()) memcpy(obj.vertices + obj.v_before[t], v.data(), sizeof(Vec4) * v.size());
    std::vector<Vec4>().swap(v);
  });
  // face indices are checked here, so that a bad file fails before the caller allocates anything
  run_parallel(nt, [&](unsigned t) {
    ObjChunk &ch = obj.chunks[t];
    // vertices visible to the faces of this chunk: all of the earlier chunks; inside the chunk the
    // interleaving of v and f lines is not tracked, so (conservatively, like the usual OBJ layout) the
    // chunk's own vertices count as well — an index beyond them is an error as in the reference
    const long nv = (long)obj.v_before[t + 1];
    for (long idx : ch.faces)
      if (idx < 1 || idx > nv) {
        ch.bad_face = true;
        return;
      }
  });
  timer.lap("gather + check");
  for (unsigned t = 0; t < nt; t++)
    if (obj.chunks[t].bad_face) return false;
  return true;
}

// Triangles of a tokenised file, flattened as skeleton.cpp:474-484 does, written by every host thread straight
// into the caller's arrays (105 MB at a million faces).
void build_triangles(const ParsedObj &obj, float *verts, float *normals, float *colors) {
  const Vec4 blue{0.0f, 0.2f, 0.4f, 0.5f};         // Loader.cpp:20
  const Vec4 translate{-0.4f, 1.15f, -0.7f, 1.0f}; // Loader.cpp:48
  const unsigned nt = (unsigned)obj.chunks.size();
  run_parallel(nt, [&](unsigned t) {
    const ObjChunk &ch = obj.chunks[t];
    for (size_t k = 0; k < ch.faces.size() / 3; k++) {
      const long a = ch.faces[3 * k], b = ch.faces[3 * k + 1], c = ch.faces[3 * k + 2];
      Tri tr{obj.vertices[(size_t)a - 1], obj.vertices[(size_t)b - 1], obj.vertices[(size_t)c - 1], Vec4{0, 0, 0, 1}, blue};
      compute_normal(tr);  // from the scaled, untransformed vertices; kept as is
      const Vec4 *vs[3] = {&tr.v0, &tr.v1, &tr.v2};
      const size_t i = obj.f_before[t] + k;
      float *o = verts + 12 * i;  // xyz of the three vertices after Loader.cpp:44-53, w = 0
      for (const Vec4 *v : vs) {
        o[0] = (-1.f) * v->x + translate.x;
        o[1] = (-1.f) * v->y + translate.y;
        o[2] = (-1.f) * v->z + translate.z;
        o[3] = 0.0f;
        o += 4;
      }
      const float nn[4] = {tr.normal.x, tr.normal.y, tr.normal.z, 0.0f};
      memcpy(normals + 4 * i, nn, sizeof nn);
      const float cc[4] = {blue.x, blue.y, blue.z, blue.w};
      memcpy(colors + 4 * i, cc, sizeof cc);
    }
  });
}

void put_le32(unsigned char *p, uint32_t v) {
  p[0] = (unsigned char)v; p[1] = (unsigned char)(v >> 8); p[2] = (unsigned char)(v >> 16); p[3] = (unsigned char)(v >> 24);
}

uint32_t hash_u32(uint32_t x) {  // deterministic noise for the synthetic mesh
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

}  // namespace

extern "C" {

int uob_test_model_count(void) { return 26; }

int uob_load_test_model(float *verts, float *normals, float *colors, int cap) {
  std::vector<Tri> tris;
  build_test_model(tris);
  return flatten(tris, verts, normals, colors, cap);
}

int uob_load_obj(const char *path, float *verts, float *normals, float *colors, int cap) {
  // the two-pass use (cap = 0 to size the buffers, then the real call) reads and tokenises the file once
  static thread_local std::string cached_path;
  static thread_local ParsedObj cached;
  if (!path) return INT_MIN;
  if (cached_path != path || cap == 0) {
    cached.clear();
    cached_path.clear();
    if (!parse_obj(path, cached)) {
      cached.clear();
      return INT_MIN;
    }
    cached_path = path;
  }
  if ((size_t)(cap < 0 ? 0 : cap) < cached.n || !verts || !normals || !colors) return -(int)cached.n;
  PhaseTimer timer;
  build_triangles(cached, verts, normals, colors);
  timer.lap("build triangles");
  const int n = (int)cached.n;
  cached.clear();  // delivered: drop the tokens
  cached_path.clear();
  return n;
}

void uob_rot_matrix(float yaw, float pitch, float rot12[12]) {
  const float cy = cosf(yaw), sy = sinf(yaw), cp = cosf(pitch), sp = sinf(pitch);
  const float m[12] = {cy, sp * sy, sy * cp, 0.0f, 0.0f, cp, -sp, 0.0f, -sy, cy * sp, cp * cy, 0.0f};
  memcpy(rot12, m, sizeof m);
}

void uob_light_step(float *light_x, int *lor) {
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

void uob_default_camera(float *focal, float cam[4], float light[4]) {
  if (focal) *focal = 2200.0f;
  if (cam) { cam[0] = 0.0f; cam[1] = 0.0f; cam[2] = -3.2f; cam[3] = 1.0f; }
  if (light) { light[0] = 0.0f; light[1] = -0.5f; light[2] = -0.7f; light[3] = 1.0f; }
}

float uob_fitted_focal(int aa, int height) { return 1100.0f * (float)aa * (float)height / 1024.0f; }

int uob_save_bmp(const char *path, const uint32_t *argb, int width, int height) {
  if (!path || !argb || width <= 0 || height <= 0) return 1;
  FILE *f = fopen(path, "wb");
  if (!f) return 2;
  const int stride = (width * 3 + 3) & ~3;
  unsigned char hdr[54];
  memset(hdr, 0, sizeof hdr);
  hdr[0] = 'B'; hdr[1] = 'M';
  put_le32(hdr + 2, 54u + (uint32_t)stride * (uint32_t)height);
  put_le32(hdr + 10, 54);
  put_le32(hdr + 14, 40);
  put_le32(hdr + 18, (uint32_t)width);
  put_le32(hdr + 22, (uint32_t)height);
  hdr[26] = 1; hdr[28] = 24;
  put_le32(hdr + 34, (uint32_t)stride * (uint32_t)height);
  fwrite(hdr, 1, sizeof hdr, f);
  std::vector<unsigned char> row((size_t)stride, 0);
  for (int y = height - 1; y >= 0; y--) {
    for (int x = 0; x < width; x++) {
      const uint32_t p = argb[(size_t)y * width + x];
      row[3 * x + 0] = (unsigned char)(p & 255);
      row[3 * x + 1] = (unsigned char)((p >> 8) & 255);
      row[3 * x + 2] = (unsigned char)((p >> 16) & 255);
    }
    fwrite(row.data(), 1, row.size(), f);
  }
  return fclose(f) == 0 ? 0 : 3;
}

int uob_save_ppm(const char *path, const uint32_t *argb, int width, int height) {
  if (!path || !argb || width <= 0 || height <= 0) return 1;
  FILE *f = fopen(path, "wb");
  if (!f) return 2;
  fprintf(f, "P6\n%d %d\n255\n", width, height);
  std::vector<unsigned char> row((size_t)width * 3);
  for (int y = 0; y < height; y++) {
    for (int x = 0; x < width; x++) {
      const uint32_t p = argb[(size_t)y * width + x];
      row[3 * x + 0] = (unsigned char)((p >> 16) & 255);
      row[3 * x + 1] = (unsigned char)((p >> 8) & 255);
      row[3 * x + 2] = (unsigned char)(p & 255);
    }
    fwrite(row.data(), 1, row.size(), f);
  }
  return fclose(f) == 0 ? 0 : 3;
}

int uob_write_icosphere_obj(const char *path, int subdiv, float radius, float noise) {
  if (!path || subdiv < 0 || subdiv > 9) return -1;
  struct P { double x, y, z; };
  std::vector<P> v;
  std::vector<int> f;
  const double t = (1.0 + sqrt(5.0)) / 2.0;
  const double iv[12][3] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
                            {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
  const int it[20][3] = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
                         {11, 10, 2}, {10, 7, 6}, {7, 1, 8}, {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8},
                         {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
  auto norm = [](P p) { const double l = sqrt(p.x * p.x + p.y * p.y + p.z * p.z); return P{p.x / l, p.y / l, p.z / l}; };
  for (auto &p : iv) v.push_back(norm(P{p[0], p[1], p[2]}));
  for (auto &tr : it) { f.push_back(tr[0]); f.push_back(tr[1]); f.push_back(tr[2]); }
  for (int s = 0; s < subdiv; s++) {
    std::unordered_map<uint64_t, int> mid;
    mid.reserve(f.size());
    auto midpoint = [&](int a, int b) {
      const uint64_t key = a < b ? ((uint64_t)a << 32) | (uint32_t)b : ((uint64_t)b << 32) | (uint32_t)a;
      auto it2 = mid.find(key);
      if (it2 != mid.end()) return it2->second;
      const P m = norm(P{(v[a].x + v[b].x) / 2, (v[a].y + v[b].y) / 2, (v[a].z + v[b].z) / 2});
      v.push_back(m);
      const int idx = (int)v.size() - 1;
      mid.emplace(key, idx);
      return idx;
    };
    std::vector<int> nf;
    nf.reserve(f.size() * 4);
    for (size_t i = 0; i < f.size(); i += 3) {
      const int a = f[i], b = f[i + 1], c = f[i + 2];
      const int ab = midpoint(a, b), bc = midpoint(b, c), ca = midpoint(c, a);
      const int q[12] = {a, ab, ca, b, bc, ab, c, ca, bc, ab, bc, ca};
      nf.insert(nf.end(), q, q + 12);
    }
    f.swap(nf);
  }
  FILE *out = fopen(path, "wb");
  if (!out) return -2;
  fprintf(out, "# icosphere subdiv=%d radius=%g noise=%g\n", subdiv, (double)radius, (double)noise);
  for (size_t i = 0; i < v.size(); i++) {
    const double u = (double)(hash_u32((uint32_t)i * 2654435761u + 12345u) >> 8) / 16777216.0;  // [0,1)
    const double r = (double)radius * (1.0 + (double)noise * (2.0 * u - 1.0));
    fprintf(out, "v %.9g %.9g %.9g\n", (double)(float)(v[i].x * r), (double)(float)(v[i].y * r), (double)(float)(v[i].z * r));
  }
  for (size_t i = 0; i < f.size(); i += 3) fprintf(out, "f %d %d %d\n", f[i] + 1, f[i + 1] + 1, f[i + 2] + 1);
  if (fclose(out) != 0) return -3;
  return (int)(f.size() / 3);
}

}  // extern "C"
