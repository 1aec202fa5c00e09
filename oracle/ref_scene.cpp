// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Compiles the reference's own scene sources from where they lie
// (/root/reference/Source/TestModelH.h, Loader.cpp, vendored GLM 0.9.7.2) and
// exposes them through a C ABI so that the product's GLM-free scene builders
// (uob_raytracer_b200/csrc/host) can be checked bit-for-bit.  Built by
// oracle/build_ref.py into oracle/_ref/libref_scene.so.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "Loader.cpp"  // pulls in TestModelH.h (reference sources, -I/root/reference/Source)

// AoS Triangle -> the three float4 arrays of skeleton.cpp:474-484.
static int flatten(const std::vector<Triangle> &tris, float *verts, float *normals, float *colors,
                   int capacity) {
  int n = (int)tris.size();
  if (n > capacity) return -n;
  for (int i = 0; i < n; i++) {
    const Triangle &t = tris[i];
    float v[12] = {t.v0.x, t.v0.y, t.v0.z, 0.0f, t.v1.x, t.v1.y, t.v1.z, 0.0f, t.v2.x, t.v2.y, t.v2.z, 0.0f};
    memcpy(verts + 12 * (size_t)i, v, sizeof(v));
    float nn[4] = {t.normal.x, t.normal.y, t.normal.z, 0.0f};
    memcpy(normals + 4 * (size_t)i, nn, sizeof(nn));
    float c[4] = {t.color.x, t.color.y, t.color.z, t.color.w};
    memcpy(colors + 4 * (size_t)i, c, sizeof(c));
  }
  return n;
}

extern "C" {

// LoadTestModel (TestModelH.h:44-219) + flatten. Returns n (26), or -n if capacity is short.
int ref_load_test_model(float *verts, float *normals, float *colors, int capacity) {
  std::vector<Triangle> tris;
  LoadTestModel(tris);
  return flatten(tris, verts, normals, colors, capacity);
}

// load_obj (Loader.cpp:11-59) + flatten.
int ref_load_obj(const char *path, float *verts, float *normals, float *colors, int capacity) {
  std::vector<Triangle> tris = load_obj(std::string(path));
  return flatten(tris, verts, normals, colors, capacity);
}

}  // extern "C"
