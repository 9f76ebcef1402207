// scene_host.h -- host-side scene description of the product (loader output, before upload).
//
// Replaces the reference's pointer graph in managed memory (Scene -> Object -> Mesh -> Triangle,
// scene.h:86-173, scene_basics.h:64-194) by flat arrays in reference triangle order.
#pragma once
#include <string>
#include <vector>

namespace ipt {

struct HostMaterial {   // mat_t, scene_basics.h:35-62 (defaults: material.h:341-351)
  float kd[3] = {0, 0, 0}, ks[3] = {0, 0, 0}, kt[3] = {0, 0, 0}, ke[3] = {0, 0, 0};
  float ns = 1.0f, ni = 1.0f, d = 1.0f;
  int illum = 0;
};

struct HostTriangle {   // Triangle, scene_basics.h:64-110, plus the per-edge planes of signedDistance (:497-503)
  int idx = 0, obj = 0, idxE = -1;
  HostMaterial mat;
  float v[3][3];
  float vn[3][3];
  bool has_vn = false;
  float n[3], c[3], area = 0;
  float plane_k = 0;        // dot(center, normal)
  float edge[3][4];         // out.xyz, d   for edges (v0,v1), (v1,v2), (v2,v0)
};

struct HostObject {     // Object, scene_basics.h:402-424
  float pos[3];
  float lo[3], hi[3];   // {pos} U vertices
  int tri_begin = 0, tri_count = 0;
};

struct HostScene {
  std::vector<HostTriangle> tris;
  std::vector<HostObject> objects;
  std::vector<int> emissives;      // global triangle index per emissive (scene.h:100-114)
  std::vector<float> light_cdf;    // running sums of area_i / area_sum (path_trace.cu:44-50)
  std::vector<float> light_pt;     // area_i / area_sum
  float area_sum = 0;
  float cam[16];                   // Camera::getInverseViewMatrix(), scene.h:75-77 (row-major)
  bool any_vn = false, any_specular = false;

  void rebuild_lights();
};

// loadScene, scene.h:177-193.  `root` (may be empty) is prepended to relative asset paths.
bool build_host_scene(const float* const* poss, const float* const* oris, const float* const* scls,
                      const char* const* obj_fs, const char* const* mtl_fs, int n, const std::string& root,
                      HostScene* out, std::string* err);

// PNG I/O (replaces stb_image / stb_image_write: path_trace.cu:233, inv_scene.h:56)
bool write_png_rgb8(const char* path, const unsigned char* rgb, int w, int h, std::string* err);
bool read_png_rgb8(const char* path, std::vector<unsigned char>* rgb, int* w, int* h, std::string* err);

}  // namespace ipt
