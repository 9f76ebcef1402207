// Table dumper around the REFERENCE's own OBJ/MTL loader (test infrastructure).
// Compiled by oracle/Makefile with -I/root/reference; the three reference headers are included
// from where they lie (never copied).  Mirrors how scene_basics.h:207-289 (ParseFromString)
// drives LoadObj: MaterialStreamReader on the given MTL stream, triangulate = true.
//
// usage: ref_loader_dump <obj> <mtl>   ->  text on stdout:
//   V <n> / one "x y z" line per vertex (hex floats)
//   N <n> / normals
//   M <n> / per material: name | Kd Ks Ke Ns Ni d illum
//   F <n> / per triangle: v0 v1 v2 material_id
#include <cstring>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include "tiny_obj_loader.h"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::filebuf fb;
  if (!fb.open(argv[1], std::ios_base::in)) return 3;
  std::istream obj(&fb);
  std::ifstream mtl(argv[2]);
  MaterialStreamReader reader(mtl);
  attrib_t attrib;
  std::vector<shape_t> shapes;
  std::vector<material_t> materials;
  std::string warn, err;
  bool ok = LoadObj(&attrib, &shapes, &materials, &warn, &err, &obj, &reader, true, true);
  if (!ok) return 4;
  printf("V %zu\n", attrib.vertices.size() / 3);
  for (size_t i = 0; i < attrib.vertices.size() / 3; i++)
    printf("%a %a %a\n", attrib.vertices[3 * i], attrib.vertices[3 * i + 1], attrib.vertices[3 * i + 2]);
  printf("N %zu\n", attrib.normals.size() / 3);
  for (size_t i = 0; i < attrib.normals.size() / 3; i++)
    printf("%a %a %a\n", attrib.normals[3 * i], attrib.normals[3 * i + 1], attrib.normals[3 * i + 2]);
  printf("M %zu\n", materials.size());
  for (size_t i = 0; i < materials.size(); i++) {
    const material_t& m = materials[i];
    printf("%s | %a %a %a %a %a %a %a %a %a %a %a %a %d\n", m.name.c_str(), m.diffuse[0], m.diffuse[1], m.diffuse[2],
           m.specular[0], m.specular[1], m.specular[2], m.emission[0], m.emission[1], m.emission[2], m.shininess, m.ior,
           m.dissolve, m.illum);
  }
  size_t nf = 0;
  for (size_t s = 0; s < shapes.size(); s++) nf += shapes[s].mesh.material_ids.size();
  printf("F %zu\n", nf);
  for (size_t s = 0; s < shapes.size(); s++) {
    const mesh_t& me = shapes[s].mesh;
    for (size_t j = 0; j < me.material_ids.size(); j++)
      printf("%d %d %d %d\n", me.indices[3 * j].vertex_index, me.indices[3 * j + 1].vertex_index,
             me.indices[3 * j + 2].vertex_index, me.material_ids[j]);
  }
  return 0;
}
