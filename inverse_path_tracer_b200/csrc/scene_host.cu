// scene_host.cu -- asset loading and per-triangle precomputation (host only).
//
// Follows the behaviour of the reference loader path
//   loadScene (scene.h:177-193) -> Scene ctor (scene.h:89-117) -> Object ctor (scene_basics.h:405-413)
//   -> Mesh ctor / ParseFromString (scene_basics.h:147-194, 207-289) -> LoadObj / LoadMtl
//   (tiny_obj_loader.h:585-933, material.h:380-780, number grammar utils.h:70-200)
// with its own scanner-style parser.  Triangle order == face order of the OBJ file with quads split
// by the shorter-diagonal rule (tiny_obj_loader.h:257-300), which defines the global triangle ids.
#include "scene_host.h"

#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

#include "ipt_math.cuh"

namespace ipt {

namespace {

// ---------------------------------------------------------------------------------------------
// Line scanner.  Numbers follow utils.h:70-200: sign, integer digits folded into a double
// (m = m*10 + d), fraction digits added as d * 10^-k (table for k < 8, pow() after), optional
// exponent applied as ldexp(m * 5^e, e); the double is then narrowed to float.
// ---------------------------------------------------------------------------------------------
class Scanner {
 public:
  explicit Scanner(const std::string& line) : cur_(line.c_str()) { skip_blanks(); }
  bool at_end() const { return *cur_ == '\0' || *cur_ == '\r' || *cur_ == '\n'; }
  char peek(int k = 0) const { return cur_[k]; }
  void advance(int k) { cur_ += k; }
  void skip_blanks() { while (*cur_ == ' ' || *cur_ == '\t') cur_++; }
  bool prefix(const char* kw) const { return strncmp(cur_, kw, strlen(kw)) == 0; }
  bool keyword(const char* kw) {   // keyword followed by a blank
    size_t n = strlen(kw);
    if (strncmp(cur_, kw, n) == 0 && (cur_[n] == ' ' || cur_[n] == '\t')) { cur_ += n; return true; }
    return false;
  }
  std::string word() {
    skip_blanks();
    const char* b = cur_;
    while (*cur_ && *cur_ != ' ' && *cur_ != '\t' && *cur_ != '\r') cur_++;
    return std::string(b, cur_);
  }
  std::string rest() const { return std::string(cur_); }
  float real() {
    skip_blanks();
    const char* b = cur_;
    while (*cur_ && *cur_ != ' ' && *cur_ != '\t' && *cur_ != '\r') cur_++;
    double v = 0.0;
    number(b, cur_, &v);
    return (float)v;
  }
  void real3(float* o) { o[0] = real(); o[1] = real(); o[2] = real(); }
  // first integer of a "v/vt/vn" group, then skip the group
  int face_index() {
    int idx = atoi(cur_);
    while (*cur_ && *cur_ != ' ' && *cur_ != '\t' && *cur_ != '\r') cur_++;
    return idx;
  }

  static bool number(const char* s, const char* e, double* out) {
    if (s >= e) return false;
    const char* p = s;
    bool negative = false, dot_first = false;
    if (*p == '+' || *p == '-') {
      negative = (*p == '-');
      p++;
      dot_first = (p != e && *p == '.');
    } else if (*p == '.') {
      dot_first = true;
    } else if (!isdigit((unsigned char)*p)) {
      return false;
    }
    double m = 0.0;
    if (!dot_first) {
      int nd = 0;
      while (p != e && isdigit((unsigned char)*p)) { m *= 10; m += (int)(*p - '0'); p++; nd++; }
      if (nd == 0) return false;
    }
    int ex = 0;
    bool want_exp = false;
    if (p != e && *p == '.') {
      p++;
      static const double tenth[8] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
      int k = 1;
      while (p != e && isdigit((unsigned char)*p)) {
        m += (int)(*p - '0') * (k < 8 ? tenth[k] : std::pow(10.0, -k));
        k++; p++;
      }
      want_exp = (p != e && (*p == 'e' || *p == 'E'));
    } else if (p != e && (*p == 'e' || *p == 'E')) {
      want_exp = true;
    }
    if (want_exp) {
      p++;
      bool eneg = false;
      if (p != e && (*p == '+' || *p == '-')) { eneg = (*p == '-'); p++; }
      else if (p == e || !isdigit((unsigned char)*p)) return false;
      int nd = 0;
      while (p != e && isdigit((unsigned char)*p)) {
        if (ex > 2147483647 / 10) return false;
        ex = ex * 10 + (int)(*p - '0');
        p++; nd++;
      }
      if (nd == 0) return false;
      if (eneg) ex = -ex;
    }
    double mag = ex ? std::ldexp(m * std::pow(5.0, ex), ex) : m;
    *out = negative ? -mag : mag;
    return true;
  }

 private:
  const char* cur_;
};

void chomp(std::string& s) {
  while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back();
}

struct MtlLibrary {
  std::vector<HostMaterial> mats;
  std::map<std::string, int> by_name;   // first definition wins (std::map::insert semantics)
};

// material.h:380-780, restricted to the fields mat_t copies (scene_basics.h:50-59)
void read_mtl(std::istream& in, MtlLibrary* lib) {
  HostMaterial cur;
  std::string cur_name;
  bool seen_d = false;
  bool seen_kd = false;   // like the reference (material.h:396-398, 447-448) this one is NOT reset by newmtl
  std::string line;
  auto flush = [&]() {
    lib->by_name.insert(std::make_pair(cur_name, (int)lib->mats.size()));
    lib->mats.push_back(cur);
  };
  while (std::getline(in, line)) {
    size_t last = line.find_last_not_of(" \t\r\n");
    if (last == std::string::npos) continue;
    line.resize(last + 1);
    Scanner sc(line);
    if (sc.at_end() || sc.peek() == '#') continue;
    if (sc.keyword("newmtl")) {
      if (!cur_name.empty()) flush();
      cur = HostMaterial();
      seen_d = false;
      sc.advance(1);
      cur_name = sc.rest();
    } else if (sc.keyword("Kd")) { sc.real3(cur.kd); seen_kd = true;
    } else if (sc.keyword("map_Kd")) { if (!seen_kd) cur.kd[0] = cur.kd[1] = cur.kd[2] = 0.6f;   // material.h:630-641
    } else if (sc.keyword("Ks")) { sc.real3(cur.ks);
    } else if (sc.keyword("Ke")) { sc.real3(cur.ke);
    } else if (sc.keyword("Kt") || sc.keyword("Tf")) { sc.real3(cur.kt);
    } else if (sc.keyword("Ni")) { cur.ni = sc.real();
    } else if (sc.keyword("Ns")) { cur.ns = sc.real();
    } else if (sc.keyword("illum")) { sc.skip_blanks(); cur.illum = atoi(sc.rest().c_str());
    } else if (sc.keyword("d")) { cur.d = sc.real(); seen_d = true;
    } else if (sc.keyword("Tr")) { if (!seen_d) cur.d = 1.0f - sc.real();
    }
  }
  flush();   // the last material is always emitted (material.h end of LoadMtl)
}

struct ObjMesh {
  std::vector<float> v, vn;
  std::vector<int> tri_v;     // 3 per triangle
  std::vector<int> tri_mat;   // material id or -1
  MtlLibrary lib;
};

// ---- polygons -> triangles ---------------------------------------------------------------------------
// What exportGroupsToShape does with triangulate = true (tiny_obj_loader.h:180-566), including the parts
// that only matter for odd input: faces wait in their group until the group is flushed (`g`, `o`, a usemtl
// that changes the material, end of file) and are range-checked against the vertex array of THAT moment;
// quads split along the shorter diagonal; polygons with more corners are ear-clipped in a 2-D projection.
class FaceGroup {
 public:
  void add(const std::vector<int>& corners, int material) {
    if (corners.size() < 3) return;
    first_.push_back((int)idx_.size());
    count_.push_back((int)corners.size());
    mat_.push_back(material);
    idx_.insert(idx_.end(), corners.begin(), corners.end());
  }
  void flush(const std::vector<float>& v, ObjMesh* out) {
    for (size_t f = 0; f < first_.size(); f++) emit_face(&idx_[first_[f]], count_[f], mat_[f], v, out);
    first_.clear(); count_.clear(); mat_.clear(); idx_.clear();
  }

 private:
  std::vector<int> first_, count_, mat_, idx_;

  static bool readable(const std::vector<float>& v, int vertex, size_t component) {
    return (size_t)vertex * 3 + component < v.size();   // a negative index wraps and fails
  }
  static void tri(ObjMesh* out, int a, int b, int c, int material) {
    out->tri_v.push_back(a); out->tri_v.push_back(b); out->tri_v.push_back(c);
    out->tri_mat.push_back(material);
  }
  // crossing-number test against the candidate ear (material.h:369-381)
  static bool inside_ear(const float x[3], const float y[3], float qx, float qy) {
    int crossings = 0;
    int j = 2;
    for (int i = 0; i < 3; j = i, i++) {
      const bool straddles = (y[i] > qy) != (y[j] > qy);
      if (straddles && qx < (x[j] - x[i]) * (qy - y[i]) / (y[j] - y[i]) + x[i]) crossings ^= 1;
    }
    return crossings != 0;
  }
  static void emit_face(const int* c, int n, int material, const std::vector<float>& v, ObjMesh* out) {
    if (n == 3) { tri(out, c[0], c[1], c[2], material); return; }
    if (n == 4) {
      for (int k = 0; k < 4; k++) if (!readable(v, c[k], 2)) return;
      const float* a = &v[3 * (size_t)c[0]]; const float* b = &v[3 * (size_t)c[1]];
      const float* cc = &v[3 * (size_t)c[2]]; const float* d = &v[3 * (size_t)c[3]];
      // shorter diagonal, plain float arithmetic in the reference's order (tiny_obj_loader.h:257-265)
      float acx = cc[0] - a[0], acy = cc[1] - a[1], acz = cc[2] - a[2];
      float bdx = d[0] - b[0], bdy = d[1] - b[1], bdz = d[2] - b[2];
      float len_ac = acx * acx + acy * acy + acz * acz;
      float len_bd = bdx * bdx + bdy * bdy + bdz * bdz;
      if (len_ac < len_bd) { tri(out, c[0], c[1], c[2], material); tri(out, c[0], c[2], c[3], material); }
      else { tri(out, c[0], c[1], c[3], material); tri(out, c[1], c[2], c[3], material); }
      return;
    }
    // projection plane from the first corner that is not degenerate (tiny_obj_loader.h:314-361)
    size_t ua = 1, va = 2;
    for (int k = 0; k < n; k++) {
      const int i0 = c[k], i1 = c[(k + 1) % n], i2 = c[(k + 2) % n];
      if (!(readable(v, i0, 2) && readable(v, i1, 2) && readable(v, i2, 2))) continue;
      const float* p0 = &v[3 * (size_t)i0]; const float* p1 = &v[3 * (size_t)i1]; const float* p2 = &v[3 * (size_t)i2];
      const float ax = p1[0] - p0[0], ay = p1[1] - p0[1], az = p1[2] - p0[2];
      const float bx = p2[0] - p1[0], by = p2[1] - p1[1], bz = p2[2] - p1[2];
      const float nx = fabsf(ay * bz - az * by), ny = fabsf(az * bx - ax * bz), nz = fabsf(ax * by - ay * bx);
      const float tiny = 1.1920929e-07f;   // numeric_limits<float>::epsilon()
      if (nx > tiny || ny > tiny || nz > tiny) {
        const bool x_dominates = nx > ny && nx > nz;
        if (!x_dominates) {
          ua = 0;
          if (nz > nx && nz > ny) va = 1;
        }
        break;
      }
    }
    // ear clipping (tiny_obj_loader.h:421-564): `at` walks the ring; a full round without an ear ends the face
    // and whatever is left of it is dropped
    std::vector<int> ring(c, c + n);
    size_t at = 0, tries_left = (size_t)n, last_size = (size_t)n;
    while (ring.size() > 3 && tries_left > 0) {
      const size_t m = ring.size();
      if (at >= m) at -= m;
      if (last_size != m) { last_size = m; tries_left = m; } else { tries_left--; }
      int ear[3];
      float x[3], y[3];
      for (int k = 0; k < 3; k++) {
        ear[k] = ring[(at + k) % m];
        const bool ok = readable(v, ear[k], ua) && readable(v, ear[k], va);
        x[k] = ok ? v[(size_t)ear[k] * 3 + ua] : 0.0f;
        y[k] = ok ? v[(size_t)ear[k] * 3 + va] : 0.0f;
      }
      const float turn = (x[1] - x[0]) * (y[2] - y[1]) - (y[1] - y[0]) * (x[2] - x[1]);
      const float signed_half = (x[0] * y[1] - y[0] * x[1]) * 0.5f;   // the reference's "area": corners 0 and 1 only
      bool is_ear = !(turn * signed_half < 0.0f);
      for (size_t o = 3; is_ear && o < m; o++) {
        const int q = ring[(at + o) % m];
        if (!(readable(v, q, ua) && readable(v, q, va))) continue;
        if (inside_ear(x, y, v[(size_t)q * 3 + ua], v[(size_t)q * 3 + va])) is_ear = false;
      }
      if (!is_ear) { at++; continue; }
      tri(out, ear[0], ear[1], ear[2], material);
      ring.erase(ring.begin() + (std::ptrdiff_t)((at + 1) % m));
    }
    if (ring.size() == 3) tri(out, ring[0], ring[1], ring[2], material);
  }
};

bool read_obj(const std::string& path, const std::string& mtl_arg, bool mtl_is_inline, ObjMesh* out, std::string* err) {
  std::ifstream in(path.c_str());
  if (!in) { *err = "Object File was not able to be opened: '" + path + "' (" + strerror(errno) + ")"; return false; }
  int active = -1;
  bool mtl_read = false;
  std::vector<int> poly;
  FaceGroup group;
  std::string line;
  size_t line_no = 0;
  while (std::getline(in, line)) {
    line_no++;
    chomp(line);
    Scanner sc(line);
    if (sc.at_end() || sc.peek() == '#') continue;
    if (sc.keyword("v")) {
      float p[3]; sc.real3(p);
      out->v.insert(out->v.end(), p, p + 3);
    } else if (sc.keyword("vn")) {
      float p[3]; sc.real3(p);
      out->vn.insert(out->vn.end(), p, p + 3);
    } else if (sc.keyword("f")) {
      poly.clear();
      sc.skip_blanks();
      const int nv = (int)(out->v.size() / 3);
      while (!sc.at_end()) {
        int raw = sc.face_index();
        if (raw == 0) { std::ostringstream os; os << "Failed parse `f' line (zero face index), " << path << ":" << line_no; *err = os.str(); return false; }
        poly.push_back(raw > 0 ? raw - 1 : nv + raw);
        sc.skip_blanks();
      }
      group.add(poly, active);
    } else if (sc.prefix("usemtl")) {
      sc.advance(6);
      std::string name = sc.word();
      std::map<std::string, int>::const_iterator it = out->lib.by_name.find(name);
      const int next = (it == out->lib.by_name.end()) ? -1 : it->second;
      if (next != active) group.flush(out->v, out);
      active = next;
    } else if (sc.keyword("g") || sc.keyword("o")) {
      group.flush(out->v, out);
    } else if (sc.keyword("mtllib")) {
      // the file name on this line is ignored: the reference always reads the MTL stream it was
      // handed (MaterialStreamReader, material.h:784-809 <- scene_basics.h:224-225)
      if (!mtl_is_inline && !mtl_read) {
        std::ifstream mtl(mtl_arg.c_str());
        if (mtl) read_mtl(mtl, &out->lib);
        mtl_read = true;
      }
    }
  }
  group.flush(out->v, out);
  return true;
}

struct Xform {
  float L[3][3];
  float t[3];
};

// T = Translate(pos) * AngleAxis(|ori|, ori/|ori|) * Scale(scl)   (scene_basics.h:149-157)
Xform make_xform(const float* pos, const float* ori, const float* scl) {
  f3 o = mk3(ori[0], ori[1], ori[2]);
  float angle = norm3(o);
  f3 ax = normalize3(o);
  float s = sinf(angle), c = cosf(angle);
  f3 sa = ax * s;
  f3 ca = ax * (1.0f - c);
  float R[3][3];
  float xy = ca.x * ax.y, xz = ca.x * ax.z, yz = ca.y * ax.z;
  R[0][0] = fmaf(ca.x, ax.x, c); R[0][1] = xy - sa.z;           R[0][2] = xz + sa.y;
  R[1][0] = xy + sa.z;           R[1][1] = fmaf(ca.y, ax.y, c); R[1][2] = yz - sa.x;
  R[2][0] = xz - sa.y;           R[2][1] = yz + sa.x;           R[2][2] = fmaf(ca.z, ax.z, c);
  Xform T;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T.L[i][j] = R[i][j] * scl[j];
    T.t[i] = pos[i];
  }
  return T;
}

f3 apply_point(const Xform& T, f3 v) {
  return mk3(dot3(T.L[0][0], T.L[0][1], T.L[0][2], v) + T.t[0], dot3(T.L[1][0], T.L[1][1], T.L[1][2], v) + T.t[1],
             dot3(T.L[2][0], T.L[2][1], T.L[2][2], v) + T.t[2]);
}

// (L^T)^-1 by cofactors in double, rounded to float (scene_basics.h:240)
void inverse_transpose(const Xform& T, float N[3][3]) {
  double m[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = T.L[j][i];
  double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
  double c01 = m[1][0] * m[2][2] - m[1][2] * m[2][0];
  double c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
  double det = m[0][0] * c00 - m[0][1] * c01 + m[0][2] * c02;
  double inv[3][3] = {
      {c00 / det, (m[0][2] * m[2][1] - m[0][1] * m[2][2]) / det, (m[0][1] * m[1][2] - m[0][2] * m[1][1]) / det},
      {(m[1][2] * m[2][0] - m[1][0] * m[2][2]) / det, (m[0][0] * m[2][2] - m[0][2] * m[2][0]) / det, (m[0][2] * m[1][0] - m[0][0] * m[1][2]) / det},
      {c02 / det, (m[0][1] * m[2][0] - m[0][0] * m[2][1]) / det, (m[0][0] * m[1][1] - m[0][1] * m[1][0]) / det}};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) N[i][j] = (float)inv[i][j];
}

// Triangle ctor (scene_basics.h:74-96) + the static part of Object::signedDistance (:497-503)
void finish_triangle(HostTriangle* t) {
  f3 v0 = mk3(t->v[0][0], t->v[0][1], t->v[0][2]);
  f3 v1 = mk3(t->v[1][0], t->v[1][1], t->v[1][2]);
  f3 v2 = mk3(t->v[2][0], t->v[2][1], t->v[2][2]);
  f3 c = mk3(0.0f, 0.0f, 0.0f);
  c = c + mk3(v0.x / 3.0f, v0.y / 3.0f, v0.z / 3.0f);
  c = c + mk3(v1.x / 3.0f, v1.y / 3.0f, v1.z / 3.0f);
  c = c + mk3(v2.x / 3.0f, v2.y / 3.0f, v2.z / 3.0f);
  f3 n = cross3(v1 - v0, v2 - v1);
  t->area = norm3(n) / 2.0f;
  n = normalize3(n);
  t->n[0] = n.x; t->n[1] = n.y; t->n[2] = n.z;
  t->c[0] = c.x; t->c[1] = c.y; t->c[2] = c.z;
  t->plane_k = dot3(c, n);
  const f3 vs[3] = {v0, v1, v2};
  for (int j = 0; j < 3; j++) {
    f3 s0 = vs[j], s1 = vs[(j + 1) % 3];
    f3 out = normalize3(cross3(s1 - s0, n));
    float d = (-dot3(out, s1 + s0)) / 2.0f;
    t->edge[j][0] = out.x; t->edge[j][1] = out.y; t->edge[j][2] = out.z; t->edge[j][3] = d;
  }
  if (!t->has_vn)
    for (int j = 0; j < 3; j++) { t->vn[j][0] = n.x; t->vn[j][1] = n.y; t->vn[j][2] = n.z; }
}

// Camera (scene.h:35-84) with the default parameters of scene.h:3-7
void default_camera(float M[16]) {
  f3 pos = mk3(0.0f, 0.0f, 0.0f);
  f3 f = normalize3(normalize3(mk3(0.0f, 0.0f, 1.0f)));
  f3 up = normalize3(mk3(0.0f, 1.0f, 0.0f));
  f3 s = normalize3(cross3(f, up));
  f3 u = normalize3(cross3(s, f));
  float ha = (float)(M_PI * (double)90.0f / 360.0);
  float view[4][4] = {{s.x, s.y, s.z, -dot3(s, pos)}, {u.x, u.y, u.z, -dot3(u, pos)}, {f.x, f.y, f.z, -dot3(f, pos)}, {0, 0, 0, 1}};
  float sc[4] = {tanf(ha), tanf(ha * 1.0f), 1.0f, 1.0f};
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) M[4 * i + j] = sc[i] * view[j][i];
}

}  // namespace

void HostScene::rebuild_lights() {
  emissives.clear();
  for (size_t i = 0; i < tris.size(); i++) {
    const float* e = tris[i].mat.ke;
    tris[i].idxE = -1;
    if (e[0] > 0.0f || e[1] > 0.0f || e[2] > 0.0f) {   // scene_basics.h:183-187
      tris[i].idxE = (int)emissives.size();
      emissives.push_back((int)i);
    }
  }
  area_sum = 0.0f;
  for (size_t k = 0; k < emissives.size(); k++) area_sum += tris[emissives[k]].area;
  light_cdf.clear();
  light_pt.clear();
  float run = 0.0f;
  for (size_t k = 0; k < emissives.size(); k++) {
    float p = tris[emissives[k]].area / area_sum;
    run += p;
    light_cdf.push_back(run);
    light_pt.push_back(p);
  }
}

bool build_host_scene(const float* const* poss, const float* const* oris, const float* const* scls,
                      const char* const* obj_fs, const char* const* mtl_fs, int n, const std::string& root,
                      HostScene* out, std::string* err) {
  auto locate = [&](const std::string& p) {
    if (root.empty() || p.empty() || p[0] == '/') return p;
    return root + "/" + p;
  };
  out->tris.clear();
  out->objects.clear();
  for (int o = 0; o < n; o++) {
    if (!poss[o] || !oris[o] || !scls[o] || !obj_fs[o] || !mtl_fs[o]) { *err = "loadScene: null object parameter"; return false; }
    const std::string mtl_arg = mtl_fs[o];
    const bool inline_mtl = !mtl_arg.empty() && mtl_arg[0] == '*';   // scene_basics.h:223
    ObjMesh mesh;
    if (!read_obj(locate(obj_fs[o]), inline_mtl ? mtl_arg : locate(mtl_arg), inline_mtl, &mesh, err)) return false;

    // inline material string "*Kd r g b*": only Kd is honoured (scene_basics.h:249-268)
    HostMaterial fallback;
    if (inline_mtl && mtl_arg.size() >= 2) {
      std::istringstream ss(mtl_arg.substr(1, mtl_arg.size() - 2));
      std::string ln;
      while (std::getline(ss, ln, '\n')) {
        if (ln.size() >= 3 && ln[0] == 'K' && (ln[2] == ' ' || ln[2] == '\t')) {
          char which = ln[1];
          Scanner sc(ln);
          sc.advance(2);
          float rgb[3];
          sc.real3(rgb);
          if (which == 'd') memcpy(fallback.kd, rgb, sizeof(rgb));
        }
      }
    }

    Xform T = make_xform(poss[o], oris[o], scls[o]);
    float N[3][3];
    inverse_transpose(T, N);
    std::vector<f3> wv, wn;
    for (size_t i = 0; i + 2 < mesh.v.size(); i += 3) wv.push_back(apply_point(T, mk3(mesh.v[i], mesh.v[i + 1], mesh.v[i + 2])));
    for (size_t i = 0; i + 2 < mesh.vn.size(); i += 3) {
      f3 q = mk3(mesh.vn[i], mesh.vn[i + 1], mesh.vn[i + 2]);
      wn.push_back(mk3(dot3(N[0][0], N[0][1], N[0][2], q), dot3(N[1][0], N[1][1], N[1][2], q), dot3(N[2][0], N[2][1], N[2][2], q)));
    }
    const bool use_vn = (wn.size() == wv.size());   // scene_basics.h:176: normals are indexed by VERTEX index

    HostObject ho;
    for (int k = 0; k < 3; k++) ho.pos[k] = ho.lo[k] = ho.hi[k] = poss[o][k];
    ho.tri_begin = (int)out->tris.size();
    const size_t nF = mesh.tri_mat.size();
    for (size_t fi = 0; fi < nF; fi++) {
      HostTriangle t;
      t.idx = (int)out->tris.size();
      t.obj = o;
      int mid = mesh.tri_mat[fi];
      t.mat = (mid >= 0) ? mesh.lib.mats[mid] : fallback;
      t.has_vn = use_vn;
      for (int j = 0; j < 3; j++) {
        int vi = mesh.tri_v[3 * fi + j];
        if (vi < 0 || (size_t)vi >= wv.size()) { *err = std::string("face references a missing vertex in ") + obj_fs[o]; return false; }
        t.v[j][0] = wv[vi].x; t.v[j][1] = wv[vi].y; t.v[j][2] = wv[vi].z;
        if (use_vn) { t.vn[j][0] = wn[vi].x; t.vn[j][1] = wn[vi].y; t.vn[j][2] = wn[vi].z; }
        for (int k = 0; k < 3; k++) { ho.lo[k] = fminf(ho.lo[k], t.v[j][k]); ho.hi[k] = fmaxf(ho.hi[k], t.v[j][k]); }
      }
      finish_triangle(&t);
      out->tris.push_back(t);
    }
    ho.tri_count = (int)out->tris.size() - ho.tri_begin;
    out->objects.push_back(ho);
  }
  out->any_vn = false;
  out->any_specular = false;
  for (size_t i = 0; i < out->tris.size(); i++) {
    out->any_vn = out->any_vn || out->tris[i].has_vn;
    const float* ks = out->tris[i].mat.ks;
    out->any_specular = out->any_specular || (ks[0] != 0.0f || ks[1] != 0.0f || ks[2] != 0.0f);
  }
  out->rebuild_lights();
  default_camera(out->cam);
  return true;
}

}  // namespace ipt
