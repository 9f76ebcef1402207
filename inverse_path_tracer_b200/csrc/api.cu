// api.cu -- the C ABI of libipt_b200.so (declarations and reference citations: include/ipt_b200.h).
//
// A scene handle owns: the host description (for getMaterials / tables), the flat device arrays of
// device_scene.h, one CUDA stream, and lazily grown scratch buffers.  No managed memory, no pointer
// graph.  Handles are kept in a registry so freeScene()/setMaterials() through either library name
// (libpt.so / libipt.so are the same file) validate what they are given.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/ipt_b200.h"
#include "device_scene.h"
#include "kernels.h"
#include "scene_handle.h"
#include "nvtx_ranges.h"
#include "ipt_math.cuh"

namespace ipt {

static thread_local std::string g_error;
static std::string g_asset_root;
static thread_local int g_device = -1;
// Loader / integrator knobs are per calling thread and are copied into the handle by loadScene(): a scene keeps the
// settings it was loaded under (and iptSceneSetIntegrator changes one scene), so threads driving different scenes --
// eight scenes on eight GPUs from one process -- cannot disturb each other.
static thread_local int g_traversal_mode = 0;
static thread_local int g_integrator = 0;          // 0 auto, 1 megakernel, 2 wavefront (BVH scenes)
static thread_local int g_wf_slots = 1 << 24;
std::atomic<unsigned long long> g_launches(0);
static std::mutex g_registry_mutex;
static std::set<void*> g_registry;

constexpr int kBruteForceMaxTris = 128;
#ifndef IPT_LEAF_SIZE
#define IPT_LEAF_SIZE 2
#endif
// leaf size of the internal triangle tree (<= 7: three bits in a descriptor).  The reference's rule is 4 (bvh.h:112, kept by iptBuildBVH);
// on the 5.2 M-triangle scene (frame ms | incoherent Grays/s, profiles/r02_config5.txt): 1: 27.9 | 5.2, 2: 26.3 | 5.5, 3: 26.5 | 5.4,
// 4: 26.9 | 5.3, 6: 28.6 | 4.8, 7: 29.3 | 4.7
constexpr int kTraversalLeafSize = IPT_LEAF_SIZE;
constexpr double kBigTriangleDiag = 0.25;   // a triangle is "large" when its box diagonal exceeds this fraction of the scene's
static thread_local int g_big_max = 128;    // at most this many large triangles leave the tree (0: none)

int fail(const std::string& msg) {
  g_error = msg;
  return -1;
}

SceneHandle* checked(void* p) {
  std::lock_guard<std::mutex> lk(g_registry_mutex);
  if (!p || !g_registry.count(p)) { g_error = "invalid scene handle"; return nullptr; }
  return (SceneHandle*)p;
}
// compute entry points: the scene must live on a GPU (there is no CPU fallback)
SceneHandle* checked_device(void* p) {
  SceneHandle* h = checked(p);
  if (h && h->host_only) { g_error = "scene was loaded host-only (iptSetDevice(-2)); compute needs a CUDA device"; return nullptr; }
  return h;
}

static void pack_materials(const HostScene& hs, std::vector<float4>* mat) {
  const size_t nT = hs.tris.size();
  mat->resize(kMatStride * nT);
  for (size_t i = 0; i < nT; i++) {
    const HostTriangle& t = hs.tris[i];
    const bool any_ks = t.mat.ks[0] != 0.0f || t.mat.ks[1] != 0.0f || t.mat.ks[2] != 0.0f;
    int flags = (t.has_vn ? TRI_HAS_VN : 0) | (any_ks ? TRI_ANY_KS : 0) | ((any_ks && t.mat.ns != 0.0f) ? TRI_SPECULAR : 0) |   // path_trace.cu:132
                (t.n[2] == -1.0f ? TRI_NEG_Z : 0);
    float fbits;
    memcpy(&fbits, &flags, 4);
    const float pi = 3.14159274f;   // (float)M_PI; Kd / pi is the indirect BSDF (path_trace.cu:15-17), divided once here
    (*mat)[0 * nT + i] = make_float4(t.mat.kd[0], t.mat.kd[1], t.mat.kd[2], t.mat.ns);
    (*mat)[1 * nT + i] = make_float4(t.mat.ke[0], t.mat.ke[1], t.mat.ke[2], fbits);
    (*mat)[2 * nT + i] = make_float4(t.mat.kd[0] / pi, t.mat.kd[1] / pi, t.mat.kd[2] / pi, 0.0f);
    (*mat)[3 * nT + i] = make_float4(t.n[0], t.n[1], t.n[2], 0.0f);
    (*mat)[4 * nT + i] = make_float4(t.mat.ks[0], t.mat.ks[1], t.mat.ks[2], t.area);
    float R[9];
    frame_rows_from_z(mk3(t.n[0], t.n[1], t.n[2]), R);   // sampleNextDir's rotation, fixed per triangle
    (*mat)[5 * nT + i] = make_float4(R[0], R[1], R[2], 0.0f);
    (*mat)[6 * nT + i] = make_float4(R[3], R[4], R[5], 0.0f);
    (*mat)[7 * nT + i] = make_float4(R[6], R[7], R[8], 0.0f);
  }
}

// plane groups of the brute-force mode: triangles with value-equal (n, k) share one plane test.  Inside a
// group, two triangles whose shared edge carries exactly negated edge planes (the two halves of a quad:
// out_b = -out_a and d_b = -d_a bit for bit, so sd_b = -sd_a exactly) are paired into one QUAD SLOT of five
// edge records {diag (a's sign), a's other two edges, b's other two edges}; the kernel evaluates five edge
// planes for the two triangles instead of six.  An unpaired triangle occupies a slot whose second half
// can never be hit (edge offset +inf, id -1).  Groups are ordered by their first triangle.
static bool edges_negated(const float* ea, const float* eb) {
  return ea[0] == -eb[0] && ea[1] == -eb[1] && ea[2] == -eb[2] && ea[3] == -eb[3];
}
// `subset`: ascending triangle ids to group (all triangles of a brute-force scene, the large ones of a BVH scene)
static void build_plane_groups(const HostScene& hs, const std::vector<int>& subset, std::vector<float4>* planes, std::vector<int2>* grange,
                               std::vector<float4>* edges, std::vector<int2>* slot_ids) {
  std::vector<std::vector<int>> members;
  for (size_t si = 0; si < subset.size(); si++) {
    const int i = subset[si];
    const HostTriangle& t = hs.tris[i];
    int g = -1;
    for (size_t k = 0; k < members.size(); k++) {
      const HostTriangle& r = hs.tris[members[k][0]];
      // value equality, so -0 and +0 components group together: a zero sign can only flip the sign of a
      // zero denom / num, and those are rejected by |denom| <= MIN_DOT resp. t <= EPSILON either way
      if (r.n[0] == t.n[0] && r.n[1] == t.n[1] && r.n[2] == t.n[2] && r.plane_k == t.plane_k) { g = (int)k; break; }
    }
    if (g < 0) { g = (int)members.size(); members.push_back(std::vector<int>()); }
    members[g].push_back(i);
  }
  planes->clear(); grange->clear(); edges->clear(); slot_ids->clear();
  auto push_edge = [&](const float* e) { edges->push_back(make_float4(e[0], e[1], e[2], e[3])); };
  const float never[4] = {0.0f, 0.0f, 0.0f, INFINITY};
  for (size_t g = 0; g < members.size(); g++) {
    const HostTriangle& r = hs.tris[members[g][0]];
    planes->push_back(make_float4(r.n[0], r.n[1], r.n[2], r.plane_k));
    int2 rg;
    rg.x = (int)slot_ids->size();
    std::vector<char> used(members[g].size(), 0);
    for (size_t ia = 0; ia < members[g].size(); ia++) {
      if (used[ia]) continue;
      used[ia] = 1;
      const HostTriangle& ta = hs.tris[members[g][ia]];
      int partner = -1, ja = -1, jb = -1;
      for (size_t ib = ia + 1; ib < members[g].size() && partner < 0; ib++) {
        if (used[ib]) continue;
        const HostTriangle& tb = hs.tris[members[g][ib]];
        for (int x = 0; x < 3 && partner < 0; x++)
          for (int y = 0; y < 3; y++)
            if (edges_negated(ta.edge[x], tb.edge[y])) { partner = (int)ib; ja = x; jb = y; break; }
      }
      int2 ids;
      ids.x = ta.idx;
      if (partner >= 0) {
        used[partner] = 1;
        const HostTriangle& tb = hs.tris[members[g][partner]];
        ids.y = tb.idx;
        push_edge(ta.edge[ja]);
        push_edge(ta.edge[(ja + 1) % 3]); push_edge(ta.edge[(ja + 2) % 3]);
        push_edge(tb.edge[(jb + 1) % 3]); push_edge(tb.edge[(jb + 2) % 3]);
      } else {
        ids.y = -1;
        push_edge(ta.edge[0]); push_edge(ta.edge[1]); push_edge(ta.edge[2]);
        push_edge(never); push_edge(never);
      }
      slot_ids->push_back(ids);
    }
    rg.y = (int)slot_ids->size();
    grange->push_back(rg);
  }
}

static int upload_scene(SceneHandle* h) {
  const HostScene& hs = h->host;
  const size_t nT = hs.tris.size(), nE = hs.emissives.size();
  std::vector<float4> geom(4 * nT), mat, light(3 * std::max<size_t>(nE, 1)), vtx(5 * nT);
  for (size_t i = 0; i < nT; i++) {
    const HostTriangle& t = hs.tris[i];
    geom[4 * i] = make_float4(t.n[0], t.n[1], t.n[2], t.plane_k);
    for (int j = 0; j < 3; j++) geom[4 * i + 1 + j] = make_float4(t.edge[j][0], t.edge[j][1], t.edge[j][2], t.edge[j][3]);
    vtx[5 * i + 0] = make_float4(t.v[0][0], t.v[0][1], t.v[0][2], t.vn[0][0]);
    vtx[5 * i + 1] = make_float4(t.v[1][0], t.v[1][1], t.v[1][2], t.vn[0][1]);
    vtx[5 * i + 2] = make_float4(t.v[2][0], t.v[2][1], t.v[2][2], t.vn[0][2]);
    vtx[5 * i + 3] = make_float4(t.vn[1][0], t.vn[1][1], t.vn[1][2], t.area);
    vtx[5 * i + 4] = make_float4(t.vn[2][0], t.vn[2][1], t.vn[2][2], 0.0f);
  }
  pack_materials(hs, &mat);
  for (size_t k = 0; k < nE; k++) {
    const HostTriangle& t = hs.tris[hs.emissives[k]];
    int ti = t.idx;
    float tb;
    memcpy(&tb, &ti, 4);
    light[3 * k + 0] = make_float4(t.v[0][0], t.v[0][1], t.v[0][2], hs.light_cdf[k]);
    light[3 * k + 1] = make_float4(t.v[1][0], t.v[1][1], t.v[1][2], hs.light_pt[k]);
    light[3 * k + 2] = make_float4(t.v[2][0], t.v[2][1], t.v[2][2], tb);
  }
  IPT_CUDA(cudaMalloc(&h->geom, std::max<size_t>(geom.size(), 1) * sizeof(float4)));
  IPT_CUDA(cudaMalloc(&h->mat, std::max<size_t>(mat.size(), 1) * sizeof(float4)));
  IPT_CUDA(cudaMalloc(&h->light, light.size() * sizeof(float4)));
  IPT_CUDA(cudaMalloc(&h->vtx, std::max<size_t>(vtx.size(), 1) * sizeof(float4)));
  if (nT) {
    IPT_CUDA(cudaMemcpy(h->geom, geom.data(), geom.size() * sizeof(float4), cudaMemcpyHostToDevice));
    IPT_CUDA(cudaMemcpy(h->mat, mat.data(), mat.size() * sizeof(float4), cudaMemcpyHostToDevice));
    IPT_CUDA(cudaMemcpy(h->vtx, vtx.data(), vtx.size() * sizeof(float4), cudaMemcpyHostToDevice));
  }
  IPT_CUDA(cudaMemcpy(h->light, light.data(), light.size() * sizeof(float4), cudaMemcpyHostToDevice));

  h->use_bvh = (h->traversal_mode == 2) || (h->traversal_mode == 0 && (int)nT > kBruteForceMaxTris);
  // the brute-force set: every triangle of a small scene; in a BVH scene the few triangles that are large against
  // the scene (room walls around a dense mesh).  Left in the tree, a large triangle inflates the boxes of every
  // node above its leaf and rays far from the mesh descend deep for nothing; outside it the tree stays tight and
  // the set doubles as an early bound (its hit is found before the traversal starts).  Only automatic BVH scenes
  // do this: a scene forced into BVH mode (iptSetTraversalMode(2)) keeps every triangle in the tree.
  std::vector<int> brute, rest;
  if (!h->use_bvh) {
    for (size_t i = 0; i < nT; i++) brute.push_back((int)i);
  } else if ((int)nT > kBruteForceMaxTris && h->big_max > 0) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = 0; i < nT; i++)
      for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], hs.tris[i].v[j][k]); hi[k] = fmaxf(hi[k], hs.tris[i].v[j][k]); }
    const double scene_diag2 = (double)(hi[0] - lo[0]) * (hi[0] - lo[0]) + (double)(hi[1] - lo[1]) * (hi[1] - lo[1]) + (double)(hi[2] - lo[2]) * (hi[2] - lo[2]);
    std::vector<std::pair<double, int>> cand;
    for (size_t i = 0; i < nT; i++) {
      const HostTriangle& t = hs.tris[i];
      double d2 = 0.0;
      for (int k = 0; k < 3; k++) {
        const double e = (double)fmaxf(fmaxf(t.v[0][k], t.v[1][k]), t.v[2][k]) - (double)fminf(fminf(t.v[0][k], t.v[1][k]), t.v[2][k]);
        d2 += e * e;
      }
      if (d2 > kBigTriangleDiag * kBigTriangleDiag * scene_diag2) cand.push_back(std::make_pair(-d2, (int)i));
    }
    std::sort(cand.begin(), cand.end());
    if ((int)cand.size() > h->big_max) cand.resize(h->big_max);
    std::vector<char> is_big(nT, 0);
    for (size_t c = 0; c < cand.size(); c++) is_big[cand[c].second] = 1;
    for (size_t i = 0; i < nT; i++) (is_big[i] ? brute : rest).push_back((int)i);
  }
  if (h->use_bvh && rest.empty() && brute.empty()) for (size_t i = 0; i < nT; i++) rest.push_back((int)i);
  h->n_big = h->use_bvh ? (int)brute.size() : 0;
  if (h->n_big > 0) {
    // triangle -> index in the brute-force set (255 = in the tree): the wavefront adjoint privatises the gradient of these
    // few, heavily hit triangles per warp in shared memory (k_wf_sweep)
    std::vector<unsigned char> slot(nT, 255);
    for (size_t i = 0; i < brute.size(); i++) slot[brute[i]] = (unsigned char)i;
    IPT_CUDA(cudaMalloc(&h->big_slot, nT));
    IPT_CUDA(cudaMemcpy(h->big_slot, slot.data(), nT, cudaMemcpyHostToDevice));
    IPT_CUDA(cudaMalloc(&h->big_tri, brute.size() * sizeof(int)));
    IPT_CUDA(cudaMemcpy(h->big_tri, brute.data(), brute.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  if (!brute.empty()) {
    std::vector<float4> planes, edges;
    std::vector<int2> grange, slot_tri;
    build_plane_groups(hs, brute, &planes, &grange, &edges, &slot_tri);
    // the kernels walk the set as one record stream (device_scene.h): {plane}{meta} per group, 5 edges + ids per slot
    std::vector<float4> stream;
    auto ibits = [](int v) { float f; memcpy(&f, &v, 4); return f; };
    for (size_t g = 0; g < planes.size(); g++) {
      const int ns = grange[g].y - grange[g].x;
      stream.push_back(planes[g]);
      stream.push_back(make_float4(ibits((2 + 6 * ns) * (int)sizeof(float4)), 0.0f, 0.0f, 0.0f));
      for (int j = grange[g].x; j < grange[g].y; j++) {
        for (int k = 0; k < 5; k++) stream.push_back(edges[5 * (size_t)j + k]);
        stream.push_back(make_float4(ibits(slot_tri[j].x), ibits(slot_tri[j].y), 0.0f, 0.0f));
      }
    }
    h->n_groups = (int)planes.size();
    h->n_stream = (int)stream.size();
    IPT_CUDA(cudaMalloc(&h->bf_stream, stream.size() * sizeof(float4)));
    IPT_CUDA(cudaMemcpy(h->bf_stream, stream.data(), stream.size() * sizeof(float4), cudaMemcpyHostToDevice));
  }
  if (h->use_bvh && nT) {
    // triangle-level tree: box = vertex bounds, split position = triangle center (the analogue of
    // Object::getBBox / getPosition, scene_basics.h:417-424, with triangles as primitives)
    const size_t nR = rest.size();
    std::vector<float> boxes(6 * nR), pos(3 * nR);
    for (size_t i = 0; i < nR; i++) {
      const HostTriangle& t = hs.tris[rest[i]];
      for (int k = 0; k < 3; k++) {
        boxes[6 * i + k] = fminf(fminf(t.v[0][k], t.v[1][k]), t.v[2][k]);
        boxes[6 * i + 3 + k] = fmaxf(fmaxf(t.v[0][k], t.v[1][k]), t.v[2][k]);
        pos[3 * i + k] = t.c[k];
      }
    }
    std::vector<BvhNodeHost> nodes;
    std::vector<int> perm;
    std::string err;
    if (nR && !build_bvh(boxes.data(), pos.data(), (int)nR, kTraversalLeafSize, &nodes, &perm, h->device, &err)) return fail("BVH build failed: " + err);
    for (size_t j = 0; j < perm.size(); j++) perm[j] = rest[perm[j]];   // tree order -> triangle ids
    std::vector<float4> packed(2 * nodes.size());
    for (size_t i = 0; i < nodes.size(); i++) {
      int ro = nodes[i].rightOffset;
      int sp = (nodes[i].start << 3) | (nodes[i].nPrims & 7);
      float a, b;
      memcpy(&a, &ro, 4);
      memcpy(&b, &sp, 4);
      // The traversal copy of each box is padded.  The reference's triangle test accepts a hit whose
      // ROUNDED hit point passes the edge planes, i.e. it tolerates a few ulps of the coordinate
      // magnitude; a box that is exact in real arithmetic can therefore exclude a hit the brute-force
      // loop accepts (seen with near axis-parallel rays through a vertex).  pad = 1e-5 * (1 + |coord|)
      // is > 40x that tolerance.  iptBuildBVH (layout parity) returns the unpadded boxes.
      float lo[3], hi[3];
      for (int k = 0; k < 3; k++) {
        const float m = fmaxf(fabsf(nodes[i].lo[k]), fabsf(nodes[i].hi[k]));
        const float pad = 1e-5f * (1.0f + m);
        lo[k] = nodes[i].lo[k] - pad;
        hi[k] = nodes[i].hi[k] + pad;
      }
      packed[2 * i] = make_float4(lo[0], lo[1], lo[2], a);
      packed[2 * i + 1] = make_float4(hi[0], hi[1], hi[2], b);
    }
    // traversal records: every inner node carries both (padded) child boxes and the child descriptors
    auto desc_of = [&](size_t c) {
      return nodes[c].rightOffset != 0 ? (int)c : ~((nodes[c].start << 3) | (nodes[c].nPrims & 7));
    };
    std::vector<float4> wide(4 * nodes.size(), make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    std::vector<int> depth(nodes.size(), 1);
    int max_depth = 1;
    for (size_t i = 0; i < nodes.size(); i++) {
      if (nodes[i].rightOffset == 0) continue;
      const size_t l = i + 1, r = i + (size_t)nodes[i].rightOffset;
      depth[l] = depth[r] = depth[i] + 1;                 // parents precede children in DFS pre-order
      max_depth = std::max(max_depth, depth[i] + 1);
      int dl = desc_of(l), dr = desc_of(r);
      float fl, fr;
      memcpy(&fl, &dl, 4);
      memcpy(&fr, &dr, 4);
      wide[4 * i + 0] = make_float4(packed[2 * l].x, packed[2 * l].y, packed[2 * l].z, fl);
      wide[4 * i + 1] = make_float4(packed[2 * l + 1].x, packed[2 * l + 1].y, packed[2 * l + 1].z, fr);
      wide[4 * i + 2] = make_float4(packed[2 * r].x, packed[2 * r].y, packed[2 * r].z, 0.0f);
      wide[4 * i + 3] = make_float4(packed[2 * r + 1].x, packed[2 * r + 1].y, packed[2 * r + 1].z, 0.0f);
    }
    if (max_depth > 64) return fail("BVH deeper than 64 levels: the traversal stack (like the reference's todo[64], bvh.h:43) cannot hold it");
    // quantised traversal records (device_scene.h): 16 bits per coordinate on the padded root box, lower bounds rounded
    // down and upper bounds up (checked against the float the kernel reconstructs), so a quantised box contains its
    // float box: the closest hit is a minimum over all triangles that pass their own test, a larger box never changes it
    std::vector<float4> quant;
#if IPT_BVH_QUANT
    quant.assign(2 * nodes.size(), make_float4(0.0f, 0.0f, 0.0f, 0.0f));
#endif
    if (!quant.empty()) {
      for (int k = 0; k < 3; k++) {
        const float lo = (&packed[0].x)[k], hi = (&packed[1].x)[k];
        // two grid steps of margin on either side: every box is padded outwards by one more step than rounding needs
        // (the kernel's folded decode is exact to half a step, intersect.cuh)
        h->bvh_qscale[k] = (hi > lo) ? (hi - lo) / 65531.0f * 1.0000005f : 1.0f;
        h->bvh_qoff[k] = lo - 2.0f * h->bvh_qscale[k];
      }
      auto qlo = [&](float v, int k) {
        int q = (int)floorf((v - h->bvh_qoff[k]) / h->bvh_qscale[k]);
        q = std::max(0, std::min(65535, q));
        while (q > 0 && fmaf((float)q, h->bvh_qscale[k], h->bvh_qoff[k]) > v) q--;
        return (unsigned)std::max(0, q - 1);
      };
      auto qhi = [&](float v, int k) {
        int q = (int)ceilf((v - h->bvh_qoff[k]) / h->bvh_qscale[k]);
        q = std::max(0, std::min(65535, q));
        while (q < 65535 && fmaf((float)q, h->bvh_qscale[k], h->bvh_qoff[k]) < v) q++;
        return (unsigned)std::min(65535, q + 1);
      };
      auto ubits = [](unsigned v) { float f; memcpy(&f, &v, 4); return f; };
      auto qbox = [&](size_t c, unsigned* w) {   // three words of the quantised (padded) box of binary node c
        const float4 blo = packed[2 * c], bhi = packed[2 * c + 1];
        const unsigned lx = qlo(blo.x, 0), ly = qlo(blo.y, 1), lz = qlo(blo.z, 2), hx = qhi(bhi.x, 0), hy = qhi(bhi.y, 1), hz = qhi(bhi.z, 2);
        w[0] = lx | (ly << 16); w[1] = lz | (hx << 16); w[2] = hy | (hz << 16);
      };
#if IPT_BVH_QUANT == 2
      // 4-wide collapse of the binary tree: the 4-node headed by binary inner node i holds the children of i's children (a
      // child of i that is a leaf stays as it is), i.e. every other level of the binary tree disappears.  Records (64 bytes,
      // one 128-byte line holds two): 4 x 3 words of quantised boxes, 4 descriptors (>= 0: 4-node index, < 0: leaf as
      // before; an empty slot is a point box at the grid corner with the descriptor of a leaf without triangles).  4-nodes
      // are numbered in depth-first order.  `held` = stack entries a ray can hold when it arrives at the node (every
      // ancestor may have pushed all its other children): the bound on the traversal stack.
      struct Node4 { size_t ch[4]; int nc; };
      std::vector<int> id4(nodes.size(), -1);
      std::vector<Node4> n4;
      int stack_need = 1;
      if (nodes[0].rightOffset != 0) {
        std::vector<std::pair<size_t, int>> todo;
        todo.emplace_back((size_t)0, 0);
        while (!todo.empty()) {
          const size_t i = todo.back().first;
          const int held = todo.back().second;
          todo.pop_back();
          id4[i] = (int)n4.size();
          Node4 nd;
          nd.nc = 0;
          const size_t lr[2] = {i + 1, i + (size_t)nodes[i].rightOffset};
          for (size_t c : lr) {
            if (nodes[c].rightOffset != 0) { nd.ch[nd.nc++] = c + 1; nd.ch[nd.nc++] = c + (size_t)nodes[c].rightOffset; }
            else nd.ch[nd.nc++] = c;
          }
          stack_need = std::max(stack_need, held + nd.nc - 1);
          for (int k = nd.nc - 1; k >= 0; k--)
            if (nodes[nd.ch[k]].rightOffset != 0) todo.emplace_back(nd.ch[k], held + nd.nc - 1);
          n4.push_back(nd);
        }
      }
      quant.assign(4 * std::max<size_t>(n4.size(), 1), make_float4(0.0f, 0.0f, 0.0f, 0.0f));
      for (size_t n = 0; n < n4.size(); n++) {
        unsigned w[16];
        for (int k = 0; k < 4; k++) {
          if (k < n4[n].nc) {
            const size_t c = n4[n].ch[k];
            qbox(c, w + 3 * k);
            w[12 + k] = (unsigned)(nodes[c].rightOffset != 0 ? id4[c] : desc_of(c));
          } else {
            w[3 * k] = 0xffffffffu; w[3 * k + 1] = 0xffffffffu; w[3 * k + 2] = 0xffffffffu;
            w[12 + k] = 0xffffffffu;   // ~0: leaf, start 0, no triangles
          }
        }
        for (int q = 0; q < 4; q++) quant[4 * n + q] = make_float4(ubits(w[4 * q]), ubits(w[4 * q + 1]), ubits(w[4 * q + 2]), ubits(w[4 * q + 3]));
      }
      max_depth = stack_need + 1;
      if (max_depth > 64) return fail("BVH too deep for the 64-entry traversal stack");
#else
      for (size_t i = 0; i < nodes.size(); i++) {
        if (nodes[i].rightOffset == 0) continue;
        const size_t ch[2] = {i + 1, i + (size_t)nodes[i].rightOffset};
        unsigned w[6];
        for (int c = 0; c < 2; c++) qbox(ch[c], w + 3 * c);
        quant[2 * i] = make_float4(ubits(w[0]), ubits(w[1]), ubits(w[2]), ubits(w[3]));
        quant[2 * i + 1] = make_float4(ubits(w[4]), ubits(w[5]), wide[4 * i + 0].w, wide[4 * i + 1].w);   // + descriptors of the left / right child
      }
#endif
    }
    IPT_CUDA(cudaMalloc(&h->bvh_q, std::max<size_t>(quant.size(), 2) * sizeof(float4)));
    if (!quant.empty()) IPT_CUDA(cudaMemcpy(h->bvh_q, quant.data(), quant.size() * sizeof(float4), cudaMemcpyHostToDevice));
    h->bvh_root_desc = nodes.empty() ? 0 : desc_of(0);
    h->bvh_depth = max_depth;
    h->n_nodes = (int)nodes.size();
#if IPT_BVH_QUANT
    const size_t n_wide_dev = 4;   // the float traversal records are not used by this build (device_scene.h): keep a token allocation
#else
    const size_t n_wide_dev = std::max<size_t>(wide.size(), 4);
#endif
    IPT_CUDA(cudaMalloc(&h->bvh_wide, n_wide_dev * sizeof(float4)));
    IPT_CUDA(cudaMalloc(&h->bvh_nodes, std::max<size_t>(packed.size(), 2) * sizeof(float4)));
    IPT_CUDA(cudaMalloc(&h->bvh_prims, std::max<size_t>(perm.size(), 1) * sizeof(int)));
    if (!nodes.empty()) {
      IPT_CUDA(cudaMemcpy(h->bvh_wide, wide.data(), std::min(n_wide_dev, wide.size()) * sizeof(float4), cudaMemcpyHostToDevice));
      IPT_CUDA(cudaMemcpy(h->bvh_nodes, packed.data(), packed.size() * sizeof(float4), cudaMemcpyHostToDevice));
      IPT_CUDA(cudaMemcpy(h->bvh_prims, perm.data(), perm.size() * sizeof(int), cudaMemcpyHostToDevice));
      // leaf triangles are read in tree order: store their records that way (slot j = triangle perm[j]), so a leaf is
      // one contiguous run and the id lookup is only needed for accepted hits.  Planes first ([n] float4), then the edge
      // planes ([3 n]): a leaf visit reads the planes of all its triangles at once and the 48 bytes of edges only for the
      // triangles whose plane test passes
      const size_t n_slots = perm.size();
      std::vector<float4> leaf_geom(4 * n_slots);
      for (size_t j = 0; j < n_slots; j++) {
        leaf_geom[j] = geom[4 * (size_t)perm[j]];
        memcpy(&leaf_geom[n_slots + 3 * j], &geom[4 * (size_t)perm[j] + 1], 3 * sizeof(float4));
      }
      IPT_CUDA(cudaMemcpy(h->geom, leaf_geom.data(), leaf_geom.size() * sizeof(float4), cudaMemcpyHostToDevice));
      h->geom_edges = h->geom + n_slots;
    }
  }
  return 0;
}

int make_plan(const SceneHandle* h, const IptRenderParams* p, Plan* plan) {
  (void)h;
  if (!p) return fail("null render params");
  if (p->width <= 0 || p->height <= 0 || p->spp <= 0) return fail("width, height and spp must be positive");
  if (!(p->p_rr > 0.0f)) return fail("p_rr must be positive");
  if (p->p_rr >= 1.0f && p->max_depth <= 0) return fail("unbounded depth needs p_rr < 1 (paths would never terminate in a closed scene)");
  int rb = p->row_begin < 0 ? 0 : p->row_begin;
  int re = (p->row_end <= 0 || p->row_end > p->height) ? p->height : p->row_end;
  if (rb >= re) return fail("empty row shard");
  if ((long long)p->width * p->height > 0x7fffffffLL) return fail("image too large");
  DevRender& rp = plan->rp;
  rp.width = p->width; rp.height = p->height; rp.spp = p->spp; rp.max_depth = p->max_depth;
  rp.p_rr = p->p_rr; rp.key0 = p->seed_lo; rp.key1 = p->seed_hi;
  rp.keys = philox_key_schedule(rp.key0, rp.key1);
  rp.inv_spp = 1.0f / (float)p->spp;
  rp.two_over_w = 2.0f / (float)p->width;
  rp.two_over_h = 2.0f / (float)p->height;
  rp.c_diff = (1.0f / 0.318309873f) / p->p_rr;   // canon v2: diffuse coeff = cos * ((1 / (1/pi)) / p_rr), path_trace.cu:137
  int gs = p->group_samples;
  if (gs < 0) {   // auto (include/ipt_b200.h): depends on the full image size only
    gs = 32;   // cap 32, not 64: the bench job as one of eight row shards runs 21.3 instead of 23.5 ms (tail of 1.3 ms work items), whole: +0.75 % (profiles/r02_group_cap.txt)
    while (gs > 1 && (long long)p->width * p->height * p->spp < ((long long)gs << 22)) gs >>= 1;
  }
  rp.group_samples = (gs <= 0 || gs > p->spp) ? p->spp : gs;
  rp.groups = (p->spp + rp.group_samples - 1) / rp.group_samples;
  // the kernels index work items (pixel, group) in 32 bits
  if ((long long)p->width * p->height * rp.groups >= (1ll << 31) - (1ll << 24))
    return fail("too many work items (width * height * ceil(spp / group_samples) must stay below 2^31): render row shards or use larger sample groups");
  const int step = p->row_step < 1 ? 1 : p->row_step;
  const int rows = (re - rb + step - 1) / step;
  rp.row_begin = rb;
  rp.row_step = step;
  rp.n_pix = rows * p->width;
  // work-item index -> (pixel, group) and pixel -> (row, column) by multiplication (an integer division per item is a
  // visible cost when items are small)
  auto magic = [](unsigned d, unsigned* m, int* shift) {   // device_scene.h: floor(x / d) = umulhi(x, m) >> shift for x < 2^31
    *m = 0; *shift = 0;
    if (d < 2) return;
    int l = 0;
    while ((1ull << l) < d) l++;
    *m = (unsigned)(((1ull << (31 + l)) / d) + 1ull);
    *shift = l - 1;
  };
  magic((unsigned)rp.groups, &rp.magic_groups, &rp.shift_groups);
  magic((unsigned)p->width, &rp.magic_width, &rp.shift_width);
  plan->rows = rows;
  return 0;
}

// ---- chunked launches ------------------------------------------------------------------------------------------
// A persistent launch ends in a tail: the work queue is empty and the warps run on for as long as their longest path
// (path lengths are geometric; ~90 bounces among 10^5 lanes), most lanes idle.  For a 1024^2 x 1024 spp frame that is
// nothing; for one rank's share of a 500^2 x 64 spp inverse iteration (20 paths per lane) it is a quarter of each kernel.
// A launch can be split into row blocks issued on separate internal streams (iptSceneSetChunks / IPT_CHUNKS), so that a
// chunk's blocks take over the SM slots the previous chunk's blocks give up.  MEASURED SLOWER and therefore off by
// default (profiles/r02_chunked_launches.txt: inverse iteration of a 1/8 shard 1.98 / 2.25 / 2.81 ms with 1 / 2 / 4 chunks,
// Cornell 512^2 x 64 spp 3.20 / 3.32 / 3.50 ms): a block frees its slot only when all four of its warps have run dry,
// which in the tail is almost as late as the end of the launch, so the chunks serialise and each adds its own tail.
int pick_chunks(const SceneHandle* h, const Plan& plan) {
  const char* env = getenv("IPT_CHUNKS");   // developer A/B
  int k = h->chunks;
  if (env && *env) k = atoi(env);
  return std::max(1, std::min(std::min(k, SceneHandle::kMaxChunks), plan.rows));
}

Fork::Fork(SceneHandle* h_, cudaStream_t st_, int k_) : h(h_), st(st_), k(k_) {}
cudaError_t Fork::begin() {
  if (k <= 1) return cudaSuccess;
  cudaError_t e = cudaEventRecord(h->ev_fork, st);
  for (int c = 0; c < k && e == cudaSuccess; c++) e = cudaStreamWaitEvent(h->cstream[c], h->ev_fork, 0);
  return e;
}
cudaError_t Fork::end() {
  if (k <= 1) return cudaSuccess;
  cudaError_t e = cudaSuccess;
  for (int c = 0; c < k && e == cudaSuccess; c++) {
    e = cudaEventRecord(h->ev_join[c], h->cstream[c]);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, h->ev_join[c], 0);
  }
  return e;
}

// rows [lr0, lr1) of the shard (shard-local row indices) that chunk c of k covers, and the render descriptor of that block
void chunk_render(const Plan& plan, int c, int k, DevRender* rp_c, int* lr0, int* lr1) {
  *lr0 = (int)((long long)plan.rows * c / k);
  *lr1 = (int)((long long)plan.rows * (c + 1) / k);
  *rp_c = plan.rp;
  rp_c->row_begin = plan.rp.row_begin + *lr0 * plan.rp.row_step;
  rp_c->n_pix = (*lr1 - *lr0) * plan.rp.width;
}

int prepare_forward(SceneHandle* h, const Plan& plan) {
  IPT_CUDA(h->partial.reserve((size_t)plan.rp.n_pix * plan.rp.groups * 3 * sizeof(float)));
  IPT_CUDA(h->counters.reserve(16 * sizeof(unsigned long long)));
  return 0;
}

// forward pass of chunk c into its block of h->partial, on `st_c`.  counters[0..2] are shared by the chunks (atomics)
// and must have been zeroed by the caller; the chunk's work counter is ctr[4 + c].
int forward_chunk(SceneHandle* h, const Plan& plan, const DevScene& sc, int c, int k, unsigned long long* counters, cudaStream_t st_c) {
  DevRender rp;
  int lr0, lr1;
  chunk_render(plan, c, k, &rp, &lr0, &lr1);
  if (lr1 <= lr0) return 0;
  unsigned long long* ctr = (unsigned long long*)h->counters.ptr;
  DevOut out;
  memset(&out, 0, sizeof(out));
  out.work_counter = reinterpret_cast<unsigned*>(ctr + 4 + c);
  out.counters = counters;
  out.partial = (float*)h->partial.ptr + (size_t)lr0 * rp.width * rp.groups * 3;
  const size_t smem = paths_smem_bytes(sc, 0);
  const int grid = h->sm_count * paths_blocks_per_sm(MODE_RENDER_ID, sc, smem);
  {
    NvtxRange r("ipt.paths.render");
    IPT_CUDA(launch_paths(MODE_RENDER_ID, sc, rp, out, grid, smem, st_c));
  }
  g_launches += 1;
  return 0;
}

int launch_forward(SceneHandle* h, const Plan& plan, float* img_dev, unsigned char* img8_dev, unsigned long long* counters_dev,
                   cudaStream_t st, bool resolve) {
  const DevScene sc = h->dev();
  const DevRender& rp = plan.rp;
  if (prepare_forward(h, plan) != 0) return -1;
  unsigned long long* ctr = (unsigned long long*)h->counters.ptr;
  IPT_CUDA(cudaMemsetAsync(ctr, 0, 16 * sizeof(unsigned long long), st));
  if (counters_dev) IPT_CUDA(cudaMemsetAsync(counters_dev, 0, 3 * sizeof(unsigned long long), st));
  unsigned long long* counters = counters_dev ? counters_dev : ctr;
  float* partial = (float*)h->partial.ptr;
  if ((sc.use_bvh && h->integrator != 1) || (!sc.use_bvh && h->integrator == 2)) {
    // large scenes: wavefront (wavefront.cu); batches of at most `cap` paths
    const long long total = (long long)rp.n_pix * rp.spp;
    const int cap = (int)std::max<long long>(1, std::min<long long>(total, (long long)h->wf_slots));
    IPT_CUDA(h->wavefront.reserve(wavefront_workspace_bytes(cap)));
    if (!h->host_live) IPT_CUDA(cudaHostAlloc((void**)&h->host_live, sizeof(int), cudaHostAllocDefault));
    const int slot = (int)(h->ev_count % SceneHandle::kEventRing);
    IPT_CUDA(cudaEventRecord(h->ev_start[slot], st));
    unsigned long long launches = 0;
    IPT_CUDA(launch_wavefront(sc, rp, h->wavefront.ptr, cap, partial, counters, h->sm_count, h->host_live, &launches, st));
    IPT_CUDA(cudaEventRecord(h->ev_stop[slot], st));
    h->ev_count++;
    g_launches += launches;
  } else {
    const int k = pick_chunks(h, plan);
    const int slot = (int)(h->ev_count % SceneHandle::kEventRing);
    IPT_CUDA(cudaEventRecord(h->ev_start[slot], st));
    Fork f(h, st, k);
    IPT_CUDA(f.begin());
    for (int c = 0; c < k; c++)
      if (forward_chunk(h, plan, sc, c, k, counters, f.stream(c)) != 0) return -1;
    IPT_CUDA(f.end());
    IPT_CUDA(cudaEventRecord(h->ev_stop[slot], st));
    h->ev_count++;
  }
  if (!resolve) return 0;
  IPT_CUDA(launch_resolve(partial, rp.n_pix, rp.groups, img_dev, img8_dev, st));
  g_launches += 1;
  return 0;
}

// gradient tiles: one [3 nT] float tile per block in global memory, updated with fire-and-forget float reductions
// (shared-memory float atomics are compare-and-swap loops on this architecture); scenes whose tiles would be too
// large fall back to double atomics on the result itself
int prepare_grad(SceneHandle* h, const Plan& plan, int k, GradLaunch* g) {
  const DevScene sc = h->dev();
  const int nT = sc.nT;
  const char* acc_env = getenv("IPT_GRAD_ACC");   // developer A/B: 0 = global double, 1 = shared tile, 2 = global tile
  g->smem = paths_smem_bytes(sc, 0);
  g->acc_mode = (3 * (size_t)nT <= 65536) ? 2 : 0;
  if (acc_env && *acc_env) g->acc_mode = atoi(acc_env);
  if (g->acc_mode == 1) g->smem = paths_smem_bytes(sc, 3 * nT);
  g->grid = h->sm_count * paths_blocks_per_sm(MODE_GRAD_ID, sc, g->smem);
  g->vdepth = plan.rp.max_depth;
  IPT_CUDA(h->counters.reserve(16 * sizeof(unsigned long long)));
  if (g->acc_mode) IPT_CUDA(h->block_acc.reserve((size_t)k * g->grid * 3 * nT * sizeof(float)));
  // adjoint vertex stacks: two halves of max_depth 32-byte vertices per resident thread (no initialisation needed:
  // a sweep only reads what its own path wrote)
  g->vstack_bytes = (size_t)g->grid * paths_block_threads() * 2 * (size_t)g->vdepth * (sc.plain ? 32 : 64);   // 64-byte vertices carry the Phong factors
  IPT_CUDA(h->vstacks.reserve((size_t)k * g->vstack_bytes));
  return 0;
}

// path adjoint of chunk c on `st_c`; h->dldpx holds dLoss/dL of the whole shard, ctr[0..2] were zeroed by the caller
int grad_chunk(SceneHandle* h, const Plan& plan, const DevScene& sc, const GradLaunch& g, int c, int k, const unsigned char* target_dev,
               double* lossgrad_dev, cudaStream_t st_c) {
  DevRender rp;
  int lr0, lr1;
  chunk_render(plan, c, k, &rp, &lr0, &lr1);
  if (lr1 <= lr0) return 0;
  const int nT = sc.nT;
  unsigned long long* ctr = (unsigned long long*)h->counters.ptr;
  DevOut out;
  memset(&out, 0, sizeof(out));
  out.work_counter = reinterpret_cast<unsigned*>(ctr + 4 + c);
  out.counters = ctr;
  out.target = target_dev;
  out.dldpx = (const float*)h->dldpx.ptr + (size_t)lr0 * rp.width * 3;
  out.acc_stride = 3 * nT;
  out.acc_in_smem = g.acc_mode;
  if (g.acc_mode) {
    out.block_acc = (float*)h->block_acc.ptr + (size_t)c * g.grid * 3 * nT;
    if (g.acc_mode == 2) IPT_CUDA(cudaMemsetAsync(out.block_acc, 0, (size_t)g.grid * 3 * nT * sizeof(float), st_c));
  } else {
    out.global_acc = lossgrad_dev + 1;
  }
  out.vdepth = g.vdepth;
  out.vstacks = (float4*)((char*)h->vstacks.ptr + (size_t)c * g.vstack_bytes);
  {
    NvtxRange r("ipt.paths.adjoint");
    IPT_CUDA(launch_paths(MODE_GRAD_ID, sc, rp, out, g.grid, g.smem, st_c));
  }
  g_launches += 1;
  return 0;
}

int launch_grad_pass(SceneHandle* h, const Plan& plan, const unsigned char* target_dev, double* lossgrad_dev, cudaStream_t st) {
  const DevScene sc = h->dev();
  const int nT = sc.nT;
  if (nT == 0) return 0;
  if (sc.use_bvh && h->integrator != 1) {
    // BVH scenes: the adjoint replays the paths through the wavefront integrator with vertex recording and sweeps every
    // batch backwards (wavefront.cu, k_wf_sweep); the record bounds the batch (vdepth x cap x 32 or 64 bytes)
    const DevRender& rp = plan.rp;
    const int vf4 = sc.plain ? 2 : 4;
    const long long total = (long long)rp.n_pix * rp.spp;
    const long long rec_budget = (long long)6 << 30;   // bytes of vertex record per batch
    long long cap_ll = std::min<long long>(std::min<long long>(total, (long long)h->wf_slots), rec_budget / ((long long)rp.max_depth * vf4 * 16));
    const int cap = (int)std::max<long long>(1, cap_ll);
    IPT_CUDA(h->wavefront.reserve(wavefront_workspace_bytes(cap)));
    IPT_CUDA(h->vstacks.reserve(wavefront_record_bytes(cap, rp.max_depth, vf4)));
    IPT_CUDA(h->counters.reserve(16 * sizeof(unsigned long long)));
    if (!h->host_live) IPT_CUDA(cudaHostAlloc((void**)&h->host_live, sizeof(int), cudaHostAllocDefault));
    unsigned long long* ctr = (unsigned long long*)h->counters.ptr;
    IPT_CUDA(cudaMemsetAsync(ctr, 0, 16 * sizeof(unsigned long long), st));
    WavefrontAdjoint adj;
    adj.record = h->vstacks.ptr;
    adj.vdepth = rp.max_depth;
    adj.dldpx = (const float*)h->dldpx.ptr;
    adj.grad = lossgrad_dev + 1;
    const int slot = (int)(h->ev_count % SceneHandle::kEventRing);
    IPT_CUDA(cudaEventRecord(h->ev_start[slot], st));
    unsigned long long launches = 0;
    IPT_CUDA(launch_wavefront(sc, rp, h->wavefront.ptr, cap, nullptr, ctr, h->sm_count, h->host_live, &launches, st, &adj));
    IPT_CUDA(cudaEventRecord(h->ev_stop[slot], st));
    h->ev_count++;
    g_launches += launches;
    return 0;
  }
  const int k = pick_chunks(h, plan);
  GradLaunch g;
  if (prepare_grad(h, plan, k, &g) != 0) return -1;
  unsigned long long* ctr = (unsigned long long*)h->counters.ptr;
  IPT_CUDA(cudaMemsetAsync(ctr, 0, 16 * sizeof(unsigned long long), st));
  const int slot = (int)(h->ev_count % SceneHandle::kEventRing);
  IPT_CUDA(cudaEventRecord(h->ev_start[slot], st));
  Fork f(h, st, k);
  IPT_CUDA(f.begin());
  for (int c = 0; c < k; c++)
    if (grad_chunk(h, plan, sc, g, c, k, target_dev, lossgrad_dev, f.stream(c)) != 0) return -1;
  IPT_CUDA(f.end());
  IPT_CUDA(cudaEventRecord(h->ev_stop[slot], st));
  h->ev_count++;
  if (g.acc_mode) {
    IPT_CUDA(launch_reduce_blocks((const float*)h->block_acc.ptr, k * g.grid, 3 * nT, lossgrad_dev + 1, 0, 0, st));
    g_launches += 1;
  }
  return 0;
}

static uint64_t legacy_seed() {
  const char* s = getenv("IPT_SEED");
  if (s && *s) return strtoull(s, nullptr, 10);
  return (uint64_t)time(nullptr);   // path_trace.cu:206
}

}  // namespace ipt

using namespace ipt;

// No C++ exception may cross the C ABI (callers are ctypes / C): entry points that allocate on the host run their body
// through this guard, which turns std::bad_alloc and friends into the library's error state.
template <class F>
static int guarded(const char* what, F&& body) {
  try {
    return body();
  } catch (const std::exception& e) {
    return fail(std::string(what) + ": " + e.what());
  } catch (...) {
    return fail(std::string(what) + ": unknown C++ exception");
  }
}

extern "C" {

const char* iptGetLastError(void) { return g_error.c_str(); }

void iptDefaultParams(IptRenderParams* p) {
  if (!p) return;
  p->width = 500; p->height = 500; p->spp = 100;   // scene.h:8-10
  p->max_depth = 0;
  p->p_rr = 0.9f;                                  // scene.h:11
  p->seed_lo = 0; p->seed_hi = 0;
  p->group_samples = -1;
  p->row_begin = 0; p->row_end = 0; p->row_step = 1;
}

int iptSetAssetRoot(const char* root) { g_asset_root = root ? root : ""; return 0; }
int iptSetDevice(int device) { g_device = device; return 0; }
int iptSetBvhLargeTriangleLimit(int max_count) {
  if (max_count < 0 || max_count > kBruteForceMaxTris) return fail("large-triangle limit must be in [0, 128]");
  g_big_max = max_count;
  return 0;
}
int iptSceneLargeTriangles(void* scene) {
  SceneHandle* h = checked(scene);
  return h ? h->n_big : -1;
}
int iptSetIntegrator(int mode, int wavefront_slots) {
  if (mode < 0 || mode > 2) return fail("integrator must be 0, 1 or 2");
  if (wavefront_slots < 0) return fail("wavefront_slots must be >= 0");
  g_integrator = mode;
  g_wf_slots = wavefront_slots > 0 ? wavefront_slots : (1 << 24);
  return 0;
}
int iptSceneSetIntegrator(void* scene, int mode, int wavefront_slots) {
  SceneHandle* h = checked(scene);
  if (!h) return -1;
  if (mode < 0 || mode > 2) return fail("integrator must be 0, 1 or 2");
  if (wavefront_slots < 0) return fail("wavefront_slots must be >= 0");
  std::lock_guard<std::recursive_mutex> lk(h->mutex);
  h->integrator = mode;
  h->wf_slots = wavefront_slots > 0 ? wavefront_slots : (1 << 24);
  return 0;
}
int iptSetTraversalMode(int mode) {
  if (mode < 0 || mode > 2) return fail("traversal mode must be 0, 1 or 2");
  g_traversal_mode = mode;
  return 0;
}
double iptMeasureCopyGBps(long long bytes, int reps) {
  if (bytes < 16 || reps < 1) { fail("iptMeasureCopyGBps: bytes >= 16 and reps >= 1 required"); return -1.0; }
  const int dev = g_device >= 0 ? g_device : 0;
  if (cudaSetDevice(dev) != cudaSuccess) { fail("iptMeasureCopyGBps: no CUDA device"); return -1.0; }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t n = (size_t)bytes / 16;
  float4 *a = nullptr, *b = nullptr;
  if (cudaMalloc(&a, n * 16) != cudaSuccess || cudaMalloc(&b, n * 16) != cudaSuccess) {
    cudaFree(a);
    fail("iptMeasureCopyGBps: allocation failed");
    return -1.0;
  }
  cudaMemset(a, 0, n * 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; i++) launch_copy16(a, b, n, sms, nullptr);
  cudaEventRecord(e0, nullptr);
  for (int i = 0; i < reps; i++) launch_copy16(a, b, n, sms, nullptr);
  cudaEventRecord(e1, nullptr);
  const cudaError_t err = cudaEventSynchronize(e1);
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, e0, e1);
  g_launches += (unsigned long long)reps + 3;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(a); cudaFree(b);
  if (err != cudaSuccess || !(ms > 0.0f)) { fail(std::string("iptMeasureCopyGBps: ") + cudaGetErrorString(err)); return -1.0; }
  return 2.0 * (double)(n * 16) * reps / (ms * 1e-3) / 1e9;
}

unsigned long long iptKernelLaunchCount(void) { return g_launches.load() + bvh_kernel_launches(); }
int iptSetBvhBuilder(int mode) {
  if (mode < 0 || mode > 2) return fail("BVH builder mode must be 0 (auto), 1 (host) or 2 (GPU)");
  set_bvh_builder_mode(mode);
  return 0;
}

static int load_scene_impl(float** poss, float** oris, float** scls, char** obj_fs, char** mtl_fs, int n, void** scenePtr, SceneHandle** in_flight) {
  if (scenePtr) *scenePtr = nullptr;
  if (!scenePtr || n < 0 || (n > 0 && (!poss || !oris || !scls || !obj_fs || !mtl_fs))) { fail("loadScene: bad arguments"); return 0; }
  SceneHandle* h = new SceneHandle();
  *in_flight = h;   // the caller disposes of it if an exception unwinds this function
  h->traversal_mode = g_traversal_mode; h->big_max = g_big_max; h->integrator = g_integrator; h->wf_slots = g_wf_slots;
  std::string err;
  if (!build_host_scene(poss, oris, scls, obj_fs, mtl_fs, n, g_asset_root, &h->host, &err)) {
    fprintf(stderr, "ipt_b200: loadScene failed: %s\n", err.c_str());
    fail(err);
    delete h;
    *in_flight = nullptr;
    return 0;
  }
  int dev = g_device;
  if (dev == -2) {   // host-only handle: loader tables / materials without touching CUDA
    h->host_only = true;
    std::lock_guard<std::mutex> lk(g_registry_mutex);
    g_registry.insert(h);
    *scenePtr = h;
    *in_flight = nullptr;
    return (int)h->host.tris.size();
  }
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) {
    fail("ipt_b200: no CUDA device available (this library has no CPU fallback)");
    fprintf(stderr, "%s\n", g_error.c_str());
    delete h;
    *in_flight = nullptr;
    return 0;
  }
  h->device = dev;
  cudaError_t e = cudaSetDevice(dev);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->busy, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  for (int i = 0; i < SceneHandle::kMaxChunks && e == cudaSuccess; i++) {
    e = cudaStreamCreateWithFlags(&h->cstream[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming);
  }
  for (int i = 0; i < SceneHandle::kEventRing && e == cudaSuccess; i++) {
    e = cudaEventCreate(&h->ev_start[i]);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev_stop[i]);
  }
  if (e != cudaSuccess) {
    fail(std::string("ipt_b200: CUDA initialisation failed (no CPU fallback): ") + cudaGetErrorString(e));
    fprintf(stderr, "%s\n", g_error.c_str());
    delete h;
    *in_flight = nullptr;
    return 0;
  }
  if (upload_scene(h) != 0) {
    fprintf(stderr, "ipt_b200: scene upload failed: %s\n", g_error.c_str());
    h->destroy();
    delete h;
    *in_flight = nullptr;
    return 0;
  }
  {
    std::lock_guard<std::mutex> lk(g_registry_mutex);
    g_registry.insert(h);
  }
  *scenePtr = h;
  *in_flight = nullptr;
  return (int)h->host.tris.size();
}

int loadScene(float** poss, float** oris, float** scls, char** obj_fs, char** mtl_fs, int n, void** scenePtr) {
  SceneHandle* in_flight = nullptr;
  try {
    return load_scene_impl(poss, oris, scls, obj_fs, mtl_fs, n, scenePtr, &in_flight);
  } catch (const std::exception& e) {   // e.g. std::bad_alloc on a huge OBJ: no exception may cross the C ABI
    fail(std::string("loadScene: ") + e.what());
  } catch (...) {
    fail("loadScene: unknown C++ exception");
  }
  fprintf(stderr, "ipt_b200: %s\n", iptGetLastError());
  if (in_flight) { in_flight->destroy(); delete in_flight; }
  if (scenePtr) *scenePtr = nullptr;
  return 0;
}

void freeScene(void* scenePtr) {
  {
    std::lock_guard<std::mutex> lk(g_registry_mutex);
    if (!scenePtr || !g_registry.erase(scenePtr)) return;
  }
  SceneHandle* h = (SceneHandle*)scenePtr;
  h->destroy();
  delete h;
}

// albedos changed on the device (iptSetMaterialsDevice, iptInverseStep) are read back before the host table is used, so
// getMaterials always reflects the last set, like the reference's (inv_path_trace.cu:210-214)
static void sync_host_kd(SceneHandle* h) {
  if (!h->kd_dirty || h->host_only) return;
  const size_t nT = h->host.tris.size();
  std::vector<float4> kd(nT);
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  if (nT && cudaMemcpy(kd.data(), h->mat, nT * sizeof(float4), cudaMemcpyDeviceToHost) == cudaSuccess)
    for (size_t i = 0; i < nT; i++) { h->host.tris[i].mat.kd[0] = kd[i].x; h->host.tris[i].mat.kd[1] = kd[i].y; h->host.tris[i].mat.kd[2] = kd[i].z; }
  h->kd_dirty = false;
}

void getMaterials(void* scenePtr, float* materials) {
  guarded("getMaterials", [&]() -> int {
    SceneHandle* h = checked(scenePtr);
    if (!h || !materials) return 0;
    std::lock_guard<std::recursive_mutex> lk(h->mutex);
    sync_host_kd(h);
    for (size_t i = 0; i < h->host.tris.size(); i++)
      for (int j = 0; j < 3; j++) materials[3 * i + j] = h->host.tris[i].mat.kd[j];
    return 0;
  });
}

void setMaterials(void* scenePtr, float* materials) {
  guarded("setMaterials", [&]() -> int {
    SceneHandle* h = checked(scenePtr);
    if (!h || !materials) return 0;
    std::lock_guard<std::recursive_mutex> lk(h->mutex);
    for (size_t i = 0; i < h->host.tris.size(); i++)
      for (int j = 0; j < 3; j++) h->host.tris[i].mat.kd[j] = materials[3 * i + j];
    if (h->host_only) return 0;
    std::vector<float4> mat;
    pack_materials(h->host, &mat);
    // the legacy call is fully synchronous like the reference's (scene.h:154-162 writes managed memory between renders):
    // wait for everything queued on the device, on whatever stream, before the table is replaced
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (!mat.empty()) cudaMemcpy(h->mat, mat.data(), mat.size() * sizeof(float4), cudaMemcpyHostToDevice);
    h->kd_dirty = false;
    return 0;
  });
}

int iptSetMaterialsDevice(void* scene, const float* kd_dev, void* stream) {
  SceneHandle* h = checked_device(scene);
  if (!h || !kd_dev) return fail("iptSetMaterialsDevice: bad arguments");
  IPT_CUDA(cudaSetDevice(h->device));
  CallGuard guard(h, (cudaStream_t)stream);
  const int nT = (int)h->host.tris.size();
  IPT_CUDA(launch_set_kd(h->mat, kd_dev, nT, (cudaStream_t)stream));
  g_launches += 1;
  h->kd_dirty = true;   // getMaterials / the triangle table read the albedos back from the device
  return 0;
}

int iptSceneInfo(void* scene, int* n_triangles, int* n_emissives, int* n_objects, int* n_bvh_nodes, int* uses_bvh) {
  SceneHandle* h = checked(scene);
  if (!h) return -1;
  if (n_triangles) *n_triangles = (int)h->host.tris.size();
  if (n_emissives) *n_emissives = (int)h->host.emissives.size();
  if (n_objects) *n_objects = (int)h->host.objects.size();
  if (n_bvh_nodes) *n_bvh_nodes = h->n_nodes;
  if (uses_bvh) *uses_bvh = h->use_bvh ? 1 : 0;
  return 0;
}

int iptGetTriangleTable(void* scene, float* verts, float* kd, float* ks, float* ke, float* ns, int* has_vn, float* normal,
                        float* center, float* area) {
  return guarded("iptGetTriangleTable", [&]() -> int {
    SceneHandle* h = checked(scene);
    if (!h) return -1;
    std::lock_guard<std::recursive_mutex> lk(h->mutex);
    sync_host_kd(h);
    for (size_t i = 0; i < h->host.tris.size(); i++) {
      const HostTriangle& t = h->host.tris[i];
      for (int j = 0; j < 3; j++) {
        if (verts) for (int k = 0; k < 3; k++) verts[9 * i + 3 * j + k] = t.v[j][k];
        if (kd) kd[3 * i + j] = t.mat.kd[j];
        if (ks) ks[3 * i + j] = t.mat.ks[j];
        if (ke) ke[3 * i + j] = t.mat.ke[j];
        if (normal) normal[3 * i + j] = t.n[j];
        if (center) center[3 * i + j] = t.c[j];
      }
      if (ns) ns[i] = t.mat.ns;
      if (has_vn) has_vn[i] = t.has_vn ? 1 : 0;
      if (area) area[i] = t.area;
    }
    return 0;
  });
}

int iptGetCamera(void* scene, float* m16) {
  SceneHandle* h = checked(scene);
  if (!h || !m16) return -1;
  memcpy(m16, h->host.cam, 16 * sizeof(float));
  return 0;
}

int iptTraceRaysDevice(void* scene, const float* rays_dev, int n, int* tri_dev, float* t_dev, void* stream) {
  SceneHandle* h = checked_device(scene);
  if (!h) return -1;
  if (n < 0 || (n > 0 && (!rays_dev || !tri_dev || !t_dev))) return fail("iptTraceRaysDevice: bad arguments");
  if (n == 0) return 0;
  IPT_CUDA(cudaSetDevice(h->device));
  CallGuard guard(h, (cudaStream_t)stream);
  IPT_CUDA(h->trace_counter.reserve(sizeof(unsigned long long)));
  IPT_CUDA(launch_trace_rays(h->dev(), rays_dev, n, tri_dev, t_dev, (unsigned long long*)h->trace_counter.ptr, h->sm_count, (cudaStream_t)stream));
  g_launches += (h->use_bvh && h->n_big > 0) ? 2 : 1;
  return 0;
}

int iptTraceRaysStatsDevice(void* scene, const float* rays_dev, int n, int* tri_dev, float* t_dev, unsigned long long* stats_dev, void* stream) {
  SceneHandle* h = checked_device(scene);
  if (!h) return -1;
  if (n <= 0 || !rays_dev || !tri_dev || !t_dev || !stats_dev) return fail("iptTraceRaysStatsDevice: bad arguments");
  if (!h->use_bvh) return fail("iptTraceRaysStatsDevice: the scene is in brute-force mode (no BVH to instrument)");
  IPT_CUDA(cudaSetDevice(h->device));
  CallGuard guard(h, (cudaStream_t)stream);
  IPT_CUDA(cudaMemsetAsync(stats_dev, 0, 2 * sizeof(unsigned long long), (cudaStream_t)stream));
  IPT_CUDA(h->trace_counter.reserve(sizeof(unsigned long long)));
  IPT_CUDA(launch_trace_rays_stats(h->dev(), rays_dev, n, tri_dev, t_dev, (unsigned long long*)h->trace_counter.ptr, stats_dev, h->sm_count,
                                   (cudaStream_t)stream));
  g_launches += (h->use_bvh && h->n_big > 0) ? 2 : 1;
  return 0;
}

int iptTraceRays(void* scene, const float* rays, int n, int* tri, float* t) {
  return guarded("iptTraceRays", [&]() -> int {
    SceneHandle* h = checked_device(scene);
    if (!h) return -1;
    if (n < 0 || (n > 0 && (!rays || !tri || !t))) return fail("iptTraceRays: bad arguments");
    if (n == 0) return 0;
    std::lock_guard<std::recursive_mutex> lk(h->mutex);
    IPT_CUDA(cudaSetDevice(h->device));
    IPT_CUDA(h->rays.reserve((size_t)n * 6 * sizeof(float)));
    IPT_CUDA(h->hits_tri.reserve((size_t)n * sizeof(int)));
    IPT_CUDA(h->hits_t.reserve((size_t)n * sizeof(float)));
    IPT_CUDA(cudaMemcpyAsync(h->rays.ptr, rays, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    if (iptTraceRaysDevice(scene, (const float*)h->rays.ptr, n, (int*)h->hits_tri.ptr, (float*)h->hits_t.ptr, h->stream) != 0) return -1;
    IPT_CUDA(cudaMemcpyAsync(tri, h->hits_tri.ptr, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    IPT_CUDA(cudaMemcpyAsync(t, h->hits_t.ptr, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    IPT_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
  });
}

int iptRenderDevice(void* scene, const IptRenderParams* p, float* img_dev, unsigned char* img8_dev, unsigned long long* counters_dev,
                    void* stream) {
  SceneHandle* h = checked_device(scene);
  if (!h) return -1;
  Plan plan;
  if (make_plan(h, p, &plan) != 0) return -1;
  IPT_CUDA(cudaSetDevice(h->device));
  CallGuard guard(h, (cudaStream_t)stream);
  return launch_forward(h, plan, img_dev, img8_dev, counters_dev, (cudaStream_t)stream);
}

int iptRender(void* scene, const IptRenderParams* p, float* img, unsigned char* img8, unsigned long long* counters) {
  return guarded("iptRender", [&]() -> int {
    SceneHandle* h = checked_device(scene);
    if (!h) return -1;
    Plan plan;
    if (make_plan(h, p, &plan) != 0) return -1;
    std::lock_guard<std::recursive_mutex> lk(h->mutex);
    IPT_CUDA(cudaSetDevice(h->device));
    const size_t n3 = (size_t)plan.rp.n_pix * 3;
    IPT_CUDA(h->img.reserve(n3 * sizeof(float)));
    IPT_CUDA(h->img8.reserve(n3));
    IPT_CUDA(h->cnt_out.reserve(4 * sizeof(unsigned long long)));
    unsigned long long* cdev = (unsigned long long*)h->cnt_out.ptr;
    if (launch_forward(h, plan, img ? (float*)h->img.ptr : nullptr, img8 ? (unsigned char*)h->img8.ptr : nullptr, cdev, h->stream) != 0) return -1;
    if (img) IPT_CUDA(cudaMemcpyAsync(img, h->img.ptr, n3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (img8) IPT_CUDA(cudaMemcpyAsync(img8, h->img8.ptr, n3, cudaMemcpyDeviceToHost, h->stream));
    if (counters) IPT_CUDA(cudaMemcpyAsync(counters, cdev, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    IPT_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
  });
}

int iptCreateImageEx(void* scene, const IptRenderParams* p, const char* png_path) {
  return guarded("iptCreateImageEx", [&]() -> int {
    if (!p || !png_path) return fail("iptCreateImageEx: bad arguments");
    IptRenderParams q = *p;
    q.row_begin = 0;
    q.row_end = 0;
    q.row_step = 1;
    std::vector<unsigned char> px((size_t)q.width * q.height * 3);
    if (iptRender(scene, &q, nullptr, px.data(), nullptr) != 0) return -1;
    std::string err;
    if (!write_png_rgb8(png_path, px.data(), q.width, q.height, &err)) return fail(err);
    return 0;
  });
}

void createImage(void* scenePtr, char* img_file) {
  guarded("createImage", [&]() -> int {
    IptRenderParams p;
    iptDefaultParams(&p);
    const uint64_t seed = legacy_seed();
    p.seed_lo = (uint32_t)seed;
    p.seed_hi = (uint32_t)(seed >> 32);
    if (iptCreateImageEx(scenePtr, &p, img_file) != 0) fprintf(stderr, "ipt_b200: createImage failed: %s\n", g_error.c_str());
    return 0;
  });
}

int iptGraphAccumulateDevice(void* scene, const IptRenderParams* p, const unsigned char* target_dev, double* raw_dev, void* stream) {
  SceneHandle* h = checked_device(scene);
  if (!h) return -1;
  if (!target_dev || !raw_dev) return fail("iptGraphAccumulateDevice: bad arguments");
  Plan plan;
  if (make_plan(h, p, &plan) != 0) return -1;
  IPT_CUDA(cudaSetDevice(h->device));
  CallGuard guard(h, (cudaStream_t)stream);
  cudaStream_t st = (cudaStream_t)stream;
  const DevScene sc = h->dev();
  const int nT = sc.nT;
  // the edge matrix is dense, (nT + 1) x nT (inv_scene.h:52-65): refuse scenes whose matrix does not fit 32-bit indexing
  // of the accumulators (nT <= 14 000 keeps (nT + 1) nT x 10 floats below 2^31) instead of wrapping
  if ((size_t)(nT + 1) * (size_t)nT * kGraphAccPerEdge >= ((size_t)1 << 31))
    return fail("createGraph: the scene is too large for the dense (nT + 1) x nT edge matrix of the reference (nT must stay below 14 000)");
  const int n_edges = (nT + 1) * nT;
  const int acc_floats = n_edges * kGraphAccPerEdge;
  IPT_CUDA(cudaMemsetAsync(raw_dev, 0, (size_t)n_edges * 16 * sizeof(double), st));
  if (nT == 0) return 0;
  IPT_CUDA(h->counters.reserve(16 * sizeof(unsigned long long)));
  unsigned long long* ctr = (unsigned long long*)h->counters.ptr;
  IPT_CUDA(cudaMemsetAsync(ctr, 0, 16 * sizeof(unsigned long long), st));
  IPT_CUDA(h->acc.reserve((size_t)acc_floats * sizeof(double)));
  IPT_CUDA(cudaMemsetAsync(h->acc.ptr, 0, (size_t)acc_floats * sizeof(double), st));
  DevOut out;
  memset(&out, 0, sizeof(out));
  out.work_counter = reinterpret_cast<unsigned*>(ctr + 4);
  out.counters = ctr;
  out.target = target_dev;
  out.acc_stride = acc_floats;
  // 1: edge tile in shared memory (float atomics there are CAS loops); 2: one tile per block in global memory,
  // updated with float reductions; 0: double atomics on one global array (scenes whose tiles are too large)
  const char* acc_env = getenv("IPT_GRAPH_ACC");   // developer A/B
  size_t smem = paths_smem_bytes(sc, acc_floats);
  const int grid_if_global = h->sm_count * paths_blocks_per_sm(MODE_GRAPH_ID, sc, paths_smem_bytes(sc, 0));
  if ((size_t)grid_if_global * acc_floats * sizeof(float) <= ((size_t)128 << 20)) out.acc_in_smem = 2;   // measured 9 % faster than 1
  else out.acc_in_smem = smem <= 160 * 1024 ? 1 : 0;
  if (acc_env && *acc_env) out.acc_in_smem = atoi(acc_env);
  if (out.acc_in_smem != 1) smem = paths_smem_bytes(sc, 0);
  const int grid = h->sm_count * paths_blocks_per_sm(MODE_GRAPH_ID, sc, smem);
  if (out.acc_in_smem) {
    IPT_CUDA(h->block_acc.reserve((size_t)grid * acc_floats * sizeof(float)));
    out.block_acc = (float*)h->block_acc.ptr;
    if (out.acc_in_smem == 2) IPT_CUDA(cudaMemsetAsync(out.block_acc, 0, (size_t)grid * acc_floats * sizeof(float), st));
  } else {
    out.global_acc = (double*)h->acc.ptr;
  }
  {
    const int slot = (int)(h->ev_count % SceneHandle::kEventRing);
    IPT_CUDA(cudaEventRecord(h->ev_start[slot], st));
    NvtxRange r("ipt.paths.graph");
    IPT_CUDA(launch_paths(MODE_GRAPH_ID, sc, plan.rp, out, grid, smem, st));
    IPT_CUDA(cudaEventRecord(h->ev_stop[slot], st));
    h->ev_count++;
  }
  g_launches += 1;
  if (out.acc_in_smem) {
    IPT_CUDA(launch_reduce_blocks(out.block_acc, grid, acc_floats, (double*)h->acc.ptr, 8, kGraphAccPerEdge, st));
    g_launches += 1;
  }
  IPT_CUDA(launch_graph_finalize((const double*)h->acc.ptr, h->mat, nT, n_edges, raw_dev, st));
  g_launches += 1;
  return 0;
}

// DataWrapper::compress + Edge::normalize (inv_scene.h:38-47, 87-115) on the raw sums
int iptGraphCompress(const double* raw, int nT, float* data) {
  if (!raw || !data || nT < 0) return fail("iptGraphCompress: bad arguments");
  const size_t size = (size_t)(nT + 1) * nT;
  float* weights = data;
  float* pixels = data + size;
  float* lights = data + 4 * size;
  std::vector<float> row(nT);
  for (int dst = 0; dst <= nT; dst++) {
    float total = 0.0f;
    for (int src = 0; src < nT; src++) {
      const double* e = raw + ((size_t)dst * nT + src) * 16;
      const float w = canon_logf((float)e[1] + 1.0f);
      const float fs = (float)e[14];
      const float div = (fs != 0.0f) ? fs : 1.0f;
      row[src] = w;
      total += w;
      for (int i = 0; i < 3; i++) {
        pixels[((size_t)dst * nT + src) * 3 + i] = (float)e[2 + i] / div;
        lights[((size_t)dst * nT + src) * 3 + i] = (float)e[8 + i] / div;
      }
    }
    for (int src = 0; src < nT; src++) weights[(size_t)dst * nT + src] = (total != 0.0f) ? row[src] / total : 0.0f;
  }
  return 0;
}

int iptGraphCompressDevice(const double* raw_dev, int nT, float* data_dev, void* stream) {
  if (!raw_dev || !data_dev || nT < 0) return fail("iptGraphCompressDevice: bad arguments");
  if (nT == 0) return 0;
  IPT_CUDA(launch_graph_compress(raw_dev, nT, data_dev, (cudaStream_t)stream));
  g_launches += 1;
  return 0;
}

int iptCreateGraphEx(void* scene, const IptRenderParams* p, const unsigned char* target, double* raw, float* data,
                     unsigned long long* counters) {
  return guarded("iptCreateGraphEx", [&]() -> int {
    SceneHandle* h = checked_device(scene);
    if (!h) return -1;
    if (!p || !target) return fail("iptCreateGraphEx: bad arguments");
    std::lock_guard<std::recursive_mutex> lk(h->mutex);
    IPT_CUDA(cudaSetDevice(h->device));
    const int nT = (int)h->host.tris.size();
    const size_t n_edges = (size_t)(nT + 1) * nT;
    const size_t tbytes = (size_t)p->width * p->height * 3;
    IPT_CUDA(h->target.reserve(tbytes));
    IPT_CUDA(h->raw.reserve(std::max<size_t>(n_edges, 1) * 16 * sizeof(double)));
    IPT_CUDA(cudaMemcpyAsync(h->target.ptr, target, tbytes, cudaMemcpyHostToDevice, h->stream));
    if (iptGraphAccumulateDevice(scene, p, (const unsigned char*)h->target.ptr, (double*)h->raw.ptr, h->stream) != 0) return -1;
    std::vector<double> host_raw(n_edges * 16);
    if (n_edges) IPT_CUDA(cudaMemcpyAsync(host_raw.data(), h->raw.ptr, n_edges * 16 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (counters) IPT_CUDA(cudaMemcpyAsync(counters, h->counters.ptr, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    IPT_CUDA(cudaStreamSynchronize(h->stream));
    if (raw && n_edges) memcpy(raw, host_raw.data(), n_edges * 16 * sizeof(double));
    if (data) return iptGraphCompress(host_raw.data(), nT, data);
    return 0;
  });
}

void createGraph(void* scenePtr, char* imgFile, float* data) {
  guarded("createGraph", [&]() -> int {
    SceneHandle* h = checked_device(scenePtr);
    if (!h || !imgFile || !data) { fprintf(stderr, "ipt_b200: createGraph: bad arguments\n"); return 0; }
    IptRenderParams p;
    iptDefaultParams(&p);
    const uint64_t seed = legacy_seed();
    p.seed_lo = (uint32_t)seed;
    p.seed_hi = (uint32_t)(seed >> 32);
    std::vector<unsigned char> target;
    int w = 0, hgt = 0;
    std::string err;
    if (!read_png_rgb8(imgFile, &target, &w, &hgt, &err)) { fail(err); fprintf(stderr, "ipt_b200: createGraph: %s\n", err.c_str()); return 0; }
    if (w != p.width || hgt != p.height) {   // the reference silently assumes 500x500 (inv_scene.h:57)
      fail("createGraph: target image must be 500x500");
      fprintf(stderr, "ipt_b200: %s (got %dx%d)\n", g_error.c_str(), w, hgt);
      return 0;
    }
    if (iptCreateGraphEx(scenePtr, &p, target.data(), nullptr, data, nullptr) != 0) fprintf(stderr, "ipt_b200: createGraph failed: %s\n", g_error.c_str());
    return 0;
  });
}

int iptRenderGradDevice(void* scene, const IptRenderParams* p, const unsigned char* target_dev, double* lossgrad_dev, float* img_dev,
                        void* stream) {
  SceneHandle* h = checked_device(scene);
  if (!h) return -1;
  if (!p || !target_dev || !lossgrad_dev) return fail("iptRenderGradDevice: bad arguments");
  IptRenderParams q = *p;
  if (q.max_depth <= 0 || q.max_depth > kGradMaxDepth) q.max_depth = kGradMaxDepth;
  Plan plan;
  if (make_plan(h, &q, &plan) != 0) return -1;
  IPT_CUDA(cudaSetDevice(h->device));
  CallGuard guard(h, (cudaStream_t)stream);
  cudaStream_t st = (cudaStream_t)stream;
  const DevScene sc = h->dev();
  const DevRender& rp = plan.rp;
  const int nT = sc.nT;
  const size_t n3 = (size_t)rp.n_pix * 3;
  float* img = img_dev;
  if (!img) {
    IPT_CUDA(h->img.reserve(n3 * sizeof(float)));
    img = (float*)h->img.ptr;
  }
  if (launch_forward(h, plan, img, nullptr, nullptr, st) != 0) return -1;
  IPT_CUDA(cudaMemsetAsync(lossgrad_dev, 0, (size_t)(1 + 3 * nT) * sizeof(double), st));
  IPT_CUDA(h->dldpx.reserve(n3 * sizeof(float)));
  const double inv_n = 1.0 / (3.0 * (double)rp.width * (double)rp.height);
  IPT_CUDA(launch_loss_dldpx(img, target_dev, rp.row_begin, rp.row_step, rp.width, rp.n_pix, inv_n, 1.0f / (float)rp.spp, (float*)h->dldpx.ptr, lossgrad_dev, st));
  g_launches += 1;
  return launch_grad_pass(h, plan, target_dev, lossgrad_dev, st);
}

int iptRenderGrad(void* scene, const IptRenderParams* p, const unsigned char* target, double* loss, double* grad, float* img) {
  return guarded("iptRenderGrad", [&]() -> int {
    SceneHandle* h = checked_device(scene);
    if (!h) return -1;
    if (!p || !target || !loss || !grad) return fail("iptRenderGrad: bad arguments");
    std::lock_guard<std::recursive_mutex> lk(h->mutex);
    IPT_CUDA(cudaSetDevice(h->device));
    const int nT = (int)h->host.tris.size();
    Plan plan;
    if (make_plan(h, p, &plan) != 0) return -1;
    const size_t tbytes = (size_t)p->width * p->height * 3;
    const size_t n3 = (size_t)plan.rp.n_pix * 3;
    IPT_CUDA(h->target.reserve(tbytes));
    IPT_CUDA(h->lossgrad.reserve((size_t)(1 + 3 * nT) * sizeof(double)));
    IPT_CUDA(h->img.reserve(n3 * sizeof(float)));
    IPT_CUDA(cudaMemcpyAsync(h->target.ptr, target, tbytes, cudaMemcpyHostToDevice, h->stream));
    if (iptRenderGradDevice(scene, p, (const unsigned char*)h->target.ptr, (double*)h->lossgrad.ptr, (float*)h->img.ptr, h->stream) != 0) return -1;
    std::vector<double> lg(1 + 3 * (size_t)nT);
    IPT_CUDA(cudaMemcpyAsync(lg.data(), h->lossgrad.ptr, lg.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (img) IPT_CUDA(cudaMemcpyAsync(img, h->img.ptr, n3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    IPT_CUDA(cudaStreamSynchronize(h->stream));
    *loss = lg[0];
    memcpy(grad, lg.data() + 1, 3 * (size_t)nT * sizeof(double));
    return 0;
  });
}

int iptPathKernelTimes(void* scene, float* ms, int max_count) {
  SceneHandle* h = checked_device(scene);
  if (!h || !ms || max_count < 0) return -1;
  IPT_CUDA(cudaSetDevice(h->device));
  const unsigned long long have = h->ev_count < (unsigned long long)SceneHandle::kEventRing ? h->ev_count : SceneHandle::kEventRing;
  const int n = (int)((unsigned long long)max_count < have ? max_count : have);
  for (int i = 0; i < n; i++) {   // ms[0] = most recent launch
    const int slot = (int)((h->ev_count - 1 - i) % SceneHandle::kEventRing);
    IPT_CUDA(cudaEventSynchronize(h->ev_stop[slot]));
    IPT_CUDA(cudaEventElapsedTime(&ms[i], h->ev_start[slot], h->ev_stop[slot]));
  }
  return n;
}

int iptLastCounters(void* scene, unsigned long long* counters) {
  SceneHandle* h = checked_device(scene);
  if (!h || !counters) return -1;
  if (!h->counters.ptr) return fail("no path kernel has been launched on this scene yet");
  IPT_CUDA(cudaSetDevice(h->device));
  IPT_CUDA(cudaMemcpy(counters, h->counters.ptr, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return 0;
}

int iptBuildBVH(const float* boxes, const float* pos, int n, int leaf_size, float* node_boxes, int* node_ints, int max_nodes, int* perm) {
  return guarded("iptBuildBVH", [&]() -> int {
    if (n < 0 || (n > 0 && (!boxes || !pos || !perm)) || !node_boxes || !node_ints) return fail("iptBuildBVH: bad arguments");
    std::vector<BvhNodeHost> nodes;
    std::vector<int> pm;
    std::string err;
    int dev = g_device;
    if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) dev = -1;
    if (!build_bvh(boxes, pos, n, leaf_size, &nodes, &pm, dev, &err)) return fail(err);
    if ((int)nodes.size() > max_nodes) return fail("iptBuildBVH: max_nodes too small");
    for (size_t i = 0; i < nodes.size(); i++) {
      for (int k = 0; k < 3; k++) { node_boxes[6 * i + k] = nodes[i].lo[k]; node_boxes[6 * i + 3 + k] = nodes[i].hi[k]; }
      node_ints[3 * i] = nodes[i].start; node_ints[3 * i + 1] = nodes[i].nPrims; node_ints[3 * i + 2] = nodes[i].rightOffset;
    }
    for (int i = 0; i < n; i++) perm[i] = pm[i];
    return (int)nodes.size();
  });
}

int iptWritePNG(const char* path, const unsigned char* rgb, int width, int height) {
  return guarded("iptWritePNG", [&]() -> int {
    std::string err;
    if (!path || !rgb || width <= 0 || height <= 0) return fail("iptWritePNG: bad arguments");
    if (!write_png_rgb8(path, rgb, width, height, &err)) return fail(err);
    return 0;
  });
}

int iptReadPNG(const char* path, unsigned char* rgb, int capacity_bytes, int* width, int* height) {
  return guarded("iptReadPNG", [&]() -> int {
    std::string err;
    std::vector<unsigned char> px;
    int w = 0, h = 0;
    if (!path) return fail("iptReadPNG: bad arguments");
    if (!read_png_rgb8(path, &px, &w, &h, &err)) return fail(err);
    if (width) *width = w;
    if (height) *height = h;
    if (rgb) {
      if ((size_t)capacity_bytes < px.size()) return fail("iptReadPNG: buffer too small");
      memcpy(rgb, px.data(), px.size());
    }
    return 0;
  });
}

}  // extern "C"
