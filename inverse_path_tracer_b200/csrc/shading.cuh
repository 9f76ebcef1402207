// shading.cuh -- BSDF / normal helpers shared by k_paths (paths.cu) and the wavefront shading kernel (wavefront.cu)
#pragma once
#include "intersect.cuh"

namespace ipt {

// ------------------------------------------------------------------------------------------------
// shading helpers
// ------------------------------------------------------------------------------------------------
// Triangle::getNormal (scene_basics.h:100-109) for triangles with per-vertex normals.
static __device__ __noinline__ f3 interpolated_normal(const float4* __restrict__ vtx, int tri, f3 p) {
  const float4 a = __ldg(&vtx[5 * tri + 0]), b = __ldg(&vtx[5 * tri + 1]), c = __ldg(&vtx[5 * tri + 2]);
  const float4 n1 = __ldg(&vtx[5 * tri + 3]), n2 = __ldg(&vtx[5 * tri + 4]);
  const f3 v0 = xyz(a), v1 = xyz(b), v2 = xyz(c);
  const f3 n0 = mk3(a.w, b.w, c.w);
  const float area = n1.w;
  const float w0 = (0.5f * norm3(cross3(v1 - p, v2 - p))) / area;
  const float w1 = (0.5f * norm3(cross3(v2 - p, v0 - p))) / area;
  const float w2 = (0.5f * norm3(cross3(v0 - p, v1 - p))) / area;
  f3 n = mk3(fmaf(n2.x, w2, fmaf(n1.x, w1, n0.x * w0)), fmaf(n2.y, w2, fmaf(n1.y, w1, n0.y * w0)), fmaf(n2.z, w2, fmaf(n1.z, w1, n0.z * w0)));
  return normalize3(n);
}

__device__ __forceinline__ f3 shading_normal(const DevScene& sc, int tri, int flags, f3 face_n, f3 p) {
  if (flags & TRI_HAS_VN) return interpolated_normal(sc.vtx, tri, p);
  return face_n;
}

// Phong lobe of BSDF (path_trace.cu:19-23); only evaluated for materials with Ks != 0.
static __device__ __noinline__ float phong_coeff(f3 nrm, f3 w, f3 w_i, float n) {
  const float k = 2.0f * dot3(nrm, w_i);
  const f3 refl = mk3(fmaf(k, nrm.x, -w_i.x), fmaf(k, nrm.y, -w_i.y), fmaf(k, nrm.z, -w_i.z));
  const float pw = canon_pow(dot3(refl, w), n);
  return ((n + 2.0f) / 2.0f / IPT_PI_F) * fmaxf(pw, 0.0f);
}

// specular branch of sampleNextDir (path_trace.cu:94,108)
static __device__ __noinline__ f3 sample_specular_dir(f3 N, float shininess, float u_phi, float u_theta, float* p_sample) {
  float sphi, cphi;
  sincos2pi(u_phi, &sphi, &cphi);
  const float cth = canon_pow(u_theta, 1.0f / (shininess + 1.0f));
  const float sth = sqrtf(fmaxf(0.0f, 1.0f - cth * cth));
  const f3 hemi = mk3(sth * cphi, sth * sphi, cth);
  const f3 nd = (N.z == -1.0f) ? -hemi : rotate_from_z(N, hemi);
  *p_sample = canon_pow((shininess + 1.0f) * cth, shininess);
  return normalize3(nd);
}

// directLighting's area-proportional pick (path_trace.cu:39-51): the first emissive triangle whose cumulative
// probability reaches u, the last one if none does.  Same index as the reference's linear scan; emissive meshes
// (thousands of lights) are searched by bisection because the cdf is non-decreasing.
__device__ __forceinline__ int pick_light(const float4* __restrict__ lights, int nE, float u) {
  if (nE <= 16) {
    int li = 0;
    while (li < nE - 1 && !(lights[3 * li].w >= u)) li++;
    return li;
  }
  int lo = 0, hi = nE - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (lights[3 * mid].w >= u) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// lights are staged into shared memory by k_paths when they fit this budget (512 emissive triangles)
constexpr int kLightSmemMax = 512;

}  // namespace ipt
