// wavefront.cu -- forward integrator for BVH scenes: path state in HBM, one traversal per bounce (trace.cu)
#include <algorithm>

#include "intersect.cuh"
#include "kernels.h"
#include "nvtx_ranges.h"
#include "shading.cuh"

namespace ipt {

// ------------------------------------------------------------------------------------------------
// Wavefront forward integrator for BVH scenes.
//
// k_paths keeps a whole path in registers, which is the right shape for scenes that fit shared memory, but
// with a BVH the traversal then runs at the occupancy of a 96-register kernel and every warp mixes rays of
// all bounce depths.  For large scenes the forward pass is split instead: one light traversal kernel
// (k_trace_rays_bvh, the persistent traversal of the ray-batch entry point) between shading kernels that keep
// the path state in HBM, compacted every bounce.  A batch is `tile_n` pixels x `ns` consecutive samples, all
// at the same bounce; arithmetic per path and the order of every sum are those of k_paths, so the image is
// bit-identical to the megakernel's and to the oracle's.
//
//   live slot j of a bounce:  state[k * cap + j], k = 0..5 (float4 each)
//     S0 = {multiplier, path index in batch}   S1 = {prevMultiplier, flags}     S2 = {L, cos(theta) of the shadow ray}
//     S3 = {L_e, p_t of the sampled light}      S4 = {L_d, sampled light id}     S5 = {BSDF(direct) of the bounce, -}
//   rays[6 * j] extension ray, rays[6 * (cap + j)] shadow ray of the previous bounce (direction 0 = no ray)
//   hits likewise at j and cap + j; the kernel that writes a ray also writes its hit among the scene's brute-force set there
//   (two hit buffers, swapped with the rays), and k_trace_rays_bvh continues from it.
// ------------------------------------------------------------------------------------------------
enum { WF_END_AFTER = 1, WF_SH_VALID = 2, WF_EXT_VALID = 4 };
#ifndef IPT_WF_BLOCK
#define IPT_WF_BLOCK 256
#endif
#ifndef IPT_WF_MINB
#define IPT_WF_MINB 3         // k_wf_shade, scenes with a Phong lobe / vertex normals: 80 registers (4 blocks = 64 registers spill 300 bytes per thread)
#endif
#ifndef IPT_WF_MINB_PLAIN
#define IPT_WF_MINB_PLAIN 3   // plain scenes: 80 registers, no spills
#endif
constexpr int kWfBlock = IPT_WF_BLOCK;

// Adjoint recording (wavefront form of k_paths<GRAD>): the replay of a batch writes one vertex per
// (bounce, path) -- V[(bounce * cap + path) * vf4 ..], vf4 = 2 float4 for plain scenes ({tri, b, light, g} {P}), 4 with a
// Phong lobe / vertex normals (+ {T} {BSDF direct}) -- and the number of vertices of every path; k_wf_sweep then runs
// Pbar_k = C_k + T_k Pbar_{k+1} backwards over each path.  V == nullptr: plain forward pass.
struct WfRecord {
  float4* V;
  int* vlen;
  int vf4;
};

struct WfBatch {
  int cap;          // slots per array
  int tile_p0;      // first local pixel of the tile
  int tile_n;       // pixels in the tile
  int s0;           // first sample of the batch
  int ns;           // samples per pixel in the batch
};

__device__ __forceinline__ void wf_path_ids(const DevRender& rp, const WfBatch& b, int path, int* pixel, int* sample, float* pix_c, float* pix_r) {
  const int ss = path / b.tile_n;
  const int pl = b.tile_p0 + (path - ss * b.tile_n);
  const int lr = pl / rp.width;
  const int c = pl - lr * rp.width;
  const int r = rp.row_begin + lr * rp.row_step;
  *pixel = r * rp.width + c;
  *sample = b.s0 + ss;
  *pix_c = (float)c;
  *pix_r = (float)r;
}

__device__ __forceinline__ void wf_store_ray(float* __restrict__ rays, size_t idx, f3 p, f3 d) {
  float2* r2 = reinterpret_cast<float2*>(rays + 6 * idx);
  r2[0] = make_float2(p.x, p.y); r2[1] = make_float2(p.z, d.x); r2[2] = make_float2(d.y, d.z);
}
__device__ __forceinline__ void wf_load_ray(const float* __restrict__ rays, size_t idx, f3* p, f3* d) {
  const float2* r2 = reinterpret_cast<const float2*>(rays + 6 * idx);
  const float2 a = r2[0], b = r2[1], c = r2[2];
  *p = mk3(a.x, a.y, b.x); *d = mk3(b.y, c.x, c.y);
}

// camera rays of a batch (path_trace.cu:155-165): path i = ss * tile_n + pixel-in-tile, live slot i
__global__ void __launch_bounds__(kWfBlock) k_wf_generate(DevScene sc, DevRender rp, WfBatch b, float4* __restrict__ state, float* __restrict__ rays,
                                                          int* __restrict__ hit_tri, float* __restrict__ hit_t, int* __restrict__ n_live,
                                                          unsigned long long* __restrict__ counters) {
  // like k_wf_shade, the kernel leaves the camera rays' hits among the scene's brute-force set for the traversal to start from
  extern __shared__ float4 wf_smem[];
  float4* wf_after;
  const SmallScene ss = stage_small_scene(sc, wf_smem, &wf_after);
  __syncthreads();
  const int n = b.tile_n * b.ns;
  const int cap = b.cap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int pixel, cur_s;
    float pix_c, pix_r;
    wf_path_ids(rp, b, i, &pixel, &cur_s, &pix_c, &pix_r);
    const u4 rn = philox4x32_10((uint32_t)pixel, (uint32_t)cur_s, 0u, IPT_STREAM_CAMERA, rp.keys);
    const float x = fmaf(pix_c + u01(rn.x), rp.two_over_w, -1.0f);
    const float y = fmaf(pix_r + u01(rn.y), -rp.two_over_h, 1.0f);
    const f3 dl = mk3(x, y, 1.0f);
    const f3 rpos = mk3(sc.cam[3], sc.cam[7], sc.cam[11]);
    const f3 rdir = normalize3(mk3(dot3(sc.cam[0], sc.cam[1], sc.cam[2], dl), dot3(sc.cam[4], sc.cam[5], sc.cam[6], dl), dot3(sc.cam[8], sc.cam[9], sc.cam[10], dl)));
    state[0 * (size_t)cap + i] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(i));
    state[1 * (size_t)cap + i] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(WF_EXT_VALID));
    state[2 * (size_t)cap + i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    state[3 * (size_t)cap + i] = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    state[4 * (size_t)cap + i] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
    state[5 * (size_t)cap + i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    wf_store_ray(rays, (size_t)i, rpos, rdir);
    wf_store_ray(rays, (size_t)cap + i, rpos, mk3(0.0f, 0.0f, 0.0f));
    if (ss.nGroups > 0) {
      HitRec he, hs;
      closest_hit_small2(ss, rpos, rdir, rpos, mk3(0.0f, 0.0f, 0.0f), he, hs);
      hit_tri[i] = he.tri; hit_t[i] = he.t;
      hit_tri[(size_t)cap + i] = hs.tri; hit_t[(size_t)cap + i] = hs.t;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *n_live = n;
    if (counters) { atomicAdd(counters + 0, (unsigned long long)n); atomicAdd(counters + 1, (unsigned long long)n); }
  }
}

// one bounce of shading for every live path: second half of directLighting of the previous bounce
// ([D] of k_paths), then radiance() of this bounce's hit ([A] of k_paths); survivors are appended to the
// next bounce's arrays
template <bool PLAIN, bool RECORD>
__global__ void __launch_bounds__(kWfBlock, PLAIN ? IPT_WF_MINB_PLAIN : IPT_WF_MINB) k_wf_shade(DevScene sc, DevRender rp, WfBatch b, int bounce, const float4* __restrict__ sin_, float4* __restrict__ sout,
                                                       const float* __restrict__ rin, float* __restrict__ rout, const int* __restrict__ hit_tri,
                                                       const float* __restrict__ hit_t, int* __restrict__ hit_tri_out, float* __restrict__ hit_t_out,
                                                       const int* __restrict__ n_in, int* __restrict__ n_out,
                                                       float* __restrict__ lsamp, unsigned long long* __restrict__ counters, WfRecord rec) {
  // the scene's brute-force set (whole small scene / large triangles of a BVH scene) in shared memory: the kernel tests the two
  // rays it writes for a surviving path against it right away (the dual-ray loop of k_paths) and stores the hits where the next
  // bounce's traversal picks them up as its starting bound -- the separate k_trace_rays pass per bounce (2.4 of 29 ms on the
  // 5.2 M-triangle scene: a re-read of every ray and nine launches) is only left for the camera rays
  extern __shared__ float4 wf_smem[];
  float4* wf_after;
  const SmallScene ss = stage_small_scene(sc, wf_smem, &wf_after);
  __syncthreads();
  const int nT = sc.nT, nE = sc.nE;
  const size_t cap = (size_t)b.cap;
  const int n = *n_in;
  const float4* __restrict__ mat = sc.mat;
  const float4* __restrict__ slight = sc.light;
  const f3 zero3 = mk3(0.0f, 0.0f, 0.0f);
  const unsigned lane = threadIdx.x & 31u;
  unsigned n_ext = 0, n_shadow = 0;
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const int j = base + (int)threadIdx.x;
    bool alive = j < n;
    float m0 = 1.0f, m1 = 1.0f, m2 = 1.0f, q0 = 1.0f, q1 = 1.0f, q2 = 1.0f;
    float L0 = 0.0f, L1 = 0.0f, L2 = 0.0f, e0 = 0.0f, e1 = 0.0f, e2 = 0.0f, d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
    float f0 = 0.0f, f1 = 0.0f, f2 = 0.0f;
    int path = 0;
    bool ext_valid = false, sh_valid = false, end_after = false;
    f3 sh_org = zero3, sh_dir = zero3, rpos = zero3, rdir = zero3;
    int sh_light = -1;
    float sh_cos = 0.0f, sh_pt = 1.0f;
    if (alive) {
      const float4 S0 = sin_[0 * cap + j], S1 = sin_[1 * cap + j], S2 = sin_[2 * cap + j], S3 = sin_[3 * cap + j], S4 = sin_[4 * cap + j], S5 = sin_[5 * cap + j];
      m0 = S0.x; m1 = S0.y; m2 = S0.z; path = __float_as_int(S0.w);
      q0 = S1.x; q1 = S1.y; q2 = S1.z;
      const int fl = __float_as_int(S1.w);
      L0 = S2.x; L1 = S2.y; L2 = S2.z; sh_cos = S2.w;
      e0 = S3.x; e1 = S3.y; e2 = S3.z; sh_pt = S3.w;
      d0 = S4.x; d1 = S4.y; d2 = S4.z; sh_light = __float_as_int(S4.w);
      f0 = S5.x; f1 = S5.y; f2 = S5.z;
      const bool was_ext = (fl & WF_EXT_VALID) != 0, was_sh = (fl & WF_SH_VALID) != 0, was_end = (fl & WF_END_AFTER) != 0;
      // ---- [D] of the previous bounce (path_trace.cu:73-88, 176-180) ----
      if (was_sh) {
        d0 = d1 = d2 = 0.0f;
        const int hs_tri = hit_tri[cap + j];
        const float hs_t = hit_t[cap + j];
        f3 so, sd;
        wf_load_ray(rin, cap + j, &so, &sd);
        if (hs_tri >= 0) {
          const float4 lmb = __ldg(&mat[1 * nT + sh_light]);
          const f3 shit = mk3(fmaf(sd.x, hs_t, so.x), fmaf(sd.y, hs_t, so.y), fmaf(sd.z, hs_t, so.z));
          const f3 lfn = xyz(__ldg(&mat[3 * nT + sh_light]));
          const f3 ln = PLAIN ? lfn : shading_normal(sc, sh_light, __float_as_int(lmb.w), lfn, shit);
          const float cos_tp = -dot3(ln, sd);
          if (!(cos_tp < 0.0f) && hs_tri == sh_light) {
            const float g = (sh_cos * cos_tp) / ((hs_t * hs_t) * sh_pt);
            d0 = f0 * (lmb.x * g); d1 = f1 * (lmb.y * g); d2 = f2 * (lmb.z * g);
            if (RECORD) {   // the connection of the vertex recorded one bounce ago is visible
              float* va = reinterpret_cast<float*>(rec.V + ((size_t)(bounce - 1) * cap + path) * rec.vf4);
              *reinterpret_cast<float2*>(va + 2) = make_float2(__int_as_float(sh_light), g);
            }
          }
        }
        L0 = fmaf(q0, e0 + d0, L0); L1 = fmaf(q1, e1 + d1, L1); L2 = fmaf(q2, e2 + d2, L2);
        q0 = m0; q1 = m1; q2 = m2;
        if (was_end) { alive = false; if (RECORD) rec.vlen[path] = bounce; }
      }
      // ---- [A] this bounce's hit (path_trace.cu:111-144) ----
      if (alive && was_ext) {
        const int tri = hit_tri[j];
        const float ht = hit_t[j];
        wf_load_ray(rin, (size_t)j, &rpos, &rdir);
        if (tri < 0) {
          L0 = fmaf(q0, e0 + d0, L0); L1 = fmaf(q1, e1 + d1, L1); L2 = fmaf(q2, e2 + d2, L2);
          alive = false;
          if (RECORD) {
            float4* vp = rec.V + ((size_t)bounce * cap + path) * rec.vf4;
            vp[0] = make_float4(__int_as_float(-1), 0.0f, __int_as_float(-1), 0.0f);
            vp[1] = make_float4(q0, q1, q2, 0.0f);
            rec.vlen[path] = bounce + 1;
          }
        } else {
          int pixel, cur_s;
          float pix_c, pix_r;
          wf_path_ids(rp, b, path, &pixel, &cur_s, &pix_c, &pix_r);
          const float4 ma = __ldg(&mat[tri]), mb = __ldg(&mat[1 * nT + tri]);
          const int flags = __float_as_int(mb.w);
          const f3 face_n = xyz(__ldg(&mat[3 * nT + tri]));
          const f3 hit = mk3(fmaf(rdir.x, ht, rpos.x), fmaf(rdir.y, ht, rpos.y), fmaf(rdir.z, ht, rpos.z));
          const f3 ns = PLAIN ? face_n : shading_normal(sc, tri, flags, face_n, hit);
          if (bounce == 0) { e0 = mb.x; e1 = mb.y; e2 = mb.z; }
          float vb = 0.0f, vt0 = 0.0f, vt1 = 0.0f, vt2 = 0.0f;   // adjoint record: d(throughput factor)/dKd, per-channel throughput factor
          const u4 rn = philox4x32_10((uint32_t)pixel, (uint32_t)cur_s, (uint32_t)bounce, IPT_STREAM_BOUNCE, rp.keys);
          if (nE > 0) {
            const float u_pick = u16(rn.x >> 16), r1 = u16(rn.y >> 16), r2 = u16(rn.y & 0xffffu);
            float up = u_pick;
            if (nE > IPT_LIGHT_PICK16_MAX) up = u01(philox4x32_10((uint32_t)pixel, (uint32_t)cur_s, (uint32_t)bounce, IPT_STREAM_PICK, rp.keys).x);
            const int li = pick_light(slight, nE, up);
            const float4 la = __ldg(&slight[3 * li]), lb = __ldg(&slight[3 * li + 1]), lc = __ldg(&slight[3 * li + 2]);
            const float sq = sqrtf(r1);
            const float c1 = 1.0f - sq, c2 = sq * (1.0f - r2), c3 = r2 * sq;
            const f3 ep = mk3(fmaf(c3, lc.x, fmaf(c2, lb.x, c1 * la.x)), fmaf(c3, lc.y, fmaf(c2, lb.y, c1 * la.y)), fmaf(c3, lc.z, fmaf(c2, lb.z, c1 * la.z)));
            const f3 toLight = normalize3(ep - hit);
            const float cos_theta = dot3(ns, toLight);
            if (!(cos_theta < 0.0f)) {
              sh_valid = true;
              sh_org = hit; sh_dir = toLight;
              sh_light = __float_as_int(lc.w);
              sh_cos = cos_theta; sh_pt = lb.w;
              f0 = ma.x; f1 = ma.y; f2 = ma.z;
              if (!PLAIN && (flags & TRI_ANY_KS)) {
                const float4 mc = __ldg(&mat[4 * nT + tri]);
                const float co = phong_coeff(ns, rdir, toLight, ma.w);
                f0 = fmaf(mc.x, co, f0); f1 = fmaf(mc.y, co, f1); f2 = fmaf(mc.z, co, f2);
              }
            }
          }
          const bool depth_ok = (rp.max_depth <= 0) || (bounce + 1 < rp.max_depth);
          if (u16(rn.x & 0xffffu) < rp.p_rr && depth_ok) {
            f3 nd;
            float p_sample = IPT_INV_PI_F;
            if (!PLAIN && (flags & TRI_SPECULAR)) {
              nd = sample_specular_dir(face_n, ma.w, u01(rn.z), u01(rn.w), &p_sample);
            } else {
              float sphi, cphi;
              sincos2pi(u01(rn.z), &sphi, &cphi);
              const float ut = u01(rn.w);
              const float cth = sqrtf(ut), sth = sqrtf(1.0f - ut);
              const f3 hemi = mk3(sth * cphi, sth * sphi, cth);
              if (flags & TRI_NEG_Z) {
                nd = -hemi;
              } else {
                const float4 r0 = __ldg(&mat[5 * nT + tri]), r1 = __ldg(&mat[6 * nT + tri]), r2 = __ldg(&mat[7 * nT + tri]);
                nd = mk3(dot3(r0.x, r0.y, r0.z, hemi), dot3(r1.x, r1.y, r1.z, hemi), dot3(r2.x, r2.y, r2.z, hemi));
              }
              nd = normalize3(nd);
            }
            const float cosn = dot3(nd, ns);
            const float4 mp = __ldg(&mat[2 * nT + tri]);
            float i0 = mp.x, i1 = mp.y, i2 = mp.z;
            float coeff = cosn * rp.c_diff;
            if (!PLAIN && (flags & TRI_ANY_KS)) {
              const float4 mc = __ldg(&mat[4 * nT + tri]);
              const float co = phong_coeff(ns, rdir, nd, ma.w);
              i0 = fmaf(mc.x, co, i0); i1 = fmaf(mc.y, co, i1); i2 = fmaf(mc.z, co, i2);
              if (flags & TRI_SPECULAR) coeff = (cosn / p_sample) / rp.p_rr;
            }
            m0 = (m0 * i0) * coeff; m1 = (m1 * i1) * coeff; m2 = (m2 * i2) * coeff;
            rpos = hit;
            rdir = nd;
            ext_valid = true;
            vb = coeff / IPT_PI_F;
            vt0 = i0 * coeff; vt1 = i1 * coeff; vt2 = i2 * coeff;
          } else {
            end_after = true;
          }
          if (RECORD) {
            float4* vp = rec.V + ((size_t)bounce * cap + path) * rec.vf4;
            vp[0] = make_float4(__int_as_float(tri), vb, __int_as_float(-1), 0.0f);
            vp[1] = make_float4(q0, q1, q2, 0.0f);
            if (rec.vf4 == 4) { vp[2] = make_float4(vt0, vt1, vt2, 0.0f); vp[3] = make_float4(f0, f1, f2, 0.0f); }
            if (!sh_valid && end_after) rec.vlen[path] = bounce + 1;
          }
          if (!sh_valid) {
            d0 = d1 = d2 = 0.0f;
            L0 = fmaf(q0, e0 + d0, L0); L1 = fmaf(q1, e1 + d1, L1); L2 = fmaf(q2, e2 + d2, L2);
            q0 = m0; q1 = m1; q2 = m2;
            if (end_after) alive = false;
          }
        }
      }
      if (!alive) {   // path finished: its radiance goes to the per-sample buffer (summed in sample order by k_wf_accumulate)
        lsamp[0 * cap + path] = L0; lsamp[1 * cap + path] = L1; lsamp[2 * cap + path] = L2;
      }
    }
    // ---- append survivors ----
    const unsigned amask = __ballot_sync(kFull, alive);
    // one atomic per warp (one per block, through shared memory and three barriers per trip, measured no faster)
    int basej = 0;
    if (amask) {
      const int leader = __ffs(amask) - 1;
      if ((int)lane == leader) basej = atomicAdd(n_out, __popc(amask));
      basej = __shfl_sync(kFull, basej, leader);
    }
    if (amask) {
      if (alive) {
        const size_t jn = (size_t)basej + __popc(amask & ((1u << lane) - 1u));
        const int fl = (ext_valid ? WF_EXT_VALID : 0) | (sh_valid ? WF_SH_VALID : 0) | (end_after ? WF_END_AFTER : 0);
        sout[0 * cap + jn] = make_float4(m0, m1, m2, __int_as_float(path));
        sout[1 * cap + jn] = make_float4(q0, q1, q2, __int_as_float(fl));
        sout[2 * cap + jn] = make_float4(L0, L1, L2, sh_cos);
        sout[3 * cap + jn] = make_float4(e0, e1, e2, sh_pt);
        sout[4 * cap + jn] = make_float4(d0, d1, d2, __int_as_float(sh_light));
        sout[5 * cap + jn] = make_float4(f0, f1, f2, 0.0f);
        wf_store_ray(rout, jn, rpos, ext_valid ? rdir : zero3);
        wf_store_ray(rout, cap + jn, sh_org, sh_valid ? sh_dir : zero3);
        n_ext += ext_valid ? 1u : 0u;
        n_shadow += sh_valid ? 1u : 0u;
        if (ss.nGroups > 0) {
          HitRec he, hs;
          closest_hit_small2(ss, rpos, ext_valid ? rdir : zero3, sh_org, sh_valid ? sh_dir : zero3, he, hs);
          hit_tri_out[jn] = he.tri; hit_t_out[jn] = he.t;
          hit_tri_out[cap + jn] = hs.tri; hit_t_out[cap + jn] = hs.t;
        }
      }
    }
  }
  if (counters) {
    for (int o = 16; o > 0; o >>= 1) {
      n_ext += __shfl_down_sync(kFull, n_ext, o);
      n_shadow += __shfl_down_sync(kFull, n_shadow, o);
    }
    if (lane == 0 && (n_ext | n_shadow)) {
      atomicAdd(counters + 1, (unsigned long long)n_ext);
      atomicAdd(counters + 2, (unsigned long long)n_shadow);
    }
  }
}

// toneMap's running sum (path_trace.cu:190-193) over the batch: samples in order, each into the partial sum of its
// sample group (a batch may start inside a group -- it then continues that group's partial -- and span several groups)
__global__ void __launch_bounds__(kWfBlock) k_wf_accumulate(const float* __restrict__ lsamp, WfBatch b, int group_samples, int groups, int spp, float inv_spp,
                                                            float* __restrict__ partial) {
  const int pl = blockIdx.x * blockDim.x + threadIdx.x;
  if (pl >= b.tile_n) return;
  const size_t cap = (size_t)b.cap;
  int ss = 0;
  while (ss < b.ns) {
    const int s = b.s0 + ss;
    const int g = s / group_samples;
    const int g_begin = g * group_samples, g_end = min(spp, g_begin + group_samples);
    float* dst = partial + ((size_t)(b.tile_p0 + pl) * groups + g) * 3;
    float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
    if (s != g_begin) { p0 = dst[0]; p1 = dst[1]; p2 = dst[2]; }
    const int n_here = min(b.ns - ss, g_end - s);
    for (int k = 0; k < n_here; k++, ss++) {
      const size_t path = (size_t)ss * b.tile_n + pl;
      p0 += lsamp[0 * cap + path] * inv_spp; p1 += lsamp[1 * cap + path] * inv_spp; p2 += lsamp[2 * cap + path] * inv_spp;
    }
    dst[0] = p0; dst[1] = p1; dst[2] = p2;
  }
}

// gradient contribution of one vertex.  In a room-plus-mesh scene most vertices lie on the few room triangles -- the scene's
// brute-force set -- and per-lane double atomics on those ~50 addresses serialise in the L2 (the config-5 adjoint took 134 ms, 100 of
// them here; 78 ms with warp aggregation in front of the global atomics).  So: lanes that hit the same triangle of the set add up by
// shuffles, the lowest of them adds the sum to the WARP's private accumulator in shared memory (one address per triangle and
// channel, written by one lane per step: no atomics), and the block flushes once at the end.  Mesh triangles (millions of
// addresses, no contention) go straight to the global array.
constexpr int kSweepBigMax = 128;   // kBruteForceMaxTris
__device__ __forceinline__ void wf_grad_add(double* __restrict__ grad, const unsigned char* __restrict__ big_slot, float* __restrict__ wacc, int tri,
                                            float g0, float g1, float g2) {
  const unsigned lane = threadIdx.x & 31u;
  const int slot = big_slot ? (int)__ldg(&big_slot[tri]) : 255;
  if (slot == 255) {
    double* gg = grad + 3 * (size_t)tri;
    atomicAdd(gg + 0, (double)g0); atomicAdd(gg + 1, (double)g1); atomicAdd(gg + 2, (double)g2);
    return;
  }
  const unsigned peers = __match_any_sync(__activemask(), slot);
  const int leader = __ffs(peers) - 1;
  for (unsigned rest = peers & (peers - 1u); rest; rest &= rest - 1u) {
    const int src = __ffs(rest) - 1;
    const float b0 = __shfl_sync(peers, g0, src), b1 = __shfl_sync(peers, g1, src), b2 = __shfl_sync(peers, g2, src);
    if ((int)lane == leader) { g0 += b0; g1 += b1; g2 += b2; }
  }
  if ((int)lane == leader) { wacc[3 * slot] += g0; wacc[3 * slot + 1] += g1; wacc[3 * slot + 2] += g2; }
  __syncwarp(peers);
}

// backward sweep of the adjoint over the recorded vertices of a batch (the deferred sweep of k_paths<GRAD>, one thread
// per path; for a fixed bounce the threads of a warp read consecutive vertices):
//   Pbar_k = (E + D_k) + T_k Pbar_{k+1},   dL/dKd[t_k] += P_k (Ke_l g_k + b_k Pbar_{k+1}),   miss: the stale D_{k-1} is re-added with P_k
// Gradient contributions are weighted by dLoss/dL of the path's pixel (dldpx, shard-local pixel index) and added to
// grad[3 nT] with double atomics (BVH scenes have too many triangles for per-block tiles).
__global__ void __launch_bounds__(kWfBlock) k_wf_sweep(DevScene sc, WfBatch b, WfRecord rec, const float* __restrict__ dldpx, double* __restrict__ grad) {
  __shared__ float s_acc[(kWfBlock / 32) * 3 * kSweepBigMax];   // one private accumulator per warp for the brute-force set
  for (int i = threadIdx.x; i < (kWfBlock / 32) * 3 * kSweepBigMax; i += blockDim.x) s_acc[i] = 0.0f;
  __syncthreads();
  float* wacc = s_acc + (threadIdx.x >> 5) * 3 * kSweepBigMax;
  const unsigned char* __restrict__ big_slot = sc.big_slot;
  const int n = b.tile_n * b.ns;
  const size_t cap = (size_t)b.cap;
  const int nT = sc.nT;
  const float4* __restrict__ mat = sc.mat;
  for (int path = blockIdx.x * blockDim.x + threadIdx.x; path < n; path += gridDim.x * blockDim.x) {
    const int len = rec.vlen[path];
    if (len <= 0) continue;
    const int ss = path / b.tile_n;
    const int pl = b.tile_p0 + (path - ss * b.tile_n);
    const float w0 = dldpx[3 * (size_t)pl], w1 = dldpx[3 * (size_t)pl + 1], w2 = dldpx[3 * (size_t)pl + 2];
    const float4 first = rec.V[(size_t)path * rec.vf4];
    float se0 = 0.0f, se1 = 0.0f, se2 = 0.0f;   // L_e: emission of the first hit (sticky, path_trace.cu:168)
    if (__float_as_int(first.x) >= 0) { const float4 ke = __ldg(&mat[1 * nT + __float_as_int(first.x)]); se0 = ke.x; se1 = ke.y; se2 = ke.z; }
    float pb0 = 0.0f, pb1 = 0.0f, pb2 = 0.0f;
    for (int k = len - 1; k >= 0; k--) {
      const float4* vp = rec.V + ((size_t)k * cap + path) * rec.vf4;
      const float4 va = vp[0], vP = vp[1];
      const int vtri = __float_as_int(va.x);
      if (vtri >= 0) {
        const float4 kd = __ldg(&mat[vtri]);
        const int vlight = __float_as_int(va.z);
        const float vb = va.y, vg = va.w;
        float k0 = 0.0f, k1 = 0.0f, k2 = 0.0f;
        if (vlight >= 0) { const float4 lke = __ldg(&mat[1 * nT + vlight]); k0 = lke.x * vg; k1 = lke.y * vg; k2 = lke.z * vg; }
        float g0 = w0 * (vP.x * fmaf(vb, pb0, k0)), g1 = w1 * (vP.y * fmaf(vb, pb1, k1)), g2 = w2 * (vP.z * fmaf(vb, pb2, k2));
        wf_grad_add(grad, big_slot, wacc, vtri, g0, g1, g2);
        if (rec.vf4 == 2) {
          pb0 = fmaf(kd.x * vb, pb0, se0 + kd.x * k0); pb1 = fmaf(kd.y * vb, pb1, se1 + kd.y * k1); pb2 = fmaf(kd.z * vb, pb2, se2 + kd.z * k2);
        } else {
          const float4 vt = vp[2], vd = vp[3];
          pb0 = fmaf(vt.x, pb0, se0 + vd.x * k0); pb1 = fmaf(vt.y, pb1, se1 + vd.y * k1); pb2 = fmaf(vt.z, pb2, se2 + vd.z * k2);
        }
      } else {
        float c0 = se0, c1 = se1, c2 = se2;
        if (k > 0) {
          const float4* up = rec.V + ((size_t)(k - 1) * cap + path) * rec.vf4;
          const float4 ua = up[0];
          const int utri = __float_as_int(ua.x), ulight = __float_as_int(ua.z);
          if (utri >= 0 && ulight >= 0) {
            const float4 lke = __ldg(&mat[1 * nT + ulight]);
            const float ug = ua.w;
            const float k0 = lke.x * ug, k1 = lke.y * ug, k2 = lke.z * ug;
            wf_grad_add(grad, big_slot, wacc, utri, w0 * (vP.x * k0), w1 * (vP.y * k1), w2 * (vP.z * k2));
            float f0, f1, f2;
            if (rec.vf4 == 2) { const float4 kd = __ldg(&mat[utri]); f0 = kd.x; f1 = kd.y; f2 = kd.z; }
            else { const float4 ud = up[3]; f0 = ud.x; f1 = ud.y; f2 = ud.z; }
            c0 += f0 * k0; c1 += f1 * k1; c2 += f2 * k2;
          }
        }
        pb0 = c0; pb1 = c1; pb2 = c2;
      }
    }
  }
  // flush the warps' private sums of the brute-force set: one double atomic per triangle and channel per block
  __syncthreads();
  const int nb = min(sc.nBig, kSweepBigMax);
  for (int i = threadIdx.x; i < 3 * nb; i += blockDim.x) {
    float a = 0.0f;
    for (int w = 0; w < kWfBlock / 32; w++) a += s_acc[w * 3 * kSweepBigMax + i];
    if (a != 0.0f) atomicAdd(grad + 3 * (size_t)sc.big_tri[i / 3] + (i % 3), (double)a);
  }
}

// ---- wavefront forward pass ------------------------------------------------------------------------------
// workspace layout (bytes, all 256-aligned): state A, state B (6 float4 x cap each), rays A, rays B
// (2 x 6 floats x cap each), hit ids A / B, hit t A / B (2 x cap each), per-sample radiance (3 x cap), control words
static size_t wf_align(size_t b) { return (b + 255) & ~(size_t)255; }
size_t wavefront_workspace_bytes(int cap) {
  const size_t c = (size_t)cap;
  return 2 * wf_align(6 * c * sizeof(float4)) + 2 * wf_align(12 * c * sizeof(float)) + 2 * wf_align(2 * c * sizeof(int)) + 2 * wf_align(2 * c * sizeof(float)) +
         wf_align(3 * c * sizeof(float)) + 256;
}
// vertex record of the adjoint replay: vdepth bounces x cap paths x vf4 float4, plus the path lengths
size_t wavefront_record_bytes(int cap, int vdepth, int vf4) { return wf_align((size_t)vdepth * cap * vf4 * sizeof(float4)) + wf_align((size_t)cap * sizeof(int)); }

cudaError_t launch_wavefront(const DevScene& sc, const DevRender& rp, void* workspace, int cap, float* partial, unsigned long long* counters,
                             int sm_count, int* host_live, unsigned long long* n_launches, cudaStream_t st, const WavefrontAdjoint* adj) {
  const size_t c = (size_t)cap;
  char* w = (char*)workspace;
  float4* state[2];
  float* rays[2];
  state[0] = (float4*)w; w += wf_align(6 * c * sizeof(float4));
  state[1] = (float4*)w; w += wf_align(6 * c * sizeof(float4));
  rays[0] = (float*)w; w += wf_align(12 * c * sizeof(float));
  rays[1] = (float*)w; w += wf_align(12 * c * sizeof(float));
  int* hit_tri[2];
  float* hit_t[2];
  hit_tri[0] = (int*)w; w += wf_align(2 * c * sizeof(int));
  hit_tri[1] = (int*)w; w += wf_align(2 * c * sizeof(int));
  hit_t[0] = (float*)w; w += wf_align(2 * c * sizeof(float));
  hit_t[1] = (float*)w; w += wf_align(2 * c * sizeof(float));
  float* lsamp = (float*)w; w += wf_align(3 * c * sizeof(float));
  unsigned long long* trace_counter = (unsigned long long*)w;   // [0]
  int* n_live = (int*)(w + 16);                                 // [0], [1]: ping-pong live counts

  const int shade_grid = sm_count * 8;
  cudaError_t e = cudaSuccess;
  WfRecord rec;
  rec.V = nullptr; rec.vlen = nullptr; rec.vf4 = sc.plain ? 2 : 4;
  if (adj) {
    rec.V = (float4*)adj->record;
    rec.vlen = (int*)((char*)adj->record + wf_align((size_t)adj->vdepth * c * rec.vf4 * sizeof(float4)));
  }

  // the four variants of the shading kernel: scenes without a Phong lobe / vertex normals (PLAIN), with the adjoint's vertex record
  auto shade = sc.plain ? (adj ? k_wf_shade<true, true> : k_wf_shade<true, false>) : (adj ? k_wf_shade<false, true> : k_wf_shade<false, false>);
  const size_t shade_smem = small_scene_smem_bytes(sc.nStream);
  e = cudaFuncSetAttribute((const void*)shade, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shade_smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_wf_generate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shade_smem);
  if (e != cudaSuccess) return e;
  const int tile_cap = std::min(rp.n_pix, cap);
  const bool unbounded = rp.max_depth <= 0;
  for (int p0 = 0; p0 < rp.n_pix; p0 += tile_cap) {
    const int tn = std::min(tile_cap, rp.n_pix - p0);
    const int nsb = std::max(1, std::min(rp.spp, cap / tn));
    {
      for (int s0 = 0; s0 < rp.spp; s0 += nsb) {
        WfBatch b;
        b.cap = cap; b.tile_p0 = p0; b.tile_n = tn; b.s0 = s0; b.ns = std::min(nsb, rp.spp - s0);
        const int n_paths = b.tile_n * b.ns;
        int cur = 0;
        {
          NvtxRange r("ipt.wavefront.generate");
          k_wf_generate<<<std::min((n_paths + kWfBlock - 1) / kWfBlock, shade_grid), kWfBlock, shade_smem, st>>>(sc, rp, b, state[0], rays[0], hit_tri[0], hit_t[0],
                                                                                                                 n_live + 0, counters);
        }
        *n_launches += 1;
        for (int bounce = 0;; bounce++) {
          if (!unbounded && bounce > rp.max_depth) break;
          e = cudaMemsetAsync(n_live + (cur ^ 1), 0, sizeof(int), st);
          if (e != cudaSuccess) return e;
          {
            NvtxRange r("ipt.wavefront.extend");
            // the kernel that wrote the rays (k_wf_generate / k_wf_shade) has left their hits among the brute-force set
            e = launch_trace_bounce(sc, rays[cur], n_live + cur, cap, hit_tri[cur], hit_t[cur], trace_counter, sm_count, n_launches, st, true);
          }
          if (e != cudaSuccess) return e;
          {
            NvtxRange r("ipt.wavefront.shade");
            shade<<<shade_grid, kWfBlock, shade_smem, st>>>(sc, rp, b, bounce, state[cur], state[cur ^ 1], rays[cur], rays[cur ^ 1], hit_tri[cur], hit_t[cur],
                                                            hit_tri[cur ^ 1], hit_t[cur ^ 1], n_live + cur, n_live + (cur ^ 1), lsamp, counters, rec);
          }
          *n_launches += 1;
          cur ^= 1;
          if (unbounded) {
            e = cudaMemcpyAsync(host_live, n_live + cur, sizeof(int), cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
            if (*host_live == 0) break;
          }
        }
        if (adj) {
          NvtxRange r("ipt.wavefront.sweep");
          k_wf_sweep<<<std::min((n_paths + kWfBlock - 1) / kWfBlock, shade_grid), kWfBlock, 0, st>>>(sc, b, rec, adj->dldpx, adj->grad);
        } else {
          NvtxRange r("ipt.wavefront.accumulate");
          k_wf_accumulate<<<(b.tile_n + kWfBlock - 1) / kWfBlock, kWfBlock, 0, st>>>(lsamp, b, rp.group_samples, rp.groups, rp.spp, rp.inv_spp, partial);
        }
        *n_launches += 1;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
      }
    }
  }
  return cudaGetLastError();
}

}  // namespace ipt
