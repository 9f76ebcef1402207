// trace.cu -- closest hit for ray batches (Scene::getIntersection, scene.h:127 -> bvh.h:37-107 ->
// scene_basics.h:426-459): iptTraceRays* and the per-bounce traversal of the wavefront integrator.
//
//   k_trace_rays       brute force over the scene's shared-memory set (whole small scene / large triangles of a BVH scene)
//   k_trace_rays_bvh   persistent, phased BVH traversal
//   k_copy16           copy-bandwidth measurement aid (iptMeasureCopyGBps)
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "intersect.cuh"
#include "kernels.h"

namespace ipt {

// Persistent ray-batch traversal.  Each lane owns one ray at a time and is a state machine (inner node /
// leaf / finished); the warp advances in three phases so that lanes in the same state execute together:
//   fetch  -- every finished lane stores its hit and claims the next ray (one atomic per warp),
//   inner  -- lanes at inner nodes keep descending while at least kInnerQuorum of them are (a lane that
//             reaches a leaf, or finishes, waits),
//   leaf   -- every lane holding a leaf tests its triangles, then pops.
// (One state machine step per lane per trip -- the previous form -- executes the union of all three bodies on
// every trip with ~9 of 32 lanes active in each.)
// Ray index space: `n_live == nullptr`: rays [0, n).  Otherwise (wavefront): n = *n_live and the batch is
// [0, n) u [cap, cap + n).  A ray with direction 0 is "no ray": a miss.
#ifndef IPT_INNER_QUORUM
#define IPT_INNER_QUORUM 14
#endif
constexpr int kInnerQuorum = IPT_INNER_QUORUM;
#ifndef IPT_FETCH_QUORUM
#define IPT_FETCH_QUORUM 8
#endif
constexpr int kFetchQuorum = IPT_FETCH_QUORUM;
#ifndef IPT_TRACE_MINB
#define IPT_TRACE_MINB 10   // resident blocks per SM the register allocation aims at (10: 48 registers, no spills; 12: 40 registers, 70 bytes of spills, 2 % slower; 16: 32 registers, 36 % slower)
#endif
#ifndef IPT_SPECULATE
#define IPT_SPECULATE 0   // 1: a lane that reaches a leaf postpones it once and keeps descending (Aila & Laine's speculative traversal)
#endif
#ifndef IPT_RAY_CHUNK
#define IPT_RAY_CHUNK 128
#endif
constexpr int kRayChunk = IPT_RAY_CHUNK;

template <bool STATS>
__global__ void __launch_bounds__(kBlock, IPT_TRACE_MINB) k_trace_rays_bvh(DevScene sc, const float* __restrict__ rays, int n_direct, const int* __restrict__ n_live, int cap,
                                                           int* __restrict__ tri, float* __restrict__ tout, unsigned long long* __restrict__ counter,
                                                           unsigned long long* __restrict__ stats) {
  extern __shared__ float4 smem[];
  const uint32_t col = smem_addr(reinterpret_cast<int*>(smem) + threadIdx.x), stride = blockDim.x * (uint32_t)sizeof(int);
  const unsigned lane = threadIdx.x & 31u;
  // a scene with a brute-force set (large triangles kept out of the tree): each ray's hit among those is already in tri / tout
  // -- written by k_trace_rays (iptTraceRays*) or by the wavefront kernel that produced the ray (k_wf_generate / k_wf_shade) --
  // and the traversal starts from that bound
  const bool seeded = sc.nGroups > 0;
  // a batch has < 2^31 rays (a wavefront bounce: 2 x cap <= 2^25), so ray indices are ints
  const int n = n_live ? *n_live : n_direct;
  const int total = n_live ? 2 * n : n;
  unsigned nodes = 0, tris = 0;
  BvhRay r;
  r.cur = kBvhDone;
  r.best_t = INFINITY; r.best_i = -1; r.sp = 0;
  int idx = -1;
  bool exhausted = false;       // the batch has no ray left for this lane
  int pend = 0;                 // IPT_SPECULATE: a postponed leaf (descriptor < 0), 0 = none
  int chunk_next = 0, chunk_end = 0;   // the warp's current chunk of ray indices (warp-uniform; a batch has < 2^31 rays)
  for (;;) {
    // ---- fetch ----
    const bool want = r.cur == kBvhDone && pend == 0 && !exhausted;
    const unsigned wm = __ballot_sync(kFull, want);
    // a fetch costs the warp ~60 instructions however few lanes take part: wait for kFetchQuorum finished lanes unless
    // the others have nothing to do either
    if (wm && (__popc(wm) >= kFetchQuorum || __ballot_sync(kFull, r.cur != kBvhDone || pend != 0) == 0u)) {
      if (want && idx >= 0) { tri[idx] = r.best_i; tout[idx] = r.best_t; idx = -1; }
      // rays are claimed from the global counter kRayChunk at a time per warp (one atomic per chunk, not per refill:
      // every warp of the grid hits the same address); a refill takes what is left of the warp's chunk first
      const int need = __popc(wm), rank = __popc(wm & ((1u << lane) - 1u));
      const int left = chunk_end - chunk_next;
      int mine;
      if (need > left) {
        const int leader = __ffs(wm) - 1;
        unsigned long long fresh = 0;
        if ((int)lane == leader) fresh = atomicAdd(counter, (unsigned long long)kRayChunk);
        fresh = __shfl_sync(kFull, fresh, leader);
        // a claim far beyond the batch (the counter keeps growing while warps drain) is clamped: it is exhausted anyway
        const int fb = (int)min(fresh, (unsigned long long)0x7fff0000u);
        mine = rank < left ? chunk_next + rank : fb + (rank - left);
        chunk_next = fb + (need - left);
        chunk_end = fb + kRayChunk;
      } else {
        mine = chunk_next + rank;
        chunk_next += need;
      }
      if (want) {
        if (mine >= total) {
          exhausted = true;
        } else {
          idx = mine < n ? mine : cap + (mine - n);
          const float2* r2 = reinterpret_cast<const float2*>(rays + 6 * (size_t)idx);
          const float2 a = __ldg(r2), b = __ldg(r2 + 1), c = __ldg(r2 + 2);
          // the seed is requested with the ray, not after the "no ray" test: one load latency per fetch instead of two
          int seed_i = -1;
          float seed_t = INFINITY;
          if (seeded) { seed_i = tri[idx]; seed_t = tout[idx]; }
          r.best_i = -1; r.best_t = INFINITY;
          if (!(b.y == 0.0f && c.x == 0.0f && c.y == 0.0f)) {
            const f3 rp = mk3(a.x, a.y, b.x), rd = mk3(b.y, c.x, c.y);
            bvh_begin(sc, r, rp, rd);
            r.best_i = seed_i; r.best_t = seed_t;
          }
        }
      }
    }
    if (__all_sync(kFull, r.cur == kBvhDone && pend == 0)) {
      if (__all_sync(kFull, exhausted)) break;
      continue;   // lanes whose new ray ended at once (no ray / root box missed) store and refetch at the top
    }
    // ---- inner nodes ----
    for (;;) {
      const bool inner = (unsigned)r.cur < (unsigned)kBvhDone;
      const unsigned im = __ballot_sync(kFull, inner);
      if (im == 0u) break;
      if (__popc(im) < kInnerQuorum) {
        // leave only if the waiting lanes will then do something: a leaf to test, or enough finished lanes for a fetch
        const unsigned leaves = __ballot_sync(kFull, r.cur < 0 || (r.cur == kBvhDone && pend != 0));
        const unsigned idle = __ballot_sync(kFull, r.cur == kBvhDone && pend == 0 && !exhausted);
        if (leaves != 0u || __popc(idle) >= kFetchQuorum) break;
      }
      if (inner) {
        bvh_step<STATS>(sc, r, col, stride, &nodes, &tris);
#if IPT_SPECULATE
        if (r.cur < 0 && pend == 0) {   // first leaf: postpone it, continue with the next stack entry
          pend = r.cur;
          r.sp -= 1;
          r.cur = kBvhDone;
          if (r.sp >= 0) r.cur = lds1i(col + (uint32_t)r.sp * stride);
          r.sp = max(r.sp, 0);
        }
#endif
      }
    }
    // ---- leaves ----
    if (r.cur < 0 || pend != 0) {
      // the direction is only needed here (the inner phase runs on 1/d): re-read it instead of keeping three registers
      // alive through the inner loop
      const float2* r2 = reinterpret_cast<const float2*>(rays + 6 * (size_t)idx);
      const float2 b = __ldg(r2 + 1), c = __ldg(r2 + 2);
      r.d = mk3(b.y, c.x, c.y);
#if IPT_BVH_QUANT
      const float2 a = __ldg(r2);   // quantised nodes: the inner phase runs on qa / qb, the origin is leaf-only as well
      r.p = mk3(a.x, a.y, b.x);
#endif
#if IPT_SPECULATE
      if (pend != 0) {
        const int packed = ~pend;
        if (STATS) tris += (unsigned)(packed & 7);
        test_leaf_lex(sc.geom, sc.geom_edges, sc.bvh_prims, packed >> 3, packed & 7, r.p, r.d, r.best_t, r.best_i);
        pend = 0;
      }
      if (r.cur < 0)
#endif
      bvh_step<STATS>(sc, r, col, stride, &nodes, &tris);
    }
  }
  if (idx >= 0) { tri[idx] = r.best_i; tout[idx] = r.best_t; }
  if (STATS) {
    atomicAdd(stats, (unsigned long long)nodes);
    atomicAdd(stats + 1, (unsigned long long)tris);
  }
}

// ------------------------------------------------------------------------------------------------
// brute-force closest hit of a ray batch against the scene's shared-memory set: the whole scene of a small scene, or
// the large-triangle set of a BVH scene (then k_trace_rays_bvh continues from these hits).  Index space as in
// k_trace_rays_bvh: [0, n_direct), or with n_live the two halves [0, n) u [cap, cap + n) of a wavefront bounce.
__global__ void __launch_bounds__(kBlock) k_trace_rays(DevScene sc, const float* __restrict__ rays, int n_direct, const int* __restrict__ n_live, int cap,
                                                       int* __restrict__ tri, float* __restrict__ tout) {
  extern __shared__ float4 smem[];
  float4* after;
  const SmallScene ss = stage_small_scene(sc, smem, &after);
  __syncthreads();
  const long long n = n_live ? (long long)*n_live : (long long)n_direct;
  const long long total = n_live ? 2 * n : n;
  // two rays per lane and trip (ray m and ray m + half) through the packed FP32x2 dual-ray loop of k_paths
  const long long half = (total + 1) / 2;
  const f3 zero3 = mk3(0.0f, 0.0f, 0.0f);
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < half; m += (long long)gridDim.x * blockDim.x) {
    const long long m2 = m + half;
    const bool two = m2 < total;
    const long long i1 = m < n ? m : (long long)cap + (m - n);
    const long long i2 = !two ? i1 : (m2 < n ? m2 : (long long)cap + (m2 - n));
    const float2* r1 = reinterpret_cast<const float2*>(rays + 6 * i1);
    const float2* r2 = reinterpret_cast<const float2*>(rays + 6 * i2);
    const float2 a1 = __ldg(r1), b1 = __ldg(r1 + 1), c1 = __ldg(r1 + 2);
    const float2 a2 = __ldg(r2), b2 = __ldg(r2 + 1), c2 = __ldg(r2 + 2);
    HitRec h1, h2;
    closest_hit_small2(ss, mk3(a1.x, a1.y, b1.x), mk3(b1.y, c1.x, c1.y), mk3(a2.x, a2.y, b2.x), two ? mk3(b2.y, c2.x, c2.y) : zero3, h1, h2);
    tri[i1] = h1.tri;
    tout[i1] = h1.t;
    if (two) { tri[i2] = h2.tri; tout[i2] = h2.t; }
  }
}

// dynamic shared memory of k_trace_rays_bvh: one stack column per thread
static size_t trace_bvh_smem_bytes(const DevScene& sc) { return bvh_stack_bytes(sc); }

// brute-force pass over the scene's shared-memory set (whole small scene / large-triangle set of a BVH scene)
static cudaError_t launch_trace_small(const DevScene& sc, const float* rays, int n_direct, const int* n_live, int cap, int* tri, float* t, int sm_count,
                                      cudaStream_t st) {
  const size_t smem = small_scene_smem_bytes(sc.nStream);
  cudaError_t e = cudaFuncSetAttribute(k_trace_rays, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long want = n_live ? (long long)sm_count * 16 : ((long long)n_direct + kBlock - 1) / kBlock;
  const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count * 16));
  k_trace_rays<<<grid, kBlock, smem, st>>>(sc, rays, n_direct, n_live, cap, tri, t);
  return cudaGetLastError();
}

// one bounce of the wavefront integrator: hits of the rays [0, n) u [cap, cap + n), n = *n_live
cudaError_t launch_trace_bounce(const DevScene& sc, const float* rays, const int* n_live, int cap, int* tri, float* t, unsigned long long* counter,
                                int sm_count, unsigned long long* n_launches, cudaStream_t st, bool seeded) {
  cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  if (sc.nGroups > 0 && !seeded) {   // `seeded`: tri / t already hold the hits among the brute-force set (written by k_wf_shade)
    e = launch_trace_small(sc, rays, 0, n_live, cap, tri, t, sm_count, st);
    if (e != cudaSuccess) return e;
    *n_launches += 1;
  }
  if (!sc.use_bvh) return cudaSuccess;   // brute-force scene driven through the wavefront (iptSceneSetIntegrator(2): A/B only)
  const size_t smem = trace_bvh_smem_bytes(sc);
  e = cudaFuncSetAttribute(k_trace_rays_bvh<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int nb = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_rays_bvh<false>, kBlock, smem);
  if (nb < 1) nb = 1;
  if (const char* ov = getenv("IPT_TRACE_BLOCKS")) nb = std::max(1, std::min(nb, atoi(ov)));   // occupancy experiments
  static bool said = false;
  if (!said && getenv("IPT_TRACE_DEBUG")) {
    said = true;
    fprintf(stderr, "ipt: k_trace_rays_bvh: %d blocks/SM, %zu B of stack per block (tree depth %d)\n", nb, smem, sc.bvh_depth);
  }
  k_trace_rays_bvh<false><<<sm_count * nb, kBlock, smem, st>>>(sc, rays, 0, n_live, cap, tri, t, counter, nullptr);
  *n_launches += 1;
  return cudaGetLastError();
}

// plain grid-stride copy, 16 bytes per thread per step (iptMeasureCopyGBps)
__global__ void __launch_bounds__(256) k_copy16(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
cudaError_t launch_copy16(const float4* src, float4* dst, size_t n, int sm_count, cudaStream_t st) {
  k_copy16<<<sm_count * 8, 256, 0, st>>>(src, dst, n);
  return cudaGetLastError();
}

cudaError_t launch_trace_rays(const DevScene& sc, const float* rays, int n, int* tri, float* t, unsigned long long* counter, int sm_count, cudaStream_t st) {
  if (sc.use_bvh) {
    if (sc.nGroups > 0) {
      cudaError_t e0 = launch_trace_small(sc, rays, n, nullptr, 0, tri, t, sm_count, st);
      if (e0 != cudaSuccess) return e0;
    }
    const size_t smem = trace_bvh_smem_bytes(sc);
    cudaError_t e = cudaFuncSetAttribute(k_trace_rays_bvh<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int nb = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_rays_bvh<false>, kBlock, smem);
    if (nb < 1) nb = 1;
    e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    k_trace_rays_bvh<false><<<sm_count * nb, kBlock, smem, st>>>(sc, rays, n, nullptr, 0, tri, t, counter, nullptr);
    return cudaGetLastError();
  }
  return launch_trace_small(sc, rays, n, nullptr, 0, tri, t, sm_count, st);
}

cudaError_t launch_trace_rays_stats(const DevScene& sc, const float* rays, int n, int* tri, float* t, unsigned long long* counter,
                                    unsigned long long* stats, int sm_count, cudaStream_t st) {
  const size_t smem = trace_bvh_smem_bytes(sc);
  cudaError_t e = cudaFuncSetAttribute(k_trace_rays_bvh<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int nb = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_rays_bvh<true>, kBlock, smem);
  if (nb < 1) nb = 1;
  e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  if (sc.nGroups > 0) {
    e = launch_trace_small(sc, rays, n, nullptr, 0, tri, t, sm_count, st);
    if (e != cudaSuccess) return e;
  }
  k_trace_rays_bvh<true><<<sm_count * nb, kBlock, smem, st>>>(sc, rays, n, nullptr, 0, tri, t, counter, stats);
  return cudaGetLastError();
}

}  // namespace ipt
