// paths.cu -- the persistent path kernel of the hot path and its epilogues.
//
//   k_paths<MODE, BVH, PLAIN>  renderSample + radiance + directLighting + BSDF + sampleNextDir (path_trace.cu:10-184)
//        MODE_RENDER : forward radiance
//        MODE_GRAPH  : Edge statistics            (inv_path_trace.cu:6-182, inv_scene.h:9-48)
//        MODE_GRAD   : dLoss/dKd by path adjoint  (new)
//   k_resolve           group partials -> pixel mean + tonemap (toneMap, path_trace.cu:186-198)
//   k_loss_dldpx / k_reduce_blocks / k_graph_finalize / k_set_kd    small epilogues
//
// Execution model: one persistent grid sized to the SM count.  Every lane owns one path
// at a time; one trip of the loop shades the previous hit, regenerates finished paths (next sample of the
// lane's work item, new items through a warp-aggregated atomic), and intersects two rays per lane in one
// pass over the scene: the shadow ray of this bounce and the extension ray of the next.  Scenes up to 128
// triangles live in shared memory as plane groups / quad slots and are intersected branch-free in packed
// FP32x2 (intersect.cuh); larger scenes use the BVH -- and, for the forward pass, the wavefront integrator of
// wavefront.cu instead of this kernel.  All arithmetic follows the canon of ipt_math.cuh (-fmad=false).
#include <algorithm>

#include "intersect.cuh"
#include "kernels.h"
#include "shading.cuh"

namespace ipt {

// One path vertex of the adjoint pass: 32 bytes = one memory sector, two 16-byte words
//   a = {tri (int bits), b, light (int bits), g}     b_ = {P0, P1, P2, unused}
// The stacks live in global memory laid out by the kernel (k_paths, MODE_GRAD) so that one vertex is one
// sector; as a thread-local array every 4-byte field of a vertex lands in a different 128-byte line.
// Scenes with a Phong lobe or per-vertex normals (template PLAIN == false) record two more words, 64 bytes in all:
//   t = {T0, T1, T2, -}  per-channel throughput factor of the bounce, (Kd/pi + Ks phong) * coeff  (0 when the path ended)
//   d = {f0, f1, f2, -}  BSDF(isDirect = true) of the bounce, Kd + Ks phong
// so that the backward sweep Pbar_k = C_k + T_k Pbar_{k+1} is exact for them too (plain scenes rebuild both from Kd).
struct AdjVertex {
  float4 a, p, t, d;
};
__device__ __forceinline__ int adj_tri(const AdjVertex& v) { return __float_as_int(v.a.x); }
__device__ __forceinline__ int adj_light(const AdjVertex& v) { return __float_as_int(v.a.z); }

// ------------------------------------------------------------------------------------------------
// the persistent path kernel
// ------------------------------------------------------------------------------------------------
// 3 resident blocks (24 warps / SM, 80 registers with ~330 B of L1-resident spills) measured 2 % faster than
// 2 blocks at 121 registers on the 1024x1024x256 Cornell render (A/B of round 1)
#ifndef IPT_MINB
#define IPT_MINB 5
#endif
// Camera rays are generated AHEAD of their use (profiles/r02_*: path regeneration ran every trip of the loop with ~9 of 32
// lanes, 120 instructions): every lane owns a ring of kRing finished camera rays of its upcoming samples in shared memory;
// when a lane that needs a ray finds its ring empty, ALL lanes with a free slot generate one (so generation runs once per
// ~2.4 trips at ~22 lanes) and taking a ray is one 16-byte load.  Entry = {direction, work item}; the generator's own
// cursor (pixel, row / column, next sample, end) lives in two more shared-memory slots of the lane.
#ifndef IPT_RING
#define IPT_RING 4       // camera rays a lane keeps ready (power of two)
#endif
#ifndef IPT_CLAIM_MAX
#define IPT_CLAIM_MAX 8  // longest run of work items a lane claims at once
#endif
constexpr int kRing = IPT_RING;
constexpr int kRingF4 = kRing + 3;   // + two generator slots + the lane's item slot {partial sum rgb, work item}
constexpr int kRingWords = kRingF4 * 4 + 1;   // 32-bit words per thread: the float4 slots + one word (last claim position)
constexpr uint32_t kRingStride = kBlock * 16u;   // bytes between two slots of one thread
// BVH is a compile-time flag: in brute-force mode every scene access is a shared-memory access with 32-bit
// addressing (a runtime choice between shared and global pointers forces generic loads)
// PLAIN is a compile-time promise that no triangle has per-vertex normals or a specular lobe (true for every
// asset the reference ships); the Phong / vertex-normal code is then not compiled into the kernel at all.
// A/B switches of round 2 (profiles/r02_ab.txt): 1 = the shipped choice
#ifndef IPT_EXPERIMENT_FAST_SINCOS
#define IPT_EXPERIMENT_FAST_SINCOS 0
#endif
#ifndef IPT_EXPERIMENT_NO_ND_NORMALIZE
#define IPT_EXPERIMENT_NO_ND_NORMALIZE 0
#endif
#ifndef IPT_FASTMATH
#define IPT_FASTMATH 1        // sqrt / normalise through the single-test sequences of ipt_math.cuh (same bits as sqrtf and /)
#endif
#if IPT_FASTMATH
#define IPT_NORMALIZE(v) normalize3_dev(v)
#define IPT_SQRT_UNIT(x) sqrt_rn_unit(x)
#define IPT_SQRT(x) sqrt_rn(x)
#else
#define IPT_NORMALIZE(v) normalize3(v)
#define IPT_SQRT_UNIT(x) sqrtf(x)
#define IPT_SQRT(x) sqrtf(x)
#endif
#ifndef IPT_SWEEP_MATCH
#define IPT_SWEEP_MATCH 0     // __match_any_sync pre-reduction of the adjoint's gradient atomics: measured, not faster (profiles/r02_adjoint_ab.txt)
#endif
#ifndef IPT_SWEEP_PER_TRIP
#define IPT_SWEEP_PER_TRIP 3
#endif
constexpr int kSweepPerTrip = IPT_SWEEP_PER_TRIP;   // adjoint sweep: at most this many vertices per lane per trip of the main loop ...
#ifndef IPT_SWEEP_QUORUM
#define IPT_SWEEP_QUORUM 1
#endif
constexpr int kSweepQuorum = IPT_SWEEP_QUORUM;      // ... while at least this many lanes of the warp have one pending
// resident blocks per SM: the forward and graph kernels of plain brute-force scenes -- every scene the reference ships --
// fit 80 registers without spills once the per-item state lives in shared memory, i.e. six blocks (24 warps) per SM
// instead of five; the adjoint and the BVH / Phong / vertex-normal variants would spill there and keep five
#ifndef IPT_MINB_FAST
#define IPT_MINB_FAST 6
#endif
template <int MODE, bool BVH, bool PLAIN>
__global__ void __launch_bounds__(kBlock, (MODE != MODE_GRAD && !BVH && PLAIN) ? IPT_MINB_FAST : IPT_MINB) k_paths(DevScene sc, DevRender rp, DevOut out) {
  extern __shared__ float4 smem[];
  const int nT = sc.nT, nE = sc.nE;
  SmallScene ss;
  ss.nGroups = 0;
  // per-thread camera-ray ring + generator state (see [B]): kRingF4 float4 columns, slot k of thread t at [k * kBlock + t]
  const uint32_t ring_a = smem_addr(smem) + threadIdx.x * 16u;
  float4* cursor = smem + kRingF4 * kBlock + kBlock / 4;   // float4 slots + one 32-bit word per thread
  int* bvh_stack = reinterpret_cast<int*>(cursor);
  if (BVH) cursor += (sc.bvh_depth * kBlock + 3) / 4;   // one stack column per thread, as deep as the tree
  // brute-force set: the whole scene, or in BVH mode the large triangles kept out of the tree
  if (!BVH || sc.nGroups > 0) ss = stage_small_scene(sc, cursor, &cursor);
  float4* smat = cursor;
  float4* slight_s = smat + (BVH ? 0 : kMatStride * nT);
  const bool lights_staged = !BVH || nE <= kLightSmemMax;   // an emissive mesh stays in global memory (BVH scenes only)
  float* sacc = reinterpret_cast<float*>(slight_s + (lights_staged ? 3 * nE : 0));
  const float4* __restrict__ slight = lights_staged ? slight_s : sc.light;
  const float4* __restrict__ mat = BVH ? sc.mat : smat;
  if (!BVH) {
    for (int i = threadIdx.x; i < kMatStride * nT; i += blockDim.x) smat[i] = sc.mat[i];
  }
  if (lights_staged) {
    for (int i = threadIdx.x; i < 3 * nE; i += blockDim.x) slight_s[i] = sc.light[i];
  }
  if (MODE != MODE_RENDER && out.acc_in_smem == 1) {
    for (int i = threadIdx.x; i < out.acc_stride; i += blockDim.x) sacc[i] = 0.0f;
  }
  __syncthreads();

  // material / light records: shared memory at explicit 32-bit addresses in brute-force mode (intersect.cuh, lds4), global in BVH mode
  const uint32_t smat_a = smem_addr(smat), slight_a = smem_addr(slight_s);
  const uint32_t nT16 = (uint32_t)nT << 4;
  auto M = [&](int field, int tri) -> float4 { return BVH ? mat[field * nT + tri] : lds4(smat_a + (uint32_t)field * nT16 + ((uint32_t)tri << 4)); };
  auto LREC = [&](int li, int k) -> float4 { return BVH ? slight[3 * li + k] : lds4(slight_a + (uint32_t)(3 * li + k) * 16u); };

  const unsigned lane = threadIdx.x & 31u;
  const unsigned n_items = (unsigned)rp.n_pix * (unsigned)rp.groups;   // < 2^31 (make_plan): 32-bit item arithmetic
  const f3 zero3 = mk3(0.0f, 0.0f, 0.0f);

  // ---- work item bookkeeping -------------------------------------------------------------------
  // the lane's work item (pixel, sample group) and the running sum of its finalised samples live in shared memory (slot
  // kRing + 2 = {part r, g, b, item}): they are touched once per path, not once per trip, and registers decide occupancy
  const uint32_t item_a = ring_a + (kRing + 2) * kRingStride;
  sts4(item_a, make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0xffffffffu)));
  int pixel = 0;                  // global pixel index of the item (RNG counter, target lookup)
  int fin_left = 0;               // samples of the item not yet finalised
  int cur_s = 0;                  // sample index of the path in flight (RNG counter)
  int rhead = 0, rcount = 0;      // camera-ray ring: first entry, entries
  bool g_done = false;            // the generator found the work queue empty
  sts4(ring_a + (kRing + 1) * kRingStride, make_float4(0.0f, 0.0f, 0.0f, 0.0f));   // generator cursor {next sample, end sample, next item, end of the claimed run}: all "used up"
  // queue position this lane saw at its last claim (sizes the next one): one more shared-memory word per thread, behind the float4 slots
  const uint32_t claim_a = smem_addr(smem) + kRingF4 * kRingStride + threadIdx.x * 4u;
  sts1i(claim_a, 0);
  // floor(i / groups) and floor(pixel / width) by multiplication (device_scene.h: exact for 31-bit dividends)
  auto decode_item = [&](unsigned my, int& pix_local, int& g, int& px, int& r, int& c) {
    pix_local = rp.groups == 1 ? (int)my : (int)(__umulhi(my, rp.magic_groups) >> rp.shift_groups);
    g = (int)(my - (unsigned)pix_local * (unsigned)rp.groups);
    const int lr = rp.width == 1 ? pix_local : (int)(__umulhi((unsigned)pix_local, rp.magic_width) >> rp.shift_width);
    c = pix_local - lr * rp.width;
    r = rp.row_begin + lr * rp.row_step;       // interleaved row shards
    px = r * rp.width + c;                     // global pixel index
  };
  // ---- path state ----------------------------------------------------------------------------------
  bool active = false;            // `h` holds the closest hit of this path's extension ray
  HitRec h;
  h.t = INFINITY; h.tri = -1;
  f3 rpos = zero3, rdir = zero3;  // extension ray (origin, direction)
  float m0 = 1.0f, m1 = 1.0f, m2 = 1.0f;     // multiplier
  float q0 = 1.0f, q1 = 1.0f, q2 = 1.0f;     // prevMultiplier
  float L0 = 0.0f, L1 = 0.0f, L2 = 0.0f;
  int e_tri = -1;                            // L_e (sticky, path_trace.cu:168) = Ke of the path's first hit, kept as that triangle's index
  float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;     // L_d (sticky on a miss)
  int bounce = 0;
  // graph mode
  float gweight = 1.0f;
  int gdst = 0;
  float pix0 = 0.0f, pix1 = 0.0f, pix2 = 0.0f;
  // grad mode
  // vertex stacks: slot s of this lane is the 32 (plain scenes) / 64 bytes at vw + s * kVS (float4 units); two halves of vdepth slots:
  // path being recorded / path being swept
  float4* vw = nullptr;
  constexpr int kVF4 = PLAIN ? 2 : 4;       // float4 words per recorded vertex
  constexpr int kVS = 32 * kVF4;            // float4 distance between two slots of one lane
  if (MODE == MODE_GRAD) vw = out.vstacks + ((size_t)(blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5)) * (2 * out.vdepth) * 32 + lane) * kVF4;
  int vbuf = 0;                              // first slot of the half the path in flight records into
  int sw_buf = 0, sw_k = -1;                 // pending sweep: its half, next vertex to process (-1: none)
  bool fin_wait = false;                     // a finished path waits for the pending sweep before it can be parked
  float pb0 = 0.0f, pb1 = 0.0f, pb2 = 0.0f;  // Pbar of the pending sweep
  float se0 = 0.0f, se1 = 0.0f, se2 = 0.0f;  // its path's L_e
  float sw0 = 0.0f, sw1 = 0.0f, sw2 = 0.0f;  // its pixel's dLoss/dL / spp
  float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f;     // dLoss/dL_pixel / spp of the item's pixel
  int path_len = 0;                          // vertices recorded for the path being finalised
  unsigned n_ext = 0, n_shadow = 0;   // per lane and launch: < 2^32 (a lane traces ~10^5 rays per second of launch); paths = n_pix * spp, counted by lane 0 of block 0
  // the light connection of the current bounce; written in [A] and read in [C] / [D] only when sh_valid, so they are
  // not reset per trip (the resets were ~30 register moves per trip at full warp width)
  f3 sh_org = zero3, sh_dir = zero3;
  int sh_light = -1, sh_tri = -1;
  float sh_cos = 0.0f, sh_pt = 1.0f, sh_gw = 0.0f;
  float f0 = 0.0f, f1 = 0.0f, f2 = 0.0f;   // BSDF(isDirect = true) of this bounce

  // 1/p_sample/p_RR/(1-P_SPEC) in double then rounded (inv_path_trace.cu:145)
  const float gscale = (float)(1.0 / (double)IPT_INV_PI_F / (double)rp.p_rr / 1.0);

  // L_e of the path in flight: the emission record of its first hit (zero before that and for non-emitters)
  auto load_le = [&](float& e0, float& e1, float& e2) {
    e0 = e1 = e2 = 0.0f;
    if (e_tri >= 0) { const float4 ke = M(1, e_tri); e0 = ke.x; e1 = ke.y; e2 = ke.z; }
  };

  // one vertex of a pending adjoint sweep: Pbar_k = C_k + Kd_k b_k Pbar_{k+1}.
  // The sweep of a finished path is not run where the path ends (there ~2 lanes of a warp would walk their
  // stacks while 30 wait): it is parked in the lane's other stack half and advanced kSweepPerTrip vertices per
  // trip of the main loop, where most lanes of the warp have one pending.
  auto sweep_step = [&]() {
    const float4* vs = vw + (size_t)sw_buf * kVS;
    const int k = sw_k--;
    AdjVertex v;
    v.a = vs[k * kVS]; v.p = vs[k * kVS + 1];
    if (!PLAIN) { v.t = vs[k * kVS + 2]; v.d = vs[k * kVS + 3]; }
    const int vtri = adj_tri(v);
    if (vtri >= 0) {
      const float4 kd = M(0, vtri);
      const int vlight = adj_light(v);
      const float vb = v.a.y, vg = v.a.w;
      float k0 = 0.0f, k1 = 0.0f, k2 = 0.0f;   // Ke_light * g
      if (vlight >= 0) { const float4 lke = M(1, vlight); k0 = lke.x * vg; k1 = lke.y * vg; k2 = lke.z * vg; }
      float g0 = v.p.x * fmaf(vb, pb0, k0), g1 = v.p.y * fmaf(vb, pb1, k1), g2 = v.p.z * fmaf(vb, pb2, k2);
#if IPT_SWEEP_MATCH
      // warp pre-reduction before the atomics (A/B, profiles/r02_adjoint_ab.txt): lanes of this sweep step that update the
      // same triangle sum their contributions by shuffles and the lowest of them issues the three reductions
      g0 *= sw0; g1 *= sw1; g2 *= sw2;
      const unsigned peers = __match_any_sync(__activemask(), vtri);
      const int leader = __ffs(peers) - 1;
      for (unsigned rest = peers & (peers - 1u); rest; rest &= rest - 1u) {
        const int src = __ffs(rest) - 1;
        const float b0 = __shfl_sync(peers, g0, src), b1 = __shfl_sync(peers, g1, src), b2 = __shfl_sync(peers, g2, src);
        if ((int)lane == leader) { g0 += b0; g1 += b1; g2 += b2; }
      }
      if ((int)lane == leader) {
        if (out.acc_in_smem == 2) { float* ga = out.block_acc + (size_t)blockIdx.x * out.acc_stride + 3 * vtri; atomicAdd(ga + 0, g0); atomicAdd(ga + 1, g1); atomicAdd(ga + 2, g2); }
        else if (out.acc_in_smem) { float* ga = sacc + 3 * vtri; atomicAdd(ga + 0, g0); atomicAdd(ga + 1, g1); atomicAdd(ga + 2, g2); }
        else { double* gg = out.global_acc + 3 * (size_t)vtri; atomicAdd(gg + 0, (double)g0); atomicAdd(gg + 1, (double)g1); atomicAdd(gg + 2, (double)g2); }
      }
#else
      if (out.acc_in_smem == 2) { float* ga = out.block_acc + (size_t)blockIdx.x * out.acc_stride + 3 * vtri; atomicAdd(ga + 0, sw0 * g0); atomicAdd(ga + 1, sw1 * g1); atomicAdd(ga + 2, sw2 * g2); }
      else if (out.acc_in_smem) { float* ga = sacc + 3 * vtri; atomicAdd(ga + 0, sw0 * g0); atomicAdd(ga + 1, sw1 * g1); atomicAdd(ga + 2, sw2 * g2); }
      else { double* gg = out.global_acc + 3 * (size_t)vtri; atomicAdd(gg + 0, (double)(sw0 * g0)); atomicAdd(gg + 1, (double)(sw1 * g1)); atomicAdd(gg + 2, (double)(sw2 * g2)); }
#endif
      if (PLAIN) {
        pb0 = fmaf(kd.x * vb, pb0, se0 + kd.x * k0); pb1 = fmaf(kd.y * vb, pb1, se1 + kd.y * k1); pb2 = fmaf(kd.z * vb, pb2, se2 + kd.z * k2);
      } else {   // recorded factors: exact with a Phong lobe as well
        pb0 = fmaf(v.t.x, pb0, se0 + v.d.x * k0); pb1 = fmaf(v.t.y, pb1, se1 + v.d.y * k1); pb2 = fmaf(v.t.z, pb2, se2 + v.d.z * k2);
      }
    } else {
      // miss: the stale L_d of vertex k-1 was re-added with P_k
      float c0 = se0, c1 = se1, c2 = se2;
      if (k > 0) {
        AdjVertex u;
        u.a = vs[(k - 1) * kVS];
        if (!PLAIN) u.d = vs[(k - 1) * kVS + 3];
        const int utri = adj_tri(u), ulight = adj_light(u);
        if (utri >= 0 && ulight >= 0) {
          const float4 lke = M(1, ulight);
          const float4 kd = M(0, utri);
          const float ug = u.a.w;
          const float k0 = lke.x * ug, k1 = lke.y * ug, k2 = lke.z * ug;
          if (out.acc_in_smem == 2) { float* ga = out.block_acc + (size_t)blockIdx.x * out.acc_stride + 3 * utri; atomicAdd(ga + 0, sw0 * (v.p.x * k0)); atomicAdd(ga + 1, sw1 * (v.p.y * k1)); atomicAdd(ga + 2, sw2 * (v.p.z * k2)); }
          else if (out.acc_in_smem) { float* ga = sacc + 3 * utri; atomicAdd(ga + 0, sw0 * (v.p.x * k0)); atomicAdd(ga + 1, sw1 * (v.p.y * k1)); atomicAdd(ga + 2, sw2 * (v.p.z * k2)); }
          else { double* gg = out.global_acc + 3 * (size_t)utri; atomicAdd(gg + 0, (double)(sw0 * (v.p.x * k0))); atomicAdd(gg + 1, (double)(sw1 * (v.p.y * k1))); atomicAdd(gg + 2, (double)(sw2 * (v.p.z * k2))); }
          if (PLAIN) { c0 += kd.x * k0; c1 += kd.y * k1; c2 += kd.z * k2; }
          else { c0 += u.d.x * k0; c1 += u.d.y * k1; c2 += u.d.z * k2; }
        }
      }
      pb0 = c0; pb1 = c1; pb2 = c2;
    }
  };

  // finalise one sample: toneMap's running sum of src[i]/SPP (path_trace.cu:190-193) / park the adjoint sweep
  auto finalize = [&]() {
    float4 part = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (MODE == MODE_RENDER) {
      part = lds4(item_a);
      part.x += L0 * rp.inv_spp; part.y += L1 * rp.inv_spp; part.z += L2 * rp.inv_spp;
    }
    if (MODE == MODE_GRAD) {
      // the previous path's sweep has not caught up (this path was shorter): finishing it here would run the sweep with
      // one or two lanes of the warp (measured: 22 % of the sweep steps, half of the sweep's instructions), so the lane
      // waits instead -- it starts no new sample until the shared per-trip steps below are through, then comes back here
      if (sw_k >= 0) { fin_wait = true; return; }
      sw_buf = vbuf; vbuf = out.vdepth - vbuf;  // the next path records into the other half
      sw_k = min(path_len, out.vdepth) - 1;
      pb0 = pb1 = pb2 = 0.0f;
      load_le(se0, se1, se2);
      sw0 = w0; sw1 = w1; sw2 = w2;
    }
    fin_left--;
    if (MODE == MODE_RENDER) {
      if (fin_left == 0) {
        float* dst = out.partial + (size_t)__float_as_uint(part.w) * 3;
        dst[0] = part.x; dst[1] = part.y; dst[2] = part.z;
      } else {
        sts4(item_a, part);
      }
    }
  };

  // one round of camera-ray generation: every lane with a free ring slot makes the ray of its next sample (path_trace.cu:
  // 155-165, Ray::transform scene_basics.h:307-319), claiming its next work item first when the current one is used up
  // (one warp-aggregated atomic).  Called by the whole warp.
  auto generate_round = [&]() {
    const bool gen = !g_done && rcount < kRing;
    float4 ga = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    int s_nx = 0, s_en = 0;
    unsigned it_next = 0, it_end = 0;
    if (gen) {
      ga = lds4(ring_a + kRing * kRingStride);           // {pixel, column, row, item}
      const float4 gb = lds4(ring_a + (kRing + 1) * kRingStride);
      s_nx = __float_as_int(gb.x); s_en = __float_as_int(gb.y);
      it_next = __float_as_uint(gb.z); it_end = __float_as_uint(gb.w);
    }
    const bool fresh = gen && s_nx == s_en;              // the generator's item is used up: next item of its run, or a new claim
    const bool claim = fresh && it_next >= it_end;
    const unsigned fm = __ballot_sync(kFull, claim);
    if (fm) {
      // guided self-scheduling: a claim is a run of n consecutive items per lane, n shrinking from 8 to 1 as the queue
      // empties (few atomics while there is plenty of work, single items at the end so the launch does not end in a tail)
      const int leader = __ffs(fm) - 1;
      const unsigned claim_base = (unsigned)__shfl_sync(kFull, lds1i(claim_a), leader);   // warp-uniform n: the leader's view
      const unsigned remaining = n_items > claim_base ? n_items - claim_base : 0u;
      const unsigned n = min(max(remaining / (gridDim.x * blockDim.x * 2u), 1u), (unsigned)IPT_CLAIM_MAX);
      unsigned base = 0;
      if ((int)lane == leader) base = atomicAdd(out.work_counter, (unsigned)__popc(fm) * n);   // overshoot <= 8 per lane: no wrap (make_plan)
      base = __shfl_sync(kFull, base, leader);
      sts1i(claim_a, (int)base);
      if (claim) {
        it_next = base + (unsigned)__popc(fm & ((1u << lane) - 1u)) * n;
        it_end = min(it_next + n, n_items);
        if (it_next >= n_items) g_done = true;
      }
    }
    if (fresh && !g_done) {
      const unsigned my = it_next++;
      int pix_local, g, px, r, c;
      decode_item(my, pix_local, g, px, r, c);
      s_nx = g * rp.group_samples;
      s_en = min(rp.spp, s_nx + rp.group_samples);
      ga = make_float4(__int_as_float(px), (float)c, (float)r, __uint_as_float(my));
      sts4(ring_a + kRing * kRingStride, ga);
    }
    if (gen && !g_done) {
      const u4 rn = philox4x32_10((uint32_t)__float_as_int(ga.x), (uint32_t)s_nx, 0u, IPT_STREAM_CAMERA, rp.keys);
      const float x = fmaf(ga.y + u01(rn.x), rp.two_over_w, -1.0f);
      const float y = fmaf(ga.z + u01(rn.y), -rp.two_over_h, 1.0f);
      const f3 dl = mk3(x, y, 1.0f);
      const f3 dir = IPT_NORMALIZE(mk3(dot3(sc.cam[0], sc.cam[1], sc.cam[2], dl), dot3(sc.cam[4], sc.cam[5], sc.cam[6], dl), dot3(sc.cam[8], sc.cam[9], sc.cam[10], dl)));
      sts4(ring_a + (uint32_t)((rhead + rcount) & (kRing - 1)) * kRingStride, make_float4(dir.x, dir.y, dir.z, ga.w));
      rcount++;
      sts4(ring_a + (kRing + 1) * kRingStride, make_float4(__int_as_float(s_nx + 1), __int_as_float(s_en), __uint_as_float(it_next), __uint_as_float(it_end)));
    }
  };

  for (;;) {
    __syncwarp();
    // =========== [A] shade the hit of the extension ray traced in the previous trip =================
    // (one radiance() call, path_trace.cu:111-144 / inv_path_trace.cu:109-150; the contribution
    //  prevMultiplier * (L_e + L_d) is applied in [D] once this bounce's shadow ray is resolved)
    bool ext_valid = false, sh_valid = false, end_after = false;
    sh_dir = zero3;   // a ray slot that carries no ray has d = 0 (denom = 0 fails every plane's MIN_DOT test)
    if (active) {
      if (h.tri < 0) {
        // miss (path_trace.cu:116): the stale L_e / L_d are re-added with the new multiplier
        if (MODE != MODE_GRAPH) { float e0, e1, e2; load_le(e0, e1, e2); L0 = fmaf(q0, e0 + d0, L0); L1 = fmaf(q1, e1 + d1, L1); L2 = fmaf(q2, e2 + d2, L2); }
        if (MODE == MODE_GRAD) {
          if (bounce < out.vdepth) { float4* vp = vw + (size_t)(vbuf + bounce) * kVS; vp[0] = make_float4(__int_as_float(-1), 0.0f, __int_as_float(-1), 0.0f); vp[1] = make_float4(q0, q1, q2, 0.0f); }
          path_len = bounce + 1;
        }
        finalize();
      } else {
        const int tri = h.tri;
        const float4 ma = M(0, tri), mb = M(1, tri);
        const int flags = PLAIN ? (__float_as_int(mb.w) & TRI_NEG_Z) : __float_as_int(mb.w);
        const f3 face_n = xyz(M(3, tri));
        const f3 hit = mk3(fmaf(rdir.x, h.t, rpos.x), fmaf(rdir.y, h.t, rpos.y), fmaf(rdir.z, h.t, rpos.z));
        const f3 ns = PLAIN ? face_n : shading_normal(sc, tri, flags, face_n, hit);
        if (MODE != MODE_GRAPH && bounce == 0) e_tri = tri;
        const u4 rn = philox4x32_10((uint32_t)pixel, (uint32_t)cur_s, (uint32_t)bounce, IPT_STREAM_BOUNCE, rp.keys);
        AdjVertex vx;
        if (MODE == MODE_GRAD) {
          vx.a = make_float4(__int_as_float(tri), 0.0f, __int_as_float(-1), 0.0f); vx.p = make_float4(q0, q1, q2, 0.0f);
          vx.t = make_float4(0.0f, 0.0f, 0.0f, 0.0f); vx.d = vx.t;
        }

        if (MODE == MODE_GRAPH) {
          // dw->update(dst, src, prev_weight, pixel, 0, prev_factors): prev_factors[DIFFUSE] == 1
          const size_t e = (size_t)gdst * nT + tri;
          if (out.acc_in_smem == 2) {
            float* acc = out.block_acc + (size_t)blockIdx.x * out.acc_stride + e * kGraphAccPerEdge;   // RED.ADD.F32, no CAS loop
            atomicAdd(acc + 0, gweight);
            atomicAdd(acc + 1, gweight * pix0); atomicAdd(acc + 2, gweight * pix1); atomicAdd(acc + 3, gweight * pix2);
            atomicAdd(reinterpret_cast<int*>(acc + 8), 1);
          } else if (out.acc_in_smem) {
            float* acc = sacc + e * kGraphAccPerEdge;
            atomicAdd(acc + 0, gweight);
            atomicAdd(acc + 1, gweight * pix0); atomicAdd(acc + 2, gweight * pix1); atomicAdd(acc + 3, gweight * pix2);
            atomicAdd(reinterpret_cast<int*>(acc + 8), 1);
          } else {
            double* ga = out.global_acc + e * kGraphAccPerEdge;
            atomicAdd(ga + 0, (double)gweight);
            atomicAdd(ga + 1, (double)(gweight * pix0)); atomicAdd(ga + 2, (double)(gweight * pix1)); atomicAdd(ga + 3, (double)(gweight * pix2));
            atomicAdd(ga + 8, 1.0);
          }
        }

        // ---- directLighting, first half (path_trace.cu:30-71 / inv_path_trace.cu:16-55): pick the light
        //      point; the shadow ray itself goes into this trip's dual intersection
        if (nE > 0) {
          const float u_pick = u16(rn.x >> 16), r1 = u16(rn.y >> 16), r2 = u16(rn.y & 0xffffu);
          int li = 0;
          if (BVH) {
            float up = u_pick;
            if (nE > IPT_LIGHT_PICK16_MAX) up = u01(philox4x32_10((uint32_t)pixel, (uint32_t)cur_s, (uint32_t)bounce, IPT_STREAM_PICK, rp.keys).x);
            li = pick_light(slight, nE, up);
          } else {
            // pick_light's linear scan (path_trace.cu:39-51) over the shared-memory table (brute-force scenes have <= 128 lights)
            while (li < nE - 1 && !(lds1f(slight_a + (uint32_t)(3 * li) * 16u + 12u) >= u_pick)) li++;
          }
          const float4 la = LREC(li, 0), lb = LREC(li, 1), lc = LREC(li, 2);
          const float sq = IPT_SQRT_UNIT(r1);   // r1 in [2^-17, 1)
          const float c1 = 1.0f - sq, c2 = sq * (1.0f - r2), c3 = r2 * sq;
          const f3 ep = mk3(fmaf(c3, lc.x, fmaf(c2, lb.x, c1 * la.x)), fmaf(c3, lc.y, fmaf(c2, lb.y, c1 * la.y)), fmaf(c3, lc.z, fmaf(c2, lb.z, c1 * la.z)));
          const f3 toLight = IPT_NORMALIZE(ep - hit);
          const float cos_theta = dot3(ns, toLight);
          // written unconditionally (plain register definitions, no predicated copies); only read when sh_valid
          sh_valid = !(cos_theta < 0.0f);
          sh_org = hit;
          if (sh_valid) sh_dir = toLight;
          sh_light = __float_as_int(lc.w); sh_tri = tri;
          sh_cos = cos_theta; sh_pt = lb.w; sh_gw = gweight;
          if (MODE != MODE_GRAPH && !PLAIN) {   // (plain scenes: BSDF(isDirect) is Kd of sh_tri, re-read in [D])
            f0 = ma.x; f1 = ma.y; f2 = ma.z;   // BSDF(isDirect = true): Kd, no 1/pi (path_trace.cu:14-17)
            if (sh_valid && (flags & TRI_ANY_KS)) {
              const float4 mc = M(4, tri);
              const float co = phong_coeff(ns, rdir, toLight, ma.w);
              f0 = fmaf(mc.x, co, f0); f1 = fmaf(mc.y, co, f1); f2 = fmaf(mc.z, co, f2);
            }
            if (MODE == MODE_GRAD) vx.d = make_float4(f0, f1, f2, 0.0f);
          }
        }

        // ---- Russian roulette + continuation (path_trace.cu:130-141) -----------------------------
        const bool depth_ok = (rp.max_depth <= 0) || (bounce + 1 < rp.max_depth);
        if (u16(rn.x & 0xffffu) < rp.p_rr && depth_ok) {
          f3 nd;
          float p_sample = IPT_INV_PI_F;
          if (!PLAIN && MODE != MODE_GRAPH && (flags & TRI_SPECULAR)) {
            nd = sample_specular_dir(face_n, ma.w, u01(rn.z), u01(rn.w), &p_sample);
          } else {
            // sampleNextDir (path_trace.cu:91-109), diffuse: the rotation FromTwoVectors(z, n) is fixed per
            // triangle and was built at load with the same arithmetic
            float sphi, cphi;
#if IPT_EXPERIMENT_FAST_SINCOS   // measurement only: what a cheaper sin/cos(2 pi u) could buy (not the canon: MUFU approximations)
            __sincosf(6.283185307179586f * u01(rn.z), &sphi, &cphi);
#else
            sincos2pi(u01(rn.z), &sphi, &cphi);
#endif
            const float ut = u01(rn.w);
            const float cth = IPT_SQRT_UNIT(ut), sth = IPT_SQRT(1.0f - ut);   // ut in [2^-33, 1]; 1 - ut may be 0
            const f3 hemi = mk3(sth * cphi, sth * sphi, cth);
            if (flags & TRI_NEG_Z) {
              nd = -hemi;
            } else {
              const float4 r0 = M(5, tri), r1 = M(6, tri), r2 = M(7, tri);
              nd = mk3(dot3(r0.x, r0.y, r0.z, hemi), dot3(r1.x, r1.y, r1.z, hemi), dot3(r2.x, r2.y, r2.z, hemi));
            }
#if !IPT_EXPERIMENT_NO_ND_NORMALIZE   // measurement only: the rotated hemisphere direction is unit length up to rounding; the canon normalises it
            nd = IPT_NORMALIZE(nd);
#endif
          }
          const float cosn = dot3(nd, ns);
          if (MODE == MODE_GRAPH) {
            gweight = gweight * cosn;
            gweight = gweight * gscale;
            gdst = tri;
          } else {
            const float4 mp = M(2, tri);   // Kd / pi, divided once at load (same IEEE division)
            float i0 = mp.x, i1 = mp.y, i2 = mp.z;         // BSDF(isDirect = false)
            float coeff = cosn * rp.c_diff;                // canon v2: (1/(1/pi))/p_rr folded on the host
            if (!PLAIN && (flags & TRI_ANY_KS)) {
              const float4 mc = M(4, tri);
              const float co = phong_coeff(ns, rdir, nd, ma.w);
              i0 = fmaf(mc.x, co, i0); i1 = fmaf(mc.y, co, i1); i2 = fmaf(mc.z, co, i2);
              if (flags & TRI_SPECULAR) coeff = (cosn / p_sample) / rp.p_rr;
            }
            m0 = (m0 * i0) * coeff; m1 = (m1 * i1) * coeff; m2 = (m2 * i2) * coeff;
            if (MODE == MODE_GRAD) {
              vx.a.y = coeff / IPT_PI_F;   // d(throughput factor)/dKd, the same with or without a Phong lobe
              if (!PLAIN) vx.t = make_float4(i0 * coeff, i1 * coeff, i2 * coeff, 0.0f);
            }
          }
          rpos = hit;
          rdir = nd;
          ext_valid = true;
        } else {
          end_after = true;
        }
        if (MODE == MODE_GRAD) {
          if (bounce < out.vdepth) {
            float4* vp = vw + (size_t)(vbuf + bounce) * kVS;
            vp[0] = vx.a; vp[1] = vx.p;
            if (!PLAIN) { vp[2] = vx.t; vp[3] = vx.d; }
          }
          path_len = bounce + 1;
        }
        bounce++;
        if (!sh_valid) {
          // no light connection to wait for: L_d = 0, apply this bounce's contribution now
          d0 = d1 = d2 = 0.0f;
          if (MODE != MODE_GRAPH) {
            float e0, e1, e2;
            load_le(e0, e1, e2);
            L0 = fmaf(q0, e0 + d0, L0); L1 = fmaf(q1, e1 + d1, L1); L2 = fmaf(q2, e2 + d2, L2);
            q0 = m0; q1 = m1; q2 = m2;
          }
          if (end_after) { finalize(); end_after = false; }
        }
      }
      active = false;
    }

    // =========== [B] path regeneration: lanes without an extension ray start their next sample ========
    // A lane whose path is ending but still waits for its last shadow ray (end_after) may already start the next sample of
    // the SAME item; at an item boundary it idles one trip (the partial sums still belong to the old item).
    // (adjoint: no new sample while a finished path waits for its slot, and no early start of the next sample while the
    //  other stack half is still being swept -- the ending path will have to wait in it)
    {
      const bool want = !ext_valid && !(MODE == MODE_GRAD && (fin_wait || (end_after && sw_k >= 0)));
      if (__ballot_sync(kFull, want && rcount == 0 && !g_done)) generate_round();
      if (want && rcount > 0) {
        const float4 e = lds4(ring_a + (uint32_t)rhead * kRingStride);
        const unsigned e_item = __float_as_uint(e.w);
        const bool new_item = e_item != (unsigned)lds1i(item_a + 12u);
        if (!new_item || (fin_left == 0 && !end_after)) {
          if (new_item) {
            int pix_local, g, r, c;
            decode_item(e_item, pix_local, g, pixel, r, c);
            sts4(item_a, make_float4(0.0f, 0.0f, 0.0f, e.w));
            cur_s = g * rp.group_samples - 1;
            fin_left = min(rp.spp, (g + 1) * rp.group_samples) - g * rp.group_samples;
            if (MODE != MODE_RENDER) {
              const unsigned char* tp = out.target + (size_t)pixel * 3;
              pix0 = (float)tp[0] / 255.0f; pix1 = (float)tp[1] / 255.0f; pix2 = (float)tp[2] / 255.0f;   // inv_scene.h:74-77
            }
            if (MODE == MODE_GRAD) {
              const float* wp = out.dldpx + (size_t)pix_local * 3;
              w0 = wp[0]; w1 = wp[1]; w2 = wp[2];
            }
          }
          cur_s++;
          rhead = (rhead + 1) & (kRing - 1);
          rcount--;
          rpos = mk3(sc.cam[3], sc.cam[7], sc.cam[11]);
          rdir = xyz(e);
          m0 = m1 = m2 = 1.0f;
          bounce = 0;
          gweight = 1.0f;
          gdst = nT;
          if (!end_after) { q0 = q1 = q2 = 1.0f; L0 = L1 = L2 = 0.0f; e_tri = -1; d0 = d1 = d2 = 0.0f; }
          ext_valid = true;
        }
      }
    }
    if (__all_sync(kFull, !ext_valid && !sh_valid && fin_left == 0 && rcount == 0 && g_done)) break;
    if (!ext_valid) rdir = zero3;   // (kept in the state itself: a select at the call is re-evaluated inside the group loop)

    // =========== [C] closest hits: extension ray and shadow ray ==========================================
    HitRec hs;
    if (!BVH) {
      closest_hit_small2(ss, rpos, rdir, sh_org, sh_dir, h, hs);
    } else {
      h.t = INFINITY; h.tri = -1;
      hs.t = INFINITY; hs.tri = -1;
      if (ss.nGroups > 0) closest_hit_small2(ss, rpos, rdir, sh_org, sh_dir, h, hs);
      if (ext_valid) h = closest_hit_bvh(sc, rpos, rdir, bvh_stack, h);
      if (sh_valid) hs = closest_hit_bvh(sc, sh_org, sh_dir, bvh_stack, hs);
    }
    n_ext += ext_valid ? 1u : 0u;
    n_shadow += sh_valid ? 1u : 0u;

    // =========== [D] second half of directLighting (path_trace.cu:73-88 / inv_path_trace.cu:57-86) =========
    if (sh_valid) {
      d0 = d1 = d2 = 0.0f;
      if (hs.tri >= 0) {
        const float4 lmb = M(1, sh_light);
        const f3 shit = mk3(fmaf(sh_dir.x, hs.t, sh_org.x), fmaf(sh_dir.y, hs.t, sh_org.y), fmaf(sh_dir.z, hs.t, sh_org.z));
        const f3 ln = PLAIN ? xyz(M(3, sh_light)) : shading_normal(sc, sh_light, __float_as_int(lmb.w), xyz(M(3, sh_light)), shit);
        const float cos_tp = -dot3(ln, sh_dir);
        if (!(cos_tp < 0.0f) && hs.tri == sh_light) {
          if (MODE == MODE_GRAPH) {
            const float w = ((sh_gw * sh_cos) * cos_tp) / ((hs.t * hs.t) * sh_pt);
            const size_t e = (size_t)sh_tri * nT + sh_light;
            if (out.acc_in_smem == 2) {
              float* acc = out.block_acc + (size_t)blockIdx.x * out.acc_stride + e * kGraphAccPerEdge;
              atomicAdd(acc + 4, w);
              atomicAdd(acc + 5, w * pix0); atomicAdd(acc + 6, w * pix1); atomicAdd(acc + 7, w * pix2);
              atomicAdd(reinterpret_cast<int*>(acc + 9), 1);
            } else if (out.acc_in_smem) {
              float* acc = sacc + e * kGraphAccPerEdge;
              atomicAdd(acc + 4, w);
              atomicAdd(acc + 5, w * pix0); atomicAdd(acc + 6, w * pix1); atomicAdd(acc + 7, w * pix2);
              atomicAdd(reinterpret_cast<int*>(acc + 9), 1);
            } else {
              double* ga = out.global_acc + e * kGraphAccPerEdge;
              atomicAdd(ga + 4, (double)w);
              atomicAdd(ga + 5, (double)(w * pix0)); atomicAdd(ga + 6, (double)(w * pix1)); atomicAdd(ga + 7, (double)(w * pix2));
              atomicAdd(ga + 9, 1.0);
            }
          } else {
            const float g = (sh_cos * cos_tp) / ((hs.t * hs.t) * sh_pt);
            if (PLAIN) { const float4 kd = M(0, sh_tri); f0 = kd.x; f1 = kd.y; f2 = kd.z; }   // BSDF(isDirect = true) = Kd (path_trace.cu:14-17)
            d0 = f0 * (lmb.x * g); d1 = f1 * (lmb.y * g); d2 = f2 * (lmb.z * g);
            if (MODE == MODE_GRAD) {
              const int k = path_len - 1;   // the vertex recorded in [A] of this trip
              if (k < out.vdepth) *reinterpret_cast<float2*>(reinterpret_cast<float*>(vw + (size_t)(vbuf + k) * kVS) + 2) = make_float2(__int_as_float(sh_light), g);
            }
          }
        }
      }
      if (MODE != MODE_GRAPH) {
        // renderSample loop body (path_trace.cu:176-180)
        float e0, e1, e2;
        load_le(e0, e1, e2);
        L0 = fmaf(q0, e0 + d0, L0); L1 = fmaf(q1, e1 + d1, L1); L2 = fmaf(q2, e2 + d2, L2);
        q0 = m0; q1 = m1; q2 = m2;
      }
      if (end_after) {
        finalize();
        // a sample started in [B] of this trip inherits clean accumulators (its multiplier is already 1); a path that
        // has to wait for its sweep slot started none, and its L_e is still needed when it is parked
        if (!(MODE == MODE_GRAD && fin_wait)) {
          L0 = L1 = L2 = 0.0f; e_tri = -1; d0 = d1 = d2 = 0.0f;
          q0 = m0; q1 = m1; q2 = m2;
        }
      }
    }
    if (MODE == MODE_GRAD) {
      // reconverge first: the lanes arrive from different arms of [D] and the point of the deferred sweep is that they
      // run it together
#pragma unroll 1
      for (int r = 0; r < kSweepPerTrip; r++) {
        const unsigned pending = __ballot_sync(kFull, sw_k >= 0);
        if (__popc(pending) < kSweepQuorum && !__any_sync(kFull, fin_wait)) break;   // a waiting lane must make progress
        if (sw_k >= 0) sweep_step();
      }
      if (fin_wait && sw_k < 0) { fin_wait = false; finalize(); }
    }
    active = ext_valid;
  }
  if (MODE == MODE_GRAD) {
    while (sw_k >= 0) sweep_step();
  }

  // ---- epilogue: counters + per-block accumulators --------------------------------------------------
  // every work item is traced to the end, so the launch traced exactly n_pix * spp paths
  unsigned long long w_paths = (blockIdx.x == 0 && threadIdx.x == 0) ? (unsigned long long)rp.n_pix * (unsigned long long)rp.spp : 0ull;
  unsigned long long w_ext = n_ext, w_shadow = n_shadow;
  for (int o = 16; o > 0; o >>= 1) {
    w_paths += __shfl_down_sync(kFull, w_paths, o);
    w_ext += __shfl_down_sync(kFull, w_ext, o);
    w_shadow += __shfl_down_sync(kFull, w_shadow, o);
  }
  if (lane == 0 && out.counters) {
    atomicAdd(out.counters + 0, w_paths);
    atomicAdd(out.counters + 1, w_ext);
    atomicAdd(out.counters + 2, w_shadow);
  }
  if (MODE != MODE_RENDER && out.acc_in_smem == 1) {
    __syncthreads();
    float* dst = out.block_acc + (size_t)blockIdx.x * out.acc_stride;
    for (int i = threadIdx.x; i < out.acc_stride; i += blockDim.x) dst[i] = sacc[i];
  }
}

// ------------------------------------------------------------------------------------------------
// toneMap (path_trace.cu:186-198): pixel = sum of group partials in group order; u8 = trunc(255 L/(1+L))
// ------------------------------------------------------------------------------------------------
__global__ void k_resolve(const float* __restrict__ partial, int n_pix, int groups, float* __restrict__ img, unsigned char* __restrict__ img8) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pix) return;
  const float* p = partial + (size_t)i * groups * 3;
  float a0 = p[0], a1 = p[1], a2 = p[2];
  for (int g = 1; g < groups; g++) { a0 += p[3 * g]; a1 += p[3 * g + 1]; a2 += p[3 * g + 2]; }
  if (img) { img[3 * (size_t)i] = a0; img[3 * (size_t)i + 1] = a1; img[3 * (size_t)i + 2] = a2; }
  if (img8) {
    img8[3 * (size_t)i] = (unsigned char)(255.0f * a0 / (1.0f + a0));
    img8[3 * (size_t)i + 1] = (unsigned char)(255.0f * a1 / (1.0f + a1));
    img8[3 * (size_t)i + 2] = (unsigned char)(255.0f * a2 / (1.0f + a2));
  }
}

// loss = sum (y - y*)^2 / (3 W H),  y = L/(1+L);  dldpx = dLoss/dL / spp   (double accumulation)
__global__ void k_loss_dldpx(const float* __restrict__ img, const unsigned char* __restrict__ target, int row_begin, int row_step, int width,
                             int n_pix, double inv_n, float inv_spp, float* __restrict__ dldpx, double* __restrict__ loss) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  if (i < n_pix) {
    const int lr = i / width;
    const size_t gpix = (size_t)(row_begin + lr * row_step) * width + (i - lr * width);
    for (int c = 0; c < 3; c++) {
      const double L = (double)img[3 * (size_t)i + c];
      const double y = L / (1.0 + L);
      const double ys = (double)target[3 * gpix + c] / 255.0;
      acc += (y - ys) * (y - ys) * inv_n;
      dldpx[3 * (size_t)i + c] = (float)(2.0 * (y - ys) * inv_n / ((1.0 + L) * (1.0 + L)) * (double)inv_spp);
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(kFull, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss, acc);
}

// sum per-block accumulators in double (block order is fixed -> deterministic given the partials)
__global__ void k_reduce_blocks(const float* __restrict__ block_acc, int n_blocks, int stride, double* __restrict__ out, int ints_from, int ints_per) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= stride) return;
  const bool is_int = ints_per > 0 && (i % ints_per) >= ints_from;
  double a = 0.0;
  for (int b = 0; b < n_blocks; b++) {
    const float v = block_acc[(size_t)b * stride + i];
    a += is_int ? (double)__float_as_int(v) : (double)v;
  }
  out[i] += a;
}

// acc10 (W_i, Wpix_i[3], W_d, Wpix_d[3], n_i, n_d) -> the reference's Edge fields (inv_scene.h:9-36):
// [n, w_sum, pixel_sum[2][3], light_sum[2][3], factors_sum[2]]; the SPECULAR lane is identically 0
// because P_SPEC == 0 (inv_scene.h:5).  Indirect updates carry factor 1, direct ones 1/pi and the
// emission of the light triangle `src` (inv_path_trace.cu:8,76-84,128).
__global__ void k_graph_finalize(const double* __restrict__ acc, const float4* __restrict__ mat, int nT, int n_edges, double* __restrict__ raw) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const double* a = acc + (size_t)e * kGraphAccPerEdge;
  double* o = raw + (size_t)e * 16;
  const int src = e % nT;
  const float4 ke = mat[1 * nT + src];
  const double ip = (double)IPT_INV_PI_F;
  o[0] += a[8] + a[9];
  o[1] += a[0] + a[4];
  o[2] += a[1] + a[5] * ip; o[3] += a[2] + a[6] * ip; o[4] += a[3] + a[7] * ip;
  const double wd = a[4] * ip;
  o[8] += (double)ke.x * wd; o[9] += (double)ke.y * wd; o[10] += (double)ke.z * wd;
  o[14] += a[0] + wd;
  if (e / nT == nT) {
    // eye row: the path starts with factors = {1, 1} (inv_path_trace.cu:172), so the first update also
    // feeds the SPECULAR lane; every later update carries factors[SPECULAR] = 0
    o[5] += a[1]; o[6] += a[2]; o[7] += a[3];
    o[15] += a[0];
  }
}

// albedo update from device memory: Kd and the load-time quotient Kd / pi (BSDF, path_trace.cu:15-17)
// DataWrapper::compress (inv_scene.h:87-115) on the device: one thread per destination row (dst = 0..nT-1, then the
// eye row), the row's weights summed in source order like the host version -- same operations, same bits
__global__ void k_graph_compress(const double* __restrict__ raw, int nT, float* __restrict__ data) {
  const int dst = blockIdx.x * blockDim.x + threadIdx.x;
  if (dst > nT) return;
  const size_t size = (size_t)(nT + 1) * nT;
  float* weights = data;
  float* pixels = data + size;
  float* lights = data + 4 * size;
  float total = 0.0f;
  for (int src = 0; src < nT; src++) {
    const double* e = raw + ((size_t)dst * nT + src) * 16;
    const float w = canon_logf((float)e[1] + 1.0f);
    const float fs = (float)e[14];
    const float div = (fs != 0.0f) ? fs : 1.0f;
    weights[(size_t)dst * nT + src] = w;
    total += w;
    for (int i = 0; i < 3; i++) {
      pixels[((size_t)dst * nT + src) * 3 + i] = (float)e[2 + i] / div;
      lights[((size_t)dst * nT + src) * 3 + i] = (float)e[8 + i] / div;
    }
  }
  for (int src = 0; src < nT; src++) {
    const float w = weights[(size_t)dst * nT + src];
    weights[(size_t)dst * nT + src] = (total != 0.0f) ? w / total : 0.0f;
  }
}

__global__ void k_set_kd(float4* __restrict__ mat, const float* __restrict__ kd, int nT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nT) return;
  const float r = kd[3 * i], g = kd[3 * i + 1], b = kd[3 * i + 2];
  float4 m0 = mat[i];
  m0.x = r; m0.y = g; m0.z = b;
  mat[i] = m0;
  mat[2 * nT + i] = make_float4(r / IPT_PI_F, g / IPT_PI_F, b / IPT_PI_F, 0.0f);
}

// ------------------------------------------------------------------------------------------------
// host-callable launchers
// ------------------------------------------------------------------------------------------------

size_t paths_smem_bytes(const DevScene& sc, int acc_floats) {
  size_t b = (size_t)acc_floats * sizeof(float) + (size_t)kRingWords * kBlock * sizeof(float);   // accumulators + camera-ray rings
  if (!sc.use_bvh || sc.nE <= kLightSmemMax) b += (size_t)3 * sc.nE * sizeof(float4);
  if (!sc.use_bvh) b += small_scene_smem_bytes(sc.nStream) + (size_t)kMatStride * sc.nT * sizeof(float4);
  else b += bvh_stack_bytes(sc) + (sc.nGroups > 0 ? small_scene_smem_bytes(sc.nStream) : 0);
  return b;
}

template <int MODE, bool BVH, bool PLAIN>
static cudaError_t launch_paths_t(const DevScene& sc, const DevRender& rp, const DevOut& out, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_paths<MODE, BVH, PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_paths<MODE, BVH, PLAIN><<<grid, kBlock, smem, st>>>(sc, rp, out);
  return cudaGetLastError();
}
template <int MODE, bool BVH, bool PLAIN>
static int blocks_per_sm_t(size_t smem) {
  int nb = 0;
  cudaFuncSetAttribute(k_paths<MODE, BVH, PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_paths<MODE, BVH, PLAIN>, kBlock, smem) != cudaSuccess || nb < 1) nb = 1;
  return nb;
}

// dispatch over (mode, BVH, plain): `what` = 0 launches, 1 queries resident blocks per SM
template <int MODE>
static int paths_dispatch(int what, const DevScene& sc, const DevRender* rp, const DevOut* out, int grid, size_t smem, cudaStream_t st, cudaError_t* err) {
  const bool bvh = sc.use_bvh != 0, plain = sc.plain != 0;
#define IPT_CASE(B, P)                                                                                    \
  if (bvh == B && plain == P) {                                                                           \
    if (what == 1) return blocks_per_sm_t<MODE, B, P>(smem);                                              \
    *err = launch_paths_t<MODE, B, P>(sc, *rp, *out, grid, smem, st);                                     \
    return 0;                                                                                             \
  }
  IPT_CASE(false, false) IPT_CASE(false, true) IPT_CASE(true, false) IPT_CASE(true, true)
#undef IPT_CASE
  return 0;
}

cudaError_t launch_paths(int mode, const DevScene& sc, const DevRender& rp, const DevOut& out, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaSuccess;
  switch (mode) {
    case MODE_RENDER: paths_dispatch<MODE_RENDER>(0, sc, &rp, &out, grid, smem, st, &e); break;
    case MODE_GRAPH: paths_dispatch<MODE_GRAPH>(0, sc, &rp, &out, grid, smem, st, &e); break;
    default: paths_dispatch<MODE_GRAD>(0, sc, &rp, &out, grid, smem, st, &e); break;
  }
  return e;
}

int paths_block_threads() { return kBlock; }

int paths_blocks_per_sm(int mode, const DevScene& sc, size_t smem) {
  cudaError_t e;
  switch (mode) {
    case MODE_RENDER: return paths_dispatch<MODE_RENDER>(1, sc, nullptr, nullptr, 0, smem, nullptr, &e);
    case MODE_GRAPH: return paths_dispatch<MODE_GRAPH>(1, sc, nullptr, nullptr, 0, smem, nullptr, &e);
    default: return paths_dispatch<MODE_GRAD>(1, sc, nullptr, nullptr, 0, smem, nullptr, &e);
  }
}

cudaError_t launch_resolve(const float* partial, int n_pix, int groups, float* img, unsigned char* img8, cudaStream_t st) {
  k_resolve<<<(n_pix + 255) / 256, 256, 0, st>>>(partial, n_pix, groups, img, img8);
  return cudaGetLastError();
}

cudaError_t launch_loss_dldpx(const float* img, const unsigned char* target, int row_begin, int row_step, int width, int n_pix, double inv_n,
                              float inv_spp, float* dldpx, double* loss, cudaStream_t st) {
  k_loss_dldpx<<<(n_pix + 255) / 256, 256, 0, st>>>(img, target, row_begin, row_step, width, n_pix, inv_n, inv_spp, dldpx, loss);
  return cudaGetLastError();
}

cudaError_t launch_reduce_blocks(const float* block_acc, int n_blocks, int stride, double* out, int ints_from, int ints_per, cudaStream_t st) {
  k_reduce_blocks<<<(stride + 255) / 256, 256, 0, st>>>(block_acc, n_blocks, stride, out, ints_from, ints_per);
  return cudaGetLastError();
}

cudaError_t launch_graph_finalize(const double* acc, const float4* mat, int nT, int n_edges, double* raw, cudaStream_t st) {
  k_graph_finalize<<<(n_edges + 255) / 256, 256, 0, st>>>(acc, mat, nT, n_edges, raw);
  return cudaGetLastError();
}

cudaError_t launch_graph_compress(const double* raw, int nT, float* data, cudaStream_t st) {
  k_graph_compress<<<(nT + 1 + 127) / 128, 128, 0, st>>>(raw, nT, data);
  return cudaGetLastError();
}
cudaError_t launch_set_kd(float4* mat, const float* kd, int nT, cudaStream_t st) {
  if (nT > 0) k_set_kd<<<(nT + 255) / 256, 256, 0, st>>>(mat, kd, nT);
  return cudaGetLastError();
}

}  // namespace ipt
