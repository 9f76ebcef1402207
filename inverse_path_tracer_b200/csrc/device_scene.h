// device_scene.h -- flat device-side scene layout and kernel launch descriptors.
//
// HBM layout.  All arrays are float4-vectorised so one triangle test is
// four 16-byte loads and the whole small scene can be staged into shared memory with float4 copies.
//   geom  [nT] : {n.xyz, k=dot(c,n)}, geom_edges [3*nT] : {e0.xyz,e0.d} {e1.xyz,e1.d} {e2.xyz,e2.d}     16 + 48 B / triangle, leaf order (BVH leaves)
//   stream [nStream] : the brute-force set as ONE record stream, walked with a single running address:
//     per plane group (triangles whose (n, k) are value-equal share one plane test)
//       {n.xyz, k}                                 the plane
//       {bytes to the next group (int bits), -, -, -}
//       per quad slot (two coplanar triangles sharing an exactly negated edge plane; five edge planes instead of six)
//         {out.xyz, d} x 5                         shared diagonal, 2 edges of a, 2 edges of b
//         {id a, id b (int bits; b = -1 unpaired), -, -}
//   mat   [8*nT] : {Kd.rgb, Ns} {Ke.rgb, flags(bits)} {Kd/pi .rgb, 0} {n.xyz, 0} {Ks.rgb, area}
//                  {R row0, 0} {R row1, 0} {R row2, 0}   (R = FromTwoVectors(z, n), path_trace.cu:102)   128 B / triangle
//   light [3*nE] : {v0.xyz, cdf} {v1.xyz, p_t} {v2.xyz, tri index(bits)}             48 B / emissive
//   vtx   [5*nT] : {v0.xyz, vn0.x} {v1.xyz, vn0.y} {v2.xyz, vn0.z} {vn1.xyz, area} {vn2.xyz, 0}
//                  (only read for triangles that carry per-vertex normals: Triangle::getNormal)
//   bvh nodes [2*nNodes] : {lo.xyz, rightOffset(bits)} {hi.xyz, start<<3|nPrims (bits)}   32 B / node (canonical DFS layout)
//   bvh wide  [4*nNodes] : {L.lo, desc L} {L.hi, desc R} {R.lo, -} {R.hi, -}             64 B / inner node (float traversal records)
//   bvh q     [2*nNodes] : 6 words of 16-bit child box coordinates + 2 child descriptors   32 B / inner node (default traversal records)
#pragma once
#include <cuda_runtime.h>

#include "ipt_math.cuh"
#include <stdint.h>

// BVH traversal records (all three built and measured in round 2, profiles/r02_config5.txt; results are identical by construction --
// a quantised box contains its float box):
//   0 = 64-byte float records of a binary node (bvh_wide)                                   29.8 ms frame, 5.0 Grays/s incoherent
//   1 = 32-byte records, both child boxes at 16 bits per coordinate on the root box (bvh_q)  29.0 - 29.4 ms, 5.3 Grays/s   <- default
//   2 = 4-wide collapse of the tree, four 16-bit child boxes in 64 bytes (bvh_q)             33.5 ms, 4.8 Grays/s
// Half the bytes per step (1) buys 2 %, half the node visits per ray (2) costs 12 %: under ncu (2) runs 3 % more instructions and
// the same number of load requests as (1) (two 256-bit loads per 64-byte node) at a lower issue rate (profiles/r02_trace_bvh_summary.txt).
// (1) also halves the node memory (215 -> 107 MB for 5.2 M triangles: inside the L2).
#ifndef IPT_BVH_QUANT
#define IPT_BVH_QUANT 1
#endif

namespace ipt {

enum : int { TRI_HAS_VN = 1, TRI_SPECULAR = 2, TRI_ANY_KS = 4, TRI_NEG_Z = 8 /* normal.z == -1: R = -I, path_trace.cu:99-100 */ };

constexpr int kMatStride = 8;   // float4 records per triangle in `mat`

struct DevScene {
  const float4* geom;          // BVH scenes: plane {n, k} of leaf slot j (slot j = triangle bvh_prims[j]), [n slots] ...
  const float4* geom_edges;    // ... and its three edge planes [3 n slots]
  const float4* stream;     // record stream of the brute-force set (see above)
  int nGroups;
  int nStream;              // float4 records in `stream`
  const float4* mat;
  const float4* light;
  const float4* vtx;
  const float4* bvh_nodes;   // null in brute-force mode; canonical 32-byte nodes (root box, layout parity)
  const float4* bvh_wide;    // 64-byte traversal records: both child boxes + child descriptors per inner node (IPT_BVH_QUANT == 0 builds)
  // 32-byte traversal records (IPT_BVH_QUANT == 1 builds): both child boxes quantised to 16 bits per coordinate on the tree's root box,
  // rounded outwards -- six words {lo_l.x|lo_l.y<<16, lo_l.z|hi_l.x<<16, hi_l.y|hi_l.z<<16, same for the right child} + the two
  // child descriptors.  coordinate = bvh_qoff + q * bvh_qscale.  Half the bytes per traversal step: one 256-bit load, and
  // the 3.4 M inner nodes of the 5.2 M-triangle scene (107 MB) fit the L2
  const float4* bvh_q;
  float bvh_qoff[3], bvh_qscale[3];
  int bvh_root_desc;         // descriptor of the root (0 = inner node 0, negative = the root is a leaf)
  int bvh_depth;             // levels of the tree = entries of the per-thread traversal stack (<= 64)
  const int* bvh_prims;      // triangle index per BVH slot
  int nT, nE, nNodes;
  const unsigned char* big_slot;   // BVH scenes: per triangle, its index in the brute-force set (255: in the tree); null without a set
  const int* big_tri;              // ... and the triangle id of every member of the set [nBig]
  int nBig;                  // BVH scenes: triangles kept out of the tree and tested brute force (`stream` describes them)
  int use_bvh;
  int plain;                 // no per-vertex normals and no specular lobe anywhere in the scene
  float cam[16];
};

struct DevRender {
  int width, height, spp, max_depth;
  float p_rr;
  float c_diff;         // (1 / (1/pi)) / p_rr : cosine-to-throughput factor of a diffuse bounce (canon v2)
  float inv_spp;        // 1 / spp (canon v3: toneMap's src[i]/SPP as a product)
  float two_over_w, two_over_h;   // camera raster scale (canon v3)
  uint32_t key0, key1;
  PhiloxKeys keys;      // round keys of (key0, key1)
  int group_samples;    // samples per work item
  int groups;           // work items per pixel
  int row_begin;        // first image row of this shard
  int row_step;         // the shard owns rows row_begin, row_begin + row_step, ... (interleaved sharding)
  int n_pix;            // pixels in this shard (compact local index = local_row * width + column)
  // floor(x / d) for x < 2^31 and d >= 2 as (__umulhi(x, magic) >> shift), magic = floor(2^(31 + l) / d) + 1, shift = l - 1,
  // l = ceil(log2 d) (a 32-bit magic number is exact for 31-bit dividends); d == 1 is handled by the caller
  unsigned magic_groups, magic_width;
  int shift_groups, shift_width;
};

// counters[0..2] = paths, extension rays, shadow rays
struct DevOut {
  unsigned* work_counter;       // next unclaimed work item (zeroed before the launch)
  unsigned long long* counters;
  float* partial;            // [n_pix * groups * 3]  forward
  const unsigned char* target;   // full-image RGB8 target (graph / grad)
  const float* dldpx;        // [n_pix * 3] dLoss/dL_pixel / spp  (grad)
  float* block_acc;          // [gridDim.x * acc_stride] per-block partial sums (graph / grad)
  double* global_acc;        // direct global accumulation when the scene does not fit shared memory
  float4* vstacks;           // adjoint vertex stacks: per warp [2 halves][vdepth][32 lanes] x 32-byte vertices (grad)
  int vdepth;                // vertices per stack half (= clamped max_depth)
  int acc_stride;
  int acc_in_smem;           // 0: double atomics on global_acc; 1: per-block tile in shared memory, flushed to block_acc;
                             // 2: float reductions (RED.ADD.F32, no CAS loop) straight into this block's tile of block_acc
};

constexpr int kGraphAccPerEdge = 10;   // W_i, Wpix_i[3], W_d, Wpix_d[3], n_i, n_d
constexpr int kGradMaxDepth = 64;      // vertex stack depth of the adjoint pass

}  // namespace ipt
