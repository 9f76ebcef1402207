// bvh_build.cu -- BVH construction with the reference's builder semantics (BVH::build, bvh.h:109-205).
//
// Node layout and primitive order must be bit-exact against bvh.h: midpoint split of the CENTROID
// box on its widest axis (ties -> lowest axis, scene_basics.h:359-369), Lomuto partition with
// in-place swaps (bvh.h:178-186), fallback mid = start + (end-start)/2 when one side is empty
// (bvh.h:188-190), leaves at <= leaf_size primitives, nodes numbered in DFS pre-order with the left
// child at parent+1 and rightOffset = distance to the right child (bvh.h:163-169, 192-200).
//
// Formulation used here (differs from the reference's explicit work stack): a subtree over
// [start,end) occupies a contiguous block of node indices, so a node's index is known as soon as
// the size of everything built before it is known.  Build proceeds level by level over *segments*:
//   1. per segment: reduce primitive boxes and centroid boxes            (parallel over primitives)
//   2. per segment: pick axis / split, evaluate the predicate            (parallel over primitives)
//   3. per segment: apply the Lomuto permutation in closed form          (parallel over primitives)
//   4. after the last level: subtree sizes bottom-up, node indices top-down (parallel over nodes)
// Closed form of step 3: let T_k be the k-th element with pos < split and
// i_k its index.  out[start+k] = in[i_k].  For a slot q >= start+nT: if in[q] is false it stays;
// if in[q] is T_k the element that ends there is found by following pos = start+k while in[pos]
// is T_j -> pos = start+j.  (Each swap moves the displaced false element to where the true one was.)
//
// Two executions of the same four steps live in this file: build_bvh_gpu (CUDA kernels, one launch
// set per tree level) and build_bvh_host (the same level-by-level formulation on one host thread,
// used for tiny inputs where kernel launches would dominate, and by the CPU-only tests).  Both are
// bit-exact against the oracle's sequential restatement of bvh.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.h"

namespace ipt {

namespace {

struct BuildNode {
  float lo[3], hi[3];
  int start, count;
  int left, right;   // build-node ids, -1 for leaves
  int subtree;       // nodes in this subtree
  int flat;          // DFS pre-order index
};

// min / max under the total order -0 < +0: independent of evaluation order, hence identical between the
// sequential host loop, the GPU's atomic reduction on order-preserving integer keys, and the oracle
inline float total_min(float a, float b) { return (b < a || (b == a && std::signbit(b))) ? b : a; }
inline float total_max(float a, float b) { return (b > a || (b == a && !std::signbit(b))) ? b : a; }

}  // namespace

static bool build_bvh_host(const float* boxes, const float* pos, int n, int leaf_size, std::vector<BvhNodeHost>* nodes, std::vector<int>* perm) {
  nodes->clear();
  perm->resize(n);
  for (int i = 0; i < n; i++) (*perm)[i] = i;
  if (n == 0) return true;

  std::vector<int>& order = *perm;
  std::vector<int> scratch(n);
  std::vector<BuildNode> bn;
  std::vector<int> level, next_level;   // build-node ids of the current level
  bn.push_back(BuildNode());
  bn[0].start = 0; bn[0].count = n; bn[0].left = bn[0].right = -1;
  level.push_back(0);

  while (!level.empty()) {
    next_level.clear();
    for (size_t li = 0; li < level.size(); li++) {
      const int id = level[li];
      const int start = bn[id].start, end = start + bn[id].count;
      // step 1: boxes
      float lo[3], hi[3], clo[3], chi[3];
      for (int k = 0; k < 3; k++) {
        lo[k] = boxes[6 * order[start] + k]; hi[k] = boxes[6 * order[start] + 3 + k];
        clo[k] = chi[k] = pos[3 * order[start] + k];
      }
      for (int q = start + 1; q < end; q++) {
        const int pr = order[q];
        for (int k = 0; k < 3; k++) {
          lo[k] = total_min(lo[k], boxes[6 * pr + k]); hi[k] = total_max(hi[k], boxes[6 * pr + 3 + k]);
          clo[k] = total_min(clo[k], pos[3 * pr + k]); chi[k] = total_max(chi[k], pos[3 * pr + k]);
        }
      }
      for (int k = 0; k < 3; k++) { bn[id].lo[k] = lo[k]; bn[id].hi[k] = hi[k]; }
      if (bn[id].count <= leaf_size) continue;
      // step 2: axis + split
      int axis = 0;
      float widest = 0.0f;
      for (int k = 0; k < 3; k++) {
        const float ext = chi[k] - clo[k];
        if (ext > widest) { widest = ext; axis = k; }
      }
      const float split = (clo[axis] + chi[axis]) / 2.0f;
      // step 3: Lomuto permutation in closed form
      int n_true = 0;
      for (int q = start; q < end; q++) if (pos[3 * order[q] + axis] < split) n_true++;
      int mid = start + n_true;
      if (n_true != 0 && n_true != end - start) {
        // rank of each true element, in order
        std::vector<int> true_rank(end - start, -1);
        int k = 0;
        for (int q = start; q < end; q++) if (pos[3 * order[q] + axis] < split) { scratch[start + k] = order[q]; true_rank[q - start] = k; k++; }
        for (int q = mid; q < end; q++) {
          int src = q;
          while (true_rank[src - start] >= 0) src = start + true_rank[src - start];   // pointer jumping
          scratch[q] = order[src];
        }
        for (int q = start; q < end; q++) order[q] = scratch[q];
      } else {
        // nothing moves when no element is true; when all are true every swap is with itself
        mid = start + (end - start) / 2;
      }
      BuildNode l, r;
      l.start = start; l.count = mid - start; l.left = l.right = -1;
      r.start = mid; r.count = end - mid; r.left = r.right = -1;
      bn[id].left = (int)bn.size(); bn.push_back(l);
      bn[id].right = (int)bn.size(); bn.push_back(r);
      next_level.push_back(bn[id].left);
      next_level.push_back(bn[id].right);
    }
    level.swap(next_level);
  }
  // step 4: subtree sizes (children always have larger ids than parents -> one reverse sweep)
  for (int id = (int)bn.size() - 1; id >= 0; id--)
    bn[id].subtree = 1 + (bn[id].left >= 0 ? bn[bn[id].left].subtree + bn[bn[id].right].subtree : 0);
  bn[0].flat = 0;
  for (size_t id = 0; id < bn.size(); id++) {
    if (bn[id].left >= 0) {
      bn[bn[id].left].flat = bn[id].flat + 1;
      bn[bn[id].right].flat = bn[id].flat + 1 + bn[bn[id].left].subtree;
    }
  }
  nodes->resize(bn.size());
  for (size_t id = 0; id < bn.size(); id++) {
    BvhNodeHost& o = (*nodes)[bn[id].flat];
    for (int k = 0; k < 3; k++) { o.lo[k] = bn[id].lo[k]; o.hi[k] = bn[id].hi[k]; }
    o.start = bn[id].start;
    o.nPrims = bn[id].count;
    o.rightOffset = bn[id].left >= 0 ? bn[bn[id].right].flat - bn[id].flat : 0;
  }
  return true;
}


// =================================================================================================
// GPU builder
// =================================================================================================
namespace {

#define BVH_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) { *err = std::string(#call) + ": " + cudaGetErrorString(e__); ok = false; goto done; } \
  } while (0)

struct GpuNode {       // build node (level order); converted to DFS pre-order at the end
  float lo[3], hi[3];
  int start, count;
  int left;            // build id of the left child (right = left + 1), -1 for a leaf
  int subtree, flat;
};

struct SegWork {       // per-segment scratch of the current level
  unsigned blo[3], bhi[3], clo[3], chi[3];   // order-preserving uint encodings of float min / max
  int axis;
  float split;
  int n_true;
  int mid;
  int is_split;
};

__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// step 1: reduce primitive boxes and centroid boxes per segment (warp-aggregated atomics)
__global__ void k_bvh_reduce(const float* __restrict__ boxes, const float* __restrict__ pos, const int* __restrict__ order,
                             const int* __restrict__ seg_of, int n, SegWork* __restrict__ work) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int seg = (q < n) ? seg_of[q] : -1;
  unsigned v[12];
  if (seg >= 0) {
    const int pr = order[q];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      v[k] = f2ord(boxes[6 * pr + k]); v[3 + k] = f2ord(boxes[6 * pr + 3 + k]);
      v[6 + k] = f2ord(pos[3 * pr + k]); v[9 + k] = v[6 + k];
    }
  }
  const unsigned active = __ballot_sync(0xffffffffu, seg >= 0);
  if (seg < 0) return;
  const unsigned peers = __match_any_sync(active, seg);
  const int leader = __ffs(peers) - 1;
  const bool whole = (peers == active) && (active == 0xffffffffu);
  if (whole) {
#pragma unroll
    for (int k = 0; k < 12; k++) {
      const bool is_min = (k < 3) || (k >= 6 && k < 9);
      v[k] = is_min ? __reduce_min_sync(0xffffffffu, v[k]) : __reduce_max_sync(0xffffffffu, v[k]);
    }
    if ((int)(threadIdx.x & 31) != leader) return;
  }
  SegWork* w = work + seg;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    atomicMin(&w->blo[k], v[k]); atomicMax(&w->bhi[k], v[3 + k]);
    atomicMin(&w->clo[k], v[6 + k]); atomicMax(&w->chi[k], v[9 + k]);
  }
}

// step 2a: per segment: node box, leaf test, split axis / coordinate (bvh.h:154-176)
__global__ void k_bvh_decide(SegWork* __restrict__ work, GpuNode* __restrict__ nodes, int first_node, int n_seg, int leaf_size) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  SegWork* w = work + s;
  GpuNode* nd = nodes + first_node + s;
  float clo[3], chi[3];
  for (int k = 0; k < 3; k++) {
    nd->lo[k] = ord2f(w->blo[k]); nd->hi[k] = ord2f(w->bhi[k]);
    clo[k] = ord2f(w->clo[k]); chi[k] = ord2f(w->chi[k]);
  }
  nd->left = -1;
  w->is_split = 0;
  w->n_true = 0;
  if (nd->count <= leaf_size) return;
  int axis = 0;
  float widest = 0.0f;
  for (int k = 0; k < 3; k++) {     // bbox_t::maxDimension: strict >, ties to the lowest axis
    const float ext = chi[k] - clo[k];
    if (ext > widest) { widest = ext; axis = k; }
  }
  w->axis = axis;
  w->split = (clo[axis] + chi[axis]) / 2.0f;
  w->is_split = 1;
}

// step 2b: predicate of every primitive slot (bvh.h:180)
__global__ void k_bvh_flags(const float* __restrict__ pos, const int* __restrict__ order, const int* __restrict__ seg_of, int n,
                            const SegWork* __restrict__ work, int* __restrict__ flag) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int seg = seg_of[q];
  int f = 0;
  if (seg >= 0 && work[seg].is_split) f = pos[3 * order[q] + work[seg].axis] < work[seg].split ? 1 : 0;
  flag[q] = f;
}

// exclusive scan, three kernels (block scan / scan of block sums / add); 1024 elements per block
__global__ void k_scan_block(const int* __restrict__ in, int* __restrict__ out, int n, int* __restrict__ sums) {
  __shared__ int warp_tot[32];
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const int v = (i < n) ? in[i] : 0;
  int x = v;
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((int)(threadIdx.x & 31) >= o) x += y; }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    int t = warp_tot[threadIdx.x];
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, t, o); if ((int)threadIdx.x >= o) t += y; }
    warp_tot[threadIdx.x] = t;
  }
  __syncthreads();
  const int base = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0;
  if (i < n) out[i] = base + x - v;
  if (threadIdx.x == 1023 && sums) sums[blockIdx.x] = base + x;
}
__global__ void k_scan_add(int* __restrict__ out, int n, const int* __restrict__ sums) {
  const int i = blockIdx.x * 1024 + threadIdx.x;
  if (i < n) out[i] += sums[blockIdx.x];
}

// step 3a: trues per segment and the split position (bvh.h:188-190 fallback included)
__global__ void k_bvh_mid(SegWork* __restrict__ work, const GpuNode* __restrict__ nodes, int first_node, int n_seg,
                          const int* __restrict__ scan, const int* __restrict__ flag, int n) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  SegWork* w = work + s;
  if (!w->is_split) return;
  const GpuNode* nd = nodes + first_node + s;
  const int start = nd->start, end = start + nd->count;
  const int before_end = (end < n) ? scan[end] : scan[n - 1] + flag[n - 1];
  const int nt = before_end - scan[start];
  w->n_true = nt;
  w->mid = (nt == 0 || nt == nd->count) ? start + nd->count / 2 : start + nt;
}

// step 3b: the Lomuto permutation in closed form.  Slot q of a split segment receives
//   q <  start + nT : the (q - start)-th true element                      (stable compaction)
//   q >= start + nT : itself when false; when true T_k, follow pos = start + k while that slot held a
//                     true element T_j -> pos = start + j  (each swap parks the displaced false there)
__global__ void k_bvh_permute(const int* __restrict__ order, int* __restrict__ order_out, const int* __restrict__ seg_of, int n,
                              const SegWork* __restrict__ work, const GpuNode* __restrict__ nodes, int first_node,
                              const int* __restrict__ flag, const int* __restrict__ scan) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int seg = seg_of[q];
  if (seg < 0) { order_out[q] = order[q]; return; }
  const SegWork* w = work + seg;
  const GpuNode* nd = nodes + first_node + seg;
  const int start = nd->start;
  if (!w->is_split || w->n_true == 0 || w->n_true == nd->count) { order_out[q] = order[q]; return; }
  const int base = scan[start];
  if (flag[q]) order_out[start + (scan[q] - base)] = order[q];    // scatter of the trues
  if (q >= start + w->n_true) {
    int src = q;
    while (flag[src]) src = start + (scan[src] - base);
    order_out[q] = order[src];
  }
}

// step 3c: children of every split segment; next-level segment id of every slot
__global__ void k_bvh_children(const SegWork* __restrict__ work, GpuNode* __restrict__ nodes, int first_node, int n_seg,
                               const int* __restrict__ split_rank, int next_first) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  const SegWork* w = work + s;
  if (!w->is_split) return;
  GpuNode* nd = nodes + first_node + s;
  const int l = next_first + 2 * split_rank[s];
  nd->left = l;
  GpuNode* a = nodes + l;
  GpuNode* b = nodes + l + 1;
  a->start = nd->start; a->count = w->mid - nd->start; a->left = -1;
  b->start = w->mid; b->count = nd->start + nd->count - w->mid; b->left = -1;
}
__global__ void k_bvh_split_flags(const SegWork* __restrict__ work, int n_seg, int* __restrict__ is_split) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_seg) is_split[s] = work[s].is_split;
}
__global__ void k_bvh_next_seg(int* __restrict__ seg_of, int n, const SegWork* __restrict__ work, const int* __restrict__ split_rank) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int seg = seg_of[q];
  if (seg < 0) return;
  const SegWork* w = work + seg;
  seg_of[q] = w->is_split ? 2 * split_rank[seg] + (q >= w->mid ? 1 : 0) : -1;
}
__global__ void k_bvh_init_work(SegWork* __restrict__ work, int n_seg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  for (int k = 0; k < 3; k++) { work[s].blo[k] = 0xffffffffu; work[s].clo[k] = 0xffffffffu; work[s].bhi[k] = 0u; work[s].chi[k] = 0u; }
}

// step 4: subtree sizes bottom-up (one launch per level, deepest first), pre-order index top-down
__global__ void k_bvh_sizes(GpuNode* __restrict__ nodes, int first, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  GpuNode* nd = nodes + first + i;
  nd->subtree = 1 + (nd->left >= 0 ? nodes[nd->left].subtree + nodes[nd->left + 1].subtree : 0);
}
__global__ void k_bvh_flat(GpuNode* __restrict__ nodes, int first, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const GpuNode* nd = nodes + first + i;
  if (nd->left >= 0) {
    nodes[nd->left].flat = nd->flat + 1;                               // left child = parent + 1 (bvh.h:192-200)
    nodes[nd->left + 1].flat = nd->flat + 1 + nodes[nd->left].subtree;
  }
}
__global__ void k_bvh_emit(const GpuNode* __restrict__ nodes, int total, BvhNodeHost* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const GpuNode* nd = nodes + i;
  BvhNodeHost o;
  for (int k = 0; k < 3; k++) { o.lo[k] = nd->lo[k]; o.hi[k] = nd->hi[k]; }
  o.start = nd->start;
  o.nPrims = nd->count;
  o.rightOffset = nd->left >= 0 ? nodes[nd->left + 1].flat - nd->flat : 0;
  out[nd->flat] = o;
}
__global__ void k_iota(int* __restrict__ a, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (v < 0) ? i : v;
}

struct ScanScratch {
  std::vector<int*> levels;
  std::vector<int> sizes;
};

// exclusive scan of `in` (n ints) into `out`; scratch levels are allocated once for the largest n
cudaError_t exclusive_scan(const int* in, int* out, int n, ScanScratch& sc, int depth, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int nb = (n + 1023) / 1024;
  int* sums = nullptr;
  if (nb > 1) {
    if ((int)sc.levels.size() <= depth) { sc.levels.resize(depth + 1, nullptr); sc.sizes.resize(depth + 1, 0); }
    if (sc.sizes[depth] < 2 * nb) {
      if (sc.levels[depth]) cudaFree(sc.levels[depth]);
      cudaError_t e = cudaMalloc(&sc.levels[depth], (size_t)2 * nb * sizeof(int));
      if (e != cudaSuccess) return e;
      sc.sizes[depth] = 2 * nb;
    }
    sums = sc.levels[depth];
  }
  k_scan_block<<<nb, 1024, 0, st>>>(in, out, n, sums);
  if (nb > 1) {
    int* sums_scanned = sums + nb;
    cudaError_t e = exclusive_scan(sums, sums_scanned, nb, sc, depth + 1, st);
    if (e != cudaSuccess) return e;
    k_scan_add<<<nb, 1024, 0, st>>>(out, n, sums_scanned);
  }
  return cudaGetLastError();
}

}  // namespace

static unsigned long long g_bvh_launches = 0;
unsigned long long bvh_kernel_launches() { return g_bvh_launches; }

static bool build_bvh_gpu(const float* boxes, const float* pos, int n, int leaf_size, std::vector<BvhNodeHost>* nodes_out,
                          std::vector<int>* perm, int device, std::string* err) {
  bool ok = true;
  float *d_boxes = nullptr, *d_pos = nullptr;
  int *d_order = nullptr, *d_order2 = nullptr, *d_seg = nullptr, *d_flag = nullptr, *d_scan = nullptr, *d_issplit = nullptr, *d_rank = nullptr;
  GpuNode* d_nodes = nullptr;
  SegWork* d_work = nullptr;
  BvhNodeHost* d_out = nullptr;
  ScanScratch scratch;
  std::vector<int> level_first, level_count;
  const int max_nodes = 2 * n;     // a binary tree over n >= 1 primitives with non-empty leaves has < 2n nodes
  const int T = 256;
  const int nbq = (n + T - 1) / T;
  cudaStream_t st = nullptr;
  int first = 0, n_seg = 1, total = 1;
  unsigned long long launches = 0;
  {
    BVH_CUDA(cudaSetDevice(device));
    BVH_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    BVH_CUDA(cudaMalloc(&d_boxes, (size_t)6 * n * sizeof(float)));
    BVH_CUDA(cudaMalloc(&d_pos, (size_t)3 * n * sizeof(float)));
    BVH_CUDA(cudaMalloc(&d_order, (size_t)n * sizeof(int)));
    BVH_CUDA(cudaMalloc(&d_order2, (size_t)n * sizeof(int)));
    BVH_CUDA(cudaMalloc(&d_seg, (size_t)n * sizeof(int)));
    BVH_CUDA(cudaMalloc(&d_flag, (size_t)n * sizeof(int)));
    BVH_CUDA(cudaMalloc(&d_scan, (size_t)n * sizeof(int)));
    BVH_CUDA(cudaMalloc(&d_issplit, (size_t)n * sizeof(int)));
    BVH_CUDA(cudaMalloc(&d_rank, (size_t)n * sizeof(int)));
    BVH_CUDA(cudaMalloc(&d_nodes, (size_t)max_nodes * sizeof(GpuNode)));
    BVH_CUDA(cudaMalloc(&d_work, (size_t)n * sizeof(SegWork)));
    BVH_CUDA(cudaMalloc(&d_out, (size_t)max_nodes * sizeof(BvhNodeHost)));
    BVH_CUDA(cudaMemcpyAsync(d_boxes, boxes, (size_t)6 * n * sizeof(float), cudaMemcpyHostToDevice, st));
    BVH_CUDA(cudaMemcpyAsync(d_pos, pos, (size_t)3 * n * sizeof(float), cudaMemcpyHostToDevice, st));
    k_iota<<<nbq, T, 0, st>>>(d_order, n, -1);
    k_iota<<<nbq, T, 0, st>>>(d_seg, n, 0);
    GpuNode root;
    memset(&root, 0, sizeof(root));
    root.start = 0; root.count = n; root.left = -1; root.flat = 0;
    BVH_CUDA(cudaMemcpyAsync(d_nodes, &root, sizeof(root), cudaMemcpyHostToDevice, st));
    launches += 2;
    while (n_seg > 0) {
      level_first.push_back(first);
      level_count.push_back(n_seg);
      const int nbs = (n_seg + T - 1) / T;
      k_bvh_init_work<<<nbs, T, 0, st>>>(d_work, n_seg);
      k_bvh_reduce<<<nbq, T, 0, st>>>(d_boxes, d_pos, d_order, d_seg, n, d_work);
      k_bvh_decide<<<nbs, T, 0, st>>>(d_work, d_nodes, first, n_seg, leaf_size);
      k_bvh_flags<<<nbq, T, 0, st>>>(d_pos, d_order, d_seg, n, d_work, d_flag);
      BVH_CUDA(exclusive_scan(d_flag, d_scan, n, scratch, 0, st));
      k_bvh_mid<<<nbs, T, 0, st>>>(d_work, d_nodes, first, n_seg, d_scan, d_flag, n);
      k_bvh_permute<<<nbq, T, 0, st>>>(d_order, d_order2, d_seg, n, d_work, d_nodes, first, d_flag, d_scan);
      std::swap(d_order, d_order2);
      k_bvh_split_flags<<<nbs, T, 0, st>>>(d_work, n_seg, d_issplit);
      BVH_CUDA(exclusive_scan(d_issplit, d_rank, n_seg, scratch, 8, st));
      int last_rank = 0, last_flag = 0;
      BVH_CUDA(cudaMemcpyAsync(&last_rank, d_rank + n_seg - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
      BVH_CUDA(cudaMemcpyAsync(&last_flag, d_issplit + n_seg - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
      BVH_CUDA(cudaStreamSynchronize(st));
      const int n_split = last_rank + last_flag;
      launches += 9;
      if (n_split == 0) break;
      if (total + 2 * n_split > max_nodes) { *err = "BVH node budget exceeded"; ok = false; goto done; }
      k_bvh_children<<<nbs, T, 0, st>>>(d_work, d_nodes, first, n_seg, d_rank, total);
      k_bvh_next_seg<<<nbq, T, 0, st>>>(d_seg, n, d_work, d_rank);
      launches += 2;
      first = total;
      n_seg = 2 * n_split;
      total += n_seg;
    }
    for (int l = (int)level_first.size() - 1; l >= 0; l--) {
      k_bvh_sizes<<<(level_count[l] + T - 1) / T, T, 0, st>>>(d_nodes, level_first[l], level_count[l]);
      launches++;
    }
    for (size_t l = 0; l < level_first.size(); l++) {
      k_bvh_flat<<<(level_count[l] + T - 1) / T, T, 0, st>>>(d_nodes, level_first[l], level_count[l]);
      launches++;
    }
    k_bvh_emit<<<(total + T - 1) / T, T, 0, st>>>(d_nodes, total, d_out);
    launches++;
    nodes_out->resize(total);
    perm->resize(n);
    BVH_CUDA(cudaMemcpyAsync(nodes_out->data(), d_out, (size_t)total * sizeof(BvhNodeHost), cudaMemcpyDeviceToHost, st));
    BVH_CUDA(cudaMemcpyAsync(perm->data(), d_order, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    BVH_CUDA(cudaStreamSynchronize(st));
    BVH_CUDA(cudaGetLastError());
  }
done:
  g_bvh_launches += launches;
  cudaFree(d_boxes); cudaFree(d_pos); cudaFree(d_order); cudaFree(d_order2); cudaFree(d_seg); cudaFree(d_flag); cudaFree(d_scan);
  cudaFree(d_issplit); cudaFree(d_rank); cudaFree(d_nodes); cudaFree(d_work); cudaFree(d_out);
  for (int* p : scratch.levels) if (p) cudaFree(p);
  if (st) cudaStreamDestroy(st);
  return ok;
}

// builder policy: 0 = auto (GPU from kGpuBuildMin primitives on), 1 = host, 2 = GPU
static int g_builder_mode = 0;
void set_bvh_builder_mode(int mode) { g_builder_mode = mode; }
constexpr int kGpuBuildMin = 2048;

bool build_bvh(const float* boxes, const float* pos, int n, int leaf_size, std::vector<BvhNodeHost>* nodes, std::vector<int>* perm,
               int device, std::string* err) {
  if (n == 0) { nodes->clear(); perm->clear(); return true; }
  const bool gpu = (g_builder_mode == 2) || (g_builder_mode == 0 && n >= kGpuBuildMin);
  if (!gpu) return build_bvh_host(boxes, pos, n, leaf_size, nodes, perm);
  if (device < 0) { *err = "GPU BVH build requested but no CUDA device is available (no CPU fallback for >= 2048 primitives)"; return false; }
  return build_bvh_gpu(boxes, pos, n, leaf_size, nodes, perm, device, err);
}

}  // namespace ipt
