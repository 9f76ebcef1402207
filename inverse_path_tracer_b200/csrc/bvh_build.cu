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
// Round 1 runs steps 1-4 on the host (single thread per segment) -- the GPU port of the same four
// steps is the next milestone; the output contract is already the final one.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "kernels.h"

namespace ipt {

namespace {

struct Segment {
  int start, end;
  int parent;      // index into the segment list of the previous level (-1 for the root)
  int is_right;    // 0 = left child, 1 = right child
};

struct BuildNode {
  float lo[3], hi[3];
  int start, count;
  int left, right;   // build-node ids, -1 for leaves
  int subtree;       // nodes in this subtree
  int flat;          // DFS pre-order index
};

}  // namespace

bool build_bvh(const float* boxes, const float* pos, int n, int leaf_size, std::vector<BvhNodeHost>* nodes, std::vector<int>* perm,
               int device, std::string* err) {
  (void)device;
  (void)err;
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
          lo[k] = fminf(lo[k], boxes[6 * pr + k]); hi[k] = fmaxf(hi[k], boxes[6 * pr + 3 + k]);
          clo[k] = fminf(clo[k], pos[3 * pr + k]); chi[k] = fmaxf(chi[k], pos[3 * pr + k]);
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

}  // namespace ipt
