// NOT PART OF THE LIBRARY.  Kept as the record of an experiment (profiles/r02_small_launches.txt): built as csrc/lpt.cu with a
// DevOut::item_order indirection in k_paths' work claim, measured on a B200 (sessions r02_s36 / r02_s37), found not to help, removed.
// It compiles only next to csrc/device_scene.h and csrc/kernels.h (with launch_item_order / item_order_bytes declared there).
//
// lpt.cu -- longest paths first: the order in which k_paths hands out the work items of a small launch
//
// A persistent launch ends in a tail as long as its longest path.  With Russian roulette and no depth limit (the reference,
// scene.h:11 / path_trace.cu:131-137) path lengths are geometric -- mean 10 bounces, the longest of 2 M paths about 140 --
// and for a small launch (a row shard of a 512 x 512 frame on one of eight GPUs: ~18 paths per lane) that one path, started
// at a random time, is a third of the launch.  But nothing about a path's length is unknown before it is traced: whether
// bounce k continues is decided by one word of the Philox block keyed (pixel, sample, k, IPT_STREAM_BOUNCE) (paths.cu [A]:
// u16(rn.x & 0xffff) < p_rr), and geometry can only end a path earlier.  So a pre-pass bounds the length of every path of the
// launch, marks the work items that hold a long one ("heavy": >= kHeavyBounces) and builds a permutation of the item queue
// that spreads the heavy items over its beginning, one per claim -- longest-processing-time-first scheduling.  k_paths reads its queue position through
// that permutation (DevOut::item_order); results do not depend on the order items are processed in (every item has its own
// partial sum), so images, counters and gradients are bit-identical with and without it (tests/test_gpu_parity.py).
//
// The pre-pass repeats the ~10 Philox calls of a path (about 4 % of the work of the render kernel), so api.cu only runs it
// for launches with fewer than kLptMaxPathsPerLane paths per resident lane, where the tail is worth more than that.
#include <algorithm>

#include "device_scene.h"
#include "kernels.h"

namespace ipt {

constexpr int kHeavyBounces = 48;     // a path at least this long makes its work item "heavy" (p_rr = 0.9: 0.6 % of the paths)
constexpr int kLptBlock = 256;

struct LptBuffers {
  unsigned* order;      // [n_items] queue position -> work item
  unsigned char* flag;  // [n_items] 1 = heavy
  unsigned* heavy;      // [cap] heavy items in the order found
  unsigned* a;          // [cap] heavy items that sit behind the head of the queue
  unsigned* b;          // [cap] light items that sit in the head of the queue
  unsigned* count;      // [4] number of heavy items, |a|, |b|, -
  unsigned cap;
};

// upper bound of the number of radiance() calls of the path (pixel, sample): path_trace.cu:131-137 as implemented in paths.cu [A]
__device__ __forceinline__ int path_length_bound(const DevRender& rp, uint32_t pixel, uint32_t sample, int stop_at) {
  int b = 0;
  for (;; b++) {
    const bool depth_ok = (rp.max_depth <= 0) || (b + 1 < rp.max_depth);
    if (!depth_ok || b + 1 >= stop_at) break;
    const u4 rn = philox4x32_10(pixel, sample, (uint32_t)b, IPT_STREAM_BOUNCE, rp.keys);
    if (!(u16(rn.x & 0xffffu) < rp.p_rr)) break;
  }
  return b + 1;
}

// identity order + heavy flags + list of the heavy items
__global__ void __launch_bounds__(kLptBlock) k_lpt_classify(DevRender rp, unsigned n_items, LptBuffers buf) {
  for (unsigned item = blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += gridDim.x * blockDim.x) {
    // decode_item of k_paths (paths.cu): item -> (shard-local pixel, sample group) -> global pixel
    const unsigned pix_local = item / (unsigned)rp.groups;
    const int g = (int)(item - pix_local * (unsigned)rp.groups);
    const int lr = (int)(pix_local / (unsigned)rp.width);
    const int c = (int)pix_local - lr * rp.width;
    const int r = rp.row_begin + lr * rp.row_step;
    const uint32_t pixel = (uint32_t)(r * rp.width + c);
    const int s0 = g * rp.group_samples, s1 = min(rp.spp, s0 + rp.group_samples);
    bool is_heavy = false;
    for (int s = s0; s < s1 && !is_heavy; s++) is_heavy = path_length_bound(rp, pixel, (uint32_t)s, kHeavyBounces) >= kHeavyBounces;
    buf.order[item] = item;
    buf.flag[item] = is_heavy ? 1 : 0;
    if (is_heavy) {
      const unsigned pos = atomicAdd(buf.count + 0, 1u);
      if (pos < buf.cap) buf.heavy[pos] = item;
    }
  }
}

// Heavy item k goes to queue position k * kLptStride: k_paths claims runs of up to 8 consecutive positions per lane (guided
// self-scheduling), so with a stride of 8 every lane's first run holds at most one heavy item and the long paths all start at
// the beginning of the launch on different lanes (heavy items packed at the head would hand two or more of them to the same
// lane, which then becomes the tail: measured, 0.58 -> 0.78 ms on a 1/8 row shard of the 512 x 512 x 64 spp frame).
// A heavy item that already sits on a target position stays; every other one swaps with the light item of a free target.
constexpr unsigned kLptStride = 8;
__global__ void __launch_bounds__(kLptBlock) k_lpt_split(unsigned n_items, LptBuffers buf) {
  const unsigned nH = buf.count[0];
  if (nH > buf.cap || (unsigned long long)nH * kLptStride > n_items) return;   // too many: the identity order stands
  for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < nH; j += gridDim.x * blockDim.x) {
    const unsigned h = buf.heavy[j];
    if (h % kLptStride != 0u || h / kLptStride >= nH) buf.a[atomicAdd(buf.count + 1, 1u)] = h;
    const unsigned t = j * kLptStride;
    if (!buf.flag[t]) buf.b[atomicAdd(buf.count + 2, 1u)] = t;
  }
}

__global__ void __launch_bounds__(kLptBlock) k_lpt_swap(unsigned n_items, LptBuffers buf) {
  const unsigned nH = buf.count[0];
  if (nH > buf.cap || (unsigned long long)nH * kLptStride > n_items) return;
  const unsigned n = min(buf.count[1], buf.count[2]);   // equal by construction
  for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const unsigned hi = buf.a[k], lo = buf.b[k];
    buf.order[lo] = hi;
    buf.order[hi] = lo;
  }
}

size_t item_order_bytes(unsigned n_items) {
  const size_t cap = std::max<size_t>(n_items / 8, 1024);
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  return al((size_t)n_items * 4) + al(n_items) + 3 * al(cap * 4) + 256;
}

// builds the permutation in `scratch` (item_order_bytes(n_items) bytes) on `st`; returns the device pointer of order[n_items]
cudaError_t launch_item_order(const DevRender& rp, unsigned n_items, void* scratch, int sm_count, const unsigned** order, cudaStream_t st) {
  const size_t cap = std::max<size_t>(n_items / 8, 1024);
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  char* w = (char*)scratch;
  LptBuffers buf;
  buf.order = (unsigned*)w; w += al((size_t)n_items * 4);
  buf.flag = (unsigned char*)w; w += al(n_items);
  buf.heavy = (unsigned*)w; w += al(cap * 4);
  buf.a = (unsigned*)w; w += al(cap * 4);
  buf.b = (unsigned*)w; w += al(cap * 4);
  buf.count = (unsigned*)w;
  buf.cap = (unsigned)cap;
  cudaError_t e = cudaMemsetAsync(buf.count, 0, 4 * sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  const int grid = (int)std::min<size_t>(((size_t)n_items + kLptBlock - 1) / kLptBlock, (size_t)sm_count * 8);
  k_lpt_classify<<<grid, kLptBlock, 0, st>>>(rp, n_items, buf);
  k_lpt_split<<<sm_count, kLptBlock, 0, st>>>(n_items, buf);
  k_lpt_swap<<<sm_count, kLptBlock, 0, st>>>(n_items, buf);
  *order = buf.order;
  return cudaGetLastError();
}

}  // namespace ipt
