// kernels.h -- host-callable launchers implemented in paths.cu / trace.cu / wavefront.cu / bvh_build.cu
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "device_scene.h"

namespace ipt {

enum { MODE_RENDER_ID = 0, MODE_GRAPH_ID = 1, MODE_GRAD_ID = 2 };

size_t paths_smem_bytes(const DevScene& sc, int acc_floats);
int paths_blocks_per_sm(int mode, const DevScene& sc, size_t smem);
int paths_block_threads();   // threads per block of k_paths
cudaError_t launch_paths(int mode, const DevScene& sc, const DevRender& rp, const DevOut& out, int grid, size_t smem, cudaStream_t st);
// wavefront forward pass for BVH scenes: batches of <= cap paths, state in `workspace`; adds into `partial` exactly
// what k_paths<RENDER> would write.  host_live: one pinned int (used only with unbounded depth).
size_t wavefront_workspace_bytes(int cap);
// adjoint replay through the wavefront (BVH scenes): `record` holds wavefront_record_bytes(cap, vdepth, vf4) bytes (vf4 = 2
// for plain scenes, 4 otherwise); every batch is replayed with vertex recording and swept backwards into grad[3 nT] (doubles,
// atomics), weighted by dldpx (shard-local pixels).  With `adj` the pass writes no partial sums.
struct WavefrontAdjoint {
  void* record;
  int vdepth;
  const float* dldpx;
  double* grad;
};
size_t wavefront_record_bytes(int cap, int vdepth, int vf4);
cudaError_t launch_wavefront(const DevScene& sc, const DevRender& rp, void* workspace, int cap, float* partial, unsigned long long* counters,
                             int sm_count, int* host_live, unsigned long long* n_launches, cudaStream_t st, const WavefrontAdjoint* adj = nullptr);
// one bounce of the wavefront integrator (trace.cu): closest hits of the rays [0, n) u [cap, cap + n), n = *n_live
cudaError_t launch_trace_bounce(const DevScene& sc, const float* rays, const int* n_live, int cap, int* tri, float* t, unsigned long long* counter,
                                int sm_count, unsigned long long* n_launches, cudaStream_t st, bool seeded);
cudaError_t launch_trace_rays(const DevScene& sc, const float* rays, int n, int* tri, float* t, unsigned long long* counter, int sm_count, cudaStream_t st);
cudaError_t launch_trace_rays_stats(const DevScene& sc, const float* rays, int n, int* tri, float* t, unsigned long long* counter,
                                    unsigned long long* stats, int sm_count, cudaStream_t st);
cudaError_t launch_copy16(const float4* src, float4* dst, size_t n, int sm_count, cudaStream_t st);
cudaError_t launch_resolve(const float* partial, int n_pix, int groups, float* img, unsigned char* img8, cudaStream_t st);
cudaError_t launch_loss_dldpx(const float* img, const unsigned char* target, int row_begin, int row_step, int width, int n_pix, double inv_n, float inv_spp,
                              float* dldpx, double* loss, cudaStream_t st);
cudaError_t launch_reduce_blocks(const float* block_acc, int n_blocks, int stride, double* out, int ints_from, int ints_per, cudaStream_t st);
cudaError_t launch_set_kd(float4* mat, const float* kd, int nT, cudaStream_t st);
cudaError_t launch_graph_compress(const double* raw, int nT, float* data, cudaStream_t st);
cudaError_t launch_graph_finalize(const double* acc, const float4* mat, int nT, int n_edges, double* raw, cudaStream_t st);

// BVH::build (bvh.h:109-205) over generic primitives.  nodes: 9 values per node
// (lo[3], hi[3] as floats; start, nPrims, rightOffset as ints); perm: primitive order.
struct BvhNodeHost {
  float lo[3], hi[3];
  int start, nPrims, rightOffset;
};
void set_bvh_builder_mode(int mode);   // 0 auto, 1 host, 2 GPU
unsigned long long bvh_kernel_launches();
bool build_bvh(const float* boxes, const float* pos, int n, int leaf_size, std::vector<BvhNodeHost>* nodes, std::vector<int>* perm,
               int device, std::string* err);

}  // namespace ipt
