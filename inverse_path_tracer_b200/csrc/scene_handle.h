// scene_handle.h -- the scene handle behind the C ABI and the helpers api.cu shares with inverse.cu.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <memory>
#include <string>

#include "../../include/ipt_b200.h"
#include "device_scene.h"
#include "kernels.h"
#include "scene_host.h"

namespace ipt {

int fail(const std::string& msg);   // sets the thread's last error, returns -1
extern std::atomic<unsigned long long> g_launches;

#define IPT_CUDA(call)                                                                                        \
  do {                                                                                                        \
    cudaError_t e__ = (call);                                                                                 \
    if (e__ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e__));                 \
  } while (0)

struct Scratch {
  void* ptr = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&ptr, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() { if (ptr) cudaFree(ptr); ptr = nullptr; bytes = 0; }
};

struct SceneHandle {
  HostScene host;
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  float4* geom = nullptr;
  float4* geom_edges = nullptr;   // BVH scenes: edge planes of the leaf slots (inside the geom allocation)
  float4* bf_stream = nullptr;   // record stream of the brute-force set (device_scene.h)
  int n_groups = 0;
  int n_stream = 0;
  float4* mat = nullptr;
  float4* light = nullptr;
  float4* vtx = nullptr;
  float4* bvh_nodes = nullptr;
  float4* bvh_wide = nullptr;
  float4* bvh_q = nullptr;
  float bvh_qoff[3] = {0, 0, 0}, bvh_qscale[3] = {1, 1, 1};
  int bvh_root_desc = 0;
  int bvh_depth = 1;
  int* bvh_prims = nullptr;
  int n_nodes = 0;
  int n_big = 0;
  unsigned char* big_slot = nullptr;
  int* big_tri = nullptr;
  bool use_bvh = false;
  bool host_only = false;
  bool kd_dirty = false;      // the device albedos were changed on the device (iptSetMaterialsDevice, iptInverseStep): getMaterials reads them back
  // knobs copied from the loading thread's settings (iptSetTraversalMode / iptSetBvhLargeTriangleLimit / iptSetIntegrator)
  int traversal_mode = 0, big_max = 128, integrator = 0, wf_slots = 1 << 24;
  // ring of CUDA event pairs bracketing the last kEventRing k_paths launches (roofline timing without
  // synchronising inside a timed loop)
  static constexpr int kEventRing = 64;
  cudaEvent_t ev_start[kEventRing] = {}, ev_stop[kEventRing] = {};
  unsigned long long ev_count = 0;   // loaded with iptSetDevice(-2): tables only, no device arrays
  int* host_live = nullptr;          // pinned word for the wavefront integrator's live count
  Scratch wavefront;
  Scratch partial, counters, cnt_out, trace_counter, block_acc, vstacks, acc, dldpx, img, img8, target, raw, lossgrad, rays, hits_tri, hits_t;
  // One call at a time per scene: every entry point holds `mutex` while it enqueues (the scratch buffers, the work
  // counter and the event ring are per handle), and calls enqueued on DIFFERENT streams are chained through `busy`
  // (recorded after each call, waited on by the next), so two calls in flight on one scene cannot share scratch.
  std::recursive_mutex mutex;
  // chunked launches (api.cu; measured slower, off unless IPT_CHUNKS asks): internal streams and fork / join events
  static constexpr int kMaxChunks = 4;
  cudaStream_t cstream[kMaxChunks] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxChunks] = {};
  int chunks = 1;
  cudaEvent_t busy = nullptr;
  bool busy_recorded = false;
  cudaStream_t busy_stream = nullptr;

  DevScene dev() const {
    DevScene d;
    d.geom = geom; d.geom_edges = geom_edges; d.stream = bf_stream; d.nGroups = n_groups; d.nStream = n_stream; d.mat = mat; d.light = light; d.vtx = vtx; d.bvh_nodes = bvh_nodes; d.bvh_wide = bvh_wide; d.bvh_q = bvh_q; for (int k = 0; k < 3; k++) { d.bvh_qoff[k] = bvh_qoff[k]; d.bvh_qscale[k] = bvh_qscale[k]; } d.bvh_root_desc = bvh_root_desc; d.bvh_depth = bvh_depth; d.bvh_prims = bvh_prims;
    d.nT = (int)host.tris.size(); d.nE = (int)host.emissives.size(); d.nNodes = n_nodes; d.nBig = n_big; d.big_slot = big_slot; d.big_tri = big_tri; d.use_bvh = use_bvh ? 1 : 0; d.plain = (!host.any_vn && !host.any_specular) ? 1 : 0;
    memcpy(d.cam, host.cam, sizeof(d.cam));
    return d;
  }
  void destroy() {
    if (host_only) return;
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    cudaFree(geom); cudaFree(bf_stream); cudaFree(mat); cudaFree(light); cudaFree(vtx); cudaFree(bvh_nodes); cudaFree(bvh_wide); cudaFree(bvh_q); cudaFree(big_slot); cudaFree(big_tri); cudaFree(bvh_prims);
    Scratch* all[] = {&wavefront, &partial, &counters, &cnt_out, &trace_counter, &block_acc, &vstacks, &acc, &dldpx, &img, &img8, &target, &raw, &lossgrad, &rays, &hits_tri, &hits_t};
    for (Scratch* s : all) s->release();
    if (host_live) cudaFreeHost(host_live);
    if (busy) cudaEventDestroy(busy);
    if (ev_fork) cudaEventDestroy(ev_fork);
    for (int i = 0; i < kMaxChunks; i++) { if (cstream[i]) { cudaStreamSynchronize(cstream[i]); cudaStreamDestroy(cstream[i]); } if (ev_join[i]) cudaEventDestroy(ev_join[i]); }
    if (stream) cudaStreamDestroy(stream);
    for (int i = 0; i < kEventRing; i++) { if (ev_start[i]) cudaEventDestroy(ev_start[i]); if (ev_stop[i]) cudaEventDestroy(ev_stop[i]); }
  }
};

// RAII guard of a device-pointer entry point: takes the handle's mutex, makes `st` wait for the scene's previous call
// (if that ran on another stream) and records `busy` on `st` when the enqueue is done.  Inside a stream capture the
// events are skipped (a capture may not wait on outside work); the captured graph then owns the ordering.
struct CallGuard {
  SceneHandle* h;
  cudaStream_t st;
  std::unique_lock<std::recursive_mutex> lk;
  bool capturing = false;
  CallGuard(SceneHandle* h_, cudaStream_t st_) : h(h_), st(st_), lk(h_->mutex) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) capturing = true;
    if (!capturing && h->busy && h->busy_recorded && h->busy_stream != st) cudaStreamWaitEvent(st, h->busy, 0);
  }
  ~CallGuard() {
    if (!capturing && h->busy) {
      if (cudaEventRecord(h->busy, st) == cudaSuccess) { h->busy_recorded = true; h->busy_stream = st; }
    }
  }
};

SceneHandle* checked(void* p);          // registry lookup; null + error for a stale / foreign pointer
SceneHandle* checked_device(void* p);   // ... and the scene must live on a GPU

struct Plan {
  DevRender rp;
  int rows;
};
int make_plan(const SceneHandle* h, const IptRenderParams* p, Plan* plan);

// chunked launches (api.cu)
struct Fork {   // `st` forks into k internal streams (begin) that join it again (end); k == 1: everything stays on st
  SceneHandle* h;
  cudaStream_t st;
  int k;
  Fork(SceneHandle* h, cudaStream_t st, int k);
  cudaError_t begin();
  cudaError_t end();
  cudaStream_t stream(int c) const { return k <= 1 ? st : h->cstream[c]; }
};
struct GradLaunch {
  size_t smem, vstack_bytes;
  int grid, acc_mode, vdepth;
};
int pick_chunks(const SceneHandle* h, const Plan& plan);
void chunk_render(const Plan& plan, int c, int k, DevRender* rp_c, int* lr0, int* lr1);
int prepare_forward(SceneHandle* h, const Plan& plan);
int forward_chunk(SceneHandle* h, const Plan& plan, const DevScene& sc, int c, int k, unsigned long long* counters, cudaStream_t st_c);
int prepare_grad(SceneHandle* h, const Plan& plan, int k, GradLaunch* g);
int grad_chunk(SceneHandle* h, const Plan& plan, const DevScene& sc, const GradLaunch& g, int c, int k, const unsigned char* target_dev,
               double* lossgrad_dev, cudaStream_t st_c);
// forward pass of the shard into h->partial (+ toneMap into img_dev / img8_dev unless resolve == false)
int launch_forward(SceneHandle* h, const Plan& plan, float* img_dev, unsigned char* img8_dev, unsigned long long* counters_dev, cudaStream_t st,
                   bool resolve = true);
// path adjoint of the shard: adds dLoss/dKd (3 nT doubles) to lossgrad_dev + 1; needs h->dldpx filled for this plan
int launch_grad_pass(SceneHandle* h, const Plan& plan, const unsigned char* target_dev, double* lossgrad_dev, cudaStream_t st);

}  // namespace ipt
