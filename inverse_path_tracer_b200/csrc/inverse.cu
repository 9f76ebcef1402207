// inverse.cu -- the inverse loop inside the library: one call = one optimisation iteration on this GPU's row shard.
//
//   iptInverseStep:  k_paths<RENDER>  ->  k_resolve_loss  ->  k_paths<GRAD>  ->  k_reduce_blocks  ->  k_inverse_update
//
// The reference has no optimiser for the albedos (its material recovery is the GCN of ipt.py); this is the loop of
// BASELINE config 3 / 4 -- forward render of the shard at `spp` samples, loss against the target image, path adjoint,
// sum of [loss, dLoss/dKd] over the GPUs that share the image, Adam step on the nT x 3 albedos, new albedos into the
// scene -- issued from C++ in one call instead of ~12 Python / torch / NCCL calls per iteration.  The exchange step is
// fused into the update kernel (k_inverse_update): every rank stores its 1 + 3 nT partial sums straight into every
// peer's mailbox over NVLink (peer pointers from CUDA IPC), raises a flag there, waits for the peers' flags in its own
// mailbox, adds the W partial vectors in rank order -- so every rank computes bit-identical sums and therefore
// bit-identical albedos, with no second broadcast -- and applies Adam.  Messages are 728 bytes (scene 0): the cost is
// one NVLink store latency instead of an NCCL launch, and no host thread is involved.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "scene_handle.h"

namespace ipt {

// toneMap + loss in one pass over the shard (k_resolve + k_loss_dldpx of paths.cu, same arithmetic):
// pixel = sum of group partials in group order; y = L / (1 + L); loss += (y - target / 255)^2 / (3 W H);
// dldpx = dLoss/dL / spp
__global__ void k_resolve_loss(const float* __restrict__ partial, int n_pix, int groups, const unsigned char* __restrict__ target, int row_begin,
                               int row_step, int width, double inv_n, float inv_spp, float* __restrict__ img, float* __restrict__ dldpx,
                               double* __restrict__ loss) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  if (i < n_pix) {
    const float* p = partial + (size_t)i * groups * 3;
    float a[3] = {p[0], p[1], p[2]};
    for (int g = 1; g < groups; g++) { a[0] += p[3 * g]; a[1] += p[3 * g + 1]; a[2] += p[3 * g + 2]; }
    const int lr = i / width;
    const size_t gpix = (size_t)(row_begin + lr * row_step) * width + (i - lr * width);
    for (int c = 0; c < 3; c++) {
      if (img) img[3 * (size_t)i + c] = a[c];
      const double L = (double)a[c];
      const double y = L / (1.0 + L);
      const double ys = (double)target[3 * gpix + c] / 255.0;
      acc += (y - ys) * (y - ys) * inv_n;
      dldpx[3 * (size_t)i + c] = (float)(2.0 * (y - ys) * inv_n / ((1.0 + L) * (1.0 + L)) * (double)inv_spp);
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss, acc);
}

// Mailbox of one rank (device memory of that rank, opened by every peer through CUDA IPC):
//   data[2][W][n]  doubles: buffer b = step parity, slot r = sending rank, n = 1 + 3 nT
//   flag[2][W]     unsigned long long: step number whose data slot [b][r] holds
// Two buffers suffice: a rank cannot run more than one step ahead of its slowest peer, because finishing step t needs
// that peer's flag for step t, which the peer raises only after it has consumed step t - 1.
struct InverseState {
  float* kd;        // [3 nT] current albedos
  float* m;         // [3 nT] Adam first moment
  float* v;         // [3 nT] Adam second moment
  double* lossgrad; // [n] this rank's partial sums, then (after the update) the global sums of the last step
  double* history;  // [cap] global loss per step (ring)
  unsigned long long* step;   // steps applied so far
  int* error;       // 1: a peer did not arrive within the timeout
};

struct InversePeers {
  double* box[8];   // mailbox base of every rank (own rank: local pointer)
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// one block.  n = 1 + 3 nT <= blockDim.x * k is handled by a strided loop.
__global__ void __launch_bounds__(256) k_inverse_update(InverseState st, InversePeers peers, int rank, int world, int nT, float4* __restrict__ mat,
                                                        float lr, float beta1, float beta2, float eps, float kd_min, float kd_max, int history_cap,
                                                        unsigned long long timeout_ns) {
  const int n = 1 + 3 * nT;
  const unsigned long long t = *st.step + 1ull;   // this step's number (1-based)
  __shared__ int s_fail;
  if (threadIdx.x == 0) s_fail = 0;
  __syncthreads();
  if (world > 1) {
    const int b = (int)(t & 1ull);
    const size_t data_per_buf = (size_t)world * n;
    // 1. my partial sums into slot [b][rank] of every mailbox (own included), straight over NVLink
    for (int r = 0; r < world; r++) {
      double* dst = peers.box[r] + (size_t)b * data_per_buf + (size_t)rank * n;
      for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = st.lossgrad[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag in every mailbox
    if ((int)threadIdx.x < world) {
      unsigned long long* flags = reinterpret_cast<unsigned long long*>(peers.box[threadIdx.x] + 2 * data_per_buf);
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flags + (size_t)b * world + rank), "l"(t) : "memory");
    }
    // 3. wait for every rank's flag in my mailbox
    if ((int)threadIdx.x < world) {
      const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peers.box[rank] + 2 * data_per_buf) + (size_t)b * world + threadIdx.x;
      const unsigned long long t0 = globaltimer_ns();
      while (ld_volatile_u64(mine) < t) {
        if (globaltimer_ns() - t0 > timeout_ns) { s_fail = 1; break; }
        __nanosleep(200);
      }
    }
    __threadfence_system();
    __syncthreads();
    if (s_fail) {
      if (threadIdx.x == 0) *st.error = 1;
      return;
    }
    // 4. global sums in rank order: identical on every rank
    const double* box = peers.box[rank] + (size_t)b * data_per_buf;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      double a = 0.0;
      for (int r = 0; r < world; r++) a += ld_volatile_f64(box + (size_t)r * n + i);
      st.lossgrad[i] = a;
    }
    __syncthreads();
  }
  // 5. Adam (torch.optim.Adam's arithmetic in float32, bias corrections in double) + clamp + albedos into the scene
  const double bc1 = 1.0 - pow((double)beta1, (double)t), bc2 = 1.0 - pow((double)beta2, (double)t);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const float w1 = (float)(1.0 - (double)beta1), w2 = (float)(1.0 - (double)beta2);
  for (int i = threadIdx.x; i < 3 * nT; i += blockDim.x) {
    const float g = (float)st.lossgrad[1 + i];
    float m = st.m[i], v = st.v[i], kd = st.kd[i];
    m = m + w1 * (g - m);
    v = v * beta2;
    v = v + (w2 * g) * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    kd = kd - step_size * (m / denom);
    kd = fminf(fmaxf(kd, kd_min), kd_max);
    st.m[i] = m; st.v[i] = v; st.kd[i] = kd;
  }
  __syncthreads();
  for (int tri = threadIdx.x; tri < nT; tri += blockDim.x) {
    const float r = st.kd[3 * tri], g = st.kd[3 * tri + 1], b = st.kd[3 * tri + 2];
    float4 m0 = mat[tri];
    m0.x = r; m0.y = g; m0.z = b;
    mat[tri] = m0;                                                                        // Kd
    mat[2 * nT + tri] = make_float4(r / IPT_PI_F, g / IPT_PI_F, b / IPT_PI_F, 0.0f);      // Kd / pi (k_set_kd)
  }
  if (threadIdx.x == 0) {
    st.history[(t - 1ull) % (unsigned long long)history_cap] = st.lossgrad[0];
    *st.step = t;
  }
}

__global__ void k_inverse_seed_kd(InverseState st, const float* __restrict__ kd_in, float4* __restrict__ mat, int nT, int reset_moments) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * nT) {
    st.kd[i] = kd_in[i];
    if (reset_moments) { st.m[i] = 0.0f; st.v[i] = 0.0f; }
  }
  if (i < nT) {
    const float r = kd_in[3 * i], g = kd_in[3 * i + 1], b = kd_in[3 * i + 2];
    float4 m0 = mat[i];
    m0.x = r; m0.y = g; m0.z = b;
    mat[i] = m0;
    mat[2 * nT + i] = make_float4(r / IPT_PI_F, g / IPT_PI_F, b / IPT_PI_F, 0.0f);
  }
  if (i == 0 && reset_moments) *st.step = 0ull;
}

constexpr int kHistoryCap = 4096;
constexpr int kMaxWorld = 8;

struct InverseHandle {
  SceneHandle* scene = nullptr;
  void* scene_key = nullptr;
  IptRenderParams params;
  IptInverseConfig cfg;
  int nT = 0, n = 0;
  const unsigned char* target_dev = nullptr;   // caller-owned, full image
  void* target_own = nullptr;                  // library-owned copy (iptInverseCreate with a host target)
  InverseState st;
  void* state_mem = nullptr;
  double* mailbox = nullptr;                   // local mailbox (world > 1)
  size_t mailbox_bytes = 0;
  InversePeers peers;
  void* opened[kMaxWorld] = {};                // peer mappings opened through CUDA IPC (closed on destroy)
  bool connected = false;
  unsigned long long steps_issued = 0;
};

static size_t mailbox_size(int world, int n) { return (size_t)2 * world * n * sizeof(double) + (size_t)2 * world * sizeof(unsigned long long); }

static int enqueue_step(InverseHandle* o, unsigned long long seed) {
  SceneHandle* h = o->scene;
  cudaStream_t st = h->stream;
  IptRenderParams q = o->params;
  q.seed_lo = (uint32_t)seed;
  q.seed_hi = (uint32_t)(seed >> 32);
  if (q.max_depth <= 0 || q.max_depth > kGradMaxDepth) q.max_depth = kGradMaxDepth;
  Plan plan;
  if (make_plan(h, &q, &plan) != 0) return -1;
  if (o->params.group_samples < 0) {
    // The size rule looks at the whole image so that shards of a RENDER agree bit for bit; here only sums over the
    // shard matter (the exchange step re-associates them anyway), so a small shard -- one of eight ranks has ~20 paths
    // per resident lane -- takes groups small enough for >= 16 work items per lane: the launch must not end in a tail
    const long long lanes = (long long)h->sm_count * 5 * paths_block_threads();
    int gs = plan.rp.group_samples;
    while (gs > 1 && (long long)plan.rp.n_pix * plan.rp.spp / gs < 16 * lanes) gs >>= 1;
    if (gs != plan.rp.group_samples) {
      q.group_samples = gs;
      if (make_plan(h, &q, &plan) != 0) return -1;
    }
  }
  const DevRender& rp = plan.rp;
  const DevScene sc = h->dev();
  const size_t n3 = (size_t)rp.n_pix * 3;
  IPT_CUDA(h->dldpx.reserve(n3 * sizeof(float)));
  if (sc.use_bvh && h->integrator != 1) {
    // BVH scenes: forward pass and adjoint replay both run through the wavefront integrator (api.cu)
    IPT_CUDA(cudaMemsetAsync(o->st.lossgrad, 0, (size_t)o->n * sizeof(double), st));
    if (launch_forward(h, plan, nullptr, nullptr, nullptr, st, /*resolve=*/false) != 0) return -1;
    const double inv_n_w = 1.0 / (3.0 * (double)rp.width * (double)rp.height);
    k_resolve_loss<<<(rp.n_pix + 255) / 256, 256, 0, st>>>((const float*)h->partial.ptr, rp.n_pix, rp.groups, o->target_dev, rp.row_begin, rp.row_step, rp.width,
                                                           inv_n_w, 1.0f / (float)rp.spp, nullptr, (float*)h->dldpx.ptr, o->st.lossgrad);
    IPT_CUDA(cudaGetLastError());
    g_launches += 1;
    if (launch_grad_pass(h, plan, o->target_dev, o->st.lossgrad, st) != 0) return -1;
    k_inverse_update<<<1, 256, 0, st>>>(o->st, o->peers, o->cfg.rank, o->cfg.world, o->nT, h->mat, o->cfg.lr, o->cfg.beta1, o->cfg.beta2, o->cfg.eps,
                                       o->cfg.kd_min, o->cfg.kd_max, kHistoryCap, (unsigned long long)(o->cfg.timeout_ms > 0 ? o->cfg.timeout_ms : 10000) * 1000000ull);
    IPT_CUDA(cudaGetLastError());
    g_launches += 1;
    o->steps_issued++;
    return 0;
  }
  const int k = pick_chunks(h, plan);
  GradLaunch g;
  if (prepare_forward(h, plan) != 0 || prepare_grad(h, plan, k, &g) != 0) return -1;
  unsigned long long* ctr = (unsigned long long*)h->counters.ptr;
  IPT_CUDA(cudaMemsetAsync(o->st.lossgrad, 0, (size_t)o->n * sizeof(double), st));
  const double inv_n = 1.0 / (3.0 * (double)rp.width * (double)rp.height);
  // one chain per chunk of rows -- forward, toneMap + loss, adjoint -- on its own stream: the adjoint of the first rows
  // fills the SMs that the forward pass of the last rows leaves idle in its tail (api.cu, "chunked launches")
  Fork f(h, st, k);
  IPT_CUDA(cudaMemsetAsync(ctr, 0, 16 * sizeof(unsigned long long), st));
  const int slot = (int)(h->ev_count % SceneHandle::kEventRing);   // iptPathKernelTimes: forward + loss + adjoint of the iteration as one interval
  IPT_CUDA(cudaEventRecord(h->ev_start[slot], st));
  IPT_CUDA(f.begin());
  for (int c = 0; c < k; c++) {
    cudaStream_t sc_st = f.stream(c);
    DevRender rc;
    int lr0, lr1;
    chunk_render(plan, c, k, &rc, &lr0, &lr1);
    if (lr1 <= lr0) continue;
    if (forward_chunk(h, plan, sc, c, k, ctr + 8, sc_st) != 0) return -1;   // ray counters of the forward pass go to ctr[8..10], the adjoint's to ctr[0..2]
    const size_t p0 = (size_t)lr0 * rc.width;
    k_resolve_loss<<<(rc.n_pix + 255) / 256, 256, 0, sc_st>>>((const float*)h->partial.ptr + p0 * rc.groups * 3, rc.n_pix, rc.groups, o->target_dev, rc.row_begin,
                                                              rc.row_step, rc.width, inv_n, 1.0f / (float)rc.spp, nullptr, (float*)h->dldpx.ptr + p0 * 3, o->st.lossgrad);
    IPT_CUDA(cudaGetLastError());
    g_launches += 1;
    IPT_CUDA(cudaMemsetAsync(ctr + 4 + c, 0, sizeof(unsigned long long), sc_st));   // the chunk's work counter, reused by its adjoint launch
    if (grad_chunk(h, plan, sc, g, c, k, o->target_dev, o->st.lossgrad, sc_st) != 0) return -1;
  }
  IPT_CUDA(f.end());
  IPT_CUDA(cudaEventRecord(h->ev_stop[slot], st));
  h->ev_count++;
  if (g.acc_mode && o->nT) {
    IPT_CUDA(launch_reduce_blocks((const float*)h->block_acc.ptr, k * g.grid, 3 * o->nT, o->st.lossgrad + 1, 0, 0, st));
    g_launches += 1;
  }
  k_inverse_update<<<1, 256, 0, st>>>(o->st, o->peers, o->cfg.rank, o->cfg.world, o->nT, h->mat, o->cfg.lr, o->cfg.beta1, o->cfg.beta2, o->cfg.eps,
                                     o->cfg.kd_min, o->cfg.kd_max, kHistoryCap, (unsigned long long)(o->cfg.timeout_ms > 0 ? o->cfg.timeout_ms : 10000) * 1000000ull);
  IPT_CUDA(cudaGetLastError());
  g_launches += 1;
  o->steps_issued++;
  return 0;
}

}  // namespace ipt

using namespace ipt;

extern "C" {

void iptInverseDefaultConfig(IptInverseConfig* c) {
  if (!c) return;
  c->lr = 0.02f; c->beta1 = 0.9f; c->beta2 = 0.999f; c->eps = 1e-8f;
  c->kd_min = 1e-3f; c->kd_max = 1.0f;
  c->rank = 0; c->world = 1;
  c->timeout_ms = 10000;
}

int iptInverseCreate(void* scene, const IptRenderParams* p, const unsigned char* target_dev, const IptInverseConfig* cfg, void** out) {
  if (out) *out = nullptr;
  SceneHandle* h = checked_device(scene);
  if (!h) return -1;
  if (!p || !target_dev || !cfg || !out) return fail("iptInverseCreate: bad arguments");
  if (cfg->world < 1 || cfg->world > kMaxWorld || cfg->rank < 0 || cfg->rank >= cfg->world) return fail("iptInverseCreate: rank / world out of range (world <= 8)");
  std::lock_guard<std::recursive_mutex> lk(h->mutex);
  IPT_CUDA(cudaSetDevice(h->device));
  InverseHandle* o = new InverseHandle();
  o->scene = h; o->scene_key = scene; o->params = *p; o->cfg = *cfg;
  o->nT = (int)h->host.tris.size();
  o->n = 1 + 3 * o->nT;
  o->target_dev = target_dev;
  const size_t nf = (size_t)3 * o->nT;
  // one allocation: kd, m, v (float) | lossgrad, history (double) | step (u64) | error (int)
  const size_t bytes = 3 * nf * sizeof(float) + 16 + ((size_t)o->n + kHistoryCap) * sizeof(double) + 2 * sizeof(unsigned long long);
  cudaError_t e = cudaMalloc(&o->state_mem, bytes);
  if (e != cudaSuccess) { delete o; return fail(std::string("iptInverseCreate: ") + cudaGetErrorString(e)); }
  cudaMemsetAsync(o->state_mem, 0, bytes, h->stream);
  char* base = (char*)o->state_mem;
  size_t off = ((3 * nf * sizeof(float)) + 15) & ~(size_t)15;
  o->st.kd = (float*)base; o->st.m = o->st.kd + nf; o->st.v = o->st.m + nf;
  o->st.lossgrad = (double*)(base + off);
  o->st.history = o->st.lossgrad + o->n;
  o->st.step = (unsigned long long*)(o->st.history + kHistoryCap);
  o->st.error = (int*)(o->st.step + 1);
  memset(&o->peers, 0, sizeof(o->peers));
  if (cfg->world > 1) {
    o->mailbox_bytes = mailbox_size(cfg->world, o->n);
    e = cudaMalloc((void**)&o->mailbox, o->mailbox_bytes);
    if (e != cudaSuccess) { cudaFree(o->state_mem); delete o; return fail(std::string("iptInverseCreate: ") + cudaGetErrorString(e)); }
    cudaMemsetAsync(o->mailbox, 0, o->mailbox_bytes, h->stream);
    o->peers.box[cfg->rank] = o->mailbox;
  } else {
    o->connected = true;
  }
  // start from the scene's current albedos
  std::vector<float> kd(nf);
  for (int i = 0; i < o->nT; i++)
    for (int j = 0; j < 3; j++) kd[3 * (size_t)i + j] = h->host.tris[i].mat.kd[j];
  if (nf) cudaMemcpyAsync(o->st.kd, kd.data(), nf * sizeof(float), cudaMemcpyHostToDevice, h->stream);
  e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { cudaFree(o->mailbox); cudaFree(o->state_mem); delete o; return fail(std::string("iptInverseCreate: ") + cudaGetErrorString(e)); }
  *out = o;
  return 0;
}

int iptInverseMailbox(void* opt, void** mailbox_dev, void* ipc_handle64) {
  InverseHandle* o = (InverseHandle*)opt;
  if (!o) return fail("iptInverseMailbox: bad arguments");
  if (mailbox_dev) *mailbox_dev = o->mailbox;
  if (ipc_handle64) {
    if (!o->mailbox) return fail("iptInverseMailbox: world == 1 has no mailbox");
    IPT_CUDA(cudaSetDevice(o->scene->device));
    cudaIpcMemHandle_t hd;
    IPT_CUDA(cudaIpcGetMemHandle(&hd, o->mailbox));
    static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(ipc_handle64, &hd, 64);
  }
  return 0;
}

int iptInverseConnect(void* opt, void* const* peer_mailboxes, const void* ipc_handles) {
  InverseHandle* o = (InverseHandle*)opt;
  if (!o || (!peer_mailboxes && !ipc_handles)) return fail("iptInverseConnect: bad arguments");
  if (o->cfg.world == 1) return 0;
  IPT_CUDA(cudaSetDevice(o->scene->device));
  for (int r = 0; r < o->cfg.world; r++) {
    if (r == o->cfg.rank) continue;
    if (peer_mailboxes) {
      if (!peer_mailboxes[r]) return fail("iptInverseConnect: null peer mailbox");
      o->peers.box[r] = (double*)peer_mailboxes[r];
    } else {
      cudaIpcMemHandle_t hd;
      memcpy(&hd, (const char*)ipc_handles + 64 * (size_t)r, 64);
      void* p = nullptr;
      IPT_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
      o->opened[r] = p;
      o->peers.box[r] = (double*)p;
    }
  }
  o->connected = true;
  return 0;
}

int iptInverseSetKd(void* opt, const float* kd_host, int reset_moments) {
  InverseHandle* o = (InverseHandle*)opt;
  if (!o || !kd_host) return fail("iptInverseSetKd: bad arguments");
  SceneHandle* h = checked_device(o->scene_key);
  if (!h) return -1;
  std::lock_guard<std::recursive_mutex> lk(h->mutex);
  IPT_CUDA(cudaSetDevice(h->device));
  const size_t nf = (size_t)3 * o->nT;
  IPT_CUDA(h->lossgrad.reserve(nf * sizeof(float) + 16));
  IPT_CUDA(cudaMemcpyAsync(h->lossgrad.ptr, kd_host, nf * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  if (o->nT) k_inverse_seed_kd<<<(int)((nf + 255) / 256), 256, 0, h->stream>>>(o->st, (const float*)h->lossgrad.ptr, h->mat, o->nT, reset_moments ? 1 : 0);
  IPT_CUDA(cudaGetLastError());
  g_launches += 1;
  for (int i = 0; i < o->nT; i++)
    for (int j = 0; j < 3; j++) h->host.tris[i].mat.kd[j] = kd_host[3 * (size_t)i + j];
  IPT_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int iptInverseStep(void* opt, unsigned long long first_seed, int n_steps) {
  InverseHandle* o = (InverseHandle*)opt;
  if (!o || n_steps < 0) return fail("iptInverseStep: bad arguments");
  SceneHandle* h = checked_device(o->scene_key);
  if (!h) return -1;
  if (!o->connected) return fail("iptInverseStep: world > 1 needs iptInverseConnect first");
  std::lock_guard<std::recursive_mutex> lk(h->mutex);
  IPT_CUDA(cudaSetDevice(h->device));
  for (int i = 0; i < n_steps; i++)
    if (enqueue_step(o, first_seed + (unsigned long long)i) != 0) return -1;
  h->kd_dirty = true;
  return 0;
}

int iptInverseState(void* opt, double* loss, float* kd_host, unsigned long long* steps, double* loss_history, int history_count) {
  InverseHandle* o = (InverseHandle*)opt;
  if (!o) return fail("iptInverseState: bad arguments");
  SceneHandle* h = checked_device(o->scene_key);
  if (!h) return -1;
  std::lock_guard<std::recursive_mutex> lk(h->mutex);
  IPT_CUDA(cudaSetDevice(h->device));
  IPT_CUDA(cudaStreamSynchronize(h->stream));
  int err = 0;
  unsigned long long t = 0;
  IPT_CUDA(cudaMemcpy(&err, o->st.error, sizeof(int), cudaMemcpyDeviceToHost));
  IPT_CUDA(cudaMemcpy(&t, o->st.step, sizeof(t), cudaMemcpyDeviceToHost));
  if (err) return fail("iptInverseStep: a peer rank did not deliver its partial sums within the timeout");
  if (steps) *steps = t;
  if (loss) {
    *loss = 0.0;
    if (t > 0) IPT_CUDA(cudaMemcpy(loss, o->st.history + (t - 1) % kHistoryCap, sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (kd_host && o->nT) IPT_CUDA(cudaMemcpy(kd_host, o->st.kd, (size_t)3 * o->nT * sizeof(float), cudaMemcpyDeviceToHost));
  if (loss_history && history_count > 0) {
    // the last `history_count` losses, oldest first (at most kHistoryCap are kept)
    std::vector<double> ring(kHistoryCap);
    IPT_CUDA(cudaMemcpy(ring.data(), o->st.history, kHistoryCap * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < history_count; i++) {
      const long long s = (long long)t - history_count + i;
      loss_history[i] = (s >= 0 && (long long)t - s <= kHistoryCap) ? ring[(size_t)(s % kHistoryCap)] : 0.0;
    }
  }
  return 0;
}

int iptInverseDestroy(void* opt) {
  InverseHandle* o = (InverseHandle*)opt;
  if (!o) return 0;
  SceneHandle* h = checked(o->scene_key);   // the scene may already be gone: then its stream is, too
  if (h) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
  }
  for (int r = 0; r < kMaxWorld; r++)
    if (o->opened[r]) cudaIpcCloseMemHandle(o->opened[r]);
  cudaFree(o->mailbox);
  cudaFree(o->state_mem);
  delete o;
  return 0;
}

}  // extern "C"
