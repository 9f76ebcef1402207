/* ipt_b200.h -- C ABI of libipt_b200.so, the B200-native drop-in for the rendering hot path of
 * bblinn2017/inverse_path_tracer.
 *
 * Part 1 is exactly what the reference's FFI for this path binds (ipt_cuda.py:7-8 loads
 * ./build/libipt.so and ./build/libpt.so and calls these six symbols through ctypes).  Both file
 * names are provided as symlinks to this one library, so a scene made through one name can be used
 * and freed through the other (ipt_cuda.py:112,134 does exactly that).
 *
 * Part 2 adds runtime parameters, raw outputs, ray batches, the material adjoint and device-pointer
 * entry points.  Every extension returns 0 on success and -1 on failure; the message is available
 * from iptGetLastError().  No entry point ever calls exit() (the reference does on load errors,
 * scene_basics.h:164,216-220).  Plain pointers and sizes only -- no torch types.
 */
#ifndef IPT_B200_H
#define IPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Part 1 -- the reference's six entry points
 * ---------------------------------------------------------------------------------------------- */

/* replaces scene.h:177-193.  poss/oris/scls: n pointers to 3 floats (position, axis-angle vector,
 * scale); obj_fs: n OBJ paths; mtl_fs: n MTL paths or inline "*Kd r g b*" strings.  All caller-owned
 * and only read during the call.  Writes an opaque library-owned handle to *scenePtr and returns
 * the triangle count (0 and a NULL handle on failure). */
int loadScene(float** poss, float** oris, float** scls, char** obj_fs, char** mtl_fs, int n, void** scenePtr);

/* replaces scene.h:195-199 */
void freeScene(void* scenePtr);

/* replaces path_trace.cu:227-234: renders 500x500 at 100 spp (scene.h:8-10), Russian roulette .9,
 * unbounded depth, Reinhard u8, and writes an 8-bit RGB PNG.  The seed is time(NULL) like the
 * reference unless the environment variable IPT_SEED is set. */
void createImage(void* scenePtr, char* img_file);

/* replaces inv_path_trace.cu:195-208: reads the 500x500 target PNG, accumulates the light-transport
 * edge statistics and writes 7*(nT+1)*nT floats [weights | pixels | lights] into caller memory. */
void createGraph(void* scenePtr, char* imgFile, float* data);

/* replace inv_path_trace.cu:210-214 and :216-221: nT*3 diffuse albedos in triangle order */
void getMaterials(void* scenePtr, float* materials);
void setMaterials(void* scenePtr, float* materials);

/* ------------------------------------------------------------------------------------------------
 * Part 2 -- extensions
 * ---------------------------------------------------------------------------------------------- */

typedef struct IptRenderParams {
  int width, height;      /* scene.h:8-9   (500, 500)  */
  int spp;                /* scene.h:10    (100)       */
  int max_depth;          /* <= 0: unbounded like the reference; else max radiance() calls per path */
  float p_rr;             /* scene.h:11    (.9f)       */
  uint32_t seed_lo, seed_hi;   /* Philox key */
  int group_samples;      /* samples accumulated sequentially per partial sum; a pixel is the sum of its group partials in
                           * group order.  0 (or > spp): spp, i.e. the reference's single running sum (path_trace.cu:190-193).
                           * -1 (the default): chosen from the FULL image size only, so every shard of an image agrees:
                           * the largest g in {32, 16, ..., 2, 1} with width * height * spp >= g * 2^22 (else 1), i.e. at
                           * least 4 Mi work items.  A group is the unit of work the persistent kernel hands to a lane:
                           * small jobs -- and row shards of them on several GPUs -- need small groups so the launch
                           * does not end in a tail of idle lanes; large jobs prefer large groups (fewer partials), but at 64
                           * a 1024 x 1024 x 1024 spp job split over eight GPUs ends in a tail worth 9 % of its time. */
  int row_begin, row_end; /* pixel-row shard [begin, end); row_end <= 0 means height */
  int row_step;           /* every row_step-th row of the shard, starting at row_begin (<= 1: every row).  Interleaved
                           * shards (begin = rank, step = world) balance the load across GPUs; outputs are compact */
} IptRenderParams;

const char* iptGetLastError(void);
void iptDefaultParams(IptRenderParams* p);
/* directory prepended to relative asset paths in loadScene ("" = current directory, the reference behaviour) */
int iptSetAssetRoot(const char* root);
/* CUDA device used by scenes loaded afterwards from this thread (default -1: current device).
 * -2 loads host-only handles (loader tables, get/setMaterials); every compute call on them fails. */
int iptSetDevice(int device);
/* 0 = auto (brute force up to 128 triangles, BVH above), 1 = force brute force, 2 = force BVH.
 * Like iptSetDevice, iptSetBvhLargeTriangleLimit and iptSetIntegrator this is a setting of the CALLING THREAD that
 * loadScene copies into the handle: a loaded scene keeps the settings it was loaded under, so threads that drive
 * different scenes (eight scenes on eight GPUs from one process) do not interfere. */
int iptSetTraversalMode(int mode);

int iptSceneInfo(void* scene, int* n_triangles, int* n_emissives, int* n_objects, int* n_bvh_nodes, int* uses_bvh);
/* loader tables (triangle order defines ids): verts[nT*9], kd/ks/ke[nT*3], ns[nT], has_vn[nT],
 * normal/center[nT*3], area[nT]; any pointer may be NULL */
int iptGetTriangleTable(void* scene, float* verts, float* kd, float* ks, float* ke, float* ns, int* has_vn,
                        float* normal, float* center, float* area);
int iptGetCamera(void* scene, float* m16);

/* Concurrency: a scene handle runs ONE call at a time.  Every entry point takes the handle's lock while it enqueues,
 * and the device-pointer (`*Device`) calls, which run on a caller-chosen stream, are chained on the device as well: each
 * waits for the scene's previous call (an event recorded after it) before it touches the handle's scratch buffers, so
 * calls issued on different streams or threads serialise instead of sharing the work counter / partial sums.  Use one
 * handle per stream (or per GPU) for concurrent work.  Inside a CUDA stream capture the chaining events are skipped. */

/* closest hit for n rays (p.xyz, d.xyz); tri = -1 and t = +inf on a miss.  Host buffers.
 * The result is the lexicographic minimum of (t, triangle id) over all triangles that pass the reference's FP32 test
 * (scene_basics.h:434-458), i.e. what a loop over all triangles returns.  BVH scenes (> 128 triangles) return the same
 * bits as that loop as long as the test's own rounding -- t is off by about 2^-24 (|n.p| + |k|) / |n.d| -- stays inside
 * the padding of the traversal boxes, 1e-5 (1 + |coordinate|): verified for origins up to 1e4 units from the scene; from
 * 1e5 units one adversarial ray in 39600 got the next triangle instead. */
int iptTraceRays(void* scene, const float* rays, int n, int* tri, float* t);
/* same on device memory, asynchronous on `stream` (a cudaStream_t) */
int iptTraceRaysDevice(void* scene, const float* rays_dev, int n, int* tri_dev, float* t_dev, void* stream);

/* instrumented BVH traversal (device memory): stats_dev[0] = child boxes fetched (32 B each), stats_dev[1] = leaf
 * triangles tested (64 B each) -- the algorithmic bytes of the L2/HBM roofline.  BVH-mode scenes only. */
int iptTraceRaysStatsDevice(void* scene, const float* rays_dev, int n, int* tri_dev, float* t_dev,
                            unsigned long long* stats_dev, void* stream);

/* forward render of the row shard.  img: rows*width*3 float radiance means (may be NULL);
 * img8: rows*width*3 tone-mapped bytes (may be NULL); counters[3] = paths, extension rays,
 * shadow rays (may be NULL).  Host buffers, synchronous. */
int iptRender(void* scene, const IptRenderParams* p, float* img, unsigned char* img8, unsigned long long* counters);
/* device buffers, asynchronous on `stream`; counters_dev is zeroed by the call */
int iptRenderDevice(void* scene, const IptRenderParams* p, float* img_dev, unsigned char* img8_dev,
                    unsigned long long* counters_dev, void* stream);
int iptCreateImageEx(void* scene, const IptRenderParams* p, const char* png_path);

/* createGraph with runtime parameters.  target: height*width*3 bytes (host).  raw (optional):
 * (nT+1)*nT*16 doubles per edge [n, w_sum, pixel_sum[2][3], light_sum[2][3], factors_sum[2]],
 * rows dst = 0..nT-1 then the eye row; data (optional): 7*(nT+1)*nT floats (reference layout). */
int iptCreateGraphEx(void* scene, const IptRenderParams* p, const unsigned char* target, double* raw, float* data,
                     unsigned long long* counters);
/* raw edge sums of this shard only, left on the device (all-reduce them, then iptGraphCompress) */
int iptGraphAccumulateDevice(void* scene, const IptRenderParams* p, const unsigned char* target_dev, double* raw_dev, void* stream);
/* DataWrapper::compress (inv_scene.h:87-115) on host arrays / on device arrays (same bits: the log is the canon's) */
int iptGraphCompress(const double* raw, int n_triangles, float* data);
int iptGraphCompressDevice(const double* raw_dev, int n_triangles, float* data_dev, void* stream);

/* NEW (no reference counterpart): loss = sum over the shard of (L/(1+L) - target/255)^2 / (3*W*H) and
 * its gradient with respect to every diffuse albedo, nT*3 doubles.  max_depth is clamped to 64.  The sampler does
 * not depend on Kd, so the estimator is a polynomial in the albedos and its gradient is exact -- with a Phong lobe
 * (Ks != 0) too: the adjoint then records the per-channel throughput factor and direct BSDF of every bounce. */
int iptRenderGrad(void* scene, const IptRenderParams* p, const unsigned char* target, double* loss, double* grad, float* img);
/* device buffers: lossgrad_dev[1 + nT*3] doubles (loss first), img_dev optional; asynchronous */
int iptRenderGradDevice(void* scene, const IptRenderParams* p, const unsigned char* target_dev, double* lossgrad_dev,
                        float* img_dev, void* stream);
/* albedos from device memory (nT*3 floats), asynchronous on `stream` */
int iptSetMaterialsDevice(void* scene, const float* kd_dev, void* stream);

/* GPU BVH builder reproducing BVH::build (bvh.h:109-205) over n primitives: boxes[n*6] (lo, hi),
 * pos[n*3] (split positions).  Outputs: node_boxes[max_nodes*6], node_ints[max_nodes*3] =
 * (start, nPrims, rightOffset), perm[n] (primitive order).  Returns the node count or -1. */
int iptBuildBVH(const float* boxes, const float* pos, int n, int leaf_size, float* node_boxes, int* node_ints,
                int max_nodes, int* perm);

/* builder used by iptBuildBVH and by BVH-mode scene loads: 0 = auto (CUDA kernels from 2048 primitives
 * on, one host thread below), 1 = host, 2 = CUDA kernels.  All produce bit-identical trees. */
int iptSetBvhBuilder(int mode);

/* BVH scenes keep their few LARGE triangles (box diagonal > 1/4 of the scene's; room walls around a dense mesh)
 * out of the tree and test them brute force from shared memory before each traversal: the tree stays tight and
 * the traversal starts with a bound.  `max_count` caps that set (default 128, 0 = every triangle goes into the
 * tree).  Applies to scenes loaded afterwards; results do not depend on it.  iptSceneLargeTriangles returns the
 * size of the set for a loaded scene (0 for brute-force scenes), or -1. */
int iptSetBvhLargeTriangleLimit(int max_count);
int iptSceneLargeTriangles(void* scene);

/* forward integrator for scenes traversed through the BVH: 0 = auto (wavefront), 1 = persistent megakernel
 * (the one brute-force scenes always use), 2 = wavefront: path state in device memory, one closest-hit
 * kernel per bounce over the compacted live paths.  `wavefront_slots` bounds the paths in flight per batch
 * (0 = default, 16 Mi).  All produce bit-identical images.  iptSetIntegrator sets the calling thread's default for
 * scenes loaded afterwards; iptSceneSetIntegrator changes one loaded scene (later render calls on it).  The
 * device-pointer render calls stay asynchronous with a bounded max_depth; with unbounded depth the wavefront reads
 * the number of live paths back after every bounce (one stream synchronisation per bounce). */
int iptSetIntegrator(int mode, int wavefront_slots);
int iptSceneSetIntegrator(void* scene, int mode, int wavefront_slots);

/* ---- part 3: the inverse loop inside the library (new; the reference's material recovery is the GCN of ipt.py) --------
 * One optimiser object per (scene, GPU).  iptInverseStep enqueues whole iterations on the scene's own stream -- forward
 * render of this rank's row shard at p->spp samples (seed = first_seed + i), loss against the target (the loss of
 * iptRenderGrad), path adjoint, sum of [loss, dLoss/dKd] over the `world` ranks that share the image, Adam step
 * (torch.optim.Adam's arithmetic, float32), clamp to [kd_min, kd_max], new albedos into the scene -- and returns without
 * waiting.  With world > 1 the sum is part of the update kernel: every rank stores its partial sums into every peer's
 * mailbox over NVLink and adds the `world` vectors in rank order, so all ranks hold bit-identical albedos without a
 * broadcast.  The mailboxes are ordinary device allocations exchanged once: iptInverseMailbox returns the local
 * pointer and/or a 64-byte CUDA IPC handle; iptInverseConnect takes either raw peer pointers (ranks in one process)
 * or the `world` x 64 bytes of IPC handles gathered from all ranks (one process per GPU).  A peer that does not arrive
 * within timeout_ms makes the next iptInverseState fail instead of hanging the GPU. */
typedef struct IptInverseConfig {
  float lr, beta1, beta2, eps;   /* Adam; defaults 0.02, 0.9, 0.999, 1e-8 */
  float kd_min, kd_max;          /* clamp after every step; defaults 1e-3, 1 */
  int rank, world;               /* this GPU's place among the GPUs sharing the image (row shard = p->row_begin/row_step) */
  int timeout_ms;                /* exchange timeout; default 10000 */
} IptInverseConfig;
void iptInverseDefaultConfig(IptInverseConfig* cfg);
/* target_dev: height*width*3 bytes on the scene's device, caller-owned, must outlive the optimiser.  Starts from the
 * scene's current albedos. */
int iptInverseCreate(void* scene, const IptRenderParams* p, const unsigned char* target_dev, const IptInverseConfig* cfg, void** opt);
int iptInverseMailbox(void* opt, void** mailbox_dev, void* ipc_handle64);
int iptInverseConnect(void* opt, void* const* peer_mailboxes, const void* ipc_handles);
/* albedos (nT*3 floats, host) into the optimiser and the scene; reset_moments != 0 also clears Adam's state and step count */
int iptInverseSetKd(void* opt, const float* kd_host, int reset_moments);
int iptInverseStep(void* opt, unsigned long long first_seed, int n_steps);
/* waits for the enqueued steps; loss = global loss of the last step, kd_host[nT*3], steps = iterations applied,
 * loss_history[history_count] = the last losses, oldest first (any pointer may be NULL) */
int iptInverseState(void* opt, double* loss, float* kd_host, unsigned long long* steps, double* loss_history, int history_count);
int iptInverseDestroy(void* opt);

int iptWritePNG(const char* path, const unsigned char* rgb, int width, int height);
int iptReadPNG(const char* path, unsigned char* rgb, int capacity_bytes, int* width, int* height);

/* device times (ms) of the most recent persistent path-kernel launches on this scene, newest first
 * (CUDA events recorded on the launching stream around each k_paths launch; ring of 64).  Blocks until
 * those launches have finished.  Returns how many were written, or -1. */
int iptPathKernelTimes(void* scene, float* ms, int max_count);
/* counters (paths, extension rays, shadow rays) of the last launch that used the scene's internal
 * counter block (graph / grad passes, or render calls given no counter buffer); synchronous */
int iptLastCounters(void* scene, unsigned long long* counters);

/* Measurement aid for the traversal roofline (SURVEY.md section 8d asks for an L2-resident copy figure, which
 * MEASURED_PEAKS.json does not hold): copies `bytes` (rounded down to 16) from one device buffer to another
 * `reps` times with a grid of SMs x 8 blocks of 256 threads (128-bit loads and stores) and returns the
 * read + write rate of the timed repetitions in GB/s, or a negative value on error.  Buffers that fit the
 * L2 together measure the L2; larger ones HBM.  Synchronous; uses the device selected by iptSetDevice. */
double iptMeasureCopyGBps(long long bytes, int reps);

/* number of kernels launched by this library since load (bench.py's gpu_launches) */
unsigned long long iptKernelLaunchCount(void);

#ifdef __cplusplus
}
#endif
#endif /* IPT_B200_H */
