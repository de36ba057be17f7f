// kernels.h -- the C ABI of the B200 wavefront path tracer (libcrt_b200.so).
//
// Part 1 is the drop-in boundary: the three entry points of the reference's
// kernels.h:6-8 with identical names, argument order, by-value PODs and
// blocking semantics (implementation reference: kernels.cu:571-680).  A host
// program written against the reference headers (reference main.cpp:94,98,138)
// links against this library unchanged; oracle/ref_driver.cpp is built both
// ways to prove it.
//
// Part 2 is strictly additive: entry points the reference lacks but its
// README scenes / BASELINE configs need (SURVEY.md 8b "gaps"): a sphere-scene
// initialiser, a ray-batch intersector, options (device, sample stream,
// deferred finalisation for the multi-GPU reduce) and counters.
//
// Error behaviour (all entry points), as reference kernels.cu:30-37: a CUDA
// failure prints "CUDA error = ... at file:line 'call'" to stderr, resets the
// device and calls exit(99).  There are no return codes on part 1.
#pragma once

#include "rt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

// ---------------------------------------------------------------- part 1 ----
// reference kernels.h:6 / kernels.cu:571.  Copies the scene to the device
// (nothing host-side is retained), allocates the frame buffer as pinned,
// device-mapped host memory (the reference uses managed memory; its caller only
// dereferences it on the host) and returns it through *fb (host-dereferenceable
// after runRenderer, valid until cleanupRenderer).  fb[j*nx+i] is linear RGB, row 0 = bottom.
void initRenderer(const kernel_scene sc, const camera cam, vec3** fb, int nx, int ny, int maxDepth);

// reference kernels.h:7 / kernels.cu:652.  Blocking.  Overwrites fb with the
// mean of `ns` samples per pixel.  tx,ty were the reference's block shape; the
// wavefront kernels choose their own launch shapes, so they are accepted and
// ignored (they never changed the image).
void runRenderer(int ns, int tx, int ty);

// reference kernels.h:8 / kernels.cu:666.  Ends the frame: every pointer
// initRenderer produced (including *fb) is dead afterwards.  The device memory,
// streams and pinned buffers stay cached inside the library for the next
// initRenderer of the process (a warm frame makes no allocation call); they are
// returned by rendererReleaseCaches(), or here when resetDeviceOnCleanup is set
// (the reference ends with cudaDeviceReset(); see renderer_options).
void cleanupRenderer();

// ---------------------------------------------------------------- part 2 ----
typedef struct renderer_options {
    int device;               // CUDA ordinal to render on; -1 = keep the current device
    unsigned int sampleStream;// g: per-pixel seed is wang_hash(pixelId + g*nx*ny) (reference kernels.cu:542 is g = 0)
    int deferFinalize;        // 0: runRenderer writes fb = sum/ns (reference).  1: keep un-normalised sums
                              //    on the device for an external reduce; call finalizeFrame afterwards
    int resetDeviceOnCleanup; // 1: cleanupRenderer ends with cudaDeviceReset() like kernels.cu:679
                              //    (default 0: fatal inside a process that shares the device, e.g. torch)
    int megaBatch;            // iterations launched between host checks of the live-path counter (0 = default)
    int reserved[3];          // [0]: path slots per pixel (0/1 = one slot per pixel = the reference's RNG streams;
                              //      k > 1 = k independent streams per pixel, ns/k samples each, throughput mode)
} renderer_options;

// Applies to the NEXT initRenderer* call.  Passing NULL restores the defaults.
void setRendererOptions(const renderer_options* opt);

// Sphere scenes (README era; no entry point exists at the reference's HEAD).
// `spheres[i]` uses `materials[i]`; both arrays are copied into __constant__
// memory (n <= 1024).  Sky is the gradient of kernels.cu:419-421, there is no
// light and no next-event estimation; everything else follows color().
void initRendererSpheres(const sphere* spheres, const material* materials, int n, const camera cam, vec3** fb,
                         int nx, int ny, int maxDepth);

// Closest-hit query of a ray batch against the scene given to initRenderer:
// the arithmetic of hit()/hitMesh()/hitBvh() (kernels.cu:325,296,154) on
// `n` rays.  origins/dirs are n*3 floats (dirs are normalised inside exactly
// as the reference's ray constructor does, ray.h:9).  Outputs: t (FLT_MAX on
// miss), triangle slot id (-1 on miss), meshID (-1 on miss).  All pointers are
// HOST pointers; copies are part of the call.
void intersectBatch(const float* origins, const float* dirs, long long n, float tMin, float tMax, float* outT,
                    int* outTriId, int* outMeshId);

// Same query with DEVICE-resident SoA buffers: rays as float4 (ox,oy,oz,tMin)
// and (dx,dy,dz,tMax), result float4 (t,u,v, triId as int bits) plus meshID.
// Used by the 64 Mi-ray microbench (BASELINE config 5).  Returns kernel ms.
float intersectBatchDevice(const void* dRayO, const void* dRayD, long long n, void* dHit, int* dMeshId);
// anyHit != 0: the shadow-ray query, hitMesh(.., isShadow = true) (kernels.cu:207,500): t = 0.0f when any triangle lies
// in (tMin, tMax), FLT_MAX otherwise; ids are -1. anyHit == 0: intersectBatchDevice.
float intersectBatchDeviceEx(const void* dRayO, const void* dRayD, long long n, void* dHit, int* dMeshId, int anyHit);

// Fills device ray buffers with the config-5 batch: the first
// half are jittered camera rays over a virtual filmW x filmH film, the second
// half are incoherent rays (origin uniform in the scene bounds, direction
// from the unit-sphere rejection sampler), seeded per ray with the
// kernels.cu:542 formula applied to the ray index.
void generateRayBatchDevice(void* dRayO, void* dRayD, long long n, int filmW, int filmH, float tMin, float tMax);

// Returns the cached device arena, streams, events and pinned buffers to the
// driver.  No-op while a frame is live (between initRenderer and cleanupRenderer).
void rendererReleaseCaches();

// The reference's BSDF library beyond the three materials material_scatter dispatches (material.h:33-143) through its
// scene presets (scene_materials.h:22-93, numbered in file order: 0 floor_coat, 1 floor_diffuse, 2 floor_checker,
// 3 model_coat, 4 model_diffuse, 5 model_glossy, 6 model_glass, 7 model_tintedglass, 8 model_sss; 9 = subsurface_bsdf with
// the sss constants), evaluated on a batch of surface points.  in: 12 floats per item {normal.xyz, t, p.xyz, inside,
// wo.xyz, rng state bits}; out: 12 floats {wi.xyz, t, throughput.xyz, specular | refracted << 1 (int bits), rng state bits
// after, 0, 0, 0}.  HOST pointers; returns 0, or -1 for an unknown preset.
int scatterBatch(int preset, long long n, const float* in, float* out);

// Self test: the light sampler's sine / cosine (the math library's fast path restated, csrc/wavefront_kernels.cuh) against
// sinf / cosf for every float in [0, 2*pi]. Returns the number of differing arguments (0 expected).
long long rendererTrigSelfTest(void);

void* rendererDeviceAlloc(size_t bytes);
void rendererDeviceFree(void* p);
void rendererCopyToHost(void* dst, const void* dSrc, size_t bytes);
void rendererCopyToDevice(void* dDst, const void* src, size_t bytes);

typedef struct renderer_stats {
    unsigned long long raysExtend;  // closest-hit rays traced by the last runRenderer (primary + secondary) = calls of hit(.., false)
    unsigned long long raysShadow;  // any-hit rays traced by the last runRenderer = calls of hit(.., true)
    unsigned long long samples;     // nx*ny*ns
    unsigned long long kernelLaunches; // kernels launched by the last runRenderer
    unsigned long long iterations;  // wavefront iterations that had work
    unsigned long long resumes;     // rays that ran out of their per-launch step budget and continued in a later launch
    unsigned long long deferred;    // shade entries postponed one iteration because the slot's shadow ray was pending
    float msTotal;                  // device time of the last runRenderer (CUDA events on the render stream)
    float msTrace, msShade, msOther; // per-kernel-family device time (only when profiling is on)
    int profiled;
} renderer_stats;

void getRendererStats(renderer_stats* out);
// 1: the wavefront alone, one synchronisation per iteration, CUDA events around each kernel family (diagnostics).
// 2: the timed configuration (wavefront + chaser), launched kernel by kernel instead of as a graph with CUDA events around
//    every trace and shade launch; msTrace / msShade are the sums, `iterations` the launch count (roofline numbers).
void setRendererProfiling(int on);

// Multi-GPU support: the per-pixel un-normalised radiance sums (float4 per
// pixel: r,g,b,unused) live on the device.  A launcher with one process per
// GPU reduces them with NCCL and calls finalizeFrame on the root.
void* getRendererAccumDevice();   // device pointer, nx*ny float4
void setRendererAccumDevice(void* dAccum); // render into a caller-owned device buffer (e.g. a torch tensor) instead
void finalizeFrame(int nsTotal);  // fb = accum / nsTotal (blocking)

// Several GPUs of one box inside the library (the reference is single-GPU). setRendererGpus(N) applies to the next
// initRenderer of the calling thread: it brings up N - 1 worker host threads, one per further device (devices d, d+1, ... from
// the caller's current device), each uploading and indexing the whole scene. runRenderer(ns) then renders ns/N samples of every
// pixel per device (device g on RNG stream g; g = 0 is the reference's stream), combines the un-normalised sums with ONE
// ncclReduce to the caller's device and writes fb = sum / ns there; cleanupRenderer ends the workers. NCCL (libnccl.so.2) is
// loaded at run time, only in this mode. Mesh scenes only. N <= 1 restores single-GPU rendering.
void setRendererGpus(int n);
int getRendererGpus(void); // devices the calling thread's renderer uses (1 = single)

// Progressive rendering (mesh scenes, one slot per pixel). A pixel's samples are ONE RNG stream (reference
// kernels.cu:542-548), so a frame can be continued exactly: runRenderer(a) followed by continueRenderer(b) leaves in fb the
// bits runRenderer(a + b) would. saveRendererCheckpoint writes the running sums and the stream positions
// ("CRTCKP01", nx, ny, samples done, stream, nx*ny float4, nx*ny uint32); loadRendererCheckpoint restores them after an
// initRenderer of the same scene / size / sample stream (the reference's wish list, TODO.txt:70). Return 0 on success.
int continueRenderer(int nsMore, int tx, int ty);
int getRendererSamplesDone();
int saveRendererCheckpoint(const char* path);
int loadRendererCheckpoint(const char* path);

// Diagnostic hook for tests: copy one per-slot array of the mesh pipeline ("rayO","rayD","atten","pcol","hit","shO","shD",
// "shL","shC","accum"; float4 per slot / pixel) to the host. Returns the number of bytes copied (0 = unknown name).
size_t rendererDebugRead(const char* name, void* dst, size_t maxBytes);

// Instrumentation for the flop side of the roofline: when on, the traversal kernels count
// internal-node visits (two slab tests each) and triangle tests.  Never on in timed runs.
void setRendererCounting(int on);
void getRendererTraversalCounts(unsigned long long* nodeVisits, unsigned long long* triTests);
// The part of the last frame that chaseKernel did (visits/tests only in counting builds); the rest was traced by traceKernel.
void getRendererChaserCounts(unsigned long long* raysExtend, unsigned long long* raysShadow, unsigned long long* nodeVisits,
                             unsigned long long* triTests);

// The acceleration structure. The reference walks the caller's bvh_node[] heap (kernels.cu:154-224). By default this
// library builds its OWN tree from the caller's triangle[] inside initRenderer (8-wide, quantised child boxes, surface-area
// heuristic: csrc/wide_bvh.h) and certifies per ray that the reference's walk returns the same hit; the few rays it cannot
// certify (exact ties, last-ulp box culls of the caller's tree) are re-traced in the reference's order over the caller's tree.
//   0 TRAVERSAL_WIDE (default)   1 TRAVERSAL_EXACT: the caller's tree in the reference's order for every ray (the checker)
//   2 TRAVERSAL_WIDE_UNCERTIFIED: own tree, no certificate, no re-trace (diagnostic: results may differ on ties)
// Applies to the next initRenderer. -1 restores the default (or CRT_TRAVERSAL=exact|wide|uncertified from the environment).
void setRendererTraversal(int mode);

typedef struct renderer_wide_info {
    int active;                 // the wide tree is in use for this scene
    int traversal;              // mode in effect
    unsigned int numNodes, numTriangles;
    int depth;                  // wide levels (= traversal stack entries per ray)
    int buildThreads;
    float buildMs;              // host build inside initRenderer (overlapped with the texture upload)
    float sahCost;
    unsigned long long lastBatchRedo;  // rays of the last intersectBatch* call answered by the exact kernel
    unsigned long long lastFrameRedo;  // rays of the last runRenderer answered by the exact kernel
} renderer_wide_info;
void getRendererWideInfo(renderer_wide_info* out);
// Test hook: copies the wide tree in use to the host (96-byte node records of csrc/wide_bvh.h; the caller's slot index of
// every leaf triangle). Returns the number of nodes (0 = no wide tree in use).
unsigned int getRendererWideTree(void* nodes, unsigned int nodeCapacity, unsigned int* triOrig, unsigned int triCapacity);

#ifdef __cplusplus
}
#endif
