// renderer_internal.h -- process-wide renderer state shared by the translation units of libcrt_b200.so.
// One context per process, like the reference's global `renderContext` (kernels.cu:145): the ABI is not
// re-entrant and neither was the reference's.
#pragma once

#include <cuda_runtime.h>

#include <vector>

#include "device_scene.cuh"
#include "mesh_pipeline.cuh"
#include "wide_traverse.cuh"
#include "kernels.h"

#define CRT_CHECK(val) crtCheckCuda((val), #val, __FILE__, __LINE__)
void crtCheckCuda(cudaError_t result, const char* func, const char* file, int line);

enum TraversalMode { TRAVERSAL_WIDE = 0, TRAVERSAL_EXACT = 1, TRAVERSAL_WIDE_UNCERTIFIED = 2 };
enum SceneKind { SCENE_NONE = 0, SCENE_MESH = 1, SCENE_SPHERES = 2 };

struct RendererContext {
    bool initialised = false;
    SceneKind kind = SCENE_NONE;
    renderer_options opts = {-1, 0u, 0, 0, 0, {0, 0, 0}};
    int numSMs = 0;
    int nx = 0, ny = 0, maxDepth = 0;
    CameraDev cam;
    LightDesc light;
    vec3* fb = nullptr; // pinned host memory handed to the caller (kernels.cu:578-580 uses managed memory)
    float* fbDevice = nullptr; // the normalised frame on the device: finalizeFrameTo() writes it, one bulk copy takes it to `fb`

    // mesh scene
    float4* triGeom = nullptr;
    float4* triShade = nullptr;
    float4* nodes = nullptr;
    float4* materials = nullptr;
    float** texData = nullptr;
    int* texWidth = nullptr;
    int* texHeight = nullptr;
    std::vector<float*> texPtrHost;
    int numTextures = 0;
    unsigned int numTriSlots = 0;
    MeshView mesh;
    // the renderer's own wide tree (wide_bvh.h); wide.stackDepth == 0: not built / not usable, exact traversal only
    WideView wide = {};
    unsigned int* batchRedo = nullptr; // ray indices the wide walk could not certify (intersectBatchDevice)
    size_t batchRedoCap = 0;
    WideBvhStats wideStats;
    const void* wideNodesDev = nullptr;        // (kept for getRendererWideTree: tests download and check the tree)
    const unsigned int* wideTriOrigDev = nullptr;
    int traversal = 0;                 // TRAVERSAL_* in effect for this scene

    // sphere scene
    int numSpheres = 0;
    int skyMode = 0;

    // wavefront state: `mp` = mesh pipeline, `wf` = sphere pipeline (accum lives in wf.accum for both)
    MeshState mp = {};
    ChaseRing ring = {};   // hand-over of lagging slots from the wavefront to chaseKernel
    cudaStream_t streamFast = nullptr; // the chaser's stream
    cudaEvent_t evLane = nullptr;
    unsigned long long* laneSums = nullptr;
    unsigned long long* batchScratch = nullptr; // cursor + counters of intersectBatchDevice
    MeshControl* hostCtlFast = nullptr; // pinned
    WfState wf = {};
    bool ownsAccum = false;
    WfControl* hostCtl = nullptr; // pinned
    cudaStream_t stream = nullptr;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    cudaGraphExec_t graphExec = nullptr;
    long long graphKey = -1;
    std::vector<unsigned char> graphSignature; // bytes of every argument the captured launches carry (crtRunMesh compares before re-using the graph)

    int samplesDone = 0; // samples per pixel in the sums (runRenderer sets, continueRenderer adds)
    int traceBlocks = 0; // persistent grid of traceKernel: one resident wave
    int wideTraceBlocks = 0;
    bool counting = false;
    unsigned long long lastNodeVisits = 0, lastTriTests = 0, lastBatchRedo = 0, lastFrameRedo = 0;
    unsigned long long chaserRays = 0, chaserShadowRays = 0, chaserNodeVisits = 0, chaserTriTests = 0; // the chaser's share of the last frame
    renderer_stats stats = {};
};

// One renderer per HOST THREAD (the reference has one per process, kernels.cu:145; a single-threaded caller sees no difference):
// the multi-GPU mode runs one host thread per device, each with its own context, arena and stream cache.
extern thread_local RendererContext g_ctx;
extern thread_local renderer_options g_opts;

void crtRunMesh(RendererContext& c, int ns, bool resume);
void crtRunSpheres(RendererContext& c, int ns);
void finalizeFrameTo(RendererContext& c, const float4* accum, float ns, cudaStream_t stream); // fb = sums / ns, on the device, then one bulk copy
