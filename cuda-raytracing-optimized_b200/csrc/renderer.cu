// renderer.cu -- the C ABI of libcrt_b200.so (include/kernels.h) for the triangle-mesh path.
//
// Replaces the host half of the reference's kernels.cu:
//   initRenderer     kernels.cu:571-650   scene upload (+ re-tiling into 16-byte tiles)
//   runRenderer      kernels.cu:652-664   one frame; here a loop of wavefront iterations, captured in a CUDA graph
//   cleanupRenderer  kernels.cu:666-680
// and keeps the reference's error behaviour (check_cuda, kernels.cu:30-37).
// There is no CPU fallback: without a CUDA device every entry point exits(99).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "kernels.h"
#include "multi_gpu.h"
#include "raybatch_kernels.cuh"
#include "renderer_internal.h"
#include "wavefront_kernels.cuh"
#include "wide_build.cuh"
#include "wide_bvh.h"
#include "wide_traverse.cuh"

void crtCheckCuda(cudaError_t result, const char* func, const char* file, int line) {
    if (result) {
        std::fprintf(stderr, "CUDA error = %s at %s:%d '%s' \n", cudaGetErrorString(result), file, line, func);
        cudaDeviceReset();
        std::exit(99);
    }
}

thread_local RendererContext g_ctx;
thread_local renderer_options g_opts = {-1, 0u, 0, 0, 0, {0, 0, 0}};
static thread_local int g_profiling = 0;
static int g_traversal = -1;
static thread_local bool g_isGpuWorker = false; // setRendererTraversal(); -1 = CRT_TRAVERSAL from the environment, else TRAVERSAL_WIDE

static f3 toF3(const vec3& v) { return mk3(v.e[0], v.e[1], v.e[2]); }

// CRT_TIMING=1: host wall time of the phases of initRenderer / runRenderer on stderr (diagnostic)
struct PhaseTimer {
    bool on;
    std::chrono::steady_clock::time_point t;
    PhaseTimer() : on(std::getenv("CRT_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        cudaDeviceSynchronize();
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[crt timing] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

// ------------------------------------------------------------ allocation --
// Device memory comes from one arena that outlives cleanupRenderer: a host program that renders frame after frame
// (initRenderer -> runRenderer -> cleanupRenderer, bench.py's e2e step) was spending 100 ms .. 1.6 s per frame in ~40
// cudaMalloc/cudaFree calls (measured with CRT_TIMING=1; cudaFree synchronises and unmaps). The first frame of a process
// allocates block by block; cleanupRenderer folds what the frame needed into ONE block that the next frame bumps through,
// so a warm frame makes no allocation call at all. Streams, events and the pinned buffers are kept the same way.
// rendererReleaseCaches() (and cleanup with resetDeviceOnCleanup) gives everything back.
struct DeviceArena {
    std::vector<std::pair<char*, size_t>> blocks; // blocks[0] = main block (may be absent), others = overflow of this frame
    bool hasMain = false;
    size_t used = 0;     // bytes handed out of the main block
    size_t wanted = 0;   // bytes requested since the last reset
    int device = -1;
};
static thread_local DeviceArena g_arena;

static void arenaFreeAll() {
    for (auto& b : g_arena.blocks) cudaFree(b.first);
    g_arena.blocks.clear();
    g_arena.hasMain = false;
    g_arena.used = g_arena.wanted = 0;
}

static void* arenaAlloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (!bytes) bytes = 256;
    g_arena.wanted += bytes;
    if (g_arena.hasMain && g_arena.used + bytes <= g_arena.blocks[0].second) {
        void* p = g_arena.blocks[0].first + g_arena.used;
        g_arena.used += bytes;
        return p;
    }
    char* p = nullptr;
    CRT_CHECK(cudaMalloc((void**)&p, bytes));
    g_arena.blocks.push_back({p, bytes});
    return p;
}

// End of a frame's lifetime: every arena pointer dies. If the frame overflowed the main block, replace the blocks by one
// block of the size the frame needed.
static void arenaReset() {
    const size_t need = g_arena.wanted;
    const bool overflowed = g_arena.blocks.size() > (g_arena.hasMain ? 1u : 0u);
    if (overflowed) {
        arenaFreeAll();
        char* p = nullptr;
        if (need && cudaMalloc((void**)&p, need) == cudaSuccess) {
            g_arena.blocks.push_back({p, need});
            g_arena.hasMain = true;
        } else {
            cudaGetLastError(); // no room for the cache: the next frame allocates block by block again
        }
    }
    g_arena.used = g_arena.wanted = 0;
}

template <class T>
static T* devAlloc(size_t count) {
    return (T*)arenaAlloc(count * sizeof(T));
}

// Streams, events and pinned host buffers, created once per process (per device).
#define CHASE_STREAMS 16
#define UPLOAD_THREADS 8           // host threads staging one upload (at most; CRT_UPLOAD_THREADS)
#define UPLOAD_SLOTS 2             // pinned chunks per thread (one being filled while the other is on the bus)
#define UPLOAD_CHUNK (4u << 20)    // bytes per chunk
struct HostCache {
    int device = -1;
    cudaStream_t stream = nullptr, streamFast = nullptr;
    cudaStream_t chaseStreams[CHASE_STREAMS] = {}; // high priority: one wave of chaseKernel each, round robin
    cudaEvent_t evLane = nullptr, evStart = nullptr, evStop = nullptr;
    void* hostCtl = nullptr;      // pinned
    void* hostCtlFast = nullptr;  // pinned
    void* fb = nullptr;           // pinned + mapped frame buffer
    size_t fbBytes = 0;
    // uploads from the caller's pageable memory (hostUpload below)
    unsigned char* upPinned = nullptr;
    cudaStream_t upStreams[UPLOAD_THREADS] = {};
    cudaEvent_t upDone[UPLOAD_THREADS][UPLOAD_SLOTS] = {};
};
static thread_local HostCache g_cache;

static void releaseCaches() {
    if (g_cache.device >= 0) {
        cudaFreeHost(g_cache.hostCtl); cudaFreeHost(g_cache.hostCtlFast); cudaFreeHost(g_cache.fb);
        cudaEventDestroy(g_cache.evLane); cudaEventDestroy(g_cache.evStart); cudaEventDestroy(g_cache.evStop);
        cudaStreamDestroy(g_cache.streamFast); cudaStreamDestroy(g_cache.stream);
        for (auto& cs : g_cache.chaseStreams) cudaStreamDestroy(cs);
        if (g_cache.upPinned) {
            cudaFreeHost(g_cache.upPinned);
            for (int t = 0; t < UPLOAD_THREADS; t++) {
                cudaStreamDestroy(g_cache.upStreams[t]);
                for (int k = 0; k < UPLOAD_SLOTS; k++) cudaEventDestroy(g_cache.upDone[t][k]);
            }
        }
    }
    g_cache = HostCache();
    arenaFreeAll();
    g_arena.device = -1;
    cudaGetLastError();
}

extern "C" void rendererReleaseCaches() {
    if (g_ctx.initialised) return; // the live frame owns arena memory: call after cleanupRenderer
    releaseCaches();
}

// fb = sums / ns (kernels.cu:568) on `stream`: the division on the device, then one bulk copy into the caller's pinned frame buffer.
void finalizeFrameTo(RendererContext& c, const float4* accum, float ns, cudaStream_t stream) {
    const unsigned int npix = (unsigned int)c.nx * (unsigned int)c.ny;
    if (!npix) return;
    finalizeKernel<<<(npix + 255) / 256, 256, 0, stream>>>(accum, c.fbDevice, npix, ns);
    CRT_CHECK(cudaMemcpyAsync(c.fb, c.fbDevice, (size_t)npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, stream));
}

// Host-to-device copy of caller-owned PAGEABLE memory (the ABI hands over plain host pointers: kernels.h). cudaMemcpy stages such
// a copy through the driver's own pinned buffer on the calling thread: ~11 GB/s measured on the benchmark scene's 137 MB, i.e.
// 12.5 of initRenderer's 16.5 ms. Here UPLOAD_THREADS host threads copy alternate chunks into pinned buffers of the library
// (cached per process like the streams) and send each chunk on a stream of their own, so staging and bus transfers overlap.
// Blocking: returns when the bytes are on the device. Small copies take the plain path.
static void uploadResources(HostCache& hc) {
    if (hc.upPinned) return;
    CRT_CHECK(cudaMallocHost(&hc.upPinned, (size_t)UPLOAD_THREADS * UPLOAD_SLOTS * UPLOAD_CHUNK));
    for (int t = 0; t < UPLOAD_THREADS; t++) {
        CRT_CHECK(cudaStreamCreateWithFlags(&hc.upStreams[t], cudaStreamNonBlocking));
        for (int k = 0; k < UPLOAD_SLOTS; k++) CRT_CHECK(cudaEventCreateWithFlags(&hc.upDone[t][k], cudaEventDisableTiming));
    }
}

static void hostUpload(HostCache& hc, void* dst, const void* src, size_t bytes) {
    static const int threads = std::getenv("CRT_UPLOAD_THREADS") ? std::max(1, std::min(UPLOAD_THREADS, std::atoi(std::getenv("CRT_UPLOAD_THREADS")))) : 6; // (measured on the benchmark scene, init in ms: 2 -> 10.3, 4 -> 8.0, 6 -> 7.3, 8 -> 7.4)
    static const size_t chunk = std::getenv("CRT_UPLOAD_CHUNK_KB") ? std::max((size_t)64 << 10, std::min((size_t)UPLOAD_CHUNK, (size_t)std::atoi(std::getenv("CRT_UPLOAD_CHUNK_KB")) << 10)) : (size_t)1 << 20; // (1 MB chunks: 7.3 ms, 4 MB: 9.1)
    if (bytes < 2 * chunk) {
        if (bytes) CRT_CHECK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
        return;
    }
    uploadResources(hc);
    int dev = 0;
    CRT_CHECK(cudaGetDevice(&dev));
    const size_t chunks = (bytes + chunk - 1) / chunk;
    auto work = [&hc, dst, src, bytes, chunks, dev](int t) {
        CRT_CHECK(cudaSetDevice(dev));
        unsigned int turn = 0;
        for (size_t ch = (size_t)t; ch < chunks; ch += (size_t)threads, turn++) {
            const int slot = (int)(turn % UPLOAD_SLOTS);
            unsigned char* stage = hc.upPinned + ((size_t)t * UPLOAD_SLOTS + slot) * UPLOAD_CHUNK;
            if (turn >= UPLOAD_SLOTS) CRT_CHECK(cudaEventSynchronize(hc.upDone[t][slot])); // the chunk sent from this buffer two turns ago
            const size_t off = ch * chunk, len = std::min(chunk, bytes - off);
            std::memcpy(stage, (const unsigned char*)src + off, len);
            CRT_CHECK(cudaMemcpyAsync((unsigned char*)dst + off, stage, len, cudaMemcpyHostToDevice, hc.upStreams[t]));
            CRT_CHECK(cudaEventRecord(hc.upDone[t][slot], hc.upStreams[t]));
        }
        CRT_CHECK(cudaStreamSynchronize(hc.upStreams[t]));
    };
    std::thread helpers[UPLOAD_THREADS - 1];
    for (int t = 1; t < threads; t++) helpers[t - 1] = std::thread(work, t);
    work(0);
    for (int t = 1; t < threads; t++) helpers[t - 1].join();
}

// The way back: device memory into the caller's pageable memory, same buffers and threads. Per thread the bus transfer of chunk
// k + 1 runs while chunk k is copied out of its pinned buffer. The caller orders `dSrc` (device-synchronised) beforehand.
static void hostDownload(HostCache& hc, void* dst, const void* dSrc, size_t bytes) {
    static const int threads = std::getenv("CRT_UPLOAD_THREADS") ? std::max(1, std::min(UPLOAD_THREADS, std::atoi(std::getenv("CRT_UPLOAD_THREADS")))) : 6;
    const size_t chunk = (size_t)1 << 20;
    if (bytes < 2 * chunk) {
        if (bytes) CRT_CHECK(cudaMemcpy(dst, dSrc, bytes, cudaMemcpyDeviceToHost));
        return;
    }
    uploadResources(hc);
    int dev = 0;
    CRT_CHECK(cudaGetDevice(&dev));
    const size_t chunks = (bytes + chunk - 1) / chunk;
    auto work = [&hc, dst, dSrc, bytes, chunks, chunk, dev](int t) {
        CRT_CHECK(cudaSetDevice(dev));
        unsigned int turn = 0;
        size_t prevOff = 0, prevLen = 0;
        for (size_t ch = (size_t)t; ch < chunks; ch += (size_t)threads, turn++) {
            const int slot = (int)(turn % UPLOAD_SLOTS);
            unsigned char* stage = hc.upPinned + ((size_t)t * UPLOAD_SLOTS + slot) * UPLOAD_CHUNK;
            const size_t off = ch * chunk, len = std::min(chunk, bytes - off);
            CRT_CHECK(cudaMemcpyAsync(stage, (const unsigned char*)dSrc + off, len, cudaMemcpyDeviceToHost, hc.upStreams[t]));
            CRT_CHECK(cudaEventRecord(hc.upDone[t][slot], hc.upStreams[t]));
            if (turn > 0) { // the previous chunk has landed in the other buffer: hand it to the caller while this one travels
                const int prev = (int)((turn - 1) % UPLOAD_SLOTS);
                CRT_CHECK(cudaEventSynchronize(hc.upDone[t][prev]));
                std::memcpy((unsigned char*)dst + prevOff, hc.upPinned + ((size_t)t * UPLOAD_SLOTS + prev) * UPLOAD_CHUNK, prevLen);
            }
            prevOff = off;
            prevLen = len;
        }
        if (turn > 0) {
            const int prev = (int)((turn - 1) % UPLOAD_SLOTS);
            CRT_CHECK(cudaEventSynchronize(hc.upDone[t][prev]));
            std::memcpy((unsigned char*)dst + prevOff, hc.upPinned + ((size_t)t * UPLOAD_SLOTS + prev) * UPLOAD_CHUNK, prevLen);
        }
    };
    std::thread helpers[UPLOAD_THREADS - 1];
    for (int t = 1; t < threads; t++) helpers[t - 1] = std::thread(work, t);
    work(0);
    for (int t = 1; t < threads; t++) helpers[t - 1].join();
}

static void freeWavefront(RendererContext& c) { // everything sized by the slot count; `accum` is per pixel and stays
    WfState& s = c.wf; // (arena memory: nothing to free one by one)
    float4* accum = s.accum;
    std::memset(&s, 0, sizeof(s));
    s.accum = accum;
    if (c.graphExec) { cudaGraphExecDestroy(c.graphExec); c.graphExec = nullptr; }
    c.graphKey = -1;
}

static void allocWavefront(RendererContext& c, unsigned int numSlots) {
    if (c.wf.numSlots == numSlots && c.wf.rayO) return;
    freeWavefront(c);
    WfState& s = c.wf;
    s.numSlots = numSlots;
    s.rayO = devAlloc<float4>(numSlots);
    s.rayD = devAlloc<float4>(numSlots);
    s.atten = devAlloc<float4>(numSlots);
    s.pcol = devAlloc<float4>(numSlots);
    s.hit = devAlloc<float4>(numSlots);
    s.shO = devAlloc<float4>(numSlots);
    s.shD = devAlloc<float4>(numSlots);
    s.shL = devAlloc<float4>(numSlots);
    s.queueA = devAlloc<unsigned int>(numSlots);
    s.queueB = devAlloc<unsigned int>(numSlots);
    s.regen = devAlloc<unsigned int>(numSlots);
    s.ctl = devAlloc<WfControl>(1);
}

static void freeMeshPipeline(RendererContext& c) {
    MeshState& s = c.mp; // (arena memory: nothing to free one by one)
    std::memset(&s, 0, sizeof(s));
    std::memset(&c.ring, 0, sizeof(c.ring));
    if (c.graphExec) { cudaGraphExecDestroy(c.graphExec); c.graphExec = nullptr; }
    c.graphKey = -1;
}

static void allocMeshPipeline(RendererContext& c, unsigned int numSlots) {
    if (c.mp.numSlots == numSlots && c.mp.rayO) return;
    freeMeshPipeline(c);
    MeshState& s = c.mp;
    s.numSlots = numSlots;
    s.rayO = devAlloc<float4>(numSlots);
    s.rayD = devAlloc<float4>(numSlots);
    s.atten = devAlloc<float4>(numSlots);
    s.pcol = devAlloc<float4>(numSlots);
    s.hit = devAlloc<float4>(numSlots);
    s.travE = devAlloc<uint2>(numSlots);
    s.shO = devAlloc<float4>(numSlots);
    s.shD = devAlloc<float4>(numSlots);
    s.shL = devAlloc<float4>(numSlots);
    s.shC = devAlloc<float4>(numSlots);
    s.travS = devAlloc<uint2>(numSlots);
    s.pending = devAlloc<unsigned char>(numSlots);
    s.ready = devAlloc<unsigned char>(numSlots);
    s.rngOut = devAlloc<unsigned int>(numSlots);
    for (int k = 0; k < 2; k++) {
        s.traceQ[k] = devAlloc<unsigned int>(2 * (size_t)numSlots); // at most one extend entry (front) and one shadow entry (back) per slot
        s.shadeQ[k] = devAlloc<unsigned int>(numSlots);
    }
    s.traceCap = 2u * numSlots;
    s.redoQ = devAlloc<unsigned int>(2 * (size_t)numSlots);
    s.ctl = devAlloc<MeshControl>(1);
    c.ring.entries[0] = devAlloc<unsigned int>(numSlots);
    c.ring.entries[1] = devAlloc<unsigned int>(numSlots);
    c.ring.ctl = devAlloc<unsigned int>(16);
    c.ring.counters = devAlloc<unsigned long long>(8);
}

extern "C" void setRendererOptions(const renderer_options* opt) {
    if (opt) g_opts = *opt;
    else g_opts = renderer_options{-1, 0u, 0, 0, 0, {0, 0, 0}};
}

extern "C" void setRendererProfiling(int on) { g_profiling = on; }

extern "C" void setRendererTraversal(int mode) { g_traversal = mode; }

static int traversalMode() {
    if (g_traversal >= 0) return g_traversal;
    const char* v = std::getenv("CRT_TRAVERSAL"); // exact | wide | uncertified (diagnostics, A/B runs)
    if (v && v[0] == 'e') return TRAVERSAL_EXACT;
    if (v && v[0] == 'u') return TRAVERSAL_WIDE_UNCERTIFIED;
    return TRAVERSAL_WIDE;
}

static void initCommon(RendererContext& c, const camera& cam, vec3** fb, int nx, int ny, int maxDepth) {
    if (g_opts.device >= 0) CRT_CHECK(cudaSetDevice(g_opts.device));
    int dev = 0;
    CRT_CHECK(cudaGetDevice(&dev));
    CRT_CHECK(cudaDeviceGetAttribute(&c.numSMs, cudaDevAttrMultiProcessorCount, dev)); // (cudaGetDeviceProperties costs milliseconds)
    c.opts = g_opts;
    c.nx = nx;
    c.ny = ny;
    c.maxDepth = maxDepth;
    c.cam.origin = toF3(cam.origin);
    c.cam.lowerLeft = toF3(cam.lower_left_corner);
    c.cam.horizontal = toF3(cam.horizontal);
    c.cam.vertical = toF3(cam.vertical);
    c.cam.u = toF3(cam.u);
    c.cam.v = toF3(cam.v);
    c.cam.w = toF3(cam.w);
    c.cam.lensRadius = cam.lens_radius;
    if (g_cache.device != dev || g_arena.device != dev) { // first frame of the process, or the caller moved to another GPU
        if (g_cache.device >= 0 && g_cache.device != dev) {
            CRT_CHECK(cudaSetDevice(g_cache.device));
            releaseCaches();
            CRT_CHECK(cudaSetDevice(dev));
        }
        if (g_cache.device < 0) {
            CRT_CHECK(cudaStreamCreateWithFlags(&g_cache.stream, cudaStreamNonBlocking));
            int prLow = 0, prHigh = 0;
            CRT_CHECK(cudaDeviceGetStreamPriorityRange(&prLow, &prHigh));
            CRT_CHECK(cudaStreamCreateWithPriority(&g_cache.streamFast, cudaStreamNonBlocking, prHigh));
            const bool chaseLow = std::getenv("CRT_CHASE_PRIORITY") && std::getenv("CRT_CHASE_PRIORITY")[0] == 'l'; // (experiment)
            for (auto& cs : g_cache.chaseStreams) CRT_CHECK(cudaStreamCreateWithPriority(&cs, cudaStreamNonBlocking, chaseLow ? prLow : prHigh));
            CRT_CHECK(cudaEventCreateWithFlags(&g_cache.evLane, cudaEventDisableTiming));
            CRT_CHECK(cudaEventCreate(&g_cache.evStart));
            CRT_CHECK(cudaEventCreate(&g_cache.evStop));
            CRT_CHECK(cudaMallocHost(&g_cache.hostCtlFast, sizeof(MeshControl)));
            CRT_CHECK(cudaMallocHost(&g_cache.hostCtl, sizeof(WfControl) > sizeof(MeshControl) ? sizeof(WfControl) : sizeof(MeshControl)));
            g_cache.device = dev;
        }
        g_arena.device = dev;
    }
    c.stream = g_cache.stream;
    c.streamFast = g_cache.streamFast;
    c.evLane = g_cache.evLane;
    c.evStart = g_cache.evStart;
    c.evStop = g_cache.evStop;
    c.hostCtlFast = (MeshControl*)g_cache.hostCtlFast;
    c.hostCtl = (WfControl*)g_cache.hostCtl;
    c.laneSums = devAlloc<unsigned long long>(8);
    c.batchScratch = devAlloc<unsigned long long>(8);
    const size_t npix = (size_t)nx * ny;
    // The frame buffer the caller reads after runRenderer (kernels.cu:578-580 uses managed memory; main.cpp:105,119 only
    // ever dereferences it on the host): pinned host memory, so the frame needs no page-fault migration (7.4 ms for 1200x800).
    // finalizeKernel writes the normalised frame into device memory and ONE bulk copy takes it across the bus: round 1 let the
    // kernel store straight into the mapped buffer, which the ncu launch list showed at 3.3 ms for 11.5 MB (4-byte stores over
    // PCIe) in every frame of every path -- the copy engine needs 0.25 ms.
    const size_t fbBytes = (npix ? npix : 1) * sizeof(vec3);
    if (g_cache.fbBytes < fbBytes) {
        if (g_cache.fb) CRT_CHECK(cudaFreeHost(g_cache.fb));
        CRT_CHECK(cudaHostAlloc(&g_cache.fb, fbBytes, cudaHostAllocDefault));
        g_cache.fbBytes = fbBytes;
    }
    c.fb = (vec3*)g_cache.fb;
    if (fb) *fb = c.fb;
    c.fbDevice = devAlloc<float>(3 * (npix ? npix : 1));
    c.wf.accum = devAlloc<float4>(npix);
    c.ownsAccum = true;
    std::memset(&c.stats, 0, sizeof(c.stats));
    c.initialised = true;
}

// ------------------------------------------------------------ wide tree on the device --
struct DeviceWideTree {
    WideNode* nodes = nullptr;
    unsigned int* triOrig = nullptr;
    unsigned int numNodes = 0, numTris = 0, numBinary = 0;
    int depth = 0, levels = 0;
    float range[3] = {0, 0, 0};
    float ms = 0.0f;
};

// Builds the wide tree from the caller's triangle records already on the device (wide_build.cuh). All work is queued on the
// default stream; the host reads one counter per level. Returns false when there is no real triangle.
static bool buildWideOnDevice(const float* dTris, unsigned int numSlots, DeviceWideTree& out) {
    const auto t0 = std::chrono::steady_clock::now();
    PhaseTimer pt;
    GpuBuild b;
    std::memset(&b, 0, sizeof(b));
    const size_t n = numSlots;
    b.n = numSlots;
    b.primSlot = devAlloc<unsigned int>(n);
    b.primLo = devAlloc<float4>(n);
    b.primHi = devAlloc<float4>(n);
    for (int k = 0; k < 2; k++) { b.order[k] = devAlloc<unsigned int>(n); b.nodeOf[k] = devAlloc<unsigned int>(n); b.active[k] = devAlloc<unsigned int>(n); }
    b.nLo = devAlloc<float4>(2 * n + 2);
    b.nHi = devAlloc<float4>(2 * n + 2);
    b.cLo = devAlloc<int>(3 * (2 * n + 2));
    b.cHi = devAlloc<int>(3 * (2 * n + 2));
    b.nFirst = devAlloc<unsigned int>(2 * n + 2);
    b.nCursor = devAlloc<unsigned int>(2 * n + 2);
    b.nSlot = devAlloc<int>(2 * n + 2);
    b.nSplit = devAlloc<uint2>(2 * n + 2);
    b.binCapacity = n / 2 + 1;
    b.binCnt = devAlloc<unsigned int>(b.binCapacity * 3 * GB_BINS);
    b.binLo = devAlloc<int>(b.binCapacity * 3 * GB_BINS * 3);
    b.binHi = devAlloc<int>(b.binCapacity * 3 * GB_BINS * 3);
    b.dpCost = devAlloc<float>(7 * (2 * n + 2));
    b.dpDec = devAlloc<unsigned char>(8 * (2 * n + 2));
    b.pending = devAlloc<unsigned int>(n + 1);
    b.depthOf = devAlloc<unsigned int>(n + 1);
    b.counters = devAlloc<unsigned int>(32);
    out.nodes = (WideNode*)arenaAlloc((n + 1) * sizeof(WideNode));
    out.triOrig = devAlloc<unsigned int>(n);

    unsigned int init[32];
    std::memset(init, 0, sizeof(init));
    for (int a = 0; a < 3; a++) {
        init[8 + a] = 0x7F800000u;  init[11 + a] = 0xFF800000u ^ 0x7FFFFFFFu;  // box min / max (order-preserving images of +inf / -inf)
        init[14 + a] = 0x7F800000u; init[17 + a] = 0xFF800000u ^ 0x7FFFFFFFu;  // centroid min / max
    }
    pt.mark("  wide build: allocation");
    CRT_CHECK(cudaMemcpy(b.counters, init, sizeof(init), cudaMemcpyHostToDevice));
    gbPrimsKernel<<<(numSlots + 255) / 256, 256>>>(dTris, numSlots, b);
    gbRootKernel<<<1, 1>>>(b);
    CRT_CHECK(cudaGetLastError());
    unsigned int h[8];
    CRT_CHECK(cudaMemcpy(h, b.counters, sizeof(h), cudaMemcpyDeviceToHost));
    const unsigned int prims = h[0];
    if (prims == 0) return false;
    float4 rootLo, rootHi;
    CRT_CHECK(cudaMemcpy(&rootLo, b.nLo, sizeof(float4), cudaMemcpyDeviceToHost));
    CRT_CHECK(cudaMemcpy(&rootHi, b.nHi, sizeof(float4), cudaMemcpyDeviceToHost));
    const float lo[3] = {rootLo.x, rootLo.y, rootLo.z}, hi[3] = {rootHi.x, rootHi.y, rootHi.z};
    float pad[3];
    for (int a = 0; a < 3; a++) {
        out.range[a] = std::max(std::max(std::fabs(lo[a]), std::fabs(hi[a])), 1e-30f);
        pad[a] = out.range[a] * WIDE_PAD_SCALE;
    }
    pt.mark("  wide build: primitives");
    // binary tree, level by level
    int cur = 0;
    unsigned int activeCount = h[2];
    const unsigned int primBlocks = (prims + 255) / 256;
    std::vector<unsigned int> levelEnd(1, 1u); // binary nodes [levelEnd[L-1], levelEnd[L]) were created by the split of level L-1
    while (activeCount > 0 && out.levels < 256) {
        gbClearBinsKernel<<<std::min<unsigned int>((activeCount * 3u * GB_BINS + 255u) / 256u, 4096u), 256>>>(b, activeCount);
        gbBinKernel<<<primBlocks, 256>>>(b, cur);
        CRT_CHECK(cudaMemsetAsync(&b.counters[2 + (cur ^ 1)], 0, sizeof(unsigned int)));
        gbSplitKernel<<<(activeCount * 32u + 255u) / 256u, 256>>>(b, cur, activeCount);
        gbPartitionKernel<<<primBlocks, 256>>>(b, cur);
        cur ^= 1;
        unsigned int hc[4];
        CRT_CHECK(cudaMemcpy(hc, b.counters, sizeof(hc), cudaMemcpyDeviceToHost));
        activeCount = hc[2 + cur];
        levelEnd.push_back(hc[1]);
        out.levels++;
    }
    CRT_CHECK(cudaGetLastError());
    pt.mark("  wide build: binary levels");
    // collapse: cost tables bottom-up (deepest binary level first), then the wide nodes level by level
    const int greedy = std::getenv("CRT_WIDE_GREEDY") ? 1 : 0; // (diagnostic: the greedy collapse the cost tables replaced)
    const float triCost = std::getenv("CRT_WIDE_TRICOST") ? (float)std::atof(std::getenv("CRT_WIDE_TRICOST")) : GB_WIDE_TRI_COST; // (tuning knob)
    for (size_t l = levelEnd.size(); l-- > 0;) {
        const unsigned int s0 = l ? levelEnd[l - 1] : 0u, s1 = levelEnd[l];
        if (s1 > s0) gbCollapseCostKernel<<<(s1 - s0 + 127) / 128, 128>>>(b, s0, s1, triCost, 0);
    }
    const unsigned int one = 1u, zero = 0u;
    CRT_CHECK(cudaMemcpy(&b.counters[4], &one, 4, cudaMemcpyHostToDevice));
    CRT_CHECK(cudaMemcpy(&b.counters[5], &zero, 4, cudaMemcpyHostToDevice));
    CRT_CHECK(cudaMemcpy(&b.counters[6], &zero, 4, cudaMemcpyHostToDevice));
    CRT_CHECK(cudaMemcpy(b.pending, &zero, 4, cudaMemcpyHostToDevice));
    CRT_CHECK(cudaMemcpy(b.depthOf, &one, 4, cudaMemcpyHostToDevice));
    unsigned int start = 0, end = 1;
    while (end > start) {
        gbCollapseKernel<<<(end - start + 63) / 64, 64>>>(b, cur, start, end, make_float3(pad[0], pad[1], pad[2]), out.nodes, out.triOrig, greedy);
        start = end;
        CRT_CHECK(cudaMemcpy(&end, &b.counters[4], sizeof(unsigned int), cudaMemcpyDeviceToHost));
    }
    CRT_CHECK(cudaGetLastError());
    pt.mark("  wide build: collapse");
    CRT_CHECK(cudaMemcpy(h, b.counters, sizeof(h), cudaMemcpyDeviceToHost));
    out.numBinary = h[1];
    out.numNodes = h[4];
    out.numTris = h[5];
    out.depth = (int)h[6];
    out.ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return out.numTris == prims;
}

extern "C" void initRenderer(const kernel_scene sc, const camera cam, vec3** fb, int nx, int ny, int maxDepth) {
    RendererContext& c = g_ctx;
    if (c.initialised) cleanupRenderer();
    PhaseTimer pt;
    initCommon(c, cam, fb, nx, ny, maxDepth);
    pt.mark("init: common");
    c.kind = SCENE_MESH;

    const mesh* m = sc.m;
    const unsigned int numSlots = m->numTris;
    // our own tree is built on host threads from the caller's triangles while this thread uploads the scene
    c.traversal = traversalMode();
    // CRT_WIDE_BUILD=host: the host builder (wide_bvh.cpp; the device build's cross-check), overlapped with the uploads
    const char* buildEnv = std::getenv("CRT_WIDE_BUILD");
    const bool hostBuild = buildEnv && buildEnv[0] == 'h';
    WideBvhHost wideHost;
    bool wideBuilt = false;
    std::thread wideBuilder;
    if (hostBuild && c.traversal != TRAVERSAL_EXACT && numSlots > 0)
        wideBuilder = std::thread([&] { wideBuilt = buildWideBvh(m->tris, numSlots, 0, wideHost); });
    // textures: float RGB, nearest lookup (kernels.cu:620-645). They are most of the scene's bytes (113 of 137 MB on the benchmark);
    // uploaded below, after the geometry, while the device re-tiles it.
    c.numTextures = sc.numTextures;
    std::vector<float*> texPtr(sc.numTextures > 0 ? sc.numTextures : 1, nullptr);
    std::vector<int> texW(texPtr.size(), 0), texH(texPtr.size(), 0);
    for (int i = 0; i < sc.numTextures; i++) {
        const stexture& t = sc.textures[i];
        texW[i] = t.width;
        texH[i] = t.height;
        texPtr[i] = (float*)arenaAlloc((size_t)t.width * t.height * 3 * sizeof(float));
    }
    c.texPtrHost = texPtr;
    c.texData = devAlloc<float*>(texPtr.size());
    c.texWidth = devAlloc<int>(texPtr.size());
    c.texHeight = devAlloc<int>(texPtr.size());
    // triangles: upload the caller's 64-byte records once, re-tile on the device, drop the staging copy
    float* staging = (float*)arenaAlloc((size_t)(numSlots ? numSlots : 1) * sizeof(triangle));
    hostUpload(g_cache, staging, m->tris, (size_t)numSlots * sizeof(triangle));
    const unsigned int primsPerLeaf = sc.numPrimitivesPerLeaf > 0 ? (unsigned int)sc.numPrimitivesPerLeaf : 1u;
    const unsigned int leafBytes = 32u * primsPerLeaf + 32u * ((primsPerLeaf + 7u) / 8u);
    const size_t numLeaves = ((size_t)numSlots + primsPerLeaf - 1) / primsPerLeaf;
    const size_t geomBytes = (numLeaves + 1) * leafBytes + 256; // (+1: the tail-mode prefetch of a leaf pair may touch the next leaf)
    c.triGeom = (float4*)arenaAlloc(geomBytes);
    CRT_CHECK(cudaMemset(c.triGeom, 0, geomBytes));
    c.triShade = devAlloc<float4>(3 * (size_t)numSlots);
    if (numSlots) {
        retileTrianglesKernel<<<(numSlots + 255) / 256, 256>>>(staging, numSlots, primsPerLeaf, leafBytes, (float*)c.triGeom, c.triShade);
        CRT_CHECK(cudaGetLastError());
    }
    c.numTriSlots = numSlots;
    pt.mark("init: triangles");

    // nodes: upload the caller's 24-byte records, re-tile into 64-byte child-pair records on the device (intersect.cuh)
    const size_t nodeBytes = (size_t)m->numBvhNodes * sizeof(bvh_node);
    const unsigned int firstLeaf = (unsigned int)(m->numBvhNodes / 2); // kernels.cu:614
    float* nodeStaging = (float*)arenaAlloc(nodeBytes + 64);
    hostUpload(g_cache, nodeStaging, m->bvh, nodeBytes);
    c.nodes = devAlloc<float4>(4 * (size_t)(firstLeaf ? firstLeaf : 1) + 8);
    if (firstLeaf) {
        swizzleNodesKernel<<<(firstLeaf + 255) / 256, 256>>>(nodeStaging, firstLeaf, c.nodes);
        CRT_CHECK(cudaGetLastError());
    }
    pt.mark("init: nodes");
    // the textures go up from a helper thread (with this thread's upload buffers, which this thread does not touch again before
    // the join) while this thread builds the tree: the build is kernels + small read-backs, the upload is host memcpy + bus
    HostCache* uploadCache = &g_cache;
    int uploadDevice = 0;
    CRT_CHECK(cudaGetDevice(&uploadDevice));
    std::thread texUploader([&sc, &texPtr, uploadCache, uploadDevice] {
        CRT_CHECK(cudaSetDevice(uploadDevice));
        for (int i = 0; i < sc.numTextures; i++) {
            const stexture& t = sc.textures[i];
            hostUpload(*uploadCache, texPtr[i], t.data, (size_t)t.width * t.height * 3 * sizeof(float));
        }
    });
    CRT_CHECK(cudaDeviceSynchronize()); // the re-tiling kernels ran on the default stream, the frame runs on c.stream

    c.mesh.nodes = c.nodes;
    c.mesh.tris = c.triGeom;
    c.mesh.firstLeaf = firstLeaf;
    c.mesh.primsPerLeaf = (unsigned int)sc.numPrimitivesPerLeaf;
    c.mesh.leafBytes = leafBytes;
    c.mesh.boundsMin = toF3(m->bounds.min);
    c.mesh.boundsMax = toF3(m->bounds.max);

    // materials (helper_structs.h:133-138) as two float4 each
    std::vector<float4> mats(2 * (size_t)(sc.numMaterials > 0 ? sc.numMaterials : 1));
    for (int i = 0; i < sc.numMaterials; i++) {
        const material& mt = sc.materials[i];
        mats[2 * i] = make_float4(mt.color.e[0], mt.color.e[1], mt.color.e[2], mt.param);
        int type = (int)mt.type, tex = mt.texId;
        float4 b;
        std::memcpy(&b.x, &type, 4);
        std::memcpy(&b.y, &tex, 4);
        b.z = b.w = 0.0f;
        mats[2 * i + 1] = b;
    }
    c.materials = devAlloc<float4>(mats.size());
    CRT_CHECK(cudaMemcpy(c.materials, mats.data(), mats.size() * sizeof(float4), cudaMemcpyHostToDevice));

    CRT_CHECK(cudaMemcpy(c.texData, texPtr.data(), texPtr.size() * sizeof(float*), cudaMemcpyHostToDevice));
    CRT_CHECK(cudaMemcpy(c.texWidth, texW.data(), texW.size() * sizeof(int), cudaMemcpyHostToDevice));
    CRT_CHECK(cudaMemcpy(c.texHeight, texH.data(), texH.size() * sizeof(int), cudaMemcpyHostToDevice));

    pt.mark("init: materials");

    // the wide tree: built on the device from the staged triangles (or by the host builder), leaf triangles re-tiled on the device
    c.wide = WideView{};
    c.wideStats = WideBvhStats();
    if (wideBuilder.joinable()) wideBuilder.join();
    if (c.traversal != TRAVERSAL_EXACT && numSlots > 0 && sc.numPrimitivesPerLeaf > 0 && firstLeaf > 0) {
        uint4* dNodes = nullptr;
        unsigned int* dOrig = nullptr;
        size_t nt = 0;
        float range[3] = {0, 0, 0};
        bool ok = false;
        if (hostBuild) {
            if (wideBuilt) {
                const size_t nn = wideHost.nodes.size();
                nt = wideHost.triOrig.size();
                dNodes = (uint4*)arenaAlloc(nn * sizeof(WideNode));
                dOrig = devAlloc<unsigned int>(nt);
                CRT_CHECK(cudaMemcpy(dNodes, wideHost.nodes.data(), nn * sizeof(WideNode), cudaMemcpyHostToDevice));
                CRT_CHECK(cudaMemcpy(dOrig, wideHost.triOrig.data(), nt * sizeof(unsigned int), cudaMemcpyHostToDevice));
                c.wideStats = wideHost.stats;
                for (int a = 0; a < 3; a++) range[a] = wideHost.range[a];
                ok = true;
            }
        } else {
            DeviceWideTree t;
            ok = buildWideOnDevice(staging, numSlots, t);
            dNodes = (uint4*)t.nodes;
            dOrig = t.triOrig;
            nt = t.numTris;
            c.wideStats.numNodes = t.numNodes;
            c.wideStats.numTris = t.numTris;
            c.wideStats.numBinaryNodes = t.numBinary;
            c.wideStats.maxDepth = t.depth;
            c.wideStats.msTotal = t.ms;
            c.wideStats.threads = 0; // built on the device
            for (int a = 0; a < 3; a++) range[a] = t.range[a];
        }
        pt.mark("init: wide tree build");
        if (ok && c.wideStats.maxDepth <= 27) {
            float4* dTriA = devAlloc<float4>(2 * nt);
            float2* dTriB = devAlloc<float2>(nt);
            unsigned int* dBad = devAlloc<unsigned int>(1);
            CRT_CHECK(cudaMemset(dBad, 0, sizeof(unsigned int)));
            wideTrianglesKernel<<<(unsigned int)((nt + 255) / 256), 256>>>(staging, dOrig, (unsigned int)nt, dTriA, dTriB);
            checkRefTreeKernel<<<(unsigned int)((m->numBvhNodes + 255) / 256), 256>>>(nodeStaging, (unsigned int)m->numBvhNodes, dBad);
            CRT_CHECK(cudaGetLastError());
            unsigned int bad = 0;
            CRT_CHECK(cudaMemcpy(&bad, dBad, sizeof(bad), cudaMemcpyDeviceToHost));
            if (bad == 0) { // (else: the caller's boxes do not nest, the certificate's premise fails: exact traversal only)
                c.wide.nodes = dNodes;
                c.wide.triA = dTriA;
                c.wide.triB = dTriB;
                c.wide.rangeX = WIDE_ORIGIN_RANGE * range[0];
                c.wide.rangeY = WIDE_ORIGIN_RANGE * range[1];
                c.wide.rangeZ = WIDE_ORIGIN_RANGE * range[2];
                // entries per thread: the tree's depth, rounded up to a multiple of 4 and at least 16 (see the carve-out note in crtRunMesh)
                // (+1: the last entry parks the current node's leaf meta bytes, wide_traverse.cuh)
                c.wide.stackDepth = std::max(16u, ((unsigned int)c.wideStats.maxDepth + 1u + 3u) & ~3u);
                if (std::getenv("CRT_WIDE_STACK")) c.wide.stackDepth = std::max((unsigned int)c.wideStats.maxDepth + 1u, (unsigned int)std::atoi(std::getenv("CRT_WIDE_STACK"))); // (diagnostic)
                c.wideNodesDev = dNodes;
                c.wideTriOrigDev = dOrig;
            }
            if (pt.on)
                std::fprintf(stderr, "[crt timing] wide tree: %u nodes, %u triangles, depth %d, %s build %.2f ms, ref tree nests: %s\n", c.wideStats.numNodes,
                             c.wideStats.numTris, c.wideStats.maxDepth, hostBuild ? "host" : "device", c.wideStats.msTotal, bad ? "NO" : "yes");
        }
    }
    pt.mark("init: wide tree upload");
    // light: RenderContext default members, kernels.cu:93-94
    c.light.center = mk3(52.514355f, 715.686951f, -272.620972f);
    c.light.radius = 50.0f;
    c.light.color = mk3(20.0f, 20.0f, 20.0f);

    // setRendererGpus(N > 1): N - 1 further devices, one host thread each (multi_gpu.h). Worker threads themselves are single.
    extern int crtRequestedGpus();
    if (crtRequestedGpus() > 1 && !g_isGpuWorker) {
        int dev = 0;
        CRT_CHECK(cudaGetDevice(&dev));
        c.opts.deferFinalize = 1;
        g_isGpuWorker = true; // (the lambda-spawned threads have their own flag; this one guards re-entry from this thread)
        crtMultiGpuStart(sc, cam, nx, ny, maxDepth, crtRequestedGpus(), dev, c.opts.sampleStream);
        g_isGpuWorker = false;
        c.opts.sampleStream = c.opts.sampleStream * (unsigned int)crtRequestedGpus();
        pt.mark("init: worker devices");
    }
    // the uploads and scene kernels above ran on the legacy stream (some from pageable memory); the frame runs on non-blocking
    // streams, which do not order against it
    texUploader.join();
    pt.mark("init: textures (helper thread) joined");
    CRT_CHECK(cudaDeviceSynchronize());
}

static bool useWideTree(const RendererContext& c) { return c.wide.stackDepth > 0 && c.traversal != TRAVERSAL_EXACT; }

static ShadeScene shadeScene(const RendererContext& c) {
    ShadeScene s;
    s.triShade = c.triShade;
    s.mats.mats = c.materials;
    s.mats.texData = c.texData;
    s.mats.texWidth = c.texWidth;
    s.mats.texHeight = c.texHeight;
    s.light = c.light;
    s.maxDepth = c.maxDepth;
    return s;
}


#ifdef WIDE_SMEM_TOP
#define WIDE_TRACE_EXTRA_SMEM (96u * WIDE_SMEM_TOP)
#else
#define WIDE_TRACE_EXTRA_SMEM 0u
#endif
template <int CUR>
static void launchWideTrace(RendererContext& c, const MeshState& mp, cudaStream_t stream, int traceBlocks) {
    const size_t smem = (size_t)c.wide.stackDepth * WIDE_TRACE_BLOCK * sizeof(uint2) + WIDE_TRACE_EXTRA_SMEM;
    const bool certify = c.traversal != TRAVERSAL_WIDE_UNCERTIFIED;
    WideView wv = c.wide;
#ifdef WIDE_SMEM_TOP
    wv.topCount = std::min<unsigned int>(WIDE_SMEM_TOP, c.wideStats.numNodes);
#endif
    if (c.counting) {
        if (certify) wideTraceKernel<true, CUR, true><<<traceBlocks, WIDE_TRACE_BLOCK, smem, stream>>>(mp, c.mesh, wv);
        else wideTraceKernel<true, CUR, false><<<traceBlocks, WIDE_TRACE_BLOCK, smem, stream>>>(mp, c.mesh, wv);
    } else {
        if (certify) wideTraceKernel<false, CUR, true><<<traceBlocks, WIDE_TRACE_BLOCK, smem, stream>>>(mp, c.mesh, wv);
        else wideTraceKernel<false, CUR, false><<<traceBlocks, WIDE_TRACE_BLOCK, smem, stream>>>(mp, c.mesh, wv);
    }
}

// One wavefront iteration on `stream`: trace (extend + shadow rays) -> shade (+ retire sample, + next camera ray).
static void launchMeshIteration(RendererContext& c, const MeshState& mp, cudaStream_t stream, int cur, int traceBlocks, int shadeBlocks,
                                cudaEvent_t* ev, int shadeThreads = WF_BLOCK) {
    if (ev) cudaEventRecord(ev[0], stream);
    if (useWideTree(c)) {
        if (cur) launchWideTrace<1>(c, mp, stream, c.wideTraceBlocks);
        else launchWideTrace<0>(c, mp, stream, c.wideTraceBlocks);
    } else if (c.counting) {
        if (cur) traceKernel<true, 1><<<traceBlocks, TRACE_BLOCK, 0, stream>>>(mp, c.mesh);
        else traceKernel<true, 0><<<traceBlocks, TRACE_BLOCK, 0, stream>>>(mp, c.mesh);
    } else {
        if (cur) traceKernel<false, 1><<<traceBlocks, TRACE_BLOCK, 0, stream>>>(mp, c.mesh);
        else traceKernel<false, 0><<<traceBlocks, TRACE_BLOCK, 0, stream>>>(mp, c.mesh);
    }
    if (ev) cudaEventRecord(ev[1], stream);
    {
        // threads per shade block (the grid keeps its thread count): the kernel appends per block behind two barriers, and the ncu
        // source page showed a fifth of its stall samples right after them; with 4 warps per barrier instead of 8 the frame takes
        // 273.4 ms instead of 279.0 (64 threads: 275.9). CRT_SHADE_BLOCK overrides.
        static const int shadeBlock = std::getenv("CRT_SHADE_BLOCK") ? std::atoi(std::getenv("CRT_SHADE_BLOCK")) : 128;
        if (shadeBlock > 0 && shadeBlock <= WF_BLOCK && shadeBlock % 32 == 0 && shadeThreads == WF_BLOCK) { shadeBlocks = shadeBlocks * WF_BLOCK / shadeBlock; shadeThreads = shadeBlock; }
    }
    meshShadeKernel<<<shadeBlocks, shadeThreads, 0, stream>>>(mp, shadeScene(c), c.cam, cur);
    if (ev) cudaEventRecord(ev[2], stream);
}

static cudaGraphExec_t captureMeshBatch(RendererContext& c, const MeshState& mp, cudaStream_t stream, int batch, int traceBlocks, int shadeBlocks,
                                        int shadeThreads = WF_BLOCK) {
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    CRT_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    for (int k = 0; k < batch; k++) launchMeshIteration(c, mp, stream, k & 1, traceBlocks, shadeBlocks, nullptr, shadeThreads);
    CRT_CHECK(cudaStreamEndCapture(stream, &graph));
    CRT_CHECK(cudaGraphInstantiate(&exec, graph, 0));
    CRT_CHECK(cudaGraphDestroy(graph));
    return exec;
}


void crtRunMesh(RendererContext& c, int ns, bool resume) {
    const unsigned int npix = (unsigned int)c.nx * (unsigned int)c.ny;
    int slotsPerPixel = c.opts.reserved[0] > 0 ? c.opts.reserved[0] : 1;
    if (ns % slotsPerPixel != 0 || resume) slotsPerPixel = 1;
    PhaseTimer pt;
    allocMeshPipeline(c, npix * (unsigned int)slotsPerPixel);
    pt.mark("run: state allocation");
    MeshState& mp = c.mp;
    mp.accum = c.wf.accum;
    mp.npix = npix;
    mp.nx = c.nx;
    mp.ny = c.ny;
    mp.samplesPerSlot = ns / slotsPerPixel;
    mp.slotsPerPixel = slotsPerPixel;
    mp.streamBase = c.opts.sampleStream;
    {
        const char* v = std::getenv("CRT_TILED_SLOTS"); // CRT_TILED_SLOTS=0: slots row by row (diagnostics)
        mp.tilesX = (c.nx % 8 == 0 && c.ny % 4 == 0 && !(v && v[0] == '0')) ? (unsigned int)c.nx / 8u : 0u;
    }
    mp.traceBudget = c.opts.reserved[1] > 0 ? c.opts.reserved[1] : TRACE_BUDGET;
    mp.traceMinActive = c.opts.reserved[2] > 0 ? c.opts.reserved[2] : (useWideTree(c) ? WIDE_TRACE_MIN_ACTIVE : TRACE_MIN_ACTIVE);
    if (!c.traceBlocks) {
        int perSM = 0;
        CRT_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, traceKernel<false, 0>, TRACE_BLOCK, 0));
        c.traceBlocks = c.numSMs * (perSM > 0 ? perSM : 1); // persistent: exactly one resident wave
    }
    if (useWideTree(c) && !c.wideTraceBlocks) {
        const size_t smem = (size_t)c.wide.stackDepth * WIDE_TRACE_BLOCK * sizeof(uint2) + WIDE_TRACE_EXTRA_SMEM;
        auto prep = [&](auto kernel) { CRT_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); };
        prep(wideTraceKernel<false, 0, true>); prep(wideTraceKernel<false, 1, true>); prep(wideTraceKernel<false, 0, false>); prep(wideTraceKernel<false, 1, false>);
        prep(wideTraceKernel<true, 0, true>); prep(wideTraceKernel<true, 1, true>); prep(wideTraceKernel<true, 0, false>); prep(wideTraceKernel<true, 1, false>);
        int perSM = 0;
        CRT_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, wideTraceKernel<false, 0, true>, WIDE_TRACE_BLOCK, smem));
        c.wideTraceBlocks = c.numSMs * (perSM > 0 ? perSM : 1);
        // Shared-memory carve-out: left to the driver. Measured on the benchmark frame (profiles/r02/carveout_sweep.txt): with the
        // driver's own choice per kernel the frame takes 357 ms at 13, 14, 16 or 20 stack entries per thread but 380-390 ms at
        // exactly 12 (6 KB per chaser block, 12 KB per trace block); stating one split for every kernel of the frame gives 363 ms at
        // best (the shade kernel then loses L1), too small a split 390-410 ms (one trace block per SM does not fit), too large 380 ms.
        // The stack is therefore sized to a multiple of 4 entries, at least 16 (initRenderer); CRT_WIDE_CARVEOUT=<percent> states a
        // split for the trace kernels (tuning knob).
        cudaFuncAttributes fa;
        CRT_CHECK(cudaFuncGetAttributes(&fa, wideTraceKernel<false, 0, true>));
        int pct = -1;
        if (std::getenv("CRT_WIDE_CARVEOUT")) pct = std::min(std::max(std::atoi(std::getenv("CRT_WIDE_CARVEOUT")), -1), 100);
        if (pct >= 0) {
            auto carve = [&](auto kernel) { CRT_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct)); };
            carve(wideTraceKernel<false, 0, true>); carve(wideTraceKernel<false, 1, true>); carve(wideTraceKernel<false, 0, false>); carve(wideTraceKernel<false, 1, false>);
            carve(wideTraceKernel<true, 0, true>); carve(wideTraceKernel<true, 1, true>); carve(wideTraceKernel<true, 0, false>); carve(wideTraceKernel<true, 1, false>);
        }
        if (std::getenv("CRT_TIMING")) std::fprintf(stderr, "[crt timing] wide trace: %d blocks/SM, %zu B shared per block, carve-out %d %%\n", perSM, (size_t)fa.sharedSizeBytes + smem, pct);
    }
    const int kernelsPerIteration = 2;
    cudaStream_t stream = c.stream;

    std::memset(&c.stats, 0, sizeof(c.stats));
    c.chaserRays = c.chaserShadowRays = c.chaserNodeVisits = c.chaserTriTests = 0;
    c.stats.samples = (unsigned long long)npix * (unsigned long long)(ns > 0 ? ns : 0);
    CRT_CHECK(cudaEventRecord(c.evStart, stream));
    if (!resume) {
        CRT_CHECK(cudaMemsetAsync(mp.accum, 0, (size_t)npix * sizeof(float4), stream));
        c.samplesDone = 0;
    }
    CRT_CHECK(cudaMemsetAsync(mp.ctl, 0, sizeof(MeshControl), stream));
    unsigned long long launches = 0;
    MeshControl* host = (MeshControl*)c.hostCtl;

    if (npix > 0 && ns > 0 && c.maxDepth > 0) {
        meshStartKernel<<<c.numSMs * 4, WF_BLOCK, 0, stream>>>(mp, c.cam, resume ? 1 : 0);
        launches += 1;
        CRT_CHECK(cudaGetLastError());

        int batch = c.opts.megaBatch > 0 ? c.opts.megaBatch : 16;
        batch = (batch + 1) & ~1; // even: the two queue sets swap roles every iteration
        const char* chaserEnv = std::getenv("CRT_CHASER"); // CRT_CHASER=0: the wavefront alone (diagnostics, tests)
        const bool lanes = g_profiling != 1 && !(chaserEnv && chaserEnv[0] == '0') && c.traceBlocks >= 2 * c.numSMs;
        if (g_profiling == 1) {
            cudaEvent_t ev[3];
            for (auto& e : ev) CRT_CHECK(cudaEventCreate(&e));
            c.stats.profiled = 1;
            const bool dump = std::getenv("CRT_DUMP_ITERATIONS") != nullptr; // per-iteration queue sizes and kernel times on stderr
            unsigned long long iter = 0;
            while (true) {
                for (int cur = 0; cur < 2; cur++) {
                    if (dump) CRT_CHECK(cudaMemcpyAsync(host, mp.ctl, sizeof(MeshControl), cudaMemcpyDeviceToHost, stream));
                    launchMeshIteration(c, mp, stream, cur, c.traceBlocks, c.numSMs * 4, ev);
                    launches += kernelsPerIteration;
                    CRT_CHECK(cudaStreamSynchronize(stream));
                    float msT, msS;
                    cudaEventElapsedTime(&msT, ev[0], ev[1]); c.stats.msTrace += msT;
                    cudaEventElapsedTime(&msS, ev[1], ev[2]); c.stats.msShade += msS;
                    if (dump) std::fprintf(stderr, "iter %llu trace %u deferred %u  trace_ms %.4f shade_ms %.4f\n", iter, host->traceCount[cur] + host->traceBack[cur],
                                           host->shadeCount[cur], msT, msS);
                    iter++;
                }
                CRT_CHECK(cudaMemcpyAsync(host, mp.ctl, sizeof(MeshControl), cudaMemcpyDeviceToHost, stream));
                CRT_CHECK(cudaStreamSynchronize(stream));
                if (host->traceCount[0] == 0 && host->traceBack[0] == 0 && host->shadeCount[0] == 0) break;
            }
            for (auto& e : ev) cudaEventDestroy(e);
        } else {
            // The wavefront leaves one trace-block slot per SM to the chaser (mesh_pipeline.cuh), which runs beside it for the whole
            // frame on its own stream. Between two batches of iterations the slots that fell behind are moved into the ring.
            auto envInt = [](const char* name, int dflt) { const char* v = std::getenv(name); return v ? std::atoi(v) : dflt; };
            const bool chase = lanes;
            const int waveBlocks = envInt("CRT_CHASE_WAVE", c.numSMs * 4);        // tuning knobs (defaults are the measured best)
            const int lastWaveBlocks = envInt("CRT_CHASE_LAST_WAVE", c.numSMs * 16);
            const int reserve = chase ? envInt("CRT_CHASE_RESERVE", 0) : 0;            // trace blocks per SM the wavefront does not ask for
            const int blocksA = c.traceBlocks - reserve * c.numSMs;
            const int shadeBlocksA = envInt("CRT_LANE_A_SHADE_BLOCKS", c.numSMs * 8);
            float lagFactor = (float)envInt("CRT_CHASE_LAG_PCT", 40) * 0.01f; // raised while the chaser has room (below)
            const float lagFactorMax = (float)envInt("CRT_CHASE_LAG_MAX_PCT", 90) * 0.01f;
            const float lagStep = (float)envInt("CRT_CHASE_LAG_STEP_PCT", 2) * 0.01f;
            const unsigned int target = (unsigned int)envInt("CRT_CHASE_TARGET", 4000);
            const float exclusiveFactor = (float)envInt("CRT_CHASE_EXCLUSIVE_PCT", 8) * 0.01f;
            const unsigned int exclusiveEvery = (unsigned int)envInt("CRT_CHASE_EXCLUSIVE_EVERY", 4); // every n-th warp of a wave serves the exclusive ring ...
            const unsigned int exclusivePairs = (1u << envInt("CRT_CHASE_EXCLUSIVE_SLOTS", 1)) - 1u;  // ... this many slots at a time
            const unsigned int moveAllBelow = (unsigned int)envInt("CRT_CHASE_MOVE_ALL", 24576);
            const unsigned int capacity = (unsigned int)envInt("CRT_CHASE_CAPACITY", 16384);
            const long long key = ((long long)mp.samplesPerSlot << 24) ^ ((long long)slotsPerPixel << 8) ^ batch ^
                                  ((long long)c.counting << 60) ^ ((long long)c.maxDepth << 40) ^ ((long long)mp.streamBase << 48) ^
                                  ((long long)mp.traceBudget << 12) ^ ((long long)mp.traceMinActive << 4) ^ ((long long)chase << 59) ^
                                  ((long long)blocksA << 30) ^ ((long long)shadeBlocksA << 17) ^ ((long long)(useWideTree(c) ? 1 + c.traversal : 0) << 56);
            // A captured graph carries its launch arguments BY VALUE (the whole MeshState with its queue and accumulator pointers,
            // the scene views, the camera), so it is re-used only when every one of those bytes is what it was at capture time --
            // a caller that changes the accumulator, the slot layout or an option between two runRenderer calls gets a new graph.
            std::vector<unsigned char> signature;
            {
                const ShadeScene sceneNow = shadeScene(c);
                const int scalars[6] = {batch, blocksA, shadeBlocksA, c.wideTraceBlocks, (int)c.counting, useWideTree(c) ? 1 + (int)c.traversal : 0};
                auto put = [&signature](const void* p, size_t n) { signature.insert(signature.end(), (const unsigned char*)p, (const unsigned char*)p + n); };
                put(&mp, sizeof(mp)); put(&c.mesh, sizeof(c.mesh)); put(&c.wide, sizeof(c.wide)); put(&sceneNow, sizeof(sceneNow)); put(&c.cam, sizeof(c.cam));
                put(scalars, sizeof(scalars));
            }
            if (!c.graphExec || c.graphKey != key || c.graphSignature != signature) {
                if (c.graphExec) { cudaGraphExecDestroy(c.graphExec); c.graphExec = nullptr; }
                c.graphExec = captureMeshBatch(c, mp, stream, batch, blocksA, shadeBlocksA);
                c.graphKey = key;
                c.graphSignature.swap(signature);
                pt.mark("run: graph capture");
            }
            const ChaseRing ring = c.ring;
            const ShadeScene scene = shadeScene(c);
            int wave = 0;
            auto launchWave = [&](int blocks) { // after a hand-over: the wave starts once the commit kernel of `stream` has run
                // a stream whose previous wave has finished, if there is one (a wave can run for most of the frame: the next wave
                // must not queue behind it); else round robin
                cudaStream_t on = g_cache.chaseStreams[wave % CHASE_STREAMS];
                for (int k = 0; k < CHASE_STREAMS; k++) {
                    cudaStream_t candidate = g_cache.chaseStreams[(wave + k) % CHASE_STREAMS];
                    if (cudaStreamQuery(candidate) == cudaSuccess) { on = candidate; break; }
                }
                cudaGetLastError(); // (cudaErrorNotReady of a busy stream is not an error)
                wave++;
                CRT_CHECK(cudaEventRecord(c.evLane, stream));
                CRT_CHECK(cudaStreamWaitEvent(on, c.evLane, 0));
                if (useWideTree(c)) {
                    const size_t smem = (size_t)c.wide.stackDepth * CHASE_BLOCK * sizeof(uint2);
                    if (c.counting) chaseKernel<true, true><<<blocks, CHASE_BLOCK, smem, on>>>(mp, c.mesh, c.wide, scene, c.cam, ring, exclusiveEvery, exclusivePairs);
                    else chaseKernel<false, true><<<blocks, CHASE_BLOCK, smem, on>>>(mp, c.mesh, c.wide, scene, c.cam, ring, exclusiveEvery, exclusivePairs);
                } else {
                    if (c.counting) chaseKernel<true, false><<<blocks, CHASE_BLOCK, 0, on>>>(mp, c.mesh, c.wide, scene, c.cam, ring, exclusiveEvery, exclusivePairs);
                    else chaseKernel<false, false><<<blocks, CHASE_BLOCK, 0, on>>>(mp, c.mesh, c.wide, scene, c.cam, ring, exclusiveEvery, exclusivePairs);
                }
                CRT_CHECK(cudaGetLastError());
                launches += 1;
            };
            if (chase) {
                CRT_CHECK(cudaMemsetAsync(ring.ctl, 0, 16 * sizeof(unsigned int), stream));
                CRT_CHECK(cudaMemsetAsync(ring.counters, 0, 8 * sizeof(unsigned long long), stream));
                CRT_CHECK(cudaMemsetAsync(c.laneSums, 0, 8 * sizeof(unsigned long long), stream));
            }
            const bool dumpLanes = std::getenv("CRT_DUMP_LANES") != nullptr;
            std::vector<cudaEvent_t> profEvents;
            unsigned int* hostRing = (unsigned int*)c.hostCtlFast;
            const auto t0 = std::chrono::steady_clock::now();
            bool last = false;
            while (true) {
                if (chase) { // the wavefront is idle here: move lagging slots from its input queues to the rings
                    CRT_CHECK(cudaMemsetAsync(c.laneSums, 0, 4 * sizeof(unsigned long long), stream));
                    laneStatsKernel<<<c.numSMs * 2, WF_BLOCK, 0, stream>>>(mp, ring, c.laneSums, lagFactor);
                    lanePartitionKernel<<<c.numSMs * 2, WF_BLOCK, 0, stream>>>(mp, ring, c.laneSums, lagFactor, exclusiveFactor, 6, moveAllBelow, capacity,
                                                                             (unsigned int)wave * 0x9E3779B9u);
                    laneCopyBackKernel<<<c.numSMs * 2, WF_BLOCK, 0, stream>>>(mp);
                    laneCommitKernel<<<1, 1, 0, stream>>>(mp.ctl, ring, c.laneSums);
                    launches += 4;
                    launchWave(last ? lastWaveBlocks : waveBlocks);
                }
                if (last) break;
                if (g_profiling == 2) { // the timed configuration, launch by launch with events around every kernel (no graph, no extra sync)
                    if (profEvents.empty()) {
                        profEvents.resize(3 * (size_t)batch);
                        for (auto& e : profEvents) CRT_CHECK(cudaEventCreate(&e));
                        c.stats.profiled = 2;
                    }
                    for (int k = 0; k < batch; k++) launchMeshIteration(c, mp, stream, k & 1, blocksA, shadeBlocksA, &profEvents[3 * (size_t)k]);
                } else {
                    CRT_CHECK(cudaGraphLaunch(c.graphExec, stream));
                }
                launches += (unsigned long long)batch * kernelsPerIteration;
                CRT_CHECK(cudaMemcpyAsync(host, mp.ctl, sizeof(MeshControl), cudaMemcpyDeviceToHost, stream));
                if (chase) CRT_CHECK(cudaMemcpyAsync(hostRing, ring.ctl, 16 * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
                CRT_CHECK(cudaStreamSynchronize(stream));
                if (g_profiling == 2) {
                    for (int k = 0; k < batch; k++) {
                        float msT = 0.0f, msS = 0.0f;
                        cudaEventElapsedTime(&msT, profEvents[3 * (size_t)k], profEvents[3 * (size_t)k + 1]);
                        cudaEventElapsedTime(&msS, profEvents[3 * (size_t)k + 1], profEvents[3 * (size_t)k + 2]);
                        c.stats.msTrace += msT;
                        c.stats.msShade += msS;
                    }
                }
                const unsigned int live = host->traceCount[0] + host->traceBack[0] + host->shadeCount[0];
                if (dumpLanes) {
                    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                    std::fprintf(stderr, "lanes t=%.2f ms  wavefront: trace %u shade %u iters %llu   chaser shared: %u/%u finished %u  exclusive: %u/%u finished %u  lag %.2f\n",
                                 ms, host->traceCount[0] + host->traceBack[0], host->shadeCount[0], host->iterations, chase ? hostRing[0] : 0u, chase ? hostRing[1] : 0u,
                                 chase ? hostRing[4] : 0u, chase ? hostRing[8] : 0u, chase ? hostRing[9] : 0u, chase ? hostRing[12] : 0u, lagFactor);
                }
                if (live == 0) break;
                if (chase) { // the chaser has room: hand over slots that lag less (they would otherwise stretch the end of the frame)
                    const unsigned int held = (hostRing[2] - hostRing[4]) + (hostRing[10] - hostRing[12]);
                    if (held < target && lagFactor < lagFactorMax) lagFactor += lagStep;
                }
                last = chase && live <= moveAllBelow; // the next hand-over takes everything that is left: no iteration follows it
            }
            for (auto& e : profEvents) cudaEventDestroy(e);
            if (chase) { // wait for the chaser's waves, collect their ray counts
                for (int k = 0; k < (wave < CHASE_STREAMS ? wave : CHASE_STREAMS); k++) CRT_CHECK(cudaStreamSynchronize(g_cache.chaseStreams[k]));
                CRT_CHECK(cudaMemcpyAsync(hostRing, ring.counters, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
                CRT_CHECK(cudaStreamSynchronize(stream));
                const unsigned long long* hc = (const unsigned long long*)hostRing;
                host->raysExtend += hc[0];
                host->raysShadow += hc[1];
                host->nodeVisits += hc[2];
                host->triTests += hc[3];
                c.chaserRays = hc[0];
                c.chaserShadowRays = hc[1];
                c.chaserNodeVisits = hc[2];
                c.chaserTriTests = hc[3];
                host->redone += hc[4];
                if (dumpLanes) {
                    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                    std::fprintf(stderr, "lanes t=%.2f ms  chaser done: %llu extend + %llu shadow rays in %d waves\n", ms, hc[0], hc[1], wave);
                }
            }
        }
    } else {
        CRT_CHECK(cudaMemcpyAsync(host, mp.ctl, sizeof(MeshControl), cudaMemcpyDeviceToHost, stream));
    }

    if (npix > 0 && ns > 0 && c.maxDepth > 0) c.samplesDone += ns;
    if (!c.opts.deferFinalize) {
        finalizeFrameTo(c, mp.accum, float(resume ? c.samplesDone : ns), stream);
        launches += 1;
    }
    CRT_CHECK(cudaEventRecord(c.evStop, stream));
    CRT_CHECK(cudaStreamSynchronize(stream));
    CRT_CHECK(cudaGetLastError());
    CRT_CHECK(cudaEventElapsedTime(&c.stats.msTotal, c.evStart, c.evStop));
    pt.mark("run: frame");
    c.stats.raysExtend = host->raysExtend;
    c.stats.raysShadow = host->raysShadow;
    c.stats.iterations = host->iterations;
    c.stats.resumes = host->resumes;
    c.stats.deferred = host->deferred;
    c.lastFrameRedo = host->redone;
    c.stats.kernelLaunches = launches;
    c.lastNodeVisits = host->nodeVisits;
    c.lastTriTests = host->triTests;
}

extern "C" void runRenderer(int ns, int tx, int ty) {
    (void)tx; (void)ty; // launch-shape hints of the megakernel; the wavefront kernels size themselves
    RendererContext& c = g_ctx;
    if (!c.initialised) {
        std::fprintf(stderr, "runRenderer called before initRenderer\n");
        std::exit(99);
    }
    if (c.kind == SCENE_MESH && crtMultiGpuCount() > 1) crtMultiGpuRun(ns, [](int own) { crtRunMesh(g_ctx, own, false); });
    else if (c.kind == SCENE_MESH) crtRunMesh(c, ns, false);
    else crtRunSpheres(c, ns);
}

// ------------------------------------------------------------ progressive --
// SURVEY.md 8f rank 3 (the reference's wish list, TODO.txt:70): because a pixel's samples are one RNG stream, a frame can be
// continued exactly -- runRenderer(a) then continueRenderer(b) gives the bits of runRenderer(a + b) -- and the state that
// makes this possible (running sums + stream positions) can be written to a file and read back into a fresh initRenderer.
extern "C" int continueRenderer(int nsMore, int tx, int ty) {
    (void)tx; (void)ty;
    RendererContext& c = g_ctx;
    if (!c.initialised || c.kind != SCENE_MESH || c.samplesDone <= 0 || !c.mp.rngOut || c.mp.slotsPerPixel != 1 || crtMultiGpuCount() > 1) return -1;
    crtRunMesh(c, nsMore, true);
    return 0;
}

extern "C" int getRendererSamplesDone() { return g_ctx.initialised ? g_ctx.samplesDone : 0; }

static unsigned int hostSlotPixel(const MeshState& st, unsigned int slot) { // slotPixel() of mesh_pipeline.cuh on the host
    const unsigned int s = slot % st.npix;
    if (st.tilesX == 0u) return s;
    const unsigned int tile = s >> 5, o = s & 31u;
    const unsigned int ty = tile / st.tilesX, tx = tile - ty * st.tilesX;
    return (ty * 4u + (o >> 3)) * (unsigned int)st.nx + tx * 8u + (o & 7u);
}

struct CheckpointHeader {
    char magic[8]; // "CRTCKP01"
    int nx, ny, samplesDone;
    unsigned int stream;
};

extern "C" int saveRendererCheckpoint(const char* path) {
    RendererContext& c = g_ctx;
    if (!c.initialised || c.kind != SCENE_MESH || c.samplesDone <= 0 || !c.mp.rngOut || c.mp.slotsPerPixel != 1) return -1;
    const size_t npix = (size_t)c.nx * c.ny;
    std::vector<float4> sums(npix);
    std::vector<unsigned int> rng(npix);
    CRT_CHECK(cudaMemcpy(sums.data(), c.wf.accum, npix * sizeof(float4), cudaMemcpyDeviceToHost));
    CRT_CHECK(cudaMemcpy(rng.data(), c.mp.rngOut, npix * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    if (c.mp.tilesX) { // the file is in pixel order, rngOut in slot order
        std::vector<unsigned int> byPixel(npix);
        for (size_t sidx = 0; sidx < npix; sidx++) byPixel[hostSlotPixel(c.mp, (unsigned int)sidx)] = rng[sidx];
        rng.swap(byPixel);
    }
    CheckpointHeader h;
    std::memcpy(h.magic, "CRTCKP01", 8);
    h.nx = c.nx; h.ny = c.ny; h.samplesDone = c.samplesDone; h.stream = c.opts.sampleStream;
    FILE* f = std::fopen(path, "wb");
    if (!f) return -2;
    const bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 && std::fwrite(sums.data(), sizeof(float4), npix, f) == npix &&
                    std::fwrite(rng.data(), sizeof(unsigned int), npix, f) == npix;
    std::fclose(f);
    return ok ? 0 : -3;
}

// After initRenderer of the same scene, size and sample stream: restores sums and stream positions; continueRenderer() goes on.
extern "C" int loadRendererCheckpoint(const char* path) {
    RendererContext& c = g_ctx;
    if (!c.initialised || c.kind != SCENE_MESH) return -1;
    FILE* f = std::fopen(path, "rb");
    if (!f) return -2;
    CheckpointHeader h;
    const size_t npix = (size_t)c.nx * c.ny;
    std::vector<float4> sums(npix);
    std::vector<unsigned int> rng(npix);
    const bool ok = std::fread(&h, sizeof(h), 1, f) == 1 && std::memcmp(h.magic, "CRTCKP01", 8) == 0 && h.nx == c.nx && h.ny == c.ny &&
                    h.stream == c.opts.sampleStream && h.samplesDone > 0 && std::fread(sums.data(), sizeof(float4), npix, f) == npix &&
                    std::fread(rng.data(), sizeof(unsigned int), npix, f) == npix;
    std::fclose(f);
    if (!ok) return -3;
    allocMeshPipeline(c, (unsigned int)npix);
    c.mp.slotsPerPixel = 1;
    c.mp.npix = (unsigned int)npix;
    c.mp.nx = c.nx;
    {
        const char* v = std::getenv("CRT_TILED_SLOTS");
        c.mp.tilesX = (c.nx % 8 == 0 && c.ny % 4 == 0 && !(v && v[0] == '0')) ? (unsigned int)c.nx / 8u : 0u;
    }
    if (c.mp.tilesX) {
        std::vector<unsigned int> bySlot(npix);
        for (size_t sidx = 0; sidx < npix; sidx++) bySlot[sidx] = rng[hostSlotPixel(c.mp, (unsigned int)sidx)];
        rng.swap(bySlot);
    }
    CRT_CHECK(cudaMemcpy(c.wf.accum, sums.data(), npix * sizeof(float4), cudaMemcpyHostToDevice));
    CRT_CHECK(cudaMemcpy(c.mp.rngOut, rng.data(), npix * sizeof(unsigned int), cudaMemcpyHostToDevice));
    c.samplesDone = h.samplesDone;
    return 0;
}

extern "C" void finalizeFrame(int nsTotal) {
    RendererContext& c = g_ctx;
    if (!c.initialised) return;
    finalizeFrameTo(c, c.wf.accum, float(nsTotal), c.stream);
    CRT_CHECK(cudaGetLastError());
    CRT_CHECK(cudaStreamSynchronize(c.stream));
}

extern "C" void* getRendererAccumDevice() { return g_ctx.initialised ? (void*)g_ctx.wf.accum : nullptr; }

extern "C" void setRendererAccumDevice(void* dAccum) {
    RendererContext& c = g_ctx;
    if (!c.initialised || !dAccum) return;
    c.wf.accum = (float4*)dAccum;
    c.ownsAccum = false;
    // the captured graph bakes the accumulator pointer into its kernel nodes (MeshState by value): capture again
    if (c.graphExec) { cudaGraphExecDestroy(c.graphExec); c.graphExec = nullptr; }
    c.graphKey = -1;
}

extern "C" void getRendererStats(renderer_stats* out) {
    if (out) *out = g_ctx.stats;
}

// Test/diagnostic hook: copies one of the mesh pipeline's per-slot arrays to the host (returns bytes copied, 0 if unknown).
extern "C" size_t rendererDebugRead(const char* name, void* dst, size_t maxBytes) {
    RendererContext& c = g_ctx;
    if (!c.initialised || !c.mp.rayO) return 0;
    const MeshState& m = c.mp;
    const void* src = nullptr;
    size_t bytes = (size_t)m.numSlots * sizeof(float4);
    const std::string n(name);
    if (n == "rayO") src = m.rayO; else if (n == "rayD") src = m.rayD; else if (n == "atten") src = m.atten;
    else if (n == "pcol") src = m.pcol; else if (n == "hit") src = m.hit; else if (n == "shO") src = m.shO;
    else if (n == "shD") src = m.shD; else if (n == "shL") src = m.shL; else if (n == "shC") src = m.shC;
    else if (n == "accum") { src = m.accum; bytes = (size_t)m.npix * sizeof(float4); }
    if (!src) return 0;
    if (bytes > maxBytes) bytes = maxBytes;
    CRT_CHECK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return bytes;
}

extern "C" void setRendererCounting(int on) {
    g_ctx.counting = on != 0;
}

extern "C" void getRendererTraversalCounts(unsigned long long* nodeVisits, unsigned long long* triTests) {
    if (nodeVisits) *nodeVisits = g_ctx.lastNodeVisits;
    if (triTests) *triTests = g_ctx.lastTriTests;
}

extern "C" void getRendererWideInfo(renderer_wide_info* out) {
    if (!out) return;
    const RendererContext& c = g_ctx;
    out->active = c.initialised && c.wide.stackDepth > 0 && c.traversal != TRAVERSAL_EXACT;
    out->traversal = c.traversal;
    out->numNodes = c.wideStats.numNodes;
    out->numTriangles = c.wideStats.numTris;
    out->depth = c.wideStats.maxDepth;
    out->buildThreads = c.wideStats.threads;
    out->buildMs = (float)c.wideStats.msTotal;
    out->sahCost = (float)c.wideStats.sahCost;
    out->lastBatchRedo = c.lastBatchRedo;
    out->lastFrameRedo = c.lastFrameRedo;
}

// Test hook: copies the wide tree in use to the host (96-byte node records, csrc/wide_bvh.h; caller's slot index per leaf
// triangle). Returns the number of nodes (0 = no wide tree); copies at most the given capacities.
extern "C" unsigned int getRendererWideTree(void* nodes, unsigned int nodeCapacity, unsigned int* triOrig, unsigned int triCapacity) {
    const RendererContext& c = g_ctx;
    if (!c.initialised || !c.wideNodesDev) return 0;
    const unsigned int nn = std::min(nodeCapacity, c.wideStats.numNodes), nt = std::min(triCapacity, c.wideStats.numTris);
    if (nodes && nn) CRT_CHECK(cudaMemcpy(nodes, c.wideNodesDev, (size_t)nn * sizeof(WideNode), cudaMemcpyDeviceToHost));
    if (triOrig && nt) CRT_CHECK(cudaMemcpy(triOrig, c.wideTriOrigDev, (size_t)nt * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    return c.wideStats.numNodes;
}

extern "C" void getRendererChaserCounts(unsigned long long* raysExtend, unsigned long long* raysShadow, unsigned long long* nodeVisits,
                                        unsigned long long* triTests) {
    if (raysExtend) *raysExtend = g_ctx.chaserRays;
    if (raysShadow) *raysShadow = g_ctx.chaserShadowRays;
    if (nodeVisits) *nodeVisits = g_ctx.chaserNodeVisits;
    if (triTests) *triTests = g_ctx.chaserTriTests;
}

extern "C" void cleanupRenderer() {
    RendererContext& c = g_ctx;
    if (!c.initialised) return;
    PhaseTimer pt;
    if (crtMultiGpuCount() > 1) crtMultiGpuStop();
    CRT_CHECK(cudaStreamSynchronize(c.stream));
    CRT_CHECK(cudaStreamSynchronize(c.streamFast));
    freeWavefront(c);
    freeMeshPipeline(c);
    c.wf.accum = nullptr;
    c.texPtrHost.clear();
    arenaReset(); // every device pointer of the frame dies here; the memory stays with the process for the next frame
    const bool reset = c.opts.resetDeviceOnCleanup != 0;
    if (c.batchRedo) cudaFree(c.batchRedo);
    c = RendererContext();
    pt.mark("cleanup");
    if (reset) { // kernels.cu:679
        releaseCaches();
        cudaDeviceReset();
    }
}

// ------------------------------------------------------------- ray batches --
extern "C" float intersectBatchDeviceEx(const void* dRayO, const void* dRayD, long long n, void* dHit, int* dMeshId, int anyHit);
extern "C" float intersectBatchDevice(const void* dRayO, const void* dRayD, long long n, void* dHit, int* dMeshId) {
    return intersectBatchDeviceEx(dRayO, dRayD, n, dHit, dMeshId, 0);
}

extern "C" float intersectBatchDeviceEx(const void* dRayO, const void* dRayD, long long n, void* dHit, int* dMeshId, int anyHit) {
    RendererContext& c = g_ctx;
    if (!c.initialised || c.kind != SCENE_MESH) {
        std::fprintf(stderr, "intersectBatchDevice needs a mesh scene (initRenderer)\n");
        std::exit(99);
    }
    unsigned long long* cursor = c.batchScratch;      // [0] cursor of the first kernel, [1] of the second
    unsigned long long* counts = c.batchScratch + 2;  // [2], [3] visits / tests, [4] rays left to the exact kernel
    CRT_CHECK(cudaMemsetAsync(c.batchScratch, 0, 5 * sizeof(unsigned long long), c.stream));
    const bool useWide = c.wide.stackDepth > 0 && c.traversal != TRAVERSAL_EXACT;
    if (useWide && c.batchRedoCap < (size_t)n) {
        if (c.batchRedo) CRT_CHECK(cudaFree(c.batchRedo)); // per-call scratch: plain allocation, not the frame's arena
        CRT_CHECK(cudaMalloc((void**)&c.batchRedo, (size_t)n * sizeof(unsigned int)));
        c.batchRedoCap = (size_t)n;
    }
    const int blocks = c.numSMs * 8;
    CRT_CHECK(cudaEventRecord(c.evStart, c.stream));
    const float4* rayO = (const float4*)dRayO;
    const float4* rayD = (const float4*)dRayD;
    if (useWide) {
        const size_t smem = (size_t)c.wide.stackDepth * WIDE_BATCH_BLOCK * sizeof(uint2);
        const bool certify = c.traversal != TRAVERSAL_WIDE_UNCERTIFIED;
        auto launch = [&](auto kernel) {
            CRT_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int perSM = 1;
            CRT_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, WIDE_BATCH_BLOCK, smem));
            kernel<<<c.numSMs * (perSM > 0 ? perSM : 1), WIDE_BATCH_BLOCK, smem, c.stream>>>(c.mesh, c.wide, c.triShade, rayO, rayD, (unsigned long long)n, (float4*)dHit,
                                                                                dMeshId, cursor, counts, c.batchRedo, c.batchScratch + 4, anyHit);
        };
        if (c.counting) { if (certify) launch(wideIntersectBatchKernel<true, true>); else launch(wideIntersectBatchKernel<true, false>); }
        else { if (certify) launch(wideIntersectBatchKernel<false, true>); else launch(wideIntersectBatchKernel<false, false>); }
        // the rays the certificate does not cover, in the reference's visiting order (device-side count: no host round trip)
        intersectBatchKernel<false><<<c.numSMs, WF_BLOCK, 0, c.stream>>>(c.mesh, c.triShade, rayO, rayD, 0ull, (float4*)dHit, dMeshId, cursor + 1, counts,
                                                                        c.batchRedo, c.batchScratch + 4, anyHit);
    } else if (c.counting) {
        intersectBatchKernel<true><<<blocks, WF_BLOCK, 0, c.stream>>>(c.mesh, c.triShade, rayO, rayD, (unsigned long long)n, (float4*)dHit, dMeshId, cursor,
                                                                    counts, nullptr, nullptr, anyHit);
    } else {
        intersectBatchKernel<false><<<blocks, WF_BLOCK, 0, c.stream>>>(c.mesh, c.triShade, rayO, rayD, (unsigned long long)n, (float4*)dHit, dMeshId, cursor,
                                                                     counts, nullptr, nullptr, anyHit);
    }
    CRT_CHECK(cudaEventRecord(c.evStop, c.stream));
    CRT_CHECK(cudaGetLastError());
    CRT_CHECK(cudaStreamSynchronize(c.stream));
    float ms = 0.0f;
    CRT_CHECK(cudaEventElapsedTime(&ms, c.evStart, c.evStop));
    unsigned long long h[3];
    CRT_CHECK(cudaMemcpy(h, counts, sizeof(h), cudaMemcpyDeviceToHost));
    c.lastNodeVisits = h[0];
    c.lastTriTests = h[1];
    c.lastBatchRedo = h[2];
    return ms;
}

extern "C" void generateRayBatchDevice(void* dRayO, void* dRayD, long long n, int filmW, int filmH, float tMin, float tMax) {
    RendererContext& c = g_ctx;
    if (!c.initialised) {
        std::fprintf(stderr, "generateRayBatchDevice called before initRenderer\n");
        std::exit(99);
    }
    const unsigned int blocks = (unsigned int)((n + 255) / 256);
    if (n > 0)
        generateRayBatchKernel<<<blocks, 256, 0, c.stream>>>((float4*)dRayO, (float4*)dRayD, (unsigned long long)n, c.cam, c.mesh.boundsMin,
                                                            c.mesh.boundsMax, filmW, filmH, tMin, tMax);
    CRT_CHECK(cudaGetLastError());
    CRT_CHECK(cudaStreamSynchronize(c.stream));
}

extern "C" void intersectBatch(const float* origins, const float* dirs, long long n, float tMin, float tMax, float* outT, int* outTriId,
                               int* outMeshId) {
    if (n <= 0) return;
    std::vector<float4> o((size_t)n), d((size_t)n);
    for (long long i = 0; i < n; i++) {
        o[i] = make_float4(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2], tMin);
        d[i] = make_float4(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], tMax);
    }
    float4 *dO = nullptr, *dD = nullptr, *dH = nullptr; // per-call buffers: plain allocations, not the frame's arena
    int* dM = nullptr;
    CRT_CHECK(cudaMalloc((void**)&dO, (size_t)n * sizeof(float4)));
    CRT_CHECK(cudaMalloc((void**)&dD, (size_t)n * sizeof(float4)));
    CRT_CHECK(cudaMalloc((void**)&dH, (size_t)n * sizeof(float4)));
    CRT_CHECK(cudaMalloc((void**)&dM, (size_t)n * sizeof(int)));
    hostUpload(g_cache, dO, o.data(), (size_t)n * sizeof(float4));
    hostUpload(g_cache, dD, d.data(), (size_t)n * sizeof(float4));
    CRT_CHECK(cudaDeviceSynchronize()); // (the query runs on a non-blocking stream)
    intersectBatchDevice(dO, dD, n, dH, dM);
    std::vector<float4> h((size_t)n);
    hostDownload(g_cache, h.data(), dH, (size_t)n * sizeof(float4));
    if (outMeshId) hostDownload(g_cache, outMeshId, dM, (size_t)n * sizeof(int));
    for (long long i = 0; i < n; i++) {
        if (outT) outT[i] = h[i].x;
        if (outTriId) std::memcpy(&outTriId[i], &h[i].w, 4);
    }
    cudaFree(dO); cudaFree(dD); cudaFree(dH); cudaFree(dM);
}

// Host pointers in and out (12 floats per item each way, layout at scatterBatchKernel). Needs no scene.
extern "C" int scatterBatch(int preset, long long n, const float* in, float* out) {
    if (preset < 0 || preset >= PRESET_COUNT || n < 0) return -1;
    if (n == 0) return 0;
    float4 *dIn = nullptr, *dOut = nullptr;
    CRT_CHECK(cudaMalloc((void**)&dIn, (size_t)n * 3 * sizeof(float4)));
    CRT_CHECK(cudaMalloc((void**)&dOut, (size_t)n * 3 * sizeof(float4)));
    CRT_CHECK(cudaMemcpy(dIn, in, (size_t)n * 3 * sizeof(float4), cudaMemcpyHostToDevice));
    scatterBatchKernel<<<(unsigned int)((n + 127) / 128), 128>>>(preset, n, dIn, dOut);
    CRT_CHECK(cudaGetLastError());
    CRT_CHECK(cudaMemcpy(out, dOut, (size_t)n * 3 * sizeof(float4), cudaMemcpyDeviceToHost));
    cudaFree(dIn);
    cudaFree(dOut);
    return 0;
}

// sinCosSmall (wavefront_kernels.cuh) against the math library's sinf / cosf for every float in [0, 2*pi]: returns the number of
// arguments for which a bit differs (expected 0), or -1 without a device.
extern "C" long long rendererTrigSelfTest() {
    unsigned long long* d = nullptr;
    if (cudaMalloc((void**)&d, sizeof(*d)) != cudaSuccess) return -1;
    CRT_CHECK(cudaMemset(d, 0, sizeof(*d)));
    const float twoPi = 6.2831855f;
    unsigned int last;
    std::memcpy(&last, &twoPi, 4);
    trigSelfTestKernel<<<1184, 256>>>(0u, last + 1u, d);
    CRT_CHECK(cudaGetLastError());
    unsigned long long h = 0;
    CRT_CHECK(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(d);
    return (long long)h;
}

extern "C" void* rendererDeviceAlloc(size_t bytes) {
    void* p = nullptr;
    CRT_CHECK(cudaMalloc(&p, bytes ? bytes : 16));
    return p;
}
extern "C" void rendererDeviceFree(void* p) { cudaFree(p); }
extern "C" void rendererCopyToHost(void* dst, const void* dSrc, size_t bytes) {
    CRT_CHECK(cudaDeviceSynchronize()); // (what a pageable cudaMemcpy on the legacy stream implied for the caller's earlier work)
    hostDownload(g_cache, dst, dSrc, bytes);
}
extern "C" void rendererCopyToDevice(void* dDst, const void* src, size_t bytes) {
    hostUpload(g_cache, dDst, src, bytes);
}

#include "spheres_path.cuh"
#include "multi_gpu.cu.inc"
