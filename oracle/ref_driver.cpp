// ref_driver.cpp -- TEST INFRASTRUCTURE. A host program written against the REFERENCE's headers
// (kernels.h -> helper_structs.h -> vec3.h, found through -I/root/reference) that drives the three C entry
// points the way the reference's main.cpp:62-139 does, on a scene produced by libcrt_host.so.
// It is linked three ways (oracle/Makefile):
//   oracle/_ref/ref_driver        + oracle/_ref/libref.so        the unmodified reference kernel (the parity oracle,
//                                                                and bench.py --impl reference)
//   oracle/_ref/ref_shim_driver   + oracle/_ref/libref_shim.so   same + the hitMesh ray-batch probe, + the
//                                                                reference-derived sphere megakernel
//   oracle/_ref/dropin_driver     + build/libcrt_b200.so         the SAME object code against OUR library: the
//                                                                drop-in proof (nothing but the link line changes)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "kernels.h"

extern "C" {
// libcrt_host.so (cuda-raytracing-optimized_b200/host/host_api.h), declared with the reference's types
void* crtSceneCreateStaircase(float detail, int texSize, int primsPerLeaf);
void* crtSceneLoadBVH(const char* path, int texSize);
const kernel_scene* crtSceneKernelScene(const void* s);
void crtStaircaseCamera(int nx, int ny, camera* out);
int crtRtiowScene(unsigned int seed, sphere* outSpheres, material* outMaterials, int cap);
void crtRtiowCamera(int nx, int ny, camera* out);
int crtWriteRef(const char* path, int nx, int ny, const vec3* fb);
#ifdef WITH_SHIM
float refIntersectBatch(const float* rayO, const float* rayD, long long n, int isShadow, float* outHit, int* outMesh);
int refScatterBatch(int preset, long long n, const float* in, float* out);
void refSpheresInit(const sphere* spheres, const material* materials, int n, const camera cam, vec3** fb, int nx, int ny, int maxDepth);
void refSpheresRun(int ns, int tx, int ty);
void refSpheresCleanup();
#endif
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int usage() {
    std::fprintf(stderr,
                 "usage:\n"
                 "  render_e2e <same arguments as render>   every step = initRenderer + runRenderer + frame read\n"
                 "  render  <detail|file.bvh> <texSize> <primsPerLeaf> <nx> <ny> <ns> <maxDepth> <warmup> <steps> <out.ref|->\n"
                 "  spheres <seed> <nx> <ny> <ns> <maxDepth> <warmup> <steps> <out.ref|->      (shim build only)\n"
                 "  batch   <detail|file.bvh> <texSize> <primsPerLeaf> <rays.bin> <isShadow> <hits.bin> (shim build only)\n");
    return 2;
}

static void* makeScene(const char* spec, int texSize, int ppl) {
    const size_t n = std::strlen(spec);
    if (n > 4 && std::strcmp(spec + n - 4, ".bvh") == 0) return crtSceneLoadBVH(spec, texSize);
    return crtSceneCreateStaircase((float)std::atof(spec), texSize, ppl);
}

int main(int argc, char** argv) {
    if (argc < 2) return usage();
    const std::string mode = argv[1];
    if (mode == "render_e2e" && argc == 12) {
        // every step = what a caller of the 3 entry points pays from host buffers to a host-readable frame:
        // initRenderer (scene upload) + runRenderer + reading the frame; cleanupRenderer (a cudaDeviceReset in the
        // reference, kernels.cu:679) is left outside the timed region for both libraries.
        const int texSize = std::atoi(argv[3]), ppl = std::atoi(argv[4]);
        const int nx = std::atoi(argv[5]), ny = std::atoi(argv[6]), ns = std::atoi(argv[7]), maxDepth = std::atoi(argv[8]);
        const int warmup = std::atoi(argv[9]), steps = std::atoi(argv[10]);
        void* scene = makeScene(argv[2], texSize, ppl);
        if (!scene) { std::fprintf(stderr, "scene creation failed\n"); return 1; }
        const kernel_scene ksc = *crtSceneKernelScene(scene);
        camera cam;
        crtStaircaseCamera(nx, ny, &cam);
        std::vector<vec3> host((size_t)nx * ny);
        std::printf("{\"mode\": \"render_e2e\", \"nx\": %d, \"ny\": %d, \"ns\": %d, \"maxDepth\": %d, \"ms\": [", nx, ny, ns, maxDepth);
        for (int i = 0; i < warmup + steps; i++) {
            vec3* fb = nullptr;
            const double t0 = now();
            initRenderer(ksc, cam, &fb, nx, ny, maxDepth);
            runRenderer(ns, 8, 8);
            std::memcpy(host.data(), fb, host.size() * sizeof(vec3));
            const double t1 = now();
            if (i >= warmup) std::printf("%s%.4f", i > warmup ? ", " : "", (t1 - t0) * 1e3);
            if (i == warmup + steps - 1 && std::strcmp(argv[11], "-") != 0) crtWriteRef(argv[11], nx, ny, host.data());
            cleanupRenderer();
        }
        std::printf("]}\n");
        return 0;
    }
    if (mode == "render" && argc == 12) {
        const int texSize = std::atoi(argv[3]), ppl = std::atoi(argv[4]);
        const int nx = std::atoi(argv[5]), ny = std::atoi(argv[6]), ns = std::atoi(argv[7]), maxDepth = std::atoi(argv[8]);
        const int warmup = std::atoi(argv[9]), steps = std::atoi(argv[10]);
        void* scene = makeScene(argv[2], texSize, ppl);
        if (!scene) { std::fprintf(stderr, "scene creation failed\n"); return 1; }
        const kernel_scene ksc = *crtSceneKernelScene(scene);
        camera cam;
        crtStaircaseCamera(nx, ny, &cam); // setup_camera, staircase_scene.h:62
        vec3* fb = nullptr;
        initRenderer(ksc, cam, &fb, nx, ny, maxDepth); // main.cpp:94
        for (int i = 0; i < warmup; i++) runRenderer(ns, 8, 8);
        std::printf("{\"mode\": \"render\", \"nx\": %d, \"ny\": %d, \"ns\": %d, \"maxDepth\": %d, \"ms\": [", nx, ny, ns, maxDepth);
        for (int i = 0; i < steps; i++) {
            const double t0 = now();
            runRenderer(ns, 8, 8); // main.cpp:98, blocking
            const double t1 = now();
            std::printf("%s%.4f", i ? ", " : "", (t1 - t0) * 1e3);
        }
        std::printf("]}\n");
        if (std::strcmp(argv[11], "-") != 0 && crtWriteRef(argv[11], nx, ny, fb) != 0) { std::fprintf(stderr, "cannot write %s\n", argv[11]); return 1; }
        std::fflush(stdout);
        cleanupRenderer(); // main.cpp:138
        return 0;
    }
#ifdef WITH_SHIM
    if (mode == "spheres" && argc == 10) {
        const unsigned seed = (unsigned)std::atoi(argv[2]);
        const int nx = std::atoi(argv[3]), ny = std::atoi(argv[4]), ns = std::atoi(argv[5]), maxDepth = std::atoi(argv[6]);
        const int warmup = std::atoi(argv[7]), steps = std::atoi(argv[8]);
        std::vector<sphere> sph(1024);
        std::vector<material> mats(1024);
        const int n = crtRtiowScene(seed, sph.data(), mats.data(), 1024);
        camera cam;
        crtRtiowCamera(nx, ny, &cam);
        vec3* fb = nullptr;
        refSpheresInit(sph.data(), mats.data(), n, cam, &fb, nx, ny, maxDepth);
        for (int i = 0; i < warmup; i++) refSpheresRun(ns, 8, 8);
        std::printf("{\"mode\": \"spheres\", \"n\": %d, \"nx\": %d, \"ny\": %d, \"ns\": %d, \"maxDepth\": %d, \"ms\": [", n, nx, ny, ns, maxDepth);
        for (int i = 0; i < steps; i++) {
            const double t0 = now();
            refSpheresRun(ns, 8, 8);
            const double t1 = now();
            std::printf("%s%.4f", i ? ", " : "", (t1 - t0) * 1e3);
        }
        std::printf("]}\n");
        if (std::strcmp(argv[9], "-") != 0 && crtWriteRef(argv[9], nx, ny, fb) != 0) return 1;
        refSpheresCleanup();
        return 0;
    }
    if (mode == "scatter" && argc == 5) { // scatter <preset> <in.bin> <out.bin>: int64 n, then n x 12 floats each way
        std::ifstream in(argv[3], std::ios::binary);
        long long n = 0;
        in.read((char*)&n, 8);
        std::vector<float> a(12 * (size_t)n), b(12 * (size_t)n);
        in.read((char*)a.data(), a.size() * 4);
        if (!in) { std::fprintf(stderr, "short read on %s\n", argv[3]); return 1; }
        if (refScatterBatch(std::atoi(argv[2]), n, a.data(), b.data()) != 0) return 1;
        std::ofstream out(argv[4], std::ios::binary);
        out.write((const char*)&n, 8);
        out.write((const char*)b.data(), b.size() * 4);
        std::printf("{\"mode\": \"scatter\", \"n\": %lld}\n", n);
        return 0;
    }
    if (mode == "batch" && argc == 8) {
        const int texSize = std::atoi(argv[3]), ppl = std::atoi(argv[4]);
        void* scene = makeScene(argv[2], texSize, ppl);
        if (!scene) { std::fprintf(stderr, "scene creation failed\n"); return 1; }
        const kernel_scene ksc = *crtSceneKernelScene(scene);
        camera cam;
        crtStaircaseCamera(64, 64, &cam);
        vec3* fb = nullptr;
        initRenderer(ksc, cam, &fb, 64, 64, 1);
        // rays.bin: int64 n, then n float4 origins(+tMin), then n float4 dirs(+tMax)
        std::ifstream in(argv[5], std::ios::binary);
        long long n = 0;
        in.read((char*)&n, 8);
        std::vector<float> ro(4 * (size_t)n), rd(4 * (size_t)n), hit(4 * (size_t)n);
        std::vector<int> meshId((size_t)n);
        in.read((char*)ro.data(), ro.size() * 4);
        in.read((char*)rd.data(), rd.size() * 4);
        if (!in) { std::fprintf(stderr, "short read on %s\n", argv[5]); return 1; }
        const float ms = refIntersectBatch(ro.data(), rd.data(), n, std::atoi(argv[6]), hit.data(), meshId.data());
        std::ofstream out(argv[7], std::ios::binary);
        out.write((const char*)&n, 8);
        out.write((const char*)hit.data(), hit.size() * 4);
        out.write((const char*)meshId.data(), meshId.size() * 4);
        std::printf("{\"mode\": \"batch\", \"n\": %lld, \"kernel_ms\": %.4f}\n", n, ms);
        std::fflush(stdout);
        cleanupRenderer();
        return 0;
    }
#endif
    return usage();
}
