// main.cpp -- host driver: the role of the reference's main.cpp:62-139 with its hard-coded run parameters
// (640x800, 256 spp, depth 64, 8x8 blocks, C:\ asset paths; main.cpp:63-71, staircase_scene.h:122,162)
// turned into arguments.  Flow is the reference's: camera -> scene -> initRenderer -> timed runRenderer ->
// PPM on stdout / file -> optional RMSE against a REF_00.01 frame -> cleanupRenderer.
//
//   crt_render [--scene staircase|rtiow|file.bvh] [--nx N] [--ny N] [--ns N] [--depth N] [--detail F]
//              [--tex N] [--textures DIR] [--ppm out.ppm|-] [--ref frame.ref] [--save-ref frame.ref] [--bvh-out file.bvh] [--gpus N] [--frames K]
//   (the reference's single positional argument, maxDepth, is still accepted: main.cpp:73-74)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "host_api.h"
#include "kernels.h"

int main(int argc, char** argv) {
    int nx = 640, ny = 800, ns = 256, maxDepth = 64, tx = 8, ty = 8, texSize = 1024; // main.cpp:65-70
    float detail = 1.0f;
    int gpus = 1, frames = 1;
    std::string scene = "staircase", ppm, refIn, refOut, bvhOut, textureDir;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--scene") scene = next();
        else if (a == "--nx") nx = std::atoi(next());
        else if (a == "--ny") ny = std::atoi(next());
        else if (a == "--ns") ns = std::atoi(next());
        else if (a == "--depth") maxDepth = std::atoi(next());
        else if (a == "--detail") detail = (float)std::atof(next());
        else if (a == "--tex") texSize = std::atoi(next());
        else if (a == "--textures") textureDir = next(); // directory with the nine PNG files of staircase_scene.h:125-133
        else if (a == "--ppm") ppm = next();
        else if (a == "--ref") refIn = next();
        else if (a == "--save-ref") refOut = next();
        else if (a == "--bvh-out") bvhOut = next();
        else if (a == "--frames") frames = std::atoi(next()); // render the frame this many times (the last one is kept)
        else if (a == "--gpus") gpus = std::atoi(next()); // sample-sharded over N devices of this box, one NCCL reduce per frame
        else if (a[0] != '-') maxDepth = (int)std::strtol(a.c_str(), NULL, 10);
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    std::fprintf(stderr, "Rendering a %dx%d image with %d samples per pixel and max depth %d in %dx%d blocks.\n", nx, ny, ns, maxDepth, tx, ty);

    vec3* fb = nullptr;
    crt_scene* sc = nullptr;
    camera cam;
    if (scene == "rtiow") {
        std::vector<sphere> sph(1024);
        std::vector<material> mats(1024);
        const int n = crtRtiowScene(1, sph.data(), mats.data(), 1024);
        crtRtiowCamera(nx, ny, &cam);
        initRendererSpheres(sph.data(), mats.data(), n, cam, &fb, nx, ny, maxDepth);
    } else {
        const bool isFile = scene.size() > 4 && scene.compare(scene.size() - 4, 4, ".bvh") == 0;
        sc = isFile ? crtSceneLoadBVH(scene.c_str(), texSize) : crtSceneCreateStaircase(detail, texSize, 5);
        if (!sc) { std::fprintf(stderr, "Failed to setup kernel scene\n"); return -1; }
        if (!textureDir.empty() && crtSceneLoadTextureDir(sc, textureDir.c_str()) != 9) {
            std::cerr << "Failed to load textures" << std::endl; // staircase_scene.h:134-137
            return -1;
        }
        const kernel_scene* ksc = crtSceneKernelScene(sc);
        std::fprintf(stderr, " there are %u triangles, and %d bvh nodes\n numPrimitivesPerNode %d\n", ksc->m->numTris, ksc->m->numBvhNodes,
                     ksc->numPrimitivesPerLeaf); // staircase_scene.h:177-179
        if (!bvhOut.empty() && crtSceneSaveBVH(sc, bvhOut.c_str()) != 0) std::fprintf(stderr, "cannot write %s\n", bvhOut.c_str());
        crtStaircaseCamera(nx, ny, &cam);
        if (gpus > 1) setRendererGpus(gpus);
        initRenderer(*ksc, cam, &fb, nx, ny, maxDepth);
    }

    for (int f = 1; f < frames; f++) runRenderer(ns, tx, ty);
    const auto start = std::chrono::steady_clock::now();
    runRenderer(ns, tx, ty);
    const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    renderer_stats st;
    getRendererStats(&st);
    std::fprintf(stderr, "took %.3f seconds on %d GPU(s) (device %.1f ms; %.1f Mrays/s, %.2f Msamples/s, %llu wavefront iterations).\n", seconds,
                 getRendererGpus(), st.msTotal, (st.raysExtend + st.raysShadow) / (st.msTotal * 1e3), st.samples / (st.msTotal * 1e3), st.iterations);

    if (!ppm.empty()) crtWritePPM(ppm.c_str(), nx, ny, fb);
    if (!refIn.empty()) { // main.cpp:108-128
        std::vector<vec3> reference((size_t)nx * ny);
        if (crtReadRef(refIn.c_str(), nx, ny, reference.data()) != 0) {
            std::fprintf(stderr, "Failed to load reference image\n");
            return -1;
        }
        std::fprintf(stderr, "RMSE = %g\n", crtRmse(fb, reference.data(), nx, ny));
    }
    if (!refOut.empty()) crtWriteRef(refOut.c_str(), nx, ny, fb); // main.cpp:130-134
    cleanupRenderer();
    if (sc) crtSceneDestroy(sc);
    return 0;
}
