// ref_spheres.cu -- TEST INFRASTRUCTURE. "Reference-derived" thread-per-pixel megakernel for the README's
// random-spheres scene, which the reference's HEAD no longer contains.
// It is assembled from the reference's OWN device functions, included from /root/reference:
//   sphereHit (intersections.h:85), material_scatter (scene_materials.h:13), get_ray (camera.h:8),
//   rnd.h, render()'s pixel loop (kernels.cu:535-569), color()'s bounce loop with Russian roulette
//   (kernels.cu:396-533), the sky gradient (kernels.cu:419-421), spheres in __constant__ memory
//   (README.md:93-104).  It is NOT the reference; it is the closest thing to it that can exist,
//   and the definition csrc/spheres_path.cuh is compared against.
#include <cuda_runtime.h>
#include <device_launch_parameters.h>
#include <iostream>

#include "rnd.h"
#include "vec3.h"
#include "camera.h"
#include "intersections.h"
#include "material.h"
#include "scene_materials.h"

#define EPSILON 0.01f
#define MAX_SPHERES 1024

// raw storage: the reference's sphere/material have user-provided constructors, which __constant__ forbids
__constant__ float4 d_sphereRaw[MAX_SPHERES];      // sizeof(sphere) == 16
__constant__ int d_materialRaw[MAX_SPHERES * 6];   // sizeof(material) == 24
#define d_spheres ((const sphere*)d_sphereRaw)
#define d_materials ((const material*)d_materialRaw)

struct SphereContext {
    vec3* fb;
    int numSpheres, nx, ny, ns, maxDepth;
    camera cam;
    unsigned int stream;
};
static SphereContext sctx;

__device__ void colorSpheres(const SphereContext& context, path& p) {
    p.attenuation = vec3(1.0, 1.0, 1.0);
    p.color = vec3(0, 0, 0);
    for (p.bounce = 0; p.bounce < context.maxDepth; p.bounce++) {
        const ray r(p.origin, p.rayDir);
        float closest = FLT_MAX;
        int id = -1;
        for (int s = 0; s < context.numSpheres; s++) {
            float t = sphereHit(d_spheres[s], r, EPSILON, closest);
            if (t < closest) { closest = t; id = s; }
        }
        if (id < 0) {
            float t = 0.5f * (p.rayDir.y() + 1.0f);
            vec3 c = (1.0f - t) * vec3(1.0, 1.0, 1.0) + t * vec3(0.5, 0.7, 1.0);
            p.color += p.attenuation * c;
            return;
        }
        intersection inters;
        inters.objId = 1;
        inters.t = closest;
        inters.p = r.point_at_parameter(closest);
        inters.normal = (inters.p - d_spheres[id].center) / d_spheres[id].radius;
        if (dot(r.direction(), inters.normal) > 0.0f) inters.normal = -inters.normal;
        inters.inside = p.inside;
        scatter_info scatter(inters);
        material_scatter(scatter, inters, p.rayDir, d_materials[id], d_materials[id].color, p.rng);
        p.origin += scatter.t * p.rayDir;
        p.rayDir = scatter.wi;
        p.attenuation *= scatter.throughput;
        p.specular = scatter.specular;
        p.inside = scatter.refracted ? !p.inside : p.inside;
        if (p.bounce > 3) {
            float m = max(p.attenuation);
            if (rnd(p.rng) > m) return;
            p.attenuation *= 1 / m;
        }
    }
}

__global__ void renderSpheres(const SphereContext context) {
    int i = threadIdx.x + blockIdx.x * blockDim.x;
    int j = threadIdx.y + blockIdx.y * blockDim.y;
    if ((i >= context.nx) || (j >= context.ny)) return;
    path p;
    uint64_t pixelId = j * context.nx + i;
    p.rng = (wang_hash(pixelId + context.stream * context.nx * context.ny) * 336343633) | 1;
    vec3 col(0, 0, 0);
    for (int s = 0; s < context.ns; s++) {
        float u = float(i + rnd(p.rng)) / float(context.nx);
        float v = float(j + rnd(p.rng)) / float(context.ny);
        ray r = get_ray(context.cam, u, v, p.rng);
        p.origin = r.origin();
        p.rayDir = r.direction();
        p.specular = false;
        p.inside = false;
        colorSpheres(context, p);
        col += p.color;
    }
    context.fb[pixelId] = col / float(context.ns);
}

static void chk(cudaError_t e, const char* what) {
    if (e) { std::cerr << "CUDA error = " << cudaGetErrorString(e) << " at " << what << "\n"; cudaDeviceReset(); exit(99); }
}

extern "C" void refSpheresInit(const sphere* spheres, const material* materials, int n, const camera cam, vec3** fb, int nx, int ny,
                               int maxDepth) {
    sctx.numSpheres = n; sctx.nx = nx; sctx.ny = ny; sctx.maxDepth = maxDepth; sctx.cam = cam; sctx.stream = 0;
    chk(cudaMallocManaged((void**)&sctx.fb, (size_t)nx * ny * sizeof(vec3)), "fb");
    *fb = sctx.fb;
    chk(cudaMemcpyToSymbol(d_sphereRaw, spheres, n * sizeof(sphere)), "spheres");
    chk(cudaMemcpyToSymbol(d_materialRaw, materials, n * sizeof(material)), "materials");
}
extern "C" void refSpheresRun(int ns, int tx, int ty) {
    sctx.ns = ns;
    dim3 blocks((sctx.nx + tx - 1) / tx, (sctx.ny + ty - 1) / ty), threads(tx, ty);
    renderSpheres<<<blocks, threads>>>(sctx);
    chk(cudaGetLastError(), "launch");
    chk(cudaDeviceSynchronize(), "sync");
}
extern "C" void refSpheresCleanup() { cudaFree(sctx.fb); }
