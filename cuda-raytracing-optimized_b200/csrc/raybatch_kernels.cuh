// raybatch_kernels.cuh -- closest-hit queries on a flat ray batch (BASELINE config 5) and its generator.
//
// The query is the arithmetic of hit() -> hitMesh() -> hitBvh() (kernels.cu:325-339, :296, :154) for one ray:
// the direction is normalised by the ray constructor (ray.h:9), the scene bounds are tested first, then the
// dual-node traversal runs.  Per ray the kernel moves 32 B in and 20 B out.
#pragma once

#include "bsdf.cuh"
#include "device_scene.cuh"
#include "rng.cuh"
#include "traverse.cuh"
#include "wide_traverse.cuh"

template <bool COUNT>
__global__ void __launch_bounds__(256, 5) intersectBatchKernel(MeshView mesh, const float4* __restrict__ triShade,
                                                            const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                            unsigned long long n, float4* __restrict__ outHit, int* __restrict__ outMesh,
                                                            unsigned long long* cursor, unsigned long long* counts,
                                                            const unsigned int* __restrict__ indices, const unsigned long long* nDevice, int anyHit) {
    // anyHit: hitMesh(.., isShadow = true): t = 0.0f when anything lies in (tMin, tMax), FLT_MAX otherwise (kernels.cu:207)
    // indices != nullptr: the rays to trace are rayO[indices[k]], k < *nDevice (the rays the wide walk could not certify)
    if (nDevice) n = *nDevice;
    __shared__ RayCold coldAll[256];
    __shared__ float tMinAll[256];
    RayCold& c = coldAll[threadIdx.x];
    const unsigned int lane = threadIdx.x & 31u;
    bool live = false, exhausted = false;
    unsigned long long index = 0;
    RayHot r;
    TravHot s;
    int steps = 0;
    unsigned int nodeVisits = 0, triTests = 0;
    r.ox = r.oy = r.oz = r.ix = r.iy = r.iz = 0.0f;
    s.idx = 0u; s.bitStack = 0u; s.closest = 0.0f;
    while (true) {
        unsigned int liveMask = __ballot_sync(0xFFFFFFFFu, live);
        if (!exhausted && __popc(liveMask) < 20) { // refill idle lanes, one atomic per warp
            const unsigned int mask = ~liveMask;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(mask));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (base + __popc(mask) >= n) exhausted = true;
            const unsigned long long i = base + __popc(mask & ((1u << lane) - 1u));
            if (!live && i < n) {
                index = indices ? (unsigned long long)indices[i] : i;
                const float4 ro = __ldg(rayO + index);
                const float4 rd = __ldg(rayD + index);
                prepRay(r, c, xyz(ro), unit(xyz(rd)), rd.w);
                tMinAll[threadIdx.x] = ro.w;
                c.rec = make_float4(0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu), 0.0f);
                live = true;
                s.idx = 1u; s.bitStack = 1u; s.closest = rd.w;
                if (!rayHitsBounds(mesh, r, rd.w)) {
                    s.idx = 0u;
                    s.closest = FLT_MAX;
                }
            }
            liveMask = __ballot_sync(0xFFFFFFFFu, live);
        }
        if (liveMask == 0u) break;
        travRound<false>(mesh, r, c, tMinAll[threadIdx.x], anyHit != 0, live, s, steps, max(1, min(TRACE_NODE_QUORUM, __popc(liveMask) >> 1)), nodeVisits, triTests);
        if (live && s.idx == 0u) {
            float t = s.closest;
            unsigned int triId = __float_as_uint(c.rec.z);
            float u = c.rec.x, v = c.rec.y;
            int meshID = -1;
            if (anyHit && t < c.dir.w) {
                t = 0.0f;
                triId = 0xFFFFFFFFu;
                u = v = 0.0f;
            } else if (t < c.dir.w) {
                meshID = __float_as_int(__ldg(triShade + 3 * triId).w);
            } else {
                t = FLT_MAX;
                triId = 0xFFFFFFFFu;
                u = v = 0.0f;
            }
            outHit[index] = make_float4(t, u, v, __uint_as_float(triId));
            outMesh[index] = meshID;
            live = false;
        }
    }
    if (COUNT) {
        atomicAdd(&counts[0], (unsigned long long)nodeVisits);
        atomicAdd(&counts[1], (unsigned long long)triTests);
    }
}

// The same query through the renderer's own wide tree (wide_traverse.cuh). Rays whose result the certificate does not
// cover are not answered here: their indices go to `redo` and the order-exact kernel above answers them.
// Dynamic shared memory: wide.stackDepth * WIDE_BATCH_BLOCK uint2.
#define WIDE_BATCH_BLOCK 128 // 7 blocks per SM at 72 registers: no spills
template <bool COUNT, bool CERTIFY>
__global__ void __launch_bounds__(WIDE_BATCH_BLOCK, 7) wideIntersectBatchKernel(MeshView mesh, WideView wide, const float4* __restrict__ triShade,
                                                                   const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                                   unsigned long long n, float4* __restrict__ outHit, int* __restrict__ outMesh,
                                                                   unsigned long long* cursor, unsigned long long* counts, unsigned int* __restrict__ redo,
                                                                   unsigned long long* redoCount, int anyHit) {
    extern __shared__ uint2 wideStackAll[];
    __shared__ RayCold coldAll[WIDE_BATCH_BLOCK];
    __shared__ float4 invAll[WIDE_BATCH_BLOCK]; // {the reference's 1 / direction, tMin}
    RayCold& c = coldAll[threadIdx.x];
    float4& invT = invAll[threadIdx.x];
    uint2* stack = wideStackAll + threadIdx.x;
    const unsigned int lane = threadIdx.x & 31u;
    bool live = false, exhausted = false;
    unsigned long long index = 0;
    WideRay r;
    WideTrav s;
    unsigned int nodeVisits = 0, triTests = 0;
    const unsigned int k3f = wideConst3F();
    r.ox = r.oy = r.oz = r.ix = r.iy = r.iz = 0.0f; r.oct = 0u;
    s.ngx = s.ngy = s.tgx = s.tgy = 0u; s.sp = -1; s.closest = 0.0f;
    while (true) {
        unsigned int liveMask = __ballot_sync(0xFFFFFFFFu, live);
        if (!exhausted && __popc(liveMask) < 20) { // refill idle lanes, one atomic per warp
            const unsigned int mask = ~liveMask;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(mask));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (base + __popc(mask) >= n) exhausted = true;
            const unsigned long long i = base + __popc(mask & ((1u << lane) - 1u));
            if (!live && i < n) {
                index = i;
                const float4 ro = __ldg(rayO + i);
                const float4 rd = __ldg(rayD + i);
                const f3 d = unit(xyz(rd)); // the ray constructor normalises (ray.h:9)
                c.dir = mk4(d, rd.w);
                c.rec = make_float4(0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu), 0.0f);
                f3 inv;
                const bool covered = wideSetup(wide, r, xyz(ro), d, anyHit != 0, inv);
                invT = mk4(inv, ro.w);
                if (!covered) {
                    redo[atomicAdd(redoCount, 1ull)] = (unsigned int)i;
                } else {
                    if (!wideHitsBounds(mesh, r, inv, rd.w)) { // hitMesh: scene bounds first (kernels.cu:297)
                        outHit[i] = make_float4(FLT_MAX, 0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu));
                        outMesh[i] = -1;
                    } else {
                        wideStart(s, rd.w);
                        live = true;
                    }
                }
            }
            liveMask = __ballot_sync(0xFFFFFFFFu, live);
        }
        if (liveMask == 0u) {
            if (exhausted) break;
            continue;
        }
        wideRound(wide, r, c, invT.w, live, s, stack, WIDE_BATCH_BLOCK, max(1, min(WIDE_NODE_QUORUM, __popc(liveMask) >> 1)), k3f, nodeVisits, triTests);
        if (live && s.sp < 0) {
            float t = s.closest;
            unsigned int triId = __float_as_uint(c.rec.z);
            float u = c.rec.x, v = c.rec.y;
            int meshID = -1;
            bool certified = true;
            if (triId != 0xFFFFFFFFu && t < c.dir.w) {
                if (CERTIFY) certified = wideCertify(mesh, r, xyz(invT), c.dir.w, t, triId);
                if (anyHit) { t = 0.0f; triId = 0xFFFFFFFFu; u = v = 0.0f; }
                else meshID = __float_as_int(__ldg(triShade + 3 * triId).w);
            } else {
                t = FLT_MAX;
                triId = 0xFFFFFFFFu;
                u = v = 0.0f;
            }
            if (certified) {
                outHit[index] = make_float4(t, u, v, __uint_as_float(triId));
                outMesh[index] = meshID;
            } else {
                redo[atomicAdd(redoCount, 1ull)] = (unsigned int)index;
            }
            live = false;
        }
    }
    if (COUNT) {
        atomicAdd(&counts[0], (unsigned long long)nodeVisits);
        atomicAdd(&counts[1], (unsigned long long)triTests);
    }
}

// First half of the batch: jittered camera rays over a filmW x filmH virtual film (coherent).
// Second half: origin uniform in the scene bounds, direction from the unit-sphere sampler (incoherent).
// Every ray has its own stream, seeded like a pixel (kernels.cu:542) from its index.
__global__ void generateRayBatchKernel(float4* __restrict__ rayO, float4* __restrict__ rayD, unsigned long long n, CameraDev cam,
                                       f3 bmin, f3 bmax, int filmW, int filmH, float tMin, float tMax) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned int rng = pathSeed((unsigned int)i);
    f3 o, d;
    if (i < n / 2) {
        const unsigned long long film = (unsigned long long)filmW * filmH;
        const unsigned long long p = i % film;
        const int px = (int)(p % filmW), py = (int)(p / filmW);
        const float u = float(px + rnd(rng)) / float(filmW);
        const float v = float(py + rnd(rng)) / float(filmH);
        o = cam.origin;
        d = unit(cam.lowerLeft + u * cam.horizontal + v * cam.vertical - cam.origin);
    } else {
        const float a = rnd(rng);
        const float b = rnd(rng);
        const float c = rnd(rng);
        o = mk3(bmin.x + a * (bmax.x - bmin.x), bmin.y + b * (bmax.y - bmin.y), bmin.z + c * (bmax.z - bmin.z));
        d = unit(randomInUnitSphere(rng));
    }
    rayO[i] = mk4(o, tMin);
    rayD[i] = mk4(d, tMax);
}

// ---- BSDF probe: one preset of scene_materials.h on a batch of surface points (tests) ----------------
// in:  3 float4 per item {normal.xyz, t} {p.xyz, inside (0/1)} {wo.xyz, rng state bits}
// out: 3 float4 per item {wi.xyz, t} {throughput.xyz, specular | refracted << 1 (as int bits)} {rng state bits after, 0, 0, 0}
__global__ void scatterBatchKernel(int preset, long long n, const float4* __restrict__ in, float4* __restrict__ out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 a = in[3 * k], b = in[3 * k + 1], c = in[3 * k + 2];
    SurfacePoint i;
    i.normal = xyz(a);
    i.t = a.w;
    i.inside = b.w != 0.0f;
    unsigned int rng = __float_as_uint(c.w);
    Scatter s; // scatter_info's constructor, helper_structs.h:45
    s.wi = mk3(0.0f, 0.0f, 0.0f);
    s.specular = false;
    s.throughput = mk3(1.0f, 1.0f, 1.0f);
    s.refracted = false;
    s.t = i.t;
    presetScatter(preset, s, i, xyz(b), xyz(c), rng);
    out[3 * k] = mk4(s.wi, s.t);
    out[3 * k + 1] = mk4(s.throughput, __int_as_float((s.specular ? 1 : 0) | (s.refracted ? 2 : 0)));
    out[3 * k + 2] = make_float4(__uint_as_float(rng), 0.0f, 0.0f, 0.0f);
}
