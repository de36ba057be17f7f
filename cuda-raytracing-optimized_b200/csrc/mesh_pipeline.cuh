// mesh_pipeline.cuh -- the triangle-mesh render path as a two-launch wavefront iteration.
//
//   traceKernel   every ray of the iteration, closest-hit (extend) and any-hit (shadow) alike, traversed by
//                 persistent warps that refill idle lanes one by one from the trace queue; a ray that exceeds its
//                 step budget parks its 24-byte state and is re-queued for the next launch (traverse.cuh)
//   shadeKernel   everything between two hit() calls of the reference's color() loop (kernels.cu:402-531): miss / light /
//                 albedo / scatter / next-event sample / Russian roulette, plus -- when the path ends -- retiring the
//                 sample into the pixel's sum (kernels.cu:558) and generating the next camera ray (kernels.cu:549-555)
//   startKernel   seeds the slots and generates sample 0 (kernels.cu:541-555)
//
// Ordering rules that keep the float sums bit-identical to the reference's sequential thread:
//   * a slot has at most ONE shadow ray in flight (`pending`); shadeKernel does not touch a slot whose shadow ray is
//     still pending, it re-queues the entry for the next iteration; so p.color receives light contributions and the
//     sky term in bounce order (kernels.cu:424,508);
//   * the shadow ray of a path's last bounce is FINAL: it carries the finished sample's colour and adds
//     colour (+ contribution when unoccluded) to the pixel sum itself, while the slot already traces its next sample;
//     the next sample cannot retire before `pending` clears, so samples reach `col` in sample order (kernels.cu:558).
//
// Queue entries are 32-bit: slot index | ENTRY_SHADOW | ENTRY_RESUME. Appends use one atomic per warp (ballot + popc).
#pragma once

#include "bsdf.cuh"
#include "device_scene.cuh"
#include "traverse.cuh"
#include "wavefront_kernels.cuh"

#define ENTRY_RESUME 0x80000000u
#define ENTRY_SHADOW 0x40000000u
#define ENTRY_SLOT_MASK 0x3FFFFFFFu

#define SHADOW_FLAG_FINAL 1u

struct MeshControl {
    unsigned int traceCount[2];  // entries in traceQ[k]
    unsigned int shadeCount[2];  // entries in shadeQ[k]
    unsigned int traceCursor;    // dynamic fetch cursor of the running traceKernel
    unsigned int blocksDone;     // shadeKernel's last block resets the consumed queues
    unsigned int pad0, pad1;
    unsigned long long raysExtend;  // finished closest-hit rays
    unsigned long long raysShadow;  // finished any-hit rays
    unsigned long long resumes;     // rays parked and continued in a later launch
    unsigned long long deferred;    // shade entries postponed because a shadow ray was pending
    unsigned long long iterations;
    unsigned long long nodeVisits;
    unsigned long long triTests;
};

struct MeshState {
    // path state, one entry per slot
    float4* rayO;   // {origin, rng}
    float4* rayD;   // {rayDir, flags}
    float4* atten;  // {attenuation, sample index}
    float4* pcol;   // {p.color, -}
    float4* hit;    // {t/closest, u, v, triId}
    uint2* travE;   // parked closest-hit traversal {idx, bitStack}
    // the slot's shadow ray
    float4* shO;    // {origin, -}
    float4* shD;    // {shadowDir, lightDist}
    float4* shL;    // {lightContribution, flags}
    float4* shC;    // FINAL only: {finished sample's colour, -}
    uint2* travS;   // parked any-hit traversal
    unsigned char* pending;
    unsigned int* traceQ[2];
    unsigned int* shadeQ[2];
    float4* accum;
    MeshControl* ctl;
    unsigned int numSlots;
    unsigned int npix;
    int nx, ny;
    int samplesPerSlot;
    int slotsPerPixel;
    unsigned int streamBase;
    int traceBudget;    // steps per ray per launch before it is parked
    int traceMinActive; // refill a warp when fewer lanes than this still traverse
};

__device__ __forceinline__ void accumulatePixel(const MeshState& st, unsigned int pixel, float r, float g, float b) {
    if (st.slotsPerPixel == 1) { // one slot per pixel: plain read-modify-write, in sample order
        float4 a = st.accum[pixel];
        a.x += r; a.y += g; a.z += b;
        st.accum[pixel] = a;
    } else {
        atomicAdd(&st.accum[pixel].x, r);
        atomicAdd(&st.accum[pixel].y, g);
        atomicAdd(&st.accum[pixel].z, b);
    }
}

// Starts sample `sample` of `slot`: kernels.cu:549-555.
__device__ __forceinline__ void startSample(const MeshState& st, const CameraDev& cam, unsigned int slot, unsigned int rng, int sample) {
    const unsigned int pixel = slot % st.npix;
    const int px = (int)(pixel % (unsigned int)st.nx), py = (int)(pixel / (unsigned int)st.nx);
    const float u = float(px + rnd(rng)) / float(st.nx);
    const float v = float(py + rnd(rng)) / float(st.ny);
    f3 o, d;
    cameraRay(cam, u, v, rng, o, d);
    st.rayO[slot] = mk4(o, __uint_as_float(rng));
    st.rayD[slot] = mk4(d, __uint_as_float(0u));
    st.atten[slot] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(sample));
    st.pcol[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

__global__ void __launch_bounds__(WF_BLOCK) meshStartKernel(MeshState st, CameraDev cam) {
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < st.numSlots; base += stride) {
        const unsigned int slot = base + laneId();
        const bool alive = slot < st.numSlots;
        if (alive) {
            const unsigned int pixel = slot % st.npix;
            const unsigned int stream = st.streamBase * (unsigned int)st.slotsPerPixel + slot / st.npix;
            st.pending[slot] = 0;
            startSample(st, cam, slot, pathSeed(pixel + stream * st.npix), 0); // kernels.cu:541-542 is stream 0
        }
        const unsigned int pos = warpAppend(alive, &st.ctl->traceCount[0]);
        if (alive) st.traceQ[0][pos] = slot;
    }
}

// ------------------------------------------------------------------- trace --
#define TRACE_BUDGET 96        // default steps per ray per launch before it is parked
#define TRACE_MIN_ACTIVE 20    // default: refill when fewer lanes than this hold a ray
#define TRACE_BLOCKS_PER_SM 5  // register cap 48: occupancy is what hides the L1/L2 latency of the node fetches

template <bool COUNT>
__global__ void __launch_bounds__(WF_BLOCK, TRACE_BLOCKS_PER_SM) traceKernel(MeshState st, MeshView mesh, int cur) {
    __shared__ RayCold coldAll[WF_BLOCK];
    RayCold& c = coldAll[threadIdx.x];
    MeshControl* ctl = st.ctl;
    const unsigned int n = ctl->traceCount[cur];
    const unsigned int* __restrict__ queue = st.traceQ[cur];
    unsigned int* __restrict__ nextTrace = st.traceQ[cur ^ 1];
    unsigned int* __restrict__ shadeQ = st.shadeQ[cur];
    const unsigned int lane = laneId();
    // Rays a warp holds at a time: 32 when the queue is long; when it is shorter than the grid (the tail of a frame) the
    // rays are spread over all warps, so that no ray waits in lockstep for a longer one in the same warp.
    // (`take` is a launch constant, but under the 48-register cap ptxas spilled it to local memory; it lives in shared
    // memory instead, read through a volatile pointer where it is used: local-memory traffic of this kernel is zero.)
    __shared__ unsigned int takeShared;
    if (threadIdx.x == 0) {
        const unsigned int totalWarps = gridDim.x * (WF_BLOCK / 32);
        takeShared = min(32u, max(1u, (n + totalWarps - 1) / totalWarps));
    }
    __syncthreads();
    const volatile unsigned int* takePtr = &takeShared;
#define take (*takePtr)
#define tail (take < 32u)
#define refillBelow (tail ? take : (unsigned int)st.traceMinActive)

    bool live = false;       // this lane holds a ray
    bool exhausted = false;  // the queue has no more entries for this warp
    bool isShadow = false;
    RayHot r;
    TravHot s;
    int steps = 0;
    unsigned int nodeVisits = 0, triTests = 0;
    r.ox = r.oy = r.oz = r.ix = r.iy = r.iz = 0.0f;
    s.idx = 0u; s.bitStack = 0u; s.closest = 0.0f;

    while (true) {
        // Lanes that still traverse. Finished rays stay in their lanes until the warp runs low on work: retiring them one
        // by one costs the whole warp a pass through the retire code per ray (measured: +25 % kernel time), so rays are
        // retired -- and idle lanes refilled -- in batches.
        bool working = live && s.idx != 0u && steps < st.traceBudget;
        unsigned int workMask = __ballot_sync(0xFFFFFFFFu, working);
        if (exhausted ? workMask == 0u : (unsigned int)__popc(workMask) < refillBelow) {
            // ---- retire finished rays, park the ones that ran out of budget
            const bool finished = live && s.idx == 0u;
            const bool park = live && !finished && steps >= st.traceBudget;
            const unsigned int entry = __float_as_uint(c.rec.w);
            const unsigned int slot = entry & ENTRY_SLOT_MASK;
            const bool toShade = finished && !isShadow;
            if (finished) {
                if (!isShadow) {
                    // hitMesh returns `closest` (== t_max when nothing was hit) or FLT_MAX; hit() tests `< t_max` (kernels.cu:330)
                    st.hit[slot] = make_float4(s.closest, c.rec.x, c.rec.y, c.rec.z);
                } else {
                    const float4 l = st.shL[slot];
                    const bool unoccluded = !(s.closest < c.dir.w); // hit(...) false: p.color += p.lightContribution (kernels.cu:500-508)
                    if (__float_as_uint(l.w) & SHADOW_FLAG_FINAL) {
                        float4 col = st.shC[slot];
                        if (unoccluded) { col.x += l.x; col.y += l.y; col.z += l.z; }
                        accumulatePixel(st, slot % st.npix, col.x, col.y, col.z); // col += p.color (kernels.cu:558)
                    } else if (unoccluded) {
                        float4 col = st.pcol[slot];
                        col.x += l.x; col.y += l.y; col.z += l.z;
                        st.pcol[slot] = col;
                    }
                    st.pending[slot] = 0;
                }
                live = false;
            }
            if (park) {
                if (isShadow) {
                    st.travS[slot] = make_uint2(s.idx, s.bitStack);
                } else {
                    st.travE[slot] = make_uint2(s.idx, s.bitStack);
                    st.hit[slot] = make_float4(s.closest, c.rec.x, c.rec.y, c.rec.z);
                }
                s.idx = 0u;
                live = false;
            }
            // queue appends and ray statistics: one atomic per warp per counter (ballot + popc)
            const unsigned int mShade = __ballot_sync(0xFFFFFFFFu, toShade);
            const unsigned int mPark = __ballot_sync(0xFFFFFFFFu, park);
            const unsigned int mShadowDone = __ballot_sync(0xFFFFFFFFu, finished && isShadow);
            unsigned int baseShade = 0, basePark = 0;
            if (lane == 0) {
                if (mShade) { baseShade = atomicAdd(&ctl->shadeCount[cur], __popc(mShade)); atomicAdd(&ctl->raysExtend, (unsigned long long)__popc(mShade)); }
                if (mPark) { basePark = atomicAdd(&ctl->traceCount[cur ^ 1], __popc(mPark)); atomicAdd(&ctl->resumes, (unsigned long long)__popc(mPark)); }
                if (mShadowDone) atomicAdd(&ctl->raysShadow, (unsigned long long)__popc(mShadowDone));
            }
            baseShade = __shfl_sync(0xFFFFFFFFu, baseShade, 0);
            basePark = __shfl_sync(0xFFFFFFFFu, basePark, 0);
            const unsigned int below = (1u << lane) - 1u;
            if (toShade) shadeQ[baseShade + __popc(mShade & below)] = slot;
            if (park) nextTrace[basePark + __popc(mPark & below)] = entry | ENTRY_RESUME;

            // ---- refill idle lanes, one atomic per warp
            if (!exhausted) {
                const unsigned int idle = ~workMask; // every lane that does not traverse has just been retired (or was empty)
                const unsigned int count = tail ? min((unsigned int)__popc(idle), take - (unsigned int)__popc(workMask)) : (unsigned int)__popc(idle);
                unsigned int base = 0;
                if (lane == 0) base = atomicAdd(&ctl->traceCursor, count);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (base + count >= n) exhausted = true; // warp-uniform: the tail of the queue has been handed out
                const unsigned int rank = __popc(idle & below);
                const unsigned int i = base + rank;
                if (!live && rank < count && i < n) {
                    const unsigned int e = queue[i];
                    const unsigned int sl = e & ENTRY_SLOT_MASK;
                    isShadow = (e & ENTRY_SHADOW) != 0u;
                    const float4 ro = isShadow ? st.shO[sl] : st.rayO[sl];
                    const float4 rd = isShadow ? st.shD[sl] : st.rayD[sl];
                    const float tMax = isShadow ? rd.w : FLT_MAX;
                    prepRay(r, c, xyz(ro), unit(xyz(rd)), tMax); // hit(): ray(p.origin, dir) normalises again (kernels.cu:326)
                    steps = 0;
                    live = true;
                    c.rec = make_float4(0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu), __uint_as_float(e));
                    if (e & ENTRY_RESUME) {
                        const uint2 t = isShadow ? st.travS[sl] : st.travE[sl];
                        s.idx = t.x;
                        s.bitStack = t.y;
                        s.closest = tMax;
                        if (!isShadow) {
                            const float4 h = st.hit[sl];
                            s.closest = h.x;
                            c.rec = make_float4(h.y, h.z, h.w, __uint_as_float(e));
                        }
                    } else {
                        s.idx = 1u; s.bitStack = 1u; s.closest = tMax;
                        if (!rayHitsBounds(mesh, r, tMax)) { // hitMesh: scene bounds first (kernels.cu:297)
                            s.idx = 0u;
                            s.closest = FLT_MAX;
                        }
                    }
                }
            }
            if (!__any_sync(0xFFFFFFFFu, live)) break;
            working = live && s.idx != 0u;
            workMask = __ballot_sync(0xFFFFFFFFu, working);
            if (workMask == 0u) continue; // e.g. every new ray missed the scene bounds: retire them
        }

        // ---- one scheduling round: node steps while enough lanes stand on nodes, then the leaves
        if (tail) // short queue: latency-bound launch, spend instructions on prefetching
            travRound<true>(mesh, r, c, RT_EPSILON, isShadow, working, s, steps, 1, nodeVisits, triTests);
        else
            travRound<false>(mesh, r, c, RT_EPSILON, isShadow, working, s, steps, max(1, min(TRACE_NODE_QUORUM, __popc(workMask) >> 1)), nodeVisits, triTests);
    }

    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) {
            nodeVisits += __shfl_xor_sync(0xFFFFFFFFu, nodeVisits, o);
            triTests += __shfl_xor_sync(0xFFFFFFFFu, triTests, o);
        }
        if (lane == 0) {
            atomicAdd(&ctl->nodeVisits, (unsigned long long)nodeVisits);
            atomicAdd(&ctl->triTests, (unsigned long long)triTests);
        }
    }
}
#undef take
#undef tail
#undef refillBelow

// ------------------------------------------------------------------- shade --
__global__ void __launch_bounds__(WF_BLOCK) meshShadeKernel(MeshState st, ShadeScene sc, CameraDev cam, int cur) {
    MeshControl* ctl = st.ctl;
    const unsigned int n = ctl->shadeCount[cur];
    const unsigned int* __restrict__ queue = st.shadeQ[cur];
    unsigned int* __restrict__ nextTrace = st.traceQ[cur ^ 1];
    unsigned int* __restrict__ nextShade = st.shadeQ[cur ^ 1];
    const unsigned int stride = gridDim.x * blockDim.x;
    unsigned int deferredCount = 0;
    for (unsigned int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
        const unsigned int i = base + laneId();
        bool traceNext = false, castsShadow = false, defer = false;
        unsigned int slot = 0;
        if (i < n) {
            slot = queue[i];
            if (st.pending[slot]) {
                defer = true; // its shadow ray is still in flight: keep bounce order, come back next iteration
            } else {
                const float4 h = st.hit[slot];
                const float4 ro = st.rayO[slot];
                const float4 rd = st.rayD[slot];
                const float4 att4 = st.atten[slot];
                f3 origin = xyz(ro), dir = xyz(rd), att = xyz(att4);
                unsigned int rng = __float_as_uint(ro.w);
                unsigned int flags = __float_as_uint(rd.w);
                const bool specularIn = (flags & PATH_FLAG_SPECULAR) != 0u;
                bool inside = (flags & PATH_FLAG_INSIDE) != 0u;
                unsigned int bounce = flags & PATH_BOUNCE_MASK;
                const f3 rdir = unit(dir); // direction of the ray hit() traced
                float4 pc = st.pcol[slot];
                bool continues = false;
                f3 shDir = mk3(0.0f, 0.0f, 0.0f), shL = mk3(0.0f, 0.0f, 0.0f);
                float lightDist = 0.0f;

                if (!(h.x < FLT_MAX)) {
                    // no mesh hit. Specular paths may still see the light sphere (kernels.cu:346-349); it ends the path
                    // without adding emission because SHADOW is defined (:440-446). Otherwise: constant grey sky (:424).
                    const bool hitsLight = specularIn && sphereHitT(sc.light.center, sc.light.radius, origin, rdir, RT_EPSILON, FLT_MAX) < FLT_MAX;
                    if (!hitsLight) {
                        const f3 add = att * mk3(0.5f, 0.5f, 0.5f);
                        pc.x += add.x; pc.y += add.y; pc.z += add.z;
                    }
                } else {
                    const unsigned int triId = __float_as_uint(h.w);
                    const float4 s0 = __ldg(sc.triShade + 3 * triId);
                    const float4 s1 = __ldg(sc.triShade + 3 * triId + 1);
                    const float4 s2 = __ldg(sc.triShade + 3 * triId + 2);
                    const int meshID = __float_as_int(s0.w);
                    SurfacePoint sp;
                    sp.normal = xyz(s0);
                    sp.t = h.x;
                    sp.inside = inside;
                    const float hu = h.y, hv = h.z;
                    // texCoords: u weights vertex 1, v weights vertex 2 (kernels.cu:337-338)
                    const float hw = (1 - hu - hv);
                    float tu = __fmaf_rn(hw, s1.x, mad2(hu, s1.z, hv, s2.x)); // hu*tc[2] + hv*tc[4] + (1-hu-hv)*tc[0]
                    float tv = __fmaf_rn(hw, s1.y, mad2(hu, s1.w, hv, s2.y));
                    if (dot(rdir, sp.normal) > 0.0f) sp.normal = -sp.normal;

                    const float4 m0 = __ldg(sc.mats.mats + 2 * meshID);
                    const float4 m1 = __ldg(sc.mats.mats + 2 * meshID + 1);
                    const int texId = __float_as_int(m1.y);
                    f3 albedo;
                    if (texId != -1) { // kernels.cu:457-471: nearest texel, frac() wrap
                        const int width = sc.mats.texWidth[texId];
                        const int height = sc.mats.texHeight[texId];
                        tu = tu - floorf(tu);
                        tv = tv - floorf(tv);
                        const int tx = (width - 1) * tu;
                        const int ty = (height - 1) * tv;
                        const int tIdx = ty * width + tx;
                        const float* td = sc.mats.texData[texId];
                        albedo = mk3(__ldg(td + tIdx * 3 + 0), __ldg(td + tIdx * 3 + 1), __ldg(td + tIdx * 3 + 2));
                    } else {
                        albedo = xyz(m0);
                    }

                    Scatter scat;
                    scat.specular = false;
                    scat.throughput = mk3(1.0f, 1.0f, 1.0f);
                    scat.refracted = false;
                    scat.t = h.x;
                    scat.wi = mk3(0.0f, 0.0f, 0.0f);
                    materialScatter(scat, sp, dir, __float_as_int(m1.x), m0.w, albedo, rng);

                    origin = origin + scat.t * dir; // kernels.cu:485 (not inters.p)
                    dir = scat.wi;
                    att = att * scat.throughput;
                    const bool specular = scat.specular;
                    inside = scat.refracted ? !inside : inside;

                    if (!specular && sampleLight(sc.light, origin, sp.normal, att, rng, shDir, shL, lightDist)) castsShadow = true;

                    continues = true;
                    if (bounce > 3u) { // Russian roulette, kernels.cu:514-526
                        const float m = maxcomp(att);
                        if (rnd(rng) > m) continues = false;
                        else att = att * (1 / m);
                    }
                    if (continues) {
                        bounce = (bounce + 1u) & PATH_BOUNCE_MASK; // p.bounce is a uint8_t (helper_structs.h:58)
                        if (!((int)bounce < sc.maxDepth)) continues = false;
                    }
                    flags = bounce | (specular ? PATH_FLAG_SPECULAR : 0u) | (inside ? PATH_FLAG_INSIDE : 0u);
                }

                if (castsShadow) {
                    st.shO[slot] = mk4(origin, 0.0f);
                    st.shD[slot] = mk4(shDir, lightDist);
                    st.shL[slot] = mk4(shL, __uint_as_float(continues ? 0u : SHADOW_FLAG_FINAL));
                    st.pending[slot] = 1;
                }
                if (continues) {
                    st.rayO[slot] = mk4(origin, __uint_as_float(rng));
                    st.rayD[slot] = mk4(dir, __uint_as_float(flags));
                    st.atten[slot] = mk4(att, att4.w);
                    traceNext = true;
                } else {
                    // the sample is finished: retire its colour (now, or by its FINAL shadow ray) and start the next one
                    if (castsShadow) st.shC[slot] = pc;
                    else accumulatePixel(st, slot % st.npix, pc.x, pc.y, pc.z);
                    const int sample = __float_as_int(att4.w) + 1;
                    if (sample < st.samplesPerSlot) {
                        startSample(st, cam, slot, rng, sample);
                        traceNext = true;
                    }
                }
            }
        }
        const unsigned int posDefer = warpAppend(defer, &ctl->shadeCount[cur ^ 1]);
        if (defer) { nextShade[posDefer] = slot; deferredCount++; }
        // extend and shadow entries of a warp go out with one atomic
        const unsigned int mE = __ballot_sync(0xFFFFFFFFu, traceNext), mS = __ballot_sync(0xFFFFFFFFu, castsShadow);
        const unsigned int total = __popc(mE) + __popc(mS);
        if (total) {
            unsigned int basePos = 0;
            if (laneId() == 0) basePos = atomicAdd(&ctl->traceCount[cur ^ 1], total);
            basePos = __shfl_sync(0xFFFFFFFFu, basePos, 0);
            const unsigned int below = (1u << laneId()) - 1u;
            if (castsShadow) nextTrace[basePos + __popc(mS & below)] = slot | ENTRY_SHADOW; // shadow rays first: they unblock the slot
            if (traceNext) nextTrace[basePos + __popc(mS) + __popc(mE & below)] = slot;
        }
    }
    if (deferredCount) atomicAdd(&ctl->deferred, (unsigned long long)deferredCount);

    // the last block to finish recycles the queues this iteration consumed
    __shared__ bool isLast;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        isLast = atomicAdd(&ctl->blocksDone, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (isLast && threadIdx.x == 0) {
        if (ctl->traceCount[cur] | ctl->shadeCount[cur]) ctl->iterations += 1;
        ctl->traceCount[cur] = 0;
        ctl->shadeCount[cur] = 0;
        ctl->traceCursor = 0;
        ctl->blocksDone = 0;
    }
}

// ------------------------------------------------------------ express lane --
// The frame's critical path is its heaviest pixels: one RNG stream per pixel means a pixel's bounces are strictly
// sequential (kernels.cu:542-548), and a pixel looking into glass needs ~5x the bounces of an average one. In a single
// wavefront those pixels advance one bounce per iteration while iterations are long (~1 ms with ~1.6 M rays), and then
// drag a ~4000-iteration tail of nearly empty launches behind the bulk. So slots that fall behind (sample index well below
// the mean of all live slots) are moved to a second, small wavefront -- same kernels, same state arrays, own queues and
// control block -- that runs concurrently on its own stream with short iterations. Results do not change: every slot is
// still processed by the same kernels in the same per-slot order, only by another queue.
__global__ void laneStatsKernel(MeshState a, unsigned long long* sums) {
    const unsigned int n1 = a.ctl->traceCount[0], n2 = a.ctl->shadeCount[0];
    unsigned long long sum = 0, cnt = 0;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
        const unsigned int entry = i < n1 ? a.traceQ[0][i] : a.shadeQ[0][i - n1];
        if (entry & ENTRY_SHADOW) continue; // count every live slot once: by its extend entry or its deferred shade entry
        sum += (unsigned long long)__float_as_int(a.atten[entry & ENTRY_SLOT_MASK].w);
        cnt += 1;
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    }
    if (laneId() == 0 && cnt) { atomicAdd(&sums[0], sum); atomicAdd(&sums[1], cnt); }
}

// Splits A's input queues (index 0): lagging slots go to B's input queues, the rest to A's spare queues (index 1).
// A slot's entries (extend, shadow, deferred shade) all take the same side: the decision reads only per-slot state.
__global__ void lanePartitionKernel(MeshState a, MeshState b, const unsigned long long* sums, float factor, int minMean, unsigned int moveAllBelow,
                                    unsigned int capB) {
    const unsigned int n1 = a.ctl->traceCount[0], n2 = a.ctl->shadeCount[0];
    const float mean = sums[1] ? (float)((double)sums[0] / (double)sums[1]) : 0.0f;
    const bool moveAll = n1 + n2 <= moveAllBelow;
    const bool allowed = moveAll || (mean >= (float)minMean && b.ctl->traceCount[0] + b.ctl->shadeCount[0] < capB);
    const float threshold = factor * mean;
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n1 + n2; base += stride) {
        const unsigned int i = base + laneId();
        const bool valid = i < n1 + n2;
        const bool isTrace = i < n1;
        unsigned int entry = 0;
        bool lag = false;
        if (valid) {
            entry = isTrace ? a.traceQ[0][i] : a.shadeQ[0][i - n1];
            lag = allowed && (moveAll || (float)__float_as_int(a.atten[entry & ENTRY_SLOT_MASK].w) < threshold);
        }
        unsigned int pos;
        pos = warpAppend(valid && isTrace && lag, &b.ctl->traceCount[0]);
        if (valid && isTrace && lag) b.traceQ[0][pos] = entry;
        pos = warpAppend(valid && isTrace && !lag, &a.ctl->traceCount[1]);
        if (valid && isTrace && !lag) a.traceQ[1][pos] = entry;
        pos = warpAppend(valid && !isTrace && lag, &b.ctl->shadeCount[0]);
        if (valid && !isTrace && lag) b.shadeQ[0][pos] = entry;
        pos = warpAppend(valid && !isTrace && !lag, &a.ctl->shadeCount[1]);
        if (valid && !isTrace && !lag) a.shadeQ[1][pos] = entry;
    }
}

__global__ void laneCopyBackKernel(MeshState a) {
    const unsigned int n1 = a.ctl->traceCount[1], n2 = a.ctl->shadeCount[1];
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
        if (i < n1) a.traceQ[0][i] = a.traceQ[1][i];
        else a.shadeQ[0][i - n1] = a.shadeQ[1][i - n1];
    }
}

__global__ void laneCommitKernel(MeshControl* ctl) {
    ctl->traceCount[0] = ctl->traceCount[1];
    ctl->shadeCount[0] = ctl->shadeCount[1];
    ctl->traceCount[1] = 0;
    ctl->shadeCount[1] = 0;
}
