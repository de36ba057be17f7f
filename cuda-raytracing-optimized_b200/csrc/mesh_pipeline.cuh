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
#include "wide_traverse.cuh"

#define ENTRY_RESUME 0x80000000u
#define ENTRY_SHADOW 0x40000000u
#define ENTRY_SLOT_MASK 0x3FFFFFFFu

#define SHADOW_FLAG_FINAL 1u  // the path ended at this bounce: the shadow ray carries the finished sample's colour
#define SHADOW_FLAG_LAST 2u   // ... and it was the slot's last sample: this shadow ray is all that is left of the slot

struct MeshControl {
    unsigned int traceCount[2];  // entries at the FRONT of traceQ[k] (extend rays, resumed / re-traced rays)
    unsigned int traceBack[2];   // entries at the BACK of traceQ[k], from its last element downwards (shadow rays): traceEntry()
    unsigned int shadeCount[2];  // entries in shadeQ[k]
    unsigned int traceCursor;    // dynamic fetch cursor of the running traceKernel
    unsigned int blocksDone;     // shadeKernel's last block resets the consumed queues
    unsigned int redoCount;      // entries in redoQ: rays the wide walk could not certify, re-traced in the reference's order
    unsigned int redoCursor;
    unsigned long long raysExtend;  // finished closest-hit rays
    unsigned long long raysShadow;  // finished any-hit rays
    unsigned long long resumes;     // rays parked and continued in a later launch
    unsigned long long deferred;    // shade entries postponed because a shadow ray was pending
    unsigned long long iterations;
    unsigned long long nodeVisits;
    unsigned long long triTests;
    unsigned long long redone;      // rays answered by the exact kernel although the wide tree was in use
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
    unsigned char* ready;  // 1 while the slot has an entry in the shade queue (its hit record waits to be shaded)
    unsigned int* rngOut;  // the slot's RNG state after its last sample of this run (continueRenderer / checkpoints start from it)
    unsigned int* traceQ[2];
    unsigned int* shadeQ[2];
    unsigned int* redoQ;   // trace entries of this iteration that go through the exact kernel (wide traversal only)
    float4* accum;
    MeshControl* ctl;
    unsigned int numSlots;
    unsigned int npix;
    int nx, ny;
    int samplesPerSlot;
    int slotsPerPixel;
    unsigned int tilesX;   // > 0: slots enumerate the frame in 8x4 pixel tiles (tilesX tiles per row); 0: row by row
    unsigned int streamBase;
    int traceBudget;    // steps per ray per launch before it is parked
    int traceMinActive; // refill a warp when fewer lanes than this still traverse
    unsigned int traceCap; // elements of traceQ[k]
};

// Entry i of a trace queue with `front` entries at its front and the rest at its back. The shade kernel appends extend rays at the
// front and shadow rays at the back, so that a trace launch walks all its closest-hit rays first (the long walks start early:
// the launch ends when its last ray ends) and then all its any-hit rays (a third as long), each kind among its own.
__device__ __forceinline__ unsigned int traceEntry(const unsigned int* __restrict__ q, unsigned int cap, unsigned int front, unsigned int i) {
    return q[i < front ? i : cap - 1u - (i - front)];
}

// The pixel a slot renders. Slots are numbered tile by tile (8 wide, 4 high) when the frame divides into such tiles, so that
// the 32 lanes of a warp -- consecutive slots wherever queues are in slot order -- are a compact patch of the image: their
// rays start together, hit the same materials and texture lines, walk the same part of the tree. Which slot renders a
// pixel changes nothing in the pixel (its RNG stream is seeded by the pixel's own index, kernels.cu:541-542).
__device__ __forceinline__ unsigned int slotPixel(const MeshState& st, unsigned int slot) {
    const unsigned int s = slot % st.npix;
    if (st.tilesX == 0u) return s;
    const unsigned int tile = s >> 5, o = s & 31u;
    const unsigned int ty = tile / st.tilesX, tx = tile - ty * st.tilesX;
    return (ty * 4u + (o >> 3)) * (unsigned int)st.nx + tx * 8u + (o & 7u);
}

__device__ __forceinline__ void accumulatePixel(const MeshState& st, unsigned int pixel, float r, float g, float b) {
    if (st.slotsPerPixel == 1) { // one slot per pixel: plain read-modify-write, in sample order (read from L2: the
        float4 a = __ldcg(&st.accum[pixel]); // previous add may have come from another SM while this kernel was running)
        a.x += r; a.y += g; a.z += b;
        st.accum[pixel] = a;
    } else {
        atomicAdd(&st.accum[pixel].x, r);
        atomicAdd(&st.accum[pixel].y, g);
        atomicAdd(&st.accum[pixel].z, b);
    }
}

// A path between two hit() calls, in registers (`path`, helper_structs.h:48-71, minus what the wavefront keeps elsewhere).
struct PathRegs {
    f3 origin, dir, att;   // rayDir is NOT normalised here: hit() normalises (kernels.cu:326), color() advances along the raw one (:485)
    unsigned int rng;
    unsigned int flags;    // bounce | PATH_FLAG_SPECULAR | PATH_FLAG_INSIDE
    int sample;            // index of the sample this path belongs to
    f3 color;              // p.color
};

// Starts sample `sample` of `slot`: kernels.cu:549-555.
__device__ __forceinline__ void startSample(const MeshState& st, const CameraDev& cam, unsigned int slot, unsigned int rng, int sample, PathRegs& p) {
    const unsigned int pixel = slotPixel(st, slot);
    const int px = (int)(pixel % (unsigned int)st.nx), py = (int)(pixel / (unsigned int)st.nx);
    const float u = float(px + rnd(rng)) / float(st.nx);
    const float v = float(py + rnd(rng)) / float(st.ny);
    cameraRay(cam, u, v, rng, p.origin, p.dir);
    p.rng = rng;
    p.flags = 0u; // bounce 0, specular = inside = false (kernels.cu:554-555)
    p.att = mk3(1.0f, 1.0f, 1.0f);
    p.sample = sample;
    p.color = mk3(0.0f, 0.0f, 0.0f);
}

__device__ __forceinline__ void storePath(const MeshState& st, unsigned int slot, const PathRegs& p, bool withColor) {
    st.rayO[slot] = mk4(p.origin, __uint_as_float(p.rng));
    st.rayD[slot] = mk4(p.dir, __uint_as_float(p.flags));
    st.atten[slot] = mk4(p.att, __int_as_float(p.sample));
    if (withColor) st.pcol[slot] = mk4(p.color, 0.0f);
}

// What shadePath decided.
struct ShadeResult {
    bool traceNext;    // `p` now holds the next extend ray of the slot (the path continues, or the next sample started)
    bool continued;    // ... it is the same path (p.color did not change)
    bool castsShadow;  // the bounce cast a shadow ray (handed to the sink)
};

// Where shadePath puts what leaves the path, as soon as it is known (so that the values do not stay in registers while the
// next camera ray is generated): the SoA arrays for meshShadeKernel, the partner lane's shared memory for chaseKernel.
//   castShadow(origin, dir, lightDist, contribution, flags, carried)  the bounce's shadow ray (kernels.cu:493-508); SHADOW_FLAG_FINAL = the
//                                                                     path ended at this bounce and the ray carries the sample's colour
//   retire(colour)                                                    the sample ended without a shadow ray: col += p.color (kernels.cu:558)
struct ArraySink {
    const MeshState& st;
    unsigned int slot;
    __device__ __forceinline__ void castShadow(const f3& origin, const f3& dir, float lightDist, const f3& contribution, unsigned int flags, const f3& carried) {
        st.shO[slot] = mk4(origin, 0.0f);
        st.shD[slot] = mk4(dir, lightDist);
        st.shL[slot] = mk4(contribution, __uint_as_float(flags));
        if (flags & SHADOW_FLAG_FINAL) st.shC[slot] = mk4(carried, 0.0f);
        st.pending[slot] = 1;
    }
    __device__ __forceinline__ void retire(const f3& c) { accumulatePixel(st, slotPixel(st, slot), c.x, c.y, c.z); }
};

// Everything between two hit() calls of color() (kernels.cu:402-531) for one path whose closest-hit record is `h`
// {t (FLT_MAX = no mesh hit), u, v, triId}: miss / light / albedo / scatter / next-event sample / Russian roulette, and
// -- when the path ends -- the start of the slot's next sample (kernels.cu:549-555). Shared by meshShadeKernel (state in
// the SoA arrays) and chaseKernel (state in registers).
template <class Sink>
__device__ __forceinline__ ShadeResult shadePath(const MeshState& st, const ShadeScene& sc, const CameraDev& cam, unsigned int slot, PathRegs& p,
                                                 const float4& h, Sink& sink) {
    ShadeResult out;
    out.traceNext = false; out.continued = false; out.castsShadow = false;
    f3 origin = p.origin, dir = p.dir, att = p.att;
    unsigned int rng = p.rng;
    unsigned int flags = p.flags;
    const bool specularIn = (flags & PATH_FLAG_SPECULAR) != 0u;
    bool inside = (flags & PATH_FLAG_INSIDE) != 0u;
    unsigned int bounce = flags & PATH_BOUNCE_MASK;
    const f3 rdir = unit(dir); // direction of the ray hit() traced
    f3 pc = p.color;
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

        if (!specular && sampleLight(sc.light, origin, sp.normal, att, rng, shDir, shL, lightDist)) out.castsShadow = true;

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

    if (out.castsShadow)
        sink.castShadow(origin, shDir, lightDist, shL,
                        continues ? 0u : (p.sample + 1 < st.samplesPerSlot ? SHADOW_FLAG_FINAL : (SHADOW_FLAG_FINAL | SHADOW_FLAG_LAST)), pc);
    if (continues) {
        p.origin = origin; p.dir = dir; p.att = att; p.rng = rng; p.flags = flags; p.color = pc;
        out.traceNext = true;
        out.continued = true;
    } else {
        // the sample is finished: retire its colour (now, or by its FINAL shadow ray) and start the next one
        if (!out.castsShadow) sink.retire(pc);
        const int sample = p.sample + 1;
        p.rng = rng; // the slot's stream runs on into the next sample (kernels.cu:542-548)
        if (sample < st.samplesPerSlot) {
            startSample(st, cam, slot, rng, sample, p);
            out.traceNext = true;
        } else {
            st.rngOut[slot] = rng; // where the pixel's stream stands: a later continueRenderer() goes on from here
        }
    }
    return out;
}

// resume = 0: seed every slot (kernels.cu:541-542). resume = 1: continue the streams where the previous run left them.
__global__ void __launch_bounds__(WF_BLOCK) meshStartKernel(MeshState st, CameraDev cam, int resume) {
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < st.numSlots; base += stride) {
        const unsigned int slot = base + laneId();
        const bool alive = slot < st.numSlots;
        if (alive) {
            const unsigned int pixel = slotPixel(st, slot);
            const unsigned int stream = st.streamBase * (unsigned int)st.slotsPerPixel + slot / st.npix;
            st.pending[slot] = 0;
            st.ready[slot] = 0;
            PathRegs p;
            startSample(st, cam, slot, resume ? st.rngOut[slot] : pathSeed(pixel + stream * st.npix), 0, p); // kernels.cu:541-542 is stream 0
            storePath(st, slot, p, true);
        }
        const unsigned int pos = warpAppend(alive, &st.ctl->traceCount[0]);
        if (alive) st.traceQ[0][pos] = slot;
    }
}

// ------------------------------------------------------------------- trace --
#define TRACE_BUDGET 128       // default steps per ray per launch before it is parked
#define TRACE_MIN_ACTIVE 20    // default: refill when fewer lanes than this hold a ray
#ifndef TRACE_BLOCK
#define TRACE_BLOCK 256
#endif
#define TRACE_BLOCKS_PER_SM (1280 / TRACE_BLOCK) // 40 warps per SM, register cap 48: occupancy is what hides the L1/L2 latency of the node fetches

// A per-warp flag byte in shared memory, addressed from %tid inside the asm block: the address is recomputed at every
// use instead of occupying a register (or a local-memory slot, which is what ptxas chose under the 48-register cap).
__device__ __forceinline__ unsigned int warpFlagLoad(unsigned int sharedBase) {
    unsigned int v;
    asm volatile("{\n\t.reg .u32 t;\n\tmov.u32 t, %%tid.x;\n\tshr.u32 t, t, 5;\n\tadd.u32 t, t, %1;\n\tld.volatile.shared.u8 %0, [t];\n\t}" : "=r"(v) : "r"(sharedBase));
    return v;
}
__device__ __forceinline__ void warpFlagSet(unsigned int sharedBase) {
    asm volatile("{\n\t.reg .u32 t;\n\t.reg .u32 one;\n\tmov.u32 t, %%tid.x;\n\tshr.u32 t, t, 5;\n\tadd.u32 t, t, %0;\n\tmov.u32 one, 1;\n\tst.volatile.shared.u8 [t], one;\n\t}" ::"r"(sharedBase) : "memory");
}

// CUR (which of the two queue sets is the input) is a template parameter so that the queue pointers are reads of the kernel's
// parameter bank and not six registers held across the traversal loop.
// REDO: the input is redoQ (what wideTraceKernel of the same iteration could not certify) instead of traceQ[CUR]; those rays
// run to their end here (no step budget: the wide pipeline has no parked rays).
template <bool COUNT, int CUR, bool REDO = false>
__global__ void __launch_bounds__(TRACE_BLOCK, TRACE_BLOCKS_PER_SM) traceKernel(MeshState st, MeshView mesh) {
    constexpr int cur = CUR;
    __shared__ RayCold coldAll[TRACE_BLOCK];
    RayCold& c = coldAll[threadIdx.x];
    MeshControl* ctl = st.ctl;
    const unsigned int* __restrict__ queue = REDO ? st.redoQ : st.traceQ[cur];
    unsigned int* __restrict__ nextTrace = st.traceQ[cur ^ 1];
    unsigned int* __restrict__ shadeQ = st.shadeQ[cur];
    const unsigned int lane = laneId();
    // Rays a warp holds at a time: 32 when the queue is long; when it is shorter than the grid (the tail of a frame) the
    // rays are spread over all warps, so that no ray waits in lockstep for a longer one in the same warp.
    // (`take` is a launch constant, but under the 48-register cap ptxas spilled it to local memory; it lives in shared
    // memory instead, read through a volatile pointer where it is used: local-memory traffic of this kernel is zero.)
    __shared__ unsigned int takeShared, countShared, frontShared;
    __shared__ unsigned char exhaustedShared[TRACE_BLOCK / 32]; // per warp: the queue has no more entries for this warp
    if (threadIdx.x == 0) {
        const unsigned int front = REDO ? ctl->redoCount : ctl->traceCount[cur];
        const unsigned int count = REDO ? front : front + ctl->traceBack[cur];
        const unsigned int totalWarps = gridDim.x * (TRACE_BLOCK / 32);
        frontShared = front;
        countShared = count;
        takeShared = min(32u, max(1u, (count + totalWarps - 1) / totalWarps));
    }
    if (threadIdx.x < TRACE_BLOCK / 32) exhaustedShared[threadIdx.x] = 0;
    __syncthreads();
    const volatile unsigned int* takePtr = &takeShared;
    const volatile unsigned int* countPtr = &countShared;
    const unsigned int exhaustedBase = (unsigned int)__cvta_generic_to_shared(exhaustedShared);
#define n (*countPtr)
#define exhausted (warpFlagLoad(exhaustedBase) != 0u)
#define take (*takePtr)
#define tail (take < 32u)
#define refillBelow (tail ? take : (unsigned int)st.traceMinActive)

    bool live = false;       // this lane holds a ray
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
        bool working = live && s.idx != 0u && (REDO || steps < st.traceBudget);
        unsigned int workMask = __ballot_sync(0xFFFFFFFFu, working);
        if (exhausted ? workMask == 0u : (unsigned int)__popc(workMask) < refillBelow) {
            // ---- retire finished rays, park the ones that ran out of budget
            const bool finished = live && s.idx == 0u;
            const bool park = !REDO && live && !finished && steps >= st.traceBudget;
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
                        accumulatePixel(st, slotPixel(st, slot), col.x, col.y, col.z); // col += p.color (kernels.cu:558)
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
                    st.pending[slot] = 2; // in flight AND parked: whoever continues the slot resumes from travS
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
                if (REDO && (mShade | mShadowDone)) atomicAdd(&ctl->redone, (unsigned long long)__popc(mShade | mShadowDone));
            }
            baseShade = __shfl_sync(0xFFFFFFFFu, baseShade, 0);
            basePark = __shfl_sync(0xFFFFFFFFu, basePark, 0);
            const unsigned int below = (1u << lane) - 1u;
            if (toShade) {
                shadeQ[baseShade + __popc(mShade & below)] = slot;
                st.ready[slot] = 1;
            }
            if (park) nextTrace[basePark + __popc(mPark & below)] = entry | ENTRY_RESUME;

            // ---- refill idle lanes, one atomic per warp
            if (!exhausted) {
                const unsigned int idle = ~workMask; // every lane that does not traverse has just been retired (or was empty)
                const unsigned int count = tail ? min((unsigned int)__popc(idle), take - (unsigned int)__popc(workMask)) : (unsigned int)__popc(idle);
                unsigned int base = 0;
                if (lane == 0) base = atomicAdd(REDO ? &ctl->redoCursor : &ctl->traceCursor, count);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (base + count >= n) warpFlagSet(exhaustedBase); // warp-uniform: the tail of the queue has been handed out
                const unsigned int rank = __popc(idle & below);
                const unsigned int i = base + rank;
                if (!live && rank < count && i < n) {
                    const unsigned int e = traceEntry(queue, st.traceCap, *(const volatile unsigned int*)&frontShared, i);
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
#undef n
#undef exhausted
#undef take
#undef tail
#undef refillBelow

// ------------------------------------------------------------- wide trace --
// The same job as traceKernel over the renderer's own wide tree (wide_traverse.cuh): persistent warps, idle lanes refilled
// from the queue with one atomic per warp, finished rays retired in batches. Every finished ray that found a triangle is
// certified against the caller's tree. What cannot be certified is appended to the NEXT iteration's trace queue with
// ENTRY_RESUME set; the lane that receives such an entry (or a ray the wide arithmetic does not cover) walks it right at the
// refill in the reference's order over the caller's tree (travRound), so its slot is simply one iteration late. (A separate
// launch for those ~14 rays per iteration put ~60 us of serial traversal latency on every iteration's critical path.)
// There is no step budget and no parking: the traversal stack lives in shared memory (wide.stackDepth entries per thread).
#ifndef WIDE_TRACE_BLOCK
#define WIDE_TRACE_BLOCK 128
#endif
#ifndef WIDE_TRACE_BLOCKS_PER_SM
#define WIDE_TRACE_BLOCKS_PER_SM 7 // 72 registers, no spills (256 x 4 = 64 registers spilled 24 bytes inside the node step; same speed)
#endif
#ifndef WIDE_ENDGAME
#define WIDE_ENDGAME 24u
#endif
#ifndef WIDE_ENDGAME_MIN
#define WIDE_ENDGAME_MIN 2u
#endif
#ifndef WIDE_TRACE_MIN_ACTIVE
#define WIDE_TRACE_MIN_ACTIVE 14 // refill when fewer lanes than this hold a ray: a refill pass costs ~300 instructions (measured: 14 < 20 < 24 < 28)
#endif
template <bool COUNT, int CUR, bool CERTIFY>
__global__ void __launch_bounds__(WIDE_TRACE_BLOCK, WIDE_TRACE_BLOCKS_PER_SM) wideTraceKernel(MeshState st, MeshView mesh, WideView wide) {
    constexpr int cur = CUR;
    extern __shared__ uint2 wideStackAll[];
    __shared__ RayCold coldAll[WIDE_TRACE_BLOCK];
    __shared__ float4 invAll[WIDE_TRACE_BLOCK]; // the reference's 1 / direction (wideSetup)
    RayCold& c = coldAll[threadIdx.x];
    float4& invT = invAll[threadIdx.x];
    uint2* stack = wideStackAll + threadIdx.x;
    MeshControl* ctl = st.ctl;
    const unsigned int* __restrict__ queue = st.traceQ[cur];
    unsigned int* __restrict__ shadeQ = st.shadeQ[cur];
    const unsigned int lane = laneId();
    // launch constants and the per-warp "queue exhausted" flag live in shared memory, read where they are used (the same
    // measure as in traceKernel: held in registers across the traversal loop they cost spills under the 64-register cap)
    __shared__ unsigned int takeShared, countShared, frontShared;
    __shared__ unsigned char exhaustedShared[WIDE_TRACE_BLOCK / 32];
    __shared__ unsigned int lastBaseShared[WIDE_TRACE_BLOCK / 32]; // the queue position this warp's last refill started at
    if (threadIdx.x == 0) {
        const unsigned int count = ctl->traceCount[cur] + ctl->traceBack[cur];
        const unsigned int totalWarps = gridDim.x * (WIDE_TRACE_BLOCK / 32);
        frontShared = ctl->traceCount[cur];
        countShared = count;
        takeShared = min(32u, max(1u, (count + totalWarps - 1) / totalWarps)); // short queue: spread the rays over all warps
    }
    if (threadIdx.x < WIDE_TRACE_BLOCK / 32) { exhaustedShared[threadIdx.x] = 0; lastBaseShared[threadIdx.x] = 0u; }
    __syncthreads();
    const volatile unsigned int* takePtr = &takeShared;
    const volatile unsigned int* countPtr = &countShared;
    const unsigned int exhaustedBase = (unsigned int)__cvta_generic_to_shared(exhaustedShared);
#define n (*countPtr)
#define exhausted (warpFlagLoad(exhaustedBase) != 0u)
#define take (*takePtr)
#define tail (take < 32u)
#define refillBelow (tail ? take : (unsigned int)st.traceMinActive)
    const unsigned int k3f = wideConst3F();
#ifdef WIDE_SMEM_TOP
    {   // the first nodes of the tree (breadth-first numbering: the root and the levels below it) copied into shared memory
        uint4* top = (uint4*)(wideStackAll + wide.stackDepth * WIDE_TRACE_BLOCK);
        for (unsigned int i = threadIdx.x; i < 6u * wide.topCount; i += WIDE_TRACE_BLOCK) top[i] = wide.nodes[i];
        __syncthreads();
    }
#endif

    bool live = false;
    WideRay r;
    WideTrav s;
    unsigned int nodeVisits = 0, triTests = 0;
    r.ox = r.oy = r.oz = r.ix = r.iy = r.iz = 0.0f; r.oct = 0u;
    s.ngx = s.ngy = s.tgx = s.tgy = 0u; s.sp = -1; s.closest = 0.0f;

    while (true) {
        bool working = live && s.sp >= 0;
        unsigned int workMask = __ballot_sync(0xFFFFFFFFu, working);
        if (exhausted ? workMask == 0u : (unsigned int)__popc(workMask) < refillBelow) {
            // ---- retire finished rays
            const bool finished = live && s.sp < 0;
            const bool isShadow = (r.oct & WIDE_FLAG_ANYHIT) != 0u;
            const unsigned int entry = __float_as_uint(c.rec.w);
            const unsigned int slot = entry & ENTRY_SLOT_MASK;
            bool redo = false;
            if (finished) {
                const unsigned int winner = __float_as_uint(c.rec.z);
                if (CERTIFY && winner != 0xFFFFFFFFu && !(r.oct & WIDE_FLAG_EXACT)) redo = !wideCertify(mesh, r, xyz(invT), c.dir.w, s.closest, winner);
                if (!redo) {
                    if (!isShadow) {
                        st.hit[slot] = make_float4(s.closest, c.rec.x, c.rec.y, c.rec.z);
                    } else {
                        // (the three records are requested together: which of the two colours is needed depends on the first)
                        const float4 l = st.shL[slot];
                        const float4 carried = st.shC[slot];
                        const float4 running = st.pcol[slot];
                        const bool unoccluded = !(s.closest < c.dir.w); // hit(...) false: p.color += p.lightContribution (kernels.cu:500-508)
                        if (__float_as_uint(l.w) & SHADOW_FLAG_FINAL) {
                            float4 col = carried;
                            if (unoccluded) { col.x += l.x; col.y += l.y; col.z += l.z; }
                            accumulatePixel(st, slotPixel(st, slot), col.x, col.y, col.z); // col += p.color (kernels.cu:558)
                        } else if (unoccluded) {
                            float4 col = running;
                            col.x += l.x; col.y += l.y; col.z += l.z;
                            st.pcol[slot] = col;
                        }
                        st.pending[slot] = 0;
                    }
                }
                live = false;
            }
            const bool toShade = finished && !redo && !isShadow;
            const unsigned int mShade = __ballot_sync(0xFFFFFFFFu, toShade);
            const unsigned int mShadowDone = __ballot_sync(0xFFFFFFFFu, finished && !redo && isShadow);
            const unsigned int mRedo = __ballot_sync(0xFFFFFFFFu, redo);
            const unsigned int mExactDone = __ballot_sync(0xFFFFFFFFu, finished && (r.oct & WIDE_FLAG_EXACT) != 0u);
            unsigned int baseShade = 0, baseRedo = 0;
            if (lane == 0) {
                if (mShade) { baseShade = atomicAdd(&ctl->shadeCount[cur], __popc(mShade)); atomicAdd(&ctl->raysExtend, (unsigned long long)__popc(mShade)); }
                if (mShadowDone) atomicAdd(&ctl->raysShadow, (unsigned long long)__popc(mShadowDone));
                if (mRedo) baseRedo = atomicAdd(&ctl->traceCount[cur ^ 1], __popc(mRedo));
                if (mExactDone) atomicAdd(&ctl->redone, (unsigned long long)__popc(mExactDone));
            }
            baseShade = __shfl_sync(0xFFFFFFFFu, baseShade, 0);
            baseRedo = __shfl_sync(0xFFFFFFFFu, baseRedo, 0);
            const unsigned int below = (1u << lane) - 1u;
            if (toShade) {
                shadeQ[baseShade + __popc(mShade & below)] = slot;
                st.ready[slot] = 1;
            }
            if (redo) st.traceQ[cur ^ 1][baseRedo + __popc(mRedo & below)] = entry | ENTRY_RESUME; // next iteration, in the reference's order

            // ---- refill idle lanes, one atomic per warp
            if (!exhausted) {
                const unsigned int idle = ~workMask;
                unsigned int count = tail ? min((unsigned int)__popc(idle), take - (unsigned int)__popc(workMask)) : (unsigned int)__popc(idle);
#ifndef WIDE_NO_ENDGAME
                {   // the end of the queue: when (by this warp's last look at the cursor) less than WIDE_ENDGAME rays per warp are left,
                    // a warp takes its share of what is left instead of filling every idle lane, so that the warps run out of work
                    // together and the launch does not wait for the few that took a full batch last
                    const volatile unsigned int* lastBase = &lastBaseShared[threadIdx.x >> 5];
                    const unsigned int left = n - min(n, *lastBase), warps = gridDim.x * (WIDE_TRACE_BLOCK / 32);
                    if (left < WIDE_ENDGAME * warps) count = min(count, max(WIDE_ENDGAME_MIN, left / warps));
                }
#endif
                unsigned int base = 0;
                if (lane == 0) base = atomicAdd(&ctl->traceCursor, count);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
#ifndef WIDE_NO_ENDGAME
                if (lane == 0) lastBaseShared[threadIdx.x >> 5] = base + count;
#endif
                if (base + count >= n) warpFlagSet(exhaustedBase); // warp-uniform: the tail of the queue has been handed out
                const unsigned int rank = __popc(idle & below);
                const unsigned int i = base + rank;
                bool exact = false, exactShadow = false; // this lane's ray is walked in the reference's order (below)
                if (!live && rank < count && i < n) {
                    const unsigned int e = traceEntry(queue, st.traceCap, *(const volatile unsigned int*)&frontShared, i);
                    const unsigned int sl = e & ENTRY_SLOT_MASK;
                    const bool shadow = (e & ENTRY_SHADOW) != 0u;
                    const float4 ro = shadow ? st.shO[sl] : st.rayO[sl];
                    const float4 rd = shadow ? st.shD[sl] : st.rayD[sl];
                    const float tMax = shadow ? rd.w : FLT_MAX;
                    const f3 d = unit(xyz(rd)); // hit(): ray(p.origin, dir) normalises again (kernels.cu:326)
                    c.dir = mk4(d, tMax);
                    c.rec = make_float4(0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu), __uint_as_float(e & ~ENTRY_RESUME));
                    f3 inv;
                    const bool covered = wideSetup(wide, r, xyz(ro), d, shadow, inv);
                    invT = mk4(inv, 0.0f);
                    wideStart(s, tMax);
                    if (!wideHitsBounds(mesh, r, inv, tMax)) { // hitMesh: scene bounds first (kernels.cu:297)
                        s.sp = -1;
                        s.closest = FLT_MAX;
                    } else if (!covered || (e & ENTRY_RESUME)) {
                        exact = true;
                        exactShadow = shadow;
                    }
                    live = true;
                }
                if (__any_sync(0xFFFFFFFFu, exact)) {
                    // the rays the certificate rejected last iteration, and rays outside the wide arithmetic's range: the caller's
                    // tree in the reference's order, by the lanes concerned, here and now (a handful per launch)
                    RayHot rh;
                    rh.ox = r.ox; rh.oy = r.oy; rh.oz = r.oz;
                    rh.ix = invT.x; rh.iy = invT.y; rh.iz = invT.z;
                    TravHot th;
                    th.idx = exact ? 1u : 0u; th.bitStack = 1u; th.closest = c.dir.w;
                    int steps = 0;
                    while (__any_sync(0xFFFFFFFFu, th.idx != 0u))
                        travRound<true>(mesh, rh, c, RT_EPSILON, exactShadow, th.idx != 0u, th, steps, 1, nodeVisits, triTests);
                    if (exact) {
                        s.closest = th.closest;
                        s.sp = -1;
                        r.oct |= WIDE_FLAG_EXACT;
                    }
                }
            }
            if (!__any_sync(0xFFFFFFFFu, live)) {
                if (exhausted) break;
                continue;
            }
            working = live && s.sp >= 0;
            workMask = __ballot_sync(0xFFFFFFFFu, working);
            if (workMask == 0u) continue; // e.g. every new ray missed the scene bounds: retire them
        }
        wideRound(wide, r, c, RT_EPSILON, working, s, stack, WIDE_TRACE_BLOCK, tail ? 1 : max(1, min(WIDE_NODE_QUORUM, __popc(workMask) >> 1)), k3f, nodeVisits, triTests);
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

#undef n
#undef exhausted
#undef take
#undef tail
#undef refillBelow

// ------------------------------------------------------------------- shade --
#ifndef SHADE_DENSE_FRACTION
#define SHADE_DENSE_FRACTION 4 // sweep all slots in slot order when more than 1/4 of them wait to be shaded
#endif
// Two ways to find the work. SPARSE: walk the shade queue (entries in the order rays happened to finish: every state access
// is a gather). DENSE, when most slots have an entry anyway (the bulk of a frame): sweep the slots in slot order and shade
// those whose `ready` flag is set -- the same set, but state loads and stores are coalesced, neighbouring lanes are
// neighbouring pixels (same materials, same texture lines), and the trace queue it writes comes out in slot order, so the
// next trace launch gets coalesced refills and rays of neighbouring pixels in one warp.
#ifndef SHADE_BLOCKS_PER_SM
#define SHADE_BLOCKS_PER_SM 4
#endif
__global__ void __launch_bounds__(WF_BLOCK, SHADE_BLOCKS_PER_SM) meshShadeKernel(MeshState st, ShadeScene sc, CameraDev cam, int cur) {
    MeshControl* ctl = st.ctl;
    const unsigned int queued = ctl->shadeCount[cur];
    const bool dense = queued > st.numSlots / SHADE_DENSE_FRACTION;
    const unsigned int n = dense ? st.numSlots : queued;
    const unsigned int* __restrict__ queue = st.shadeQ[cur];
    unsigned int* __restrict__ nextTrace = st.traceQ[cur ^ 1];
    unsigned int* __restrict__ nextShade = st.shadeQ[cur ^ 1];
    const unsigned int stride = gridDim.x * blockDim.x;
    unsigned int deferredCount = 0;
    __shared__ unsigned int appendS[WF_BLOCK / 32], appendE[WF_BLOCK / 32], appendC[WF_BLOCK / 32], appendBaseE, appendBaseS;
    for (unsigned int blockBase = blockIdx.x * blockDim.x; blockBase < n; blockBase += stride) { // (same trip count for every warp of the block)
        const unsigned int i = blockBase + threadIdx.x;
        bool traceNext = false, castsShadow = false, defer = false, cameraRay = false;
        unsigned int slot = 0;
        if (i < n) {
            // every load the entry needs is issued before the first one is looked at: `ready`, `pending` and the five state records
            // do not depend on each other (in the dense sweep the slot is the index itself), so the kernel waits for ONE memory
            // latency here instead of three in a row (flag -> flag -> state; the ncu source page showed the waits on exactly those)
            slot = dense ? i : queue[i];
            const unsigned char isReady = dense ? st.ready[i] : (unsigned char)1;
            const unsigned char isPending = st.pending[slot];
            const float4 h = st.hit[slot];
            const float4 ro = st.rayO[slot];
            const float4 rd = st.rayD[slot];
            const float4 att4 = st.atten[slot];
            const float4 pc4 = st.pcol[slot];
            if (!isReady) {
                // (dense sweep: the slot has no entry this iteration)
            } else if (isPending) {
                defer = true; // its shadow ray is still in flight: keep bounce order, come back next iteration
            } else {
                PathRegs p;
                p.origin = xyz(ro); p.dir = xyz(rd); p.att = xyz(att4);
                p.rng = __float_as_uint(ro.w);
                p.flags = __float_as_uint(rd.w);
                p.sample = __float_as_int(att4.w);
                p.color = xyz(pc4);
                ArraySink sink{st, slot};
                const ShadeResult r = shadePath(st, sc, cam, slot, p, h, sink);
                traceNext = r.traceNext;
                castsShadow = r.castsShadow;
                if (traceNext) storePath(st, slot, p, !r.continued);
                cameraRay = traceNext && !r.continued; // the slot's next sample starts: a camera ray
                st.ready[slot] = 0;
            }
        }
        const unsigned int posDefer = warpAppend(defer, &ctl->shadeCount[cur ^ 1]);
        if (defer) { nextShade[posDefer] = slot; deferredCount++; }
        // the entries of the BLOCK go out with one atomic per end of the trace queue: shadow entries to the back; to the front the
        // block's camera rays (first rays of new samples: neighbouring pixels, coherent walks) and then its bounce rays, so that a
        // refill of the trace kernel (14-18 lanes) gets rays of one kind (measured: camera rays first -1 %)
        const bool bounceRay = traceNext && !cameraRay;
        const unsigned int mE = __ballot_sync(0xFFFFFFFFu, bounceRay), mS = __ballot_sync(0xFFFFFFFFu, castsShadow), mC = __ballot_sync(0xFFFFFFFFu, cameraRay);
        const unsigned int warp = threadIdx.x >> 5;
        if (laneId() == 0) { appendS[warp] = __popc(mS); appendE[warp] = __popc(mE); appendC[warp] = __popc(mC); }
        __syncthreads();
        unsigned int totalS = 0, totalE = 0, totalC = 0, beforeS = 0, beforeE = 0, beforeC = 0;
        for (unsigned int k = 0; k < (blockDim.x >> 5); k++) {
            const unsigned int cs = appendS[k], ce = appendE[k], cc = appendC[k];
            if (k < warp) { beforeS += cs; beforeE += ce; beforeC += cc; }
            totalS += cs; totalE += ce; totalC += cc;
        }
        if (threadIdx.x == 0) { // extend entries at the front of the queue, shadow entries at its back (traceEntry)
            if (totalE + totalC) appendBaseE = atomicAdd(&ctl->traceCount[cur ^ 1], totalE + totalC);
            if (totalS) appendBaseS = atomicAdd(&ctl->traceBack[cur ^ 1], totalS);
        }
        __syncthreads();
        const unsigned int below = (1u << laneId()) - 1u;
        if (castsShadow) nextTrace[st.traceCap - 1u - (appendBaseS + beforeS + __popc(mS & below))] = slot | ENTRY_SHADOW;
        if (cameraRay) nextTrace[appendBaseE + beforeC + __popc(mC & below)] = slot; // the block's camera rays (neighbouring pixels) first
        if (bounceRay) nextTrace[appendBaseE + totalC + beforeE + __popc(mE & below)] = slot;
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
        if (ctl->traceCount[cur] | ctl->traceBack[cur] | ctl->shadeCount[cur]) ctl->iterations += 1;
        ctl->traceCount[cur] = 0;
        ctl->traceBack[cur] = 0;
        ctl->shadeCount[cur] = 0;
        ctl->traceCursor = 0;
        ctl->blocksDone = 0;
        ctl->redoCount = 0;
        ctl->redoCursor = 0;
    }
}

// ------------------------------------------------------------------ chaser --
// The frame's critical path is its heaviest pixels: one RNG stream per pixel means a pixel's bounces are strictly
// sequential (kernels.cu:542-548), and a pixel looking into glass needs ~10x the bounces of an average one. In the
// wavefront those pixels advance one bounce per iteration while iterations are long (~1.1 ms with ~1.6 M rays), and then
// drag a ~4000-iteration tail of nearly empty launches behind the bulk. So slots that fall behind (sample index well below
// the mean of all live slots) leave the wavefront for good: between two batches of iterations lanePartitionKernel moves
// them into a ring, and chaseKernel -- launched as a small wave after every hand-over, on high-priority streams beside the
// wavefront; a warp leaves as soon as the ring is empty and its slots are done -- takes them from there and runs each to its
// last sample without ever leaving the SM:
//   * a warp owns up to 16 slots; lane i (< 16) traces the slot's extend ray and shades, lane i + 16 traces its shadow ray,
//     so that both rays of a bounce are in flight at the same time exactly as in the wavefront;
//   * a bounce costs its own traversal plus its own shading: no launch, no queue, no step budget, and no waiting for the
//     longest ray of 8000 others (the wavefront express lane this replaces needed 190-250 us per bounce beside the bulk);
//   * the arithmetic is the same code (travRound, shadePath), the per-slot order of operations is the same (the next bounce
//     is shaded only after the previous bounce's shadow ray has been applied; a FINAL shadow ray adds the finished sample to
//     the pixel while the next sample is already being traced), so frames stay bit-identical.
// State hand-over: the wavefront's last writes to a moved slot precede the ring's `tail` update by a kernel boundary (and a
// wave is launched after every update, so no entry waits for a warp that has already left); the
// chaser reads the slot with L2 loads (__ldcg: its SM's L1 may hold lines from before) and never writes it back.
#ifndef CHASE_BLOCK
#define CHASE_BLOCK 64     // small blocks: a block gives its registers back when both its warps are done
#endif
#ifndef CHASE_MIN_BLOCKS
#define CHASE_MIN_BLOCKS 10 // register cap 96: no spills (at 12 blocks / 80 registers the wide chaser spilled 88 bytes)
#endif
#define CHASE_SLOTS_PER_WARP 16
#define CHASE_ENTRY_SHADE 0x40000000u // ring entry: the slot waits to be shaded (ENTRY_RESUME keeps its meaning: parked extend ray)
#define CHASE_ENTRY_DRAIN 0xC0000000u // ring entry: the slot has no path any more, only the FINAL shadow ray of its last sample

struct ChaseRing {
    // two rings: [0] shared (a warp takes up to 16 slots), [1] exclusive (the most lagging slots: one per warp, lowest latency)
    unsigned int* entries[2];        // numSlots entries each: a slot enters at most once per frame
    unsigned int* ctl;               // per ring r at ctl[8*r + ..]: [0] head (claimed) [1] tail (published) [2] reserved (appended) [4] slots finished
    unsigned long long* counters;    // [0] extend rays [1] shadow rays [2] node visits [3] triangle tests (counting builds) [4] rays re-traced exactly
};

struct ChasePathSm {   // a main lane's path between two bounces
    float4 oRng, dFlags, attSample, color;
};
struct ChaseShadowSm { // a partner lane's shadow ray
    float4 o, dDist, lFlags, carried;
};

struct ChaseSink {
    const MeshState& st;
    unsigned int slot;
    ChaseShadowSm* partner;
    __device__ __forceinline__ void castShadow(const f3& origin, const f3& dir, float lightDist, const f3& contribution, unsigned int flags, const f3& carried) {
        partner->o = mk4(origin, 0.0f);
        partner->dDist = mk4(dir, lightDist);
        partner->lFlags = mk4(contribution, __uint_as_float(flags));
        partner->carried = mk4(carried, 0.0f);
    }
    __device__ __forceinline__ void retire(const f3& c) { accumulatePixel(st, slotPixel(st, slot), c.x, c.y, c.z); }
};

__device__ __forceinline__ unsigned int ldVolatile(const unsigned int* p) { return *(const volatile unsigned int*)p; }

// One lane's traversal in the chaser: the wide walk with the certificate, and the order-exact walk (the caller's tree in
// the reference's order) either as the only traversal (WIDE = false) or, inline, for the rays the certificate rejects.
template <bool WIDE>
struct ChaseLane {
    // exact walk (always available)
    RayHot r;
    TravHot s;
    // wide walk
    WideRay wr;
    WideTrav ws;
    bool exactOnly; // wide: this ray is not covered by the wide arithmetic

    __device__ __forceinline__ void clear() {
        r.ox = r.oy = r.oz = r.ix = r.iy = r.iz = 0.0f;
        s.idx = 0u; s.bitStack = 0u; s.closest = 0.0f;
        wr.ox = wr.oy = wr.oz = wr.ix = wr.iy = wr.iz = 0.0f; wr.oct = 0u;
        ws.ngx = ws.ngy = ws.tgx = ws.tgy = 0u; ws.sp = -1; ws.closest = 0.0f;
        exactOnly = false;
    }
    // a fresh ray (hit(): the direction is normalised again, kernels.cu:326; hitMesh: scene bounds first, :297)
    __device__ __forceinline__ void start(const MeshView& mesh, const WideView& wide, RayCold& c, const f3& o, const f3& dRaw, float tMax, bool anyHit) {
        const f3 d = unit(dRaw);
        prepRay(r, c, o, d, tMax);
        c.rec = make_float4(0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu), 0.0f);
        const bool inBounds = rayHitsBounds(mesh, r, tMax);
        if (WIDE) {
            f3 inv;
            exactOnly = !wideSetup(wide, wr, o, d, anyHit, inv); // (inv == {r.ix, r.iy, r.iz}: prepRay made the same divisions)
            wideStart(ws, tMax);
            if (!inBounds || exactOnly) { ws.sp = -1; ws.closest = inBounds ? tMax : FLT_MAX; }
            if (!inBounds) exactOnly = false; // a miss of the scene bounds is final
        }
        s.idx = 1u; s.bitStack = 1u; s.closest = tMax;
        if (!inBounds) { s.idx = 0u; s.closest = FLT_MAX; }
    }
    // a ray the wavefront parked (exact pipeline only: the wide pipeline has no step budget)
    __device__ __forceinline__ void resume(RayCold& c, const f3& o, const f3& dRaw, float tMax, const uint2& t, float closest) {
        prepRay(r, c, o, unit(dRaw), tMax);
        s.idx = t.x; s.bitStack = t.y; s.closest = closest;
    }
    // nothing to trace: the record in c.rec / `closest` is already the answer
    __device__ __forceinline__ void done(float closest) {
        s.idx = 0u; s.bitStack = 0u; s.closest = closest;
        if (WIDE) { ws.sp = -1; ws.closest = closest; exactOnly = false; }
    }
    __device__ __forceinline__ bool working() const { return WIDE ? ws.sp >= 0 : s.idx != 0u; }
    __device__ __forceinline__ float closest() const { return WIDE ? ws.closest : s.closest; }
};

template <bool COUNT, bool WIDE>
__global__ void __launch_bounds__(CHASE_BLOCK, CHASE_MIN_BLOCKS) chaseKernel(MeshState st, MeshView mesh, WideView wide, ShadeScene sc, CameraDev cam, ChaseRing ring,
                                                                              unsigned int exclusiveEvery, unsigned int exclusivePairs) {
    extern __shared__ uint2 wideStackAll[];
    __shared__ RayCold coldAll[CHASE_BLOCK];
    __shared__ ChasePathSm pathAll[CHASE_BLOCK / 2];
    __shared__ ChaseShadowSm shadowAll[CHASE_BLOCK / 2];
    const unsigned int lane = laneId();
    const unsigned int warpInBlock = threadIdx.x >> 5;
    const bool exclusive = ((blockIdx.x * (CHASE_BLOCK / 32) + warpInBlock) % exclusiveEvery) == 0u;
    unsigned int* const rctl = ring.ctl + (exclusive ? 8 : 0);
    const unsigned int* const rentries = exclusive ? ring.entries[1] : ring.entries[0]; // (a select, not an index: an indexed read would copy the parameter struct to local memory)
    const unsigned int pairLimit = exclusive ? exclusivePairs : 0xFFFFu; // mask of the pairs this warp may fill
    const bool isMain = lane < CHASE_SLOTS_PER_WARP;
    const unsigned int pairIdx = warpInBlock * CHASE_SLOTS_PER_WARP + (lane & (CHASE_SLOTS_PER_WARP - 1u)); // shared by a main lane and its partner
    RayCold& c = coldAll[threadIdx.x];
    ChasePathSm& path = pathAll[pairIdx];
    ChaseShadowSm& shadow = shadowAll[pairIdx];
    uint2* stack = wideStackAll + threadIdx.x;
    const unsigned int k3f = wideConst3F();

    // lane state
    enum { IDLE = 0, TRACE = 1, READY = 2, DRAIN = 3 };
    //   main lane:    IDLE no slot | TRACE extend ray in flight | READY extend ray done, to be shaded | DRAIN no more samples, partner still busy
    //   partner lane: IDLE         | TRACE shadow ray in flight | READY shadow ray done, result not yet applied
    int state = IDLE;
    unsigned int slot = 0;
    bool unoccluded = false;
    ChaseLane<WIDE> t;
    t.clear();
    int steps = 0;
    unsigned int nodeVisits = 0, triTests = 0, doneExtend = 0, doneShadow = 0, redone = 0;

    while (true) {
        // ---- 1. idle pairs take slots from the ring (one compare-and-swap per warp)
        const unsigned int busyMask = __ballot_sync(0xFFFFFFFFu, state != IDLE);
        const unsigned int pairBusy = (busyMask | (busyMask >> CHASE_SLOTS_PER_WARP)) & 0xFFFFu; // a pair is free when both lanes are idle
        const unsigned int freePairs = ~pairBusy & pairLimit;
        unsigned int got = 0, base = 0;
        if (freePairs != 0u) {
            if (lane == 0) {
                const unsigned int tail = ldVolatile(&rctl[1]);
                const unsigned int head = ldVolatile(&rctl[0]);
                if (head < tail) {
                    const unsigned int want = min((unsigned int)__popc(freePairs), tail - head);
                    if (atomicCAS(&rctl[0], head, head + want) == head) { got = want; base = head; }
                    else got = 0xFFFFFFFFu; // lost the race: try again
                }
            }
            got = __shfl_sync(0xFFFFFFFFu, got, 0);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
        }
        const bool retry = got == 0xFFFFFFFFu;
        if (retry) got = 0;
        bool ingested = false;
        unsigned int entry = 0;
        if (got != 0u) {
            __threadfence(); // entries and slot state were published before `tail`
            const unsigned int rank = __popc(freePairs & ((1u << (lane & 15u)) - 1u));
            if (((freePairs >> (lane & 15u)) & 1u) && rank < got) {
                entry = __ldcg(&rentries[base + rank]);
                slot = entry & ENTRY_SLOT_MASK;
                ingested = true;
            }
        }
        if (ingested) {
            steps = 0;
            if (isMain) {
                const float4 ro = __ldcg(&st.rayO[slot]);
                const float4 rd = __ldcg(&st.rayD[slot]);
                path.oRng = ro;
                path.dFlags = rd;
                path.attSample = __ldcg(&st.atten[slot]);
                path.color = __ldcg(&st.pcol[slot]);
                if ((entry & CHASE_ENTRY_DRAIN) == CHASE_ENTRY_DRAIN) {
                    c.rec = make_float4(0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu), 0.0f);
                    t.done(FLT_MAX);
                    state = DRAIN;
                } else if (entry & CHASE_ENTRY_SHADE) {
                    const float4 h = __ldcg(&st.hit[slot]);
                    c.rec = make_float4(h.y, h.z, h.w, 0.0f);
                    t.done(h.x);
                    state = READY;
                } else if (!WIDE && (entry & ENTRY_RESUME)) { // (the wide pipeline parks nothing; a parked ray would simply start over)
                    const uint2 tr = __ldcg(&st.travE[slot]);
                    const float4 h = __ldcg(&st.hit[slot]);
                    t.resume(c, xyz(ro), xyz(rd), FLT_MAX, tr, h.x);
                    c.rec = make_float4(h.y, h.z, h.w, 0.0f);
                    state = TRACE;
                } else {
                    t.start(mesh, wide, c, xyz(ro), xyz(rd), FLT_MAX, false);
                    state = TRACE;
                }
            } else {
                const unsigned int pend = (unsigned int)__ldcg(&st.pending[slot]);
                if (pend != 0u) {
                    const float4 so = __ldcg(&st.shO[slot]);
                    const float4 sd = __ldcg(&st.shD[slot]);
                    shadow.o = so;
                    shadow.dDist = sd;
                    shadow.lFlags = __ldcg(&st.shL[slot]);
                    shadow.carried = __ldcg(&st.shC[slot]);
                    if (!WIDE && pend == 2u) {
                        const uint2 tr = __ldcg(&st.travS[slot]);
                        t.resume(c, xyz(so), xyz(sd), sd.w, tr, sd.w);
                        c.rec = make_float4(0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu), 0.0f);
                    } else {
                        t.start(mesh, wide, c, xyz(so), xyz(sd), sd.w, true);
                    }
                    state = TRACE;
                }
            }
        }

        // ---- 2. one scheduling round of traversal for every ray in flight
        const bool working = state == TRACE && t.working();
        if (__any_sync(0xFFFFFFFFu, working)) {
            if (WIDE) wideRound(wide, t.wr, c, RT_EPSILON, working, t.ws, stack, CHASE_BLOCK, 1, k3f, nodeVisits, triTests);
            else travRound<true>(mesh, t.r, c, RT_EPSILON, !isMain, working, t.s, steps, 1, nodeVisits, triTests);
        }
        if (WIDE) {
            // a finished wide walk must be certified; what is not (and what the wide arithmetic does not cover) is walked in
            // the reference's order right here, by the lanes concerned, before the ray counts as finished
            bool needExact = false;
            if (state == TRACE && t.ws.sp < 0) {
                const unsigned int winner = __float_as_uint(c.rec.z);
                needExact = t.exactOnly || (winner != 0xFFFFFFFFu && !wideCertify(mesh, t.wr, mk3(t.r.ix, t.r.iy, t.r.iz), c.dir.w, t.ws.closest, winner));
            }
            if (__any_sync(0xFFFFFFFFu, needExact)) {
                if (needExact) { // from the root (start() has already tested the scene bounds)
                    redone++;
                    c.rec = make_float4(0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu), 0.0f);
                    t.s.idx = 1u; t.s.bitStack = 1u; t.s.closest = c.dir.w;
                }
                while (__any_sync(0xFFFFFFFFu, needExact && t.s.idx != 0u))
                    travRound<true>(mesh, t.r, c, RT_EPSILON, !isMain, needExact && t.s.idx != 0u, t.s, steps, 1, nodeVisits, triTests);
                if (needExact) { t.ws.closest = t.s.closest; t.exactOnly = false; }
            }
        }

        // ---- 3. finished rays
        if (state == TRACE && !t.working()) {
            state = READY;
            if (isMain) doneExtend++;
            else { doneShadow++; unoccluded = !(t.closest() < c.dir.w); } // hit(...) false: p.color += p.lightContribution (kernels.cu:500-508)
        }

        // ---- 4. a main lane whose extend ray is done and whose partner is not tracing: apply the shadow result, then shade
        const unsigned int readyMask = __ballot_sync(0xFFFFFFFFu, state == READY);
        const unsigned int traceMask = __ballot_sync(0xFFFFFFFFu, state == TRACE);
        const unsigned int drainMask = __ballot_sync(0xFFFFFFFFu, state == DRAIN);
        const unsigned int unoccMask = __ballot_sync(0xFFFFFFFFu, unoccluded);
        const unsigned int partnerBit = 1u << ((lane & 15u) + CHASE_SLOTS_PER_WARP);
        const bool act = isMain && (((readyMask | drainMask) >> lane) & 1u) && !(traceMask & partnerBit);
        bool cast = false, freed = false;
        if (act) {
            if (readyMask & partnerBit) { // the previous bounce's shadow ray has finished
                const float4 l = shadow.lFlags;
                const bool lit = (unoccMask & partnerBit) != 0u;
                if (__float_as_uint(l.w) & SHADOW_FLAG_FINAL) {
                    float4 col = shadow.carried;
                    if (lit) { col.x += l.x; col.y += l.y; col.z += l.z; }
                    accumulatePixel(st, slotPixel(st, slot), col.x, col.y, col.z); // col += p.color (kernels.cu:558)
                } else if (lit) {
                    float4 col = path.color;
                    col.x += l.x; col.y += l.y; col.z += l.z;
                    path.color = col;
                }
            }
            if (state == DRAIN) {
                state = IDLE;
                freed = true;
            } else {
                PathRegs p;
                p.origin = xyz(path.oRng); p.dir = xyz(path.dFlags); p.att = xyz(path.attSample);
                p.rng = __float_as_uint(path.oRng.w);
                p.flags = __float_as_uint(path.dFlags.w);
                p.sample = __float_as_int(path.attSample.w);
                p.color = xyz(path.color);
                const float4 h = make_float4(t.closest(), c.rec.x, c.rec.y, c.rec.z);
                ChaseSink sink{st, slot, &shadow};
                const ShadeResult res = shadePath(st, sc, cam, slot, p, h, sink);
                cast = res.castsShadow;
                if (res.traceNext) {
                    path.oRng = mk4(p.origin, __uint_as_float(p.rng));
                    path.dFlags = mk4(p.dir, __uint_as_float(p.flags));
                    path.attSample = mk4(p.att, __int_as_float(p.sample));
                    path.color = mk4(p.color, 0.0f);
                    t.start(mesh, wide, c, p.origin, p.dir, FLT_MAX, false);
                    steps = 0;
                    state = TRACE;
                } else if (cast) {
                    state = DRAIN; // the last sample's FINAL shadow ray is still to be traced
                } else {
                    state = IDLE;
                    freed = true;
                }
            }
        }
        __syncwarp();
        // partners: the applied result is consumed; a freshly cast shadow ray starts
        const unsigned int actMask = __ballot_sync(0xFFFFFFFFu, act);
        const unsigned int castMask = __ballot_sync(0xFFFFFFFFu, cast);
        const unsigned int freedMask = __ballot_sync(0xFFFFFFFFu, freed);
        if (!isMain) {
            const unsigned int mainBit = 1u << (lane & 15u);
            if ((actMask & mainBit) && state == READY) { state = IDLE; unoccluded = false; }
            if (castMask & mainBit) {
                const float4 so = shadow.o;
                const float4 sd = shadow.dDist;
                t.start(mesh, wide, c, xyz(so), xyz(sd), sd.w, true);
                steps = 0;
                state = TRACE;
            }
        }
        if (freedMask != 0u && lane == 0) atomicAdd(&rctl[4], (unsigned int)__popc(freedMask));

        // ---- 5. nothing in flight and nothing to take: this warp is done (the next hand-over brings its own wave)
        if (!retry && __ballot_sync(0xFFFFFFFFu, state != IDLE) == 0u) {
            unsigned int empty = 0;
            if (lane == 0) empty = ldVolatile(&rctl[0]) >= ldVolatile(&rctl[1]) ? 1u : 0u;
            if (__shfl_sync(0xFFFFFFFFu, empty, 0)) break;
        }
    }

    for (int o = 16; o > 0; o >>= 1) {
        doneExtend += __shfl_xor_sync(0xFFFFFFFFu, doneExtend, o);
        doneShadow += __shfl_xor_sync(0xFFFFFFFFu, doneShadow, o);
        redone += __shfl_xor_sync(0xFFFFFFFFu, redone, o);
        if (COUNT) {
            nodeVisits += __shfl_xor_sync(0xFFFFFFFFu, nodeVisits, o);
            triTests += __shfl_xor_sync(0xFFFFFFFFu, triTests, o);
        }
    }
    if (lane == 0) {
        if (doneExtend) atomicAdd(&ring.counters[0], (unsigned long long)doneExtend);
        if (doneShadow) atomicAdd(&ring.counters[1], (unsigned long long)doneShadow);
        if (redone) atomicAdd(&ring.counters[4], (unsigned long long)redone);
        if (COUNT) {
            atomicAdd(&ring.counters[2], (unsigned long long)nodeVisits);
            atomicAdd(&ring.counters[3], (unsigned long long)triTests);
        }
    }
}

// Statistics of the live slots of the wavefront's input queues (index 0) for the hand-over that follows:
//   sums[0], sums[1]  sum and count of the slots' sample indices (their mean measures the frame's progress)
//   sums[2]           slots the chaser holds right now (one snapshot, so that every thread of the partition decides alike)
//   sums[3]           slots whose sample index is below factor * (mean of the previous hand-over, sums[4])
//   sums[4]           that mean, as float bits (written by laneCommitKernel; survives from hand-over to hand-over)
__global__ void laneStatsKernel(MeshState a, ChaseRing ring, unsigned long long* sums, float factor) {
    const unsigned int front = a.ctl->traceCount[0], n1 = front + a.ctl->traceBack[0], n2 = a.ctl->shadeCount[0];
    if (blockIdx.x == 0 && threadIdx.x == 0)
        sums[2] = (unsigned long long)(ring.ctl[2] - ldVolatile(&ring.ctl[4])) + (unsigned long long)(ring.ctl[8 + 2] - ldVolatile(&ring.ctl[8 + 4]));
    const float prevThreshold = factor * __uint_as_float((unsigned int)sums[4]);
    unsigned long long sum = 0, cnt = 0, under = 0;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
        const unsigned int entry = i < n1 ? traceEntry(a.traceQ[0], a.traceCap, front, i) : a.shadeQ[0][i - n1];
        if (entry & ENTRY_SHADOW) continue; // count every live slot once: by its extend entry or its deferred shade entry
        const int sample = __float_as_int(a.atten[entry & ENTRY_SLOT_MASK].w);
        sum += (unsigned long long)sample;
        cnt += 1;
        if ((float)sample < prevThreshold) under += 1;
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        under += __shfl_xor_sync(0xFFFFFFFFu, under, o);
    }
    if (laneId() == 0 && cnt) { atomicAdd(&sums[0], sum); atomicAdd(&sums[1], cnt); if (under) atomicAdd(&sums[3], under); }
}

// Splits the wavefront's input queues (index 0): lagging slots go to the chaser's rings, the rest to the spare queues
// (index 1). A slot lags when its sample index is below `factor` x the mean; when more slots lag than the chaser has room
// for, a hash of the slot number picks which of them go now (every entry of a slot reads the same per-slot values, so all
// entries of a slot take the same side). Slots below `exclusiveFactor` x the mean always go, into the exclusive ring.
// A slot enters a ring once, by its extend entry or its deferred shade entry; its shadow entry is dropped
// (the chaser finds the shadow ray's state in pending[]: 1 = not started, 2 = parked) unless the slot has nothing else
// left (SHADOW_FLAG_LAST).
__global__ void lanePartitionKernel(MeshState a, ChaseRing ring, const unsigned long long* sums, float factor, float exclusiveFactor, int minMean,
                                    unsigned int moveAllBelow, unsigned int capacity, unsigned int salt) {
    const unsigned int front = a.ctl->traceCount[0], n1 = front + a.ctl->traceBack[0], n2 = a.ctl->shadeCount[0];
    const float mean = sums[1] ? (float)((double)sums[0] / (double)sums[1]) : 0.0f;
    const bool moveAll = n1 + n2 <= moveAllBelow;
    const bool started = mean >= (float)minMean;
    const unsigned long long held = sums[2], under = sums[3];
    const unsigned long long room = held < (unsigned long long)capacity ? (unsigned long long)capacity - held : 0ull;
    const unsigned int admit = under <= room ? 0x1000000u : (unsigned int)((double)room / (double)under * 16777216.0); // of 2^24
    const float threshold = factor * mean, thresholdX = exclusiveFactor * mean;
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n1 + n2; base += stride) {
        const unsigned int i = base + laneId();
        const bool valid = i < n1 + n2;
        const bool isTrace = i < n1;
        unsigned int entry = 0;
        bool lag = false, lagX = false;
        if (valid) {
            entry = isTrace ? traceEntry(a.traceQ[0], a.traceCap, front, i) : a.shadeQ[0][i - n1];
            const unsigned int slot = entry & ENTRY_SLOT_MASK;
            const float sample = (float)__float_as_int(a.atten[slot].w);
            lagX = started && sample < thresholdX;
            lag = moveAll || lagX || (started && sample < threshold && (wangHash(slot ^ salt) & 0xFFFFFFu) < admit);
        }
        const bool isShadowEntry = isTrace && (entry & ENTRY_SHADOW) != 0u;
        // a shadow entry is all that is left of a slot whose last sample has ended: then IT takes the slot into the ring
        const bool lastShadow = valid && lag && isShadowEntry && (__float_as_uint(a.shL[entry & ENTRY_SLOT_MASK].w) & SHADOW_FLAG_LAST) != 0u;
        const bool toRing = valid && lag && (!isShadowEntry || lastShadow);
        const unsigned int ringEntry = lastShadow ? ((entry & ENTRY_SLOT_MASK) | CHASE_ENTRY_DRAIN) : (isTrace ? entry : (entry | CHASE_ENTRY_SHADE));
        unsigned int pos;
        pos = warpAppend(toRing && !lagX, &ring.ctl[2]);
        if (toRing && !lagX) ring.entries[0][pos] = ringEntry;
        pos = warpAppend(toRing && lagX, &ring.ctl[8 + 2]);
        if (toRing && lagX) ring.entries[1][pos] = ringEntry;
        if (toRing && !isTrace) a.ready[entry & ENTRY_SLOT_MASK] = 0; // the wavefront's dense shade sweep must not see the slot any more
        pos = warpAppend(valid && isTrace && !lag && !isShadowEntry, &a.ctl->traceCount[1]);
        if (valid && isTrace && !lag && !isShadowEntry) a.traceQ[1][pos] = entry;
        pos = warpAppend(valid && isTrace && !lag && isShadowEntry, &a.ctl->traceBack[1]); // shadow entries stay at the back
        if (valid && isTrace && !lag && isShadowEntry) a.traceQ[1][a.traceCap - 1u - pos] = entry;
        pos = warpAppend(valid && !isTrace && !lag, &a.ctl->shadeCount[1]);
        if (valid && !isTrace && !lag) a.shadeQ[1][pos] = entry;
    }
}

__global__ void laneCopyBackKernel(MeshState a) {
    const unsigned int front = a.ctl->traceCount[1], n1 = front + a.ctl->traceBack[1], n2 = a.ctl->shadeCount[1];
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
        if (i < n1) {
            const unsigned int at = i < front ? i : a.traceCap - 1u - (i - front); // both ends keep their places
            a.traceQ[0][at] = a.traceQ[1][at];
        } else {
            a.shadeQ[0][i - n1] = a.shadeQ[1][i - n1];
        }
    }
}

// Makes the split visible: the wavefront continues with the spare queues' contents, the chaser sees the new ring entries
// (everything the earlier kernels of this stream wrote precedes the `tail` store).
__global__ void laneCommitKernel(MeshControl* ctl, ChaseRing ring, unsigned long long* sums) {
    sums[4] = (unsigned long long)__float_as_uint(sums[1] ? (float)((double)sums[0] / (double)sums[1]) : 0.0f);
    ctl->traceCount[0] = ctl->traceCount[1];
    ctl->traceBack[0] = ctl->traceBack[1];
    ctl->shadeCount[0] = ctl->shadeCount[1];
    ctl->traceCount[1] = 0;
    ctl->traceBack[1] = 0;
    ctl->shadeCount[1] = 0;
    __threadfence();
    *(volatile unsigned int*)&ring.ctl[1] = ring.ctl[2];
    *(volatile unsigned int*)&ring.ctl[8 + 1] = ring.ctl[8 + 2];
}
