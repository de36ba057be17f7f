// spheres_path.cuh -- the README-era random-spheres scene (BASELINE configs 1 and 2) on the wavefront kernels.
//
// The reference's HEAD has no sphere renderer: only the pieces survive --
// sphereHit (intersections.h:85-104), material_scatter (scene_materials.h:13-20), get_ray with a lens
// (camera.h:8-12), rnd.h, the commented sky gradient (kernels.cu:419-421) and README.md:93-104's
// "spheres in constant memory".  This file assembles exactly those pieces around color()'s loop
// (kernels.cu:396-533): per-bounce closest sphere, normal (p - c)/r flipped towards the ray,
// material_scatter, origin += t * rayDir, Russian roulette after bounce 3; a miss adds
// attenuation * gradient(rayDir.y) and ends the path.  No light, no shadow rays.
// oracle/ref_spheres.cu is the same definition as a thread-per-pixel megakernel built from the
// reference's own headers; tests compare the two.
// An iteration is two launches: extendSpheresBvhKernel (closest sphere) and shadeSpheresKernel (scatter, retire, next camera
// ray, queue swap).
// Included at the end of renderer.cu (one translation unit: the kernels of wavefront_kernels.cuh are shared).
#pragma once

#define MAX_SPHERES 1024

__constant__ float4 c_spheres[MAX_SPHERES]; // {center.xyz, radius}: 16 KB, read with a warp-uniform index

// Closest sphere: the brute-force loop every lane walks in lock step (constant-cache broadcast).
__global__ void __launch_bounds__(WF_BLOCK) extendSpheresKernel(WfState st, const unsigned int* __restrict__ queue, int numSpheres) {
    WfControl* ctl = st.ctl;
    const unsigned int n = ctl->countActive;
    while (true) {
        unsigned int base = 0;
        if (laneId() == 0) base = atomicAdd(&ctl->cursorExtend, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const unsigned int i = base + laneId();
        if (i < n) {
            const unsigned int slot = queue[i];
            const float4 ro = st.rayO[slot];
            const float4 rd = st.rayD[slot];
            const f3 o = xyz(ro);
            const f3 d = unit(xyz(rd));
            float closest = FLT_MAX;
            unsigned int id = 0xFFFFFFFFu;
            for (int s = 0; s < numSpheres; s++) {
                const float4 sp = c_spheres[s];
                const float t = sphereHitT(xyz(sp), sp.w, o, d, RT_EPSILON, closest);
                if (t < closest) {
                    closest = t;
                    id = (unsigned int)s;
                }
            }
            st.hit[slot] = make_float4(closest, 0.0f, 0.0f, __uint_as_float(id));
        }
    }
}

// ---- sphere BVH ----------------------------------------------------------------------------------------------------
// 488 tests per ray is what the README-era kernel did; the result of that loop is `min over spheres of r_s`, where r_s is
// the sphere's first root in (t_min, inf) (sphereHit, intersections.h:85-104, tries the near root, then the far one), ties
// going to the lowest index (the loop accepts on strict `<` in index order). That is a property of the set of spheres, not
// of the loop, so any traversal that (a) never skips a sphere whose r_s could win and (b) breaks ties by index returns the
// same bits. The spheres are put into a small BVH at init (host, median split, one sphere per leaf, boxes padded by 1 % of the
// radius + 1e-3: three orders of magnitude more than the rounding of r_s for these sizes); spheres too large for that
// margin (radius > 100: the ground sphere, whose roots lose ~1e-3 to cancellation) stay in a list that every ray tests.
// The tree is stored depth-first with skip links (no stack): node i's subtree is [i+1, skip_i) -- once per ray octant, each
// threading visiting the child on the ray's near side of the split axis first, so that a hit found early culls the far side
// (8 x 243 nodes x 32 B: L1-resident).
struct SphereBvh {
    const float4* __restrict__ nodes;   // 2 per node: {bmin.xyz, skip}{bmax.xyz, first | count << 24 (0xFFFFFFFF = internal)}
    const float4* __restrict__ spheres; // leaf order: {center.xyz, radius}
    const unsigned int* __restrict__ ids; // leaf order -> index the caller gave the sphere
    unsigned int numNodes;
    unsigned int numAlways;             // spheres[0 .. numAlways) are tested by every ray
    unsigned int octantMask;            // 7: nodes[] holds 8 threadings of the tree, one per ray octant (near child first); 0: one
};

__device__ __forceinline__ void testSphere(const SphereBvh& b, unsigned int k, const f3& o, const f3& d, float& closest, unsigned int& id) {
    const float4 sp = __ldg(b.spheres + k);
    const float t = sphereHitT(xyz(sp), sp.w, o, d, RT_EPSILON, FLT_MAX); // r_s: does not depend on the current closest
    if (t < FLT_MAX) {
        const unsigned int s = __ldg(b.ids + k);
        if (t < closest || (t == closest && s < id)) {
            closest = t;
            id = s;
        }
    }
}

// The closest sphere along (o, d): {t, id} (FLT_MAX / ~0 = none).
__device__ __forceinline__ void closestSphere(const SphereBvh& bvh, const f3& o, const f3& d, float& closest, unsigned int& id, unsigned int& boxTests,
                                              unsigned int& sphereTests) {
    // (direction components below 1e-20 are clamped, keeping their sign: no infinite reciprocal, so no inf - inf in the fused slab
    // test below; the boxes' padding times 1e20 dwarfs what the clamp moves)
    const f3 dc = mk3(fabsf(d.x) < 1e-20f ? copysignf(1e-20f, d.x) : d.x, fabsf(d.y) < 1e-20f ? copysignf(1e-20f, d.y) : d.y,
                      fabsf(d.z) < 1e-20f ? copysignf(1e-20f, d.z) : d.z);
    const f3 inv = mk3(1.0f / dc.x, 1.0f / dc.y, 1.0f / dc.z);
    const f3 oi = mk3(o.x * inv.x, o.y * inv.y, o.z * inv.z);
    closest = FLT_MAX;
    id = 0xFFFFFFFFu;
    for (unsigned int k = 0; k < bvh.numAlways; k++) testSphere(bvh, k, o, d, closest, id);
    sphereTests += bvh.numAlways;
    const unsigned int octant = ((dc.x < 0.0f ? 1u : 0u) | (dc.y < 0.0f ? 2u : 0u) | (dc.z < 0.0f ? 4u : 0u)) & bvh.octantMask;
    const float4* __restrict__ nodes = bvh.nodes + (size_t)octant * 2u * bvh.numNodes;
    unsigned int node = 0;
    while (true) {
        // box phase: every lane walks on until it stands in a leaf (or its walk ends); the leaves' sphere tests then run for all
        // those lanes together instead of one lane at a time inside the walk
        unsigned int leaf = 0u; // count << 24 | first; 0 = none
        while (node < bvh.numNodes) {
            float4 lo, hi;
            ldg256(nodes + 2 * node, lo, hi); // one 256-bit load per node
            boxTests++;
            // conservative slab test, plane * (1/d) - origin * (1/d) as one fused multiply-add per plane: against (plane - origin) / d
            // it errs by |origin / d| * 2^-23 at most, 1/100 or less of what the boxes' padding (1e-3 + 1e-5 |centre|, times 1/d) allows;
            const float x0 = __fmaf_rn(lo.x, inv.x, -oi.x), x1 = __fmaf_rn(hi.x, inv.x, -oi.x);
            const float y0 = __fmaf_rn(lo.y, inv.y, -oi.y), y1 = __fmaf_rn(hi.y, inv.y, -oi.y);
            const float z0 = __fmaf_rn(lo.z, inv.z, -oi.z), z1 = __fmaf_rn(hi.z, inv.z, -oi.z);
            const float tEnter = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
            const float tExit = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
            const bool hitBox = tEnter <= tExit * 1.00001f + 1e-5f && tEnter <= closest;
            const unsigned int word = __float_as_uint(hi.w);
            node = hitBox ? node + 1 : __float_as_uint(lo.w); // into the subtree, or past it
            if (hitBox && word != 0xFFFFFFFFu) {
                leaf = word;
                break;
            }
        }
        if (leaf == 0u) break;
        const unsigned int first = leaf & 0xFFFFFFu, count = leaf >> 24;
        testSphere(bvh, first, o, d, closest, id); // (leaves hold one sphere unless CRT_SPHERES_LEAF says otherwise)
        if (count > 1u) for (unsigned int k = 1; k < count; k++) testSphere(bvh, first + k, o, d, closest, id);
        sphereTests += count;
    }
}

__device__ __forceinline__ void closestSphere(const SphereBvh& bvh, const f3& o, const f3& d, float& closest, unsigned int& id) {
    unsigned int boxTests = 0, sphereTests = 0; // (never read: the compiler drops the counting)
    closestSphere(bvh, o, d, closest, id, boxTests, sphereTests);
}

__global__ void __launch_bounds__(WF_BLOCK) extendSpheresBvhKernel(WfState st, const unsigned int* __restrict__ queue, SphereBvh bvh) {
    WfControl* ctl = st.ctl;
    const unsigned int n = ctl->countActive;
    while (true) {
        unsigned int base = 0;
        if (laneId() == 0) base = atomicAdd(&ctl->cursorExtend, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const unsigned int i = base + laneId();
        if (i < n) {
            const unsigned int slot = queue[i];
            const float4 ro = st.rayO[slot];
            const float4 rd = st.rayD[slot];
            float closest;
            unsigned int id;
            closestSphere(bvh, xyz(ro), unit(xyz(rd)), closest, id);
            st.hit[slot] = make_float4(closest, 0.0f, 0.0f, __uint_as_float(id));
        }
    }
}

// ---- one path, in registers ------------------------------------------------------------------------------------------
// color()'s loop body (kernels.cu:402-531, without light and shadow rays) on a path held in registers; the wavefront kernels
// load / store it around these two functions, the persistent kernel keeps it in registers for the whole pixel.
struct SpherePath {
    f3 origin, dir, att, col;
    unsigned int rng, flags; // flags: PATH_BOUNCE_MASK | PATH_FLAG_*
};

// The next camera ray of `pixel` from the path's RNG stream (kernels.cu:549-555).
__device__ __forceinline__ void startSpherePath(SpherePath& p, const CameraDev& cam, int nx, int ny, unsigned int pixel) {
    const int px = (int)(pixel % (unsigned int)nx), py = (int)(pixel / (unsigned int)nx);
    const float u = float(px + rnd(p.rng)) / float(nx);
    const float v = float(py + rnd(p.rng)) / float(ny);
    cameraRay(cam, u, v, p.rng, p.origin, p.dir);
    p.flags = 0u; // bounce 0, specular = inside = false (kernels.cu:554-555)
    p.att = mk3(1.0f, 1.0f, 1.0f);
    p.col = mk3(0.0f, 0.0f, 0.0f);
}

// Everything between two closest-sphere queries of one path whose hit record is `h`: sky gradient on a miss, scatter, Russian
// roulette. Returns whether the path has another ray to trace; when it has not, p.col is the sample's colour.
// `rdir` = unit(p.dir), the direction the ray was traced along (the caller has it already).
__device__ __forceinline__ bool scatterSpherePath(SpherePath& p, const float4* __restrict__ mats, int maxDepth, const float4& h, const f3& rdir) {
    if (!(h.x < FLT_MAX)) {
        // sky gradient, kernels.cu:419-421
        const float t = 0.5f * (p.dir.y + 1.0f);
        const f3 c = (1.0f - t) * mk3(1.0f, 1.0f, 1.0f) + t * mk3(0.5f, 0.7f, 1.0f);
        const f3 add = p.att * c;
        p.col.x += add.x; p.col.y += add.y; p.col.z += add.z;
        return false;
    }
    bool inside = (p.flags & PATH_FLAG_INSIDE) != 0u;
    unsigned int bounce = p.flags & PATH_BOUNCE_MASK;
    const unsigned int id = __float_as_uint(h.w);
    const float4 sp = c_spheres[id];
    const f3 hp = p.origin + h.x * rdir; // point_at_parameter on the traced (normalised) ray
    SurfacePoint s;
    s.normal = (hp - xyz(sp)) / sp.w;
    s.t = h.x;
    s.inside = inside;
    if (dot(rdir, s.normal) > 0.0f) s.normal = -s.normal;
    const float4 m0 = __ldg(mats + 2 * id);
    const float4 m1 = __ldg(mats + 2 * id + 1);
    Scatter scat;
    scat.specular = false;
    scat.throughput = mk3(1.0f, 1.0f, 1.0f);
    scat.refracted = false;
    scat.t = h.x;
    scat.wi = mk3(0.0f, 0.0f, 0.0f);
    materialScatter(scat, s, p.dir, __float_as_int(m1.x), m0.w, xyz(m0), p.rng);
    p.origin = p.origin + scat.t * p.dir;
    p.dir = scat.wi;
    p.att = p.att * scat.throughput;
    inside = scat.refracted ? !inside : inside;
    bool continues = true;
    if (bounce > 3u) {
        const float m = maxcomp(p.att);
        if (rnd(p.rng) > m) continues = false;
        else p.att = p.att * (1 / m);
    }
    if (continues) {
        bounce = (bounce + 1u) & PATH_BOUNCE_MASK;
        if (!((int)bounce < maxDepth)) continues = false;
    }
    p.flags = bounce | (scat.specular ? PATH_FLAG_SPECULAR : 0u) | (inside ? PATH_FLAG_INSIDE : 0u);
    return continues;
}

// Shade + retire + regenerate + advance in one launch (an iteration is extend, then this): when a path ends its colour goes
// into the pixel (col += p.color, kernels.cu:558, in sample order: one slot per pixel) and the slot's next sample starts
// right here (kernels.cu:549-555) instead of in a separate raygen pass; the last block to finish swaps the queues.
// Returns whether the slot has another ray to trace. Shared by shadeSpheresKernel and finishSpheresKernel.
__device__ __forceinline__ bool shadeSphereSlot(const WfState& st, const float4* __restrict__ mats, int maxDepth, const CameraDev& cam, int nx, int ny,
                                                int samplesPerSlot, int slotsPerPixel, unsigned int npix, unsigned int slot, const float4& h) {
    const float4 ro = st.rayO[slot];
    const float4 rd = st.rayD[slot];
    const float4 att4 = st.atten[slot];
    const float4 pc = st.pcol[slot];
    SpherePath p;
    p.origin = xyz(ro); p.dir = xyz(rd); p.att = xyz(att4); p.col = xyz(pc);
    p.rng = __float_as_uint(ro.w);
    p.flags = __float_as_uint(rd.w);
    int sample = __float_as_int(att4.w);
    bool continues = scatterSpherePath(p, mats, maxDepth, h, unit(p.dir));
    if (!continues) {
        // the sample is finished: col += p.color (kernels.cu:558), then the slot's next sample
        const unsigned int pixel = slot % npix;
        if (slotsPerPixel == 1) {
            float4 a = st.accum[pixel];
            a.x += p.col.x; a.y += p.col.y; a.z += p.col.z;
            st.accum[pixel] = a;
        } else {
            atomicAdd(&st.accum[pixel].x, p.col.x);
            atomicAdd(&st.accum[pixel].y, p.col.y);
            atomicAdd(&st.accum[pixel].z, p.col.z);
        }
        sample += 1;
        if (sample < samplesPerSlot) {
            startSpherePath(p, cam, nx, ny, pixel);
            st.pcol[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            continues = true;
        }
    }
    if (continues) {
        st.rayO[slot] = mk4(p.origin, __uint_as_float(p.rng));
        st.rayD[slot] = mk4(p.dir, __uint_as_float(p.flags));
        st.atten[slot] = mk4(p.att, __int_as_float(sample));
    }
    return continues;
}

__global__ void __launch_bounds__(WF_BLOCK) shadeSpheresKernel(WfState st, const float4* __restrict__ mats, int maxDepth,
                                                               const unsigned int* __restrict__ queue, unsigned int* __restrict__ nextQueue,
                                                               CameraDev cam, int nx, int ny, int samplesPerSlot, int slotsPerPixel) {
    WfControl* ctl = st.ctl;
    const unsigned int n = ctl->countActive;
    const unsigned int npix = (unsigned int)nx * (unsigned int)ny;
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
        const unsigned int i = base + laneId();
        bool continues = false;
        unsigned int slot = 0;
        if (i < n) {
            slot = queue[i];
            continues = shadeSphereSlot(st, mats, maxDepth, cam, nx, ny, samplesPerSlot, slotsPerPixel, npix, slot, st.hit[slot]);
        }
        const unsigned int posNext = warpAppend(continues, &ctl->countNext);
        if (continues) nextQueue[posNext] = slot;
    }

    // the last block to finish advances the iteration (what advanceKernel did)
    __shared__ bool isLast;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        isLast = atomicAdd(&ctl->blocksDone, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (isLast && threadIdx.x == 0) {
        ctl->raysExtend += ctl->countActive;
        if (ctl->countActive) ctl->iterations += 1;
        ctl->countActive = ctl->countNext;
        ctl->countNext = 0;
        ctl->cursorExtend = 0;
        ctl->blocksDone = 0;
    }
}

// The tail of a frame: when few slots are left (the heaviest pixels: every pixel's samples are one sequential chain,
// kernels.cu:542-548), iterating two launches per bounce over a nearly empty queue is mostly launch latency (1011 iterations for
// a 220-iteration bulk). One thread per remaining slot then runs the slot to its last sample with the same two device functions.
__global__ void __launch_bounds__(128) finishSpheresKernel(WfState st, const float4* __restrict__ mats, int maxDepth, const unsigned int* __restrict__ queue,
                                                           SphereBvh bvh, CameraDev cam, int nx, int ny, int samplesPerSlot, int slotsPerPixel) {
    WfControl* ctl = st.ctl;
    const unsigned int n = ctl->countActive;
    const unsigned int npix = (unsigned int)nx * (unsigned int)ny;
    unsigned long long rays = 0;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned int slot = queue[i];
        bool continues = true;
        while (continues) {
            const float4 ro = st.rayO[slot];
            const float4 rd = st.rayD[slot];
            float closest;
            unsigned int id;
            closestSphere(bvh, xyz(ro), unit(xyz(rd)), closest, id);
            rays++;
            continues = shadeSphereSlot(st, mats, maxDepth, cam, nx, ny, samplesPerSlot, slotsPerPixel, npix, slot, make_float4(closest, 0.0f, 0.0f, __uint_as_float(id)));
        }
    }
    for (int o = 16; o > 0; o >>= 1) rays += __shfl_xor_sync(0xFFFFFFFFu, rays, o);
    if (laneId() == 0 && rays) atomicAdd(&ctl->raysExtend, rays);
}

__global__ void finishSpheresDoneKernel(WfControl* ctl) { ctl->countActive = 0; }

// ---- the persistent kernel ---------------------------------------------------------------------------------------------
// In this scene a path slot ALWAYS has a next ray until its pixel is finished (no shadow rays, a finished sample starts the
// next one), so the wavefront's queues have nothing to compact but finished pixels -- and every bounce pays a round trip of
// the path state through memory, two launches, and a tail of 1 011 nearly empty iterations for the heaviest pixel (its samples
// are one sequential chain, kernels.cu:542-548). Here a LANE owns a work item (one slot of one pixel) from its first camera
// ray to its last sample with the path in registers, and takes the next item from a global cursor the moment it is done:
// lanes never wait for a pixel other than their own, the frame ends when the last started item ends. Same device functions,
// same per-pixel order of every floating-point operation as the wavefront kernels (tests compare the frames bit for bit).
#define SPH_MEGA_BLOCK 128 // (64 x 16, 256 x 4 and 32 x 32 measure the same; 128 x 6 with 80 registers is 5 % slower)
#define SPH_MEGA_BLOCKS_PER_SM 8
#ifndef SPH_ITEM_CHUNK
#define SPH_ITEM_CHUNK 32u // (config 2: 8 -> 28.4 ms, 16 -> 26.8, 32 -> 26.9, 64 -> 29.5, 128 -> 33.3; one atomic per need: 28.7)
#endif

// The shape of closestSphere's loops matters more than their content: the same walk written as "at most N box steps, then look
// what the lane needs" (a counted inner loop) ran at HALF the speed (73 ms) -- ptxas then no longer re-joins the lanes before
// the leaf tests. Keep the plain nested loops.
// (Measured and not kept, profiles/r02/ab_summary.json: a per-lane state machine that runs, each turn, whichever piece -- box
// steps, leaf tests, shading -- most lanes of the warp wait for. Lane use rises, the scheduling ballots and the wider live state
// cost more than that returns: 42-45 ms against 34 ms for the plain loop below.)
template <bool COUNT>
__global__ void __launch_bounds__(SPH_MEGA_BLOCK, SPH_MEGA_BLOCKS_PER_SM)
spheresMegaKernel(WfState st, const float4* __restrict__ mats, int maxDepth, SphereBvh bvh, CameraDev cam, int nx, int ny, int samplesPerSlot,
                  int slotsPerPixel, unsigned int streamBase, unsigned int numItems) {
    WfControl* ctl = st.ctl;
    const unsigned int npix = (unsigned int)nx * (unsigned int)ny;
    SpherePath p;
    f3 sum = mk3(0.0f, 0.0f, 0.0f);
    unsigned int pixel = 0;
    int sample = 0;
    bool live = false, exhausted = false;
    unsigned int poolBase = 0, poolLeft = 0; // the warp's pool of work items (warp-uniform)
    unsigned int rays = 0, trips = 0;
    unsigned long long boxTests = 0, sphereTests = 0; // COUNT only
    while (true) {
        // work items come from the global cursor SPH_ITEM_CHUNK at a time into a pool of the warp; lanes draw from the pool with a
        // ballot (at 10 spp some lane of a warp finishes a pixel in almost every turn: one global atomic per turn, with its round
        // trip in front of every lane, halved the ray rate of config 1)
        if (!exhausted || poolLeft != 0u) {
            unsigned int need = __ballot_sync(0xFFFFFFFFu, !live);
            while (need != 0u) {
                if (poolLeft == 0u) {
                    if (exhausted) break;
                    unsigned int base = 0;
                    if (laneId() == 0) base = atomicAdd(&ctl->cursorExtend, SPH_ITEM_CHUNK);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    exhausted = base + SPH_ITEM_CHUNK >= numItems;
                    if (base >= numItems) break;
                    poolBase = base;
                    poolLeft = min(SPH_ITEM_CHUNK, numItems - base);
                }
                const unsigned int take = min((unsigned int)__popc(need), poolLeft);
                const unsigned int rank = __popc(need & ((1u << laneId()) - 1u));
                if (!live && rank < take) {
                    const unsigned int item = poolBase + rank;
                    live = true;
                    pixel = item % npix;
                    const unsigned int stream = streamBase * (unsigned int)slotsPerPixel + item / npix;
                    p.rng = pathSeed(pixel + stream * npix); // kernels.cu:541-542 (stream 0)
                    sample = 0;
                    sum = mk3(0.0f, 0.0f, 0.0f);
                    startSpherePath(p, cam, nx, ny, pixel);
                }
                poolBase += take;
                poolLeft -= take;
                need = __ballot_sync(0xFFFFFFFFu, !live);
            }
        }
        if (!__any_sync(0xFFFFFFFFu, live)) break;
        trips++;
        if (live) {
            float closest;
            unsigned int id;
            const f3 rdir = unit(p.dir); // once per ray: the walk and the shading both use it
            if (COUNT) {
                unsigned int nb = 0, ns = 0;
                closestSphere(bvh, p.origin, rdir, closest, id, nb, ns);
                boxTests += nb;
                sphereTests += ns;
            } else {
                closestSphere(bvh, p.origin, rdir, closest, id);
            }
            rays++;
            p.col = mk3(0.0f, 0.0f, 0.0f); // a sphere path only collects light when it ends (the sky): nothing to carry between rays
            if (!scatterSpherePath(p, mats, maxDepth, make_float4(closest, 0.0f, 0.0f, __uint_as_float(id)), rdir)) {
                sum.x += p.col.x; sum.y += p.col.y; sum.z += p.col.z; // col += p.color, in sample order (kernels.cu:558)
                sample++;
                if (sample < samplesPerSlot) {
                    startSpherePath(p, cam, nx, ny, pixel);
                } else {
                    if (slotsPerPixel == 1) {
                        st.accum[pixel] = make_float4(sum.x, sum.y, sum.z, 0.0f);
                    } else {
                        atomicAdd(&st.accum[pixel].x, sum.x);
                        atomicAdd(&st.accum[pixel].y, sum.y);
                        atomicAdd(&st.accum[pixel].z, sum.z);
                    }
                    live = false;
                }
            }
        }
    }
    unsigned long long warpRays = rays;
    for (int o = 16; o > 0; o >>= 1) warpRays += __shfl_xor_sync(0xFFFFFFFFu, warpRays, o);
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) {
            boxTests += __shfl_xor_sync(0xFFFFFFFFu, boxTests, o);
            sphereTests += __shfl_xor_sync(0xFFFFFFFFu, sphereTests, o);
        }
        if (laneId() == 0) {
            atomicAdd(&ctl->nodeVisits, boxTests);
            atomicAdd(&ctl->triTests, sphereTests);
        }
    }
    if (laneId() == 0) {
        if (warpRays) atomicAdd(&ctl->raysExtend, warpRays);
        atomicMax(&ctl->iterations, (unsigned long long)trips); // rays of the busiest warp's longest lane chain
    }
}

// Host side of the sphere BVH (layout: SphereBvh above). Median split of the centroids along the longest axis.
static SphereBvh g_sphereBvh;

static void buildSphereBvh(RendererContext& c, const std::vector<float4>& sp, int n) {
    std::vector<unsigned int> always, rest;
    for (int i = 0; i < n; i++) (sp[i].w > 100.0f ? always : rest).push_back((unsigned int)i);
    std::vector<float4> leafSpheres;
    std::vector<unsigned int> leafIds;
    for (unsigned int i : always) { leafSpheres.push_back(sp[i]); leafIds.push_back(i); }
    const size_t maxLeaf = std::getenv("CRT_SPHERES_LEAF") ? (size_t)std::max(1, std::atoi(std::getenv("CRT_SPHERES_LEAF"))) : 1; // measured on config 2: 1 -> 34.0 ms, 2 -> 35.0, 4 -> 36.9
    const bool ordered = !(std::getenv("CRT_SPHERES_ORDERED") && std::getenv("CRT_SPHERES_ORDERED")[0] == '0');
    // the tree, then one depth-first threading of it per ray octant
    struct Node { float bmin[3], bmax[3]; int left, right, axis; unsigned int leafWord; };
    std::vector<Node> tree;
    struct Builder {
        const std::vector<float4>& sp;
        std::vector<Node>& tree;
        std::vector<float4>& leafSpheres;
        std::vector<unsigned int>& leafIds;
        size_t maxLeaf;
        bool sah;
        int build(std::vector<unsigned int>& idx, size_t lo, size_t hi) {
            Node nd;
            float cmin[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmax[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (int a = 0; a < 3; a++) { nd.bmin[a] = FLT_MAX; nd.bmax[a] = -FLT_MAX; }
            for (size_t k = lo; k < hi; k++) {
                const float4 s = sp[idx[k]];
                const float ce[3] = {s.x, s.y, s.z};
                for (int a = 0; a < 3; a++) {
                    const float pad = s.w * 1.01f + 1e-3f + 1e-5f * std::fabs(ce[a]);
                    nd.bmin[a] = std::min(nd.bmin[a], ce[a] - pad);
                    nd.bmax[a] = std::max(nd.bmax[a], ce[a] + pad);
                    cmin[a] = std::min(cmin[a], ce[a]);
                    cmax[a] = std::max(cmax[a], ce[a]);
                }
            }
            nd.left = nd.right = -1;
            nd.axis = 0;
            nd.leafWord = 0xFFFFFFFFu;
            const int me = (int)tree.size();
            tree.push_back(nd);
            if (hi - lo <= maxLeaf) {
                tree[me].leafWord = (unsigned int)leafSpheres.size() | ((unsigned int)(hi - lo) << 24);
                for (size_t k = lo; k < hi; k++) { leafSpheres.push_back(sp[idx[k]]); leafIds.push_back(idx[k]); }
            } else {
                // split: the plane of least surface-area cost over all three axes and all positions (a few hundred spheres: the full
                // sweep costs nothing); CRT_SPHERES_SAH=0: the median of the longest axis (round 1's tree)
                int axis = 0;
                for (int a = 1; a < 3; a++) if (cmax[a] - cmin[a] > cmax[axis] - cmin[axis]) axis = a;
                size_t mid = (lo + hi) / 2;
                auto key = [&](unsigned int p, int a) { return a == 0 ? sp[p].x : a == 1 ? sp[p].y : sp[p].z; };
                if (sah) {
                    const size_t n = hi - lo;
                    double best = 1e300;
                    std::vector<unsigned int> order(idx.begin() + lo, idx.begin() + hi), bestOrder;
                    std::vector<double> rightArea(n + 1, 0.0);
                    for (int a = 0; a < 3; a++) {
                        std::sort(order.begin(), order.end(), [&](unsigned int p, unsigned int q) { return key(p, a) < key(q, a) || (key(p, a) == key(q, a) && p < q); });
                        auto grow = [&](float* mn, float* mx, unsigned int p) {
                            const float4 s4 = sp[p];
                            const float ce[3] = {s4.x, s4.y, s4.z};
                            for (int d = 0; d < 3; d++) { mn[d] = std::min(mn[d], ce[d] - s4.w); mx[d] = std::max(mx[d], ce[d] + s4.w); }
                        };
                        auto area = [](const float* mn, const float* mx) {
                            const double ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
                            return ex * ey + ey * ez + ez * ex;
                        };
                        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
                        for (size_t k = n; k-- > 1;) { grow(mn, mx, order[k]); rightArea[k] = area(mn, mx); }
                        float ln[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, lx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
                        for (size_t k = 1; k < n; k++) {
                            grow(ln, lx, order[k - 1]);
                            const double cost = area(ln, lx) * (double)k + rightArea[k] * (double)(n - k);
                            if (cost < best) { best = cost; mid = lo + k; axis = a; bestOrder = order; }
                        }
                    }
                    std::copy(bestOrder.begin(), bestOrder.end(), idx.begin() + lo);
                } else {
                    std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](unsigned int p, unsigned int q) {
                        return key(p, axis) < key(q, axis) || (key(p, axis) == key(q, axis) && p < q);
                    });
                }
                const int l = build(idx, lo, mid), r = build(idx, mid, hi);
                tree[me].left = l; tree[me].right = r; tree[me].axis = axis;
            }
            return me;
        }
    } builder{sp, tree, leafSpheres, leafIds, maxLeaf, !(std::getenv("CRT_SPHERES_SAH") && std::getenv("CRT_SPHERES_SAH")[0] == '0')};
    if (!rest.empty()) builder.build(rest, 0, rest.size());
    const unsigned int numNodes = (unsigned int)tree.size();
    const unsigned int numOrders = ordered ? 8u : 1u;
    std::vector<float4> nodes;
    struct Emitter {
        const std::vector<Node>& tree;
        std::vector<float4>& nodes;
        unsigned int octant;
        size_t base;
        void emit(int t) { // the lower side first, unless the ray travels towards lower coordinates on the split axis
            const Node& nd = tree[t];
            const size_t me = nodes.size();
            nodes.push_back(make_float4(nd.bmin[0], nd.bmin[1], nd.bmin[2], 0.0f));
            nodes.push_back(make_float4(nd.bmax[0], nd.bmax[1], nd.bmax[2], 0.0f));
            if (nd.left >= 0) {
                const bool flip = ((octant >> nd.axis) & 1u) != 0u;
                emit(flip ? nd.right : nd.left);
                emit(flip ? nd.left : nd.right);
            }
            const unsigned int skip = (unsigned int)((nodes.size() - base) / 2);
            std::memcpy(&nodes[me].w, &skip, 4);
            std::memcpy(&nodes[me + 1].w, &nd.leafWord, 4);
        }
    };
    for (unsigned int oct = 0; oct < numOrders; oct++) {
        Emitter e{tree, nodes, oct, nodes.size()};
        if (numNodes) e.emit(0);
    }
    if (leafSpheres.empty()) { leafSpheres.push_back(make_float4(0, 0, 0, 0)); leafIds.push_back(0); }
    if (nodes.empty()) { nodes.push_back(make_float4(0, 0, 0, 0)); nodes.push_back(make_float4(0, 0, 0, 0)); }
    float4* dNodes = (float4*)arenaAlloc(nodes.size() * sizeof(float4));
    float4* dSpheres = (float4*)arenaAlloc(leafSpheres.size() * sizeof(float4));
    unsigned int* dIds = (unsigned int*)arenaAlloc(leafIds.size() * sizeof(unsigned int));
    CRT_CHECK(cudaMemcpy(dNodes, nodes.data(), nodes.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CRT_CHECK(cudaMemcpy(dSpheres, leafSpheres.data(), leafSpheres.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CRT_CHECK(cudaMemcpy(dIds, leafIds.data(), leafIds.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
    g_sphereBvh.nodes = dNodes;
    g_sphereBvh.spheres = dSpheres;
    g_sphereBvh.ids = dIds;
    g_sphereBvh.numNodes = numNodes;
    g_sphereBvh.numAlways = (unsigned int)always.size();
    g_sphereBvh.octantMask = ordered ? 7u : 0u;
    (void)c;
}

extern "C" void initRendererSpheres(const sphere* spheres, const material* materials, int n, const camera cam, vec3** fb, int nx,
                                    int ny, int maxDepth) {
    RendererContext& c = g_ctx;
    if (n < 0 || n > MAX_SPHERES) {
        std::fprintf(stderr, "initRendererSpheres: %d spheres exceed the constant-memory table (%d)\n", n, MAX_SPHERES);
        std::exit(99);
    }
    if (c.initialised) cleanupRenderer();
    initCommon(c, cam, fb, nx, ny, maxDepth);
    c.kind = SCENE_SPHERES;
    c.numSpheres = n;
    std::vector<float4> sp((size_t)(n > 0 ? n : 1)), mats(2 * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        sp[i] = make_float4(spheres[i].center.e[0], spheres[i].center.e[1], spheres[i].center.e[2], spheres[i].radius);
        const material& mt = materials[i];
        mats[2 * i] = make_float4(mt.color.e[0], mt.color.e[1], mt.color.e[2], mt.param);
        int type = (int)mt.type, tex = -1;
        float4 b;
        std::memcpy(&b.x, &type, 4);
        std::memcpy(&b.y, &tex, 4);
        b.z = b.w = 0.0f;
        mats[2 * i + 1] = b;
    }
    CRT_CHECK(cudaMemcpyToSymbol(c_spheres, sp.data(), (size_t)n * sizeof(float4)));
    buildSphereBvh(c, sp, n);
    c.materials = (float4*)arenaAlloc(mats.size() * sizeof(float4));
    CRT_CHECK(cudaMemcpy(c.materials, mats.data(), mats.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CRT_CHECK(cudaDeviceSynchronize()); // the uploads ran on the legacy stream, the frame runs on non-blocking streams
}

static bool g_spheresBrute = false; // CRT_SPHERES_BRUTE=1: all spheres from __constant__ for every ray (tests compare the two)

static void launchSphereIteration(RendererContext& c, cudaStream_t stream, unsigned int* qCur, unsigned int* qNext, int samplesPerSlot,
                                  int slotsPerPixel) {
    const int grid = c.numSMs * 8;
    if (g_spheresBrute) extendSpheresKernel<<<grid, WF_BLOCK, 0, stream>>>(c.wf, qCur, c.numSpheres); // the README-era loop, kept as the BVH's checker
    else extendSpheresBvhKernel<<<grid, WF_BLOCK, 0, stream>>>(c.wf, qCur, g_sphereBvh);
    shadeSpheresKernel<<<grid, WF_BLOCK, 0, stream>>>(c.wf, c.materials, c.maxDepth, qCur, qNext, c.cam, c.nx, c.ny, samplesPerSlot, slotsPerPixel);
}

void crtRunSpheres(RendererContext& c, int ns) {
    const unsigned int npix = (unsigned int)c.nx * (unsigned int)c.ny;
    int slotsPerPixel = c.opts.reserved[0] > 0 ? c.opts.reserved[0] : 1;
    if (ns % slotsPerPixel != 0) slotsPerPixel = 1;
    const int samplesPerSlot = ns / slotsPerPixel;
    {
        const char* v = std::getenv("CRT_SPHERES_BRUTE");
        g_spheresBrute = v && v[0] == '1';
    }
    allocWavefront(c, npix * (unsigned int)slotsPerPixel);
    cudaStream_t stream = c.stream;
    std::memset(&c.stats, 0, sizeof(c.stats));
    c.stats.samples = (unsigned long long)npix * (unsigned long long)(ns > 0 ? ns : 0);
    CRT_CHECK(cudaEventRecord(c.evStart, stream));
    CRT_CHECK(cudaMemsetAsync(c.wf.accum, 0, (size_t)npix * sizeof(float4), stream));
    CRT_CHECK(cudaMemsetAsync(c.wf.ctl, 0, sizeof(WfControl), stream));
    unsigned long long launches = 0;
    const bool wavefront = g_spheresBrute || (std::getenv("CRT_SPHERES_WAVEFRONT") && std::getenv("CRT_SPHERES_WAVEFRONT")[0] == '1');
    if (npix > 0 && ns > 0 && c.maxDepth > 0 && !wavefront) {
        // the product path: one persistent launch (the wavefront kernels below stay as its checker: CRT_SPHERES_WAVEFRONT=1)
        const int grid = c.numSMs * SPH_MEGA_BLOCKS_PER_SM;
        if (c.counting) // setRendererCounting: box and sphere tests per ray for the bench's flop model (not a timed configuration)
            spheresMegaKernel<true><<<grid, SPH_MEGA_BLOCK, 0, stream>>>(c.wf, c.materials, c.maxDepth, g_sphereBvh, c.cam, c.nx, c.ny, samplesPerSlot, slotsPerPixel,
                                                                         c.opts.sampleStream, npix * (unsigned int)slotsPerPixel);
        else
            spheresMegaKernel<false><<<grid, SPH_MEGA_BLOCK, 0, stream>>>(c.wf, c.materials, c.maxDepth, g_sphereBvh, c.cam, c.nx, c.ny, samplesPerSlot, slotsPerPixel,
                                                                          c.opts.sampleStream, npix * (unsigned int)slotsPerPixel);
        launches += 1;
        CRT_CHECK(cudaMemcpyAsync(c.hostCtl, c.wf.ctl, sizeof(WfControl), cudaMemcpyDeviceToHost, stream));
    } else if (npix > 0 && ns > 0 && c.maxDepth > 0) {
        const int grid = c.numSMs * 8;
        raygenKernel<true><<<grid, WF_BLOCK, 0, stream>>>(c.wf, c.cam, c.wf.queueA, c.nx, c.ny, samplesPerSlot, slotsPerPixel,
                                                          c.opts.sampleStream);
        advanceKernel<<<1, 1, 0, stream>>>(c.wf.ctl);
        launches += 2;
        int batch = c.opts.megaBatch > 0 ? c.opts.megaBatch : 16;
        batch = (batch + 1) & ~1;
        const unsigned int finishBelow = std::getenv("CRT_SPHERES_FINISH") ? (unsigned int)std::atoi(std::getenv("CRT_SPHERES_FINISH")) : 131072u; // 0: wavefront to the end (measured on config 2: 0 -> 70.6 ms, 16 Ki -> 57.8, 48 Ki -> 56.1, 128 Ki -> 53.6, 400 k -> 55.8)
        const long long key = ((long long)samplesPerSlot << 24) ^ ((long long)slotsPerPixel << 8) ^ batch ^ (1LL << 61) ^
                              ((long long)c.maxDepth << 40) ^ ((long long)g_spheresBrute << 58);
        if (!c.graphExec || c.graphKey != key) {
            if (c.graphExec) { cudaGraphExecDestroy(c.graphExec); c.graphExec = nullptr; }
            cudaGraph_t graph;
            CRT_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            for (int k = 0; k < batch; k++) {
                const bool flip = (k & 1) != 0;
                launchSphereIteration(c, stream, flip ? c.wf.queueB : c.wf.queueA, flip ? c.wf.queueA : c.wf.queueB, samplesPerSlot,
                                      slotsPerPixel);
            }
            CRT_CHECK(cudaStreamEndCapture(stream, &graph));
            CRT_CHECK(cudaGraphInstantiate(&c.graphExec, graph, 0));
            CRT_CHECK(cudaGraphDestroy(graph));
            c.graphKey = key;
        }
        while (true) {
            CRT_CHECK(cudaGraphLaunch(c.graphExec, stream));
            launches += (unsigned long long)batch * 2;
            CRT_CHECK(cudaMemcpyAsync(c.hostCtl, c.wf.ctl, sizeof(WfControl), cudaMemcpyDeviceToHost, stream));
            CRT_CHECK(cudaStreamSynchronize(stream));
            if (c.hostCtl->countActive == 0) break;
            if (!g_spheresBrute && c.hostCtl->countActive <= finishBelow) { // the tail: one thread per remaining slot (batch is even: the live queue is queueA)
                finishSpheresKernel<<<(c.hostCtl->countActive + 127) / 128, 128, 0, stream>>>(c.wf, c.materials, c.maxDepth, c.wf.queueA, g_sphereBvh, c.cam, c.nx,
                                                                                            c.ny, samplesPerSlot, slotsPerPixel);
                finishSpheresDoneKernel<<<1, 1, 0, stream>>>(c.wf.ctl);
                launches += 2;
                CRT_CHECK(cudaMemcpyAsync(c.hostCtl, c.wf.ctl, sizeof(WfControl), cudaMemcpyDeviceToHost, stream));
                CRT_CHECK(cudaStreamSynchronize(stream));
                break;
            }
        }
    } else {
        CRT_CHECK(cudaMemcpyAsync(c.hostCtl, c.wf.ctl, sizeof(WfControl), cudaMemcpyDeviceToHost, stream));
    }
    if (!c.opts.deferFinalize) {
        finalizeFrameTo(c, c.wf.accum, float(ns), stream);
        launches += 1;
    }
    CRT_CHECK(cudaEventRecord(c.evStop, stream));
    CRT_CHECK(cudaStreamSynchronize(stream));
    CRT_CHECK(cudaGetLastError());
    CRT_CHECK(cudaEventElapsedTime(&c.stats.msTotal, c.evStart, c.evStop));
    c.stats.raysExtend = c.hostCtl->raysExtend;
    c.lastNodeVisits = c.hostCtl->nodeVisits; // box tests / sphere tests of a counted frame (getRendererTraversalCounts)
    c.lastTriTests = c.hostCtl->triTests;
    c.stats.raysShadow = 0;
    c.stats.iterations = c.hostCtl->iterations;
    c.stats.kernelLaunches = launches;
}
