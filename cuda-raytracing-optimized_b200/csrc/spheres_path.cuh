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
// same bits. The spheres are put into a small BVH at init (host, median split, <= 4 per leaf, boxes padded by 1 % of the
// radius + 1e-3: three orders of magnitude more than the rounding of r_s for these sizes); spheres too large for that
// margin (radius > 100: the ground sphere, whose roots lose ~1e-3 to cancellation) stay in a list that every ray tests.
// The tree is stored depth-first with skip links (no stack): node i's subtree is [i+1, skip_i).
struct SphereBvh {
    const float4* __restrict__ nodes;   // 2 per node: {bmin.xyz, skip}{bmax.xyz, first | count << 24 (0xFFFFFFFF = internal)}
    const float4* __restrict__ spheres; // leaf order: {center.xyz, radius}
    const unsigned int* __restrict__ ids; // leaf order -> index the caller gave the sphere
    unsigned int numNodes;
    unsigned int numAlways;             // spheres[0 .. numAlways) are tested by every ray
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
__device__ __forceinline__ void closestSphere(const SphereBvh& bvh, const f3& o, const f3& d, float& closest, unsigned int& id) {
    const f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    closest = FLT_MAX;
    id = 0xFFFFFFFFu;
    for (unsigned int k = 0; k < bvh.numAlways; k++) testSphere(bvh, k, o, d, closest, id);
    unsigned int node = 0;
    while (node < bvh.numNodes) {
        const float4 lo = __ldg(bvh.nodes + 2 * node);
        const float4 hi = __ldg(bvh.nodes + 2 * node + 1);
        // conservative slab test: fminf/fmaxf drop the NaN of 0 * inf (an axis the ray is parallel to)
        const float x0 = (lo.x - o.x) * inv.x, x1 = (hi.x - o.x) * inv.x;
        const float y0 = (lo.y - o.y) * inv.y, y1 = (hi.y - o.y) * inv.y;
        const float z0 = (lo.z - o.z) * inv.z, z1 = (hi.z - o.z) * inv.z;
        const float tEnter = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
        const float tExit = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
        const bool hitBox = tEnter <= tExit * 1.00001f + 1e-5f && tEnter <= closest;
        const unsigned int leaf = __float_as_uint(hi.w);
        if (!hitBox) {
            node = __float_as_uint(lo.w); // skip the subtree
        } else {
            if (leaf != 0xFFFFFFFFu) {
                const unsigned int first = leaf & 0xFFFFFFu, count = leaf >> 24;
                for (unsigned int k = 0; k < count; k++) testSphere(bvh, first + k, o, d, closest, id);
            }
            node = node + 1;
        }
    }
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

// Shade + retire + regenerate + advance in one launch (an iteration is extend, then this): when a path ends its colour goes
// into the pixel (col += p.color, kernels.cu:558, in sample order: one slot per pixel) and the slot's next sample starts
// right here (kernels.cu:549-555) instead of in a separate raygen pass; the last block to finish swaps the queues.
// Everything between two closest-sphere queries of one path slot whose hit record is `h` (color()'s loop body, kernels.cu:402-531,
// without light and shadow rays): sky gradient on a miss, scatter, Russian roulette; when the path ends, col += p.color and the
// slot's next camera ray. Returns whether the slot has another ray to trace. Shared by shadeSpheresKernel and finishSpheresKernel.
__device__ __forceinline__ bool shadeSphereSlot(const WfState& st, const float4* __restrict__ mats, int maxDepth, const CameraDev& cam, int nx, int ny,
                                                int samplesPerSlot, int slotsPerPixel, unsigned int npix, unsigned int slot, const float4& h) {
    bool continues = false;
            const float4 ro = st.rayO[slot];
            const float4 rd = st.rayD[slot];
            f3 origin = xyz(ro), dir = xyz(rd);
            unsigned int rng = __float_as_uint(ro.w);
            unsigned int flags = __float_as_uint(rd.w);
            bool inside = (flags & PATH_FLAG_INSIDE) != 0u;
            unsigned int bounce = flags & PATH_BOUNCE_MASK;
            const float4 att4 = st.atten[slot];
            f3 att = xyz(att4);
            float4 pc = st.pcol[slot];
            if (!(h.x < FLT_MAX)) {
                // sky gradient, kernels.cu:419-421
                const float t = 0.5f * (dir.y + 1.0f);
                const f3 c = (1.0f - t) * mk3(1.0f, 1.0f, 1.0f) + t * mk3(0.5f, 0.7f, 1.0f);
                const f3 add = att * c;
                pc.x += add.x; pc.y += add.y; pc.z += add.z;
            } else {
                const unsigned int id = __float_as_uint(h.w);
                const float4 sp = c_spheres[id];
                const f3 rdir = unit(dir);
                const f3 p = origin + h.x * rdir; // point_at_parameter on the traced (normalised) ray
                SurfacePoint s;
                s.normal = (p - xyz(sp)) / sp.w;
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
                materialScatter(scat, s, dir, __float_as_int(m1.x), m0.w, xyz(m0), rng);
                origin = origin + scat.t * dir;
                dir = scat.wi;
                att = att * scat.throughput;
                inside = scat.refracted ? !inside : inside;
                continues = true;
                if (bounce > 3u) {
                    const float m = maxcomp(att);
                    if (rnd(rng) > m) continues = false;
                    else att = att * (1 / m);
                }
                if (continues) {
                    bounce = (bounce + 1u) & PATH_BOUNCE_MASK;
                    if (!((int)bounce < maxDepth)) continues = false;
                }
                flags = bounce | (scat.specular ? PATH_FLAG_SPECULAR : 0u) | (inside ? PATH_FLAG_INSIDE : 0u);
            }
            if (continues) {
                st.rayO[slot] = mk4(origin, __uint_as_float(rng));
                st.rayD[slot] = mk4(dir, __uint_as_float(flags));
                st.atten[slot] = mk4(att, att4.w);
            } else {
                // the sample is finished: col += p.color (kernels.cu:558), then the slot's next sample
                const unsigned int pixel = slot % npix;
                if (slotsPerPixel == 1) {
                    float4 a = st.accum[pixel];
                    a.x += pc.x; a.y += pc.y; a.z += pc.z;
                    st.accum[pixel] = a;
                } else {
                    atomicAdd(&st.accum[pixel].x, pc.x);
                    atomicAdd(&st.accum[pixel].y, pc.y);
                    atomicAdd(&st.accum[pixel].z, pc.z);
                }
                const int sample = __float_as_int(att4.w) + 1;
                if (sample < samplesPerSlot) {
                    const int px = (int)(pixel % (unsigned int)nx), py = (int)(pixel / (unsigned int)nx);
                    const float u = float(px + rnd(rng)) / float(nx);
                    const float v = float(py + rnd(rng)) / float(ny);
                    f3 o, d;
                    cameraRay(cam, u, v, rng, o, d);
                    st.rayO[slot] = mk4(o, __uint_as_float(rng));
                    st.rayD[slot] = mk4(d, __uint_as_float(0u)); // bounce 0, specular = inside = false (kernels.cu:554-555)
                    st.atten[slot] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(sample));
                    st.pcol[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    continues = true;
                }
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

// Host side of the sphere BVH (layout: SphereBvh above). Median split of the centroids along the longest axis.
static SphereBvh g_sphereBvh;

static void buildSphereBvh(RendererContext& c, const std::vector<float4>& sp, int n) {
    std::vector<unsigned int> always, rest;
    for (int i = 0; i < n; i++) (sp[i].w > 100.0f ? always : rest).push_back((unsigned int)i);
    std::vector<float4> nodes, leafSpheres;
    std::vector<unsigned int> leafIds;
    for (unsigned int i : always) { leafSpheres.push_back(sp[i]); leafIds.push_back(i); }
    struct Builder {
        const std::vector<float4>& sp;
        std::vector<float4>& nodes;
        std::vector<float4>& leafSpheres;
        std::vector<unsigned int>& leafIds;
        void build(std::vector<unsigned int>& idx, size_t lo, size_t hi) {
            float bmin[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, bmax[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            float cmin[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmax[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (size_t k = lo; k < hi; k++) {
                const float4 s = sp[idx[k]];
                const float ce[3] = {s.x, s.y, s.z};
                for (int a = 0; a < 3; a++) {
                    const float pad = s.w * 1.01f + 1e-3f + 1e-5f * std::fabs(ce[a]);
                    bmin[a] = std::min(bmin[a], ce[a] - pad);
                    bmax[a] = std::max(bmax[a], ce[a] + pad);
                    cmin[a] = std::min(cmin[a], ce[a]);
                    cmax[a] = std::max(cmax[a], ce[a]);
                }
            }
            const size_t me = nodes.size() / 2;
            nodes.push_back(make_float4(bmin[0], bmin[1], bmin[2], 0.0f));
            nodes.push_back(make_float4(bmax[0], bmax[1], bmax[2], 0.0f));
            unsigned int leafWord = 0xFFFFFFFFu;
            if (hi - lo <= 4) {
                leafWord = (unsigned int)leafSpheres.size() | ((unsigned int)(hi - lo) << 24);
                for (size_t k = lo; k < hi; k++) { leafSpheres.push_back(sp[idx[k]]); leafIds.push_back(idx[k]); }
            } else {
                int axis = 0;
                for (int a = 1; a < 3; a++) if (cmax[a] - cmin[a] > cmax[axis] - cmin[axis]) axis = a;
                const size_t mid = (lo + hi) / 2;
                std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](unsigned int p, unsigned int q) {
                    const float a = axis == 0 ? sp[p].x : axis == 1 ? sp[p].y : sp[p].z;
                    const float b = axis == 0 ? sp[q].x : axis == 1 ? sp[q].y : sp[q].z;
                    return a < b || (a == b && p < q);
                });
                build(idx, lo, mid);
                build(idx, mid, hi);
            }
            const unsigned int skip = (unsigned int)(nodes.size() / 2);
            std::memcpy(&nodes[2 * me].w, &skip, 4);
            std::memcpy(&nodes[2 * me + 1].w, &leafWord, 4);
        }
    } builder{sp, nodes, leafSpheres, leafIds};
    if (!rest.empty()) builder.build(rest, 0, rest.size());
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
    g_sphereBvh.numNodes = rest.empty() ? 0u : (unsigned int)(nodes.size() / 2);
    g_sphereBvh.numAlways = (unsigned int)always.size();
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
    if (npix > 0 && ns > 0 && c.maxDepth > 0) {
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
        if (npix) finalizeKernel<<<(npix + 255) / 256, 256, 0, stream>>>(c.wf.accum, (float*)c.fb, npix, float(ns));
        launches += 1;
    }
    CRT_CHECK(cudaEventRecord(c.evStop, stream));
    CRT_CHECK(cudaStreamSynchronize(stream));
    CRT_CHECK(cudaGetLastError());
    CRT_CHECK(cudaEventElapsedTime(&c.stats.msTotal, c.evStart, c.evStop));
    c.stats.raysExtend = c.hostCtl->raysExtend;
    c.stats.raysShadow = 0;
    c.stats.iterations = c.hostCtl->iterations;
    c.stats.kernelLaunches = launches;
}
