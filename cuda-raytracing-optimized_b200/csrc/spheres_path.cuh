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

__global__ void __launch_bounds__(WF_BLOCK) shadeSpheresKernel(WfState st, const float4* __restrict__ mats, int maxDepth,
                                                               const unsigned int* __restrict__ queue, unsigned int* __restrict__ nextQueue) {
    WfControl* ctl = st.ctl;
    const unsigned int n = ctl->countActive;
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
        const unsigned int i = base + laneId();
        bool continues = false, ended = false;
        unsigned int slot = 0;
        if (i < n) {
            slot = queue[i];
            const float4 h = st.hit[slot];
            const float4 ro = st.rayO[slot];
            const float4 rd = st.rayD[slot];
            f3 origin = xyz(ro), dir = xyz(rd);
            unsigned int rng = __float_as_uint(ro.w);
            unsigned int flags = __float_as_uint(rd.w);
            bool inside = (flags & PATH_FLAG_INSIDE) != 0u;
            unsigned int bounce = flags & PATH_BOUNCE_MASK;
            const float4 att4 = st.atten[slot];
            f3 att = xyz(att4);
            if (!(h.x < FLT_MAX)) {
                // sky gradient, kernels.cu:419-421
                const float t = 0.5f * (dir.y + 1.0f);
                const f3 c = (1.0f - t) * mk3(1.0f, 1.0f, 1.0f) + t * mk3(0.5f, 0.7f, 1.0f);
                float4 pc = st.pcol[slot];
                const f3 add = att * c;
                pc.x += add.x; pc.y += add.y; pc.z += add.z;
                st.pcol[slot] = pc;
                ended = true;
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
                ended = !continues;
                flags = bounce | (scat.specular ? PATH_FLAG_SPECULAR : 0u) | (inside ? PATH_FLAG_INSIDE : 0u);
                st.rayO[slot] = mk4(origin, __uint_as_float(rng));
                if (continues) {
                    st.rayD[slot] = mk4(dir, __uint_as_float(flags));
                    st.atten[slot] = mk4(att, att4.w);
                }
            }
        }
        const unsigned int posNext = warpAppend(continues, &ctl->countNext);
        if (continues) nextQueue[posNext] = slot;
        const unsigned int posRegen = warpAppend(ended, &ctl->countRegen);
        if (ended) st.regen[posRegen] = slot;
    }
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
    c.materials = (float4*)arenaAlloc(mats.size() * sizeof(float4));
    CRT_CHECK(cudaMemcpy(c.materials, mats.data(), mats.size() * sizeof(float4), cudaMemcpyHostToDevice));
}

static void launchSphereIteration(RendererContext& c, cudaStream_t stream, unsigned int* qCur, unsigned int* qNext, int samplesPerSlot,
                                  int slotsPerPixel) {
    const int grid = c.numSMs * 8;
    extendSpheresKernel<<<grid, WF_BLOCK, 0, stream>>>(c.wf, qCur, c.numSpheres);
    shadeSpheresKernel<<<grid, WF_BLOCK, 0, stream>>>(c.wf, c.materials, c.maxDepth, qCur, qNext);
    raygenKernel<false><<<grid, WF_BLOCK, 0, stream>>>(c.wf, c.cam, qNext, c.nx, c.ny, samplesPerSlot, slotsPerPixel, c.opts.sampleStream);
    advanceKernel<<<1, 1, 0, stream>>>(c.wf.ctl);
}

void crtRunSpheres(RendererContext& c, int ns) {
    const unsigned int npix = (unsigned int)c.nx * (unsigned int)c.ny;
    int slotsPerPixel = c.opts.reserved[0] > 0 ? c.opts.reserved[0] : 1;
    if (ns % slotsPerPixel != 0) slotsPerPixel = 1;
    const int samplesPerSlot = ns / slotsPerPixel;
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
        const long long key = ((long long)samplesPerSlot << 24) ^ ((long long)slotsPerPixel << 8) ^ batch ^ (1LL << 61) ^
                              ((long long)c.maxDepth << 40);
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
            launches += (unsigned long long)batch * 4;
            CRT_CHECK(cudaMemcpyAsync(c.hostCtl, c.wf.ctl, sizeof(WfControl), cudaMemcpyDeviceToHost, stream));
            CRT_CHECK(cudaStreamSynchronize(stream));
            if (c.hostCtl->countActive == 0) break;
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
