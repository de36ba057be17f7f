// ref_shim.cu -- TEST INFRASTRUCTURE. Builds oracle/_ref/libref_shim.so: the reference's kernels.cu compiled
// from where it lies (found through -I/root/reference, never copied) plus two probes that call the
// reference's own device functions on caller-supplied inputs:
//   refIntersectBatch   hit()'s ray construction (ray.h:9) + hitMesh() (kernels.cu:296) on a ray batch
//   refRenderCount      nothing of ours: just re-exports the 3 entry points (they come with the include)
//   refScatterBatch     the scene presets of scene_materials.h:22-93 (and subsurface_bsdf, material.h:94) on a batch of
//                       surface points: the reference's own BSDF library, for csrc/bsdf.cuh's restatement of it
// The probes use the global `renderContext` that the reference's initRenderer fills (kernels.cu:145,571).
#include "kernels.cu"

__global__ void shimIntersectKernel(const RenderContext context, const float4* rayO, const float4* rayD, long long n, int isShadow,
                                    float4* outHit, int* outMesh) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 o = rayO[i], d = rayD[i];
    const ray r(vec3(o.x, o.y, o.z), vec3(d.x, d.y, d.z));
    tri_hit th;
    th.triId = 0xFFFFFFFFu;
    th.u = 0;
    th.v = 0;
    float t = hitMesh(r, context, o.w, d.w, th, true, isShadow != 0);
    int meshID = -1;
    if (t < d.w) {
        if (isShadow) {
            t = 0.0f;
            th.triId = 0xFFFFFFFFu;
            th.u = th.v = 0;
        } else {
            meshID = context.tris[th.triId].meshID;
        }
    } else {
        t = FLT_MAX;
        th.triId = 0xFFFFFFFFu;
        th.u = th.v = 0;
    }
    outHit[i] = make_float4(t, th.u, th.v, __uint_as_float(th.triId));
    outMesh[i] = meshID;
}

// Host pointers in and out; rays as float4 {o.xyz,tMin} {d.xyz,tMax}. Returns kernel milliseconds.
extern "C" float refIntersectBatch(const float* rayO, const float* rayD, long long n, int isShadow, float* outHit, int* outMesh) {
    float4 *dO, *dD, *dH;
    int* dM;
    checkCudaErrors(cudaMalloc(&dO, n * sizeof(float4)));
    checkCudaErrors(cudaMalloc(&dD, n * sizeof(float4)));
    checkCudaErrors(cudaMalloc(&dH, n * sizeof(float4)));
    checkCudaErrors(cudaMalloc(&dM, n * sizeof(int)));
    checkCudaErrors(cudaMemcpy(dO, rayO, n * sizeof(float4), cudaMemcpyHostToDevice));
    checkCudaErrors(cudaMemcpy(dD, rayD, n * sizeof(float4), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    shimIntersectKernel<<<(unsigned)((n + 63) / 64), 64>>>(renderContext, dO, dD, n, isShadow, dH, dM);
    cudaEventRecord(e1);
    checkCudaErrors(cudaGetLastError());
    checkCudaErrors(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    checkCudaErrors(cudaMemcpy(outHit, dH, n * sizeof(float4), cudaMemcpyDeviceToHost));
    checkCudaErrors(cudaMemcpy(outMesh, dM, n * sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(dO); cudaFree(dD); cudaFree(dH); cudaFree(dM);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

// ---- BSDF probe (same item layout as scatterBatch in include/kernels.h) ------------------------------------------------
__global__ void shimScatterKernel(int preset, long long n, const float4* in, float4* out) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 a = in[3 * k], b = in[3 * k + 1], c = in[3 * k + 2];
    intersection i;
    i.objId = 0;
    i.meshID = 0;
    i.triID = 0;
    i.normal = vec3(a.x, a.y, a.z);
    i.t = a.w;
    i.p = vec3(b.x, b.y, b.z);
    i.inside = b.w != 0.0f;
    i.texCoords[0] = i.texCoords[1] = 0.0f;
    const vec3 wo(c.x, c.y, c.z);
    rand_state rng = __float_as_uint(c.w);
    scatter_info s(i);
    s.wi = vec3(0, 0, 0);
    switch (preset) {
        case 0: floor_coat_scatter(s, i, wo, rng); break;
        case 1: floor_diffuse_scatter(s, i, wo, rng); break;
        case 2: floor_checker_scatter(s, i, wo, rng); break;
        case 3: model_coat_scatter(s, i, wo, rng); break;
        case 4: model_diffuse_scatter(s, i, wo, rng); break;
        case 5: model_glossy_scatter(s, i, wo, rng); break;
        case 6: model_glass_scatter(s, i, wo, rng); break;
        case 7: model_tintedglass_scatter(s, i, wo, rng); break;
        case 8: model_sss_scatter(s, i, wo, rng); break;
        default: subsurface_bsdf(s, i, wo, vec3(0.9f, 0.3f, 0.02f), 2.0f, rng); break;
    }
    out[3 * k] = make_float4(s.wi.x(), s.wi.y(), s.wi.z(), s.t);
    out[3 * k + 1] = make_float4(s.throughput.x(), s.throughput.y(), s.throughput.z(), __int_as_float((s.specular ? 1 : 0) | (s.refracted ? 2 : 0)));
    out[3 * k + 2] = make_float4(__uint_as_float((unsigned int)rng), 0.0f, 0.0f, 0.0f);
}

extern "C" int refScatterBatch(int preset, long long n, const float* in, float* out) {
    float4 *dIn, *dOut;
    checkCudaErrors(cudaMalloc(&dIn, n * 3 * sizeof(float4)));
    checkCudaErrors(cudaMalloc(&dOut, n * 3 * sizeof(float4)));
    checkCudaErrors(cudaMemcpy(dIn, in, n * 3 * sizeof(float4), cudaMemcpyHostToDevice));
    shimScatterKernel<<<(unsigned)((n + 63) / 64), 64>>>(preset, n, dIn, dOut);
    checkCudaErrors(cudaGetLastError());
    checkCudaErrors(cudaDeviceSynchronize());
    checkCudaErrors(cudaMemcpy(out, dOut, n * 3 * sizeof(float4), cudaMemcpyDeviceToHost));
    cudaFree(dIn);
    cudaFree(dOut);
    return 0;
}
