// ref_shim.cu -- TEST INFRASTRUCTURE. Builds oracle/_ref/libref_shim.so: the reference's kernels.cu compiled
// from where it lies (found through -I/root/reference, never copied) plus two probes that call the
// reference's own device functions on caller-supplied inputs:
//   refIntersectBatch   hit()'s ray construction (ray.h:9) + hitMesh() (kernels.cu:296) on a ray batch
//   refRenderCount      nothing of ours: just re-exports the 3 entry points (they come with the include)
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
