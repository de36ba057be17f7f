// bsdf.cuh -- material scatter (the "material" kernel's arithmetic).
//   schlick / refract / reflect   material.h:9-25
//   diffuse_bsdf                  material.h:27-31
//   glossy_bsdf                   material.h:46-53
//   fresnel_layer                 material.h:55-60  (draws one number only when not totally reflected)
//   dielectric_bsdf               material.h:73-92
//   coat_bsdf                     material.h:62-70  (library; unused by the staircase table)
//   material_scatter dispatch     scene_materials.h:13-20
// RNG draw order inside a bounce (SURVEY.md section 7 "hard parts") is kept:
// DIFFUSE 3 per rejection round; METAL the same only when fuzz > 1e-4; GLASS
// at most one.
#pragma once

#include "rng.cuh"
#include "vecmath.cuh"

struct Scatter {  // scatter_info, helper_structs.h:38-46
    f3 wi;
    f3 throughput;
    float t;
    bool specular;
    bool refracted;
};

struct SurfacePoint {  // the fields of `intersection` (helper_structs.h:16-36) the BSDFs read
    f3 normal;         // faces the ray
    float t;
    bool inside;
};

__device__ __forceinline__ float schlick(float cosine, float refIdx) {
    float r0 = (1.0f - refIdx) / (1.0f + refIdx);
    r0 = r0 * r0;
    return r0 + (1.0f - r0) * powf((1.0f - cosine), 5.0f);
}

__device__ __forceinline__ f3 refractDir(const f3& uv, const f3& n, float etaiOverEtat) {
    float cosTheta = fminf(dot(-uv, n), 1.0f);
    f3 rOutParallel = etaiOverEtat * (uv + cosTheta * n);
    float sq = sqlen(rOutParallel);
    f3 rOutPerp = sq >= 1.0f ? mk3(0.0f, 0.0f, 0.0f) : -sqrtf(1.0f - sq) * n;
    return rOutParallel + rOutPerp;
}

__device__ __forceinline__ f3 reflectDir(const f3& v, const f3& n) { return v - 2.0f * dot(v, n) * n; }

__device__ __forceinline__ void diffuseBsdf(Scatter& out, const SurfacePoint& i, const f3& albedo, unsigned int& rng) {
    out.wi = unit(i.normal + randomInUnitSphere(rng));
    out.throughput = albedo;
    out.specular = false;
}

__device__ __forceinline__ void glossyBsdf(Scatter& out, const SurfacePoint& i, const f3& wo, const f3& tint, float fuzz,
                                           unsigned int& rng) {
    f3 reflected = reflectDir(wo, i.normal);
    if (fuzz > 0.0001f) reflected = reflected + fuzz * randomInUnitSphere(rng);
    out.wi = unit(reflected);
    out.throughput = out.throughput * tint;
    out.specular = true;
}

__device__ __forceinline__ bool fresnelLayer(const SurfacePoint& i, const f3& wo, float ior, unsigned int& rng) {
    float etaiOverEtat = i.inside ? ior : (1.0f / ior);
    float cosTheta = fminf(dot(-wo, i.normal), 1.0f);
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    return (etaiOverEtat * sinTheta > 1.0f || rnd(rng) < schlick(cosTheta, etaiOverEtat));
}

__device__ __forceinline__ void dielectricBsdf(Scatter& out, const SurfacePoint& i, const f3& wo, float layerIor, const f3& glossyTint,
                                               float glossyFuzz, const f3& absorption, unsigned int& rng) {
    if (i.inside) {
        f3 e = -absorption * i.t;
        out.throughput = mk3(expf(e.x), expf(e.y), expf(e.z));
    }
    if (fresnelLayer(i, wo, layerIor, rng)) {
        glossyBsdf(out, i, wo, glossyTint, glossyFuzz, rng);
    } else {
        float etaiOverEtat = i.inside ? layerIor : (1.0f / layerIor);
        out.wi = unit(refractDir(wo, i.normal, etaiOverEtat));
        out.refracted = true;
    }
    out.specular = true;
}

__device__ __forceinline__ void coatBsdf(Scatter& out, const SurfacePoint& i, const f3& wo, float layerIor, const f3& glossyTint,
                                         float glossyFuzz, const f3& diffuseAlbedo, unsigned int& rng) {
    if (fresnelLayer(i, wo, layerIor, rng))
        glossyBsdf(out, i, wo, glossyTint, glossyFuzz, rng);
    else
        diffuseBsdf(out, i, diffuseAlbedo, rng);
}

// material types: helper_structs.h:127-131
#define MAT_DIFFUSE 0
#define MAT_METAL 1
#define MAT_GLASS 2

__device__ __forceinline__ void materialScatter(Scatter& out, const SurfacePoint& i, const f3& wo, int type, float param,
                                                const f3& color, unsigned int& rng) {
    if (type == MAT_DIFFUSE)
        diffuseBsdf(out, i, color, rng);
    else if (type == MAT_METAL)
        glossyBsdf(out, i, wo, color, param, rng);
    else
        dielectricBsdf(out, i, wo, param, color, 0.0f, mk3(0.0f, 0.0f, 0.0f), rng);
}
