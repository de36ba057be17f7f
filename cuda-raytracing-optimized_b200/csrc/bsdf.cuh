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

// reflect (material.h:23-25) with the dot product supplied by the caller: v - (2*d)*n, one fused op per component
// (the reference's build: FADD d+d, then FFMA -n*k + v).
__device__ __forceinline__ f3 reflectWithDot(const f3& v, const f3& n, float d) {
    const float k = __fadd_rn(d, d);
    return mk3(__fmaf_rn(-n.x, k, v.x), __fmaf_rn(-n.y, k, v.y), __fmaf_rn(-n.z, k, v.z));
}

__device__ __forceinline__ void diffuseBsdf(Scatter& out, const SurfacePoint& i, const f3& albedo, unsigned int& rng) {
    out.wi = unit(i.normal + randomInUnitSphere(rng));
    out.throughput = albedo;
    out.specular = false;
}

// glossy_bsdf (material.h:46-53); `dotWoN` = dot(wo, normal) as the caller's context compiles it
__device__ __forceinline__ void glossyBsdf(Scatter& out, const SurfacePoint& i, const f3& wo, float dotWoN, const f3& tint, float fuzz,
                                           unsigned int& rng) {
    f3 reflected = reflectWithDot(wo, i.normal, dotWoN);
    if (fuzz > 0.0001f) {
        const f3 s = randomInUnitSphere(rng);
        reflected = mk3(__fmaf_rn(fuzz, s.x, reflected.x), __fmaf_rn(fuzz, s.y, reflected.y), __fmaf_rn(fuzz, s.z, reflected.z));
    }
    out.wi = unit(reflected);
    out.throughput = out.throughput * tint;
    out.specular = true;
}

// dielectric_bsdf (material.h:73-92) with fresnel_layer (:55-60) and refract (:15-21) folded in.
// In the reference's compiled kernel dot(-wo, n) [fresnel_layer and refract] and dot(wo, n) [the reflect branch] share
// their rounded products wo.x*n.x and wo.z*n.z (common-subexpression elimination), which fixes the shape of both:
//     cos_theta = ((-wo.y*n.y) - P1) - P2,     dot(wo, n) = P2 + (wo.y*n.y + P1),     P1 = rn(wo.x*n.x), P2 = rn(wo.z*n.z)
// with the y product fused. Written out so that no compiler stage can choose differently.
__device__ __forceinline__ void dielectricBsdf(Scatter& out, const SurfacePoint& i, const f3& wo, float layerIor, const f3& glossyTint,
                                               float glossyFuzz, const f3& absorption, unsigned int& rng) {
    if (i.inside) {
        f3 e = -absorption * i.t;
        out.throughput = mk3(expf(e.x), expf(e.y), expf(e.z));
    }
    const f3& n = i.normal;
    const float p1 = __fmul_rn(wo.x, n.x), p2 = __fmul_rn(wo.z, n.z);
    const float etaiOverEtat = i.inside ? layerIor : (1.0f / layerIor);
    const float cosTheta = fminf(__fsub_rn(__fmaf_rn(-wo.y, n.y, -p1), p2), 1.0f);
    const float sinTheta = sqrtf(__fmaf_rn(-cosTheta, cosTheta, 1.0f));
    if (__fmul_rn(etaiOverEtat, sinTheta) > 1.0f || rnd(rng) < schlick(cosTheta, etaiOverEtat)) { // short-circuit: no draw on TIR
        glossyBsdf(out, i, wo, __fadd_rn(p2, __fmaf_rn(wo.y, n.y, p1)), glossyTint, glossyFuzz, rng);
    } else {
        // refract(wo, n, eta): eta*(wo + cos*n), then the perpendicular part added as a separate (unfused) term
        const f3 par = mk3(__fmul_rn(etaiOverEtat, __fmaf_rn(cosTheta, n.x, wo.x)), __fmul_rn(etaiOverEtat, __fmaf_rn(cosTheta, n.y, wo.y)),
                           __fmul_rn(etaiOverEtat, __fmaf_rn(cosTheta, n.z, wo.z)));
        const float sq = sqlen(par);
        f3 perp = mk3(0.0f, 0.0f, 0.0f);
        if (!(sq >= 1.0f)) {
            const float k = -sqrtf(__fsub_rn(1.0f, sq));
            perp = mk3(__fmul_rn(k, n.x), __fmul_rn(k, n.y), __fmul_rn(k, n.z));
        }
        out.wi = unit(mk3(__fadd_rn(par.x, perp.x), __fadd_rn(par.y, perp.y), __fadd_rn(par.z, perp.z)));
        out.refracted = true;
    }
    out.specular = true;
}

// fresnel_layer (material.h:55-60) on its own, for layered BSDFs outside the staircase table
__device__ __forceinline__ bool fresnelLayer(const SurfacePoint& i, const f3& wo, float ior, unsigned int& rng) {
    const float etaiOverEtat = i.inside ? ior : (1.0f / ior);
    const float cosTheta = fminf(-dot(wo, i.normal), 1.0f);
    const float sinTheta = sqrtf(__fmaf_rn(-cosTheta, cosTheta, 1.0f));
    return (__fmul_rn(etaiOverEtat, sinTheta) > 1.0f || rnd(rng) < schlick(cosTheta, etaiOverEtat));
}

// coat_bsdf (material.h:62-70): library BSDF, unused by the staircase material table
__device__ __forceinline__ void coatBsdf(Scatter& out, const SurfacePoint& i, const f3& wo, float layerIor, const f3& glossyTint,
                                         float glossyFuzz, const f3& diffuseAlbedo, unsigned int& rng) {
    if (fresnelLayer(i, wo, layerIor, rng))
        glossyBsdf(out, i, wo, dot(wo, i.normal), glossyTint, glossyFuzz, rng);
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
        glossyBsdf(out, i, wo, dot(wo, i.normal), color, param, rng);
    else
        dielectricBsdf(out, i, wo, param, color, 0.0f, mk3(0.0f, 0.0f, 0.0f), rng);
}
