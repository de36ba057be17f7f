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

// ---- the rest of the reference's BSDF library (material.h) and its scene presets (scene_materials.h:22-93) -------------
// None of these is reachable from the staircase material table (material_scatter only dispatches DIFFUSE / METAL / GLASS);
// they are the reference's material "API surface" (SURVEY.md 8f rank 2). Checked against the reference's own device
// functions through oracle/ref_shim.cu (scatterBatch below, tests/test_parity_gpu.py::test_bsdf_library_vs_reference).
// checker_layer (material.h:33-36)
__device__ __forceinline__ bool checkerLayer(const f3& p, float frequency) {
    const float sines = sinf(frequency * p.x) * sinf(frequency * p.y) * sinf(frequency * p.z);
    return sines < 0;
}

// scatter_checker (material.h:39-44)
__device__ __forceinline__ void scatterChecker(Scatter& out, const SurfacePoint& i, const f3& p, float frequency, const f3& albedo1,
                                               const f3& albedo2, unsigned int& rng) {
    if (checkerLayer(p, frequency)) diffuseBsdf(out, i, albedo1, rng);
    else diffuseBsdf(out, i, albedo2, rng);
}

// subsurface_bsdf (material.h:94-113): free-flight distance inside the medium, isotropic scattering
__device__ __forceinline__ void subsurfaceBsdf(Scatter& out, const SurfacePoint& i, const f3& wo, const f3& absorption, float scatteringDistance,
                                               unsigned int& rng) {
    bool scattered = false;
    if (i.inside) {
        const float d = -logf(rnd(rng)) / scatteringDistance;
        if (d < i.t) {
            scattered = true;
            out.t = d;
        }
        const f3 e = -absorption * out.t;
        out.throughput = mk3(expf(e.x), expf(e.y), expf(e.z));
    }
    if (scattered) {
        out.wi = randomInUnitSphere(rng);
    } else {
        out.wi = wo; // ray doesn't change direction
        out.refracted = true;
    }
    out.specular = true;
}

// subsurface_dielectric_bsdf (material.h:115-143)
__device__ __forceinline__ void subsurfaceDielectricBsdf(Scatter& out, const SurfacePoint& i, const f3& wo, float layerIor, const f3& glossyTint,
                                                         float glossyFuzz, const f3& absorption, float scatteringDistance, unsigned int& rng) {
    bool scattered = false;
    if (i.inside) {
        const float d = -logf(rnd(rng)) / scatteringDistance;
        if (d < i.t) {
            scattered = true;
            out.t = d;
        }
        const f3 e = -absorption * out.t;
        out.throughput = mk3(expf(e.x), expf(e.y), expf(e.z));
    }
    if (scattered) {
        out.wi = randomInUnitSphere(rng);
        out.specular = true;
    } else {
        // the dielectric interface: fresnel_layer, then glossy reflection or refraction (the arithmetic of dielectricBsdf
        // above, which cannot be called here because it would overwrite the throughput when inside)
        const f3& n = i.normal;
        const float p1 = __fmul_rn(wo.x, n.x), p2 = __fmul_rn(wo.z, n.z);
        const float etaiOverEtat = i.inside ? layerIor : (1.0f / layerIor);
        const float cosTheta = fminf(__fsub_rn(__fmaf_rn(-wo.y, n.y, -p1), p2), 1.0f);
        const float sinTheta = sqrtf(__fmaf_rn(-cosTheta, cosTheta, 1.0f));
        if (__fmul_rn(etaiOverEtat, sinTheta) > 1.0f || rnd(rng) < schlick(cosTheta, etaiOverEtat)) {
            glossyBsdf(out, i, wo, __fadd_rn(p2, __fmaf_rn(wo.y, n.y, p1)), glossyTint, glossyFuzz, rng);
        } else {
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
}

// hexColor (scene_materials.h:6-11): the division is by the double literal 255.0 in the reference
__device__ __forceinline__ f3 hexColor(int hexValue) {
    const float r = (float)((hexValue >> 16) & 0xFF), g = (float)((hexValue >> 8) & 0xFF), b = (float)(hexValue & 0xFF);
    return mk3(r / 255.0f, g / 255.0f, b / 255.0f);
}

// The presets of scene_materials.h:22-93, by number (the order they appear in that file).
enum ScenePreset {
    PRESET_FLOOR_COAT = 0, PRESET_FLOOR_DIFFUSE = 1, PRESET_FLOOR_CHECKER = 2, PRESET_MODEL_COAT = 3, PRESET_MODEL_DIFFUSE = 4,
    PRESET_MODEL_GLOSSY = 5, PRESET_MODEL_GLASS = 6, PRESET_MODEL_TINTEDGLASS = 7, PRESET_MODEL_SSS = 8, PRESET_SUBSURFACE = 9,
    PRESET_COUNT = 10
};

__device__ __forceinline__ void presetScatter(int preset, Scatter& out, const SurfacePoint& i, const f3& p, const f3& wo, unsigned int& rng) {
    const f3 white = mk3(1.0f, 1.0f, 1.0f);
    const f3 model = mk3(0.0972942f, 0.0482054f, 0.000273194f);
    switch (preset) {
        case PRESET_FLOOR_COAT: coatBsdf(out, i, wo, 1.5f, white, 0.0f, hexColor(0x511845), rng); break;           // :22-28
        case PRESET_FLOOR_DIFFUSE: diffuseBsdf(out, i, hexColor(0x511845), rng); break;                            // :30-33
        case PRESET_FLOOR_CHECKER: scatterChecker(out, i, p, 0.2f, hexColor(0x511845), hexColor(0xff5733), rng); break; // :35-44
        case PRESET_MODEL_COAT: coatBsdf(out, i, wo, 1.1f, white, 0.0f, model, rng); break;                        // :46-52
        case PRESET_MODEL_DIFFUSE: diffuseBsdf(out, i, model, rng); break;                                        // :54-57
        case PRESET_MODEL_GLOSSY: glossyBsdf(out, i, wo, dot(wo, i.normal), white, 0.0f, rng); break;              // :59-63
        case PRESET_MODEL_GLASS: dielectricBsdf(out, i, wo, 1.1f, white, 0.0f, mk3(0.0f, 0.0f, 0.0f), rng); break; // :65-71
        case PRESET_MODEL_TINTEDGLASS: {                                                                           // :73-81
            const f3 absorption = mk3(-logf(model.x) / 10.0f, -logf(model.y) / 10.0f, -logf(model.z) / 10.0f);
            dielectricBsdf(out, i, wo, 1.1f, white, 0.0f, absorption, rng);
            break;
        }
        case PRESET_MODEL_SSS: subsurfaceDielectricBsdf(out, i, wo, 1.333f, white, 0.0f, mk3(0.9f, 0.3f, 0.02f), 2.0f, rng); break; // :83-93
        default: subsurfaceBsdf(out, i, wo, mk3(0.9f, 0.3f, 0.02f), 2.0f, rng); break;                              // material.h:94 with the SSS constants
    }
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
