// cpu_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A scalar (OpenMP over pixel rows) CPU restatement of the reference's render path, used only by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker / baseline.  Nothing
// under cuda-raytracing-optimized_b200/ links, loads or calls it.
//
// Each function names the reference lines it restates.  Two deliberate differences from a host compile
// of the reference headers:
//   * random vectors are drawn x, then y, then z in separate statements (g++ would evaluate
//     `vec3(rnd(s), rnd(s), rnd(s))` right to left; the reference's device build draws left to right);
//   * built with -ffp-contract=off, so results do not depend on the host CPU.  The GPU fuses
//     multiply-adds, so float results agree with the CUDA reference to rounding, not bit for bit;
//     integer results (RNG, ids away from edges) agree exactly.
// Pinning: tests/test_oracle_golden.py checks this file against golden vectors produced by the
// reference's own CUDA kernel (oracle/_ref/libref*.so) on a B200 -- see tests/golden/README.md.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/rt_types.h"

namespace {

struct V3 {
    float x, y, z;
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 v3(const vec3& v) { return V3{v.e[0], v.e[1], v.e[2]}; }
// vec3.h:62-104: operator shapes
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator*(float t, V3 v) { return v3(t * v.x, t * v.y, t * v.z); }
inline V3 operator*(V3 v, float t) { return v3(t * v.x, t * v.y, t * v.z); }
inline V3 operator/(V3 v, float t) { return v3(v.x / t, v.y / t, v.z / t); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { // vec3.h:100-104
    return v3((a.y * b.z - a.z * b.y), (-(a.x * b.z - a.z * b.x)), (a.x * b.y - a.y * b.x));
}
inline float sqlen(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline float len(V3 a) { return std::sqrt(sqlen(a)); }
inline V3 unit(V3 a) { return a / len(a); } // vec3.h:194
inline float at(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// ------------------------------------------------------------------- rnd.h --
inline uint32_t wang_hash(uint32_t seed) { // rnd.h:31-39
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}
inline uint32_t xor_shift_32(uint32_t& state) { // rnd.h:5-13
    uint32_t x = state;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 15;
    state = x;
    return x;
}
inline float rnd(uint32_t& state) { return (xor_shift_32(state) & 0xFFFFFF) / 16777216.0f; } // rnd.h:15-18
inline uint32_t path_seed(uint32_t id) { return (wang_hash(id) * 336343633u) | 1u; }           // kernels.cu:542

V3 random_in_unit_disk(uint32_t& s) { // rnd.h:20-27
    V3 p;
    do {
        float a = rnd(s);
        float b = rnd(s);
        p = 2.0f * v3(a, b, 0) - v3(1, 1, 0);
    } while (dot(p, p) >= 1.0f);
    return p;
}
V3 random_in_unit_sphere(uint32_t& s) { // rnd.h:43-49
    V3 p;
    do {
        float a = rnd(s);
        float b = rnd(s);
        float c = rnd(s);
        p = 2.0f * v3(a, b, c) - v3(1, 1, 1);
    } while (sqlen(p) >= 1.0f);
    return p;
}

// ------------------------------------------------------------------- ray.h --
struct Ray {
    V3 A, B;
    Ray(V3 a, V3 b) : A(a), B(unit(b)) {} // ray.h:9 normalises
    V3 at(float t) const { return A + t * B; }
};

Ray get_ray(const camera& c, float s, float t, uint32_t& state) { // camera.h:8-12
    V3 rd = c.lens_radius * random_in_unit_disk(state);
    V3 offset = v3(c.u) * rd.x + v3(c.v) * rd.y;
    return Ray(v3(c.origin) + offset, v3(c.lower_left_corner) + s * v3(c.horizontal) + t * v3(c.vertical) - v3(c.origin) - offset);
}

// --------------------------------------------------------- intersections.h --
bool hit_bbox(V3 bmin, V3 bmax, const Ray& r, float t_max) { // intersections.h:7-23
    float t_min = 0.001f;
    for (int a = 0; a < 3; a++) {
        float invD = 1.0f / at(r.B, a);
        float t0 = (at(bmin, a) - at(r.A, a)) * invD;
        float t1 = (at(bmax, a) - at(r.A, a)) * invD;
        if (invD < 0.0f) std::swap(t0, t1);
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max < t_min) return false;
    }
    return true;
}
float hit_bbox_dist(V3 bmin, V3 bmax, const Ray& r, float t_max) { // intersections.h:25-41
    float t_min = 0.001f;
    for (int a = 0; a < 3; a++) {
        float invD = 1.0f / at(r.B, a);
        float t0 = (at(bmin, a) - at(r.A, a)) * invD;
        float t1 = (at(bmax, a) - at(r.A, a)) * invD;
        if (invD < 0.0f) std::swap(t0, t1);
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max < t_min) return FLT_MAX;
    }
    return t_min;
}
float triangleHit(const triangle& tri, const Ray& r, float t_min, float t_max, float& hitU, float& hitV) { // intersections.h:54-83
    const float EPS = 0.0000001f;
    V3 v0 = v3(tri.v[0]);
    V3 edge1 = v3(tri.v[1]) - v0;
    V3 edge2 = v3(tri.v[2]) - v0;
    V3 h = cross(r.B, edge2);
    float a = dot(edge1, h);
    if (a > -EPS && a < EPS) return FLT_MAX;
    float f = 1.0f / a;
    V3 s = r.A - v0;
    float u = f * dot(s, h);
    if (u < 0.0f || u > 1.0f) return FLT_MAX;
    V3 q = cross(s, edge1);
    float v = f * dot(r.B, q);
    if (v < 0.0f || u + v > 1.0f) return FLT_MAX;
    float t = f * dot(edge2, q);
    if (t > t_min && t < t_max) {
        hitU = u;
        hitV = v;
        return t;
    }
    return FLT_MAX;
}
float sphereHit(V3 center, float radius, const Ray& r, float t_min, float t_max) { // intersections.h:85-104
    V3 oc = r.A - center;
    float a = dot(r.B, r.B);
    float b = dot(oc, r.B);
    float c = dot(oc, oc) - radius * radius;
    float discriminant = b * b - a * c;
    if (discriminant > 0) {
        float temp = (-b - std::sqrt(discriminant)) / a;
        if (temp < t_max && temp > t_min) return temp;
        temp = (-b + std::sqrt(discriminant)) / a;
        if (temp < t_max && temp > t_min) return temp;
    }
    return FLT_MAX;
}

// -------------------------------------------------------------- kernels.cu --
#define EPSILON 0.01f // kernels.cu:19

struct Counters {
    uint64_t primary = 0, secondary = 0, shadow = 0, nodeVisits = 0, triTests = 0;
};

struct Ctx { // RenderContext, kernels.cu:69-145
    const triangle* tris;
    const bvh_node* bvh;
    uint32_t firstLeafIdx, numPrimitivesPerLeaf;
    bbox bounds;
    int nx, ny, ns, maxDepth;
    camera cam;
    V3 lightCenter = v3(52.514355f, 715.686951f, -272.620972f); // kernels.cu:93
    float lightRadius = 50.0f;
    V3 lightColor = v3(20.0f, 20.0f, 20.0f); // kernels.cu:94
    const material* materials;
    const stexture* textures;
};

struct TriHit {
    uint32_t triId;
    float u, v;
};

inline int ffs32(uint32_t x) { return __builtin_ffs((int)x); }
inline void pop_bitstack(uint32_t& bitStack, int& idx) { // kernels.cu:148-152
    int m = ffs32(bitStack) - 1;
    bitStack = (bitStack >> m) ^ 1;
    idx = (idx >> m) ^ 1;
}

#ifdef ORACLE_STEPLOG
// Design aid (oracle/sched_sim.py): the sequence of steps every traversal takes -- 0 = dual-node step, 1+k = leaf visit
// with k triangle tests, 255 = end of ray (254 = end of an any-hit ray). Single-threaded runs only.
std::vector<uint8_t>* g_stepLog = nullptr;
#define STEPLOG(x) do { if (g_stepLog) g_stepLog->push_back((uint8_t)(x)); } while (0)
#else
#define STEPLOG(x) do { } while (0)
#endif

float hitBvh(const Ray& r, const Ctx& c, float t_min, float t_max, TriHit& rec, bool isShadow, Counters* cnt) { // kernels.cu:154-224
    int idx = 1;
    float closest = t_max;
    uint32_t bitStack = 1;
    while (idx) {
        if ((uint32_t)idx < c.firstLeafIdx) {
            int idx2 = idx << 1;
            const bvh_node& left = c.bvh[idx2];
            const bvh_node& right = c.bvh[idx2 + 1];
            if (cnt) cnt->nodeVisits++;
            STEPLOG(0);
            float leftHit = hit_bbox_dist(v3(left.a), v3(left.b), r, closest);
            bool traverseLeft = leftHit < closest;
            float rightHit = hit_bbox_dist(v3(right.a), v3(right.b), r, closest);
            bool traverseRight = rightHit < closest;
            bool swap = rightHit < leftHit;
            if (traverseLeft && traverseRight) {
                idx = idx2 + (swap ? 1 : 0);
                bitStack = (bitStack << 1) + 1;
            } else if (traverseLeft || traverseRight) {
                idx = idx2 + (swap ? 1 : 0);
                bitStack = bitStack << 1;
            } else {
                pop_bitstack(bitStack, idx);
            }
        } else {
            int first = (idx - c.firstLeafIdx) * c.numPrimitivesPerLeaf;
#ifdef ORACLE_STEPLOG
            if (g_stepLog) {
                uint32_t k = 0;
                while (k < c.numPrimitivesPerLeaf && !std::isinf(c.tris[first + k].v[0].e[0])) k++;
                STEPLOG(1 + k);
            }
#endif
            for (uint32_t i = 0; i < c.numPrimitivesPerLeaf; i++) {
                const triangle& tri = c.tris[first + i];
                if (std::isinf(tri.v[0].e[0])) break;
                if (cnt) cnt->triTests++;
                float u, v;
                float hitT = triangleHit(tri, r, t_min, closest, u, v);
                if (hitT < closest) {
                    if (isShadow) { STEPLOG(254); return 0.0f; }
                    closest = hitT;
                    rec.triId = first + i;
                    rec.u = u;
                    rec.v = v;
                }
            }
            pop_bitstack(bitStack, idx);
        }
    }
    STEPLOG(isShadow ? 254 : 255);
    return closest;
}

float hitMesh(const Ray& r, const Ctx& c, float t_min, float t_max, TriHit& rec, bool isShadow, Counters* cnt) { // kernels.cu:296-323
    if (!hit_bbox(v3(c.bounds.min), v3(c.bounds.max), r, t_max)) return FLT_MAX;
    return hitBvh(r, c, t_min, t_max, rec, isShadow, cnt);
}

enum { NONE = 0, TRIMESH = 1, PLANE = 2, LIGHT = 3 }; // kernels.cu:40-45

struct Inter { // intersection, helper_structs.h:16-36
    unsigned objId;
    unsigned char meshID;
    float t;
    V3 normal;
    bool inside;
    float texCoords[2];
};

struct Path { // helper_structs.h:48-71
    V3 origin, rayDir, color, shadowDir, lightContribution, attenuation;
    bool specular, inside;
    uint8_t bounce;
    uint32_t rng;
};

bool hit(const Ctx& c, const Path& p, float t_max, bool isShadow, Inter& inters, Counters* cnt) { // kernels.cu:325-360
    const Ray r = isShadow ? Ray(p.origin, p.shadowDir) : Ray(p.origin, p.rayDir);
    TriHit triHit{0, 0, 0};
    inters.objId = NONE;
    if ((inters.t = hitMesh(r, c, EPSILON, t_max, triHit, isShadow, cnt)) < t_max) {
        if (isShadow) return true;
        inters.objId = TRIMESH;
        const triangle& tri = c.tris[triHit.triId];
        inters.meshID = tri.meshID;
        inters.normal = unit(cross(v3(tri.v[1]) - v3(tri.v[0]), v3(tri.v[2]) - v3(tri.v[0])));
        inters.texCoords[0] = (triHit.u * tri.texCoords[1 * 2 + 0] + triHit.v * tri.texCoords[2 * 2 + 0] +
                               (1 - triHit.u - triHit.v) * tri.texCoords[0 * 2 + 0]);
        inters.texCoords[1] = (triHit.u * tri.texCoords[1 * 2 + 1] + triHit.v * tri.texCoords[2 * 2 + 1] +
                               (1 - triHit.u - triHit.v) * tri.texCoords[0 * 2 + 1]);
    } else {
        if (isShadow) return false;
        if (p.specular && sphereHit(c.lightCenter, c.lightRadius, r, EPSILON, t_max) < t_max) {
            inters.objId = LIGHT;
            return true;
        }
    }
    if (inters.objId != NONE) {
        if (dot(r.B, inters.normal) > 0.0f) inters.normal = -inters.normal;
        return true;
    }
    return false;
}

// -------------------------------------------------------------- material.h --
struct Scatter { // scatter_info, helper_structs.h:38-46
    V3 wi;
    bool specular;
    V3 throughput;
    bool refracted;
    float t;
};

float schlick(float cosine, float ref_idx) { // material.h:9-13
    float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
    r0 = r0 * r0;
    return r0 + (1.0f - r0) * std::pow((1.0f - cosine), 5.0f);
}
V3 refract(V3 uv, V3 n, float etai_over_etat) { // material.h:15-21
    float cos_theta = std::fmin(dot(-uv, n), 1.0f);
    V3 r_out_parallel = etai_over_etat * (uv + cos_theta * n);
    float sq = sqlen(r_out_parallel);
    V3 r_out_perp = sq >= 1.0f ? v3(0, 0, 0) : -std::sqrt(1.0f - sq) * n;
    return r_out_parallel + r_out_perp;
}
V3 reflect(V3 v, V3 n) { return v - 2.0f * dot(v, n) * n; } // material.h:23-25

void diffuse_bsdf(Scatter& out, const Inter& i, V3 albedo, uint32_t& rng) { // material.h:27-31
    out.wi = unit(i.normal + random_in_unit_sphere(rng));
    out.throughput = albedo;
    out.specular = false;
}
void glossy_bsdf(Scatter& out, const Inter& i, V3 wo, V3 tint, float fuzz, uint32_t& rng) { // material.h:46-53
    V3 reflected = reflect(wo, i.normal);
    if (fuzz > 0.0001f) reflected = reflected + fuzz * random_in_unit_sphere(rng);
    out.wi = unit(reflected);
    out.throughput = out.throughput * tint;
    out.specular = true;
}
bool fresnel_layer(const Inter& i, V3 wo, float ior, uint32_t& rng) { // material.h:55-60
    float etai_over_etat = i.inside ? ior : (1.0f / ior);
    float cos_theta = std::fmin(dot(-wo, i.normal), 1.0f);
    float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
    return (etai_over_etat * sin_theta > 1.0f || rnd(rng) < schlick(cos_theta, etai_over_etat));
}
void dielectric_bsdf(Scatter& out, const Inter& i, V3 wo, float layer_ior, V3 glossy_tint, float glossy_fuzz, V3 absorption,
                     uint32_t& rng) { // material.h:73-92
    if (i.inside) {
        V3 e = -absorption * i.t;
        out.throughput = v3(std::exp(e.x), std::exp(e.y), std::exp(e.z));
    }
    if (fresnel_layer(i, wo, layer_ior, rng)) {
        glossy_bsdf(out, i, wo, glossy_tint, glossy_fuzz, rng);
    } else {
        float etai_over_etat = i.inside ? layer_ior : (1.0f / layer_ior);
        out.wi = unit(refract(wo, i.normal, etai_over_etat));
        out.refracted = true;
    }
    out.specular = true;
}
void material_scatter(Scatter& out, const Inter& i, V3 wo, const material& mat, V3 color, uint32_t& rng) { // scene_materials.h:13-20
    if (mat.type == DIFFUSE) diffuse_bsdf(out, i, color, rng);
    else if (mat.type == METAL) glossy_bsdf(out, i, wo, color, mat.param, rng);
    else dielectric_bsdf(out, i, wo, mat.param, color, 0.0f, v3(0, 0, 0), rng);
}

// ---- the rest of material.h and the presets of scene_materials.h:22-93 (unused by the staircase table) ----
bool checker_layer(V3 p, float frequency) { // material.h:33-36
    const float sines = std::sin(frequency * p.x) * std::sin(frequency * p.y) * std::sin(frequency * p.z);
    return sines < 0;
}
void coat_bsdf(Scatter& out, const Inter& i, V3 wo, float layer_ior, V3 glossy_tint, float glossy_fuzz, V3 diffuse_albedo,
               uint32_t& rng) { // material.h:62-70
    if (fresnel_layer(i, wo, layer_ior, rng)) glossy_bsdf(out, i, wo, glossy_tint, glossy_fuzz, rng);
    else diffuse_bsdf(out, i, diffuse_albedo, rng);
}
void subsurface_bsdf(Scatter& out, const Inter& i, V3 wo, V3 absorption, float scatteringDistance, uint32_t& rng) { // material.h:94-113
    bool scattered = false;
    if (i.inside) {
        const float d = -std::log(rnd(rng)) / scatteringDistance;
        if (d < i.t) { scattered = true; out.t = d; }
        const V3 e = -absorption * out.t;
        out.throughput = v3(std::exp(e.x), std::exp(e.y), std::exp(e.z));
    }
    if (scattered) {
        out.wi = random_in_unit_sphere(rng);
    } else {
        out.wi = wo;
        out.refracted = true;
    }
    out.specular = true;
}
void subsurface_dielectric_bsdf(Scatter& out, const Inter& i, V3 wo, float layer_ior, V3 glossy_tint, float glossy_fuzz, V3 absorption,
                                float scatteringDistance, uint32_t& rng) { // material.h:115-143
    bool scattered = false;
    if (i.inside) {
        const float d = -std::log(rnd(rng)) / scatteringDistance;
        if (d < i.t) { scattered = true; out.t = d; }
        const V3 e = -absorption * out.t;
        out.throughput = v3(std::exp(e.x), std::exp(e.y), std::exp(e.z));
    }
    if (scattered) {
        out.wi = random_in_unit_sphere(rng);
    } else if (fresnel_layer(i, wo, layer_ior, rng)) {
        glossy_bsdf(out, i, wo, glossy_tint, glossy_fuzz, rng);
    } else {
        const float etai_over_etat = i.inside ? layer_ior : (1.0f / layer_ior);
        out.wi = unit(refract(wo, i.normal, etai_over_etat));
        out.refracted = true;
    }
    out.specular = true;
}
V3 hexColor(int hexValue) { // scene_materials.h:6-11
    const float r = (float)((hexValue >> 16) & 0xFF), g = (float)((hexValue >> 8) & 0xFF), b = (float)(hexValue & 0xFF);
    return v3((float)(r / 255.0), (float)(g / 255.0), (float)(b / 255.0));
}
void preset_scatter(int preset, Scatter& out, const Inter& i, V3 p, V3 wo, uint32_t& rng) { // scene_materials.h:22-93, in file order
    const V3 white = v3(1, 1, 1), model = v3(0.0972942f, 0.0482054f, 0.000273194f);
    switch (preset) {
        case 0: coat_bsdf(out, i, wo, 1.5f, white, 0.0f, hexColor(0x511845), rng); break;
        case 1: diffuse_bsdf(out, i, hexColor(0x511845), rng); break;
        case 2: diffuse_bsdf(out, i, checker_layer(p, 0.2f) ? hexColor(0x511845) : hexColor(0xff5733), rng); break;
        case 3: coat_bsdf(out, i, wo, 1.1f, white, 0.0f, model, rng); break;
        case 4: diffuse_bsdf(out, i, model, rng); break;
        case 5: glossy_bsdf(out, i, wo, white, 0.0f, rng); break;
        case 6: dielectric_bsdf(out, i, wo, 1.1f, white, 0.0f, v3(0, 0, 0), rng); break;
        case 7: {
            const V3 absorption = v3(-std::log(model.x) / 10.0f, -std::log(model.y) / 10.0f, -std::log(model.z) / 10.0f);
            dielectric_bsdf(out, i, wo, 1.1f, white, 0.0f, absorption, rng);
            break;
        }
        case 8: subsurface_dielectric_bsdf(out, i, wo, 1.333f, white, 0.0f, v3(0.9f, 0.3f, 0.02f), 2.0f, rng); break;
        default: subsurface_bsdf(out, i, wo, v3(0.9f, 0.3f, 0.02f), 2.0f, rng); break;
    }
}

bool generateShadowRay(const Ctx& c, Path& p, const Inter& inters, float& lightDist) { // kernels.cu:363-393
    const V3 sw = unit(c.lightCenter - p.origin);
    const V3 su = unit(cross(std::fabs(sw.x) > 0.01f ? v3(0, 1, 0) : v3(1, 0, 0), sw));
    const V3 sv = cross(sw, su);
    const float cosAMax = std::sqrt(1.0f - c.lightRadius * c.lightRadius / sqlen(p.origin - c.lightCenter));
    if (std::isnan(cosAMax)) return false;
    const float eps1 = rnd(p.rng);
    const float eps2 = rnd(p.rng);
    const float cosA = 1.0f - eps1 + eps1 * cosAMax;
    const float sinA = std::sqrt(1.0f - cosA * cosA);
    const float phi = (float)(2 * M_PI * eps2);
    const V3 l = su * std::cos(phi) * sinA + sv * std::sin(phi) * sinA + sw * cosA;
    const float dotl = dot(l, inters.normal);
    if (dotl <= 0) return false;
    p.shadowDir = unit(l);
    const float omega = (float)(2 * M_PI * (1.0f - cosAMax));
    p.lightContribution = p.attenuation * c.lightColor * dotl * omega / (float)M_PI;
    lightDist = len(c.lightCenter - p.origin) - c.lightRadius;
    return true;
}

void color(const Ctx& c, Path& p, Counters* cnt) { // kernels.cu:396-533
    p.attenuation = v3(1, 1, 1);
    p.color = v3(0, 0, 0);
    for (p.bounce = 0; p.bounce < c.maxDepth; p.bounce++) {
        if (cnt) { if (p.bounce == 0) cnt->primary++; else cnt->secondary++; }
        Inter inters;
        if (!hit(c, p, FLT_MAX, false, inters, cnt)) {
            p.color = p.color + p.attenuation * v3(0.5f, 0.5f, 0.5f); // kernels.cu:424
            return;
        }
        if (inters.objId == LIGHT) return; // kernels.cu:433-446: SHADOW is defined, no emission added
        inters.inside = p.inside;
        Scatter scatter;
        scatter.specular = false;
        scatter.throughput = v3(1, 1, 1);
        scatter.refracted = false;
        scatter.t = inters.t;
        scatter.wi = v3(0, 0, 0);
        const material& mat = c.materials[inters.meshID];
        V3 albedo;
        if (mat.texId != -1) { // kernels.cu:457-471
            const stexture& tex = c.textures[mat.texId];
            float tu = inters.texCoords[0];
            tu = tu - std::floor(tu);
            float tv = inters.texCoords[1];
            tv = tv - std::floor(tv);
            const int tx = (tex.width - 1) * tu;
            const int ty = (tex.height - 1) * tv;
            const int tIdx = ty * tex.width + tx;
            albedo = v3(tex.data[tIdx * 3 + 0], tex.data[tIdx * 3 + 1], tex.data[tIdx * 3 + 2]);
        } else {
            albedo = v3(mat.color);
        }
        material_scatter(scatter, inters, p.rayDir, mat, albedo, p.rng);
        p.origin = p.origin + scatter.t * p.rayDir; // kernels.cu:485
        p.rayDir = scatter.wi;
        p.attenuation = p.attenuation * scatter.throughput;
        p.specular = scatter.specular;
        p.inside = scatter.refracted ? !p.inside : p.inside;
        float lightDist;
        if (!p.specular && generateShadowRay(c, p, inters, lightDist)) {
            if (cnt) cnt->shadow++;
            Inter sh;
            if (!hit(c, p, lightDist, true, sh, cnt)) p.color = p.color + p.lightContribution;
        }
        if (p.bounce > 3) { // kernels.cu:514-526
            float m = std::fmax(p.attenuation.x, std::fmax(p.attenuation.y, p.attenuation.z));
            if (rnd(p.rng) > m) return;
            p.attenuation = p.attenuation * (1 / m);
        }
    }
}

// ----------------------------------------------- sphere scenes (reference-derived, see csrc/spheres_path.cuh) --
void colorSpheres(const sphere* sph, const material* mats, int n, int maxDepth, Path& p, Counters* cnt) {
    p.attenuation = v3(1, 1, 1);
    p.color = v3(0, 0, 0);
    for (p.bounce = 0; p.bounce < maxDepth; p.bounce++) {
        if (cnt) { if (p.bounce == 0) cnt->primary++; else cnt->secondary++; }
        const Ray r(p.origin, p.rayDir);
        float closest = FLT_MAX;
        int id = -1;
        for (int s = 0; s < n; s++) {
            float t = sphereHit(v3(sph[s].center), sph[s].radius, r, EPSILON, closest);
            if (t < closest) { closest = t; id = s; }
        }
        if (id < 0) { // kernels.cu:419-421
            float t = 0.5f * (p.rayDir.y + 1.0f);
            V3 c = (1.0f - t) * v3(1.0f, 1.0f, 1.0f) + t * v3(0.5f, 0.7f, 1.0f);
            p.color = p.color + p.attenuation * c;
            return;
        }
        Inter inters;
        inters.objId = TRIMESH;
        inters.t = closest;
        inters.normal = (r.at(closest) - v3(sph[id].center)) / sph[id].radius;
        if (dot(r.B, inters.normal) > 0.0f) inters.normal = -inters.normal;
        inters.inside = p.inside;
        Scatter scatter;
        scatter.specular = false;
        scatter.throughput = v3(1, 1, 1);
        scatter.refracted = false;
        scatter.t = inters.t;
        scatter.wi = v3(0, 0, 0);
        material_scatter(scatter, inters, p.rayDir, mats[id], v3(mats[id].color), p.rng);
        p.origin = p.origin + scatter.t * p.rayDir;
        p.rayDir = scatter.wi;
        p.attenuation = p.attenuation * scatter.throughput;
        p.specular = scatter.specular;
        p.inside = scatter.refracted ? !p.inside : p.inside;
        if (p.bounce > 3) {
            float m = std::fmax(p.attenuation.x, std::fmax(p.attenuation.y, p.attenuation.z));
            if (rnd(p.rng) > m) return;
            p.attenuation = p.attenuation * (1 / m);
        }
    }
}

Ctx makeCtx(const kernel_scene* sc, const camera* cam, int nx, int ny, int ns, int maxDepth) {
    Ctx c;
    c.tris = sc->m->tris;
    c.bvh = sc->m->bvh;
    c.firstLeafIdx = (uint32_t)(sc->m->numBvhNodes / 2); // kernels.cu:614
    c.numPrimitivesPerLeaf = (uint32_t)sc->numPrimitivesPerLeaf;
    c.bounds = sc->m->bounds;
    c.nx = nx; c.ny = ny; c.ns = ns; c.maxDepth = maxDepth;
    if (cam) c.cam = *cam;
    c.materials = sc->materials;
    c.textures = sc->textures;
    return c;
}

} // namespace

extern "C" {

// The BSDF presets on a batch of surface points; item layout as scatterBatch (include/kernels.h): in 12 floats
// {normal.xyz, t, p.xyz, inside, wo.xyz, rng bits}, out 12 floats {wi.xyz, t, throughput.xyz, flags, rng bits after, 0, 0, 0}.
int oracleScatterBatch(int preset, long long n, const float* in, float* out) {
    if (preset < 0 || preset > 9) return -1;
    for (long long k = 0; k < n; k++) {
        const float* a = in + 12 * k;
        Inter i;
        i.objId = TRIMESH;
        i.meshID = 0;
        i.normal = v3(a[0], a[1], a[2]);
        i.t = a[3];
        i.inside = a[7] != 0.0f;
        i.texCoords[0] = i.texCoords[1] = 0.0f;
        uint32_t rng;
        std::memcpy(&rng, &a[11], 4);
        Scatter s;
        s.wi = v3(0, 0, 0);
        s.specular = false;
        s.throughput = v3(1, 1, 1);
        s.refracted = false;
        s.t = i.t;
        preset_scatter(preset, s, i, v3(a[4], a[5], a[6]), v3(a[8], a[9], a[10]), rng);
        float* o = out + 12 * k;
        o[0] = s.wi.x; o[1] = s.wi.y; o[2] = s.wi.z; o[3] = s.t;
        o[4] = s.throughput.x; o[5] = s.throughput.y; o[6] = s.throughput.z;
        const int flags = (s.specular ? 1 : 0) | (s.refracted ? 2 : 0);
        std::memcpy(&o[7], &flags, 4);
        std::memcpy(&o[8], &rng, 4);
        o[9] = o[10] = o[11] = 0.0f;
    }
    return 0;
}

// render(), kernels.cu:535-569, for rows rowBegin, rowBegin+rowStride, ... < rowEnd (other rows of fb are left untouched). counters (may be NULL): primary, secondary, shadow, nodeVisits, triTests.
void oracleRender(const kernel_scene* sc, const camera* cam, int nx, int ny, int ns, int maxDepth, unsigned int sampleStream,
                  int rowBegin, int rowEnd, int rowStride, vec3* fb, unsigned long long* counters) {
    const Ctx c = makeCtx(sc, cam, nx, ny, ns, maxDepth);
    Counters total;
#pragma omp parallel
    {
        Counters local;
#pragma omp for schedule(dynamic, 1)
        for (int j = rowBegin; j < rowEnd; j += rowStride)
            for (int i = 0; i < nx; i++) {
                Path p;
                uint32_t pixelId = (uint32_t)(j * nx + i);
                p.rng = path_seed(pixelId + sampleStream * (uint32_t)(nx * ny));
                V3 col = v3(0, 0, 0);
                for (int s = 0; s < ns; s++) {
                    float u = float(i + rnd(p.rng)) / float(nx);
                    float v = float(j + rnd(p.rng)) / float(ny);
                    Ray r = get_ray(c.cam, u, v, p.rng);
                    p.origin = r.A;
                    p.rayDir = r.B;
                    p.specular = false;
                    p.inside = false;
                    color(c, p, counters ? &local : nullptr);
                    col = col + p.color;
                }
                col = col / float(ns);
                fb[pixelId].e[0] = col.x; fb[pixelId].e[1] = col.y; fb[pixelId].e[2] = col.z;
            }
#pragma omp critical
        {
            total.primary += local.primary; total.secondary += local.secondary; total.shadow += local.shadow;
            total.nodeVisits += local.nodeVisits; total.triTests += local.triTests;
        }
    }
    if (counters) {
        counters[0] = total.primary; counters[1] = total.secondary; counters[2] = total.shadow;
        counters[3] = total.nodeVisits; counters[4] = total.triTests;
    }
}

void oracleRenderSpheres(const sphere* sph, const material* mats, int n, const camera* cam, int nx, int ny, int ns, int maxDepth,
                         unsigned int sampleStream, int rowBegin, int rowEnd, int rowStride, vec3* fb, unsigned long long* counters) {
    Counters total;
#pragma omp parallel
    {
        Counters local;
#pragma omp for schedule(dynamic, 1)
        for (int j = rowBegin; j < rowEnd; j += rowStride)
            for (int i = 0; i < nx; i++) {
                Path p;
                uint32_t pixelId = (uint32_t)(j * nx + i);
                p.rng = path_seed(pixelId + sampleStream * (uint32_t)(nx * ny));
                V3 col = v3(0, 0, 0);
                for (int s = 0; s < ns; s++) {
                    float u = float(i + rnd(p.rng)) / float(nx);
                    float v = float(j + rnd(p.rng)) / float(ny);
                    Ray r = get_ray(*cam, u, v, p.rng);
                    p.origin = r.A;
                    p.rayDir = r.B;
                    p.specular = false;
                    p.inside = false;
                    colorSpheres(sph, mats, n, maxDepth, p, counters ? &local : nullptr);
                    col = col + p.color;
                }
                col = col / float(ns);
                fb[pixelId].e[0] = col.x; fb[pixelId].e[1] = col.y; fb[pixelId].e[2] = col.z;
            }
#pragma omp critical
        { total.primary += local.primary; total.secondary += local.secondary; }
    }
    if (counters) { counters[0] = total.primary; counters[1] = total.secondary; counters[2] = 0; counters[3] = 0; counters[4] = 0; }
}

// hit()/hitMesh() on a ray batch: rays as float4 {o.xyz,tMin} {d.xyz,tMax}; out float4 {t,u,v,triId bits}, meshID.
// anyHit != 0: the isShadow walk; out t = 0 when occluded, FLT_MAX otherwise.
void oracleIntersectBatch(const kernel_scene* sc, const float* rayO, const float* rayD, long long n, int anyHit, float* outHit,
                          int* outMesh, unsigned long long* counters) {
    const Ctx c = makeCtx(sc, nullptr, 0, 0, 0, 0);
    Counters total;
#pragma omp parallel
    {
        Counters local;
#pragma omp for schedule(dynamic, 1024)
        for (long long i = 0; i < n; i++) {
            const Ray r(v3(rayO[4 * i], rayO[4 * i + 1], rayO[4 * i + 2]), v3(rayD[4 * i], rayD[4 * i + 1], rayD[4 * i + 2]));
            const float tMin = rayO[4 * i + 3], tMax = rayD[4 * i + 3];
            TriHit th{0xFFFFFFFFu, 0, 0};
            float t = hitMesh(r, c, tMin, tMax, th, anyHit != 0, counters ? &local : nullptr);
            int meshID = -1;
            if (t < tMax) {
                if (anyHit) { t = 0.0f; th = TriHit{0xFFFFFFFFu, 0, 0}; }
                else meshID = c.tris[th.triId].meshID;
            } else {
                t = FLT_MAX;
                th = TriHit{0xFFFFFFFFu, 0, 0};
            }
            outHit[4 * i] = t; outHit[4 * i + 1] = th.u; outHit[4 * i + 2] = th.v;
            std::memcpy(&outHit[4 * i + 3], &th.triId, 4);
            if (outMesh) outMesh[i] = meshID;
        }
#pragma omp critical
        { total.nodeVisits += local.nodeVisits; total.triTests += local.triTests; }
    }
    if (counters) { counters[3] = total.nodeVisits; counters[4] = total.triTests; }
}

// RNG known-answer helpers
unsigned int oracleWangHash(unsigned int x) { return wang_hash(x); }
unsigned int oraclePathSeed(unsigned int pixelId) { return path_seed(pixelId); }
unsigned int oracleXorShift(unsigned int* state) { return xor_shift_32(*state); }
float oracleRnd(unsigned int* state) { return rnd(*state); }
void oracleUnitSphere(unsigned int* state, float out[3]) { V3 p = random_in_unit_sphere(*state); out[0] = p.x; out[1] = p.y; out[2] = p.z; }
void oracleUnitDisk(unsigned int* state, float out[3]) { V3 p = random_in_unit_disk(*state); out[0] = p.x; out[1] = p.y; out[2] = p.z; }
void oracleGetRay(const camera* c, float s, float t, unsigned int* state, float o[3], float d[3]) {
    Ray r = get_ray(*c, s, t, *state);
    o[0] = r.A.x; o[1] = r.A.y; o[2] = r.A.z; d[0] = r.B.x; d[1] = r.B.y; d[2] = r.B.z;
}
float oracleTriangleHit(const triangle* tri, const float o[3], const float d[3], float tMin, float tMax, float* u, float* v) {
    Ray r(v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]));
    return triangleHit(*tri, r, tMin, tMax, *u, *v);
}
float oracleSphereHit(const sphere* s, const float o[3], const float d[3], float tMin, float tMax) {
    Ray r(v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]));
    return sphereHit(v3(s->center), s->radius, r, tMin, tMax);
}
float oracleBoxDist(const float bmin[3], const float bmax[3], const float o[3], const float d[3], float tMax) {
    Ray r(v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]));
    return hit_bbox_dist(v3(bmin[0], bmin[1], bmin[2]), v3(bmax[0], bmax[1], bmax[2]), r, tMax);
}
#ifdef ORACLE_STEPLOG
// start logging (call with OMP_NUM_THREADS=1), then fetch: returns the number of bytes; copies min(n, cap) of them
void oracleStepLogStart() {
    delete g_stepLog;
    g_stepLog = new std::vector<uint8_t>();
}
unsigned long long oracleStepLogFetch(unsigned char* out, unsigned long long cap) {
    if (!g_stepLog) return 0;
    const unsigned long long n = g_stepLog->size();
    if (out) memcpy(out, g_stepLog->data(), n < cap ? n : cap);
    return n;
}
#endif
int oracleNumThreads() {
    int n = 1;
#ifdef _OPENMP
#pragma omp parallel
    {
#pragma omp master
        n = omp_get_num_threads();
    }
#endif
    return n;
}

} // extern "C"
