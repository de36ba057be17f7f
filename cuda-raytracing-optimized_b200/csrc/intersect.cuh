// intersect.cuh -- ray/box, ray/triangle, ray/sphere tests and the BVH walk.
//
// Arithmetic follows the reference so that hit ids are bit-exact:
//   slab test            intersections.h:7-41   ((b - o) * (1.0f/d), box t_min fixed at 0.001f)
//   Moller-Trumbore      intersections.h:54-83  (reject |a| < 1e-7, u, v, u+v, t in (t_min,t_max))
//   sphere               intersections.h:85-104 (both roots)
//   dual-node traversal  kernels.cu:148-224     (near child first, tie -> left, bit-stack pop)
// Layout differences (the device layout is ours):
//   * nodes: 64 bytes per internal node index i, holding the boxes of its children L = 2i and R = 2i+1 as three float4
//         [0] {Lmin.x, Lmax.x, Rmin.x, Rmax.x}   [1] the same for y   [2] for z   [3] padding
//     so that one 32-byte load (LDG.E.ENL2.256) and one 16-byte load fetch both children from one 128-byte line
//     (2 L1 requests per step instead of 3; an earlier 96-byte record with a pre-swapped copy per direction sign cost
//     no selects but a third of the L1 lines held dead copies). The slab test's swap of t0/t1 when invD < 0
//     (intersections.h:30) is a select after the two products;
//   * triangles: packed per leaf. A leaf of N slots is N 32-byte blocks {v0.xyz, e1.xyz, e2.xy} followed by one tail block
//     with the N values e2.z (padded to a multiple of 32 bytes): 192 bytes for N = 5 instead of 5 x 48, one 256-bit load
//     plus one 32-bit load (all N of them hit one sector) per test instead of three 128-bit loads, and 9 registers per tile
//     instead of 12. e1 = v1-v0 and e2 = v2-v0 are the subtractions triangleHit does first; precomputing them does not
//     change a bit. +inf in v0.x marks an unused leaf slot.
#pragma once

#include <cfloat>

#include "vecmath.cuh"

#define RT_EPSILON 0.01f // kernels.cu:19


struct MeshView {
    const float4* __restrict__ nodes; // 4 float4 per internal node index (see travNodeStep / swizzleNodesKernel)
    const float4* __restrict__ tris;  // packed leaves, leafBytes each (see above)
    unsigned int leafBytes;           // 32 * N + 32 * ceil(N / 8)
    unsigned int firstLeaf;
    unsigned int primsPerLeaf;
    f3 boundsMin, boundsMax;
};

struct RayPrep {
    f3 o;    // origin
    f3 d;    // unit direction (ray.h:9)
    f3 inv;  // 1.0f / d  (IEEE)
};

// One axis of the slab loop (intersections.h:27-36):
//     t0 = (bmin - o) * invD; t1 = (bmax - o) * invD; if (invD < 0) swap(t0, t1);
//     t_min = t0 > t_min ? t0 : t_min;  t_max = t1 < t_max ? t1 : t_max;
// The swap is a per-ray constant (sign of invD), so the near/far plane is selected before the multiply: the two products
// are the same two products. The ternaries keep t_min / t_max when t0 / t1 is NaN (0 * inf); fmaxf / fminf return their
// non-NaN operand, i.e. the same value, in one min/max instruction instead of a compare plus a select.
__device__ __forceinline__ void slabAxis(float bmin, float bmax, float o, float inv, float& tMin, float& tMax) {
    const bool neg = inv < 0.0f;
    const float t0 = ((neg ? bmax : bmin) - o) * inv;
    const float t1 = ((neg ? bmin : bmax) - o) * inv;
    tMin = fmaxf(tMin, t0);
    tMax = fminf(tMax, t1);
}

// hit_bbox_dist: entry distance, or FLT_MAX. tMin only grows and tMax only shrinks, so testing
// once after the third axis equals the reference's per-axis early return.
__device__ __forceinline__ float boxDist(const f3& bmin, const f3& bmax, const RayPrep& r, float tMax) {
    float tMin = 0.001f;
    slabAxis(bmin.x, bmax.x, r.o.x, r.inv.x, tMin, tMax);
    slabAxis(bmin.y, bmax.y, r.o.y, r.inv.y, tMin, tMax);
    slabAxis(bmin.z, bmax.z, r.o.z, r.inv.z, tMin, tMax);
    return tMax < tMin ? FLT_MAX : tMin;
}

// hit_bbox (intersections.h:7-23): same slabs, boolean result.
__device__ __forceinline__ bool boxHit(const f3& bmin, const f3& bmax, const RayPrep& r, float tMax) {
    float tMin = 0.001f;
    slabAxis(bmin.x, bmax.x, r.o.x, r.inv.x, tMin, tMax);
    slabAxis(bmin.y, bmax.y, r.o.y, r.inv.y, tMin, tMax);
    slabAxis(bmin.z, bmax.z, r.o.z, r.inv.z, tMin, tMax);
    return !(tMax < tMin);
}

// triangleHit on a prepared tile. Returns FLT_MAX or t in (tMin, tMax). Every operation is pinned (see vecmath.cuh).
__device__ __forceinline__ float triHit(const f3& v0, const f3& edge1, const f3& edge2, const RayPrep& r, float tMin, float tMax,
                                        float& hitU, float& hitV) {
    const float EPS = 0.0000001f;
    const f3 h = cross(r.d, edge2);
    const float a = dot(edge1, h);
    if (a > -EPS && a < EPS) return FLT_MAX;
    const float f = __frcp_rn(a); // 1.0 / a, correctly rounded: the same float as the reference's division, fewer instructions
    const f3 s = subPinned(r.o, v0);
    const float u = __fmul_rn(f, dot(s, h));
    if (u < 0.0f || u > 1.0f) return FLT_MAX;
    const f3 q = cross(s, edge1);
    const float v = __fmul_rn(f, dot(r.d, q));
    if (v < 0.0f || __fadd_rn(u, v) > 1.0f) return FLT_MAX;
    const float t = __fmul_rn(f, dot(edge2, q));
    if (t > tMin && t < tMax) {
        hitU = u;
        hitV = v;
        return t;
    }
    return FLT_MAX;
}

__device__ __forceinline__ float sphereHitT(const f3& center, float radius, const f3& o, const f3& d, float tMin, float tMax) {
    const f3 oc = o - center;
    const float a = dot(d, d);
    const float b = dot(oc, d);
    const float c = __fsub_rn(dot(oc, oc), __fmul_rn(radius, radius)); // compiled as a separate multiply and subtract
    const float discriminant = __fmaf_rn(b, b, -__fmul_rn(a, c));
    if (discriminant > 0) {
        float temp = (-b - sqrtf(discriminant)) / a;
        if (temp < tMax && temp > tMin) return temp;
        temp = (-b + sqrtf(discriminant)) / a;
        if (temp < tMax && temp > tMin) return temp;
    }
    return FLT_MAX;
}
