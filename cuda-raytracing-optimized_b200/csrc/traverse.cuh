// traverse.cuh -- resumable, warp-cooperative BVH traversal in the reference's visiting order.
//
// Same decisions and arithmetic as hitBvh (kernels.cu:154-224): both children of an internal node are slab-tested
// against the current closest hit, the nearer one is entered first (tie -> left), the other is remembered as one bit
// of a 32-bit trail, and a pop jumps straight to the pending sibling (pop_bitstack, kernels.cu:148-152).  Because the
// tree is an implicit complete heap and the trail is a bit-stack, the WHOLE traversal state of a ray is
//     { idx, bitStack, closest, triId, u, v }                                  (24 bytes)
// (+ the origin and direction it was set up with), so a ray can stop after a bounded number of steps and continue in a
// later launch with no stack to spill.  The first
// ncu capture (profiles/r01/a_extend_ncu_summary.txt) showed why that matters: every launch of the one-shot walk waited
// for a single straggler ray (~1000 steps) with 5.8 of 32 lanes active.
//
// Loop shape ("while-while", Aila & Laine 2009): every lane first walks internal nodes until it stands on a leaf (or is
// done), then the leaf's <= N triangles are tested; the expensive leaf body is entered once per round instead of being
// predicated into every node step.
#pragma once

#include "intersect.cuh"

#ifndef TRACE_NODE_QUORUM
#define TRACE_NODE_QUORUM 16
#endif // node steps run while this many lanes (or half of the rays the warp holds) stand on nodes

// A ray while it is being traversed. The fields the node loop touches every step stay in registers (RayHot, TravHot);
// the rest -- direction (needed by the triangle test only), the hit record and the queue entry -- lives in shared memory,
// one RayCold per thread: the second profile of the leaner node step (profiles/r01/c_trace_ncu_summary.txt) was bound by
// L1/L2 latency at 36 % occupancy (64 registers), so registers are what buys more resident warps.
struct RayHot {
    float ox, oy, oz;                 // origin
    float ix, iy, iz;                 // 1.0f / direction (IEEE), direction normalised by the ray constructor (ray.h:9)
};

struct TravHot {
    unsigned int idx;       // current node (0 = traversal finished / lane idle)
    unsigned int bitStack;  // trail of pending siblings, sentinel bit on top
    float closest;          // current t_max
};

struct RayCold {
    float4 dir;  // {unit direction, original t_max}
    float4 rec;  // {u, v, triId bits, user word (queue entry / ray index)}
};

__device__ __forceinline__ void prepRay(RayHot& r, RayCold& c, const f3& o, const f3& dirNormalised, float tMax) {
    r.ox = o.x; r.oy = o.y; r.oz = o.z;
    r.ix = 1.0f / dirNormalised.x; r.iy = 1.0f / dirNormalised.y; r.iz = 1.0f / dirNormalised.z;
    c.dir = make_float4(dirNormalised.x, dirNormalised.y, dirNormalised.z, tMax);
}

// hit_bbox (intersections.h:7-23) against the scene bounds, hitMesh's early out (kernels.cu:297)
__device__ __forceinline__ bool rayHitsBounds(const MeshView& m, const RayHot& r, float tMax) {
    RayPrep p;
    p.o = mk3(r.ox, r.oy, r.oz);
    p.inv = mk3(r.ix, r.iy, r.iz);
    return boxHit(m.boundsMin, m.boundsMax, p, tMax);
}

__device__ __forceinline__ void travPop(TravHot& s) {
    const int m = __ffs(s.bitStack) - 1;
    s.bitStack = (s.bitStack >> m) ^ 1u;
    s.idx = (s.idx >> m) ^ 1u;
}

__device__ __forceinline__ void prefetchL1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// 32-byte and 16-byte loads through the read-only path (sm_100 has 256-bit global loads: LDG.E.ENL2.256)
__device__ __forceinline__ void ldg256(const void* p, float4& a, float4& b) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
        : "l"(p));
}

// One internal-node step (kernels.cu:163-197): both children's slab tests (hit_bbox_dist, intersections.h:25-41), then
// descend to the nearer child / remember the other / pop.
//
// Node record (intersect.cuh): 64 bytes per internal node index, {Lmin, Lmax, Rmin, Rmax} per axis, fetched with one
// 32-byte and one 16-byte load (2 L1 requests instead of 3: the kernel is bound by L1 wavefronts). The slab test's
// `if (invD < 0) swap(t0, t1)` becomes a select of {near, far} after the two products. hit_bbox_dist returns FLT_MAX on a
// miss and its t_min otherwise, and the caller compares that with `closest`; with tMax = min(closest, far...) folded in:
//     traverse  <=>  !(tMax < tMin) && tMin < closest  <=>  tMin <= tMax && tMin < closest
//     swap (rightHit < leftHit) matters only when a child is entered: both -> tMinR < tMinL; only right -> 1; only left -> 0.
// fmaxf / fminf drop a NaN operand (0 * inf on an axis the ray is parallel to) exactly like the reference's
// `t0 > t_min ? t0 : t_min`, and their result does not depend on the order of the operands.
//
// PREFETCH: whichever child is entered next, its record lies in the 128 bytes that start at byte 128*idx (children 2*idx
// and 2*idx+1 are adjacent), or -- on the last internal level -- its triangles lie in the 2*N tiles that start at leaf
// 2*idx - firstLeaf. Requesting those lines into L1 now overlaps the next step's memory latency with this step's slab
// tests. Used when a launch has too few rays to hide latency with other warps (the tail of a frame).
template <bool PREFETCH>
__device__ __forceinline__ void travNodeStep(const MeshView& m, const RayHot& r, TravHot& s) {
    const char* rec = (const char*)m.nodes + 64u * s.idx;
    float4 qx, qy;
    ldg256(rec, qx, qy);
    const float4 qz = __ldg((const float4*)(rec + 32));
    if (PREFETCH) {
        const unsigned int child = 2u * s.idx;
        if (child < m.firstLeaf) {
            prefetchL1((const char*)m.nodes + 64u * child);
        } else {
            const char* p = (const char*)m.tris + (size_t)(child - m.firstLeaf) * m.leafBytes;
            const unsigned int bytes = 2u * m.leafBytes;
            for (unsigned int o = 0; o < bytes; o += 128u) prefetchL1((const char*)p + o);
            prefetchL1((const char*)p + bytes - 16u);
        }
    }
    const float2 xl = slabPair(qx.x, qx.y, -r.ox, r.ix), xr = slabPair(qx.z, qx.w, -r.ox, r.ix);
    const float2 yl = slabPair(qy.x, qy.y, -r.oy, r.iy), yr = slabPair(qy.z, qy.w, -r.oy, r.iy);
    const float2 zl = slabPair(qz.x, qz.y, -r.oz, r.iz), zr = slabPair(qz.z, qz.w, -r.oz, r.iz);
    const bool nx = r.ix < 0.0f, ny = r.iy < 0.0f, nz = r.iz < 0.0f; // `if (invD < 0) swap(t0, t1)` (intersections.h:30)
    const float tMinL = fmaxf(fmaxf(nx ? xl.y : xl.x, ny ? yl.y : yl.x), fmaxf(nz ? zl.y : zl.x, 0.001f));
    const float tMinR = fmaxf(fmaxf(nx ? xr.y : xr.x, ny ? yr.y : yr.x), fmaxf(nz ? zr.y : zr.x, 0.001f));
    const float tMaxL = fminf(fminf(nx ? xl.x : xl.y, ny ? yl.x : yl.y), fminf(nz ? zl.x : zl.y, s.closest));
    const float tMaxR = fminf(fminf(nx ? xr.x : xr.y, ny ? yr.x : yr.y), fminf(nz ? zr.x : zr.y, s.closest));
    const bool traverseLeft = tMinL <= tMaxL && tMinL < s.closest;
    const bool traverseRight = tMinR <= tMaxR && tMinR < s.closest;
    if (traverseLeft || traverseRight) {
        const bool swap = traverseRight && (!traverseLeft || tMinR < tMinL);
        s.idx = 2u * s.idx + (swap ? 1u : 0u);
        s.bitStack = (s.bitStack << 1) + ((traverseLeft && traverseRight) ? 1u : 0u);
    } else {
        travPop(s);
    }
}

// One leaf visit (kernels.cu:198-217). An any-hit ray that finds a triangle is finished: closest = 0.0f is what the
// reference returns (kernels.cu:207).
__device__ __forceinline__ void travLeafStep(const MeshView& m, const RayHot& r, RayCold& c, float tMin, bool anyHit, TravHot& s,
                                             unsigned int& triTests) {
    const unsigned int first = (s.idx - m.firstLeaf) * m.primsPerLeaf;
    const char* leaf = (const char*)m.tris + (size_t)(s.idx - m.firstLeaf) * m.leafBytes;
    {   // the leaf is contiguous (leafBytes): request all of its lines before the first test
        for (unsigned int o = 128u; o < m.leafBytes; o += 128u) prefetchL1(leaf + o);
        prefetchL1(leaf + m.leafBytes - 16u);
    }
    RayPrep rp;
    rp.o = mk3(r.ox, r.oy, r.oz);
    rp.d = xyz(c.dir);
    for (unsigned int i = 0; i < m.primsPerLeaf; i++) {
        // 8 of the tile's 9 floats with one 256-bit load, e2.z from the leaf's tail block (one sector for the whole leaf)
        float4 t0, t1, t2;
        ldg256(leaf + 32u * i, t0, t1);
        t2.x = __ldg((const float*)(leaf + 32u * m.primsPerLeaf) + i);
        if (isinf(t0.x)) break;
        triTests++;
        float u, v;
        const float hitT = triHit(mk3(t0.x, t0.y, t0.z), mk3(t0.w, t1.x, t1.y), mk3(t1.z, t1.w, t2.x), rp, tMin, s.closest, u, v);
        if (hitT < s.closest) {
            if (anyHit) {
                s.closest = 0.0f;
                s.idx = 0;
                return;
            }
            s.closest = hitT;
            c.rec.x = u;
            c.rec.y = v;
            c.rec.z = __uint_as_float(first + i);
        }
    }
    travPop(s);
}

// One scheduling round of a warp (all 32 lanes call it together; it does not change any lane's own sequence of steps):
//   * node steps are issued while at least `quorum` lanes stand on internal nodes; lanes that already reached a leaf wait
//     for that long and no longer -- waiting for ALL lanes to reach a leaf (plain while-while, Aila & Laine 2009) left 8
//     of 32 lanes active in the node loop (profiles/r01/b_trace_ncu_summary.txt); oracle/sched_sim.cpp replays logged
//     traversals through this policy and its alternatives;
//   * then every lane that stands on a leaf tests its triangles.
// `steps` counts node steps + 2 per leaf visit, for the caller's parking budget; lanes with `on` == false sit the round out.
template <bool PREFETCH>
__device__ __forceinline__ void travRound(const MeshView& m, const RayHot& r, RayCold& c, float tMin, bool anyHit, bool on, TravHot& s,
                                          int& steps, int quorum, unsigned int& nodeVisits, unsigned int& triTests) {
    while (true) {
        const bool atNode = on && (s.idx - 1u) < (m.firstLeaf - 1u); // idx != 0 && idx < firstLeaf
        if (__popc(__ballot_sync(0xFFFFFFFFu, atNode)) < quorum) break;
        if (atNode) {
            travNodeStep<PREFETCH>(m, r, s);
            nodeVisits++;
            steps++;
        }
    }
    if (on && s.idx >= m.firstLeaf) {
        travLeafStep(m, r, c, tMin, anyHit, s, triTests);
        steps += 2;
    }
}
