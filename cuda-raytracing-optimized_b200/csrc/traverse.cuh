// traverse.cuh -- resumable, warp-cooperative BVH traversal in the reference's visiting order.
//
// Same decisions and arithmetic as hitBvh (kernels.cu:154-224): both children of an internal node are slab-tested
// against the current closest hit, the nearer one is entered first (tie -> left), the other is remembered as one bit
// of a 32-bit trail, and a pop jumps straight to the pending sibling (pop_bitstack, kernels.cu:148-152).  Because the
// tree is an implicit complete heap and the trail is a bit-stack, the WHOLE traversal state of a ray is
//     { idx, bitStack, closest, triId, u, v }                                  (24 bytes)
// so a ray can stop after a bounded number of steps and continue in a later launch with no stack to spill.  The first
// ncu capture (profiles/r01/a_extend_ncu_summary.txt) showed why that matters: every launch of the one-shot walk waited
// for a single straggler ray (~1000 steps) with 5.8 of 32 lanes active.
//
// Loop shape ("while-while", Aila & Laine 2009): every lane first walks internal nodes until it stands on a leaf (or is
// done), then the leaf's <= N triangles are tested; the expensive leaf body is entered once per round instead of being
// predicated into every node step.
#pragma once

#include "intersect.cuh"

struct TravState {
    unsigned int idx;       // current node (0 = traversal finished)
    unsigned int bitStack;  // trail of pending siblings, sentinel bit on top
    float closest;          // current t_max
    unsigned int triId;
    float u, v;
};

__device__ __forceinline__ void travInit(TravState& s, float tMax) {
    s.idx = 1;
    s.bitStack = 1;
    s.closest = tMax;
    s.triId = 0xFFFFFFFFu;
    s.u = 0.0f;
    s.v = 0.0f;
}

__device__ __forceinline__ void travPop(TravState& s) {
    const int m = __ffs(s.bitStack) - 1;
    s.bitStack = (s.bitStack >> m) ^ 1u;
    s.idx = (s.idx >> m) ^ 1u;
}

__device__ __forceinline__ void prefetchL1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// One internal-node step (kernels.cu:163-197).
// PREFETCH: whichever child is entered next, its record lies in the 96 bytes that start at float4 index 6*idx (children
// 2*idx and 2*idx+1 are adjacent), or -- on the last internal level -- its triangles lie in the 2*N tiles that start at
// leaf 2*idx - firstLeaf. Requesting those lines into L1 now overlaps the next step's memory latency with this step's
// slab tests. Used when a launch has too few rays to hide latency with other warps (the tail of a frame).
template <bool PREFETCH>
__device__ __forceinline__ void travNodeStep(const MeshView& m, const RayPrep& r, TravState& s) {
    const float4 a = __ldg(m.nodes + 3 * s.idx);
    const float4 b = __ldg(m.nodes + 3 * s.idx + 1);
    const float4 c = __ldg(m.nodes + 3 * s.idx + 2);
    if (PREFETCH) {
        const unsigned int child = 2u * s.idx;
        if (child < m.firstLeaf) {
            const float4* p = m.nodes + 3 * child;
            prefetchL1(p);
            prefetchL1(p + 5);
        } else {
            const float4* p = m.tris + 3 * ((child - m.firstLeaf) * m.primsPerLeaf);
            const unsigned int bytes = 2u * m.primsPerLeaf * 48u;
            for (unsigned int o = 0; o < bytes; o += 128u) prefetchL1((const char*)p + o);
            prefetchL1((const char*)p + bytes - 16u);
        }
    }
    const float leftHit = boxDist(mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), r, s.closest);
    const float rightHit = boxDist(mk3(b.z, b.w, c.x), mk3(c.y, c.z, c.w), r, s.closest);
    const bool traverseLeft = leftHit < s.closest;
    const bool traverseRight = rightHit < s.closest;
    const unsigned int swap = rightHit < leftHit ? 1u : 0u;
    if (traverseLeft || traverseRight) {
        s.idx = 2 * s.idx + swap;
        s.bitStack = (s.bitStack << 1) + ((traverseLeft && traverseRight) ? 1u : 0u);
    } else {
        travPop(s);
    }
}

// One leaf visit (kernels.cu:198-217). Returns true when an any-hit ray is finished (occluded).
__device__ __forceinline__ bool travLeafStep(const MeshView& m, const RayPrep& r, float tMin, bool anyHit, TravState& s,
                                             unsigned int& triTests) {
    const unsigned int first = (s.idx - m.firstLeaf) * m.primsPerLeaf;
    {   // the leaf's tiles are contiguous (N * 48 bytes): request all of its lines before the first test
        const char* p = (const char*)(m.tris + 3 * first);
        const unsigned int bytes = m.primsPerLeaf * 48u;
        for (unsigned int o = 128u; o < bytes; o += 128u) prefetchL1(p + o);
        prefetchL1(p + bytes - 16u);
    }
    for (unsigned int i = 0; i < m.primsPerLeaf; i++) {
        // all 48 bytes of the tile are requested together (one round trip); unused slots are readable padding
        const float4 t0 = __ldg(m.tris + 3 * (first + i));
        const float4 t1 = __ldg(m.tris + 3 * (first + i) + 1);
        const float4 t2 = __ldg(m.tris + 3 * (first + i) + 2);
        if (isinf(t0.x)) break;
        triTests++;
        float u, v;
        const float hitT = triHit(mk3(t0.x, t0.y, t0.z), mk3(t0.w, t1.x, t1.y), mk3(t1.z, t1.w, t2.x), r, tMin, s.closest, u, v);
        if (hitT < s.closest) {
            if (anyHit) {
                s.closest = 0.0f; // the reference returns 0.0f here (kernels.cu:207)
                s.idx = 0;
                return true;
            }
            s.closest = hitT;
            s.triId = first + i;
            s.u = u;
            s.v = v;
        }
    }
    travPop(s);
    return false;
}

// Run one lane's traversal for at most `budget` steps (a node step costs 1, a leaf visit 2); `steps` accumulates.
// Lanes of a warp call this together. Scheduling inside the warp (it does not change any lane's own sequence of steps):
//   * node steps are issued while at least TRAV_NODE_QUORUM lanes stand on internal nodes; lanes that already reached a
//     leaf wait for that long and no longer -- waiting for ALL lanes to reach a leaf (plain while-while) left 8 of 32
//     lanes active in the node loop (profiles/r01/b_trace_ncu_summary.txt), a quorum of 16 doubles that in simulation;
//   * then the leaves are processed for every lane that stands on one.
// Returns when fewer than `minActive` lanes of the warp still have work (finished, out of budget or idle): the caller
// retires / refills lanes and calls again.
#define TRAV_NODE_QUORUM 16

template <bool PREFETCH>
__device__ __forceinline__ void travRun(const MeshView& m, const RayPrep& r, float tMin, bool anyHit, bool live, TravState& s,
                                        int& steps, int budget, int minActive, unsigned int& nodeVisits, unsigned int& triTests) {
    while (true) {
        bool work = live && s.idx != 0u && steps < budget;
        if (__popc(__ballot_sync(0xFFFFFFFFu, work)) < minActive) break;
        // node phase
        while (true) {
            const bool atNode = work && s.idx < m.firstLeaf;
            const unsigned int nodeMask = __ballot_sync(0xFFFFFFFFu, atNode);
            if (nodeMask == 0u) break;
            const unsigned int leafMask = __ballot_sync(0xFFFFFFFFu, work && s.idx >= m.firstLeaf);
            if (__popc(nodeMask) < TRAV_NODE_QUORUM && leafMask != 0u) break;        // let the waiting lanes test their leaves
            if (__popc(nodeMask) + __popc(leafMask) < minActive) break;              // too few lanes left: let the caller refill
            if (atNode) {
                travNodeStep<PREFETCH>(m, r, s);
                nodeVisits++;
                steps++;
                work = s.idx != 0u && steps < budget;
            }
        }
        // leaf phase
        if (work && s.idx >= m.firstLeaf) {
            travLeafStep(m, r, tMin, anyHit, s, triTests);
            steps += 2;
        }
    }
}
