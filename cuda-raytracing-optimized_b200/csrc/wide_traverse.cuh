// wide_traverse.cuh -- traversal of the renderer's own 8-wide quantised BVH (wide_bvh.h) plus the certificate that keeps the
// result THE REFERENCE'S result.
//
// What the reference computes (hitBvh, kernels.cu:154-224) is "the smallest triangleHit() over the triangles its walk
// reaches"; its tree decides the answer only through (i) which of two exactly tied triangles is met first and (ii) box
// tests that cull, by a last-ulp rounding, a leaf whose triangle test would have passed. The fast walk below evaluates the
// SAME triangle test (triHit, intersect.cuh) over a conservative tree, i.e. it finds the true minimum over all triangles,
// and then proves that the reference finds the same triangle:
//   * near-tie flag: box culling and triangle acceptance use closest * (1 + 1e-5), so every triangle within that margin of
//     the winner is seen; if there is one, the ray is flagged;
//   * leaf certificate: the reference tests the winner's leaf iff the leaf's box (and every ancestor's) passes
//     hit_bbox_dist(box, ray, closest) < closest with ITS `closest` at that moment, which is >= min(t_max, 1.00001 * t_win)
//     when no near-tie exists. Parent boxes contain their children's and the slab arithmetic (intersections.h:25-41) is
//     monotonic in the box coordinates, so the leaf's box decides for all its ancestors: ONE reference-arithmetic box test
//     per found hit. If it fails, the ray is flagged.
//   * any-hit rays: `closest` stays t_max until the first hit, so the same test with t_max certifies "occluded".
//   * a miss needs no certificate: the reference's answer is always one of the triangle tests that pass.
// Flagged rays (measured: ~3 in 100 000) are re-traced by the order-exact kernel (traverse.cuh), which is also the checker
// in the tests. The containment assumption about the caller's tree is verified on the device at initRenderer
// (checkRefTreeKernel); if it does not hold, every ray takes the exact kernel.
#pragma once

#include "traverse.cuh"
#include "wide_bvh.h"

#define WIDE_TIE_MARGIN 1.00001f
#ifndef WIDE_NODE_QUORUM
#define WIDE_NODE_QUORUM 12 // node steps run while this many lanes (or half of the rays the warp holds) stand on nodes (measured: 8 / 12 / 16 / 20)
#endif
#define WIDE_FLAG_TIE 0x100u     // WideRay::oct: a second triangle within the margin of the current best was seen
#define WIDE_FLAG_ANYHIT 0x200u
#define WIDE_FLAG_EXACT 0x400u   // the record was produced by the order-exact walk: nothing to certify

struct WideView {
    const uint4* __restrict__ nodes;   // 96-byte records (wide_bvh.h)
    const float4* __restrict__ triA;   // per leaf triangle {v0.xyz, e1.x}{e1.yz, e2.xy}
    const float2* __restrict__ triB;   //                   {e2.z, caller's slot index}
    float rangeX, rangeY, rangeZ;      // |origin| limit of the fast path (WIDE_ORIGIN_RANGE x coordinate range)
    unsigned int stackDepth;           // entries per thread (>= wide tree depth); 0 = no wide tree: exact kernel only
    unsigned int topCount;             // (WIDE_SMEM_TOP experiment: nodes the launching kernel staged in shared memory; else 0)
};

struct WideRay {
    float ox, oy, oz;
    float ix, iy, iz;    // 1 / direction with |direction| clamped to >= 1e-20 (no inf * 0 in t = m * A + B)
    unsigned int oct;    // bits 0..2: 4 = d.x >= 0, 2 = d.y >= 0, 1 = d.z >= 0 ("octinv"); WIDE_FLAG_*
};

struct WideTrav {
    unsigned int ngx, ngy;  // node group: child base | hits in priority order (bits 24..31) + imask (bits 0..7)
    unsigned int tgx, tgy;  // triangle group: the node's triangle base | mask of the hit leaf slots still to test (8 bits)
    int sp;                 // entries on the stack; -1 = traversal finished
    float closest;
};

__device__ __forceinline__ void ldg256u(const void* p, uint4& a, uint4& b) {
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
        : "l"(p));
}

// Plane decode, two children per PRMT. The bytes (0x80 | q) of children 2k and 2k+1 and two constant 0x3F bytes form the
// word [3F][Q 2k][3F][Q 2k+1]: read as a float it is 1 + q/128 for child 2k (plus the low half as 0.248..0.25 of a grid
// step of excess mantissa, which the builder has already accounted for when it rounded child 2k's planes outward:
// wide_bvh.cpp), shifted left by 16 it is exactly 1 + q/128 for child 2k+1. The 0x3F bytes come from a register the compiler
// cannot see through, so that the selector stays an immediate (otherwise ptxas keeps re-materialising selector registers).
__device__ __forceinline__ unsigned int wideConst3F() {
    unsigned int k;
    asm volatile("mov.u32 %0, 0x3F3F3F3F;" : "=r"(k));
    return k;
}
template <int PAIR>
__device__ __forceinline__ unsigned int planePair(unsigned int w, unsigned int k3f) {
    unsigned int d;
    if (PAIR == 0) asm("prmt.b32 %0, %1, %2, 0x4041;" : "=r"(d) : "r"(w), "r"(k3f));
    else asm("prmt.b32 %0, %1, %2, 0x4243;" : "=r"(d) : "r"(w), "r"(k3f));
    return d;
}
// {t(child 2k), t(child 2k+1)} = {m, m'} * a + b in one packed FMA (sm_100 FFMA2)
__device__ __forceinline__ float2 slabPairFma(unsigned int pair, float a, float b) {
    float2 r;
    const unsigned int second = pair << 16;
    asm("{\n\t"
        ".reg .b64 p, q, s;\n\t"
        "mov.b64 p, {%2, %3};\n\t"
        "mov.b64 q, {%4, %4};\n\t"
        "mov.b64 s, {%5, %5};\n\t"
        "fma.rn.f32x2 p, p, q, s;\n\t"
        "mov.b64 {%0, %1}, p;\n\t"
        "}"
        : "=f"(r.x), "=f"(r.y)
        : "r"(pair), "r"(second), "f"(a), "f"(b));
    return r;
}

// Sets the ray up. `inv` receives the reference's own reciprocal direction (IEEE, may be infinite): the scene-bounds test at
// the start and the certificate at the end use it, so the three divisions are done once per ray. Returns false when the
// fast path does not take the ray (origin outside the range the padding covers, or a non-finite direction): the caller
// hands it to the exact kernel.
__device__ __forceinline__ bool wideSetup(const WideView& w, WideRay& r, const f3& o, const f3& d, bool anyHit, f3& inv) {
    r.ox = o.x; r.oy = o.y; r.oz = o.z;
    inv = mk3(__frcp_rn(d.x), __frcp_rn(d.y), __frcp_rn(d.z)); // = 1.0f / d, correctly rounded (the reference's invD, intersections.h:27)
    const float big = 1.0f / 1e-20f; // |direction| is clamped to >= 1e-20
    r.ix = fabsf(d.x) < 1e-20f ? copysignf(big, d.x) : inv.x;
    r.iy = fabsf(d.y) < 1e-20f ? copysignf(big, d.y) : inv.y;
    r.iz = fabsf(d.z) < 1e-20f ? copysignf(big, d.z) : inv.z;
    // near / far by the SIGN BIT of the component (-0.0 clamps to -1e-20: its reciprocal is negative)
    r.oct = (signbit(d.x) ? 0u : 4u) | (signbit(d.y) ? 0u : 2u) | (signbit(d.z) ? 0u : 1u) | (anyHit ? WIDE_FLAG_ANYHIT : 0u);
    const bool finite = (fabsf(d.x) <= 1.0f) && (fabsf(d.y) <= 1.0f) && (fabsf(d.z) <= 1.0f); // false for NaN
    return finite && fabsf(o.x) <= w.rangeX && fabsf(o.y) <= w.rangeY && fabsf(o.z) <= w.rangeZ;
}

// hit_bbox (intersections.h:7-23) against the scene bounds, hitMesh's early out (kernels.cu:297)
__device__ __forceinline__ bool wideHitsBounds(const MeshView& m, const WideRay& r, const f3& inv, float tMax) {
    RayPrep p;
    p.o = mk3(r.ox, r.oy, r.oz);
    p.inv = inv;
    return boxHit(m.boundsMin, m.boundsMax, p, tMax);
}

__device__ __forceinline__ void wideStart(WideTrav& s, float tMax) {
    s.ngx = 0u;
    s.ngy = 0x01000000u; // a virtual group whose only hit child is node 0 (imask 0: relative index 0)
    s.tgx = 0u; s.tgy = 0u;
    s.sp = 0;
    s.closest = tMax;
}

__device__ __forceinline__ void widePop(WideTrav& s, const uint2* stack, unsigned int stride) {
    if (s.sp > 0) {
        s.sp--;
        const uint2 g = stack[(unsigned int)s.sp * stride];
        s.ngx = g.x; s.ngy = g.y;
    } else {
        s.sp = -1;
    }
}

// One wide-node step: take the nearest pending child of the node group, test its 8 children, form the new groups.
// Instruction budget (the kernel is bound by instruction issue, mostly the ALU pipe: profiles/r02): per node 24 PRMT,
// 24 shifts, 24 packed FMAs, 32 min/max, 8 subtractions whose sign bits are funnelled into the miss mask.
__device__ __forceinline__ void wideNodeStep(const WideView& w, const WideRay& r, WideTrav& s, uint2* stack, unsigned int stride, unsigned int k3f) {
    const unsigned int bit = 31u - (unsigned int)__clz(s.ngy);
    s.ngy &= ~(1u << bit);
    if (s.ngy > 0x00FFFFFFu) {
        stack[(unsigned int)s.sp * stride] = make_uint2(s.ngx, s.ngy);
        s.sp++;
    }
    const unsigned int slot = (bit - 24u) ^ (r.oct & 7u);
    const unsigned int idx = s.ngx + __popc(s.ngy & ((1u << slot) - 1u) & 0xFFu);
    const char* rec = (const char*)w.nodes + 96ull * idx;
    uint4 h0, h1, q0, q1, q2, q3;
#ifdef WIDE_SMEM_TOP // (experiment, profiles/r02/neg_*: the first WIDE_SMEM_TOP nodes staged in shared memory by the kernel)
    extern __shared__ uint2 wideStackAll[];
    const uint4* top = (const uint4*)(wideStackAll + w.stackDepth * blockDim.x);
    if (idx < w.topCount) {
        h0 = top[6u * idx]; h1 = top[6u * idx + 1u]; q0 = top[6u * idx + 2u]; q1 = top[6u * idx + 3u]; q2 = top[6u * idx + 4u]; q3 = top[6u * idx + 5u];
    } else
#endif
    {
        ldg256u(rec, h0, h1);        // p.xyz, e|imask, childBase, triBase, meta lo, meta hi
        ldg256u(rec + 32, q0, q1);   // qlo x, x', y, y' | z, z', qhi x, x'
        ldg256u(rec + 64, q2, q3);   // qhi y, y', z, z' | scale xyz, spare
    }
    const float ax = __uint_as_float(q3.x) * r.ix, ay = __uint_as_float(q3.y) * r.iy, az = __uint_as_float(q3.z) * r.iz;
    const float bx = __fmaf_rn(__uint_as_float(h0.x) - r.ox, r.ix, -ax);
    const float by = __fmaf_rn(__uint_as_float(h0.y) - r.oy, r.iy, -ay);
    const float bz = __fmaf_rn(__uint_as_float(h0.z) - r.oz, r.iz, -az);
    const float limit = (r.oct & WIDE_FLAG_ANYHIT) ? s.closest : s.closest * WIDE_TIE_MARGIN;
    const bool px = (r.oct & 4u) != 0u, py = (r.oct & 2u) != 0u, pz = (r.oct & 1u) != 0u;
    unsigned int miss = 0u;
#define WIDE_HALF(NX, FX, NY, FY, NZ, FZ)                                                                               \
    {                                                                                                                    \
        const float2 nxa = slabPairFma(planePair<0>(NX, k3f), ax, bx), nxb = slabPairFma(planePair<1>(NX, k3f), ax, bx); \
        const float2 fxa = slabPairFma(planePair<0>(FX, k3f), ax, bx), fxb = slabPairFma(planePair<1>(FX, k3f), ax, bx); \
        const float2 nya = slabPairFma(planePair<0>(NY, k3f), ay, by), nyb = slabPairFma(planePair<1>(NY, k3f), ay, by); \
        const float2 fya = slabPairFma(planePair<0>(FY, k3f), ay, by), fyb = slabPairFma(planePair<1>(FY, k3f), ay, by); \
        const float2 nza = slabPairFma(planePair<0>(NZ, k3f), az, bz), nzb = slabPairFma(planePair<1>(NZ, k3f), az, bz); \
        const float2 fza = slabPairFma(planePair<0>(FZ, k3f), az, bz), fzb = slabPairFma(planePair<1>(FZ, k3f), az, bz); \
        WIDE_CHILD(nxb.y, nyb.y, nzb.y, fxb.y, fyb.y, fzb.y)                                                             \
        WIDE_CHILD(nxb.x, nyb.x, nzb.x, fxb.x, fyb.x, fzb.x)                                                             \
        WIDE_CHILD(nxa.y, nya.y, nza.y, fxa.y, fya.y, fza.y)                                                             \
        WIDE_CHILD(nxa.x, nya.x, nza.x, fxa.x, fya.x, fza.x)                                                             \
    }
    // hit <=> max(near x, y, z, 0) <= min(far x, y, z, limit); the sign bit of (far - near) is shifted into the miss mask
#define WIDE_CHILD(TNX, TNY, TNZ, TFX, TFY, TFZ)                                                  \
    {                                                                                             \
        const float tn = fmaxf(fmaxf(TNX, TNY), fmaxf(TNZ, 0.0f));                                \
        const float tf = fminf(fminf(TFX, TFY), fminf(TFZ, limit));                               \
        miss = __funnelshift_l(__float_as_uint(__fsub_rn(tf, tn)), miss, 1);                      \
    }
    // children 7..4 first: the mask is shifted left once per child, so child j ends in bit j
    WIDE_HALF(px ? q0.y : q1.w, px ? q1.w : q0.y, py ? q0.w : q2.y, py ? q2.y : q0.w, pz ? q1.y : q2.w, pz ? q2.w : q1.y)
    WIDE_HALF(px ? q0.x : q1.z, px ? q1.z : q0.x, py ? q0.z : q2.x, py ? q2.x : q0.z, pz ? q1.x : q2.z, pz ? q2.z : q1.x)
#undef WIDE_CHILD
#undef WIDE_HALF
    const unsigned int hits = ~miss & 0xFFu;
    const unsigned int imask = h0.w >> 24;
    // inner hits -> priority order: bit (slot ^ octinv), three conditional swaps of an 8-bit mask
    unsigned int prio = hits & imask;
    if (r.oct & 1u) prio = ((prio & 0x55u) << 1) | ((prio >> 1) & 0x55u);
    if (r.oct & 2u) prio = ((prio & 0x33u) << 2) | ((prio >> 2) & 0x33u);
    if (r.oct & 4u) prio = ((prio & 0x0Fu) << 4) | ((prio >> 4) & 0x0Fu);
    s.ngx = h1.x;
    s.ngy = (prio << 24) | imask;
#ifdef WIDE_PREFETCH // (experiment, profiles/r02/neg_*: request the node that will be entered next into L1 now)
    if (prio != 0u) {
        const unsigned int nb = 31u - (unsigned int)__clz(prio << 24);
        const unsigned int ns = (nb - 24u) ^ (r.oct & 7u);
        const char* nrec = (const char*)w.nodes + 96ull * (h1.x + __popc(imask & ((1u << ns) - 1u)));
        prefetchL1(nrec);
        prefetchL1(nrec + 64);
    }
#endif
    // leaf hits: the slots go to the triangle group; their triangle ranges (meta = count << 5 | offset) are decoded in the
    // triangle phase, from the node's meta bytes parked in the thread's LAST stack entry (which the tree never reaches:
    // stackDepth > tree depth). Decoding here would make the lanes that continue with nodes wait for it.
    const unsigned int leafHits = hits & ~imask;
    s.tgx = h1.y;
    s.tgy = leafHits;
    if (leafHits != 0u) {
        stack[(w.stackDepth - 1u) * stride] = make_uint2(h1.z, h1.w);
#ifdef WIDE_TRI_PREFETCH // (experiment, profiles/r02/ab_summary.json: request the first triangle of the slot the triangle phase takes first: +2.6 %)
        const unsigned int sl = (unsigned int)__ffs((int)leafHits) - 1u;
        const unsigned int m = ((sl < 4u ? h1.z : h1.w) >> (8u * (sl & 3u))) & 31u;
        prefetchL1(w.triA + 2ull * (h1.y + m));
        prefetchL1(w.triB + (h1.y + m));
#endif
    } else if (s.ngy <= 0x00FFFFFFu) widePop(s, stack, stride);
}

// The triangles of the lane's triangle group: every hit leaf slot of the last node, 1..3 triangles each (kernels.cu:200-216 with
// the near-tie margin). c.rec = {u, v, winner slot, user}.
__device__ __forceinline__ void wideTriPhase(const WideView& w, WideRay& r, RayCold& c, float tMin, WideTrav& s, const uint2* stack,
                                             unsigned int stride, unsigned int& triTests) {
    RayPrep rp;
    rp.o = mk3(r.ox, r.oy, r.oz);
    rp.d = xyz(c.dir);
    const bool anyHit = (r.oct & WIDE_FLAG_ANYHIT) != 0u;
    const uint2 meta = stack[(w.stackDepth - 1u) * stride];
    // One hit leaf slot (1..3 triangles) per phase; further hit slots of the node wait for the warp's next triangle phase, where
    // they meet other lanes' first slots instead of running at 6-8 lanes now (measured: -1.7 %; one TRIANGLE per phase: +7 %).
    unsigned int slots = s.tgy & (0u - s.tgy);
    s.tgy &= s.tgy - 1u;
    while (slots) {
        const unsigned int sl = (unsigned int)__ffs((int)slots) - 1u;
        slots &= slots - 1u;
        const unsigned int m = ((sl < 4u ? meta.x : meta.y) >> (8u * (sl & 3u))) & 0xFFu;
        unsigned int k = s.tgx + (m & 31u), left = m >> 5;
        do {
            float4 t0, t1;
            ldg256(w.triA + 2ull * k, t0, t1);
            const float2 t2 = __ldg(w.triB + k);
            triTests++;
            const float limit = anyHit ? s.closest : s.closest * WIDE_TIE_MARGIN;
            float u, v;
            const float hitT = triHit(mk3(t0.x, t0.y, t0.z), mk3(t0.w, t1.x, t1.y), mk3(t1.z, t1.w, t2.x), rp, tMin, limit, u, v);
            if (hitT < limit) {
                if (anyHit) { // kernels.cu:207: the first hit ends an any-hit walk
                    s.closest = hitT;
                    c.rec.z = t2.y;
                    s.tgy = 0u;
                    s.sp = -1;
                    return;
                }
                if (hitT < s.closest) {
                    // new best; the previous one (and whatever was accepted before it) may lie within the margin of the new one
                    r.oct = (s.closest <= hitT * WIDE_TIE_MARGIN) ? (r.oct | WIDE_FLAG_TIE) : (r.oct & ~WIDE_FLAG_TIE);
                    s.closest = hitT;
                    c.rec.x = u;
                    c.rec.y = v;
                    c.rec.z = t2.y;
                } else if (__float_as_uint(t2.y) != __float_as_uint(c.rec.z)) {
                    r.oct |= WIDE_FLAG_TIE;
                }
            }
            k++;
        } while (--left != 0u);
    }
    if (s.tgy == 0u && s.ngy <= 0x00FFFFFFu) widePop(s, stack, stride);
}

// One scheduling round of a warp (all 32 lanes call it together), the policy of travRound (traverse.cuh): node steps are
// issued while at least `quorum` lanes stand on nodes, then every lane that holds triangles tests them.
// Invariant between calls: a lane that is not finished (sp >= 0) has triangles (tgy != 0) or a pending node (ngy hits).
__device__ __forceinline__ void wideRound(const WideView& w, WideRay& r, RayCold& c, float tMin, bool on, WideTrav& s, uint2* stack,
                                          unsigned int stride, int quorum, unsigned int k3f, unsigned int& nodeVisits, unsigned int& triTests) {
    while (true) {
        const bool atNode = on && s.sp >= 0 && s.tgy == 0u;
        if (__popc(__ballot_sync(0xFFFFFFFFu, atNode)) < quorum) break;
        if (atNode) {
            wideNodeStep(w, r, s, stack, stride, k3f);
            nodeVisits++;
        }
    }
    if (on && s.sp >= 0 && s.tgy != 0u) wideTriPhase(w, r, c, tMin, s, stack, stride, triTests);
}

// The certificate (header comment). `winner` = caller's slot index of the triangle the walk found, `inv` the reference's
// reciprocal direction (wideSetup), tMax the ray's own limit, `closest` the winner's t (closest-hit rays). True: the
// reference returns the same hit.
__device__ __forceinline__ bool wideCertify(const MeshView& m, const WideRay& r, const f3& inv, float tMax, float closest, unsigned int winner) {
    if (r.oct & WIDE_FLAG_TIE) return false;
    const unsigned int leaf = m.firstLeaf + winner / m.primsPerLeaf;
    const float4* rec = m.nodes + 4ull * (leaf >> 1); // {Lmin, Lmax, Rmin, Rmax} per axis for the children of leaf >> 1
    const float4 qx = __ldg(rec), qy = __ldg(rec + 1), qz = __ldg(rec + 2);
    const bool right = (leaf & 1u) != 0u;
    RayPrep p;
    p.o = mk3(r.ox, r.oy, r.oz);
    p.inv = inv;
    const float bound = (r.oct & WIDE_FLAG_ANYHIT) ? tMax : fminf(tMax, closest * WIDE_TIE_MARGIN);
    const float entry = boxDist(mk3(right ? qx.z : qx.x, right ? qy.z : qy.x, right ? qz.z : qz.x),
                                mk3(right ? qx.w : qx.y, right ? qy.w : qy.y, right ? qz.w : qz.y), p, bound);
    return entry < bound;
}

// ---- scene preparation ---------------------------------------------------------------------------------------------------
// Leaf triangles in the wide tree's order, from the caller's 64-byte records (the same subtractions triangleHit starts with).
__global__ void wideTrianglesKernel(const float* __restrict__ src, const unsigned int* __restrict__ triOrig, unsigned int n, float4* __restrict__ triA,
                                    float2* __restrict__ triB) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int orig = triOrig[i];
    const float* t = src + 16ull * orig;
    const f3 v0 = mk3(t[0], t[1], t[2]), v1 = mk3(t[3], t[4], t[5]), v2 = mk3(t[6], t[7], t[8]);
    const f3 e1 = v1 - v0, e2 = v2 - v0;
    triA[2ull * i] = make_float4(v0.x, v0.y, v0.z, e1.x);
    triA[2ull * i + 1] = make_float4(e1.y, e1.z, e2.x, e2.y);
    triB[i] = make_float2(e2.z, __uint_as_float(orig));
}

// The certificate's premise about the CALLER's tree: every node's box contains its children's (bvh_node[] of numBvhNodes
// entries, root = 1). bad[0] counts violations; a tree that fails is traversed by the exact kernel only.
__global__ void checkRefTreeKernel(const float* __restrict__ nodes, unsigned int numNodes, unsigned int* bad) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x + 4u; // children of nodes >= 2 (the root box itself is never tested, kernels.cu:157-170)
    if (i >= numNodes) return;
    const float* c = nodes + 6ull * i;
    const float* p = nodes + 6ull * (i >> 1);
    bool ok = true;
    for (int a = 0; a < 3; a++) ok = ok && (!(c[a] <= c[3 + a]) || (p[a] <= c[a] && c[3 + a] <= p[3 + a])); // (an empty child box constrains nothing)
    if (!ok) atomicAdd(bad, 1u);
}
