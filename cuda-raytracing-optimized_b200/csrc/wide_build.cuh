// wide_build.cuh -- the wide tree (wide_bvh.h) built ON THE DEVICE, inside initRenderer, from the caller's triangles.
//
// The host builder (wide_bvh.cpp) costs ~45 ms of 16 host threads for the 311 740-triangle benchmark mesh -- part of every
// end-to-end frame -- and for that price bins along one axis only. Here the same construction runs level by level on the GPU:
//   binary tree   top-down, every node of a level at once: 32 bins x 3 axes per node filled with atomics (min / max on
//                 order-preserving integer images of the floats), one warp per node scans the bins and picks the plane of
//                 least surface-area cost, one thread per triangle moves it to its child's range. ~35 levels x 4 launches.
//   collapse      level by level, one thread per wide node: open the largest child until there are 8, assign octant slots,
//                 quantise outward (the host code, statement by statement), allocate children / leaf triangles with two atomics.
// The TOPOLOGY is deterministic (bins are order-independent); the order of triangles inside a leaf and the numbering of
// nodes depend on the order in which atomics land, which no result depends on (equal-t triangles are the certificate's
// business, wide_traverse.cuh).
// Validation: tests download the tree and run the host-side structural check; every render / ray-batch parity test runs on it.
#pragma once

#include <cuda_runtime.h>

#include "vecmath.cuh"
#include "wide_bvh.h"

#define GB_BINS 32
#define GB_NODE_COST 1.0f
#define GB_TRI_COST 1.0f
#define GB_WIDE_NODE_COST 1.0f // collapse cost model (wide_bvh.cpp): one wide-node visit against one triangle test
#define GB_WIDE_TRI_COST 0.3f
#define GB_LEAF_BIT 0x80000000u

struct GpuBuild {
    unsigned int n;                 // capacity: triangle slots
    // primitives
    unsigned int* primSlot;         // caller's slot of primitive p
    float4* primLo;                 // box min (w unused)
    float4* primHi;
    unsigned int* order[2];         // primitives in node-contiguous order (ping-pong per level)
    unsigned int* nodeOf[2];        // binary node of a position
    // binary nodes (2n)
    float4* nLo;                    // box min, w = left child (0 = leaf)
    float4* nHi;                    // box max, w = primitives below the node
    int* cLo;                       // centroid bounds, order-preserving ints, 3 per node
    int* cHi;
    unsigned int* nFirst;
    unsigned int* nCursor;          // primitives already moved into the node's range
    int* nSlot;                     // index in the current active list, -1 = not being split
    uint2* nSplit;                  // {axis (3 = by position), plane}
    unsigned int* active[2];
    // bins: [active][3][GB_BINS]
    unsigned int* binCnt;
    int* binLo;                     // 3 per bin
    int* binHi;
    size_t binCapacity;             // active nodes the bin arrays hold
    // collapse
    float* dpCost;                  // 7 per binary node: C(n, 1..7) (wide_bvh.cpp)
    unsigned char* dpDec;           // 8 per binary node: the arg-mins
    unsigned int* pending;          // binary node of wide node w
    unsigned int* depthOf;
    // counters: [0] primitives [1] binary nodes [2],[3] active counts [4] wide nodes [5] leaf triangles [6] depth [8..13] root box ints [14..19] root centroid ints
    unsigned int* counters;
};

__device__ __forceinline__ int gbEnc(float f) { // order-preserving float -> int (an involution on the bit pattern)
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float gbDec(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__device__ __forceinline__ int gbBinOf(float c, float cmin, float scale) {
    const int k = (int)((c - cmin) * scale);
    return k < 0 ? 0 : (k >= GB_BINS ? GB_BINS - 1 : k);
}
__device__ __forceinline__ float gbScale(float ext) { // 0 = this axis cannot separate the centroids
    if (!(ext > 0.0f)) return 0.0f;
    const float s = (float)GB_BINS * 0.999f / ext;
    return isfinite(s) ? s : 0.0f;
}
__device__ __forceinline__ float gbHalfArea(const float* lo, const float* hi) {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (!(dx >= 0.0f && dy >= 0.0f && dz >= 0.0f)) return 0.0f;
    return dx * dy + dy * dz + dz * dx;
}

// ---- primitives -------------------------------------------------------------------------------------------------------------
__global__ void gbPrimsKernel(const float* __restrict__ tris, unsigned int numSlots, GpuBuild b) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int lane = threadIdx.x & 31u;
    bool real = false;
    float lo[3], hi[3];
    if (i < numSlots) {
        const float* t = tris + 16ull * i;
        real = !isinf(t[0]);
        for (int a = 0; a < 3; a++) {
            lo[a] = fminf(fminf(t[a], t[3 + a]), t[6 + a]);
            hi[a] = fmaxf(fmaxf(t[a], t[3 + a]), t[6 + a]);
        }
    }
    const unsigned int mask = __ballot_sync(0xFFFFFFFFu, real);
    if (mask == 0u) return;
    unsigned int base = 0;
    const unsigned int leader = __ffs(mask) - 1;
    if (lane == leader) base = atomicAdd(&b.counters[0], __popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (real) {
        const unsigned int p = base + __popc(mask & ((1u << lane) - 1u));
        b.primSlot[p] = i;
        b.primLo[p] = make_float4(lo[0], lo[1], lo[2], 0.0f);
        b.primHi[p] = make_float4(hi[0], hi[1], hi[2], 0.0f);
        b.order[0][p] = p;
        b.nodeOf[0][p] = 0u;
        int* root = (int*)b.counters + 8;
        for (int a = 0; a < 3; a++) {
            const float c = 0.5f * (lo[a] + hi[a]);
            atomicMin(&root[a], gbEnc(lo[a]));
            atomicMax(&root[3 + a], gbEnc(hi[a]));
            atomicMin(&root[6 + a], gbEnc(c));
            atomicMax(&root[9 + a], gbEnc(c));
        }
    }
}

__global__ void gbRootKernel(GpuBuild b) {
    const unsigned int n = b.counters[0];
    const int* root = (const int*)b.counters + 8;
    b.nLo[0] = make_float4(gbDec(root[0]), gbDec(root[1]), gbDec(root[2]), __uint_as_float(0u));
    b.nHi[0] = make_float4(gbDec(root[3]), gbDec(root[4]), gbDec(root[5]), __uint_as_float(n));
    for (int a = 0; a < 3; a++) { b.cLo[a] = root[6 + a]; b.cHi[a] = root[9 + a]; }
    b.nFirst[0] = 0u;
    b.counters[1] = 1u;
    b.counters[3] = 0u;
    if (n >= 2u) { b.nSlot[0] = 0; b.active[0][0] = 0u; b.counters[2] = 1u; }
    else { b.nSlot[0] = -1; b.counters[2] = 0u; }
}

// ---- one level of the binary build ------------------------------------------------------------------------------------------
__global__ void gbClearBinsKernel(GpuBuild b, unsigned int activeCount) {
    const size_t total = (size_t)activeCount * 3 * GB_BINS;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        b.binCnt[i] = 0u;
        for (int a = 0; a < 3; a++) { b.binLo[3 * i + a] = 0x7F800000; b.binHi[3 * i + a] = (int)0xFF800000 ^ 0x7FFFFFFF; } // +inf / -inf images
    }
}

__global__ void gbBinKernel(GpuBuild b, int cur) {
    const unsigned int n = b.counters[0];
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int node = b.nodeOf[cur][i];
    const int slot = b.nSlot[node];
    if (slot < 0) return;
    const unsigned int p = b.order[cur][i];
    const float4 lo4 = b.primLo[p], hi4 = b.primHi[p];
    const float lo[3] = {lo4.x, lo4.y, lo4.z}, hi[3] = {hi4.x, hi4.y, hi4.z};
    for (int a = 0; a < 3; a++) {
        const float cmin = gbDec(b.cLo[3 * node + a]);
        const float scale = gbScale(gbDec(b.cHi[3 * node + a]) - cmin);
        if (scale == 0.0f) continue;
        const int k = gbBinOf(0.5f * (lo[a] + hi[a]), cmin, scale);
        const size_t bin = ((size_t)slot * 3 + a) * GB_BINS + k;
        // lanes of this warp that hit the same bin combine their update (the input order is spatially coherent, so the top
        // levels would otherwise serialise thousands of atomics on a few addresses)
        const unsigned int peers = __match_any_sync(__activemask(), (unsigned long long)bin);
        const unsigned int leader = __ffs(peers) - 1;
        int l0 = __reduce_min_sync(peers, gbEnc(lo[0])), l1 = __reduce_min_sync(peers, gbEnc(lo[1])), l2 = __reduce_min_sync(peers, gbEnc(lo[2]));
        int h0 = __reduce_max_sync(peers, gbEnc(hi[0])), h1 = __reduce_max_sync(peers, gbEnc(hi[1])), h2 = __reduce_max_sync(peers, gbEnc(hi[2]));
        if ((threadIdx.x & 31u) == leader) {
            atomicAdd(&b.binCnt[bin], (unsigned int)__popc(peers));
            atomicMin(&b.binLo[3 * bin], l0); atomicMin(&b.binLo[3 * bin + 1], l1); atomicMin(&b.binLo[3 * bin + 2], l2);
            atomicMax(&b.binHi[3 * bin], h0); atomicMax(&b.binHi[3 * bin + 1], h1); atomicMax(&b.binHi[3 * bin + 2], h2);
        }
    }
}

struct GbBox {
    float lo[3], hi[3];
    unsigned int cnt;
};
__device__ __forceinline__ void gbMerge(GbBox& a, const GbBox& o) {
    for (int k = 0; k < 3; k++) { a.lo[k] = fminf(a.lo[k], o.lo[k]); a.hi[k] = fmaxf(a.hi[k], o.hi[k]); }
    a.cnt += o.cnt;
}
__device__ __forceinline__ GbBox gbShflUp(const GbBox& v, int d) {
    GbBox r;
    for (int k = 0; k < 3; k++) { r.lo[k] = __shfl_up_sync(0xFFFFFFFFu, v.lo[k], d); r.hi[k] = __shfl_up_sync(0xFFFFFFFFu, v.hi[k], d); }
    r.cnt = __shfl_up_sync(0xFFFFFFFFu, v.cnt, d);
    return r;
}
__device__ __forceinline__ GbBox gbShflDown(const GbBox& v, int d) {
    GbBox r;
    for (int k = 0; k < 3; k++) { r.lo[k] = __shfl_down_sync(0xFFFFFFFFu, v.lo[k], d); r.hi[k] = __shfl_down_sync(0xFFFFFFFFu, v.hi[k], d); }
    r.cnt = __shfl_down_sync(0xFFFFFFFFu, v.cnt, d);
    return r;
}
__device__ __forceinline__ GbBox gbShfl(const GbBox& v, int src) {
    GbBox r;
    for (int k = 0; k < 3; k++) { r.lo[k] = __shfl_sync(0xFFFFFFFFu, v.lo[k], src); r.hi[k] = __shfl_sync(0xFFFFFFFFu, v.hi[k], src); }
    r.cnt = __shfl_sync(0xFFFFFFFFu, v.cnt, src);
    return r;
}

// One warp per active node, lane = bin: for each axis the boxes of bins 0..k and k+1..31, the SAH cost of that plane, the best
// of all (axis, plane); then the node becomes a leaf (<= 3 triangles and no split pays for itself) or gets two children.
__global__ void gbSplitKernel(GpuBuild b, int cur, unsigned int activeCount) {
    const unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned int lane = threadIdx.x & 31u;
    if (warp >= activeCount) return;
    const unsigned int node = b.active[cur][warp];
    const float4 nlo = b.nLo[node], nhi = b.nHi[node];
    const unsigned int count = __float_as_uint(nhi.w);
    const float inf = __int_as_float(0x7F800000);
    float bestCost = inf;
    unsigned int bestKey = 0xFFFFFFFFu; // axis * 32 + plane
    GbBox bestL, bestR;
    for (int k = 0; k < 3; k++) { bestL.lo[k] = bestR.lo[k] = inf; bestL.hi[k] = bestR.hi[k] = -inf; }
    bestL.cnt = bestR.cnt = 0u;
    for (int a = 0; a < 3; a++) {
        const size_t bin = ((size_t)warp * 3 + a) * GB_BINS + lane;
        GbBox v;
        v.cnt = b.binCnt[bin];
        for (int k = 0; k < 3; k++) { v.lo[k] = gbDec(b.binLo[3 * bin + k]); v.hi[k] = gbDec(b.binHi[3 * bin + k]); }
        GbBox pre = v, suf = v;
        for (int d = 1; d < 32; d <<= 1) {
            const GbBox up = gbShflUp(pre, d), dn = gbShflDown(suf, d);
            if ((int)lane >= d) gbMerge(pre, up);
            if ((int)lane + d < 32) gbMerge(suf, dn);
        }
        GbBox right = gbShflDown(suf, 1); // bins lane+1 .. 31
        const bool valid = lane < 31u && pre.cnt > 0u && pre.cnt < count;
        if (valid) {
            const float cost = gbHalfArea(pre.lo, pre.hi) * (float)pre.cnt + gbHalfArea(right.lo, right.hi) * (float)right.cnt;
            if (cost < bestCost) { bestCost = cost; bestKey = (unsigned int)a * 32u + lane; bestL = pre; bestR = right; }
        }
    }
    // warp argmin (ties: lowest key)
    float wc = bestCost;
    unsigned int wk = bestKey;
    for (int d = 16; d > 0; d >>= 1) {
        const float oc = __shfl_xor_sync(0xFFFFFFFFu, wc, d);
        const unsigned int ok = __shfl_xor_sync(0xFFFFFFFFu, wk, d);
        if (oc < wc || (oc == wc && ok < wk)) { wc = oc; wk = ok; }
    }
    const bool found = wk != 0xFFFFFFFFu;
    const int src = found ? (int)(wk & 31u) : 0;
    GbBox L = gbShfl(bestL, src), R = gbShfl(bestR, src);
    // (the lane that owns the winning plane also owns the winning axis: its best is the warp's best)
    if (lane != 0) return;
    const float nl[3] = {nlo.x, nlo.y, nlo.z}, nh[3] = {nhi.x, nhi.y, nhi.z};
    const float area = gbHalfArea(nl, nh);
    if (count <= WIDE_MAX_LEAF_TRIS) {
        const float leafCost = GB_TRI_COST * (float)count * area;
        if (!found || !(GB_NODE_COST * area + GB_TRI_COST * wc < leafCost)) { b.nSlot[node] = -1; return; } // stays a leaf
    }
    unsigned int axis = 3u, plane = 0u;
    const unsigned int first = b.nFirst[node];
    if (found) {
        axis = wk >> 5; plane = wk & 31u;
    } else { // all centroids coincide: split the range in the middle (children inherit the node's box)
        L.cnt = count / 2u; R.cnt = count - L.cnt;
        for (int k = 0; k < 3; k++) { L.lo[k] = R.lo[k] = nl[k]; L.hi[k] = R.hi[k] = nh[k]; }
    }
    const unsigned int left = atomicAdd(&b.counters[1], 2u);
    b.nLo[node].w = __uint_as_float(left); // (nHi.w keeps the number of primitives below the node)
    b.nSplit[node] = make_uint2(axis, plane);
    for (int s = 0; s < 2; s++) {
        const GbBox& cb = s ? R : L;
        const unsigned int child = left + s;
        b.nLo[child] = make_float4(cb.lo[0], cb.lo[1], cb.lo[2], __uint_as_float(0u));
        b.nHi[child] = make_float4(cb.hi[0], cb.hi[1], cb.hi[2], __uint_as_float(cb.cnt));
        b.nFirst[child] = first + (s ? L.cnt : 0u);
        b.nCursor[child] = 0u;
        for (int k = 0; k < 3; k++) { b.cLo[3 * child + k] = 0x7F800000; b.cHi[3 * child + k] = (int)0xFF800000 ^ 0x7FFFFFFF; }
        if (cb.cnt >= 2u) {
            const unsigned int slot = atomicAdd(&b.counters[2 + (cur ^ 1)], 1u);
            b.active[cur ^ 1][slot] = child;
            b.nSlot[child] = (int)slot;
        } else {
            b.nSlot[child] = -1;
        }
    }
}

// One thread per primitive position: primitives of a node that was split move into their child's range.
__global__ void gbPartitionKernel(GpuBuild b, int cur) {
    const unsigned int n = b.counters[0];
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int node = b.nodeOf[cur][i];
    const unsigned int p = b.order[cur][i];
    const unsigned int left = b.nSlot[node] >= 0 ? __float_as_uint(b.nLo[node].w) : 0u;
    if (left == 0u) { // not split at this level: stays where it is
        b.order[cur ^ 1][i] = p;
        b.nodeOf[cur ^ 1][i] = node;
        return;
    }
    const uint2 split = b.nSplit[node];
    const float4 lo4 = b.primLo[p], hi4 = b.primHi[p];
    const float c[3] = {0.5f * (lo4.x + hi4.x), 0.5f * (lo4.y + hi4.y), 0.5f * (lo4.z + hi4.z)};
    unsigned int side;
    if (split.x < 3u) {
        const float cmin = gbDec(b.cLo[3 * node + split.x]);
        const float scale = gbScale(gbDec(b.cHi[3 * node + split.x]) - cmin);
        side = (unsigned int)gbBinOf(c[split.x], cmin, scale) > split.y ? 1u : 0u;
    } else {
        side = (i - b.nFirst[node]) >= __float_as_uint(b.nHi[left].w) ? 1u : 0u;
    }
    const unsigned int child = left + side;
    const unsigned int peers = __match_any_sync(__activemask(), child);
    const unsigned int leader = __ffs(peers) - 1;
    const unsigned int lane = threadIdx.x & 31u;
    const int c0 = gbEnc(c[0]), c1 = gbEnc(c[1]), c2 = gbEnc(c[2]);
    const int m0 = __reduce_min_sync(peers, c0), m1 = __reduce_min_sync(peers, c1), m2 = __reduce_min_sync(peers, c2);
    const int x0 = __reduce_max_sync(peers, c0), x1 = __reduce_max_sync(peers, c1), x2 = __reduce_max_sync(peers, c2);
    unsigned int base = 0;
    if (lane == leader) {
        base = atomicAdd(&b.nCursor[child], (unsigned int)__popc(peers));
        atomicMin(&b.cLo[3 * child], m0); atomicMin(&b.cLo[3 * child + 1], m1); atomicMin(&b.cLo[3 * child + 2], m2);
        atomicMax(&b.cHi[3 * child], x0); atomicMax(&b.cHi[3 * child + 1], x1); atomicMax(&b.cHi[3 * child + 2], x2);
    }
    base = __shfl_sync(peers, base, leader);
    const unsigned int dest = b.nFirst[child] + base + __popc(peers & ((1u << lane) - 1u));
    b.order[cur ^ 1][dest] = p;
    b.nodeOf[cur ^ 1][dest] = child;
}

// after the last level the nodes that were active but not split are leaves: nothing to do (nLo.w == 0 marks a leaf)

// ---- collapse to 8-wide nodes -----------------------------------------------------------------------------------------------
// The dynamic programme of wide_bvh.cpp (Ylitie, Karras, Laine 2017, section 4.1) for the binary nodes [start, end) of one
// level, deepest level first: C(n, i) = least cost of the subtree of n represented by at most i wide-tree roots.
__global__ void gbCollapseCostKernel(GpuBuild b, unsigned int start, unsigned int end, float triCost, int greedy) {
    const unsigned int node = start + blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= end) return;
    const float4 lo = b.nLo[node], hi = b.nHi[node];
    const float l3[3] = {lo.x, lo.y, lo.z}, h3[3] = {hi.x, hi.y, hi.z};
    const float area = gbHalfArea(l3, h3);
    const unsigned int span = __float_as_uint(hi.w), left = __float_as_uint(lo.w);
    const float inf = __int_as_float(0x7F800000);
    const float leaf = (span <= WIDE_MAX_LEAF_TRIS && !(greedy && left != 0u)) ? area * (float)span * triCost : inf;
    float* C = b.dpCost + 7ull * node;
    unsigned char* D = b.dpDec + 8ull * node;
    if (left == 0u) {
        for (int i = 0; i < 7; i++) { C[i] = leaf; D[i] = 0; }
        D[7] = 0;
        return;
    }
    float L[7], R[7];
    for (int i = 0; i < 7; i++) { L[i] = b.dpCost[7ull * left + i]; R[i] = b.dpCost[7ull * (left + 1u) + i]; }
    float dist[9];
    unsigned char kbest[9];
#pragma unroll
    for (int j = 2; j <= 8; j++) {
        dist[j] = inf;
        kbest[j] = 1;
#pragma unroll
        for (int k = 1; k <= 7; k++) {
            if (k < j - 7 || k > j - 1) continue;
            const float v = L[k - 1] + R[j - k - 1];
            if (v < dist[j]) { dist[j] = v; kbest[j] = (unsigned char)k; }
        }
    }
    const float inner = dist[8] + area * GB_WIDE_NODE_COST;
    D[7] = kbest[8];
    float prev;
    if (leaf <= inner) { prev = leaf; D[0] = 0; }
    else { prev = inner; D[0] = 1; }
    C[0] = prev;
#pragma unroll
    for (int i = 2; i <= 7; i++) {
        if (dist[i] < prev) { prev = dist[i]; D[i - 1] = kbest[i]; }
        else D[i - 1] = 0;
        C[i - 1] = prev;
    }
}

__device__ __forceinline__ unsigned char gbQuantByte(int q) { return (unsigned char)(0x80 | (q < 0 ? 0 : (q > 127 ? 127 : q))); }

// One thread per wide node of the level [start, end): the statements of wide_bvh.cpp's collapse (children, octant slots, grid,
// outward quantisation with the even-slot excess), children and leaf triangles allocated with one atomic each.
__global__ void gbCollapseKernel(GpuBuild b, int finalOrder, unsigned int start, unsigned int end, float3 pad, WideNode* __restrict__ out,
                                 unsigned int* __restrict__ triOrig, int greedy) {
    const unsigned int w = start + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= end) return;
    const unsigned int node2 = b.pending[w];
    const unsigned int depth = b.depthOf[w];
    atomicMax(&b.counters[6], depth);
    const float padv[3] = {pad.x, pad.y, pad.z};
    unsigned int child[8]; // binary nodes, | GB_LEAF_BIT = emitted as one leaf (possibly an inner binary node with <= 3 triangles)
    int nc = 0;
    {
        const unsigned int left = __float_as_uint(b.nLo[node2].w);
        if (left == 0u) {
            child[nc++] = node2 | GB_LEAF_BIT; // the whole tree is one leaf
        } else if (greedy) { // (diagnostic: open the inner child with the largest surface area until there are 8)
            child[nc++] = left; child[nc++] = left + 1u;
            while (nc < 8) {
                int pick = -1;
                float bestArea = -1.0f;
                for (int k = 0; k < nc; k++) {
                    const float4 lo = b.nLo[child[k]], hi = b.nHi[child[k]];
                    if (__float_as_uint(lo.w) != 0u) {
                        const float l3[3] = {lo.x, lo.y, lo.z}, h3[3] = {hi.x, hi.y, hi.z};
                        const float ar = gbHalfArea(l3, h3);
                        if (ar > bestArea) { bestArea = ar; pick = k; }
                    }
                }
                if (pick < 0) break;
                const unsigned int l = __float_as_uint(b.nLo[child[pick]].w);
                child[pick] = l;
                child[nc++] = l + 1u;
            }
            for (int k = 0; k < nc; k++)
                if (__float_as_uint(b.nLo[child[k]].w) == 0u) child[k] |= GB_LEAF_BIT;
        } else { // follow the arg-mins of D(node, 8)
            unsigned int stackNode[16];
            int stackBudget[16];
            int sp = 0;
            const int k8 = b.dpDec[8ull * node2 + 7];
            stackNode[sp] = left + 1u; stackBudget[sp++] = 8 - k8;
            stackNode[sp] = left; stackBudget[sp++] = k8;
            while (sp > 0) {
                sp--;
                const unsigned int nd = stackNode[sp];
                int budget = stackBudget[sp];
                const unsigned char* D = b.dpDec + 8ull * nd;
                while (budget > 1 && D[budget - 1] == 0) budget--;
                if (budget == 1) {
                    child[nc++] = D[0] ? nd : (nd | GB_LEAF_BIT);
                } else {
                    const int k = D[budget - 1];
                    const unsigned int cl = __float_as_uint(b.nLo[nd].w);
                    stackNode[sp] = cl + 1u; stackBudget[sp++] = budget - k;
                    stackNode[sp] = cl; stackBudget[sp++] = k;
                }
            }
        }
    }
    float nbLo[3] = {__int_as_float(0x7F800000), __int_as_float(0x7F800000), __int_as_float(0x7F800000)};
    float nbHi[3] = {-nbLo[0], -nbLo[1], -nbLo[2]};
    for (int k = 0; k < nc; k++) {
        const float4 lo = b.nLo[child[k] & ~GB_LEAF_BIT], hi = b.nHi[child[k] & ~GB_LEAF_BIT];
        nbLo[0] = fminf(nbLo[0], lo.x); nbLo[1] = fminf(nbLo[1], lo.y); nbLo[2] = fminf(nbLo[2], lo.z);
        nbHi[0] = fmaxf(nbHi[0], hi.x); nbHi[1] = fmaxf(nbHi[1], hi.y); nbHi[2] = fmaxf(nbHi[2], hi.z);
    }
    // octant slots: greedy on dot(child centre - node centre, slot direction)
    unsigned int slotChild[8];
    for (int s = 0; s < 8; s++) slotChild[s] = 0xFFFFFFFFu;
    unsigned int placed = 0u;
    for (int round = 0; round < nc; round++) {
        int bk = -1, bs = -1;
        float bv = -__int_as_float(0x7F800000);
        for (int k = 0; k < nc; k++) {
            if (placed & (1u << k)) continue;
            const float4 lo = b.nLo[child[k] & ~GB_LEAF_BIT], hi = b.nHi[child[k] & ~GB_LEAF_BIT];
            const float d0 = 0.5f * (lo.x + hi.x) - 0.5f * (nbLo[0] + nbHi[0]);
            const float d1 = 0.5f * (lo.y + hi.y) - 0.5f * (nbLo[1] + nbHi[1]);
            const float d2 = 0.5f * (lo.z + hi.z) - 0.5f * (nbLo[2] + nbHi[2]);
            for (int s = 0; s < 8; s++) {
                if (slotChild[s] != 0xFFFFFFFFu) continue;
                const float sc = ((s & 4) ? d0 : -d0) + ((s & 2) ? d1 : -d1) + ((s & 1) ? d2 : -d2);
                if (sc > bv) { bv = sc; bk = k; bs = s; }
            }
        }
        placed |= 1u << bk;
        slotChild[bs] = child[bk];
    }
    WideNode wn;
    {
        unsigned int* z = (unsigned int*)&wn;
        for (int k = 0; k < 24; k++) z[k] = 0u;
    }
    double step[3];
    for (int a = 0; a < 3; a++) {
        const double lo = (double)nbLo[a] - (double)padv[a], hi = (double)nbHi[a] + (double)padv[a];
        int k = (int)ceil(log2(fmax(hi - lo, 1e-300) / 126.0));
        if (k < -100) k = -100;
        float p;
        while (true) {
            step[a] = ldexp(1.0, k);
            p = (float)(lo - 0.5 * step[a]);
            if ((double)p > lo - 0.3 * step[a]) p = nextafterf(p, -__int_as_float(0x7F800000));
            if ((lo - (double)p) / step[a] >= 0.3 && (hi - (double)p) / step[a] <= 126.7) break;
            k++;
        }
        wn.p[a] = p;
        wn.e[a] = (unsigned char)min(max(k + 7 + 127, 0), 255);
        wn.scale[a] = (float)ldexp(1.0, k + 7);
    }
    unsigned int innerCount = 0u, triCount = 0u;
    for (int s = 7; s >= 0; s--) { // odd slots before the even slot that reads them as excess mantissa
        for (int a = 0; a < 3; a++) { wn.qlo[a][s] = gbQuantByte(127); wn.qhi[a][s] = gbQuantByte(0); }
        if (slotChild[s] == 0xFFFFFFFFu) continue;
        const float4 clo = b.nLo[slotChild[s] & ~GB_LEAF_BIT], chi = b.nHi[slotChild[s] & ~GB_LEAF_BIT];
        const float cl[3] = {clo.x, clo.y, clo.z}, ch[3] = {chi.x, chi.y, chi.z};
        if (!(slotChild[s] & GB_LEAF_BIT)) innerCount++;
        else triCount += __float_as_uint(chi.w);
        for (int a = 0; a < 3; a++) {
            const double lo = (double)cl[a] - (double)padv[a], hi = (double)ch[a] + (double)padv[a];
            const double inv = 1.0 / step[a];
            const double gl = (s & 1) ? 0.0 : (double)(0x3F00 | wn.qlo[a][s + 1]) / 65536.0;
            const double gh = (s & 1) ? 0.0 : (double)(0x3F00 | wn.qhi[a][s + 1]) / 65536.0;
            int ql = (int)floor((lo - (double)wn.p[a]) * inv - gl);
            int qh = (int)ceil((hi - (double)wn.p[a]) * inv - gh);
            while ((double)wn.p[a] + (ql + gl) * step[a] > lo) ql--;
            while ((double)wn.p[a] + (qh + gh) * step[a] < hi) qh++;
            wn.qlo[a][s] = gbQuantByte(ql);
            wn.qhi[a][s] = gbQuantByte(qh);
        }
    }
    const unsigned int childBase = innerCount ? atomicAdd(&b.counters[4], innerCount) : 0u;
    const unsigned int triBase = triCount ? atomicAdd(&b.counters[5], triCount) : 0u;
    wn.childBase = childBase;
    wn.triBase = triBase;
    unsigned int triOffset = 0u, inner = 0u;
    for (int s = 0; s < 8; s++) {
        if (slotChild[s] == 0xFFFFFFFFu) continue;
        const unsigned int c2 = slotChild[s] & ~GB_LEAF_BIT;
        if (slotChild[s] & GB_LEAF_BIT) {
            const unsigned int cnt = __float_as_uint(b.nHi[c2].w), first = b.nFirst[c2];
            wn.meta[s] = (unsigned char)((cnt << 5) | triOffset);
            for (unsigned int t = 0; t < cnt; t++) triOrig[triBase + triOffset + t] = b.primSlot[b.order[finalOrder][first + t]];
            triOffset += cnt;
        } else {
            wn.imask |= (unsigned char)(1u << s);
            wn.meta[s] = (unsigned char)(0x20 | (24 + s));
            b.pending[childBase + inner] = c2;
            b.depthOf[childBase + inner] = depth + 1u;
            inner++;
        }
    }
    out[w] = wn;
}
