// wide_bvh.h -- the renderer's OWN acceleration structure: an 8-wide BVH with 7-bit quantised child boxes.
//
// The reference walks the caller's median-split binary heap (kernels.cu:154-224) and its author's first wish is a better
// tree (TODO.txt:574,590-596). initRenderer therefore builds, from the caller's triangle[] alone, a surface-area-heuristic
// binary tree, collapses it to 8-wide nodes and stores every wide node in ONE 96-byte record (three 32-byte loads):
//
//   bytes  0..11  px, py, pz    origin of the node's quantisation grid (floats)
//         12..14  ex, ey, ez    biased exponent byte of 128 * step  (step = 2^k: the grid spacing of that axis; informative)
//             15  imask         bit s = slot s is an inner (wide) node
//         16..19  childBase     index of the first inner child; inner child in slot s = childBase + popc(imask & ((1<<s)-1))
//         20..23  triBase       index of the node's first leaf triangle (triangles of all leaf slots are contiguous)
//         24..31  meta[8]       per slot: 0 = empty; inner: 0x20 | (24 + s); leaf: (count << 5) | offset, count 1..3, offset < 24
//         32..79  qlo x[8] y[8] z[8], qhi x[8] y[8] z[8]     bytes 0x80 | q, q in 0..127
//         80..91  sx, sy, sz    128 * step as floats (what the traversal multiplies by 1/d)
//         92..95  spare
//
// A stored byte b = 0x80 | q placed in bits 16..23 of a float whose top byte is 0x3F reads as m = 1 + q/128, so one PRMT
// turns a plane byte into a float and the slab distance is ONE fused multiply-add:  t = m * A + B  with
//   A = 128 * step / d,   B = (p - o) / d - A       (plane = p + q * step).
// Child boxes are conservative: each is grown by `pad` (below) before it is rounded outward to the grid, and `pad` exceeds
// the rounding error of that arithmetic for every ray whose origin lies within 4x the scene's coordinate range (others
// are traced by the order-exact kernel), so a child box is never missed by the arithmetic.
//
// Slots are assigned by octant (child centre relative to the node centre), so that "slot index XOR ray octant" orders the
// hit children front to back without sorting (Ylitie, Karras, Laine 2017, "Efficient incoherent ray traversal on GPUs
// through compressed wide BVHs" -- the published technique this layout follows; the code is ours).
//
// What keeps results identical to the reference (whose answer depends on ITS tree only through exact ties and last-ulp
// box culls) is the certificate in wide_traverse.cuh; this file only builds.
#pragma once

#include <cstdint>
#include <vector>

#include "rt_types.h"

struct WideNode {
    float p[3];
    uint8_t e[3];
    uint8_t imask;
    uint32_t childBase;
    uint32_t triBase;
    uint8_t meta[8];
    uint8_t qlo[3][8];
    uint8_t qhi[3][8];
    float scale[3];
    uint32_t spare;
};
static_assert(sizeof(WideNode) == 96, "WideNode layout");

#ifndef WIDE_MAX_LEAF_TRIS
#define WIDE_MAX_LEAF_TRIS 3
#endif
#define WIDE_PAD_SCALE (1.0f / 262144.0f) // 2^-18 of the largest |coordinate| per axis (error bound: DESIGN.md)
#define WIDE_ORIGIN_RANGE 4.0f            // rays whose |origin| exceeds this multiple of the coordinate range use the exact kernel

struct WideBvhStats {
    uint32_t numNodes = 0, numTris = 0, numBinaryNodes = 0;
    int maxDepth = 0;        // wide levels, root = 1: the traversal stack needs maxDepth entries
    double sahCost = 0.0;    // of the binary tree
    double msBinary = 0.0, msCollapse = 0.0, msTotal = 0.0;
    int threads = 0;
};

struct WideBvhHost {
    std::vector<WideNode> nodes;    // [0] = root
    std::vector<uint32_t> triOrig;  // leaf triangle k is the caller's slot triOrig[k]
    float pad[3] = {0, 0, 0};
    float range[3] = {0, 0, 0};     // largest |coordinate| per axis (of the triangles)
    WideBvhStats stats;
};

// Builds from the caller's triangle slots (slots whose v[0].x is +inf are unused leaf padding, kernels.cu:202).
// threads <= 0: hardware concurrency (at most 16). Returns false when there is nothing to build (no real triangle).
bool buildWideBvh(const triangle* tris, uint32_t numSlots, int threads, WideBvhHost& out);
