// bvh_builder.cpp -- complete-binary-tree BVH in the layout the reference
// traversal assumes (kernels.cu:154-224, :614) and its BVH_00.04 container
// (staircase_scene.h:75-101).  The reference's builder is a sibling project
// that is not in the tree (cuda-raytracing-optimized.sln:8), so this one is
// written from the invariants the traversal and the loader imply:
//   * implicit heap: root = 1, children of i are 2i and 2i+1, slot 0 unused,
//     numBvhNodes = 2^(L+1), every leaf on level L, firstLeafIdx = numBvhNodes/2;
//   * leaf k owns triangle slots [k*N, k*N+N), N = numPrimitivesPerLeaf; unused
//     slots are marked with v[0].x = +inf (kernels.cu:202);
//   * split = median of the centroids along the longest axis of the node box,
//     lower side to the left child (TODO.txt:235-238, helper_structs.h:106);
//   * BUILD_SAH (opt-in): the tree must stay complete, so a child can take at most
//     (leaves under it) x N triangles; among the sweep positions that respect that
//     for both children, on all three axes, the one with the least
//     area(L)*|L| + area(R)*|R| is taken (the author's own next step, TODO.txt:574).
// The sort uses a total order (key, then original index) so the tree is the
// same on every machine.
#include "bvh_builder.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>

namespace crt {

namespace {

struct Builder {
    const std::vector<triangle>& src;
    std::vector<int> order;       // permutation being partitioned
    std::vector<float> cent[3];   // centroids per axis
    BuiltMesh& out;
    int leafLevel;
    int primsPerLeaf;
    int firstLeaf;
    BuildMode mode = BUILD_MEDIAN;
    std::vector<float> areaL; // SAH sweep scratch

    static float halfArea(const bvh_node& b) {
        const float dx = b.b.e[0] - b.a.e[0], dy = b.b.e[1] - b.a.e[1], dz = b.b.e[2] - b.a.e[2];
        return dx * dy + dy * dz + dz * dx;
    }

    void sortAxis(int lo, int hi, int axis) {
        const std::vector<float>& c = cent[axis];
        std::sort(order.begin() + lo, order.begin() + hi, [&c](int x, int y) {
            if (c[x] != c[y]) return c[x] < c[y];
            return x < y;
        });
    }

    // Least-cost (axis, position) among the splits both children have room for; leaves `order[lo, hi)` sorted on that axis.
    int sahSplit(int lo, int hi, int level) {
        const int n = hi - lo;
        const long long cap = (long long)(1LL << (leafLevel - level - 1)) * primsPerLeaf; // triangles a child can hold
        const int kMin = (int)std::max<long long>(1, n - cap), kMax = (int)std::min<long long>(n - 1, cap);
        const float inf = std::numeric_limits<float>::infinity();
        float bestCost = inf;
        int bestAxis = -1, bestK = (n + 1) / 2;
        areaL.resize((size_t)n + 1);
        for (int axis = 0; axis < 3; axis++) {
            sortAxis(lo, hi, axis);
            bvh_node box;
            box.a = vec3(inf, inf, inf);
            box.b = vec3(-inf, -inf, -inf);
            for (int i = 0; i < n; i++) { // areaL[k] = area of the first k triangles
                grow(box, src[order[lo + i]]);
                areaL[i + 1] = halfArea(box);
            }
            box.a = vec3(inf, inf, inf);
            box.b = vec3(-inf, -inf, -inf);
            for (int k = n - 1; k >= kMin; k--) { // right side = triangles k .. n-1
                grow(box, src[order[lo + k]]);
                if (k > kMax) continue;
                const float cost = areaL[k] * (float)k + halfArea(box) * (float)(n - k);
                if (cost < bestCost || (cost == bestCost && (axis < bestAxis || (axis == bestAxis && k < bestK)))) {
                    bestCost = cost;
                    bestAxis = axis;
                    bestK = k;
                }
            }
        }
        if (bestAxis < 0) { bestAxis = 0; bestK = std::min(std::max((n + 1) / 2, kMin), kMax); }
        if (bestAxis != 2) sortAxis(lo, hi, bestAxis);
        return lo + bestK;
    }

    Builder(const std::vector<triangle>& s, BuiltMesh& o) : src(s), out(o) {}

    static void grow(bvh_node& b, const triangle& t) {
        for (int k = 0; k < 3; k++)
            for (int a = 0; a < 3; a++) {
                b.a.e[a] = std::min(b.a.e[a], t.v[k].e[a]);
                b.b.e[a] = std::max(b.b.e[a], t.v[k].e[a]);
            }
    }

    void build(int node, int lo, int hi, int level) {
        const float inf = std::numeric_limits<float>::infinity();
        bvh_node box;
        box.a = vec3(inf, inf, inf);
        box.b = vec3(-inf, -inf, -inf); // empty box: every slab test fails (t_max < t_min)
        for (int i = lo; i < hi; i++) grow(box, src[order[i]]);
        out.nodes[node] = box;

        if (level == leafLevel) {
            const int first = (node - firstLeaf) * primsPerLeaf;
            int n = hi - lo;
            for (int i = 0; i < n; i++) out.tris[first + i] = src[order[lo + i]];
            return; // remaining slots keep the +inf sentinel
        }

        const int n = hi - lo;
        if (mode == BUILD_SAH && n >= 2) {
            const int mid = sahSplit(lo, hi, level);
            build(2 * node, lo, mid, level + 1);
            build(2 * node + 1, mid, hi, level + 1);
            return;
        }
        int axis = 0;
        if (n > 0) {
            float ext[3] = {box.b.e[0] - box.a.e[0], box.b.e[1] - box.a.e[1], box.b.e[2] - box.a.e[2]};
            axis = (ext[0] >= ext[1]) ? 0 : 1; // max_component(), vec3.h:117-120
            axis = (ext[axis] >= ext[2]) ? axis : 2;
            const std::vector<float>& c = cent[axis];
            std::sort(order.begin() + lo, order.begin() + hi, [&c](int x, int y) {
                if (c[x] != c[y]) return c[x] < c[y];
                return x < y;
            });
        }
        const int mid = lo + (n + 1) / 2;
        build(2 * node, lo, mid, level + 1);
        build(2 * node + 1, mid, hi, level + 1);
    }
};

} // namespace

bool buildBvh(const std::vector<triangle>& tris, int primsPerLeaf, BuiltMesh& out, BuildMode mode) {
    if (primsPerLeaf < 1) return false;
    const int n = (int)tris.size();
    int level = 0;
    while ((long long)(1LL << level) * primsPerLeaf < n) level++;
    if (level > 30) return false; // the reference's bit-stack is 32 bits wide (kernels.cu:157)

    const int numLeaves = 1 << level;
    out.primsPerLeaf = primsPerLeaf;
    out.numRealTris = n;
    out.nodes.assign((size_t)2 * numLeaves, bvh_node());
    triangle pad;
    std::memset(&pad, 0, sizeof(pad));
    pad.v[0].e[0] = std::numeric_limits<float>::infinity();
    out.tris.assign((size_t)numLeaves * primsPerLeaf, pad);

    Builder b(tris, out);
    b.leafLevel = level;
    b.primsPerLeaf = primsPerLeaf;
    b.firstLeaf = numLeaves;
    b.mode = mode;
    b.order.resize(n);
    for (int i = 0; i < n; i++) b.order[i] = i;
    for (int a = 0; a < 3; a++) {
        b.cent[a].resize(n);
        for (int i = 0; i < n; i++)
            b.cent[a][i] = (tris[i].v[0].e[a] + tris[i].v[1].e[a] + tris[i].v[2].e[a]) * (1.0f / 3.0f);
    }
    b.build(1, 0, n, 0);

    // slot 0 is never read by the traversal; keep it well defined for hashing/IO
    out.nodes[0] = out.nodes[1];
    out.bounds.min = out.nodes[1].a;
    out.bounds.max = out.nodes[1].b;
    return true;
}

static const char kMagic[10] = {'B', 'V', 'H', '_', '0', '0', '.', '0', '4', '\0'};

bool saveBvhFile(const char* path, const BuiltMesh& m) {
    FILE* f = std::fopen(path, "wb");
    if (!f) return false;
    const int numTris = (int)m.tris.size();
    const int numNodes = (int)m.nodes.size();
    bool ok = std::fwrite(kMagic, 1, sizeof(kMagic), f) == sizeof(kMagic);
    ok = ok && std::fwrite(&numTris, sizeof(int), 1, f) == 1;
    ok = ok && std::fwrite(m.tris.data(), sizeof(triangle), numTris, f) == (size_t)numTris;
    ok = ok && std::fwrite(&numNodes, sizeof(int), 1, f) == 1;
    ok = ok && std::fwrite(m.nodes.data(), sizeof(bvh_node), numNodes, f) == (size_t)numNodes;
    ok = ok && std::fwrite(&m.bounds.min, sizeof(vec3), 1, f) == 1;
    ok = ok && std::fwrite(&m.bounds.max, sizeof(vec3), 1, f) == 1;
    ok = ok && std::fwrite(&m.primsPerLeaf, sizeof(int), 1, f) == 1;
    std::fclose(f);
    return ok;
}

bool loadBvhFile(const char* path, BuiltMesh& m) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    char magic[sizeof(kMagic)];
    int numTris = 0, numNodes = 0;
    bool ok = std::fread(magic, 1, sizeof(magic), f) == sizeof(magic) && std::memcmp(magic, kMagic, sizeof(kMagic)) == 0;
    ok = ok && std::fread(&numTris, sizeof(int), 1, f) == 1 && numTris >= 0;
    if (ok) {
        m.tris.resize(numTris);
        ok = std::fread(m.tris.data(), sizeof(triangle), numTris, f) == (size_t)numTris;
    }
    ok = ok && std::fread(&numNodes, sizeof(int), 1, f) == 1 && numNodes >= 2 && (numNodes & (numNodes - 1)) == 0;
    if (ok) {
        m.nodes.resize(numNodes);
        ok = std::fread(m.nodes.data(), sizeof(bvh_node), numNodes, f) == (size_t)numNodes;
    }
    ok = ok && std::fread(&m.bounds.min, sizeof(vec3), 1, f) == 1;
    ok = ok && std::fread(&m.bounds.max, sizeof(vec3), 1, f) == 1;
    ok = ok && std::fread(&m.primsPerLeaf, sizeof(int), 1, f) == 1 && m.primsPerLeaf >= 1;
    std::fclose(f);
    if (!ok) return false;
    // the traversal reads leaf k at [k*N, k*N+N): refuse files that would run off the end
    if ((long long)(numNodes / 2) * m.primsPerLeaf > numTris) return false;
    m.numRealTris = 0;
    for (const triangle& t : m.tris)
        if (!std::isinf(t.v[0].e[0])) m.numRealTris++;
    return true;
}

} // namespace crt
