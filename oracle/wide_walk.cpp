// oracle/wide_walk.cpp -- TEST INFRASTRUCTURE (never linked into the product).
//
// A scalar CPU walk of the product's own 8-wide quantised BVH (csrc/wide_bvh.h), statement by statement the algorithm of
// csrc/wide_traverse.cuh: group stack, octant ordering, PRMT-style plane decode (1 + q/128), t = fma(m, A, B), near-tie
// tracking and the reference-leaf certificate. It exists so that the BUILDER and the traversal LOGIC can be checked in the
// GPU-less authoring container against the CPU oracle (cpu_oracle.cpp, which restates the reference's hitBvh,
// kernels.cu:154-224): every ray the walk does not flag must return exactly the oracle's triangle, t, u, v.
// The triangle test and the reference box test are the oracle's own (oracleTriangleHit / oracleBoxDist), so both walks
// compare the same floats.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../cuda-raytracing-optimized_b200/csrc/wide_bvh.h"

extern "C" {
float oracleTriangleHit(const triangle* tri, const float o[3], const float d[3], float tMin, float tMax, float* u, float* v);
float oracleBoxDist(const float bmin[3], const float bmax[3], const float o[3], const float d[3], float tMax);
}

namespace {

inline float asFloat(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t bfind(uint32_t x) { return 31u - (uint32_t)__builtin_clz(x); }

struct Walker {
    const WideBvhHost& w;
    const kernel_scene* sc;
    uint64_t nodeVisits = 0, triTests = 0, flagged = 0, maxStack = 0;

    // returns t (FLT_MAX = miss); flag: 1 = the certificate failed (the exact kernel must re-trace this ray)
    float walk(const float o[3], const float dRaw[3], float tMin, float tMax, bool anyHit, uint32_t& triId, float& hu, float& hv, int& flag) {
        const mesh* m = sc->m;
        flag = 0;
        triId = 0xFFFFFFFFu;
        hu = hv = 0.0f;
        // rays the fast path does not take: far or non-finite origins / directions
        const float len = std::sqrt(dRaw[0] * dRaw[0] + dRaw[1] * dRaw[1] + dRaw[2] * dRaw[2]);
        float d[3] = {dRaw[0] / len, dRaw[1] / len, dRaw[2] / len};
        for (int a = 0; a < 3; a++)
            if (!(std::fabs(o[a]) <= WIDE_ORIGIN_RANGE * w.range[a]) || !std::isfinite(d[a])) { flag = 1; return FLT_MAX; }
        // hitMesh: scene bounds first (kernels.cu:297), the reference's own arithmetic
        {
            // hit_bbox and hit_bbox_dist agree on hit / miss; oracleBoxDist returns FLT_MAX on a miss
            if (!(oracleBoxDist(m->bounds.min.e, m->bounds.max.e, o, dRaw, tMax) < FLT_MAX)) return FLT_MAX;
        }
        float idir[3];
        bool neg[3];
        for (int a = 0; a < 3; a++) {
            float da = d[a];
            if (std::fabs(da) < 1e-20f) da = std::copysign(1e-20f, da);
            idir[a] = 1.0f / da;
            neg[a] = std::signbit(da); // (-0.0 clamps to -1e-20: near / far follow the sign bit)
        }
        const uint32_t octinv = (neg[0] ? 0u : 4u) | (neg[1] ? 0u : 2u) | (neg[2] ? 0u : 1u);
        const float delta = 1.00001f;
        float closest = tMax, closestPad = anyHit ? tMax : tMax * delta;
        bool tie = false;
        uint32_t winner = 0xFFFFFFFFu; // new (leaf order) index of the winning triangle

        struct Group { uint32_t x, y; };
        std::vector<Group> stack;
        Group ng{0u, 0u};
        bool rootPending = true; // the first step loads node 0 directly
        Group tg{0u, 0u};
        uint32_t metaLo = 0, metaHi = 0;
        while (true) {
            if (rootPending || ng.y > 0x00FFFFFFu) {
                uint32_t nodeIdx;
                if (rootPending) {
                    rootPending = false;
                    nodeIdx = 0;
                    ng.y = 0;
                } else {
                    const uint32_t bit = bfind(ng.y);
                    ng.y &= ~(1u << bit);
                    if (ng.y > 0x00FFFFFFu) { stack.push_back(ng); if (stack.size() > maxStack) maxStack = stack.size(); }
                    const uint32_t slot = (bit - 24u) ^ octinv;
                    const uint32_t imask = ng.y & 0xFFu;
                    nodeIdx = ng.x + (uint32_t)__builtin_popcount(imask & ((1u << slot) - 1u));
                }
                const WideNode& n = w.nodes[nodeIdx];
                nodeVisits++;
                float A[3], Bc[3];
                for (int a = 0; a < 3; a++) {
                    A[a] = n.scale[a] * idir[a];
                    Bc[a] = std::fmaf(n.p[a] - o[a], idir[a], -A[a]);
                }
                uint32_t hitSlots = 0;
                for (int s = 0; s < 8; s++) {
                    if (n.meta[s] == 0) continue;
                    float tn = 0.0f, tf = closestPad;
                    for (int a = 0; a < 3; a++) {
                        const uint8_t bn = neg[a] ? n.qhi[a][s] : n.qlo[a][s];
                        const uint8_t bf = neg[a] ? n.qlo[a][s] : n.qhi[a][s];
                        // planePair (wide_traverse.cuh): an even slot's byte is read together with the next slot's byte as excess mantissa
                        const uint32_t gn = (s & 1) ? 0u : (0x3F00u | (neg[a] ? n.qhi[a][s + 1] : n.qlo[a][s + 1]));
                        const uint32_t gf = (s & 1) ? 0u : (0x3F00u | (neg[a] ? n.qlo[a][s + 1] : n.qhi[a][s + 1]));
                        const float mn = asFloat(0x3F000000u | ((uint32_t)bn << 16) | gn);
                        const float mf = asFloat(0x3F000000u | ((uint32_t)bf << 16) | gf);
                        tn = std::fmax(tn, std::fmaf(mn, A[a], Bc[a]));
                        tf = std::fmin(tf, std::fmaf(mf, A[a], Bc[a]));
                    }
                    if (tn <= tf) hitSlots |= 1u << s;
                }
                // inner hits in priority order (bit 24 + (slot ^ octinv)), leaf hits as a slot mask
                uint32_t inner = hitSlots & n.imask, prio = 0;
                for (int s = 0; s < 8; s++)
                    if (inner & (1u << s)) prio |= 1u << (24u + ((uint32_t)s ^ octinv));
                ng.x = n.childBase;
                ng.y = prio | n.imask;
                tg.x = n.triBase;
                tg.y = hitSlots & ~(uint32_t)n.imask;
                std::memcpy(&metaLo, n.meta, 4);
                std::memcpy(&metaHi, n.meta + 4, 4);
            }
            // triangles of the hit leaf slots
            bool finished = false;
            while (tg.y) {
                const uint32_t s = (uint32_t)__builtin_ctz(tg.y);
                tg.y &= tg.y - 1u;
                const uint32_t meta = ((s < 4 ? metaLo : metaHi) >> (8u * (s & 3u))) & 0xFFu;
                const uint32_t first = tg.x + (meta & 31u), count = meta >> 5;
                for (uint32_t k = first; k < first + count; k++) {
                    const uint32_t orig = w.triOrig[k];
                    float u, v;
                    triTests++;
                    const float t = oracleTriangleHit(&m->tris[orig], o, dRaw, tMin, closestPad, &u, &v);
                    if (t < closestPad) { // (FLT_MAX on a miss)
                        if (anyHit) { closest = t; winner = orig; finished = true; break; }
                        if (t < closest) {
                            tie = closest <= t * delta; // the previous best (and everything seen before) is within the margin of the new one
                            closest = t;
                            closestPad = t * delta;
                            winner = orig;
                            hu = u; hv = v;
                        } else if (orig != winner) {
                            tie = true;
                        }
                    }
                }
                if (finished) break;
            }
            if (finished) break;
            if (ng.y <= 0x00FFFFFFu) {
                if (stack.empty()) break;
                ng = stack.back();
                stack.pop_back();
            }
        }
        if (winner == 0xFFFFFFFFu) return FLT_MAX; // no triangle test passes for any triangle: the reference misses too
        // certificate: the reference enters the winner's leaf iff its box test passes with t_max = its `closest` at that
        // time, which is >= min(tMax, winner's t * delta) when no other triangle lies within the margin; parent boxes
        // contain their children's, and the slab arithmetic is monotonic, so the leaf's box decides for all ancestors.
        const int N = sc->numPrimitivesPerLeaf;
        const uint32_t leaf = (uint32_t)m->numBvhNodes / 2u + winner / (uint32_t)N;
        const float bound = anyHit ? tMax : std::fmin(tMax, closest * delta);
        const float entry = oracleBoxDist(m->bvh[leaf].a.e, m->bvh[leaf].b.e, o, dRaw, bound);
        if (tie || !(entry < bound)) { flag = 1; flagged++; }
        triId = winner;
        return anyHit ? 0.0f : closest;
    }
};

} // namespace

extern "C" {

// Builds the wide tree for `sc` and walks the batch. rays as float4 {o.xyz,tMin} {d.xyz,tMax}; outHit float4 {t,u,v,triId bits};
// outFlag 1 = certificate failed. counters: [0] node visits [1] triangle tests [2] flagged rays [3] deepest stack
// [4] wide nodes [5] wide depth [6] leaf triangles [7] build microseconds.  Returns 0, or -1 when the scene has no triangle.
int wideWalkBatch(const kernel_scene* sc, const float* rayO, const float* rayD, long long n, int anyHit, float* outHit, unsigned char* outFlag,
                  unsigned long long* counters, int threads) {
    WideBvhHost w;
    if (!buildWideBvh(sc->m->tris, sc->m->numTris, threads, w)) return -1;
    Walker wk{w, sc};
    for (long long i = 0; i < n; i++) {
        uint32_t id;
        float u, v;
        int flag;
        float t = wk.walk(rayO + 4 * i, rayD + 4 * i, rayO[4 * i + 3], rayD[4 * i + 3], anyHit != 0, id, u, v, flag);
        if (!(t < rayD[4 * i + 3])) { t = FLT_MAX; id = 0xFFFFFFFFu; u = v = 0.0f; }
        if (anyHit) { id = 0xFFFFFFFFu; u = v = 0.0f; }
        outHit[4 * i] = t; outHit[4 * i + 1] = u; outHit[4 * i + 2] = v;
        std::memcpy(&outHit[4 * i + 3], &id, 4);
        outFlag[i] = (unsigned char)flag;
    }
    if (counters) {
        counters[0] = wk.nodeVisits; counters[1] = wk.triTests; counters[2] = wk.flagged; counters[3] = wk.maxStack;
        counters[4] = w.stats.numNodes; counters[5] = (unsigned long long)w.stats.maxDepth; counters[6] = w.stats.numTris;
        counters[7] = (unsigned long long)(w.stats.msTotal * 1000.0);
    }
    return 0;
}

static int checkTree(const kernel_scene* sc, const WideBvhHost& w);

// Structural check of a build: every real triangle appears exactly once; every child box (decoded in double) contains the
// padded boxes of everything below it. Returns 0 when sound, else a negative code.
int wideCheckStructure(const kernel_scene* sc, int threads, double* buildMs, unsigned long long* numNodes) {
    WideBvhHost w;
    if (!buildWideBvh(sc->m->tris, sc->m->numTris, threads, w)) return -1;
    if (buildMs) *buildMs = w.stats.msTotal;
    if (numNodes) *numNodes = w.stats.numNodes;
    return checkTree(sc, w);
}

// The same check for a tree built elsewhere (the device build, downloaded through getRendererWideTree). The padding is the
// builder's rule: 2^-18 of the largest |coordinate| of the triangles per axis. `depthOut` receives the tree's depth.
int wideCheckGiven(const kernel_scene* sc, const void* nodes, unsigned int numNodes, const unsigned int* triOrig, unsigned int numTris, int* depthOut) {
    WideBvhHost w;
    w.nodes.assign((const WideNode*)nodes, (const WideNode*)nodes + numNodes);
    w.triOrig.assign(triOrig, triOrig + numTris);
    float range[3] = {1e-30f, 1e-30f, 1e-30f};
    const mesh* m = sc->m;
    for (uint32_t i = 0; i < m->numTris; i++) {
        if (std::isinf(m->tris[i].v[0].e[0])) continue;
        for (int v = 0; v < 3; v++)
            for (int a = 0; a < 3; a++) range[a] = std::fmax(range[a], std::fabs(m->tris[i].v[v].e[a]));
    }
    for (int a = 0; a < 3; a++) { w.range[a] = range[a]; w.pad[a] = range[a] * WIDE_PAD_SCALE; }
    if (depthOut) { // depth by a walk from the root
        std::vector<int> depth(numNodes, 0);
        int deepest = 0;
        if (numNodes) depth[0] = 1;
        for (unsigned int k = 0; k < numNodes; k++) {
            const WideNode& n = w.nodes[k];
            if (depth[k] > deepest) deepest = depth[k];
            for (int s = 0; s < 8; s++)
                if (n.imask & (1u << s)) {
                    const uint32_t c = n.childBase + (uint32_t)__builtin_popcount(n.imask & ((1u << s) - 1u));
                    if (c < numNodes) depth[c] = depth[k] + 1;
                }
        }
        *depthOut = deepest;
    }
    return checkTree(sc, w);
}
}

static int checkTree(const kernel_scene* sc, const WideBvhHost& w) {
    const mesh* m = sc->m;
    std::vector<unsigned char> seen(m->numTris, 0);
    for (uint32_t id : w.triOrig) {
        if (id >= m->numTris || seen[id]) return -2;
        seen[id] = 1;
    }
    for (uint32_t i = 0; i < m->numTris; i++)
        if (!std::isinf(m->tris[i].v[0].e[0]) && !seen[i]) return -3;
    // bottom-up exact boxes (nodes are emitted breadth first: children have larger indices)
    const size_t nn = w.nodes.size();
    std::vector<double> lo(3 * nn), hi(3 * nn);
    for (size_t k = nn; k-- > 0;) {
        const WideNode& n = w.nodes[k];
        double blo[3] = {1e300, 1e300, 1e300}, bhi[3] = {-1e300, -1e300, -1e300};
        for (int s = 0; s < 8; s++) {
            if (!n.meta[s]) continue;
            double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
            if (n.imask & (1u << s)) {
                const uint32_t c = n.childBase + (uint32_t)__builtin_popcount(n.imask & ((1u << s) - 1u));
                if (c <= k || c >= nn) return -4;
                for (int a = 0; a < 3; a++) { clo[a] = lo[3 * c + a]; chi[a] = hi[3 * c + a]; }
            } else {
                const uint32_t first = n.triBase + (n.meta[s] & 31u), count = n.meta[s] >> 5;
                if (count == 0 || count > WIDE_MAX_LEAF_TRIS || first + count > w.triOrig.size()) return -5;
                for (uint32_t t = first; t < first + count; t++)
                    for (int v = 0; v < 3; v++)
                        for (int a = 0; a < 3; a++) {
                            clo[a] = std::fmin(clo[a], (double)m->tris[w.triOrig[t]].v[v].e[a]);
                            chi[a] = std::fmax(chi[a], (double)m->tris[w.triOrig[t]].v[v].e[a]);
                        }
            }
            for (int a = 0; a < 3; a++) {
                const double step = std::ldexp(1.0, (int)n.e[a] - 127 - 7);
                const double gl = (s & 1) ? 0.0 : (double)(0x3F00 | n.qlo[a][s + 1]) / 65536.0, gh = (s & 1) ? 0.0 : (double)(0x3F00 | n.qhi[a][s + 1]) / 65536.0;
                const double qlo = (double)n.p[a] + ((n.qlo[a][s] & 127) + gl) * step, qhi = (double)n.p[a] + ((n.qhi[a][s] & 127) + gh) * step;
                if (qlo > clo[a] - w.pad[a] || qhi < chi[a] + w.pad[a]) return -6;
                blo[a] = std::fmin(blo[a], clo[a]);
                bhi[a] = std::fmax(bhi[a], chi[a]);
            }
        }
        for (int a = 0; a < 3; a++) { lo[3 * k + a] = blo[a]; hi[3 * k + a] = bhi[a]; }
    }
    return 0;
}
