// wide_bvh.cpp -- host builder of the 8-wide quantised BVH (format and rationale: wide_bvh.h).
//
//   1. binned surface-area-heuristic binary build (16 bins x 3 axes on triangle centroids), leaves of <= 3 triangles.
//      The top of the tree (nodes with many triangles) is built by the calling thread with parallel binning / partition
//      passes; the subtrees below are independent tasks for a small spin-waiting thread pool;
//   2. greedy collapse to 8-wide nodes (always open the inner child with the largest surface area);
//   3. octant slot assignment, breadth-first emission (inner children of a node contiguous, leaf triangles contiguous),
//      7-bit outward quantisation of the padded child boxes.
// It runs inside initRenderer, so its wall time is part of the end-to-end frame: ~10-20 ms for the 311 740-triangle
// benchmark mesh on 8+ host cores (WideBvhStats.msTotal; CRT_TIMING=1 prints it).
#include "wide_bvh.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <condition_variable>
#include <limits>
#include <mutex>
#include <thread>

namespace {

constexpr int kBins = 16;
constexpr float kInf = std::numeric_limits<float>::infinity();
// SAH constants: cost of opening one more binary node relative to one triangle test (tuned on ray batches of the
// benchmark scene with the CPU walker, oracle/wide_walk.cpp)
constexpr float kNodeCost = 1.0f;
constexpr float kTriCost = 1.0f;
// cost model of the collapse to 8-wide nodes: one wide-node visit against one triangle test
static float kWideNode = 1.0f, kWideTri = 0.3f;

struct Box {
    float lo[3], hi[3];
};
inline void boxInit(Box& b) {
    for (int a = 0; a < 3; a++) { b.lo[a] = kInf; b.hi[a] = -kInf; }
}
inline void boxGrow(Box& b, const Box& o) {
    for (int a = 0; a < 3; a++) { b.lo[a] = std::min(b.lo[a], o.lo[a]); b.hi[a] = std::max(b.hi[a], o.hi[a]); }
}
inline void boxGrowPoint(Box& b, const float* p) {
    for (int a = 0; a < 3; a++) { b.lo[a] = std::min(b.lo[a], p[a]); b.hi[a] = std::max(b.hi[a], p[a]); }
}
inline float halfArea(const Box& b) {
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    if (!(dx >= 0.0f && dy >= 0.0f && dz >= 0.0f)) return 0.0f; // empty
    return dx * dy + dy * dz + dz * dx;
}

struct Prim {
    Box b;
    float c[3];
    uint32_t id;
};

struct Node2 {
    Box b;
    uint32_t left;   // inner: children at left, left + 1
    uint32_t first;  // prims[first, first + span) lie below this node
    uint32_t count;  // leaf: number of triangles; 0 = inner
    uint32_t span;
};

struct Bins {
    Box box[kBins];
    uint32_t cnt[kBins];
    void clear() {
        for (int k = 0; k < kBins; k++) { boxInit(box[k]); cnt[k] = 0; }
    }
    void merge(const Bins& o) {
        for (int k = 0; k < kBins; k++) { boxGrow(box[k], o.box[k]); cnt[k] += o.cnt[k]; }
    }
};

// A small pool for the build's parallel passes. Workers spin briefly for the next job (the passes follow each other within
// microseconds) and then sleep on a condition variable, so an oversubscribed host does not burn its cores on polling.
struct Pool {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable wake, finished;
    std::atomic<uint32_t> generation{0}, next{0};
    uint32_t jobCount = 0, done = 0, leftDrain = 0;
    bool quit = false;
    std::function<void(uint32_t)> job;

    explicit Pool(int threads) {
        for (int t = 1; t < threads; t++)
            workers.emplace_back([this] {
                uint32_t seen = 0;
                while (true) {
                    for (int spin = 0; spin < 2000 && generation.load(std::memory_order_acquire) == seen; spin++) {}
                    {
                        std::unique_lock<std::mutex> lk(m);
                        wake.wait(lk, [&] { return quit || generation.load(std::memory_order_relaxed) != seen; });
                        if (quit) return;
                        seen = generation.load(std::memory_order_relaxed);
                    }
                    const uint32_t did = drain();
                    std::lock_guard<std::mutex> lk(m);
                    done += did;
                    leftDrain++;
                    if (done >= jobCount && leftDrain == workers.size()) finished.notify_one();
                }
            });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(m);
            quit = true;
        }
        wake.notify_all();
        for (auto& w : workers) w.join();
    }
    uint32_t drain() {
        uint32_t did = 0;
        while (true) {
            const uint32_t i = next.fetch_add(1, std::memory_order_acq_rel);
            if (i >= jobCount) break;
            job(i);
            did++;
        }
        return did;
    }
    // Runs fn(0..count-1) on the pool and the calling thread; returns when all are done and every worker is idle again.
    void run(uint32_t count, std::function<void(uint32_t)> fn) {
        if (count == 0) return;
        if (workers.empty() || count == 1) {
            for (uint32_t i = 0; i < count; i++) fn(i);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m);
            job = std::move(fn);
            jobCount = count;
            done = 0;
            leftDrain = 0;
            next.store(0);
            generation.fetch_add(1, std::memory_order_release);
        }
        wake.notify_all();
        const uint32_t did = drain();
        std::unique_lock<std::mutex> lk(m);
        done += did;
        finished.wait(lk, [&] { return done >= jobCount && leftDrain == workers.size(); });
    }
    int size() const { return (int)workers.size() + 1; }
};

struct Split {
    int axis = -1;
    int plane = 0;      // prims with bin <= plane go left
    float scale = 0.0f; // bin = min(kBins - 1, (c - cmin) * scale)
    float cmin = 0.0f;
    Box lbox, rbox, lcbox, rcbox;
    uint32_t lcount = 0;
    float cost = kInf;
};

inline int binOf(float c, float cmin, float scale) {
    const int k = (int)((c - cmin) * scale);
    return k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
}

struct Builder {
    std::vector<Prim> prims, scratch;
    std::vector<Node2> nodes;
    std::atomic<uint32_t> nextNode{1};
    Pool* pool = nullptr;
    uint32_t parallelAbove = 0;

    struct Task {
        uint32_t node, lo, hi;
        Box cbox;
    };
    std::vector<Task> tasks;

    void binRange(uint32_t lo, uint32_t hi, int a, float cmin, float scale, Bins& bins) const {
        for (uint32_t i = lo; i < hi; i++) {
            const Prim& p = prims[i];
            const int k = binOf(p.c[a], cmin, scale);
            boxGrow(bins.box[k], p.b);
            bins.cnt[k]++;
        }
    }

    // Binned SAH along the axis on which the centroids spread most (one pass over the range; the children's boxes and
    // centroid bounds are gathered by the partition pass).
    Split findSplit(uint32_t lo, uint32_t hi, const Box& cbox, bool parallel) {
        Split best;
        int a = 0;
        float ext = cbox.hi[0] - cbox.lo[0];
        for (int k = 1; k < 3; k++)
            if (cbox.hi[k] - cbox.lo[k] > ext) { ext = cbox.hi[k] - cbox.lo[k]; a = k; }
        const float scale = ext > 0.0f ? (float)kBins * 0.999f / ext : 0.0f;
        if (!(scale > 0.0f) || !std::isfinite(scale)) return best;
        Bins bins;
        bins.clear();
        const uint32_t n = hi - lo;
        if (parallel) {
            const uint32_t chunks = (uint32_t)pool->size() * 2;
            std::vector<Bins> local(chunks);
            pool->run(chunks, [&](uint32_t k) {
                local[k].clear();
                binRange(lo + (uint32_t)((uint64_t)n * k / chunks), lo + (uint32_t)((uint64_t)n * (k + 1) / chunks), a, cbox.lo[a], scale, local[k]);
            });
            for (auto& l : local) bins.merge(l);
        } else {
            binRange(lo, hi, a, cbox.lo[a], scale, bins);
        }
        float rarea[kBins];
        uint32_t rcnt[kBins];
        Box acc;
        boxInit(acc);
        uint32_t c = 0;
        for (int k = kBins - 1; k > 0; k--) {
            boxGrow(acc, bins.box[k]);
            c += bins.cnt[k];
            rarea[k] = halfArea(acc);
            rcnt[k] = c;
        }
        boxInit(acc);
        c = 0;
        for (int k = 0; k < kBins - 1; k++) {
            boxGrow(acc, bins.box[k]);
            c += bins.cnt[k];
            if (c == 0 || c == n) continue;
            const float cost = halfArea(acc) * (float)c + rarea[k + 1] * (float)rcnt[k + 1];
            if (cost < best.cost) {
                best.cost = cost;
                best.axis = a;
                best.plane = k;
            }
        }
        if (best.axis < 0) return best;
        best.scale = scale;
        best.cmin = cbox.lo[a];
        boxInit(best.lbox); boxInit(best.rbox);
        for (int k = 0; k < kBins; k++) {
            if (k <= best.plane) { boxGrow(best.lbox, bins.box[k]); best.lcount += bins.cnt[k]; }
            else boxGrow(best.rbox, bins.box[k]);
        }
        return best;
    }

    // Small ranges: exact sweep over sorted orders instead of binning (clearing the bins costs more than the range): all
    // three axes up to 4 triangles, the axis of largest centroid spread above. Leaves prims[lo, hi) sorted along the chosen
    // axis; the split is after `s.lcount` of them (s.plane = -1 marks this form).
    static constexpr uint32_t kSmall = 12;
    Split findSplitSmall(uint32_t lo, uint32_t hi, const Box& cbox) {
        const uint32_t n = hi - lo;
        Split best;
        uint8_t order[kSmall], bestOrder[kSmall];
        float larea[kSmall];
        int bestK = 0;
        int axisLo = 0, axisHi = 3;
        if (n > 4) {
            int a = 0;
            float ext = cbox.hi[0] - cbox.lo[0];
            for (int k = 1; k < 3; k++)
                if (cbox.hi[k] - cbox.lo[k] > ext) { ext = cbox.hi[k] - cbox.lo[k]; a = k; }
            axisLo = a; axisHi = a + 1;
        }
        for (int a = axisLo; a < axisHi; a++) {
            for (uint32_t i = 0; i < n; i++) { // insertion sort by centroid (ties: triangle id, so the tree does not depend on the input order)
                const Prim& p = prims[lo + i];
                uint32_t j = i;
                while (j > 0) {
                    const Prim& q = prims[lo + order[j - 1]];
                    if (!(q.c[a] > p.c[a] || (q.c[a] == p.c[a] && q.id > p.id))) break;
                    order[j] = order[j - 1];
                    j--;
                }
                order[j] = (uint8_t)i;
            }
            Box acc;
            boxInit(acc);
            for (uint32_t i = 0; i + 1 < n; i++) { boxGrow(acc, prims[lo + order[i]].b); larea[i] = halfArea(acc); }
            boxInit(acc);
            bool better = false;
            for (uint32_t i = n - 1; i > 0; i--) {
                boxGrow(acc, prims[lo + order[i]].b);
                const float cost = larea[i - 1] * (float)i + halfArea(acc) * (float)(n - i);
                if (cost < best.cost) { best.cost = cost; best.axis = a; bestK = (int)i; better = true; }
            }
            if (better) std::memcpy(bestOrder, order, n);
        }
        if (best.axis < 0) return best;
        Prim tmp[kSmall];
        for (uint32_t i = 0; i < n; i++) tmp[i] = prims[lo + bestOrder[i]];
        for (uint32_t i = 0; i < n; i++) prims[lo + i] = tmp[i];
        best.plane = -1;
        best.lcount = (uint32_t)bestK;
        boxInit(best.lbox); boxInit(best.rbox); boxInit(best.lcbox); boxInit(best.rcbox);
        for (uint32_t i = 0; i < n; i++) {
            if (i < best.lcount) { boxGrow(best.lbox, tmp[i].b); boxGrowPoint(best.lcbox, tmp[i].c); }
            else { boxGrow(best.rbox, tmp[i].b); boxGrowPoint(best.rcbox, tmp[i].c); }
        }
        return best;
    }

    // Partitions prims[lo, hi) by the split and gathers the centroid bounds of both sides (s.lcbox / s.rcbox).
    uint32_t partitionRange(uint32_t lo, uint32_t hi, Split& s, bool parallel) {
        if (s.plane < 0) return lo + s.lcount; // findSplitSmall already ordered the range
        const int a = s.axis;
        auto goesLeft = [&](const Prim& p) { return binOf(p.c[a], s.cmin, s.scale) <= s.plane; };
        boxInit(s.lcbox); boxInit(s.rcbox);
        if (!parallel) {
            uint32_t i = lo, j = hi;
            while (true) {
                while (i < j && goesLeft(prims[i])) { boxGrowPoint(s.lcbox, prims[i].c); i++; }
                while (i < j && !goesLeft(prims[j - 1])) { j--; boxGrowPoint(s.rcbox, prims[j].c); }
                if (i >= j) break;
                std::swap(prims[i], prims[j - 1]);
            }
            return i;
        }
        // out of place: count per chunk, prefix, scatter into scratch, copy back
        const uint32_t chunks = (uint32_t)pool->size() * 2, n = hi - lo;
        std::vector<uint32_t> lcount(chunks + 1, 0);
        std::vector<Box> lcb(chunks), rcb(chunks);
        auto bound = [&](uint32_t k) { return lo + (uint32_t)((uint64_t)n * k / chunks); };
        pool->run(chunks, [&](uint32_t k) {
            uint32_t c = 0;
            Box l, r;
            boxInit(l); boxInit(r);
            for (uint32_t i = bound(k); i < bound(k + 1); i++) {
                if (goesLeft(prims[i])) { c++; boxGrowPoint(l, prims[i].c); }
                else boxGrowPoint(r, prims[i].c);
            }
            lcount[k + 1] = c;
            lcb[k] = l; rcb[k] = r;
        });
        for (uint32_t k = 0; k < chunks; k++) { lcount[k + 1] += lcount[k]; boxGrow(s.lcbox, lcb[k]); boxGrow(s.rcbox, rcb[k]); }
        const uint32_t mid = lo + lcount[chunks];
        pool->run(chunks, [&](uint32_t k) {
            uint32_t l = lo + lcount[k], r = mid + (bound(k) - lo - lcount[k]);
            for (uint32_t i = bound(k); i < bound(k + 1); i++) {
                if (goesLeft(prims[i])) scratch[l++] = prims[i];
                else scratch[r++] = prims[i];
            }
        });
        pool->run(chunks, [&](uint32_t k) { std::memcpy(&prims[bound(k)], &scratch[bound(k)], (size_t)(bound(k + 1) - bound(k)) * sizeof(Prim)); });
        return mid;
    }

    void makeLeaf(uint32_t node, uint32_t lo, uint32_t hi) {
        nodes[node].left = 0;
        nodes[node].first = lo;
        nodes[node].count = hi - lo;
        nodes[node].span = hi - lo;
    }

    // nodes[node].b is set by the caller. `top`: called on the building thread with the pool idle (may run parallel passes
    // and defers small ranges as tasks); otherwise plain recursion.
    void build(uint32_t node, uint32_t lo, uint32_t hi, const Box& cbox, bool top) {
        const uint32_t n = hi - lo;
        if (n == 1) { makeLeaf(node, lo, hi); return; }
        if (top && n < parallelAbove) {
            tasks.push_back(Task{node, lo, hi, cbox});
            return;
        }
        const bool parallel = top && pool->size() > 1;
        Split s = n <= kSmall ? findSplitSmall(lo, hi, cbox) : findSplit(lo, hi, cbox, parallel);
        const float area = halfArea(nodes[node].b);
        if (n <= WIDE_MAX_LEAF_TRIS) {
            // leaf unless the split pays for the extra node
            const float leafCost = kTriCost * (float)n * area;
            if (s.axis < 0 || !(kNodeCost * area + kTriCost * s.cost < leafCost)) { makeLeaf(node, lo, hi); return; }
        }
        uint32_t mid;
        Box lbox, rbox, lcbox, rcbox;
        if (s.axis >= 0) {
            mid = partitionRange(lo, hi, s, parallel);
            lbox = s.lbox; rbox = s.rbox; lcbox = s.lcbox; rcbox = s.rcbox;
        } else {
            // all centroids coincide: split the range in the middle
            mid = lo + n / 2;
            boxInit(lbox); boxInit(rbox); boxInit(lcbox); boxInit(rcbox);
            for (uint32_t i = lo; i < mid; i++) { boxGrow(lbox, prims[i].b); boxGrowPoint(lcbox, prims[i].c); }
            for (uint32_t i = mid; i < hi; i++) { boxGrow(rbox, prims[i].b); boxGrowPoint(rcbox, prims[i].c); }
        }
        const uint32_t left = nextNode.fetch_add(2, std::memory_order_relaxed);
        nodes[node].left = left;
        nodes[node].first = lo;
        nodes[node].count = 0;
        nodes[node].span = n;
        nodes[left].b = lbox;
        nodes[left + 1].b = rbox;
        build(left, lo, mid, lcbox, top);
        build(left + 1, mid, hi, rcbox, top);
    }
};

inline uint8_t quantByte(int q) { return (uint8_t)(0x80 | (q < 0 ? 0 : (q > 127 ? 127 : q))); }

} // namespace

bool buildWideBvh(const triangle* tris, uint32_t numSlots, int threads, WideBvhHost& out) {
    if (getenv("WB_CP")) kWideTri = (float)atof(getenv("WB_CP"));
    const auto t0 = std::chrono::steady_clock::now();
    out.nodes.clear();
    out.triOrig.clear();
    out.stats = WideBvhStats();
    if (threads <= 0) {
        threads = (int)std::thread::hardware_concurrency();
        if (threads <= 0) threads = 1;
        if (threads > 16) threads = 16;
    }
    Builder B;
    uint32_t n = 0;
    Box rootBox, rootCbox;
    double sah = 0.0;
    std::chrono::steady_clock::time_point t1;
    {
    Pool pool(threads);
    B.pool = &pool;
    out.stats.threads = threads;

    // ---- primitives (parallel): real triangles only
    std::vector<uint32_t> real;
    real.reserve(numSlots);
    for (uint32_t i = 0; i < numSlots; i++)
        if (!std::isinf(tris[i].v[0].e[0])) real.push_back(i);
    n = (uint32_t)real.size();
    if (n == 0) return false;
    B.prims.resize(n);
    B.scratch.resize(n);
    const uint32_t chunks = (uint32_t)threads * 4;
    std::vector<Box> cb(chunks), bb(chunks);
    pool.run(chunks, [&](uint32_t k) {
        Box c, b;
        boxInit(c); boxInit(b);
        for (uint32_t i = (uint32_t)((uint64_t)n * k / chunks); i < (uint32_t)((uint64_t)n * (k + 1) / chunks); i++) {
            const triangle& t = tris[real[i]];
            Prim& p = B.prims[i];
            boxInit(p.b);
            for (int v = 0; v < 3; v++) boxGrowPoint(p.b, t.v[v].e);
            for (int a = 0; a < 3; a++) p.c[a] = 0.5f * (p.b.lo[a] + p.b.hi[a]);
            p.id = real[i];
            boxGrow(b, p.b);
            boxGrowPoint(c, p.c);
        }
        cb[k] = c; bb[k] = b;
    });
    boxInit(rootBox); boxInit(rootCbox);
    for (uint32_t k = 0; k < chunks; k++) { boxGrow(rootBox, bb[k]); boxGrow(rootCbox, cb[k]); }
    for (int a = 0; a < 3; a++) {
        out.range[a] = std::max(std::max(std::fabs(rootBox.lo[a]), std::fabs(rootBox.hi[a])), 1e-30f);
        out.pad[a] = out.range[a] * WIDE_PAD_SCALE;
    }

    // ---- binary SAH tree
    B.nodes.resize(2 * (size_t)n + 2);
    B.parallelAbove = std::max<uint32_t>(8192u, n / 32u);
    B.nodes[0].b = rootBox;
    B.build(0, 0, n, rootCbox, true);
    std::sort(B.tasks.begin(), B.tasks.end(), [](const Builder::Task& x, const Builder::Task& y) { return (x.hi - x.lo) > (y.hi - y.lo); });
    pool.run((uint32_t)B.tasks.size(), [&](uint32_t k) {
        const Builder::Task& t = B.tasks[k];
        B.build(t.node, t.lo, t.hi, t.cbox, false);
    });
    const uint32_t numNodes2 = B.nextNode.load();
    t1 = std::chrono::steady_clock::now();
    out.stats.numBinaryNodes = numNodes2;

    // ---- collapse + emit, level by level (breadth-first numbering: the inner children of a node are contiguous, in slot
    //      order, and so are the leaf triangles of a node). Per level: (1) every node picks its children and their slots in
    //      parallel, (2) a prefix sum hands out child and triangle indices, (3) every node is quantised and written in parallel.
    // Which binary nodes become the (up to 8) children of a wide node is decided by the dynamic programme of Ylitie, Karras
    // and Laine 2017 (section 4.1): C(n, i) = least SAH cost of the subtree of n represented by at most i wide-tree roots,
    //   C(n, 1) = min(leaf: area * triangles * kWideTri if <= 3 triangles, inner: D(n, 8) + area * kWideNode)
    //   C(n, i) = min(D(n, i), C(n, i-1)),   D(n, j) = min over k of C(left, k) + C(right, j - k)
    // evaluated bottom-up (children have larger indices than their parent); `dec` keeps the arg-mins. Compared with opening
    // the largest child greedily this fills the nodes near the leaves (2.9x fewer wide nodes on the benchmark mesh).
    const std::vector<Node2>& N2 = B.nodes;
    std::vector<float> dpCost(7 * (size_t)numNodes2);
    std::vector<uint8_t> dec(8 * (size_t)numNodes2); // [0] 1 = inner, 0 = leaf; [i-1], i = 2..7: k of D(n, i) or 0 = "as i-1"; [7]: k of D(n, 8)
    for (uint32_t node = numNodes2; node-- > 0;) {
        const Node2& nd = N2[node];
        float* C = &dpCost[7 * (size_t)node];
        uint8_t* D = &dec[8 * (size_t)node];
        const float area = halfArea(nd.b);
        const float leaf = nd.span <= WIDE_MAX_LEAF_TRIS ? area * (float)nd.span * kWideTri : kInf;
        if (nd.count) {
            for (int i = 0; i < 7; i++) { C[i] = leaf; D[i] = 0; }
            D[7] = 0;
            continue;
        }
        const float* L = &dpCost[7 * (size_t)nd.left];
        const float* R = L + 7;
        float dist[9];
        uint8_t kbest[9];
        for (int j = 2; j <= 8; j++) {
            dist[j] = kInf;
            kbest[j] = 1;
            for (int k = std::max(1, j - 7); k <= std::min(7, j - 1); k++) {
                const float v = L[k - 1] + R[j - k - 1];
                if (v < dist[j]) { dist[j] = v; kbest[j] = (uint8_t)k; }
            }
        }
        const float inner = dist[8] + area * kWideNode;
        D[7] = kbest[8];
        if (leaf <= inner) { C[0] = leaf; D[0] = 0; }
        else { C[0] = inner; D[0] = 1; }
        for (int i = 2; i <= 7; i++) {
            if (dist[i] < C[i - 2]) { C[i - 1] = dist[i]; D[i - 1] = kbest[i]; }
            else { C[i - 1] = C[i - 2]; D[i - 1] = 0; }
        }
    }
    const uint32_t kLeafBit = 0x80000000u; // Plan::child: the binary node is emitted as ONE leaf (it may be an inner binary node with <= 3 triangles)

    struct Plan {
        uint32_t node2;
        uint32_t child[8];      // binary node per SLOT (or ~0u), | kLeafBit
        uint32_t innerCount, triCount;
        uint32_t childBase, triBase;
    };
    std::vector<Plan> level(1), next;
    level[0].node2 = 0;
    out.nodes.clear();
    out.triOrig.resize(n);
    uint32_t triCursor = 0;
    int depth = 0;
    std::vector<double> sahPart;
    while (!level.empty()) {
        depth++;
        const uint32_t count = (uint32_t)level.size();
        const uint32_t grain = 64, jobs = (count + grain - 1) / grain;
        // (1) children and slots
        pool.run(jobs, [&](uint32_t j) {
            for (uint32_t i = j * grain; i < std::min(count, (j + 1) * grain); i++) {
                Plan& pl = level[i];
                uint32_t child[8];
                int nc = 0;
                if (N2[pl.node2].count) {
                    child[nc++] = pl.node2 | kLeafBit; // the whole tree is one leaf
                } else { // follow the arg-mins of D(node, 8)
                    struct Item { uint32_t node; int budget; };
                    Item stack[16];
                    int sp = 0;
                    const uint32_t l = N2[pl.node2].left;
                    const int k8 = dec[8 * (size_t)pl.node2 + 7];
                    stack[sp++] = Item{l + 1, 8 - k8};
                    stack[sp++] = Item{l, k8};
                    while (sp > 0) {
                        Item it = stack[--sp];
                        const uint8_t* D = &dec[8 * (size_t)it.node];
                        while (it.budget > 1 && D[it.budget - 1] == 0) it.budget--;
                        if (it.budget == 1) {
                            child[nc++] = D[0] ? it.node : (it.node | kLeafBit);
                        } else {
                            const int k = D[it.budget - 1];
                            const uint32_t cl = N2[it.node].left;
                            stack[sp++] = Item{cl + 1, it.budget - k};
                            stack[sp++] = Item{cl, k};
                        }
                    }
                }
                Box nb;
                boxInit(nb);
                for (int k = 0; k < nc; k++) boxGrow(nb, N2[child[k] & ~kLeafBit].b);
                // octant slot assignment: greedy on dot(child centre - node centre, slot direction)
                float score[8][8];
                for (int k = 0; k < nc; k++) {
                    float d[3];
                    for (int a = 0; a < 3; a++) d[a] = 0.5f * (N2[child[k] & ~kLeafBit].b.lo[a] + N2[child[k] & ~kLeafBit].b.hi[a]) - 0.5f * (nb.lo[a] + nb.hi[a]);
                    for (int s = 0; s < 8; s++) score[k][s] = ((s & 4) ? d[0] : -d[0]) + ((s & 2) ? d[1] : -d[1]) + ((s & 1) ? d[2] : -d[2]);
                }
                bool placed[8] = {false, false, false, false, false, false, false, false};
                for (int s = 0; s < 8; s++) pl.child[s] = ~0u;
                for (int round = 0; round < nc; round++) {
                    int bk = -1, bs = -1;
                    float bv = -kInf;
                    for (int k = 0; k < nc; k++) {
                        if (placed[k]) continue;
                        for (int s = 0; s < 8; s++)
                            if (pl.child[s] == ~0u && score[k][s] > bv) { bv = score[k][s]; bk = k; bs = s; }
                    }
                    placed[bk] = true;
                    pl.child[bs] = child[bk];
                }
                pl.innerCount = pl.triCount = 0;
                for (int s = 0; s < 8; s++)
                    if (pl.child[s] != ~0u) {
                        if (pl.child[s] & kLeafBit) pl.triCount += N2[pl.child[s] & ~kLeafBit].span;
                        else pl.innerCount++;
                    }
            }
        });
        // (2) indices
        const uint32_t levelStart = (uint32_t)out.nodes.size();
        uint32_t childCursor = levelStart + count;
        for (uint32_t i = 0; i < count; i++) {
            level[i].childBase = childCursor;
            level[i].triBase = triCursor;
            childCursor += level[i].innerCount;
            triCursor += level[i].triCount;
        }
        out.nodes.resize(levelStart + count);
        next.resize(childCursor - (levelStart + count));
        sahPart.assign(jobs, 0.0);
        // (3) quantise and write
        pool.run(jobs, [&](uint32_t j) {
            double sahLocal = 0.0;
            for (uint32_t i = j * grain; i < std::min(count, (j + 1) * grain); i++) {
                const Plan& pl = level[i];
                Box nb;
                boxInit(nb);
                for (int s = 0; s < 8; s++)
                    if (pl.child[s] != ~0u) boxGrow(nb, N2[pl.child[s] & ~kLeafBit].b);
                // grid. The traversal reads the plane byte of an EVEN slot together with the byte of the next slot as excess
                // mantissa (wide_traverse.cuh planePair): its plane lies (q + g) steps above the origin, 0.248 <= g < 0.25.
                // Half a step of slack below the lowest plane and above the highest keeps every q inside 0..127.
                WideNode w;
                std::memset(&w, 0, sizeof(w));
                double step[3];
                for (int a = 0; a < 3; a++) {
                    const double lo = (double)nb.lo[a] - (double)out.pad[a], hi = (double)nb.hi[a] + (double)out.pad[a];
                    int k = (int)std::ceil(std::log2(std::max(hi - lo, 1e-300) / 126.0));
                    if (k < -100) k = -100;
                    float p;
                    while (true) {
                        step[a] = std::ldexp(1.0, k);
                        p = (float)(lo - 0.5 * step[a]);
                        if ((double)p > lo - 0.3 * step[a]) p = std::nextafter(p, -kInf);
                        if ((lo - (double)p) / step[a] >= 0.3 && (hi - (double)p) / step[a] <= 126.7) break;
                        k++;
                    }
                    w.p[a] = p;
                    w.e[a] = (uint8_t)std::min(std::max(k + 7 + 127, 0), 255);
                    w.scale[a] = (float)std::ldexp(1.0, k + 7); // A = 128 * step / d
                }
                for (int s = 7; s >= 0; s--) { // odd slots before the even slot that reads them as excess mantissa
                    for (int a = 0; a < 3; a++) { w.qlo[a][s] = quantByte(127); w.qhi[a][s] = quantByte(0); } // empty slot: inverted box
                    if (pl.child[s] == ~0u) continue;
                    const Node2& c = N2[pl.child[s] & ~kLeafBit];
                    for (int a = 0; a < 3; a++) {
                        const double lo = (double)c.b.lo[a] - (double)out.pad[a], hi = (double)c.b.hi[a] + (double)out.pad[a];
                        const double inv = 1.0 / step[a];
                        // excess of this slot's planes in steps (exact: the neighbour's bytes are final)
                        const double gl = (s & 1) ? 0.0 : (double)(0x3F00 | w.qlo[a][s + 1]) / 65536.0;
                        const double gh = (s & 1) ? 0.0 : (double)(0x3F00 | w.qhi[a][s + 1]) / 65536.0;
                        int ql = (int)std::floor((lo - (double)w.p[a]) * inv - gl);
                        int qh = (int)std::ceil((hi - (double)w.p[a]) * inv - gh);
                        while ((double)w.p[a] + (ql + gl) * step[a] > lo) ql--;
                        while ((double)w.p[a] + (qh + gh) * step[a] < hi) qh++;
                        w.qlo[a][s] = quantByte(ql);
                        w.qhi[a][s] = quantByte(qh);
                    }
                }
                w.childBase = pl.childBase;
                w.triBase = pl.triBase;
                uint32_t triOffset = 0, inner = 0;
                for (int s = 0; s < 8; s++) { // emit in slot order
                    if (pl.child[s] == ~0u) continue;
                    const Node2& c = N2[pl.child[s] & ~kLeafBit];
                    if (pl.child[s] & kLeafBit) {
                        w.meta[s] = (uint8_t)((c.span << 5) | triOffset);
                        for (uint32_t t = 0; t < c.span; t++) out.triOrig[pl.triBase + triOffset + t] = B.prims[c.first + t].id;
                        triOffset += c.span;
                        sahLocal += (double)halfArea(c.b) * c.span * kWideTri;
                    } else {
                        w.imask |= (uint8_t)(1u << s);
                        w.meta[s] = (uint8_t)(0x20 | (24 + s));
                        next[pl.childBase - (levelStart + count) + inner].node2 = pl.child[s];
                        inner++;
                    }
                }
                sahLocal += (double)halfArea(nb) * kWideNode;
                out.nodes[levelStart + i] = w;
            }
            sahPart[j] = sahLocal;
        });
        for (double v : sahPart) sah += v;
        level.swap(next);
        next.clear();
    }
    out.stats.maxDepth = depth;
    out.triOrig.resize(triCursor);
    B.pool = nullptr;
    } // (the pool's workers spin while they wait: it lives only as long as the build)
    const auto t2 = std::chrono::steady_clock::now();
    out.stats.numNodes = (uint32_t)out.nodes.size();
    out.stats.numTris = (uint32_t)out.triOrig.size();
    out.stats.sahCost = halfArea(rootBox) > 0.0f ? sah / (double)halfArea(rootBox) : 0.0;
    out.stats.msBinary = std::chrono::duration<double, std::milli>(t1 - t0).count();
    out.stats.msCollapse = std::chrono::duration<double, std::milli>(t2 - t1).count();
    out.stats.msTotal = std::chrono::duration<double, std::milli>(t2 - t0).count();
    return true;
}
