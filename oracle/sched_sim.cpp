// sched_sim.cpp -- TEST/DESIGN INFRASTRUCTURE (not product code).
// Replays the per-ray step sequences logged by the CPU oracle (liboracle_steplog.so, ORACLE_STEPLOG) through models of
// warp scheduling policies for traceKernel and reports issue slots per ray and lane utilisation. Used to choose the
// in-warp scheduling of csrc/traverse.cuh before spending GPU time (see DESIGN.md "What the profiles changed").
//
//   sched_sim <steplog.bin> [shuffle=1]
//
// Step log: 0 = dual-node step, 1+k = leaf visit with k triangle tests, 254/255 = end of ray.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

struct Ray {
    uint32_t begin, end; // into steps
};
static std::vector<uint8_t> steps;
static std::vector<Ray> rays;

struct Cost {
    double node = 66, leafBase = 30, tri = 60, refill = 160, overheadRound = 6;
};

struct Result {
    double slots = 0;       // warp issue slots
    double laneSlots = 0;   // useful lane-slots (active lanes x slots)
    double nodeSlots = 0, nodeLane = 0, leafSlots = 0, leafLane = 0, refillSlots = 0;
    uint64_t raysDone = 0;
    void print(const char* name) const {
        printf("%-46s slots/ray %8.1f  util %5.1f%%  node util %5.1f%% (%4.1f%% of slots)  leaf util %5.1f%% (%4.1f%%)  refill %4.1f%%\n", name,
               slots / raysDone, 100.0 * laneSlots / (32.0 * slots), 100.0 * nodeLane / (32.0 * nodeSlots), 100.0 * nodeSlots / slots,
               100.0 * leafLane / (32.0 * leafSlots), 100.0 * leafSlots / slots, 100.0 * refillSlots / slots);
    }
};

// ---- policy A: one ray per lane, node quorum Q, refill when fewer than M lanes have work
static Result simA(const Cost& c, int Q, int M, size_t firstRay, size_t numRays) {
    Result r;
    size_t next = firstRay, last = firstRay + numRays;
    uint32_t pos[32], end[32];
    bool live[32] = {};
    while (true) {
        int nlive = 0;
        for (int l = 0; l < 32; l++) nlive += live[l];
        if (next < last && nlive < (M > 0 ? M : 1) + 0) {
            // refill
            for (int l = 0; l < 32 && next < last; l++)
                if (!live[l]) {
                    pos[l] = rays[next].begin; end[l] = rays[next].end; next++;
                    live[l] = pos[l] < end[l];
                    if (!live[l]) r.raysDone++;
                }
            r.slots += c.refill; r.refillSlots += c.refill;
            r.laneSlots += c.refill * (32 - nlive);
            continue;
        }
        if (nlive == 0) { if (next >= last) break; else continue; }
        const int minAct = next < last ? M : 1;
        // run until fewer than minAct have work
        while (true) {
            int work = 0;
            for (int l = 0; l < 32; l++) work += live[l];
            if (work < minAct || work == 0) break;
            // node phase
            while (true) {
                int nNode = 0, nLeaf = 0;
                for (int l = 0; l < 32; l++) if (live[l]) { if (steps[pos[l]] == 0) nNode++; else nLeaf++; }
                if (nNode == 0) break;
                if (nNode < Q && nLeaf) break;
                if (nNode + nLeaf < minAct) break;
                r.slots += c.node; r.nodeSlots += c.node; r.nodeLane += c.node * nNode; r.laneSlots += c.node * nNode;
                for (int l = 0; l < 32; l++) if (live[l] && steps[pos[l]] == 0) { pos[l]++; if (pos[l] >= end[l]) { live[l] = false; r.raysDone++; } }
            }
            // leaf phase
            int nLeaf = 0, maxk = 0, sumk = 0;
            for (int l = 0; l < 32; l++) if (live[l] && steps[pos[l]] != 0) { nLeaf++; int k = steps[pos[l]] - 1; maxk = std::max(maxk, k); sumk += k; }
            if (nLeaf) {
                const double s = c.leafBase + c.tri * maxk;
                r.slots += s; r.leafSlots += s;
                const double useful = c.leafBase * nLeaf + c.tri * sumk;
                r.leafLane += useful; r.laneSlots += useful;
                for (int l = 0; l < 32; l++) if (live[l] && steps[pos[l]] != 0) { pos[l]++; if (pos[l] >= end[l]) { live[l] = false; r.raysDone++; } }
            }
            r.slots += c.overheadRound;
        }
    }
    return r;
}

// ---- policy B: K rays per lane (a pool of 32K per warp); each round the warp runs the phase most lanes can join.
// triPerRound: the leaf phase tests ONE triangle per lane per round (cursor kept) instead of the whole leaf.
static Result simB(const Cost& c, int K, int refillThreshold, bool triPerRound, double nodeBias, size_t firstRay, size_t numRays) {
    Result r;
    size_t next = firstRay, last = firstRay + numRays;
    std::vector<uint32_t> pos(32 * K), end(32 * K), triLeft(32 * K, 0);
    std::vector<char> live(32 * K, 0);
    const double poolCost = 4; // picking the ray + LDS/STS of its state
    while (true) {
        int nodeReady = 0, leafReady = 0, emptyLanes = 0, empties = 0, nlive = 0;
        int pickN[32], pickL[32];
        for (int l = 0; l < 32; l++) {
            pickN[l] = pickL[l] = -1;
            bool hasEmpty = false;
            for (int k = 0; k < K; k++) {
                const int e = l + 32 * k;
                if (!live[e]) { hasEmpty = true; empties++; continue; }
                nlive++;
                if (steps[pos[e]] == 0) { if (pickN[l] < 0) pickN[l] = e; }
                else if (pickL[l] < 0) pickL[l] = e;
            }
            nodeReady += pickN[l] >= 0; leafReady += pickL[l] >= 0; emptyLanes += hasEmpty;
        }
        const bool canRefill = next < last;
        if (nlive == 0 && !canRefill) break;
        if (canRefill && (emptyLanes >= refillThreshold || nlive == 0)) {
            // refill one empty entry per lane
            int filled = 0;
            for (int l = 0; l < 32 && next < last; l++)
                for (int k = 0; k < K; k++) {
                    const int e = l + 32 * k;
                    if (!live[e]) {
                        pos[e] = rays[next].begin; end[e] = rays[next].end; next++;
                        live[e] = pos[e] < end[e];
                        if (!live[e]) r.raysDone++;
                        filled++;
                        break;
                    }
                }
            r.slots += c.refill; r.refillSlots += c.refill; r.laneSlots += c.refill * filled;
            continue;
        }
        if (nodeReady > 0 && (nodeBias >= 2.0 ? (nodeReady >= (int)nodeBias || leafReady == 0) : nodeReady * nodeBias >= leafReady)) {
            const double s = c.node + poolCost;
            r.slots += s; r.nodeSlots += s; r.nodeLane += s * nodeReady; r.laneSlots += s * nodeReady;
            for (int l = 0; l < 32; l++) if (pickN[l] >= 0) { const int e = pickN[l]; pos[e]++; if (pos[e] >= end[e]) { live[e] = 0; r.raysDone++; } }
        } else if (leafReady > 0) {
            if (triPerRound) {
                const double s = c.tri + poolCost + 6;
                int act = 0;
                for (int l = 0; l < 32; l++) if (pickL[l] >= 0) {
                    const int e = pickL[l];
                    if (triLeft[e] == 0) triLeft[e] = steps[pos[e]] - 1;
                    if (triLeft[e] > 0) { triLeft[e]--; act++; }
                    if (triLeft[e] == 0) { pos[e]++; if (pos[e] >= end[e]) { live[e] = 0; r.raysDone++; } }
                }
                r.slots += s; r.leafSlots += s; r.leafLane += s * act; r.laneSlots += s * act;
            } else {
                int maxk = 0, sumk = 0;
                for (int l = 0; l < 32; l++) if (pickL[l] >= 0) { const int k = steps[pos[pickL[l]]] - 1; maxk = std::max(maxk, k); sumk += k; }
                const double s = c.leafBase + poolCost + c.tri * maxk;
                const double useful = (c.leafBase + poolCost) * leafReady + c.tri * sumk;
                r.slots += s; r.leafSlots += s; r.leafLane += useful; r.laneSlots += useful;
                for (int l = 0; l < 32; l++) if (pickL[l] >= 0) { const int e = pickL[l]; pos[e]++; if (pos[e] >= end[e]) { live[e] = 0; r.raysDone++; } }
            }
        }
        r.slots += c.overheadRound;
    }
    return r;
}

template <class F>
static Result overWarps(F f, size_t raysPerWarp) {
    Result total;
    for (size_t first = 0; first + raysPerWarp <= rays.size(); first += raysPerWarp) {
        const Result r = f(first, raysPerWarp);
        total.slots += r.slots; total.laneSlots += r.laneSlots; total.nodeSlots += r.nodeSlots; total.nodeLane += r.nodeLane;
        total.leafSlots += r.leafSlots; total.leafLane += r.leafLane; total.refillSlots += r.refillSlots; total.raysDone += r.raysDone;
    }
    return total;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: sched_sim steplog.bin [shuffle]\n"); return 1; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 1; }
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    steps.resize(n);
    if (fread(steps.data(), 1, n, f) != (size_t)n) return 1;
    fclose(f);
    uint32_t b = 0;
    uint64_t nodeSteps = 0, leafSteps = 0, tris = 0;
    for (uint32_t i = 0; i < (uint32_t)n; i++) {
        if (steps[i] >= 254) { rays.push_back({b, i}); b = i + 1; }
        else if (steps[i] == 0) nodeSteps++;
        else { leafSteps++; tris += steps[i] - 1; }
    }
    printf("%zu rays: %.1f node steps, %.2f leaf visits, %.2f triangle tests per ray\n", rays.size(), (double)nodeSteps / rays.size(),
           (double)leafSteps / rays.size(), (double)tris / rays.size());
    const int shuffle = argc > 2 ? atoi(argv[2]) : 1;
    if (shuffle == 1) { std::mt19937 g(1); std::shuffle(rays.begin(), rays.end(), g); }
    else if (shuffle > 1) { // shuffle inside windows (queue order keeps some coherence)
        std::mt19937 g(1);
        for (size_t i = 0; i + shuffle <= rays.size(); i += shuffle) std::shuffle(rays.begin() + i, rays.begin() + i + shuffle, g);
    }
    Cost c;
    if (getenv("C_NODE")) c.node = atof(getenv("C_NODE"));
    if (getenv("C_TRI")) c.tri = atof(getenv("C_TRI"));
    if (getenv("C_REFILL")) c.refill = atof(getenv("C_REFILL"));
    const double ideal = c.node * nodeSteps / rays.size() + c.leafBase * leafSteps / rays.size() + c.tri * tris / rays.size();
    printf("cost model: node %.0f, leaf %.0f + %.0f/tri, refill %.0f;  ideal (100%% lanes) = %.1f slots/ray/32\n", c.node, c.leafBase, c.tri, c.refill, ideal / 32);
    const size_t rpw = 2048; // rays a persistent warp processes
    char name[128];
    for (int Q : {32, 16, 8}) for (int M : {1, 12, 20, 26}) {
        snprintf(name, sizeof name, "A: 1 ray/lane quorum %d minActive %d", Q, M);
        overWarps([&](size_t a, size_t b2) { return simA(c, Q, M, a, b2); }, rpw).print(name);
    }
    for (int K : {2, 3, 4}) for (int thr : {16}) for (int tpr = 0; tpr < 2; tpr++) for (double bias : {1.0, 20.0, 24.0, 28.0, 30.0}) {
        snprintf(name, sizeof name, "B: %d rays/lane refill>=%d %s bias %.1f", K, thr, tpr ? "tri/round" : "leaf/round", bias);
        overWarps([&](size_t a, size_t b2) { return simB(c, K, thr, tpr != 0, bias, a, b2); }, rpw).print(name);
    }
    return 0;
}
