// bvh_builder.h -- see bvh_builder.cpp.
#pragma once

#include <vector>

#include "rt_types.h"

namespace crt {

struct BuiltMesh {
    std::vector<triangle> tris;  // numLeaves * primsPerLeaf slots, +inf-padded
    std::vector<bvh_node> nodes; // 2 * numLeaves entries, [0] unused
    bbox bounds;
    int primsPerLeaf = 0;
    int numRealTris = 0;
};

// MEDIAN: the reference author's rule (median of the centroids along the longest axis). SAH: every split is the sweep
// position and axis of least surface-area-heuristic cost among those the complete tree still has room for (SURVEY 8f rank 1,
// TODO.txt:574,590-596): same BVH_00.04 layout, same traversal, same hits -- fewer node visits.
enum BuildMode { BUILD_MEDIAN = 0, BUILD_SAH = 1 };

bool buildBvh(const std::vector<triangle>& tris, int primsPerLeaf, BuiltMesh& out, BuildMode mode = BUILD_MEDIAN);
bool saveBvhFile(const char* path, const BuiltMesh& m);
bool loadBvhFile(const char* path, BuiltMesh& m);

} // namespace crt
