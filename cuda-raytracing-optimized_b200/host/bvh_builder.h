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

bool buildBvh(const std::vector<triangle>& tris, int primsPerLeaf, BuiltMesh& out);
bool saveBvhFile(const char* path, const BuiltMesh& m);
bool loadBvhFile(const char* path, BuiltMesh& m);

} // namespace crt
