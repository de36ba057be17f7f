// rt_types.h -- plain-data types that cross the C ABI of the render path.
//
// Every struct here is LAYOUT-IDENTICAL to the type of the same name in the
// reference's helper_structs.h / vec3.h, because `initRenderer` receives
// `kernel_scene` and `camera` BY VALUE and follows host pointers into `mesh`,
// `triangle`, `bvh_node`, `material` and `stexture` arrays that the caller
// built with the reference's own headers (reference main.cpp:78-94).
//   vec3          reference vec3.h:9-40            12 B, align 4
//   bbox          reference helper_structs.h:73    24 B
//   triangle      reference helper_structs.h:81    64 B (v@0, texCoords@36, meshID@60)
//   bvh_node      reference helper_structs.h:98    24 B (min xyz, max xyz)
//   mesh          reference helper_structs.h:112   56 B
//   material      reference helper_structs.h:133   24 B
//   stexture      reference helper_structs.h:140   16 B
//   plane/sphere  reference helper_structs.h:160,168
//   camera        reference helper_structs.h:191   88 B
//   kernel_scene  reference helper_structs.h:217   64 B
// The sizes/offsets are locked by the static_asserts at the bottom; the
// numbers are the ones measured on the reference headers.
//
// The types are deliberately dumb: no methods with arithmetic. All arithmetic
// on the hot path lives in csrc/ (device) and oracle/ (CPU checker) where the
// rounding sequence is pinned explicitly.
#pragma once

#include <cstddef>
#include <cstdint>

#ifdef __CUDACC__
#define RT_HD __host__ __device__
#else
#define RT_HD
#endif

struct vec3 {
    float e[3];

    RT_HD vec3() : e{0.0f, 0.0f, 0.0f} {}
    RT_HD vec3(float x, float y, float z) : e{x, y, z} {}
    RT_HD float x() const { return e[0]; }
    RT_HD float y() const { return e[1]; }
    RT_HD float z() const { return e[2]; }
    RT_HD float operator[](int i) const { return e[i]; }
    RT_HD float& operator[](int i) { return e[i]; }
};

struct bbox {
    vec3 min;
    vec3 max;
};

struct triangle {
    vec3 v[3];
    float texCoords[6];   // (s,t) for vertex 0, 1, 2
    unsigned char meshID; // index into kernel_scene::materials
};

struct bvh_node {
    vec3 a; // box min
    vec3 b; // box max
};

// NOTE: the reference's mesh owns tris/bvh and deletes them in its destructor;
// this mirror does not own anything (the ABI only ever reads through it).
struct mesh {
    triangle* tris;
    uint32_t numTris; // number of triangle SLOTS (leaf padding included)
    bvh_node* bvh;
    int numBvhNodes;  // power of two; slot 0 unused, root = 1
    bbox bounds;
};

enum material_type { DIFFUSE = 0, METAL = 1, GLASS = 2 };

struct material {
    material_type type;
    vec3 color;
    float param; // METAL: fuzz, GLASS: index of refraction
    int texId;   // -1 = use `color`
};

struct stexture {
    float* data; // width*height RGB float triples, row 0 first
    int width;
    int height;
};

struct plane {
    vec3 norm;
    vec3 point;
};

struct sphere {
    vec3 center;
    float radius;
};

struct camera {
    vec3 origin;
    vec3 lower_left_corner;
    vec3 horizontal;
    vec3 vertical;
    vec3 u, v, w;
    float lens_radius;
};

struct kernel_scene {
    mesh* m;
    plane floor; // carried for ABI compatibility; dead on the reference path (kernels.cu:341-344)
    material* materials;
    int numMaterials;
    stexture* textures;
    int numTextures;
    int numPrimitivesPerLeaf;
};

static_assert(sizeof(vec3) == 12 && alignof(vec3) == 4, "vec3 layout");
static_assert(sizeof(bbox) == 24, "bbox layout");
static_assert(sizeof(triangle) == 64, "triangle layout");
static_assert(offsetof(triangle, v) == 0 && offsetof(triangle, texCoords) == 36 &&
              offsetof(triangle, meshID) == 60, "triangle offsets");
static_assert(sizeof(bvh_node) == 24, "bvh_node layout");
static_assert(sizeof(mesh) == 56, "mesh layout");
static_assert(offsetof(mesh, tris) == 0 && offsetof(mesh, numTris) == 8 && offsetof(mesh, bvh) == 16 &&
              offsetof(mesh, numBvhNodes) == 24 && offsetof(mesh, bounds) == 28, "mesh offsets");
static_assert(sizeof(material) == 24, "material layout");
static_assert(offsetof(material, type) == 0 && offsetof(material, color) == 4 &&
              offsetof(material, param) == 16 && offsetof(material, texId) == 20, "material offsets");
static_assert(sizeof(stexture) == 16, "stexture layout");
static_assert(sizeof(plane) == 24 && sizeof(sphere) == 16, "plane/sphere layout");
static_assert(sizeof(camera) == 88, "camera layout");
static_assert(offsetof(camera, origin) == 0 && offsetof(camera, lower_left_corner) == 12 &&
              offsetof(camera, horizontal) == 24 && offsetof(camera, vertical) == 36 &&
              offsetof(camera, u) == 48 && offsetof(camera, v) == 60 && offsetof(camera, w) == 72 &&
              offsetof(camera, lens_radius) == 84, "camera offsets");
static_assert(sizeof(kernel_scene) == 64, "kernel_scene layout");
static_assert(offsetof(kernel_scene, m) == 0 && offsetof(kernel_scene, floor) == 8 &&
              offsetof(kernel_scene, materials) == 32 && offsetof(kernel_scene, numMaterials) == 40 &&
              offsetof(kernel_scene, textures) == 48 && offsetof(kernel_scene, numTextures) == 56 &&
              offsetof(kernel_scene, numPrimitivesPerLeaf) == 60, "kernel_scene offsets");
