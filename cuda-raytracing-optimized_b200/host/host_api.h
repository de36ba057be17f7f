// host_api.h -- C ABI of libcrt_host.so: the host side of the render path.
//
// The reference keeps scene/camera setup on the host (main.cpp:62-139,
// staircase_scene.h:62-184, camera ctor helper_structs.h:194-206) and hands
// plain structs to the device library.  This library is that host side, rebuilt:
//   * it GENERATES its inputs (the reference's assets are not shipped:
//     staircase_scene.h:122,162 point at C:\Users\...), i.e. a procedural
//     staircase-class triangle mesh, 9 procedural textures and the 20-entry
//     material table of staircase_scene.h:141-160;
//   * it builds the complete-binary-tree BVH the reference traversal expects
//     (kernels.cu:154-224, firstLeafIdx = numBvhNodes/2) and reads/writes it in
//     the reference's BVH_00.04 container (staircase_scene.h:75-101);
//   * it restates the camera constructor, the RTIOW sphere scene driven by the
//     host LCG of main.cpp:17-20, and the PPM / REF_00.01 frame files.
// Everything is exposed with plain pointers so that C++, ctypes and the
// oracle's reference driver can share one scene bit for bit.
#pragma once

#include "rt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crt_scene crt_scene; // opaque; owns tris, bvh, materials, textures

// detail: 1.0 ~ 258k triangles (2^16 leaves x 5 slots, 17 levels). Smaller = coarser
// tessellation of the same rooms/objects (for tests). texSize: texture edge in texels.
crt_scene* crtSceneCreateStaircase(float detail, int texSize, int primsPerLeaf);
// The same scene with the BVH built by the surface-area heuristic inside the same complete-tree layout (buildMode 1;
// 0 = the reference author's median split, what crtSceneCreateStaircase does). Same file format, same hits.
crt_scene* crtSceneCreateStaircaseEx(float detail, int texSize, int primsPerLeaf, int buildMode);
// Mesh from a BVH_00.04 file + the procedural textures/material table.
crt_scene* crtSceneLoadBVH(const char* path, int texSize);
// Scene from caller triangles (copied); builds the BVH. materials/textures = staircase tables.
crt_scene* crtSceneFromTriangles(const triangle* tris, int n, int primsPerLeaf, int texSize);
// Textures from PNG files (loadTexture, staircase_scene.h:103-118: RGB, rows flipped, byte / 255.0f; decoder: host/png_reader.cpp).
// index 0..8 = WoodFloor, Wallpaper, Woodpanel, Painting1-3, WoodChair, Fabric, BrushedAluminium (staircase_scene.h:125-133).
int crtSceneLoadTexturePNG(crt_scene* s, int index, const char* path); // 0 on success, -1: unreadable / not a PNG (texture unchanged)
int crtSceneLoadTextureDir(crt_scene* s, const char* dir);             // the nine file names of load_scene; returns how many loaded
int crtSceneTextureInfo(const crt_scene* s, int index, int* width, int* height, const float** data);
// PNG image in memory -> width, height and (when rgbOut != NULL) width*height*3 bytes; -1: malformed, -2: capacity too small.
int crtDecodePNG(const unsigned char* bytes, unsigned long long n, int flipVertically, int* width, int* height, unsigned char* rgbOut,
                 unsigned long long capacity);
void crtSceneDestroy(crt_scene* s);
int crtSceneSaveBVH(const crt_scene* s, const char* path); // 0 on success
const kernel_scene* crtSceneKernelScene(const crt_scene* s);
int crtSceneNumRealTriangles(const crt_scene* s);          // without leaf padding
unsigned long long crtSceneHash(const crt_scene* s);       // FNV-1a over tris, nodes, materials, textures

// camera(lookfrom, lookat, vup, vfov, aspect, aperture, focus): helper_structs.h:194-206
void crtMakeCamera(const float lookfrom[3], const float lookat[3], const float vup[3], float vfov, float aspect,
                   float aperture, float focusDist, camera* out);
// setup_camera(nx, ny): staircase_scene.h:62-73
void crtStaircaseCamera(int nx, int ny, camera* out);

// RTIOW random-spheres scene (README.md:3-6 lineage): 1 ground + 22x22 small + 3 big = 488
// spheres, materials drawn with the LCG of main.cpp:17-20. Returns the count written (<= cap).
int crtRtiowScene(unsigned int seed, sphere* outSpheres, material* outMaterials, int cap);
void crtRtiowCamera(int nx, int ny, camera* out);

// Frame files. PPM: staircase_scene.h:22-43 (sRGB, rows top-down). REF_00.01: main.cpp:25-60.
int crtWritePPM(const char* path, int nx, int ny, const vec3* fb);
int crtWriteRef(const char* path, int nx, int ny, const vec3* fb);
int crtReadRef(const char* path, int nx, int ny, vec3* fb);
unsigned int crtLinearToSRGB(float x);
// sqrt(mean over pixels and channels of squared error): main.cpp:108-128
double crtRmse(const vec3* a, const vec3* b, int nx, int ny);

#ifdef __cplusplus
}
#endif
