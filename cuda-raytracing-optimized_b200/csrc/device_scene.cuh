// device_scene.cuh -- what initRenderer leaves in HBM, and the wavefront state.
//
// HBM layout (all arrays 256-byte aligned by cudaMalloc, every hot load 16 B):
//   nodes     float4[3 * numBvhNodes/2 ...]  the caller's bvh_node[] bytes; internal node i's two
//                                             children are the 48 bytes at float4 index 3i
//   triGeom   float4[3 * numSlots]           {v0.xyz,e1.x}{e1.yz,e2.xy}{e2.z,-,-,-}   (traversal)
//   triShade  float4[3 * numSlots]           {n.xyz, meshID}{tc0..tc3}{tc4,tc5,-,-}    (shade only)
//   materials float4[2 * numMaterials]       {color.rgb, param}{type, texId, -, -}
//   textures  float[ w*h*3 ] each            nearest-texel RGB exactly as uploaded
// Path state is structure-of-arrays, one float4 stream per group of fields that a kernel
// reads together (helper_structs.h:48-71 `path`, :16-36 `intersection`):
//   rayO  {origin.xyz, rng}          rayD {rayDir.xyz, flags}       (extend reads these two)
//   atten {attenuation.rgb, sample#} pcol {color.rgb, -}            (shade / accumulate)
//   hit   {t, u, v, triId}                                          (extend -> shade)
// Shadow rays are a compacted queue of three float4 per entry:
//   shO {origin.xyz, slot}           shD {shadowDir.xyz, lightDist}  shL {lightContribution.rgb, -}
#pragma once

#include <cuda_runtime.h>

#include "intersect.cuh"

#define PATH_FLAG_SPECULAR 0x100u
#define PATH_FLAG_INSIDE 0x200u
#define PATH_BOUNCE_MASK 0xFFu

struct DeviceMaterials {
    const float4* __restrict__ mats; // 2 per material
    const float* const* __restrict__ texData;
    const int* __restrict__ texWidth;
    const int* __restrict__ texHeight;
};

struct LightDesc { // RenderContext defaults, kernels.cu:93-94
    f3 center;
    float radius;
    f3 color;
};

struct CameraDev { // camera, helper_structs.h:191-215
    f3 origin, lowerLeft, horizontal, vertical, u, v, w;
    float lensRadius;
};

struct WfControl {
    unsigned int countActive;  // entries in the current extend queue
    unsigned int countNext;    // entries appended for the next iteration
    unsigned int countShadow;
    unsigned int countRegen;
    unsigned int cursorExtend; // dynamic fetch cursors
    unsigned int cursorShadow;
    unsigned int blocksDone;   // shadeSpheresKernel's last block advances the iteration
    unsigned int pad1;
    unsigned long long raysExtend;
    unsigned long long raysShadow;
    unsigned long long iterations;
    unsigned long long nodeVisits; // only in counting builds
    unsigned long long triTests;
};

struct WfState {
    float4* rayO;
    float4* rayD;
    float4* atten;
    float4* pcol;
    float4* hit;
    float4* shO;
    float4* shD;
    float4* shL;
    unsigned int* queueA;
    unsigned int* queueB;
    unsigned int* regen;
    float4* accum;  // per pixel, un-normalised radiance sum
    WfControl* ctl;
    unsigned int numSlots;
};
