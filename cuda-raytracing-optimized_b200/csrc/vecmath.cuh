// vecmath.cuh -- float3 arithmetic for the device kernels.
//
// Image parity with the reference kernel (SURVEY.md 8d "parity gates") needs
// the same IEEE operation sequence, so every helper keeps the operand order and
// grouping of the reference's vec3.h (operator shapes at vec3.h:62-104, cross
// :100-104, length :33, unit_vector :194): a*b+c*d groupings decide where nvcc
// places FMAs.  Division and sqrt stay IEEE (`/`, sqrtf; no fast-math), as in
// the reference build (vcxproj: no -use_fast_math).
#pragma once

#include <cuda_runtime.h>

struct f3 {
    float x, y, z;
};

__host__ __device__ __forceinline__ f3 mk3(float x, float y, float z) {
    f3 r;
    r.x = x; r.y = y; r.z = z;
    return r;
}
__device__ __forceinline__ f3 operator+(const f3& a, const f3& b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(const f3& a, const f3& b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator-(const f3& a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ f3 operator*(const f3& a, const f3& b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 operator*(float t, const f3& v) { return mk3(t * v.x, t * v.y, t * v.z); }
__device__ __forceinline__ f3 operator*(const f3& v, float t) { return mk3(t * v.x, t * v.y, t * v.z); }
__device__ __forceinline__ f3 operator/(const f3& v, float t) { return mk3(v.x / t, v.y / t, v.z / t); }
__device__ __forceinline__ float dot(const f3& a, const f3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross(const f3& a, const f3& b) {
    return mk3((a.y * b.z - a.z * b.y), (-(a.x * b.z - a.z * b.x)), (a.x * b.y - a.y * b.x));
}
__device__ __forceinline__ float sqlen(const f3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ float length(const f3& a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ f3 unit(const f3& v) { return v / length(v); }
__device__ __forceinline__ float maxcomp(const f3& v) { return fmaxf(v.x, fmaxf(v.y, v.z)); }

__device__ __forceinline__ f3 xyz(const float4& v) { return mk3(v.x, v.y, v.z); }
__device__ __forceinline__ float4 mk4(const f3& v, float w) { return make_float4(v.x, v.y, v.z, w); }
