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
// dot / cross / squared length are PINNED with intrinsics. For `a*b + c*d` and `a*b - c*d` C++ does not say which
// product is fused: nvcc decides per context (we saw length() contract as fma(y,y,x*x) in one kernel and fma(x,x,y*y) in
// another) and leaves `a*b - c*d` to ptxas, which decides per kernel. The forms below are the ones in the reference's
// compiled render kernel (read from SASS): second product rounded on its own, first product fused onto it, further terms
// fused on top. Intrinsics are never re-fused or re-associated by either compiler stage.
__device__ __forceinline__ float dot(const f3& a, const f3& b) {
    return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)));
}
__device__ __forceinline__ f3 cross(const f3& a, const f3& b) {
    return mk3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), -__fmaf_rn(a.x, b.z, -__fmul_rn(a.z, b.x)),
               __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float sqlen(const f3& a) { return dot(a, a); }
__device__ __forceinline__ float length(const f3& a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ f3 unit(const f3& v) { return v / length(v); }
__device__ __forceinline__ float maxcomp(const f3& v) { return fmaxf(v.x, fmaxf(v.y, v.z)); }

__device__ __forceinline__ f3 xyz(const float4& v) { return mk3(v.x, v.y, v.z); }
__device__ __forceinline__ float4 mk4(const f3& v, float w) { return make_float4(v.x, v.y, v.z, w); }

__device__ __forceinline__ f3 subPinned(const f3& a, const f3& b) {
    return mk3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z));
}
// a*p + b*q with the second product rounded on its own (the reference's compiled form)
__device__ __forceinline__ float mad2(float a, float p, float b, float q) { return __fmaf_rn(a, p, __fmul_rn(b, q)); }

// {(a - o) * inv, (b - o) * inv} with both operations IEEE round-to-nearest (the slab test's `(box - origin) * invD`,
// intersections.h:28-29): sm_100's packed FADD2 / FMUL2 do the two lanes in one issue slot each. `negO` is -o: a + (-o)
// and a - o are the same IEEE operation.
__device__ __forceinline__ float2 slabPair(float a, float b, float negO, float inv) {
    float2 r;
    asm("{\n\t"
        ".reg .b64 p, q, s;\n\t"
        "mov.b64 p, {%2, %3};\n\t"
        "mov.b64 q, {%4, %4};\n\t"
        "mov.b64 s, {%5, %5};\n\t"
        "add.rn.f32x2 p, p, q;\n\t"
        "mul.rn.f32x2 p, p, s;\n\t"
        "mov.b64 {%0, %1}, p;\n\t"
        "}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a), "f"(b), "f"(negO), "f"(inv));
    return r;
}
