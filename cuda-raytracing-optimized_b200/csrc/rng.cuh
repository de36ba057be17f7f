// rng.cuh -- the reference's custom RNG (rnd.h), integer-exact.
//   wang_hash        rnd.h:31-39
//   xor_shift_32     rnd.h:5-13   (shifts 13, 17, 15)
//   rnd              rnd.h:15-18  (low 24 bits / 2^24, exact int->float)
//   seed             kernels.cu:542  (wang_hash(pixelId) * 336343633) | 1
//   unit disk/sphere rnd.h:20-27, 41-49
// The reference builds sample vectors as vec3(rnd(s), rnd(s), rnd(s)), whose
// evaluation order C++ leaves open; its nvcc device build draws x first, then
// y, then z.  Here every draw is its own statement.
#pragma once

#include "vecmath.cuh"

__device__ __forceinline__ unsigned int wangHash(unsigned int seed) {
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}

__device__ __forceinline__ unsigned int pathSeed(unsigned int id) { return (wangHash(id) * 336343633u) | 1u; }

__device__ __forceinline__ float rnd(unsigned int& state) {
    unsigned int x = state;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 15;
    state = x;
    return (float)(x & 0xFFFFFFu) / 16777216.0f;
}

__device__ __forceinline__ f3 randomInUnitDisk(unsigned int& state) {
    f3 p;
    do {
        float a = rnd(state);
        float b = rnd(state);
        p = 2.0f * mk3(a, b, 0.0f) - mk3(1.0f, 1.0f, 0.0f);
    } while (dot(p, p) >= 1.0f);
    return p;
}

__device__ __forceinline__ f3 randomInUnitSphere(unsigned int& state) {
    f3 p;
    do {
        float a = rnd(state);
        float b = rnd(state);
        float c = rnd(state);
        p = 2.0f * mk3(a, b, c) - mk3(1.0f, 1.0f, 1.0f);
    } while (sqlen(p) >= 1.0f);
    return p;
}
