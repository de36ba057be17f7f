// host_api.cpp -- host side of the render path (see host_api.h).
//
// What it restates from the reference (host-only code, nothing here runs on
// the device):
//   camera ctor              helper_structs.h:194-206
//   setup_camera             staircase_scene.h:62-73
//   material table           staircase_scene.h:141-160
//   LinearToSRGB / writePPM  staircase_scene.h:22-43
//   REF_00.01 save/load      main.cpp:25-60, RMSE main.cpp:108-128
//   host LCG random_float    main.cpp:17-20
// What it adds because the reference's assets do not ship
// (staircase_scene.h:122,162 are C:\ paths): a procedural staircase-class
// mesh and nine procedural textures with the same roles as the PNGs.
#include "host_api.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "bvh_builder.h"
#include "png_reader.h"

namespace {

const float kPi = 3.14159265358979323846f;

// ------------------------------------------------------------- vec helpers --
inline vec3 add(vec3 a, vec3 b) { return vec3(a.e[0] + b.e[0], a.e[1] + b.e[1], a.e[2] + b.e[2]); }
inline vec3 sub(vec3 a, vec3 b) { return vec3(a.e[0] - b.e[0], a.e[1] - b.e[1], a.e[2] - b.e[2]); }
inline vec3 mul(vec3 a, float s) { return vec3(a.e[0] * s, a.e[1] * s, a.e[2] * s); }
inline float dot(vec3 a, vec3 b) { return a.e[0] * b.e[0] + a.e[1] * b.e[1] + a.e[2] * b.e[2]; }
inline vec3 cross(vec3 a, vec3 b) {
    return vec3(a.e[1] * b.e[2] - a.e[2] * b.e[1], -(a.e[0] * b.e[2] - a.e[2] * b.e[0]), a.e[0] * b.e[1] - a.e[1] * b.e[0]);
}
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3 unit(vec3 a) {
    float l = length(a);
    return vec3(a.e[0] / l, a.e[1] / l, a.e[2] / l);
}

// --------------------------------------------------------- mesh generation --
// Material slots of staircase_scene.h:141-160.
enum Mat {
    M_BLACK = 0, M_BRASS, M_ALU, M_CANDLE, M_SEAT, M_GLASS, M_GOLD, M_SHADE, M_MAGNOLIA, M_PAINT1, M_PAINT2,
    M_PAINT3, M_STEEL, M_WALLPAPER, M_WHITEPAINT, M_PLASTIC, M_WOODCHAIR, M_FLOOR, M_WOODLAMP, M_STAIRS
};

struct MeshGen {
    std::vector<triangle> tris;
    float detail;

    int seg(float n) const { return std::max(3, (int)std::lround(n * detail)); }
    int cells(float n) const { return std::max(1, (int)std::lround(n * detail)); }

    void tri(vec3 a, vec3 b, vec3 c, float ua, float va, float ub, float vb, float uc, float vc, int m) {
        triangle t;
        t.v[0] = a; t.v[1] = b; t.v[2] = c;
        t.texCoords[0] = ua; t.texCoords[1] = va;
        t.texCoords[2] = ub; t.texCoords[3] = vb;
        t.texCoords[4] = uc; t.texCoords[5] = vc;
        t.meshID = (unsigned char)m;
        // texCoords end at byte 60, meshID at 60: keep the 3 tail bytes defined for hashing / file IO
        std::memset(((unsigned char*)&t) + 61, 0, 3);
        tris.push_back(t);
    }

    // parametric grid: p(u,v), u,v in [0,1]; nu x nv cells; texcoords = (u*su, v*sv)
    template <class F>
    void grid(F p, int nu, int nv, float su, float sv, int m, bool flip = false) {
        for (int j = 0; j < nv; j++)
            for (int i = 0; i < nu; i++) {
                float u0 = (float)i / nu, u1 = (float)(i + 1) / nu;
                float v0 = (float)j / nv, v1 = (float)(j + 1) / nv;
                vec3 a = p(u0, v0), b = p(u1, v0), c = p(u1, v1), d = p(u0, v1);
                if (!flip) {
                    tri(a, b, c, u0 * su, v0 * sv, u1 * su, v0 * sv, u1 * su, v1 * sv, m);
                    tri(a, c, d, u0 * su, v0 * sv, u1 * su, v1 * sv, u0 * su, v1 * sv, m);
                } else {
                    tri(a, c, b, u0 * su, v0 * sv, u1 * su, v1 * sv, u1 * su, v0 * sv, m);
                    tri(a, d, c, u0 * su, v0 * sv, u0 * su, v1 * sv, u1 * su, v1 * sv, m);
                }
            }
    }

    // planar quad o + u*eu + v*ev
    void quad(vec3 o, vec3 eu, vec3 ev, int nu, int nv, float su, float sv, int m) {
        grid([&](float u, float v) { return add(o, add(mul(eu, u), mul(ev, v))); }, nu, nv, su, sv, m);
    }

    // axis-aligned box, all six faces, `n` cells along the longest edge
    void box(vec3 lo, vec3 hi, float n, int m, float texScale = 1.0f) {
        vec3 d = sub(hi, lo);
        float longest = std::max(d.e[0], std::max(d.e[1], d.e[2]));
        auto c = [&](float len) { return std::max(1, (int)std::lround(n * detail * len / longest)); };
        int nx = c(d.e[0]), ny = c(d.e[1]), nz = c(d.e[2]);
        float s = texScale;
        quad(vec3(lo.e[0], lo.e[1], hi.e[2]), vec3(d.e[0], 0, 0), vec3(0, d.e[1], 0), nx, ny, s, s, m); // +z
        quad(vec3(hi.e[0], lo.e[1], lo.e[2]), vec3(-d.e[0], 0, 0), vec3(0, d.e[1], 0), nx, ny, s, s, m); // -z
        quad(vec3(hi.e[0], lo.e[1], hi.e[2]), vec3(0, 0, -d.e[2]), vec3(0, d.e[1], 0), nz, ny, s, s, m); // +x
        quad(vec3(lo.e[0], lo.e[1], lo.e[2]), vec3(0, 0, d.e[2]), vec3(0, d.e[1], 0), nz, ny, s, s, m);  // -x
        quad(vec3(lo.e[0], hi.e[1], hi.e[2]), vec3(d.e[0], 0, 0), vec3(0, 0, -d.e[2]), nx, nz, s, s, m); // +y
        quad(vec3(lo.e[0], lo.e[1], lo.e[2]), vec3(d.e[0], 0, 0), vec3(0, 0, d.e[2]), nx, nz, s, s, m);  // -y
    }

    void sphereMesh(vec3 c, float r, float nLong, int m, float squashY = 1.0f) {
        int nu = seg(nLong), nv = std::max(2, nu / 2);
        grid([&](float u, float v) {
            float phi = 2 * kPi * u, th = kPi * v;
            return vec3(c.e[0] + r * std::sin(th) * std::cos(phi), c.e[1] - r * squashY * std::cos(th),
                        c.e[2] + r * std::sin(th) * std::sin(phi));
        }, nu, nv, 2.0f, 1.0f, m, true);
    }

    // surface of revolution around +y through `c`: radius(v), height(v)
    template <class R, class H>
    void lathe(vec3 c, R radius, H height, float nAround, float nUp, int m, float su = 2.0f, float sv = 1.0f) {
        int nu = seg(nAround), nv = cells(nUp);
        grid([&](float u, float v) {
            float phi = 2 * kPi * u, r = radius(v);
            return vec3(c.e[0] + r * std::cos(phi), c.e[1] + height(v), c.e[2] + r * std::sin(phi));
        }, nu, nv, su, sv, m, true);
    }

    void cylinder(vec3 base, float r, float h, float nAround, float nUp, int m) {
        lathe(base, [=](float) { return r; }, [=](float v) { return h * v; }, nAround, nUp, m);
        // caps as thin cones to the axis
        lathe(base, [=](float v) { return r * (1 - v); }, [=](float) { return h; }, nAround, 1, m);
        lathe(base, [=](float v) { return r * v; }, [=](float) { return 0.0f; }, nAround, 1, m);
    }

    // cylinder between two arbitrary points (hand rails)
    void tube(vec3 a, vec3 b, float r, float nAround, float nAlong, int m) {
        vec3 w = unit(sub(b, a));
        vec3 up = std::fabs(w.e[1]) < 0.9f ? vec3(0, 1, 0) : vec3(1, 0, 0);
        vec3 u = unit(cross(up, w)), v = cross(w, u);
        float len = length(sub(b, a));
        int nu = seg(nAround), nv = cells(nAlong);
        grid([&](float s, float t) {
            float phi = 2 * kPi * s;
            vec3 ring = add(mul(u, r * std::cos(phi)), mul(v, r * std::sin(phi)));
            return add(add(a, mul(w, len * t)), ring);
        }, nu, nv, 2.0f, len / 100.0f, m);
    }

    void torus(vec3 c, float R, float r, float nMajor, float nMinor, int m) {
        int nu = seg(nMajor), nv = seg(nMinor);
        grid([&](float u, float v) {
            float a = 2 * kPi * u, b = 2 * kPi * v;
            float q = R + r * std::cos(b);
            return vec3(c.e[0] + q * std::cos(a), c.e[1] + r * std::sin(b), c.e[2] + q * std::sin(a));
        }, nu, nv, 4.0f, 1.0f, m);
    }
};

// The room is laid out around the reference camera (staircase_scene.h:63-64:
// eye (5.56,173.68,494.52) looking down -z) and the reference light
// (kernels.cu:93: centre (52.5,715.7,-272.6), radius 50), so both keep their
// hard-wired values: a hall 700 wide, 900 tall, 1300 deep, with the light
// hanging free under the ceiling above the upper landing.
void buildStaircase(MeshGen& g) {
    const float x0 = -340, x1 = 360, y0 = 0, y1 = 900, z0 = -740, z1 = 560;
    const float W = x1 - x0, H = y1 - y0, D = z1 - z0;

    // shell (inward-facing order does not matter: the path tracer flips normals, kernels.cu:354)
    g.quad(vec3(x0, y0, z0), vec3(W, 0, 0), vec3(0, 0, D), g.cells(70), g.cells(130), 7, 13, M_FLOOR);
    g.quad(vec3(x0, y1, z0), vec3(W, 0, 0), vec3(0, 0, D), g.cells(35), g.cells(65), 1, 1, M_WHITEPAINT);
    g.quad(vec3(x0, y0, z0), vec3(W, 0, 0), vec3(0, H, 0), g.cells(70), g.cells(90), 7, 9, M_WALLPAPER);  // back
    g.quad(vec3(x0, y0, z1), vec3(W, 0, 0), vec3(0, H, 0), g.cells(35), g.cells(45), 7, 9, M_MAGNOLIA);   // behind camera
    g.quad(vec3(x0, y0, z0), vec3(0, 0, D), vec3(0, H, 0), g.cells(104), g.cells(72), 13, 9, M_WALLPAPER); // left
    g.quad(vec3(x1, y0, z0), vec3(0, 0, D), vec3(0, H, 0), g.cells(104), g.cells(72), 13, 9, M_WALLPAPER); // right
    // wood panelling (dado) on the side walls, 2 units proud of the wall
    g.quad(vec3(x0 + 2, y0, z0), vec3(0, 0, D), vec3(0, 110, 0), g.cells(130), g.cells(11), 13, 1, M_STAIRS);
    g.quad(vec3(x1 - 2, y0, z0), vec3(0, 0, D), vec3(0, 110, 0), g.cells(130), g.cells(11), 13, 1, M_STAIRS);

    // the staircase: 16 steps rising towards -z on the left half, then a landing
    const int nSteps = 16;
    const float sx0 = -300, sx1 = -40, rise = 22, run = 34, zStart = 150;
    for (int s = 0; s < nSteps; s++) {
        float zs = zStart - run * s;
        // No two faces of the scene are coplanar AND overlapping (exact ties would make the closest hit depend on the
        // traversal order): boxes that touch either interpenetrate or keep a 0.2 gap, and nothing ends exactly on the floor.
        g.box(vec3(sx0, y0 - 0.5f, zs - run + 0.2f), vec3(sx1, rise * (s + 1), zs), 26, M_STAIRS, 2.0f);
        // nosing strip, sunk one unit into the tread
        g.box(vec3(sx0 + 0.3f, rise * (s + 1) - 1.0f, zs - 3), vec3(sx1 + 4, rise * (s + 1) + 2.5f, zs + 3), 26, M_WOODCHAIR, 2.0f);
    }
    const float landY = rise * nSteps, landZ1 = zStart - run * nSteps;
    g.box(vec3(sx0, landY - 12, z0 + 2), vec3(x1 - 60, landY, landZ1), 48, M_STAIRS, 4.0f); // landing slab
    // balusters + hand rail along the open side of the flight
    for (int s = 0; s < nSteps; s++) {
        float zs = zStart - run * (s + 0.5f);
        g.cylinder(vec3(sx1 - 8, rise * (s + 1) - 0.5f, zs), 2.2f, 78.5f, 16, 6, (s & 1) ? M_BRASS : M_BLACK);
    }
    g.tube(vec3(sx1 - 8, rise + 80, zStart - run * 0.5f), vec3(sx1 - 8, landY + 80, landZ1 - run * 0.5f + run), 4.5f, 20, 60,
           M_WOODCHAIR);
    // balustrade along the landing edge
    for (int k = 0; k < 18; k++) {
        float x = sx1 + 10 + k * 18.0f;
        g.cylinder(vec3(x, landY - 0.5f, landZ1 - 6), 2.2f, 78.5f, 16, 6, (k & 1) ? M_GOLD : M_BLACK);
    }
    g.tube(vec3(sx1 - 8, landY + 80, landZ1 - 6), vec3(sx1 + 10 + 17 * 18.0f + 10, landY + 80, landZ1 - 6), 4.5f, 20, 40,
           M_WOODCHAIR);
    g.sphereMesh(vec3(sx1 - 8, rise + 92, zStart - run * 0.5f), 9, 40, M_BRASS); // newel cap

    // paintings: frame (box) + canvas, 3 units off the wall
    struct P { float x, y, z, w, h; int m; int wall; };
    const P paintings[3] = {{x1 - 3, 330, 120, 190, 130, M_PAINT1, 1}, {x1 - 3, 330, -160, 150, 190, M_PAINT2, 1},
                            {60, 560, z0 + 3, 260, 170, M_PAINT3, 0}};
    for (const P& p : paintings) {
        if (p.wall == 1) { // on the right wall, facing -x
            g.box(vec3(p.x - 4, p.y - 8, p.z - p.w / 2 - 8), vec3(p.x, p.y + p.h + 8, p.z + p.w / 2 + 8), 12, M_BLACK);
            g.quad(vec3(p.x - 5, p.y, p.z + p.w / 2), vec3(0, 0, -p.w), vec3(0, p.h, 0), g.cells(24), g.cells(24), 1, 1, p.m);
        } else { // on the back wall, facing +z
            g.box(vec3(p.x - p.w / 2 - 8, p.y - 8, p.z), vec3(p.x + p.w / 2 + 8, p.y + p.h + 8, p.z + 4), 12, M_BLACK);
            g.quad(vec3(p.x - p.w / 2, p.y, p.z + 5), vec3(p.w, 0, 0), vec3(0, p.h, 0), g.cells(24), g.cells(24), 1, 1, p.m);
        }
    }

    // two chairs: seat, back, four legs
    for (int c = 0; c < 2; c++) {
        float cx = 150 + 110 * c, cz = -40 + 170 * c;
        g.box(vec3(cx - 24, 44, cz - 24), vec3(cx + 24, 50, cz + 24), 10, M_WOODCHAIR);
        g.box(vec3(cx - 22, 49.5f, cz - 22), vec3(cx + 22, 56, cz + 22), 10, M_SEAT);
        g.box(vec3(cx - 23.5f, 49, cz - 28), vec3(cx + 23.5f, 112, cz - 23), 12, M_WOODCHAIR);
        for (int l = 0; l < 4; l++)
            g.cylinder(vec3(cx + ((l & 1) ? 20.f : -20.f), -0.5f, cz + ((l & 2) ? 20.f : -20.f)), 2.6f, 45, 14, 4, M_WOODCHAIR);
    }

    // side table with glass vase, candles, steel and gold balls
    g.cylinder(vec3(170, -0.5f, 60), 5, 71, 24, 6, M_ALU);
    g.cylinder(vec3(170, 70, 60), 46, 4, 64, 1, M_ALU);
    g.lathe(vec3(170, 74.5f, 60), [](float v) { return 7 + 9 * std::sin(2.6f * v) + 3 * v; }, [](float v) { return 58 * v; }, 96, 64,
            M_GLASS); // open vase (single sheet: a thin glass shell)
    g.sphereMesh(vec3(196, 86, 48), 11.5f, 96, M_STEEL);
    g.sphereMesh(vec3(148, 84, 80), 9.5f, 96, M_GOLD);
    for (int k = 0; k < 3; k++) g.cylinder(vec3(156.f + 9 * k, 74.5f, 36.f + 5 * k), 2.4f, 18.f + 5 * k, 20, 3, M_CANDLE);

    // big glass ball and plastic ball on the floor in front of the camera, steel torus
    g.sphereMesh(vec3(50, 36.5f, 110), 36, 176, M_GLASS);
    g.sphereMesh(vec3(110, 24.5f, 190), 24, 128, M_PLASTIC);
    g.torus(vec3(10, 10.5f, 200), 30, 10, 160, 48, M_STEEL);

    // floor lamp: wooden pole, fabric shade (open cone frustum), brass finial
    g.cylinder(vec3(300, -0.5f, -60), 16, 5.5f, 48, 1, M_WOODLAMP);
    g.cylinder(vec3(300, 4.5f, -60), 3.2f, 290, 24, 24, M_WOODLAMP);
    g.lathe(vec3(300, 295, -60), [](float v) { return 52 - 22 * v; }, [](float v) { return 70 * v; }, 128, 32, M_SHADE);
    g.sphereMesh(vec3(300, 372, -60), 5, 32, M_BRASS);

    // chandelier ring under the light (does not occlude it: the light is above)
    g.torus(vec3(52.5f, 610, -272.6f), 70, 4, 256, 24, M_GOLD);
    for (int k = 0; k < 12; k++) {
        float a = 2 * kPi * k / 12;
        vec3 base(52.5f + 70 * std::cos(a), 613.5f, -272.6f + 70 * std::sin(a));
        g.cylinder(base, 2.0f, 16, 16, 2, M_CANDLE);
    }
    // a draped cloth (heightfield) on the landing: fine tessellation = most of the triangle budget
    g.grid([&](float u, float v) {
        float x = 20 + 260 * u, z = z0 + 60 + 220 * v;
        float y = landY + 2.5f + 5.5f * (1 + std::sin(0.11f * x) * std::cos(0.13f * z)) + 2.0f * std::sin(0.37f * x + 0.29f * z);
        return vec3(x, y, z);
    }, g.cells(170), g.cells(150), 6, 5, M_SHADE);
}

// ----------------------------------------------------------------- textures --
inline float fract(float x) { return x - std::floor(x); }
inline float hash2(int x, int y, int s) {
    unsigned h = (unsigned)x * 0x9E3779B1u ^ (unsigned)y * 0x85EBCA77u ^ (unsigned)s * 0xC2B2AE3Du;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return (h & 0xFFFFFF) / 16777216.0f;
}
float valueNoise(float x, float y, int seed, int period) {
    int xi = (int)std::floor(x), yi = (int)std::floor(y);
    float fx = x - xi, fy = y - yi;
    fx = fx * fx * (3 - 2 * fx); fy = fy * fy * (3 - 2 * fy);
    auto h = [&](int a, int b) { return hash2(((a % period) + period) % period, ((b % period) + period) % period, seed); };
    float a = h(xi, yi), b = h(xi + 1, yi), c = h(xi, yi + 1), d = h(xi + 1, yi + 1);
    return (a + (b - a) * fx) + ((c + (d - c) * fx) - (a + (b - a) * fx)) * fy;
}
float fbm(float u, float v, int seed, int basePeriod) {
    float s = 0, amp = 0.5f;
    int p = basePeriod;
    for (int o = 0; o < 4; o++) { s += amp * valueNoise(u * p, v * p, seed + o, p); amp *= 0.5f; p *= 2; }
    return s;
}

// texture slots of staircase_scene.h:126-134
void makeTexture(int id, int n, std::vector<float>& out) {
    out.resize((size_t)n * n * 3);
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) {
            float u = (x + 0.5f) / n, v = (y + 0.5f) / n;
            float r = 0.5f, g = 0.5f, b = 0.5f;
            switch (id) {
            case 0: { // WoodFloor: planks
                int plank = (int)(u * 8);
                float grain = fbm(u * 2, v * 24 + plank * 3.7f, 11, 8);
                float gap = (fract(u * 8) < 0.03f || fract(v * 2 + plank * 0.37f) < 0.01f) ? 0.35f : 1.0f;
                float t = 0.55f + 0.45f * grain;
                r = 0.46f * t * gap; g = 0.27f * t * gap; b = 0.12f * t * gap;
            } break;
            case 1: { // Wallpaper: damask-like stripes
                float s = 0.5f + 0.5f * std::sin(2 * kPi * u * 12);
                float m = 0.5f + 0.5f * std::sin(2 * kPi * (u * 6 + 0.25f * std::sin(2 * kPi * v * 6)));
                float t = 0.75f + 0.25f * s * m;
                r = 0.62f * t; g = 0.58f * t; b = 0.47f * t;
            } break;
            case 2: { // Woodpanel
                float grain = fbm(u * 16, v * 2, 23, 8);
                float t = 0.5f + 0.5f * grain;
                r = 0.36f * t; g = 0.20f * t; b = 0.09f * t;
            } break;
            case 3: case 4: case 5: { // paintings: colour fields
                float a = fbm(u, v, 31 + id, 4), c = fbm(u + 0.37f, v + 0.11f, 47 + id, 4), d = fbm(u, v, 59 + id, 8);
                r = 0.15f + 0.8f * a; g = 0.15f + 0.8f * c; b = 0.15f + 0.8f * d;
                if (id == 4) { r *= 0.6f; b = std::min(1.0f, b * 1.2f); }
                if (id == 5) { g *= 0.7f; }
            } break;
            case 6: { // WoodChair
                float grain = fbm(u * 3, v * 20, 71, 8);
                float t = 0.6f + 0.4f * grain;
                r = 0.30f * t; g = 0.16f * t; b = 0.07f * t;
            } break;
            case 7: { // Fabric: weave
                float w = 0.5f + 0.25f * std::sin(2 * kPi * u * 64) + 0.25f * std::sin(2 * kPi * v * 64);
                float t = 0.8f + 0.2f * w;
                r = 0.80f * t; g = 0.74f * t; b = 0.62f * t;
            } break;
            default: { // BrushedAluminium
                float streak = fbm(u * 1, v * 64, 83, 8);
                float t = 0.78f + 0.2f * streak;
                r = 0.91f * t; g = 0.92f * t; b = 0.92f * t;
            } break;
            }
            size_t o = ((size_t)y * n + x) * 3;
            // quantise like an 8-bit PNG decoded by loadTexture (staircase_scene.h:111-113: data[i] / 255.0f)
            out[o + 0] = (float)(unsigned char)std::min(255.0f, std::max(0.0f, r * 255.0f + 0.5f)) / 255.0f;
            out[o + 1] = (float)(unsigned char)std::min(255.0f, std::max(0.0f, g * 255.0f + 0.5f)) / 255.0f;
            out[o + 2] = (float)(unsigned char)std::min(255.0f, std::max(0.0f, b * 255.0f + 0.5f)) / 255.0f;
        }
}

material mk(material_type t, float r, float g, float b, float param, int tex) {
    material m;
    m.type = t; m.color = vec3(r, g, b); m.param = param; m.texId = tex;
    return m;
}

} // namespace

struct crt_scene {
    crt::BuiltMesh built;
    mesh m;
    std::vector<material> materials;
    std::vector<std::vector<float>> texData;
    std::vector<stexture> textures;
    kernel_scene ks;

    void finish(int texSize) {
        // staircase_scene.h:141-160
        materials = {
            mk(DIFFUSE, 0.01f, 0.01f, 0.01f, 0, -1),                 // Black
            mk(METAL, 0.27f, 0.254f, 0.15f, 0.01f, -1),             // Brass
            mk(METAL, 0, 0, 0, 0, 8),                               // BrushedAluminium
            mk(DIFFUSE, 1, 1, 1, 0, -1),                            // Candles
            mk(DIFFUSE, 0.117647f, 0.054902f, 0.0666667f, 0, -1),   // ChairSeat
            mk(GLASS, 1, 1, 1, 1.45f, -1),                          // Glass
            mk(METAL, 1.0f, 0.95f, 0.35f, 0.05f, -1),               // Gold
            mk(DIFFUSE, 0, 0, 0, 0, 7),                             // Lampshade
            mk(DIFFUSE, 0.578596f, 0.578596f, 0.578596f, 0, -1),    // MagnoliaPaint
            mk(DIFFUSE, 0, 0, 0, 0, 3),                             // Painting1
            mk(DIFFUSE, 0, 0, 0, 0, 4),                             // Painting2
            mk(DIFFUSE, 0, 0, 0, 0, 5),                             // Painting3
            mk(METAL, 1, 1, 1, 0.1f, -1),                           // StainlessSteel
            mk(DIFFUSE, 0, 0, 0, 0, 1),                             // wallpaper
            mk(DIFFUSE, 0.578596f, 0.578596f, 0.578596f, 0, -1),    // whitePaint
            mk(DIFFUSE, 1, 1, 1, 0, -1),                            // WhitePlastic
            mk(DIFFUSE, 0, 0, 0, 0, 6),                             // WoodChair
            mk(DIFFUSE, 0, 0, 0, 0, 0),                             // woodFloor
            mk(DIFFUSE, 0, 0, 0, 0, 6),                             // WoodLamp
            mk(DIFFUSE, 0, 0, 0, 0, 2),                             // woodstairs
        };
        texData.resize(9);
        textures.resize(9);
        for (int i = 0; i < 9; i++) {
            makeTexture(i, texSize, texData[i]);
            textures[i].data = texData[i].data();
            textures[i].width = texSize;
            textures[i].height = texSize;
        }
        m.tris = built.tris.data();
        m.numTris = (uint32_t)built.tris.size();
        m.bvh = built.nodes.data();
        m.numBvhNodes = (int)built.nodes.size();
        m.bounds = built.bounds;
        std::memset(&ks, 0, sizeof(ks));
        ks.m = &m;
        ks.floor = plane();
        ks.materials = materials.data();
        ks.numMaterials = (int)materials.size();
        ks.textures = textures.data();
        ks.numTextures = (int)textures.size();
        ks.numPrimitivesPerLeaf = built.primsPerLeaf;
    }
};

extern "C" {

crt_scene* crtSceneCreateStaircaseEx(float detail, int texSize, int primsPerLeaf, int buildMode) {
    if (!(detail > 0.0f) || texSize < 1 || primsPerLeaf < 1 || buildMode < 0 || buildMode > 1) return nullptr;
    MeshGen g;
    g.detail = detail;
    buildStaircase(g);
    crt_scene* s = new crt_scene();
    if (!crt::buildBvh(g.tris, primsPerLeaf, s->built, buildMode ? crt::BUILD_SAH : crt::BUILD_MEDIAN)) { delete s; return nullptr; }
    s->finish(texSize);
    return s;
}

crt_scene* crtSceneCreateStaircase(float detail, int texSize, int primsPerLeaf) {
    return crtSceneCreateStaircaseEx(detail, texSize, primsPerLeaf, 0);
}

crt_scene* crtSceneLoadBVH(const char* path, int texSize) {
    crt_scene* s = new crt_scene();
    if (!crt::loadBvhFile(path, s->built)) { delete s; return nullptr; }
    s->finish(texSize);
    return s;
}

crt_scene* crtSceneFromTriangles(const triangle* tris, int n, int primsPerLeaf, int texSize) {
    if (n < 0 || primsPerLeaf < 1 || texSize < 1) return nullptr;
    std::vector<triangle> v(tris, tris + n);
    crt_scene* s = new crt_scene();
    if (!crt::buildBvh(v, primsPerLeaf, s->built)) { delete s; return nullptr; }
    s->finish(texSize);
    return s;
}

void crtSceneDestroy(crt_scene* s) { delete s; }

// loadTexture, staircase_scene.h:103-118: decode as RGB with the rows flipped (stbi_set_flip_vertically_on_load(true), :120),
// every byte / 255.0f. The scene keeps the floats; the device library uploads them at the next initRenderer.
int crtSceneLoadTexturePNG(crt_scene* s, int index, const char* path) {
    if (!s || index < 0 || index >= (int)s->textures.size()) return -1;
    int w = 0, h = 0;
    std::vector<uint8_t> rgb;
    if (!crt::readPngFile(path, true, w, h, rgb)) return -1;
    std::vector<float>& dst = s->texData[index];
    dst.resize(rgb.size());
    for (size_t i = 0; i < rgb.size(); i++) dst[i] = rgb[i] / 255.0f;
    s->textures[index].data = dst.data();
    s->textures[index].width = w;
    s->textures[index].height = h;
    return 0;
}

// The nine files load_scene reads, in its order (staircase_scene.h:125-133), from `dir`. Returns how many were loaded; the
// others keep their procedural stand-ins.
int crtSceneLoadTextureDir(crt_scene* s, const char* dir) {
    static const char* kNames[9] = {"WoodFloor.png", "Wallpaper.png", "Woodpanel.png", "Painting1.png", "Painting2.png",
                                    "Painting3.png", "WoodChair.png", "Fabric.png", "BrushedAluminium.png"};
    int loaded = 0;
    for (int i = 0; i < 9; i++) {
        const std::string path = std::string(dir) + "/" + kNames[i];
        if (crtSceneLoadTexturePNG(s, i, path.c_str()) == 0) loaded++;
    }
    return loaded;
}

int crtDecodePNG(const unsigned char* bytes, unsigned long long n, int flipVertically, int* width, int* height, unsigned char* rgbOut,
                 unsigned long long capacity) {
    int w = 0, h = 0;
    std::vector<uint8_t> rgb;
    if (!crt::decodePng(bytes, (size_t)n, flipVertically != 0, w, h, rgb)) return -1;
    if (width) *width = w;
    if (height) *height = h;
    if (rgbOut) {
        if (capacity < rgb.size()) return -2;
        std::memcpy(rgbOut, rgb.data(), rgb.size());
    }
    return 0;
}

int crtSceneTextureInfo(const crt_scene* s, int index, int* width, int* height, const float** data) {
    if (!s || index < 0 || index >= (int)s->textures.size()) return -1;
    if (width) *width = s->textures[index].width;
    if (height) *height = s->textures[index].height;
    if (data) *data = s->textures[index].data;
    return 0;
}

int crtSceneSaveBVH(const crt_scene* s, const char* path) { return crt::saveBvhFile(path, s->built) ? 0 : -1; }

const kernel_scene* crtSceneKernelScene(const crt_scene* s) { return &s->ks; }

int crtSceneNumRealTriangles(const crt_scene* s) { return s->built.numRealTris; }

unsigned long long crtSceneHash(const crt_scene* s) {
    unsigned long long h = 1469598103934665603ULL;
    auto mix = [&h](const void* p, size_t n) {
        const unsigned char* b = (const unsigned char*)p;
        for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ULL; }
    };
    // hash field by field: struct padding bytes are not part of the value
    for (const triangle& t : s->built.tris) { mix(&t, 61); }
    mix(s->built.nodes.data(), s->built.nodes.size() * sizeof(bvh_node));
    for (const material& m : s->materials) mix(&m, sizeof(material));
    for (const auto& t : s->texData) mix(t.data(), t.size() * sizeof(float));
    return h;
}

void crtMakeCamera(const float lookfrom[3], const float lookat[3], const float vup[3], float vfov, float aspect,
                   float aperture, float focusDist, camera* out) {
    // helper_structs.h:194-206, same operation order
    vec3 from(lookfrom[0], lookfrom[1], lookfrom[2]), at(lookat[0], lookat[1], lookat[2]), up(vup[0], vup[1], vup[2]);
    out->lens_radius = aperture / 2.0f;
    float theta = vfov * kPi / 180.0f;
    float half_height = std::tan(theta / 2.0f);
    float half_width = aspect * half_height;
    out->origin = from;
    out->w = unit(sub(from, at));
    out->u = unit(cross(up, out->w));
    out->v = cross(out->w, out->u);
    out->lower_left_corner =
        sub(sub(sub(from, mul(out->u, half_width * focusDist)), mul(out->v, half_height * focusDist)), mul(out->w, focusDist));
    out->horizontal = mul(out->u, 2.0f * half_width * focusDist);
    out->vertical = mul(out->v, 2.0f * half_height * focusDist);
}

void crtStaircaseCamera(int nx, int ny, camera* out) {
    // staircase_scene.h:62-73 (double literals narrowed to float by the vec3 ctor)
    const float from[3] = {(float)5.555139, (float)173.679901, (float)494.515045};
    const float at[3] = {(float)5.555139, (float)173.679901, (float)493.515045};
    const float up[3] = {0, 1, 0};
    float dist = length(sub(vec3(from[0], from[1], from[2]), vec3(at[0], at[1], at[2])));
    crtMakeCamera(from, at, up, 42.0f, float(nx) / float(ny), 0.0f, dist, out);
}

static float lcg(unsigned int& state) { // main.cpp:17-20
    state = (214013 * state + 2531011);
    return (float)((state >> 16) & 0x7FFF) / 32767;
}

int crtRtiowScene(unsigned int seed, sphere* outS, material* outM, int cap) {
    // The README-era scene (README.md:3-6): ground + 22x22 small + 3 big spheres, drawn with the host LCG.
    int n = 0;
    auto put = [&](vec3 c, float r, material m) {
        if (n < cap) { outS[n].center = c; outS[n].radius = r; outM[n] = m; }
        n++;
    };
    unsigned int st = seed;
    put(vec3(0, -1000.0f, -1), 1000, mk(DIFFUSE, 0.5f, 0.5f, 0.5f, 0, -1));
    for (int a = -11; a < 11; a++)
        for (int b = -11; b < 11; b++) {
            float choose = lcg(st);
            float cx = a + lcg(st), cz = b + lcg(st);
            vec3 c(cx, 0.2f, cz);
            if (choose < 0.8f) {
                float r = lcg(st) * lcg(st), g = lcg(st) * lcg(st), bl = lcg(st) * lcg(st);
                put(c, 0.2f, mk(DIFFUSE, r, g, bl, 0, -1));
            } else if (choose < 0.95f) {
                float r = 0.5f * (1.0f + lcg(st)), g = 0.5f * (1.0f + lcg(st)), bl = 0.5f * (1.0f + lcg(st));
                float fuzz = 0.5f * lcg(st);
                put(c, 0.2f, mk(METAL, r, g, bl, fuzz, -1));
            } else {
                put(c, 0.2f, mk(GLASS, 1, 1, 1, 1.5f, -1));
            }
        }
    put(vec3(0, 1, 0), 1.0f, mk(GLASS, 1, 1, 1, 1.5f, -1));
    put(vec3(-4, 1, 0), 1.0f, mk(DIFFUSE, 0.4f, 0.2f, 0.1f, 0, -1));
    put(vec3(4, 1, 0), 1.0f, mk(METAL, 0.7f, 0.6f, 0.5f, 0.0f, -1));
    return n;
}

void crtRtiowCamera(int nx, int ny, camera* out) {
    const float from[3] = {13, 2, 3}, at[3] = {0, 0, 0}, up[3] = {0, 1, 0};
    crtMakeCamera(from, at, up, 30.0f, float(nx) / float(ny), 0.1f, 10.0f, out);
}

unsigned int crtLinearToSRGB(float x) { // staircase_scene.h:22-30
    x = std::fmax(x, 0.0f);
    x = std::fmax(1.055f * std::pow(x, 0.416666667f) - 0.055f, 0.0f);
    unsigned int u = (unsigned int)(x * 255.9f);
    return u < 255u ? u : 255u;
}

int crtWritePPM(const char* path, int nx, int ny, const vec3* fb) { // staircase_scene.h:32-43
    FILE* f = std::strcmp(path, "-") == 0 ? stdout : std::fopen(path, "w");
    if (!f) return -1;
    std::fprintf(f, "P3\n%d %d\n255\n", nx, ny);
    for (int j = ny - 1; j >= 0; j--)
        for (int i = 0; i < nx; i++) {
            const vec3& c = fb[(size_t)j * nx + i];
            std::fprintf(f, "%u %u %u\n", crtLinearToSRGB(c.e[0]), crtLinearToSRGB(c.e[1]), crtLinearToSRGB(c.e[2]));
        }
    if (f != stdout) std::fclose(f);
    return 0;
}

static const char kRefMagic[10] = {'R', 'E', 'F', '_', '0', '0', '.', '0', '1', '\0'};

int crtWriteRef(const char* path, int nx, int ny, const vec3* fb) { // main.cpp:25-33
    FILE* f = std::fopen(path, "wb");
    if (!f) return -1;
    bool ok = std::fwrite(kRefMagic, 1, 10, f) == 10 && std::fwrite(&nx, 4, 1, f) == 1 && std::fwrite(&ny, 4, 1, f) == 1 &&
              std::fwrite(fb, sizeof(vec3), (size_t)nx * ny, f) == (size_t)nx * ny;
    std::fclose(f);
    return ok ? 0 : -1;
}

int crtReadRef(const char* path, int nx, int ny, vec3* fb) { // main.cpp:36-60
    FILE* f = std::fopen(path, "rb");
    if (!f) return -1;
    char magic[10];
    int inNx = 0, inNy = 0;
    bool ok = std::fread(magic, 1, 10, f) == 10 && std::memcmp(magic, kRefMagic, 10) == 0 && std::fread(&inNx, 4, 1, f) == 1 &&
              std::fread(&inNy, 4, 1, f) == 1;
    if (ok && (inNx != nx || inNy != ny)) { std::fclose(f); return -2; }
    ok = ok && std::fread(fb, sizeof(vec3), (size_t)nx * ny, f) == (size_t)nx * ny;
    std::fclose(f);
    return ok ? 0 : -1;
}

double crtRmse(const vec3* a, const vec3* b, int nx, int ny) { // main.cpp:116-125
    double error = 0.0;
    for (size_t i = 0; i < (size_t)nx * ny; i++)
        for (int c = 0; c < 3; c++) {
            double d = (double)(a[i].e[c] - b[i].e[c]);
            error += d * d / 3.0;
        }
    return std::sqrt(error / ((double)nx * ny));
}

} // extern "C"
