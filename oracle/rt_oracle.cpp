// rt_oracle.cpp -- CPU restatement of the reference's ray-generation / intersection / shading path.
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity ORACLE of the repo: a from-scratch, scalar
// C++ restatement of windy32/win32-ray-tracing-demo's hot path (src/RayTracingOpt), each routine
// citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it; the product (CUDA) path never does.
//
// PINNING: tests/test_oracle.py runs identical jobs through this restatement and through
// oracle/_ref/libref.so (the unmodified reference sources compiled headless by build_ref.sh) and
// demands bit-equality of triangle streams, accelerator structure hashes, per-primary-ray hit ids,
// distances and traversal sequences, Whitted images and erand48 Monte-Carlo images.  Where the
// reference is absent (the GPU box) the same outputs are pinned by the committed fixtures in
// tests/golden/ (generated from libref.so by tests/golden/make_golden.py).
//
// All arithmetic is IEEE float32 evaluated in the reference's operation order; build with
// -ffp-contract=off and without -ffast-math / -march=native (oracle/Makefile).
#include <algorithm>
#include <chrono>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>
#include <omp.h>

#include "oracle_abi.h"
#ifdef RTB_PRETEST_CHECK
#include "../win32-ray-tracing-demo_b200/csrc/rtb_pretest.h"
#endif

namespace {

// ------------------------------------------------------------------------------------------
// float3 algebra -- reference Vector.cpp:48-86, Point.cpp:20-23
// ------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
const float PI_F = 3.14159265359f; // Vector.h:8

inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline V3 normalize(V3 a) { return a * (1 / sqrtf(a.x * a.x + a.y * a.y + a.z * a.z)); } // Vector.cpp:68-71
inline float length(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
inline float comp(const V3 &a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
inline float &comp(V3 &a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct RayO { V3 o, d; };
inline V3 at(const RayO &r, float t) { return r.o + r.d * t; } // Ray.h:24-27

// ------------------------------------------------------------------------------------------
// RNG -- erand48 (reference erand48.h:53-81: 48-bit LCG a=0x5DEECE66D c=0xB) and the
// counter-based generator the CUDA path uses (Philox4x32-10), so that Monte-Carlo parity can be
// checked path-for-path as well as statistically.
// ------------------------------------------------------------------------------------------
struct Rng
{
    int kind;        // ORACLE_RNG_*
    uint64_t x48;    // erand48 state
    uint32_t key[2]; // philox key
    uint32_t ctr[4]; // philox counter (ctr[0] = block index)
    uint32_t buf[4];
    int used;        // lanes of buf consumed (4 = refill)
};

inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++)
    {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

inline void rngSeedRow(Rng &g, int y)
{ // MainWindow.cpp:273  Xi = {0, 0, (ushort)(y*y*y)}; xseed[2] is the most significant word
    g.kind = ORACLE_RNG_ERAND48;
    const uint32_t y3 = (uint32_t)y * (uint32_t)y * (uint32_t)y;
    g.x48 = (uint64_t)(uint16_t)y3 << 32;
}

inline void rngSeedSample(Rng &g, uint64_t seed, uint32_t pixel, uint32_t sample)
{ // include/rtb.h "counter-based RNG": key = (pixel, sample), counter = (block, seed_lo, seed_hi, 0)
    g.kind = ORACLE_RNG_COUNTER;
    g.key[0] = pixel; g.key[1] = sample;
    g.ctr[0] = 0; g.ctr[1] = (uint32_t)seed; g.ctr[2] = (uint32_t)(seed >> 32); g.ctr[3] = 0;
    g.used = 4;
}

// Counter mode, the stream of the CUDA path (csrc/rtb_kernels.cuh: k_montecarlo): ONE Philox block per traced ray -- block 0 of
// a sample carries the pixel jitter, block k belongs to the k-th ray of the sample in the order the rays are traced -- and the
// four values of a block are addressed by SLOT: 0 Russian roulette, 1 p_type, 2 / 3 the hemisphere pair r1 / r2, 2 the
// reflect-or-transmit test (a vertex draws either the pair or that test).  A fixed block per ray keeps the GPU's draws
// converged (every lane computes its block at the top of the loop) where a sequential stream made each lane refill at its
// own time.  erand48 mode (the reference's stream) ignores blocks and slots: the draws stay sequential in the reference's order.
inline void rngRay(Rng &g)
{
    if (g.kind != ORACLE_RNG_COUNTER) return;
    philox4x32_10(g.ctr, g.key, g.buf);
    g.ctr[0]++;
}
inline double rngNext(Rng &g);
inline double rngDraw(Rng &g, int slot)
{
    if (g.kind != ORACLE_RNG_COUNTER) return rngNext(g);
    return (double)((float)(g.buf[slot] >> 8) * (1.0f / 16777216.0f));
}

inline double rngNext(Rng &g)
{
    if (g.kind == ORACLE_RNG_ERAND48)
    { // erand48.h:53-81
        g.x48 = (g.x48 * 0x5DEECE66Dull + 0xBull) & 0xFFFFFFFFFFFFull;
        return ldexp((double)(g.x48 & 0xFFFF), -48) + ldexp((double)((g.x48 >> 16) & 0xFFFF), -32) +
               ldexp((double)((g.x48 >> 32) & 0xFFFF), -16);
    }
    if (g.used == 4)
    {
        philox4x32_10(g.ctr, g.key, g.buf);
        g.ctr[0]++;
        g.used = 0;
    }
    return (double)((float)(g.buf[g.used++] >> 8) * (1.0f / 16777216.0f));
}

// ------------------------------------------------------------------------------------------
// Materials -- Material.cpp, SolidColorMaterial.cpp:11-19, CheckerMaterial.cpp:11-23,
// RadianceCheckerMaterial.cpp:12-29, PhongMaterial.cpp:13-29, GlassMaterial.cpp:3-7
// ------------------------------------------------------------------------------------------
enum { MAT_SOLID = 0, MAT_CHECKER = 1, MAT_RADIANCE_CHECKER = 2, MAT_PHONG = 3 };
enum { DIR_XOZ = 0, DIR_XOY = 1, DIR_YOZ = 2 };

struct Mat
{
    int kind;
    float diffusiveness, reflectiveness, refractiveness, refractive_index;
    V3 a;        // solid: local colour; phong: diffuse
    V3 b;        // solid: emission;     phong: specular
    float scale; // checker scale
    float p;     // radiance-checker radiance / phong shininess
    int dir;
};

Mat solid(V3 local, V3 emission, float d, float r, float t)
{
    Mat m; memset(&m, 0, sizeof(m));
    m.kind = MAT_SOLID; m.a = local; m.b = emission;
    m.diffusiveness = d; m.reflectiveness = r; m.refractiveness = t;
    return m;
}
Mat glass(float index = 1.46) { Mat m = solid(v3(1, 1, 1), v3(0, 0, 0), 0, 0, 1); m.refractive_index = index; return m; }
Mat checker(float scale, int dir = DIR_XOZ, float refl = 0)
{ // CheckerMaterial.cpp:4-9: Material(1 - reflectiveness, reflectiveness, 0)
    Mat m = solid(v3(0, 0, 0), v3(0, 0, 0), 1 - refl, refl, 0);
    m.kind = MAT_CHECKER; m.scale = scale; m.dir = dir;
    return m;
}
Mat radianceChecker(float radiance, float scale, int dir = DIR_XOZ)
{
    Mat m = solid(v3(0, 0, 0), v3(0, 0, 0), 1, 0, 0);
    m.kind = MAT_RADIANCE_CHECKER; m.scale = scale; m.p = radiance; m.dir = dir;
    return m;
}
Mat phong(V3 diffuse, V3 specular, float shininess, float refl = 0)
{
    Mat m = solid(diffuse, specular, 1 - refl, refl, 0);
    m.kind = MAT_PHONG; m.p = shininess;
    return m;
}

inline float checkerParity(const Mat &m, V3 pos)
{ // CheckerMaterial.cpp:13-21: d = abs(floor(u*s) + floor(v*s)); d = fmod(d, 2)
    float d;
    if (m.dir == DIR_XOZ) d = fabsf(floorf(pos.x * m.scale) + floorf(pos.z * m.scale));
    else if (m.dir == DIR_YOZ) d = fabsf(floorf(pos.y * m.scale) + floorf(pos.z * m.scale));
    else d = fabsf(floorf(pos.x * m.scale) + floorf(pos.y * m.scale));
    return (float)fmod((double)d, 2.0);
}

V3 matLocal(const Mat &m, const RayO &ray, V3 pos, V3 normal)
{
    switch (m.kind)
    {
    case MAT_SOLID: return m.a;
    case MAT_CHECKER: return checkerParity(m, pos) < 1 ? v3(0.15f, 0.15f, 0.15f) : v3(1, 1, 1);
    case MAT_RADIANCE_CHECKER: return v3(0.15f, 0.15f, 0.15f);
    default:
    { // PhongMaterial.cpp:13-29 (uses the stored, unflipped normal; white light from (-1,1,1))
        const V3 lightDir = normalize(v3(-1, 1, 1));
        float NdotL = dot(normal, lightDir);
        NdotL = (NdotL < 0.0f) ? 0.0f : NdotL;
        const V3 H = normalize(lightDir - ray.d);
        float NdotH = dot(normal, H);
        NdotH = (NdotH < 0.0f) ? 0.0f : NdotH;
        const V3 diffuseTerm = m.a * NdotL;
        const V3 specularTerm = m.b * powf(NdotH, m.p);
        return mul(v3(1, 1, 1), diffuseTerm + specularTerm);
    }
    }
}

V3 matEmission(const Mat &m, V3 pos)
{
    if (m.kind == MAT_SOLID) return m.b;
    if (m.kind == MAT_RADIANCE_CHECKER) return checkerParity(m, pos) < 1 ? v3(m.p, m.p, m.p) : v3(0, 0, 0);
    return v3(0, 0, 0); // Material.cpp:19-22
}

// ------------------------------------------------------------------------------------------
// Primitives -- Plane.cpp:3-34, Sphere.cpp:10-37, Triangle.cpp:17-121, Grid.cpp:30-116
// ------------------------------------------------------------------------------------------
struct Hit { bool hit; int id; float t; V3 pos, n; int mat; };

struct Tri { V3 a, b, c, n; int mat; };

struct Counters { long long rays, tris, steps; };
struct Probe { std::vector<int> *seq; Counters *cnt; };

inline float det3(float a11, float a12, float a13, float a21, float a22, float a23, float a31, float a32, float a33)
{ // Triangle.cpp:25-36
    return a11 * a22 * a33 + a12 * a23 * a31 + a13 * a21 * a32 - a13 * a22 * a31 - a11 * a23 * a32 - a12 * a21 * a33;
}

inline bool triIntersect(const Tri &T, const RayO &ray, float &tOut, Probe *pr)
{ // Triangle.cpp:38-121 (Cramer's rule; early-out order det -> t -> beta -> gamma)
    if (pr && pr->cnt) pr->cnt->tris++;
    const float m11 = T.a.x - T.b.x, m21 = T.a.y - T.b.y, m31 = T.a.z - T.b.z;
    const float m12 = T.a.x - T.c.x, m22 = T.a.y - T.c.y, m32 = T.a.z - T.c.z;
    const float m13 = ray.d.x, m23 = ray.d.y, m33 = ray.d.z;
    const float b1 = T.a.x - ray.o.x, b2 = T.a.y - ray.o.y, b3 = T.a.z - ray.o.z;
    const float detM = det3(m11, m12, m13, m21, m22, m23, m31, m32, m33);
    if ((double)fabsf(detM) < 1e-10) return false; // double-typed compare, Triangle.cpp:90
    const float t = det3(m11, m12, b1, m21, m22, b2, m31, m32, b3) / detM;
    if (t < 0.0005f) return false;
    const float beta = det3(b1, m12, m13, b2, m22, m23, b3, m32, m33) / detM;
    if (beta < -0.0001f || beta > 1.0001f) return false;
    const float gamma = det3(m11, b1, m13, m21, b2, m23, m31, b3, m33) / detM;
    if (gamma < -0.0001f || gamma > 1.0001f || 1 - beta - gamma < -0.0001f || 1 - beta - gamma > 1.0001f) return false;
    tOut = t;
    return true;
}

struct PlaneO { V3 normal, position; float dist; };
struct SphereO { V3 center; float radius; };

inline bool planeIntersect(const PlaneO &P, const RayO &ray, float &tOut)
{ // Plane.cpp:9-34
    const V3 op = P.position - ray.o;
    const bool back = dot(P.normal, op) > 0;
    const V3 n = back ? P.normal : P.normal * -1;
    if (dot(n, ray.d) > 0)
    {
        const float distance = dot(op, n) / dot(ray.d, n);
        if (distance >= 0.0005f) { tOut = distance; return true; }
    }
    return false;
}

inline bool sphereIntersect(const SphereO &S, const RayO &ray, float &tOut)
{ // Sphere.cpp:10-37 (assumes a unit direction)
    const V3 co = ray.o - S.center;
    const float b = dot(ray.d, co);
    float delta = b * b - (dot(co, co) - S.radius * S.radius);
    if (delta >= 0)
    {
        delta = sqrtf(delta);
        if (-b + delta >= 0.0005f)
        {
            tOut = (-b - delta >= 0.0005f) ? -b - delta : -b + delta;
            return true;
        }
    }
    return false;
}

inline void updateEntryExit(float &entry, float &exit, float v)
{ // Grid.cpp:30-67
    if (entry == FLT_MAX && exit == FLT_MAX) entry = v;
    else if (exit == FLT_MAX)
    {
        if (v > entry) exit = v;
        else { exit = entry; entry = v; }
    }
    else
    {
        if (v < entry) { exit = entry; entry = v; }
        else if (v < exit) exit = v;
    }
}

inline bool boxIntersect(V3 nearP, V3 size, const RayO &ray, float &entry, float &exit)
{ // Grid.cpp:70-116
    const V3 farP = nearP + size;
    entry = FLT_MAX; exit = FLT_MAX;
    for (int axis = 0; axis < 3; axis++)
    {
        const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
        if ((double)fabsf(comp(ray.d, axis)) > 1e-10)
        {
            float distance = (comp(nearP, axis) - comp(ray.o, axis)) / comp(ray.d, axis);
            V3 p = at(ray, distance);
            if (comp(p, a1) >= comp(nearP, a1) && comp(p, a1) <= comp(farP, a1) &&
                comp(p, a2) >= comp(nearP, a2) && comp(p, a2) <= comp(farP, a2))
                updateEntryExit(entry, exit, distance);
            distance = (comp(farP, axis) - comp(ray.o, axis)) / comp(ray.d, axis);
            p = at(ray, distance);
            if (comp(p, a1) >= comp(nearP, a1) && comp(p, a1) <= comp(farP, a1) &&
                comp(p, a2) >= comp(nearP, a2) && comp(p, a2) <= comp(farP, a2))
                updateEntryExit(entry, exit, distance);
        }
    }
    return entry != FLT_MAX;
}

inline void triBounds(const Tri &t, V3 &mn, V3 &mx)
{ // Triangle.cpp:201-237
    mn = v3(std::min(std::min(t.a.x, t.b.x), t.c.x), std::min(std::min(t.a.y, t.b.y), t.c.y), std::min(std::min(t.a.z, t.b.z), t.c.z));
    mx = v3(std::max(std::max(t.a.x, t.b.x), t.c.x), std::max(std::max(t.a.y, t.b.y), t.c.y), std::max(std::max(t.a.z, t.b.z), t.c.z));
}

// ------------------------------------------------------------------------------------------
// Tunnel: tessellation (TunnelGenerator.cpp:6-366) and accelerators (Tunnel.cpp)
// ------------------------------------------------------------------------------------------
struct KdNodeO { int axis; float split; int left, right; std::vector<int> list; V3 mn, mx; };

struct TunnelO
{
    bool ptBuilders = false; // PerformanceTest's KdTreeAcc builder (event-sweep SAH + automatic termination)
    bool exactGrid = false;  // bin with Triangle::intersectWithGrid (Tunnel.cpp:435-445, the compiled-out branch)
    int algorithm;
    std::vector<Tri> tris; // surface[seg][j] flattened in (seg, j) order
    // grid (Tunnel.h:51-67)
    V3 origin; float csx, csy, csz; int nx, ny, nz;
    std::vector<std::vector<int>> cells;
    // k-d tree (Tunnel.h:75-93); nodes are stored in creation = pre-order
    std::vector<KdNodeO> nodes;
    int leaves; long long leafRefs; int maxDepth;
    // convex accelerator (PerformanceTest/ConvexAcc.h): cross section, path, ring normals, tables
    float width = 0, height = 0;
    int trisPerSegment = 0;
    std::vector<V3> cs, path, nvs;
    std::vector<V3> edge;                 // {A, B, C}: A * x + B * y + C > 0 inside
    std::vector<unsigned char> cellStatus; // [100][100]: 0 Hit, 1 Partial, 2 Miss
    std::vector<short> cellRange;          // [100][100][2]
    std::vector<unsigned short> yAxis;     // [100][360][2 * edges]
    int cxTable = 100;                     // cells per side of the lookup table
    bool cxRoundBins = true;               // direction bin: (int)(deg + 0.5) % 360 (PerformanceTest) or (int)deg (RayTracingOpt)
};
struct RayCtx { bool inTunnel; int segment; }; // RayContext.h:5-13

bool ringConvexity(const std::vector<V3> &front, const std::vector<V3> &rear, std::vector<int> &conn)
{ // TunnelGenerator.cpp:6-176 createPolyhedron; conn: 0 BC, 1 AD, 2 Both, 3 Invalid
    const float TOL = 0.001f;
    const size_t n = front.size();
    int invalid = 0;
    conn.clear();
    // "every other vertex of both rings lies behind the plane (base, normal)" -- lines 46-64 etc.
    auto othersBehind = [&](V3 nrm, V3 base, size_t j, size_t k) {
        for (size_t m = 0; m < n; m++)
            if (m != j && m != k)
            {
                const V3 e = normalize(front[m] - base), f = normalize(rear[m] - base);
                if (dot(nrm, e) > TOL || dot(nrm, f) > TOL) return false;
            }
        return true;
    };
    for (size_t j = 0; j < n; j++)
    {
        const size_t k = (j + 1) % n;
        const V3 A = front[j], B = front[k], C = rear[j], D = rear[k];
        const V3 nCBA = normalize(cross(B - C, A - B)), nCDB = normalize(cross(D - C, B - D));
        bool ok1 = true;
        if (dot(nCBA, normalize(D - A)) > TOL) ok1 = false;
        if (ok1) ok1 = othersBehind(nCBA, A, j, k);
        if (dot(nCDB, normalize(A - C)) > TOL) ok1 = false;
        if (ok1) ok1 = othersBehind(nCDB, C, j, k);
        const V3 nADB = normalize(cross(D - A, B - D)), nACD = normalize(cross(C - A, D - C));
        bool ok2 = true;
        if (dot(nADB, normalize(C - A)) > TOL) ok2 = false;
        if (ok2) ok2 = othersBehind(nADB, A, j, k);
        if (dot(nACD, normalize(B - A)) > TOL) ok2 = false;
        if (ok2) ok2 = othersBehind(nACD, A, j, k);
        if (ok1 && !ok2) conn.push_back(0);
        else if (!ok1 && ok2) conn.push_back(1);
        else if (ok1 && ok2) conn.push_back(2);
        else { conn.push_back(3); invalid++; }
    }
    return invalid == 0;
}

inline Tri makeTri(V3 a, V3 b, V3 c, int mat)
{ // Triangle.cpp:17-23: normal = (b-a) x (c-b), normalised
    Tri t; t.a = a; t.b = b; t.c = c; t.n = normalize(cross(b - a, c - b)); t.mat = mat;
    return t;
}

// `pt`: the generator of the PerformanceTest program (src/PerformanceTest/TunnelGenerator.cpp:176-342).  Same cross
// section and path; the ring at path vertex i is turned by the angle of the AVERAGED directions of the two
// segments meeting there (its "normal vectors", lines 243-279) instead of the segment's own direction + delta.
void generateTunnel(TunnelO &T, float rectWidth, float rectHeight, float archHeight, float pathRadius,
                    float pathAngle, int archSegments, int pathSegments, int groundMat, int wallMat, bool pt = false)
{ // TunnelGenerator.cpp:197-366
    std::vector<V3> cs;
    cs.push_back(v3(rectWidth * 0.5f, 0.0f, 0.0f));
    for (int i = 0; i <= archSegments; i++)
    {
        const float angle = PI_F * i / archSegments;
        cs.push_back(v3(cosf(angle) * rectWidth * 0.5f, sinf(angle) * archHeight + rectHeight, 0.0f));
    }
    cs.push_back(v3(-rectWidth * 0.5f, 0.0f, 0.0f));

    std::vector<std::vector<Tri>> surface(pathSegments);
    std::vector<V3> nvs; // PerformanceTest only
    if (pt)
    {
        const int N = pathSegments;
        std::vector<V3> path(N + 1);
        for (int i = 0; i <= N; i++)
        {
            const float theta = pathAngle * i / pathSegments;
            path[i] = v3(pathRadius * (1.0f - cosf(theta)), 0.0f, -pathRadius * sinf(theta));
        }
        for (int i = 1; i <= N; i++)
        {
            const V3 dir = normalize(path[i] - path[i - 1]);
            V3 prevDir = dir, nextDir = dir;
            if (i > 1) prevDir = normalize(path[i - 1] - path[i - 2]);
            if (i < N) nextDir = normalize(path[i + 1] - path[i]);
            const V3 v1 = normalize((prevDir + dir) * 0.5f), v2 = normalize((dir + nextDir) * 0.5f);
            if (i == 1) nvs.push_back(v1);
            nvs.push_back(v2);
        }
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < pathSegments; i++)
    {
        const float theta1 = pathAngle * i / pathSegments;
        const float theta2 = pathAngle * (i + 1) / pathSegments;
        const V3 p1 = v3(pathRadius * (1.0f - cosf(theta1)), 0.0f, -pathRadius * sinf(theta1));
        const V3 p2 = v3(pathRadius * (1.0f - cosf(theta2)), 0.0f, -pathRadius * sinf(theta2));
        const float delta = (i == pathSegments - 1) ? 0 : pathAngle / pathSegments;
        const V3 fwd = v3(0, 0, -1), seg = p2 - p1;
        float offsetAngle1 = acosf(dot(fwd, seg) * (1.0f / (length(fwd) * length(seg)))); // Vector.cpp:88-91
        float offsetAngle2 = offsetAngle1 + delta;
        if (pt)
        {
            offsetAngle1 = acosf(dot(fwd, nvs[i]) * (1.0f / (length(fwd) * length(nvs[i]))));
            offsetAngle2 = acosf(dot(fwd, nvs[i + 1]) * (1.0f / (length(fwd) * length(nvs[i + 1]))));
        }
        std::vector<V3> front, rear;
        for (size_t j = 0; j < cs.size(); j++)
        {
            const V3 p = cs[j];
            front.push_back(v3(p.x * cosf(offsetAngle1) - p.z * sinf(offsetAngle1), p.y,
                               p.x * sinf(offsetAngle1) + p.z * cosf(offsetAngle1)) + (p1 - v3(0, 0, 0)));
            rear.push_back(v3(p.x * cosf(offsetAngle2) - p.z * sinf(offsetAngle2), p.y,
                              p.x * sinf(offsetAngle2) + p.z * cosf(offsetAngle2)) + (p2 - v3(0, 0, 0)));
        }
        std::vector<int> conn;
        if (!ringConvexity(front, rear, conn)) continue; // non-convex segment contributes nothing (313-317)
        for (size_t j = 0; j < cs.size(); j++)
        {
            const V3 A = front[j], B = front[(j + 1) % front.size()], C = rear[j], D = rear[(j + 1) % rear.size()];
            const int mat = (j == cs.size() - 1) ? groundMat : wallMat;
            if (conn[j] == 1) { surface[i].push_back(makeTri(A, C, D, mat)); surface[i].push_back(makeTri(A, D, B, mat)); }
            else { surface[i].push_back(makeTri(C, D, B, mat)); surface[i].push_back(makeTri(C, B, A, mat)); }
        }
    }
    T.tris.clear();
    for (int i = 0; i < pathSegments; i++) T.tris.insert(T.tris.end(), surface[i].begin(), surface[i].end());
    T.width = rectWidth; T.height = rectHeight + archHeight;
    T.cs = cs; T.nvs = nvs;
    T.trisPerSegment = 2 * (int)cs.size();
    T.path.clear();
    for (int i = 0; i <= pathSegments; i++)
    {
        const float theta = pathAngle * i / pathSegments;
        T.path.push_back(v3(pathRadius * (1.0f - cosf(theta)), 0.0f, -pathRadius * sinf(theta)));
    }
}

void tunnelBounds(const TunnelO &T, V3 &mn, V3 &mx)
{ // Tunnel.cpp:351-370 / 487-506
    mn = v3(FLT_MAX, FLT_MAX, FLT_MAX); mx = v3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (size_t i = 0; i < T.tris.size(); i++)
    {
        V3 a, b; triBounds(T.tris[i], a, b);
        mn = v3(std::min(mn.x, a.x), std::min(mn.y, a.y), std::min(mn.z, a.z));
        mx = v3(std::max(mx.x, b.x), std::max(mx.y, b.y), std::max(mx.z, b.z));
    }
}

// Triangle::intersectWithGrid and its helpers, Triangle.cpp:123-199: separating-axis test of a triangle against one grid
// cell (three box axes, the triangle normal, nine edge cross products).  Compiled out at its call site in the reference
// (Tunnel.cpp:435-445); restated here for the exact-binning OPTION (oracle_job.grid_exact), checked against the reference
// rebuilt with that branch enabled (oracle/_ref/libref_sat.so).
inline float satMin(const V3 *pts, int n, V3 axis)
{ // Triangle.cpp:124-132
    float m = FLT_MAX;
    for (int i = 0; i < n; i++) m = std::min(m, axis.x * pts[i].x + axis.y * pts[i].y + axis.z * pts[i].z);
    return m;
}
inline float satMax(const V3 *pts, int n, V3 axis)
{ // Triangle.cpp:134-142
    float m = -FLT_MAX;
    for (int i = 0; i < n; i++) m = std::max(m, axis.x * pts[i].x + axis.y * pts[i].y + axis.z * pts[i].z);
    return m;
}
inline bool satOnAxis(const V3 *g, const V3 *t, V3 axis)
{ // Triangle.cpp:144-149
    if (satMin(g, 8, axis) > satMax(t, 3, axis)) return false;
    if (satMax(g, 8, axis) < satMin(t, 3, axis)) return false;
    return true;
}
inline V3 satCross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); } // Vector.cpp:79-82
bool triIntersectsCell(const Tri &T, V3 pos, V3 size)
{ // Triangle.cpp:152-199
    const V3 g[8] = {v3(pos.x + 0, pos.y + 0, pos.z + 0), v3(pos.x + 0, pos.y + 0, pos.z + size.z),
                     v3(pos.x + 0, pos.y + size.y, pos.z + 0), v3(pos.x + 0, pos.y + size.y, pos.z + size.z),
                     v3(pos.x + size.x, pos.y + 0, pos.z + 0), v3(pos.x + size.x, pos.y + 0, pos.z + size.z),
                     v3(pos.x + size.x, pos.y + size.y, pos.z + 0), v3(pos.x + size.x, pos.y + size.y, pos.z + size.z)};
    const V3 t[3] = {T.a, T.b, T.c};
    if (!satOnAxis(g, t, v3(1, 0, 0))) return false;
    if (!satOnAxis(g, t, v3(0, 1, 0))) return false;
    if (!satOnAxis(g, t, v3(0, 0, 1))) return false;
    if (!satOnAxis(g, t, T.n)) return false;
    const V3 e1 = v3(T.b.x - T.a.x, T.b.y - T.a.y, T.b.z - T.a.z), e2 = v3(T.c.x - T.b.x, T.c.y - T.b.y, T.c.z - T.b.z),
             e3 = v3(T.a.x - T.c.x, T.a.y - T.c.y, T.a.z - T.c.z);
    const V3 box[3] = {v3(1, 0, 0), v3(0, 1, 0), v3(0, 0, 1)};
    for (int b = 0; b < 3; b++)
    {
        if (!satOnAxis(g, t, satCross(box[b], e1))) return false;
        if (!satOnAxis(g, t, satCross(box[b], e2))) return false;
        if (!satOnAxis(g, t, satCross(box[b], e3))) return false;
    }
    return true;
}

void initGrid(TunnelO &T)
{ // Tunnel.cpp:346-465
    V3 mn, mx; tunnelBounds(T, mn, mx);
    const float width = mx.x - mn.x, height = mx.y - mn.y, depth = mx.z - mn.z;
    if (T.algorithm == 1)
    {
        const float maxLength = std::max(std::max(width, height), depth);
        const float size = maxLength / 399;
        T.origin = v3(mn.x - size / 2, mn.y - size / 2, mn.z - size / 2);
        T.csx = T.csy = T.csz = size;
        T.nx = (int)(width / T.csx + 1.5f); T.ny = (int)(height / T.csy + 1.5f); T.nz = (int)(depth / T.csz + 1.5f);
    }
    else
    {
        T.csx = width / 399; T.csy = height / 399; T.csz = depth / 399;
        T.origin = v3(mn.x - T.csx / 2, mn.y - T.csy / 2, mn.z - T.csz / 2);
        T.nx = T.ny = T.nz = 400;
    }
    T.cells.clear();
    T.cells.resize((size_t)T.nx * T.ny * T.nz);
    for (size_t m = 0; m < T.tris.size(); m++)
    { // AABB-overlap binning (the exact SAT test is disabled in the reference, Tunnel.cpp:435-445)
        V3 a, b; triBounds(T.tris[m], a, b);
        const int xb = (int)((a.x - T.origin.x) / T.csx), yb = (int)((a.y - T.origin.y) / T.csy), zb = (int)((a.z - T.origin.z) / T.csz);
        const int xe = (int)((b.x - T.origin.x) / T.csx), ye = (int)((b.y - T.origin.y) / T.csy), ze = (int)((b.z - T.origin.z) / T.csz);
        for (int i = xb; i <= xe; i++)
            for (int j = yb; j <= ye; j++)
                for (int k = zb; k <= ze; k++)
                { // Tunnel.cpp:435-445: the exact branch (option) or the simple one (what the reference ships)
                    if (T.exactGrid && !triIntersectsCell(T.tris[m], v3(T.origin.x + i * T.csx, T.origin.y + j * T.csy, T.origin.z + k * T.csz),
                                                          v3(T.csx, T.csy, T.csz))) continue;
                    T.cells[((size_t)i * T.ny + j) * T.nz + k].push_back((int)m);
                }
    }
}

struct CentroidLess
{ // Tunnel.cpp:519-544
    const std::vector<Tri> *tris; int axis;
    bool operator()(int p, int q) const
    {
        const Tri &s = (*tris)[p], &t = (*tris)[q];
        const float c1 = (comp(s.a, axis) + comp(s.b, axis) + comp(s.c, axis)) / 3;
        const float c2 = (comp(t.a, axis) + comp(t.b, axis) + comp(t.c, axis)) / 3;
        return c1 < c2;
    }
};

float splitMedian(TunnelO &T, int axis, std::vector<int> &list)
{ // Tunnel.cpp:649-669 (std::sort reorders the caller's list; children inherit that order)
    CentroidLess less = {&T.tris, axis};
    std::sort(list.begin(), list.end(), less);
    const Tri &t = T.tris[list[list.size() / 2]];
    return (comp(t.a, axis) + comp(t.b, axis) + comp(t.c, axis)) / 3;
}

float splitSAH(TunnelO &T, int node, const std::vector<int> &list, int &bestAxis)
{ // Tunnel.cpp:671-784: 99 uniform candidates per axis, first strict minimum wins
    float minSAH = FLT_MAX, minSplit = 0;
    const V3 mn = T.nodes[node].mn, mx = T.nodes[node].mx;
    for (int axis = 0; axis < 3; axis++)
    {
        const int N = 100;
        float cand[N];
        int lc[N], rc[N];
        for (int i = 1; i < N; i++) cand[i] = comp(mn, axis) + (comp(mx, axis) - comp(mn, axis)) * i / N;
        const int nextAxis = (axis + 1) % 3, prevAxis = (axis + 2) % 3;
#pragma omp parallel for schedule(static) if (list.size() > 4096)
        for (int i = 1; i < N; i++)
        {
            const float s = cand[i];
            int l = 0, r = 0;
            for (size_t j = 0; j < list.size(); j++)
            {
                const Tri &t = T.tris[list[j]];
                if (comp(t.a, axis) < s || comp(t.b, axis) < s || comp(t.c, axis) < s) l++;
                if (comp(t.a, axis) >= s || comp(t.b, axis) >= s || comp(t.c, axis) >= s) r++;
            }
            lc[i] = l; rc[i] = r;
        }
        for (int i = 1; i < N; i++)
        {
            const V3 boxSize = mx - mn;
            const float leftWidth = cand[i] - comp(mn, axis), rightWidth = comp(mx, axis) - cand[i];
            const float height = comp(boxSize, nextAxis), depth = comp(boxSize, prevAxis);
            const float SAH = (leftWidth * height + leftWidth * depth + height * depth) * lc[i] +
                              (rightWidth * height + rightWidth * depth + height * depth) * rc[i];
            if (SAH < minSAH) { minSAH = SAH; minSplit = cand[i]; bestAxis = axis; }
        }
    }
    return minSplit;
}

// PerformanceTest/KdTreeAcc.cpp:177-274: exact SAH by a sweep over the sorted bounding-box events of every
// axis.  Events sort by (position, type) with End < Planar < Start (KdTreeAcc.cpp:32-36); equal keys are
// interchangeable, so std::sort's instability cannot change a count.  Cost = 1 + 1.5 * ((SAL/SA) * NL +
// (SAR/SA) * (NR + NP)), first strict minimum wins (axis 0 -> 2, positions ascending).
struct SweepEvent { float position; int type; }; // type: 0 End, 1 Planar, 2 Start
inline bool sweepLess(const SweepEvent &a, const SweepEvent &b)
{
    return (a.position < b.position) || ((a.position == b.position) && a.type < b.type);
}

float splitSAHSweep(TunnelO &T, int node, const std::vector<int> &list, int &bestAxis, float &minSAH)
{
    minSAH = FLT_MAX;
    float minPosition = 0;
    const V3 mn = T.nodes[node].mn, mx = T.nodes[node].mx;
    for (int axis = 0; axis < 3; axis++)
    {
        std::vector<SweepEvent> events;
        events.reserve(2 * list.size());
        for (size_t i = 0; i < list.size(); i++)
        {
            V3 a, b; triBounds(T.tris[list[i]], a, b);
            if (comp(a, axis) == comp(b, axis)) events.push_back(SweepEvent{comp(a, axis), 1});
            else { events.push_back(SweepEvent{comp(a, axis), 2}); events.push_back(SweepEvent{comp(b, axis), 0}); }
        }
        std::sort(events.begin(), events.end(), sweepLess);
        int NL = 0, NP = 0, NR = (int)list.size();
        for (size_t i = 0; i < events.size();)
        {
            const float position = events[i].position;
            int PS = 0, PE = 0, PP = 0;
            while (i < events.size() && events[i].position == position && events[i].type == 0) { PE += 1; i += 1; }
            while (i < events.size() && events[i].position == position && events[i].type == 1) { PP += 1; i += 1; }
            while (i < events.size() && events[i].position == position && events[i].type == 2) { PS += 1; i += 1; }
            NP = PP; NR -= PP; NR -= PE;
            const int nextAxis = (axis + 1) % 3, prevAxis = (axis + 2) % 3;
            const V3 boxSize = mx - mn;
            const float width = comp(mx, axis) - comp(mn, axis);
            const float leftWidth = position - comp(mn, axis);
            const float rightWidth = comp(mx, axis) - position;
            const float height = comp(boxSize, nextAxis);
            const float depth = comp(boxSize, prevAxis);
            const float SAL = leftWidth * height + leftWidth * depth + height * depth;
            const float SAR = rightWidth * height + rightWidth * depth + height * depth;
            const float SA = width * height + width * depth + height * depth;
            const float SAH = 1 + 1.5f * ((SAL / SA) * NL + SAR / SA * (NR + NP));
            if (SAH < minSAH) { minSAH = SAH; minPosition = position; bestAxis = axis; }
            NL += PS; NL += PP; NP = 0;
        }
    }
    return minPosition;
}

void buildKd(TunnelO &T, int node, std::vector<int> &list, int depth)
{ // Tunnel.cpp:546-638; PerformanceTest/KdTreeAcc.cpp:38-144 when T.ptBuilders
    if (depth > T.maxDepth) T.maxDepth = depth;
    if (list.size() <= 8 || depth > 18)
    {
        T.nodes[node].axis = 3;
        T.nodes[node].split = 0.0f;
        T.nodes[node].left = T.nodes[node].right = -1;
        T.nodes[node].list = list;
        T.leaves++;
        T.leafRefs += (long long)list.size();
        return;
    }
    int axis = 0;
    float median;
    if (T.algorithm == 3) { axis = depth % 3; median = splitMedian(T, axis, list); }
    else if (T.ptBuilders)
    {
        float sah;
        median = splitSAHSweep(T, node, list, axis, sah);
        if (sah > 1.5f * list.size())
        { // automatic termination, KdTreeAcc.cpp:77-88
            T.nodes[node].axis = 3;
            T.nodes[node].split = 0.0f;
            T.nodes[node].left = T.nodes[node].right = -1;
            T.nodes[node].list = list;
            T.leaves++;
            T.leafRefs += (long long)list.size();
            return;
        }
    }
    else median = splitSAH(T, node, list, axis);
    std::vector<int> leftPart, rightPart;
    for (size_t i = 0; i < list.size(); i++)
    { // straddlers go to both sides (619-634)
        const Tri &t = T.tris[list[i]];
        if (comp(t.a, axis) < median || comp(t.b, axis) < median || comp(t.c, axis) < median) leftPart.push_back(list[i]);
        if (comp(t.a, axis) >= median || comp(t.b, axis) >= median || comp(t.c, axis) >= median) rightPart.push_back(list[i]);
    }
    T.nodes[node].axis = axis;
    T.nodes[node].split = median;
    const V3 mn = T.nodes[node].mn, mx = T.nodes[node].mx;
    const int L = (int)T.nodes.size();
    T.nodes.push_back(KdNodeO());
    T.nodes[node].left = L;
    T.nodes[L].mn = mn; T.nodes[L].mx = mx; comp(T.nodes[L].mx, axis) = median;
    buildKd(T, L, leftPart, depth + 1);
    const int R = (int)T.nodes.size();
    T.nodes.push_back(KdNodeO());
    T.nodes[node].right = R;
    T.nodes[R].mn = mn; T.nodes[R].mx = mx; comp(T.nodes[R].mn, axis) = median;
    buildKd(T, R, rightPart, depth + 1);
}

void initKd(TunnelO &T)
{ // Tunnel.cpp:467-517
    std::vector<int> list(T.tris.size());
    for (size_t i = 0; i < list.size(); i++) list[i] = (int)i;
    T.nodes.clear();
    T.nodes.push_back(KdNodeO());
    tunnelBounds(T, T.nodes[0].mn, T.nodes[0].mx);
    T.leaves = 0; T.leafRefs = 0; T.maxDepth = 0;
    buildKd(T, 0, list, 0);
}

#ifdef RTB_PRETEST_CHECK
// Checker of the product's conservative rejection test (win32-ray-tracing-demo_b200/csrc/rtb_pretest.h), built only into
// librt_oracle_pretest.so (make pretest; tests/test_pretest.py): the same host+device function the kernels run is
// evaluated next to the exact test on every (ray, triangle) pair of a frame.  A "violation" is a pair the pre-test
// rejects although the exact test would have made it the nearest hit of its list.
struct PretestStats { long long tests, candidates, updates, violations, lists, lists1, lists2, unsure; };
PretestStats g_pre = {0, 0, 0, 0, 0, 0, 0, 0};
int g_pre_use_nearest = 0;
#endif

// nearest hit of a triangle list, first-in-list wins ties (strict <)
inline bool nearestInList(const TunnelO &T, const std::vector<int> &list, const RayO &ray, float lo, float hi,
                          bool window, int &triOut, float &tOut, Probe *pr)
{
    float minDistance = FLT_MAX;
    bool found = false;
#ifdef RTB_PRETEST_CHECK
    PretestStats st = {0, 0, 0, 0, 1, 0, 0, 0};
    const float dmx = rtb_pre::dirMax(ray.d.x, ray.d.y, ray.d.z);
    const float Lp = rtb_pre::lowBound(window ? lo : -FLT_MAX);
    int cands = 0;
#endif
    for (size_t i = 0; i < list.size(); i++)
    {
        float t;
#ifdef RTB_PRETEST_CHECK
        const Tri &K = T.tris[list[i]];
        const float rec[9] = {K.a.x, K.a.y, K.a.z, K.b.x, K.b.y, K.b.z, K.c.x, K.c.y, K.c.z};
        const rtb_pre::PreTri P = rtb_pre::makePreTri(rec);
        // the kernels defer the exact test of a candidate to the end of its list, so their H is the leaf window
        // alone (g_pre_use_nearest = 0); = 1 also feeds the nearest hit so far, the tightest H the header allows
        const float Hp = rtb_pre::highBound(window ? hi : FLT_MAX, g_pre_use_nearest ? minDistance : FLT_MAX);
        const bool rejected = window ? rtb_pre::sureReject<true>(P, ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z, dmx, Lp, Hp)
                                     : rtb_pre::sureReject<false>(P, ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z, dmx, Lp, Hp);
        st.tests++;
        if (!rejected) { st.candidates++; cands++; }
        // with L' = +inf the t test fires whenever it is allowed to: the call returns "the sign of det_M is certain"
        if (!rejected && !rtb_pre::sureReject<false>(P, ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z, dmx, INFINITY, Hp)) st.unsure++;
        const bool wouldUpdate = triIntersect(K, ray, t, nullptr) && !(window && !(t >= lo && t <= hi)) && t < minDistance;
        if (wouldUpdate) st.updates++;
        if (wouldUpdate && rejected) st.violations++;
#endif
        if (!triIntersect(T.tris[list[i]], ray, t, pr)) continue;
        if (window && !(t >= lo && t <= hi)) continue;
        if (t < minDistance) { minDistance = t; triOut = list[i]; found = true; }
    }
#ifdef RTB_PRETEST_CHECK
    st.lists1 = cands >= 1; st.lists2 = cands >= 2;
#pragma omp critical(pretest_stats)
    {
        g_pre.tests += st.tests; g_pre.candidates += st.candidates; g_pre.updates += st.updates;
        g_pre.violations += st.violations; g_pre.lists += st.lists; g_pre.lists1 += st.lists1; g_pre.lists2 += st.lists2; g_pre.unsure += st.unsure;
    }
#endif
    tOut = minDistance;
    return found;
}

inline void indexInGrid(const TunnelO &T, V3 p, int &i, int &j, int &k)
{ // Tunnel.cpp:806-817
    i = (int)((p.x - T.origin.x) / T.csx); j = (int)((p.y - T.origin.y) / T.csy); k = (int)((p.z - T.origin.z) / T.csz);
    if (i < 0) i = 0;
    if (j < 0) j = 0;
    if (k < 0) k = 0;
    if (i > T.nx - 1) i = T.nx - 1;
    if (j > T.ny - 1) j = T.ny - 1;
    if (k > T.nz - 1) k = T.nz - 1;
}

bool gridIntersect(const TunnelO &T, const RayO &ray, int &triOut, float &tOut, Probe *pr)
{ // Tunnel.cpp:819-970 (3D-DDA; accepts the nearest hit of the current cell's list even outside the cell)
    const V3 nearP = T.origin;
    const V3 extent = v3(T.csx * T.nx, T.csy * T.ny, T.csz * T.nz);
    const V3 farP = T.origin + extent;
    int ci, cj, ck; float cd; V3 cp;
    if (ray.o.x < nearP.x || ray.o.x > farP.x || ray.o.y < nearP.y || ray.o.y > farP.y || ray.o.z < nearP.z || ray.o.z > farP.z)
    {
        float entry, exit;
        if (!boxIntersect(T.origin, extent, ray, entry, exit)) return false;
        cd = entry; cp = at(ray, entry);
        indexInGrid(T, cp, ci, cj, ck);
    }
    else { cp = ray.o; cd = 0; indexInGrid(T, ray.o, ci, cj, ck); }
    while (true)
    {
        const int cell = (ci * T.ny + cj) * T.nz + ck;
        if (pr) { if (pr->seq) pr->seq->push_back(cell); if (pr->cnt) pr->cnt->steps++; }
        if (nearestInList(T, T.cells[cell], ray, 0, 0, false, triOut, tOut, pr)) return true;
        const V3 p1 = T.origin + v3(ci * T.csx, cj * T.csy, ck * T.csz);
        const V3 p2 = T.origin + v3((ci + 1) * T.csx, (cj + 1) * T.csy, (ck + 1) * T.csz);
        float dx, dy, dz; int di, dj, dk;
        if (ray.d.x > 0) { dx = (p2.x - cp.x) / ray.d.x; di = 1; } else { dx = (cp.x - p1.x) / -ray.d.x; di = -1; }
        if (ray.d.y > 0) { dy = (p2.y - cp.y) / ray.d.y; dj = 1; } else { dy = (cp.y - p1.y) / -ray.d.y; dj = -1; }
        if (ray.d.z > 0) { dz = (p2.z - cp.z) / ray.d.z; dk = 1; } else { dz = (cp.z - p1.z) / -ray.d.z; dk = -1; }
        if (dx < dy && dx < dz) { ci += di; cd += dx; }
        else if (dy < dz) { cj += dj; cd += dy; }
        else { ck += dk; cd += dz; }
        cp = at(ray, cd);
        if (ci < 0 || ci > T.nx - 1 || cj < 0 || cj > T.ny - 1 || ck < 0 || ck > T.nz - 1) break;
    }
    return false;
}

bool kdIntersect(const TunnelO &T, const RayO &ray, int &triOut, float &tOut, Probe *pr)
{ // Tunnel.cpp:1163-1297 (Havran TA_rec_B, 50-entry stack)
    struct Elem { int node; float t; V3 pb; int prev; };
    float a, b, t;
    if (!boxIntersect(T.nodes[0].mn, T.nodes[0].mx - T.nodes[0].mn, ray, a, b)) return false; // Grid(near, far): size = far - near
    Elem stack[50];
    int cur = 0, farChild = -1;
    int enPt = 0;
    stack[enPt].t = a;
    if (a >= 0) stack[enPt].pb = ray.o + ray.d * a; else stack[enPt].pb = ray.o;
    int exPt = 1;
    stack[exPt].t = b;
    stack[exPt].pb = ray.o + ray.d * b;
    stack[exPt].node = -1;
    while (cur != -1)
    {
        while (T.nodes[cur].axis != 3)
        {
            if (pr) { if (pr->seq) pr->seq->push_back(cur); if (pr->cnt) pr->cnt->steps++; }
            const float splitVal = T.nodes[cur].split;
            const int axis = T.nodes[cur].axis, nextAxis = (axis + 1) % 3, prevAxis = (axis + 2) % 3;
            if (comp(stack[enPt].pb, axis) <= splitVal)
            {
                if (comp(stack[exPt].pb, axis) <= splitVal) { cur = T.nodes[cur].left; continue; }
                if (comp(stack[exPt].pb, axis) == splitVal) { cur = T.nodes[cur].right; continue; }
                farChild = T.nodes[cur].right; cur = T.nodes[cur].left;
            }
            else
            {
                if (splitVal < comp(stack[exPt].pb, axis)) { cur = T.nodes[cur].right; continue; }
                farChild = T.nodes[cur].left; cur = T.nodes[cur].right;
            }
            t = (splitVal - comp(ray.o, axis)) / comp(ray.d, axis);
            const int tmp = exPt++;
            if (exPt == enPt) exPt += 1;
            stack[exPt].prev = tmp;
            stack[exPt].t = t;
            stack[exPt].node = farChild;
            comp(stack[exPt].pb, axis) = splitVal;
            comp(stack[exPt].pb, nextAxis) = comp(ray.o, nextAxis) + t * comp(ray.d, nextAxis);
            comp(stack[exPt].pb, prevAxis) = comp(ray.o, prevAxis) + t * comp(ray.d, prevAxis);
        }
        if (pr) { if (pr->seq) pr->seq->push_back(cur); if (pr->cnt) pr->cnt->steps++; }
        if (nearestInList(T, T.nodes[cur].list, ray, stack[enPt].t - 0.001f, stack[exPt].t + 0.001f, true, triOut, tOut, pr)) return true;
        enPt = exPt;
        cur = stack[exPt].node;
        exPt = stack[enPt].prev;
    }
    return false;
}

bool linearIntersect(const TunnelO &T, const RayO &ray, int &triOut, float &tOut, Probe *pr)
{ // Tunnel.cpp:786-804
    float minDistance = FLT_MAX;
    bool found = false;
    for (size_t i = 0; i < T.tris.size(); i++)
    {
        float t;
        if (triIntersect(T.tris[i], ray, t, pr) && t < minDistance) { minDistance = t; triOut = (int)i; found = true; }
    }
    tOut = minDistance;
    return found;
}

// ------------------------------------------------------------------------------------------
// Scene, camera, settings -- GeometrySet.cpp, Camera.cpp:4-26, RenderSetting.h:40-78
// ------------------------------------------------------------------------------------------
enum { PRIM_PLANE = 0, PRIM_SPHERE = 1, PRIM_TRIANGLE = 2, PRIM_TUNNEL = 3 };
// ------------------------------------------------------------------------------------------
// Convex accelerator -- PerformanceTest/ConvexAcc.cpp (the author's own method): the tunnel is a chain
// of convex polyhedra between consecutive cross-section polygons; a ray inside walks from polygon to
// polygon (a 2D point-in-convex-polygon test through a 100x100 lookup table) until it leaves through a
// wall, where it tests that segment's triangles -- all of them in list order (ConvexSimple) or in the
// order of a (height, angle) table (Convex) -- and returns the FIRST accepted one, not the nearest.
// ------------------------------------------------------------------------------------------
// table resolution: 100 in PerformanceTest/ConvexAcc.h:12, 400 in RayTracingOpt (Tunnel.cpp:255-284)
inline int f2i(float f) { return (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : (int)0x80000000; } // cvttss2si

inline bool convexCorner(const TunnelO &T, float x, float y)
{ // ConvexAcc.cpp:103-118: inside all edges (with the 0.0001 leak margin)
    for (size_t j = 0; j < T.edge.size(); j++)
        if (T.edge[j].x * x + T.edge[j].y * y + T.edge[j].z < 0.0001f) return false;
    return true;
}

void initConvex(TunnelO &T)
{ // PerformanceTest/ConvexAcc.cpp:181-271; RayTracingOpt/Tunnel.cpp:135-344 when !T.ptBuilders
    const size_t n = T.cs.size();
    T.cxTable = T.ptBuilders ? 100 : 400;
    T.cxRoundBins = T.ptBuilders;
    if (!T.ptBuilders)
    { // RayTracingOpt derives the ring normals here: the direction of the segment that starts at the vertex (Tunnel.cpp:139-150)
        T.nvs.clear();
        for (size_t i = 0; i < T.path.size(); i++)
        {
            if (i + 1 < T.path.size()) T.nvs.push_back(normalize(T.path[i + 1] - T.path[i]));
            else T.nvs.push_back(normalize((T.path[i] + (T.path[i] - T.path[i - 1])) - T.path[i]));
        }
    }
    const int CONVEX_TABLE = T.cxTable;
    T.edge.clear();
    for (size_t i = 0; i < n; i++)
    {
        const V3 p1 = T.cs[i], p2 = T.cs[(i + 1) % n];
        T.edge.push_back(v3(p1.y - p2.y, p2.x - p1.x, p1.x * p2.y - p2.x * p1.y));
    }
    T.cellStatus.assign(CONVEX_TABLE * CONVEX_TABLE, 2);
    T.cellRange.assign(CONVEX_TABLE * CONVEX_TABLE * 2, -1);
    for (int i = 0; i < CONVEX_TABLE; i++)
        for (int j = 0; j < CONVEX_TABLE; j++)
        {
            const float cellWidth = T.width / (CONVEX_TABLE - 1.0f), cellHeight = T.height / (CONVEX_TABLE - 1.0f);
            const float cx = i * cellWidth - T.width / 2, cy = j * cellHeight;
            const float px[4] = {cx + -cellWidth / 2, cx + cellWidth / 2, cx + -cellWidth / 2, cx + cellWidth / 2};
            const float py[4] = {cy + -cellHeight / 2, cy + -cellHeight / 2, cy + cellHeight / 2, cy + cellHeight / 2};
            int hit = 0;
            for (int c = 0; c < 4; c++) hit += convexCorner(T, px[c], py[c]) ? 1 : 0;
            short mn = -1, mx = -1;
            unsigned char status;
            if (hit == 4) status = 0;
            else if (hit == 0) status = 2;
            else
            { // Partial: the range of edges that may cross the cell (ConvexAcc.cpp:136-176)
                status = 1;
                mn = SHRT_MAX; mx = SHRT_MIN;
                for (size_t e = 0; e < n; e++)
                {
                    const V3 start = T.cs[e], end = T.cs[(e + 1) % n];
                    float sd[4];
                    for (int c = 0; c < 4; c++) sd[c] = T.edge[e].x * px[c] + T.edge[e].y * py[c] + T.edge[e].z;
                    if ((sd[0] > 0.0001 && sd[1] > 0.0001 && sd[2] > 0.0001 && sd[3] > 0.0001) ||
                        (sd[0] < -0.0001 && sd[1] < -0.0001 && sd[2] < -0.0001 && sd[3] < -0.0001)) continue; // double compares
                    if (start.x > px[3] && end.x > px[3]) continue;
                    if (start.x < px[0] && end.x < px[0]) continue;
                    if (start.y > py[3] && end.y > py[3]) continue;
                    if (start.y < py[0] && end.y < py[0]) continue;
                    if ((short)e > mx) mx = (short)e;
                    if ((short)e < mn) mn = (short)e;
                }
            }
            if (status == 1 && !T.ptBuilders) { mn = 0; mx = (short)(n - 1); } // RayTracingOpt tests every edge (inPolygon, Tunnel.cpp:102-114)
            T.cellStatus[i * CONVEX_TABLE + j] = status;
            T.cellRange[(i * CONVEX_TABLE + j) * 2] = mn;
            T.cellRange[(i * CONVEX_TABLE + j) * 2 + 1] = mx;
        }
    T.yAxis.clear();
    if (T.algorithm != 5) return;
    T.yAxis.resize((size_t)100 * 360 * 2 * n);
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = 0; y < 100; y++)
        for (int iAngle = 0; iAngle < 360; iAngle++)
        { // ConvexAcc.cpp:236-270: edges ordered by how far (in angle) they are from the ray (0, height*(y+.5)/100) + angle
            const float ty = T.height * (y + 0.5f) / 100.0f;
            const float targetAngle = iAngle / 360.0f * PI_F * 2;
            std::multimap<float, int> mapping;
            for (size_t e = 0; e < n; e++)
            {
                const V3 p1 = T.cs[e], p2 = T.cs[(e + 1) % n];
                const V3 v1 = v3(p1.x - 0, p1.y - ty, p1.z - 0), v2 = v3(p2.x - 0, p2.y - ty, p2.z - 0);
                const V3 v = v3(cosf(targetAngle), sinf(targetAngle), 0);
                float delta;
                if (v1.x * v.y - v1.y * v.x > 0 && v.x * v2.y - v.y * v2.x > 0) delta = 0;
                else
                {
                    const float mxp = (p1.x + p2.x) / 2, myp = (p1.y + p2.y) / 2;
                    const float angle = atan2f(myp - ty, mxp - 0.0f);
                    delta = targetAngle - angle;
                    delta = (delta > PI_F) ? delta - 2 * PI_F : delta;
                    delta = fabsf(delta);
                }
                mapping.insert(std::multimap<float, int>::value_type(delta, (int)e));
            }
            unsigned short *row = &T.yAxis[((size_t)y * 360 + iAngle) * 2 * n];
            size_t k = 0;
            for (std::multimap<float, int>::iterator it = mapping.begin(); it != mapping.end(); ++it)
            {
                row[k++] = (unsigned short)(it->second * 2);
                row[k++] = (unsigned short)(it->second * 2 + 1);
            }
        }
}

// ConvexAcc.cpp:37-84; `distance` is only written when the ray crosses the polygon's plane
inline bool convexAtOrigin(const TunnelO &T, V3 o, V3 d, float &distance)
{
    if (o.z * d.z >= 0) return false;
    distance = (0 - o.z) / d.z;
    const V3 p = o + d * distance;
    const int CONVEX_TABLE = T.cxTable;
    const float cellWidth = T.width / (CONVEX_TABLE - 1.0f), cellHeight = T.height / (CONVEX_TABLE - 1.0f);
    int i = f2i((p.x + T.width / 2) / cellWidth + 0.5f), j = f2i(p.y / cellHeight + 0.5f);
    i = std::max(i, 0); j = std::max(j, 0);
    i = std::min(i, CONVEX_TABLE - 1); j = std::min(j, CONVEX_TABLE - 1);
    const unsigned char st = T.cellStatus[i * CONVEX_TABLE + j];
    if (st == 0) return true;
    if (st == 2) return false;
    const int begin = T.cellRange[(i * CONVEX_TABLE + j) * 2], end = T.cellRange[(i * CONVEX_TABLE + j) * 2 + 1];
    for (int e = begin; e <= end; e++)
        if (T.edge[e].x * p.x + T.edge[e].y * p.y + T.edge[e].z < 0.0001f) return false;
    return true;
}

// ConvexAcc.cpp:7-35: bring the ray into the frame of polygon `index` (rotation about y + translation)
inline bool convexPolygon(const TunnelO &T, const RayO &ray, int index, V3 &origin, V3 &dir, float &distance)
{
    const V3 p = T.path[index], n = T.nvs[index];
    const float theta = PI_F - atan2f(n.x, n.z);
    const float c = cosf(theta), s = sinf(theta);
    const V3 q = v3(ray.o.x + (0 - p.x), ray.o.y + (0 - p.y), ray.o.z + (0 - p.z));
    V3 newOrigin = v3(c * q.x + 0 * q.y + s * q.z, 0 * q.x + 1 * q.y + 0 * q.z, -s * q.x + 0 * q.y + c * q.z);
    V3 newDir = v3(c * ray.d.x + 0 * ray.d.y + s * ray.d.z, 0 * ray.d.x + 1 * ray.d.y + 0 * ray.d.z, -s * ray.d.x + 0 * ray.d.y + c * ray.d.z);
    const bool hit = convexAtOrigin(T, newOrigin, newDir, distance);
    if (!hit)
    {
        newOrigin.z = 0;
        newDir.z = 0;
        newDir = normalize(newDir);
        origin = newOrigin;
        dir = newDir;
    }
    return hit;
}

bool linearIntersect(const TunnelO &T, const RayO &ray, int &triOut, float &tOut, Probe *pr);

// ConvexAcc.cpp:273-415.  Returns the first accepted triangle of the wall segment the ray leaves through.
bool convexIntersect(const TunnelO &T, const RayO &ray, RayCtx &ctx, int &triOut, float &tOut, Probe *pr)
{
    RayO adv = ray;
    float distance = 0; // uninitialised in the reference; only read after a call that wrote it, except through
                        // the stale-value path described at convexAtOrigin
    const int N = (int)T.path.size() - 1;
    if (!ctx.inTunnel)
    {
        V3 no, nd;
        if (convexPolygon(T, ray, 0, no, nd, distance) && dot(ray.d, T.nvs[0]) > 0)
        {
            ctx.inTunnel = true; ctx.segment = 0;
            adv.o = at(ray, distance); adv.d = ray.d;
        }
        else if (convexPolygon(T, ray, N, no, nd, distance) && dot(ray.d, T.nvs[N]) < 0)
        {
            ctx.inTunnel = true; ctx.segment = N - 1;
            adv.o = at(ray, distance); adv.d = ray.d;
        }
        else return linearIntersect(T, ray, triOut, tOut, pr);
    }
    const int begin = ctx.segment;
    if (!(dot(ray.d, T.nvs[begin]) > 0)) return false; // "Backward": unimplemented in the reference (lines 405-410)
    for (int i = begin + 1; i <= N; i++)
    {
        V3 no, nd;
        if (!convexPolygon(T, adv, i, no, nd, distance))
        { // leaves through the wall of segment i - 1
            const int base = (i - 1) * T.trisPerSegment;
            if (T.algorithm == 6)
            {
                for (int j = 0; j < T.trisPerSegment; j++)
                {
                    float t;
                    if (triIntersect(T.tris[base + j], ray, t, pr)) { ctx.segment = i - 1; triOut = base + j; tOut = t; return true; }
                }
            }
            else
            {
                const float y = no.y - no.x * nd.y / nd.x;
                int index = f2i(99.0f * y / T.height + 0.5f);
                index = std::max(0, index); index = std::min(99, index);
                float fAngle = atan2f(nd.y, nd.x);
                fAngle = (fAngle < 0) ? fAngle + PI_F * 2 : fAngle;
                int iAngle = T.cxRoundBins ? f2i(fAngle / PI_F * 180.0f + 0.5f) % 360 : f2i(fAngle / PI_F * 180.0f);
                iAngle = std::max(0, iAngle); iAngle = std::min(359, iAngle);
                const unsigned short *row = &T.yAxis[((size_t)index * 360 + iAngle) * T.trisPerSegment];
                for (int j = 0; j < T.trisPerSegment; j++)
                {
                    float t;
                    if (triIntersect(T.tris[base + row[j]], ray, t, pr)) { ctx.segment = i - 1; triOut = base + row[j]; tOut = t; return true; }
                }
                adv.o = at(adv, distance);
            }
        }
        else adv.o = at(adv, distance);
    }
    ctx.inTunnel = false;
    return false;
}

struct Prim { int type; PlaneO plane; SphereO sphere; Tri tri; int mat; };

struct Camera { V3 eye, front, up, right; float ratio, xcenter, fov, fovScale, forward; };

Camera makeCamera(V3 eye, V3 front, V3 up, float ratio, float fov, float forward)
{ // Camera.cpp:4-18 (front is normalised in place before right/up are derived)
    Camera c;
    c.eye = eye;
    front = normalize(front);
    c.front = front;
    c.ratio = ratio; c.fov = fov; c.forward = forward;
    c.right = normalize(cross(front, up));
    c.up = normalize(cross(c.right, front));
    c.xcenter = ratio * 0.5f;
    c.fovScale = tanf(fov * (PI_F * 0.5f / 180)) * 2;
    return c;
}

inline RayO generateRay(const Camera &c, float x, float y)
{ // Camera.cpp:20-26
    const V3 r = c.right * ((x - c.xcenter) * c.fovScale);
    const V3 u = c.up * ((y - 0.5f) * c.fovScale);
    const V3 dir = normalize(c.front + r + u);
    RayO ray = {c.eye + dir * c.forward, dir};
    return ray;
}

struct Setting { bool mc; int maxDepth, terminationDepth, singleTracingDepth; };
Setting settingOf(int which)
{
    switch (which)
    {
    case ORACLE_SETTING_HIGHSPEED: { Setting s = {true, 6, 2, 0}; return s; }
    case ORACLE_SETTING_HIGHQUALITY: { Setting s = {true, 8, INT_MAX, INT_MAX}; return s; }
    case ORACLE_SETTING_DEFAULT: { Setting s = {true, INT_MAX, 5, 2}; return s; }
    default: { Setting s = {false, 20, INT_MAX, 0}; return s; }
    }
}

struct Scene
{
    std::vector<Prim> prims;
    std::vector<Mat> mats;
    TunnelO tunnel;
    bool hasTunnel;
    Camera cam;
    Setting setting;
};

int addMat(Scene &s, const Mat &m) { s.mats.push_back(m); return (int)s.mats.size() - 1; }
void addPlane(Scene &s, V3 normal, float dist, int mat)
{ // Plane.cpp:3-7
    Prim p; memset(&p, 0, sizeof(p));
    p.type = PRIM_PLANE; p.plane.normal = normal; p.plane.dist = dist;
    p.plane.position = v3(0, 0, 0) + normal * dist; p.mat = mat;
    s.prims.push_back(p);
}
void addSphere(Scene &s, V3 c, float r, int mat)
{
    Prim p; memset(&p, 0, sizeof(p));
    p.type = PRIM_SPHERE; p.sphere.center = c; p.sphere.radius = r; p.mat = mat;
    s.prims.push_back(p);
}

bool addStl(Scene &s, const char *path, int mat, V3 offset)
{ // GeometrySet.cpp:33-86 with the identity matrix Script3 passes (Scripts.cpp:146-151)
    FILE *fp = fopen(path, "rb");
    if (!fp) return false;
    char header[80]; int32_t count = 0;
    if (fread(header, 80, 1, fp) != 1 || fread(&count, 4, 1, fp) != 1) { fclose(fp); return false; }
    for (int i = 0; i < count; i++)
    {
        float f[12]; int16_t attr;
        if (fread(f, 4, 12, fp) != 12 || fread(&attr, 2, 1, fp) != 1) { fclose(fp); return false; }
        auto xform = [&](V3 p) { // Matrix.cpp:21-28 with m = identity, then + offset
            return v3(1 * p.x + 0 * p.y + 0 * p.z, 0 * p.x + 1 * p.y + 0 * p.z, 0 * p.x + 0 * p.y + 1 * p.z) + offset;
        };
        Prim p; memset(&p, 0, sizeof(p));
        p.type = PRIM_TRIANGLE; p.mat = mat;
        p.tri.a = xform(v3(f[3], f[4], f[5])); p.tri.b = xform(v3(f[6], f[7], f[8])); p.tri.c = xform(v3(f[9], f[10], f[11]));
        const V3 n = normalize(v3(f[0], f[1], f[2]));
        p.tri.n = v3(1 * n.x + 0 * n.y + 0 * n.z, 0 * n.x + 1 * n.y + 0 * n.z, 0 * n.x + 0 * n.y + 1 * n.z);
        p.tri.mat = mat;
        s.prims.push_back(p);
    }
    fclose(fp);
    return true;
}

bool buildPreset(Scene &s, const oracle_job *job)
{ // Scripts.cpp:22-278
    s.hasTunnel = false;
    const V3 black = v3(0, 0, 0), white = v3(1, 1, 1);
    if (job->preset == 1)
    {
        addPlane(s, v3(0, 1, 0), 0, addMat(s, radianceChecker(1.2f, 0.025f)));
        addSphere(s, v3(-10, 15, -30), 15, addMat(s, solid(white, black, 1, 0, 0)));
        addSphere(s, v3(20, 10, -20), 10, addMat(s, solid(white, black, 1, 0, 0)));
        s.cam = makeCamera(v3(0, 15, 30), v3(0, 0, -1), v3(0, 1, 0), 1.3333f, 65, 0);
        s.setting = settingOf(ORACLE_SETTING_DEFAULT);
    }
    else if (job->preset == 2 || job->preset == 3)
    {
        addPlane(s, v3(1, 0, 0), 1, addMat(s, solid(v3(0.75f, 0.25f, 0.25f), black, 1, 0, 0)));
        addPlane(s, v3(1, 0, 0), 99, addMat(s, solid(v3(0.25f, 0.25f, 0.75f), black, 1, 0, 0)));
        addPlane(s, v3(0, 1, 0), 0, addMat(s, solid(v3(0.75f, 0.75f, 0.75f), black, 1, 0, 0)));
        addPlane(s, v3(0, 1, 0), 81.6f, addMat(s, solid(v3(0.75f, 0.75f, 0.75f), black, 1, 0, 0)));
        addPlane(s, v3(0, 0, 1), 0, addMat(s, solid(v3(0.75f, 0.75f, 0.75f), black, 1, 0, 0)));
        addPlane(s, v3(0, 0, 1), 170, addMat(s, solid(v3(0, 0, 0), black, 1, 0, 0)));
        addSphere(s, v3(50, 681.6f - 0.27f, 81.6f), 600, addMat(s, solid(black, v3(24, 24, 24), 1, 0, 0)));
        if (job->preset == 2)
        {
            addSphere(s, v3(27, 16.5f, 47), 16.5f, addMat(s, solid(white, black, 0, 1, 0)));
            addSphere(s, v3(73, 16.5f, 78), 16.5f, addMat(s, glass()));
        }
        else if (!addStl(s, job->stl_path ? job->stl_path : "ball.stl", addMat(s, glass()), v3(50, 0, 40)))
            return false;
        s.cam = makeCamera(v3(50, 52, 295), v3(0, -0.045f, -1), v3(0, 1, -0.045f), 1.3333f, 28, 140.0f);
        s.setting = settingOf(ORACLE_SETTING_DEFAULT);
    }
    else
    {
        const bool wide = job->preset == 4;
        addSphere(s, wide ? v3(74.12f, 15, -96.59f) : v3(5000, 15, -5000), 15, addMat(s, phong(v3(1, 0, 0), white, 16)));
        addPlane(s, v3(0, 1, 0), -0.01f, addMat(s, solid(v3(0.25, 0.25, 0.25), black, 1, 0, 0)));
        addPlane(s, v3(0, 1, 0), 1000, addMat(s, solid(white, black, 1, 0, 0)));
        const int ground = addMat(s, checker(0.05f));
        const int wall = addMat(s, solid(black, black, 0.333f, 0.667f, 0));
        s.tunnel.algorithm = job->algorithm;
        generateTunnel(s.tunnel, 50, 25, 25, wide ? 100.0f : 5000.0f, wide ? PI_F * 0.416667f : PI_F * 0.5f,
                       job->segments, job->segments, ground, wall);
        Prim p; memset(&p, 0, sizeof(p));
        p.type = PRIM_TUNNEL;
        s.prims.push_back(p);
        s.hasTunnel = true;
        s.cam = makeCamera(v3(0, 25, 20), v3(0, 0, -1), v3(0, 1, 0), 1.3333f, 65, 0.0f);
        s.setting = settingOf(ORACLE_SETTING_SIMPLE);
    }
    return true;
}

Hit sceneIntersect(const Scene &s, const RayO &ray, Probe *pr, RayCtx *ctx = nullptr)
{ // GeometrySet.cpp:95-110 (strict <, insertion order) + Tunnel.cpp:1299-1309 dispatch
    if (pr && pr->cnt) pr->cnt->rays++;
    Hit best; memset(&best, 0, sizeof(best));
    best.hit = false; best.id = -1;
    float minDistance = FLT_MAX;
    const int nTop = (int)s.prims.size();
    for (int i = 0; i < nTop; i++)
    {
        const Prim &p = s.prims[i];
        float t; Hit h; h.hit = false;
        if (p.type == PRIM_PLANE)
        {
            if (planeIntersect(p.plane, ray, t)) { h.hit = true; h.id = i; h.t = t; h.pos = at(ray, t); h.n = p.plane.normal; h.mat = p.mat; }
        }
        else if (p.type == PRIM_SPHERE)
        {
            if (sphereIntersect(p.sphere, ray, t)) { h.hit = true; h.id = i; h.t = t; h.pos = at(ray, t); h.n = normalize(h.pos - p.sphere.center); h.mat = p.mat; }
        }
        else if (p.type == PRIM_TRIANGLE)
        {
            if (triIntersect(p.tri, ray, t, pr)) { h.hit = true; h.id = i; h.t = t; h.pos = at(ray, t); h.n = p.tri.n; h.mat = p.mat; }
        }
        else
        {
            int tri = -1; bool ok;
            const TunnelO &T = s.tunnel;
            if (T.algorithm == 1 || T.algorithm == 2) ok = gridIntersect(T, ray, tri, t, pr);
            else if (T.algorithm == 3 || T.algorithm == 4) ok = kdIntersect(T, ray, tri, t, pr);
            else if (T.algorithm == 5 || T.algorithm == 6)
            { // the ray's context is mutated by the tunnel whatever the other geometries return (Ray.h:11-16)
                RayCtx fresh = {false, -1};
                ok = convexIntersect(T, ray, ctx ? *ctx : fresh, tri, t, pr);
            }
            else ok = linearIntersect(T, ray, tri, t, pr);
            if (ok) { h.hit = true; h.id = nTop + tri; h.t = t; h.pos = at(ray, t); h.n = T.tris[tri].n; h.mat = T.tris[tri].mat; }
        }
        if (h.hit && h.t < minDistance) { minDistance = h.t; best = h; }
    }
    return best;
}

// ------------------------------------------------------------------------------------------
// Shading -- MainWindow.cpp:69-143 (trace) and 145-249 (radiance)
// ------------------------------------------------------------------------------------------
struct Fresnel { RayO refl; V3 tdir; bool tir; float Re, Tr, P, RP, TP; };

inline Fresnel refraction(const RayO &r, V3 p, V3 n, V3 nl, float nt)
{ // MainWindow.cpp:111-133 == 212-231
    Fresnel f; memset(&f, 0, sizeof(f));
    f.refl.o = p; f.refl.d = r.d - n * 2 * dot(n, r.d);
    const bool into = dot(n, nl) > 0;
    const float nc = 1;
    const float nnt = into ? nc / nt : nt / nc;
    const float ddn = dot(r.d, nl);
    const float cos2t = 1 - nnt * nnt * (1 - ddn * ddn);
    f.tir = cos2t < 0;
    if (f.tir) return f;
    f.tdir = normalize(r.d * nnt - n * ((into ? 1 : -1) * (ddn * nnt + sqrtf(cos2t))));
    const float a = nt - nc, b = nt + nc;
    const float R0 = a * a / (b * b);
    const float c = 1 - (into ? -ddn : dot(f.tdir, n));
    f.Re = R0 + (1 - R0) * c * c * c * c * c;
    f.Tr = 1 - f.Re;
    f.P = 0.25f + 0.5f * f.Re;
    f.RP = f.Re / f.P;
    f.TP = f.Tr / (1 - f.P);
    return f;
}

V3 trace(const Scene &s, RayO r, int depth, Probe *pr, RayCtx ctx = RayCtx{false, -1})
{
    const Hit res = sceneIntersect(s, r, pr, &ctx);
    if (!res.hit) return v3(0, 0, 0);
    const Mat &m = s.mats[res.mat];
    const V3 p = res.pos, n = res.n;
    const V3 nl = (dot(n, r.d) < 0) ? n : n * -1;
    const V3 local = matLocal(m, r, p, res.n);
    if (++depth > s.setting.maxDepth) return v3(0, 0, 0);
    if (depth > 100) return v3(0, 0, 0);
    V3 diffusive = v3(0, 0, 0), reflective = v3(0, 0, 0), refractive = v3(0, 0, 0); // zero-initialised
    if (m.diffusiveness > 0) diffusive = local;
    if (m.reflectiveness > 0)
    {
        RayO nr = {p, r.d - nl * 2 * dot(nl, r.d)};
        reflective = trace(s, nr, depth, pr, ctx); // newRay.context = r.context, MainWindow.cpp:105
    }
    if (m.refractiveness > 0)
    {
        const Fresnel f = refraction(r, p, n, nl, m.refractive_index);
        if (f.tir) refractive = trace(s, f.refl, depth, pr);
        else
        {
            const V3 a = trace(s, f.refl, depth, pr) * f.Re;
            RayO tr = {p, f.tdir};
            refractive = a + trace(s, tr, depth, pr) * f.Tr;
        }
    }
    return diffusive * m.diffusiveness + reflective * m.reflectiveness + refractive * m.refractiveness;
}

V3 radiance(const Scene &s, RayO r, int depth, Rng &rng, Probe *pr)
{
    rngRay(rng); // counter mode: this ray's block
    const Hit res = sceneIntersect(s, r, pr);
    if (!res.hit) return v3(0, 0, 0);
    const Mat &m = s.mats[res.mat];
    const V3 p = res.pos, n = res.n;
    const V3 nl = (dot(n, r.d) < 0) ? n : n * -1;
    V3 local = matLocal(m, r, p, res.n);
    const V3 emission = matEmission(m, p);
    const float maxColor = (local.x + local.y + local.z) * 0.333333f;
    if (++depth > s.setting.maxDepth) return emission;
    if (depth > s.setting.terminationDepth)
    {
        if (rngDraw(rng, 0) < maxColor) local = local * (1 / maxColor);
        else return emission;
    }
    if (depth > 100) return emission;
    const float p_type = (float)rngDraw(rng, 1);
    if (m.diffusiveness > 0 && p_type < m.diffusiveness)
    { // uniform hemisphere, unit weight (MainWindow.cpp:185-200)
        const float r1 = (float)rngDraw(rng, 2), r2 = (float)rngDraw(rng, 3);
        const float theta = 2 * PI_F * r1;
        const float phi = acosf(r2);
        const V3 w = nl;
        const V3 u = ((double)fabsf(w.x) > 0.1) ? normalize(cross(v3(0, 1, 0), w)) : normalize(cross(v3(1, 0, 0), w));
        const V3 v = cross(w, u);
        const V3 dir = u * (cosf(theta) * sinf(phi)) + v * (sinf(theta) * sinf(phi)) + w * cosf(phi);
        RayO nr = {p, dir};
        return emission + mul(local, radiance(s, nr, depth, rng, pr));
    }
    if (m.reflectiveness > 0 && p_type >= m.diffusiveness && p_type <= m.diffusiveness + m.reflectiveness)
    {
        RayO nr = {p, r.d - nl * 2 * dot(nl, r.d)};
        return emission + mul(local, radiance(s, nr, depth, rng, pr));
    }
    if (m.refractiveness > 0 && p_type > m.diffusiveness + m.reflectiveness)
    {
        const Fresnel f = refraction(r, p, n, nl, m.refractive_index);
        if (f.tir) return emission + mul(local, radiance(s, f.refl, depth, rng, pr));
        RayO tr = {p, f.tdir};
        if (depth > s.setting.singleTracingDepth)
        {
            if ((float)rngDraw(rng, 2) < f.P) return radiance(s, f.refl, depth, rng, pr) * f.RP;
            return radiance(s, tr, depth, rng, pr) * f.TP;
        }
        // C++ leaves the evaluation order of the two operands of `+` unspecified; the g++ 13 build of
        // the reference (oracle/_ref) evaluates the TRANSMITTED subtree first, which fixes the order
        // in which the two subtrees consume the per-row erand48 stream (MainWindow.cpp:242-243).
        const V3 b = radiance(s, tr, depth, rng, pr) * f.Tr;
        const V3 a = radiance(s, f.refl, depth, rng, pr) * f.Re;
        return a + b;
    }
    return emission;
}

void render(const Scene &s, const oracle_job *job, float *rgb, Counters *total)
{ // MainWindow.cpp:251-303
    const int width = job->width, height = job->height, samples = job->samples;
    const float dx = 1.0f / height, dy = 1.0f / height;
    std::vector<Counters> cnt(omp_get_max_threads() + 1);
    memset(cnt.data(), 0, cnt.size() * sizeof(Counters));
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = 0; y < height; y++)
    {
        Probe pr = {nullptr, &cnt[omp_get_thread_num()]};
        Rng rng;
        rngSeedRow(rng, y);
        for (int x = 0; x < width; x++)
        {
            const size_t index = (size_t)x * height + y;
            V3 c;
            if (s.setting.mc)
            {
                c = v3(0, 0, 0);
                double m[6] = {0, 0, 0, 0, 0, 0};
                for (int i = 0; i < samples; i++)
                {
                    if (job->rng == ORACLE_RNG_COUNTER) rngSeedSample(rng, job->seed, (uint32_t)(y * width + x), (uint32_t)i);
                    rngRay(rng); // counter mode: block 0 of the sample
                    const float r1 = (float)rngDraw(rng, 0), r2 = (float)rngDraw(rng, 1);
                    const float sx = (x + r1) * dx, sy = 1 - (y + r2) * dy;
                    const V3 L = radiance(s, generateRay(s.cam, sx, sy), 0, rng, &pr);
                    c = c + L * (1.0f / samples);
                    m[0] += L.x; m[1] += L.y; m[2] += L.z;
                    m[3] += (double)L.x * L.x; m[4] += (double)L.y * L.y; m[5] += (double)L.z * L.z;
                }
                if (job->moments && rgb)
                    for (int k = 0; k < 6; k++) job->moments[6 * index + k] = m[k];
            }
            else
            {
                const float sx = (x + 0.5f) * dx, sy = 1 - (y + 0.5f) * dy;
                c = trace(s, generateRay(s.cam, sx, sy), 0, &pr);
            }
            if (rgb) { rgb[3 * index] = c.x; rgb[3 * index + 1] = c.y; rgb[3 * index + 2] = c.z; }
        }
    }
    total->rays = total->tris = total->steps = 0;
    for (size_t i = 0; i < cnt.size(); i++) { total->rays += cnt[i].rays; total->tris += cnt[i].tris; total->steps += cnt[i].steps; }
}

// ------------------------------------------------------------------------------------------
// canonical structure hashes (same definition in ref_driver.cpp and the product binding)
// ------------------------------------------------------------------------------------------
inline void hmix(uint64_t &h, uint32_t v) { h = (h ^ v) * 0x100000001b3ull; }
inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
const uint64_t H0 = 0xcbf29ce484222325ull;

void hashKd(const TunnelO &T, int node, uint64_t &h)
{
    const KdNodeO &n = T.nodes[node];
    if (n.axis == 3)
    {
        hmix(h, 3); hmix(h, (uint32_t)n.list.size());
        for (size_t i = 0; i < n.list.size(); i++) hmix(h, (uint32_t)n.list[i]);
        return;
    }
    hmix(h, (uint32_t)n.axis); hmix(h, fbits(n.split));
    hashKd(T, n.left, h);
    hashKd(T, n.right, h);
}

// convex accelerator tables: frames (path vertex, ring normal, cos / sin of the ring rotation), edges, cell table, order table
uint64_t hashConvex(const TunnelO &T)
{
    uint64_t h = H0;
    hmix(h, 0x43565800u);
    hmix(h, (uint32_t)T.path.size()); hmix(h, (uint32_t)T.cs.size());
    hmix(h, fbits(T.width)); hmix(h, fbits(T.height));
    for (size_t i = 0; i < T.path.size(); i++)
    {
        const float theta = PI_F - atan2f(T.nvs[i].x, T.nvs[i].z);
        const float v[8] = {T.path[i].x, T.path[i].y, T.path[i].z, T.nvs[i].x, T.nvs[i].y, T.nvs[i].z, cosf(theta), sinf(theta)};
        for (int q = 0; q < 8; q++) hmix(h, fbits(v[q]));
    }
    for (size_t e = 0; e < T.edge.size(); e++) { hmix(h, fbits(T.edge[e].x)); hmix(h, fbits(T.edge[e].y)); hmix(h, fbits(T.edge[e].z)); }
    for (size_t i = 0; i < T.cellStatus.size(); i++) hmix(h, T.cellStatus[i]);
    for (size_t i = 0; i < T.cellRange.size(); i++) hmix(h, (uint32_t)(uint16_t)T.cellRange[i]);
    for (size_t i = 0; i < T.yAxis.size(); i++) hmix(h, T.yAxis[i]);
    return h;
}

} // namespace

// ------------------------------------------------------------------------------------------
// PerformanceTest console benchmark -- reference src/PerformanceTest/main.cpp:29-59 (trace),
// 61-81 (scene), 143-162 (ray loop), Camera.cpp:15-21
// ------------------------------------------------------------------------------------------
extern "C" int rt_oracle_bounce(oracle_bounce_job *job)
{
    if (!job || job->n < 0 || job->algorithm < 0 || job->algorithm > 6 || !job->xy) return -1;
    if (job->algorithm > 4 && !job->pt_builders) return -1; // the convex accelerator needs that program's ring normals
    Scene s;
    s.hasTunnel = true;
    const int dummy = addMat(s, solid(v3(0, 0, 0), v3(0, 0, 0), 1, 0, 0));
    s.tunnel.algorithm = job->algorithm;
    s.tunnel.ptBuilders = job->pt_builders != 0;
    generateTunnel(s.tunnel, 50, 25, 25, job->radius, job->angle, job->arch_seg, job->path_seg, dummy, dummy, job->pt_builders != 0);
    Prim p; memset(&p, 0, sizeof(p));
    p.type = PRIM_TUNNEL;
    s.prims.push_back(p);
    // exit plane, main.cpp:71-78
    const V3 normal = v3(sinf(job->angle), 0, -cosf(job->angle));
    addPlane(s, normal, job->radius * sinf(job->angle), dummy);
    const auto t0 = std::chrono::steady_clock::now();
    if (job->algorithm == 1 || job->algorithm == 2) initGrid(s.tunnel);
    else if (job->algorithm == 3 || job->algorithm == 4) initKd(s.tunnel);
    else if (job->algorithm == 5 || job->algorithm == 6) initConvex(s.tunnel);
    job->prepare_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    memset(job->stats, 0, sizeof(job->stats));
    job->struct_hash = 0;
    job->stats[ORACLE_STAT_N_TRIS] = (int64_t)s.tunnel.tris.size();
    if (job->algorithm == 1 || job->algorithm == 2)
    {
        const TunnelO &T = s.tunnel;
        uint64_t h = H0;
        hmix(h, 0x47524944u);
        hmix(h, (uint32_t)T.nx); hmix(h, (uint32_t)T.ny); hmix(h, (uint32_t)T.nz);
        hmix(h, fbits(T.origin.x)); hmix(h, fbits(T.origin.y)); hmix(h, fbits(T.origin.z));
        hmix(h, fbits(T.csx)); hmix(h, fbits(T.csy)); hmix(h, fbits(T.csz));
        job->stats[ORACLE_STAT_GRID_X] = T.nx; job->stats[ORACLE_STAT_GRID_Y] = T.ny; job->stats[ORACLE_STAT_GRID_Z] = T.nz;
        for (size_t c = 0; c < T.cells.size(); c++)
        {
            const std::vector<int> &l = T.cells[c];
            if (l.empty()) continue;
            job->stats[ORACLE_STAT_CELLS_NONEMPTY]++;
            job->stats[ORACLE_STAT_CELL_ENTRIES] += (int64_t)l.size();
            if ((int64_t)l.size() > job->stats[ORACLE_STAT_CELL_MAX]) job->stats[ORACLE_STAT_CELL_MAX] = (int64_t)l.size();
            hmix(h, (uint32_t)c); hmix(h, (uint32_t)l.size());
            for (size_t i = 0; i < l.size(); i++) hmix(h, (uint32_t)l[i]);
        }
        job->struct_hash = h;
    }
    if (job->algorithm == 5 || job->algorithm == 6) job->struct_hash = hashConvex(s.tunnel);
    if (job->algorithm == 3 || job->algorithm == 4)
    {
        const TunnelO &T = s.tunnel;
        uint64_t h = H0;
        hmix(h, 0x4b445452u);
        hmix(h, fbits(T.nodes[0].mn.x)); hmix(h, fbits(T.nodes[0].mn.y)); hmix(h, fbits(T.nodes[0].mn.z));
        hmix(h, fbits(T.nodes[0].mx.x)); hmix(h, fbits(T.nodes[0].mx.y)); hmix(h, fbits(T.nodes[0].mx.z));
        hashKd(T, 0, h);
        job->struct_hash = h;
        job->stats[ORACLE_STAT_N_TRIS] = (int64_t)T.tris.size();
        job->stats[ORACLE_STAT_KD_NODES] = (int64_t)T.nodes.size();
        job->stats[ORACLE_STAT_KD_LEAVES] = T.leaves;
        job->stats[ORACLE_STAT_KD_LEAF_REFS] = T.leafRefs;
        job->stats[ORACLE_STAT_KD_MAX_DEPTH] = T.maxDepth;
    }
    // PT camera: eye (0,25,5), front (0,0,-1), up (0,1,0); Camera.cpp:5-13
    V3 front = normalize(v3(0, 0, -1));
    const V3 right = normalize(cross(front, v3(0, 1, 0)));
    const V3 up = normalize(cross(right, front));
    const V3 eye = v3(0, 25, 5);
    long long total = 0;
    omp_set_num_threads(job->threads > 0 ? job->threads : omp_get_num_procs());
    const auto t1 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
    for (int i = 0; i < job->n; i++)
    {
        const float x = job->xy[2 * i], y = job->xy[2 * i + 1];
        const V3 r = right * ((x - 0.5f) * 1.274f), u = up * ((y - 0.5f) * 1.0f); // Camera.cpp:17-18
        RayO ray = {eye, normalize(front + r + u)};
        int depth = 0, reached = 0, lastId = -1;
        V3 lastPos = v3(0, 0, 0);
        RayCtx ctx = {false, -1}; // newRay.context = r.context, main.cpp:57
        while (true)
        { // main.cpp:29-59
            const Hit h = sceneIntersect(s, ray, nullptr, &ctx);
            total++;
            if (!h.hit) { lastId = -1; break; }
            lastId = h.id; lastPos = h.pos;
            const V3 nl = (dot(h.n, ray.d) < 0) ? h.n : h.n * -1;
            if (++depth > job->max_depth) break;
            if (s.prims[h.id < (int)s.prims.size() ? h.id : 0].type == PRIM_PLANE && h.id < (int)s.prims.size()) { reached = 1; break; }
            const V3 v = ray.d - nl * 2 * dot(nl, ray.d);
            ray.o = h.pos; ray.d = v;
        }
        if (job->reached) job->reached[i] = reached;
        if (job->depth) job->depth[i] = depth;
        if (job->last_id) job->last_id[i] = lastId;
        if (job->last_pos) { job->last_pos[3 * i] = lastPos.x; job->last_pos[3 * i + 1] = lastPos.y; job->last_pos[3 * i + 2] = lastPos.z; }
    }
    job->trace_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    job->total_rays = total;
    return 0;
}

#ifdef RTB_PRETEST_CHECK
// out[8] = tests, candidates, updates, violations, lists, lists with >= 1 candidate, lists with >= 2, candidates whose
// det_M sign was uncertain (grazing rays); reset on read
// Random stress of the rejection test away from the preset scenes: triangles of widely varying size and aspect,
// rays aimed at a point of the triangle's plane close to (inside or just outside) its boundary, from near and far,
// head-on to grazing, unit and non-unit directions.  out[4] = pairs, exact accepts, candidates, violations.
extern "C" void rt_oracle_pretest_fuzz(uint64_t seed, long long n, long long *out)
{
    long long acc = 0, cand = 0, viol = 0;
#pragma omp parallel for reduction(+ : acc, cand, viol) schedule(static)
    for (long long i = 0; i < n; i++)
    {
        uint32_t ctr[4] = {(uint32_t)i, (uint32_t)(i >> 32), 0, 0}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, r[4];
        float u[20];
        for (int b = 0; b < 5; b++)
        {
            ctr[2] = (uint32_t)b;
            philox4x32_10(ctr, key, r);
            for (int q = 0; q < 4; q++) u[4 * b + q] = (float)((r[q] >> 8) * (1.0 / 16777216.0));
        }
        const float scale = powf(10.0f, -3.0f + 7.0f * u[0]);            // triangle size 1e-3 .. 1e4
        const float aspect = powf(10.0f, -3.0f * u[1]);                   // down to 1000 : 1 slivers
        const float offset = (u[2] < 0.5f) ? 0.0f : powf(10.0f, 5.0f * u[3]) - 1.0f; // scene offset up to 1e5
        Tri K;
        K.a = v3(offset + scale * (u[4] - 0.5f), offset * 0.5f + scale * (u[5] - 0.5f), scale * (u[6] - 0.5f) - offset);
        K.b = K.a + v3(scale * (u[7] - 0.5f), scale * (u[8] - 0.5f), scale * (u[9] - 0.5f));
        K.c = K.a + v3(scale * aspect * (u[10] - 0.5f), scale * aspect * (u[11] - 0.5f), scale * aspect * (u[12] - 0.5f));
        K.n = v3(0, 1, 0); K.mat = 0;
        // target point in barycentric coordinates around the boundary, half of the time within 1e-3 of an edge
        float be = u[13] * 1.2f - 0.1f, ga = u[14] * 1.2f - 0.1f;
        if (u[15] < 0.25f) be = (u[13] - 0.5f) * 2e-3f;
        else if (u[15] < 0.5f) ga = (u[14] - 0.5f) * 2e-3f;
        else if (u[15] < 0.75f) ga = 1.0f - be + (u[14] - 0.5f) * 2e-3f;
        const V3 target = K.a + (K.b - K.a) * be + (K.c - K.a) * ga;
        V3 dir = normalize(v3(u[16] - 0.5f, u[17] - 0.5f, u[18] - 0.5f));
        const V3 nrm = cross(K.b - K.a, K.c - K.a);
        if (u[19] < 0.3f)
        { // grazing: tilt the direction into the triangle's plane
            const float nl = length(nrm);
            if (nl > 0) { const V3 nn = nrm * (1 / nl); dir = normalize(dir - nn * (dot(dir, nn) * (1.0f - 1e-3f * u[16]))); }
        }
        const float dist = scale * powf(10.0f, -2.0f + 5.0f * u[17]);    // 0.01 .. 1000 triangle sizes away
        RayO ray;
        ray.o = target - dir * dist;
        ray.d = (u[18] < 0.8f) ? dir : dir * (0.25f + 3.0f * u[19]);
        const float rec[9] = {K.a.x, K.a.y, K.a.z, K.b.x, K.b.y, K.b.z, K.c.x, K.c.y, K.c.z};
        const rtb_pre::PreTri P = rtb_pre::makePreTri(rec);
        const float dmx = rtb_pre::dirMax(ray.d.x, ray.d.y, ray.d.z);
        float t;
        const bool hit = triIntersect(K, ray, t, nullptr);
        // windows: none, and one that just contains / just excludes the hit
        float lo = -FLT_MAX, hi = FLT_MAX;
        if (hit && u[12] < 0.5f) { lo = t * (1.0f - 1e-3f * (u[11] - 0.3f)); hi = t * (1.0f + 1e-3f * (u[10] - 0.3f)); }
        const bool inWindow = hit && t >= lo && t <= hi;
        const bool rejected = rtb_pre::sureReject(P, ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z, dmx,
                                                  rtb_pre::lowBound(lo), rtb_pre::highBound(hi, FLT_MAX));
        acc += inWindow; cand += !rejected; viol += (inWindow && rejected);
    }
    out[0] = n; out[1] = acc; out[2] = cand; out[3] = viol;
}
// The BOUND itself, not only the decisions built on it.  For random pairs (same generator as above) the four
// determinants are evaluated three ways -- the reference's float association (det3), the cheap FMA form of
// rtb_pretest.h (restated here operation by operation; the decisions derived from this restatement are compared with
// rtb_pre::sureReject so that the two cannot drift apart) and exactly, in binary128 (a product of three floats has
// 72 significant bits) -- and the header's claims are measured:
//   out[0] = max |det_ref - Det| / (gamma_7 S)           claim: <= 1  (S = sum of the absolute triple products)
//   out[1] = max |det'    - Det| / (gamma_5 S)           claim: <= 1
//   out[2] = max |det' - det_ref| / kappa (M, B, G)      claim: <= 1, the header leaves a factor >= 2.6
//   out[3] = max |detT' - detT_ref| / kT                 claim: <= 1
//   out[4] = pairs on which the restated decisions differ from rtb_pre::sureReject (must be 0)
// Pairs whose determinants underflow (|S| < 1e-30) are skipped: the header's floor covers them separately.
extern "C" void rt_oracle_pretest_bound_fuzz(uint64_t seed, long long n, double *out)
{
    typedef __float128 Q;
    const double u = ldexp(1.0, -24), g7 = 7 * u / (1 - 7 * u), g5 = 5 * u / (1 - 5 * u);
    double mRef = 0, mCheap = 0, mKap = 0, mKt = 0;
    long long drift = 0;
#pragma omp parallel for reduction(max : mRef, mCheap, mKap, mKt) reduction(+ : drift) schedule(static)
    for (long long i = 0; i < n; i++)
    {
        uint32_t ctr[4] = {(uint32_t)i, (uint32_t)(i >> 32), 0, 0}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, r[4];
        float v[20];
        for (int b = 0; b < 5; b++)
        {
            ctr[2] = (uint32_t)b;
            philox4x32_10(ctr, key, r);
            for (int q = 0; q < 4; q++) v[4 * b + q] = (float)((r[q] >> 8) * (1.0 / 16777216.0));
        }
        const float scale = powf(10.0f, -3.0f + 7.0f * v[0]), aspect = powf(10.0f, -3.0f * v[1]);
        const float offset = (v[2] < 0.5f) ? 0.0f : powf(10.0f, 5.0f * v[3]) - 1.0f;
        const V3 a = v3(offset + scale * (v[4] - 0.5f), offset * 0.5f + scale * (v[5] - 0.5f), scale * (v[6] - 0.5f) - offset);
        const V3 bb = a + v3(scale * (v[7] - 0.5f), scale * (v[8] - 0.5f), scale * (v[9] - 0.5f));
        const V3 cc = a + v3(scale * aspect * (v[10] - 0.5f), scale * aspect * (v[11] - 0.5f), scale * aspect * (v[12] - 0.5f));
        const float be = v[13] * 1.2f - 0.1f, ga = v[14] * 1.2f - 0.1f;
        const V3 target = a + (bb - a) * be + (cc - a) * ga;
        V3 dir = normalize(v3(v[16] - 0.5f, v[17] - 0.5f, v[18] - 0.5f));
        if (v[19] < 0.3f)
        {
            const V3 nrm = cross(bb - a, cc - a);
            const float nl = length(nrm);
            if (nl > 0) { const V3 nn = nrm * (1 / nl); dir = normalize(dir - nn * (dot(dir, nn) * (1.0f - 1e-3f * v[16]))); }
        }
        const float dist = scale * powf(10.0f, -2.0f + 5.0f * v[17]);
        const V3 o = target - dir * dist;
        const V3 d = (v[18] < 0.8f) ? dir : dir * (0.25f + 3.0f * v[19]);
        const float rec[9] = {a.x, a.y, a.z, bb.x, bb.y, bb.z, cc.x, cc.y, cc.z};
        const rtb_pre::PreTri T = rtb_pre::makePreTri(rec);
        // ---- the float inputs both sides share
        const float e1[3] = {T.e1x, T.e1y, T.e1z}, e2[3] = {T.e2x, T.e2y, T.e2z}, dd[3] = {d.x, d.y, d.z};
        const float b[3] = {T.ax - o.x, T.ay - o.y, T.az - o.z};
        // ---- reference association (Triangle.cpp:25-36, 84-106)
        const float rM = det3(e1[0], e2[0], dd[0], e1[1], e2[1], dd[1], e1[2], e2[2], dd[2]);
        const float rT = det3(e1[0], e2[0], b[0], e1[1], e2[1], b[1], e1[2], e2[2], b[2]);
        const float rB = det3(b[0], e2[0], dd[0], b[1], e2[1], dd[1], b[2], e2[2], dd[2]);
        const float rG = det3(e1[0], b[0], dd[0], e1[1], b[1], dd[1], e1[2], b[2], dd[2]);
        // ---- rtb_pretest.h, operation by operation
        const float px = fmaf(e2[1], dd[2], -(e2[2] * dd[1])), py = fmaf(e2[2], dd[0], -(e2[0] * dd[2])), pz = fmaf(e2[0], dd[1], -(e2[1] * dd[0]));
        const float cM = fmaf(e1[2], pz, fmaf(e1[1], py, e1[0] * px)), cB = fmaf(b[2], pz, fmaf(b[1], py, b[0] * px));
        const float qx = fmaf(b[1], e1[2], -(b[2] * e1[1])), qy = fmaf(b[2], e1[0], -(b[0] * e1[2])), qz = fmaf(b[0], e1[1], -(b[1] * e1[0]));
        const float cT = fmaf(e2[2], qz, fmaf(e2[1], qy, e2[0] * qx)), cGn = fmaf(dd[2], qz, fmaf(dd[1], qy, dd[0] * qx));
        const float bn = fabsf(b[0]) + fabsf(b[1]) + fabsf(b[2]);
        const float dmx = rtb_pre::dirMax(d.x, d.y, d.z);
        const float kap = dmx * fmaf(T.ee, bn, T.a1e), kT = T.a1e * bn;
        // ---- exact
        auto det = [](const float *c0, const float *c1, const float *c2, Q &S) {
            const Q t[6] = {(Q)c0[0] * c1[1] * c2[2], (Q)c1[0] * c2[1] * c0[2], (Q)c2[0] * c0[1] * c1[2],
                            (Q)c2[0] * c1[1] * c0[2], (Q)c0[0] * c2[1] * c1[2], (Q)c1[0] * c0[1] * c2[2]};
            S = 0;
            for (int k = 0; k < 6; k++) S += t[k] < 0 ? -t[k] : t[k];
            return t[0] + t[1] + t[2] - t[3] - t[4] - t[5];
        };
        Q sM, sT, sB, sG;
        const Q xM = det(e1, e2, dd, sM), xT = det(e1, e2, b, sT), xB = det(b, e2, dd, sB), xG = det(e1, b, dd, sG);
        auto ab = [](Q x) { return (double)(x < 0 ? -x : x); };
        const struct { float ref, cheap; Q exact, S; double bound; bool isT; } D[4] = {
            {rM, cM, xM, sM, kap, false}, {rB, cB, xB, sB, kap, false}, {rG, -cGn, xG, sG, kap, false}, {rT, cT, xT, sT, kT, true}};
        for (int k = 0; k < 4; k++)
        {
            const double S = (double)D[k].S;
            if (!(S > 1e-30) || !(S < 1e30)) continue;
            mRef = fmax(mRef, ab((Q)D[k].ref - D[k].exact) / (g7 * S));
            mCheap = fmax(mCheap, ab((Q)D[k].cheap - D[k].exact) / (g5 * S));
            const double diff = fabs((double)D[k].cheap - (double)D[k].ref) / (double)D[k].bound;
            if (D[k].isT) mKt = fmax(mKt, diff); else mKap = fmax(mKap, diff);
        }
        // ---- the restatement above must decide like the header
        const float Lp = rtb_pre::lowBound(-FLT_MAX), Hp = rtb_pre::highBound(dist * 2.0f, FLT_MAX);
        const float m = fabsf(cM);
        const uint32_t sg = rtb_pre::signOf(cM);
        const float Bt = rtb_pre::xorSign(cB, sg), Gt = rtb_pre::xorSign(cGn, sg ^ 0x80000000u), Tt = rtb_pre::xorSign(cT, sg);
        const float mk = m + kap, mlo = m - kap;
        const float hiK = fmaf(RTB_PRE_CH, mk, kap), loK = -fmaf(RTB_PRE_CL, mk, kap);
        const bool mine = ((Bt < loK) | (Bt > hiK) | (Gt < loK) | (Gt > hiK) | (Bt + Gt > hiK + kap) | (Tt + kT < Lp * mlo) | (Tt - kT > Hp * mk)) & (mlo > 0.f);
        drift += mine != rtb_pre::sureReject<true>(T, o.x, o.y, o.z, d.x, d.y, d.z, dmx, Lp, Hp);
    }
    out[0] = mRef; out[1] = mCheap; out[2] = mKap; out[3] = mKt; out[4] = (double)drift;
}

extern "C" void rt_oracle_pretest_mode(int use_nearest) { g_pre_use_nearest = use_nearest; }
extern "C" void rt_oracle_pretest_stats(long long *out)
{
    const long long v[8] = {g_pre.tests, g_pre.candidates, g_pre.updates, g_pre.violations, g_pre.lists, g_pre.lists1, g_pre.lists2, g_pre.unsure};
    memcpy(out, v, sizeof(v));
    g_pre = PretestStats{0, 0, 0, 0, 0, 0, 0, 0};
}
#endif

extern "C" int rt_oracle_run(oracle_job *job)
{
    if (!job || job->preset < 1 || job->preset > 5) return -1;
    if (job->algorithm < 0 || job->algorithm > 6) return -2;
    if (job->width <= 0 || job->height <= 0) return -3;
    Scene s;
    if (!buildPreset(s, job)) return -4;
    if (job->setting != ORACLE_SETTING_PRESET) s.setting = settingOf(job->setting);

    memset(job->stats, 0, sizeof(job->stats));
    job->struct_hash = 0; job->tri_hash = 0; job->prepare_ms = 0;
    const int nTop = (int)s.prims.size();
    job->stats[ORACLE_STAT_N_TOP] = nTop;
    if (s.hasTunnel)
    {
        TunnelO &T = s.tunnel;
        T.exactGrid = job->grid_exact != 0;
        const auto t0 = std::chrono::steady_clock::now();
        if (T.algorithm == 1 || T.algorithm == 2) initGrid(T);
        else if (T.algorithm == 3 || T.algorithm == 4) initKd(T);
        else if (T.algorithm == 5 || T.algorithm == 6) initConvex(T);
        job->prepare_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        job->stats[ORACLE_STAT_N_TRIS] = (int64_t)T.tris.size();
        uint64_t th = H0;
        for (size_t k = 0; k < T.tris.size(); k++)
        {
            const Tri &t = T.tris[k];
            const float v[12] = {t.a.x, t.a.y, t.a.z, t.b.x, t.b.y, t.b.z, t.c.x, t.c.y, t.c.z, t.n.x, t.n.y, t.n.z};
            const int mat = (t.mat == T.tris.back().mat) ? 1 : 0; // the last triangle of a ring is ground
            for (int q = 0; q < 12; q++) hmix(th, fbits(v[q]));
            hmix(th, (uint32_t)mat);
            if (job->tri_out && (int)k < job->tri_cap) memcpy(job->tri_out + 12 * k, v, sizeof(v));
            if (job->tri_mat && (int)k < job->tri_cap) job->tri_mat[k] = mat;
        }
        job->tri_hash = th;
        if (T.algorithm == 1 || T.algorithm == 2)
        {
            uint64_t h = H0;
            hmix(h, 0x47524944u);
            hmix(h, (uint32_t)T.nx); hmix(h, (uint32_t)T.ny); hmix(h, (uint32_t)T.nz);
            hmix(h, fbits(T.origin.x)); hmix(h, fbits(T.origin.y)); hmix(h, fbits(T.origin.z));
            hmix(h, fbits(T.csx)); hmix(h, fbits(T.csy)); hmix(h, fbits(T.csz));
            job->stats[ORACLE_STAT_GRID_X] = T.nx; job->stats[ORACLE_STAT_GRID_Y] = T.ny; job->stats[ORACLE_STAT_GRID_Z] = T.nz;
            for (size_t c = 0; c < T.cells.size(); c++)
            {
                const std::vector<int> &l = T.cells[c];
                if (l.empty()) continue;
                job->stats[ORACLE_STAT_CELLS_NONEMPTY]++;
                job->stats[ORACLE_STAT_CELL_ENTRIES] += (int64_t)l.size();
                if ((int64_t)l.size() > job->stats[ORACLE_STAT_CELL_MAX]) job->stats[ORACLE_STAT_CELL_MAX] = (int64_t)l.size();
                hmix(h, (uint32_t)c); hmix(h, (uint32_t)l.size());
                for (size_t i = 0; i < l.size(); i++) hmix(h, (uint32_t)l[i]);
            }
            job->struct_hash = h;
        }
        else if (T.algorithm == 5 || T.algorithm == 6) job->struct_hash = hashConvex(T);
        else if (T.algorithm == 3 || T.algorithm == 4)
        {
            uint64_t h = H0;
            hmix(h, 0x4b445452u);
            hmix(h, fbits(T.nodes[0].mn.x)); hmix(h, fbits(T.nodes[0].mn.y)); hmix(h, fbits(T.nodes[0].mn.z));
            hmix(h, fbits(T.nodes[0].mx.x)); hmix(h, fbits(T.nodes[0].mx.y)); hmix(h, fbits(T.nodes[0].mx.z));
            hashKd(T, 0, h);
            job->struct_hash = h;
            job->stats[ORACLE_STAT_KD_NODES] = (int64_t)T.nodes.size();
            job->stats[ORACLE_STAT_KD_LEAVES] = T.leaves;
            job->stats[ORACLE_STAT_KD_LEAF_REFS] = T.leafRefs;
            job->stats[ORACLE_STAT_KD_MAX_DEPTH] = T.maxDepth;
        }
    }

    const int width = job->width, height = job->height;
    if (job->hit_id || job->hit_t || job->seq_len || job->seq_hash || job->seq_buf)
    {
        const float dx = 1.0f / height, dy = 1.0f / height;
#pragma omp parallel for schedule(dynamic, 1)
        for (int y = 0; y < height; y++)
        {
            std::vector<int> rec;
            for (int x = 0; x < width; x++)
            {
                const float sx = (x + 0.5f) * dx, sy = 1 - (y + 0.5f) * dy;
                rec.clear();
                Probe pr = {&rec, nullptr};
                const Hit h = sceneIntersect(s, generateRay(s.cam, sx, sy), &pr);
                const size_t p = (size_t)y * width + x;
                if (job->hit_id) job->hit_id[p] = h.hit ? h.id : -1;
                if (job->hit_t) job->hit_t[p] = h.hit ? h.t : -1.0f;
                if (job->seq_len) job->seq_len[p] = (int)rec.size();
                if (job->seq_hash)
                {
                    uint64_t hh = H0;
                    for (size_t i = 0; i < rec.size(); i++) hmix(hh, (uint32_t)rec[i]);
                    job->seq_hash[p] = hh;
                }
                if (job->seq_buf)
                    for (int i = 0; i < job->seq_cap; i++) job->seq_buf[p * job->seq_cap + i] = i < (int)rec.size() ? rec[i] : -1;
            }
        }
    }

    job->n_rays = job->n_tri_tests = job->n_steps = 0;
    job->render_ms = 0;
    if (job->rgb || job->repeat > 0)
    {
        omp_set_num_threads(job->threads > 0 ? job->threads : omp_get_num_procs());
        const int reps = job->repeat > 0 ? job->repeat : 1;
        double best = 1e300;
        Counters total = {0, 0, 0};
        for (int r = 0; r < reps; r++)
        {
            const auto t0 = std::chrono::steady_clock::now();
            std::vector<float> tmp;
            float *dst = r == 0 ? job->rgb : nullptr;
            if (r == 0 && job->rgb8 && !dst) { tmp.resize((size_t)width * height * 3); dst = tmp.data(); }
            render(s, job, dst, &total);
            if (r == 0 && job->rgb8)
                for (size_t i = 0; i < (size_t)width * height * 3; i++)
                { // MainWindow.cpp:305-311: saturate (upper clamp only), then (int)(c * 255)
                    float c = dst[i];
                    c = (c > 1.0f) ? 1.0f : c;
                    job->rgb8[i] = (uint8_t)(int)(c * 255);
                }
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (job->render_ms_all) job->render_ms_all[r] = ms;
            if (ms < best) best = ms;
        }
        job->render_ms = best;
        job->n_rays = total.rays; job->n_tri_tests = total.tris; job->n_steps = total.steps;
    }
    return 0;
}
