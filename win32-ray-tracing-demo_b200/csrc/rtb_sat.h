// rtb_sat.h -- exact triangle / grid-cell overlap test for the grid builders (host and device).
//
// The reference bins a triangle into every cell its bounding box overlaps; the exact test it also carries,
// Triangle::intersectWithGrid (reference Triangle.cpp:123-199: separating axes -- the three box axes, the triangle
// normal, the nine box-edge x triangle-edge cross products), is compiled out at its only call site
// (Tunnel.cpp:435-445, with the note that the simple way builds ~8x faster and traverses ~20 % slower).  This is that
// test, expression by expression in float, as an OPTION of both grid builders (Tunnel::exactGridBinning,
// rtb_flat_scene::grid_build_exact): fewer references per cell, same hits.  Plain C++ so that the host builder
// (host/rt_tunnel.cpp, -ffp-contract=off) and the device builder (rtb_build_grid.cuh, -fmad=false) evaluate the same
// operations and produce the same lists.
#pragma once
#include <float.h>

#if defined(__CUDACC__)
#define RTB_SAT_FN __host__ __device__ __forceinline__
#else
#define RTB_SAT_FN inline
#endif

namespace rtb_sat {

// projections of the 8 cell corners and of the 3 vertices on `axis` do not separate?  (intersectOnAxis, Triangle.cpp:144-149;
// getMin / getMax use std::min / std::max: `v < m ? v : m`, kept literally)
RTB_SAT_FN bool overlapOnAxis(const float *t, const float *lo, const float *hi, float ax, float ay, float az)
{
    float gmin = FLT_MAX, gmax = -FLT_MAX;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < 8; c++)
    { // corner order of Triangle.cpp:155-162: (0,0,0) (0,0,z) (0,y,0) (0,y,z) (x,0,0) (x,0,z) (x,y,0) (x,y,z)
        const float X = (c & 4) ? hi[0] : lo[0], Y = (c & 2) ? hi[1] : lo[1], Z = (c & 1) ? hi[2] : lo[2];
        const float d = ax * X + ay * Y + az * Z; // Vector::dot(Point), Vector.cpp:74-77
        gmin = d < gmin ? d : gmin;
        gmax = gmax < d ? d : gmax;
    }
    float tmin = FLT_MAX, tmax = -FLT_MAX;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int v = 0; v < 3; v++)
    {
        const float d = ax * t[3 * v] + ay * t[3 * v + 1] + az * t[3 * v + 2];
        tmin = d < tmin ? d : tmin;
        tmax = tmax < d ? d : tmax;
    }
    if (gmin > tmax) return false;
    if (gmax < tmin) return false;
    return true;
}

// t: a, b, c, normal (12 floats); cell: position and size per axis (the reference builds the cell as
// Grid(origin + Vector(i * size, j * size, k * size), Vector(size, size, size)) -- the caller passes exactly those floats)
RTB_SAT_FN bool triangleOverlapsCell(const float *t, const float pos[3], const float size[3])
{
    // gridPoints: pos + Vector(0 | size.x, 0 | size.y, 0 | size.z), Point::operator+ (Point.cpp:20-23)
    const float lo[3] = {pos[0] + 0.0f, pos[1] + 0.0f, pos[2] + 0.0f};
    const float hi[3] = {pos[0] + size[0], pos[1] + size[1], pos[2] + size[2]};
    if (!overlapOnAxis(t, lo, hi, 1, 0, 0)) return false;
    if (!overlapOnAxis(t, lo, hi, 0, 1, 0)) return false;
    if (!overlapOnAxis(t, lo, hi, 0, 0, 1)) return false;
    if (!overlapOnAxis(t, lo, hi, t[9], t[10], t[11])) return false; // the stored normal
    // edges: Vector(a, b) = b - a, Vector(b, c), Vector(c, a) (Vector.cpp:11-16); box edges x edge with Vector::cross
    // (y * b.z - z * b.y, z * b.x - x * b.z, x * b.y - y * b.x), written out with the unit vectors' zeros and ones
    const float e[3][3] = {{t[3] - t[0], t[4] - t[1], t[5] - t[2]}, {t[6] - t[3], t[7] - t[4], t[8] - t[5]}, {t[0] - t[6], t[1] - t[7], t[2] - t[8]}};
    for (int k = 0; k < 3; k++) // boxEdge1 = (1, 0, 0)
        if (!overlapOnAxis(t, lo, hi, 0.0f * e[k][2] - 0.0f * e[k][1], 0.0f * e[k][0] - 1.0f * e[k][2], 1.0f * e[k][1] - 0.0f * e[k][0])) return false;
    for (int k = 0; k < 3; k++) // boxEdge2 = (0, 1, 0)
        if (!overlapOnAxis(t, lo, hi, 1.0f * e[k][2] - 0.0f * e[k][1], 0.0f * e[k][0] - 0.0f * e[k][2], 0.0f * e[k][1] - 1.0f * e[k][0])) return false;
    for (int k = 0; k < 3; k++) // boxEdge3 = (0, 0, 1)
        if (!overlapOnAxis(t, lo, hi, 0.0f * e[k][2] - 1.0f * e[k][1], 1.0f * e[k][0] - 0.0f * e[k][2], 0.0f * e[k][1] - 0.0f * e[k][0])) return false;
    return true;
}

} // namespace rtb_sat
