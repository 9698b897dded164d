// rt_tunnel.cpp -- tunnel tessellation and the HOST accelerator builders.
//
// The builders produce exactly the structures the reference builds (same cells, same trees, same
// per-cell / per-leaf triangle order -- checked through the canonical structure hash against the
// reference's own build), but emit them directly in the flattened device layout of include/rtb.h
// and use asymptotically cheaper algorithms:
//   * grids: (cell, triangle) pairs are generated in triangle order and sorted by cell, giving the
//     CSR cell lists and the sparse cell directory without ever materialising the 64 M cell array
//     of the 400^3 flat grid (reference Tunnel.cpp:346-465 allocates one std::vector per cell);
//   * SAH: the left/right counts of all 99 candidate planes of an axis come from one binary search
//     per triangle into the candidate array plus a prefix sum (O(N log 99) instead of the
//     reference's O(99 N) rescans, Tunnel.cpp:739-767); subtrees are built as OpenMP tasks.
// Float arithmetic that decides structure (bounds, cell indices, candidate positions, SAH costs)
// follows the reference expression by expression; compile with -ffp-contract=off.
#include "rt.h"
#include "../csrc/rtb_sat.h"

#include <algorithm>
#include <cfloat>
#include <cstring>
#include <omp.h>

namespace rt {

// ---- tessellation (reference TunnelGenerator.cpp) ----------------------------------------------
namespace {

struct Ring { std::vector<Point> v; };

// Reference TunnelGenerator.cpp:6-176: for each quad (A,B on the front ring, C,D on the rear ring)
// decide which diagonal keeps the segment's hull convex.  0 = BC, 1 = AD, 2 = both, 3 = invalid.
bool classifyQuads(const Ring &front, const Ring &rear, std::vector<unsigned char> &conn)
{
    const float tol = 0.001f;
    const size_t n = front.v.size();
    conn.assign(n, 3);
    bool convex = true;
    // every other vertex of both rings must lie on or behind the plane through `base` with normal `nrm`
    auto hullSide = [&](const Vector &nrm, const Point &base, size_t j, size_t k) {
        for (size_t m = 0; m < n; m++)
        {
            if (m == j || m == k) continue;
            Vector e(base, front.v[m]), f(base, rear.v[m]);
            e.norm();
            f.norm();
            if (nrm.dot(e) > tol || nrm.dot(f) > tol) return false;
        }
        return true;
    };
    auto unit = [](const Point &s, const Point &e) { Vector v(s, e); v.norm(); return v; };
    for (size_t j = 0; j < n; j++)
    {
        const size_t k = (j + 1) % n;
        const Point &A = front.v[j], &B = front.v[k], &C = rear.v[j], &D = rear.v[k];
        Vector nCBA = Vector(C, B).cross(Vector(B, A)); nCBA.norm();
        Vector nCDB = Vector(C, D).cross(Vector(D, B)); nCDB.norm();
        bool bc = !(nCBA.dot(unit(A, D)) > tol);
        if (bc) bc = hullSide(nCBA, A, j, k);
        if (nCDB.dot(unit(C, A)) > tol) bc = false;
        if (bc) bc = hullSide(nCDB, C, j, k);
        Vector nADB = Vector(A, D).cross(Vector(D, B)); nADB.norm();
        Vector nACD = Vector(A, C).cross(Vector(C, D)); nACD.norm();
        bool ad = !(nADB.dot(unit(A, C)) > tol);
        if (ad) ad = hullSide(nADB, A, j, k);
        if (nACD.dot(unit(A, B)) > tol) ad = false;
        if (ad) ad = hullSide(nACD, A, j, k);
        conn[j] = bc ? (ad ? 2 : 0) : (ad ? 1 : 3);
        if (!bc && !ad) convex = false;
    }
    return convex;
}

TunnelTriangle makeTriangle(const Point &a, const Point &b, const Point &c, int material)
{
    TunnelTriangle t;
    t.a = a; t.b = b; t.c = c;
    t.normal = Vector(a, b).cross(Vector(b, c));
    t.normal.norm(); // Triangle.cpp:17-23
    t.material = material;
    return t;
}

} // namespace

bool TunnelGenerator::create(float rectWidth, float rectHeight, float archHeight, float pathRadius, float pathAngle,
                             int archSegments, int pathSegments, GeometrySet &scene, Ptr<Material> groundMaterial,
                             Ptr<Material> wallMaterial, Tunnel::Algorithm algorithm)
{
    Tunnel *tunnel = new Tunnel();
    tunnel->gridOnDevice = Tunnel::gridOnDeviceDefault;
    tunnel->exactGridBinning = Tunnel::exactGridBinningDefault;
    tunnel->kdOnDevice = Tunnel::kdOnDeviceDefault;
    tunnel->height = rectHeight + archHeight;
    tunnel->width = rectWidth;
    tunnel->algorithm = algorithm;
    tunnel->groundMaterial = groundMaterial;
    tunnel->wallMaterial = wallMaterial;

    // cross-section: rectangle + half ellipse, counter-clockwise from the lower right corner (226-244)
    std::vector<Point> section;
    section.push_back(Point(rectWidth * 0.5f, 0.0f, 0.0f));
    for (int i = 0; i <= archSegments; i++)
    {
        const float angle = PI * i / archSegments;
        section.push_back(Point(std::cos(angle) * rectWidth * 0.5f, std::sin(angle) * archHeight + rectHeight, 0.0f));
    }
    section.push_back(Point(-rectWidth * 0.5f, 0.0f, 0.0f));

    auto pathPoint = [&](int i) {
        const float theta = pathAngle * i / pathSegments;
        return Point(pathRadius * (1.0f - std::cos(theta)), 0.0f, -pathRadius * std::sin(theta));
    };
    tunnel->path.push_back(pathPoint(0));
    for (int i = 0; i < pathSegments; i++) tunnel->path.push_back(pathPoint(i + 1));
    tunnel->surface.assign(pathSegments, std::vector<TunnelTriangle>());

    // PerformanceTest's generator (src/PerformanceTest/TunnelGenerator.cpp:243-279) turns the ring at path vertex i
    // by the angle of the averaged directions of the two segments that meet there
    std::vector<Vector> ringNormal;
    if (performanceTestVariant)
    {
        const int N = pathSegments;
        for (int i = 1; i <= N; i++)
        {
            const Vector dir = Vector(tunnel->path[i - 1], tunnel->path[i]).norm();
            const Vector prevDir = i > 1 ? Vector(tunnel->path[i - 2], tunnel->path[i - 1]).norm() : dir;
            const Vector nextDir = i < N ? Vector(tunnel->path[i], tunnel->path[i + 1]).norm() : dir;
            if (i == 1) ringNormal.push_back(((prevDir + dir) * 0.5f).norm());
            ringNormal.push_back(((dir + nextDir) * 0.5f).norm());
        }
        tunnel->performanceTestBuilders = true;
        tunnel->ringNormals = ringNormal;
    }
    tunnel->crossSection = section;

#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < pathSegments; i++)
    {
        const Point p1 = pathPoint(i), p2 = pathPoint(i + 1);
        const float delta = (i == pathSegments - 1) ? 0 : pathAngle / pathSegments;
        float angle1 = Vector(0, 0, -1).angleTo(Vector(p1, p2));
        float angle2 = angle1 + delta;
        if (performanceTestVariant)
        {
            angle1 = Vector(0, 0, -1).angleTo(ringNormal[i]);
            angle2 = Vector(0, 0, -1).angleTo(ringNormal[i + 1]);
        }
        // place the section at both ends of the segment: rotate about y, then translate (296-307)
        auto place = [&](float angle, const Point &at, Ring &ring) {
            for (const Point &p : section)
                ring.v.push_back(Point(p.x * std::cos(angle) - p.z * std::sin(angle), p.y,
                                       p.x * std::sin(angle) + p.z * std::cos(angle)) + Vector(Point(0, 0, 0), at));
        };
        Ring front, rear;
        place(angle1, p1, front);
        place(angle2, p2, rear);
        std::vector<unsigned char> conn;
        if (!classifyQuads(front, rear, conn)) continue; // a non-convex segment contributes no triangles (313-317)
        const size_t n = section.size();
        for (size_t j = 0; j < n; j++)
        {
            const Point &A = front.v[j], &B = front.v[(j + 1) % n], &C = rear.v[j], &D = rear.v[(j + 1) % n];
            const int mat = (j == n - 1) ? 1 : 0; // the closing quad is the ground (331)
            if (conn[j] == 1)
            {
                tunnel->surface[i].push_back(makeTriangle(A, C, D, mat));
                tunnel->surface[i].push_back(makeTriangle(A, D, B, mat));
            }
            else
            {
                tunnel->surface[i].push_back(makeTriangle(C, D, B, mat));
                tunnel->surface[i].push_back(makeTriangle(C, B, A, mat));
            }
        }
    }
    scene.add(tunnel);
    return true;
}

// ---- Tunnel ------------------------------------------------------------------------------------
bool Tunnel::gridOnDeviceDefault = false;
bool Tunnel::exactGridBinningDefault = false;
bool Tunnel::kdOnDeviceDefault = false;

size_t Tunnel::triangleCount() const
{
    size_t n = 0;
    for (const auto &s : surface) n += s.size();
    return n;
}

void Tunnel::collect(std::vector<TunnelTriangle> &flat) const
{
    flat.clear();
    flat.reserve(triangleCount());
    for (const auto &s : surface) flat.insert(flat.end(), s.begin(), s.end());
}

namespace {

struct Bounds { float mn[3], mx[3]; };

inline void triangleBounds(const TunnelTriangle &t, float mn[3], float mx[3])
{ // reference Triangle.cpp:201-237
    for (int a = 0; a < 3; a++)
    {
        mn[a] = std::min(std::min(t.a[a], t.b[a]), t.c[a]);
        mx[a] = std::max(std::max(t.a[a], t.b[a]), t.c[a]);
    }
}

Bounds sceneBounds(const std::vector<TunnelTriangle> &tris)
{ // reference Tunnel.cpp:351-370
    Bounds b = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
    for (const TunnelTriangle &t : tris)
    {
        float mn[3], mx[3];
        triangleBounds(t, mn, mx);
        for (int a = 0; a < 3; a++) { b.mn[a] = std::min(b.mn[a], mn[a]); b.mx[a] = std::max(b.mx[a], mx[a]); }
    }
    return b;
}

} // namespace

void Tunnel::init()
{ // reference Tunnel.cpp:116-133
    std::vector<TunnelTriangle> tris;
    collect(tris);
    gridWords_.clear(); gridCellStart_.clear(); gridCellTris_.clear(); kdNodes_.clear(); kdLeafTris_.clear();
    cxFrames_.clear(); cxEdges_.clear(); cxCellStatus_.clear(); cxCellRange_.clear(); cxOrder_.clear();
    stats = BuildStats();
    if ((algorithm == RegularGrid || algorithm == FlatGrid) && gridOnDevice) { /* built by rtb_scene_upload */ }
    else if (algorithm == RegularGrid || algorithm == FlatGrid) initGrid(tris);
    else if (algorithm == KdTreeSAH && kdOnDevice && !performanceTestBuilders) { /* built by rtb_scene_upload */ }
    else if (algorithm == KdTreeStandard || algorithm == KdTreeSAH) initKdTree(tris);
    else if (algorithm == Convex || algorithm == ConvexSimple) initConvex();
    built_ = true;
    invalidate();
}

// ---- convex accelerator tables (reference PerformanceTest/ConvexAcc.cpp:181-271) ------------------
// Everything the device walk needs is a table built here with the host's libm, evaluated exactly as the
// reference evaluates it (per intersect call, from the same inputs): per path vertex the rotation that maps
// its ring onto the plane z = 0; the half-plane form of the cross-section edges; a 100 x 100 lookup over the
// cross-section's bounding rectangle saying whether a cell is inside, outside or which edges cut it; and, for
// Convex, the order in which a wall segment's triangles are tried for each of 100 heights x 360 directions.
void Tunnel::initConvex()
{
    const size_t n = crossSection.size(), nPath = path.size();
    if (n < 3 || nPath < 2) return; // flatten() emits no tables; the upload rejects the scene
    // The two reference programs differ in three details (RayTracingOpt/Tunnel.cpp:135-344 vs PerformanceTest/ConvexAcc.cpp):
    // ring normals (PerformanceTest's generator averages the adjoining segments; RayTracingOpt takes the direction of
    // the segment that starts at the vertex, here), table resolution (100 vs 400) and which edges a Partial cell
    // tests (a precomputed range vs all of them); plus the rounding of the direction bin (flatten: cx_round_bins).
    const bool pt = performanceTestBuilders;
    if (!pt)
    {
        ringNormals.clear();
        for (size_t i = 0; i < nPath; i++)
        {
            if (i + 1 < nPath) ringNormals.push_back(Vector(path[i], path[i + 1]).norm());
            else ringNormals.push_back(Vector(path[i], path[i] + Vector(path[i - 1], path[i])).norm());
        }
    }
    if (ringNormals.size() != nPath) return;
    cxFrames_.resize(nPath * 8);
    for (size_t i = 0; i < nPath; i++)
    {
        const Vector &nv = ringNormals[i];
        const float theta = PI - std::atan2(nv.x, nv.z); // ConvexAcc.cpp:14
        float *f = &cxFrames_[i * 8];
        f[0] = path[i].x; f[1] = path[i].y; f[2] = path[i].z;
        f[3] = nv.x; f[4] = nv.y; f[5] = nv.z;
        f[6] = std::cos(theta); f[7] = std::sin(theta);
    }
    cxEdges_.resize(n * 3);
    for (size_t e = 0; e < n; e++)
    { // inside <=> A*x + B*y + C > 0 for the counter-clockwise polygon
        const Point &p1 = crossSection[e], &p2 = crossSection[(e + 1) % n];
        cxEdges_[3 * e] = p1.y - p2.y;
        cxEdges_[3 * e + 1] = p2.x - p1.x;
        cxEdges_[3 * e + 2] = p1.x * p2.y - p2.x * p1.y;
    }
    auto side = [&](size_t e, float x, float y) { return cxEdges_[3 * e] * x + cxEdges_[3 * e + 1] * y + cxEdges_[3 * e + 2]; };
    auto inside = [&](float x, float y) {
        for (size_t e = 0; e < n; e++)
            if (side(e, x, y) < 0.0001f) return false;
        return true;
    };
    const int R = pt ? 100 : 400;
    cxTable_ = R;
    cxCellStatus_.assign((size_t)R * R, 2);
    cxCellRange_.assign((size_t)R * R * 2, -1);
    const float cellW = width / (R - 1.0f), cellH = height / (R - 1.0f);
    for (int i = 0; i < R; i++)
        for (int j = 0; j < R; j++)
        {
            const float cx = i * cellW - width / 2, cy = j * cellH;
            const float x0 = cx + -cellW / 2, x1 = cx + cellW / 2, y0 = cy + -cellH / 2, y1 = cy + cellH / 2;
            const float xs[4] = {x0, x1, x0, x1}, ys[4] = {y0, y0, y1, y1};
            int cornersInside = 0;
            for (int c = 0; c < 4; c++) cornersInside += inside(xs[c], ys[c]) ? 1 : 0;
            const size_t cell = (size_t)i * R + j;
            if (cornersInside == 4) { cxCellStatus_[cell] = 0; continue; }
            if (cornersInside == 0) { cxCellStatus_[cell] = 2; continue; }
            cxCellStatus_[cell] = 1;
            if (!pt) { cxCellRange_[cell * 2] = 0; cxCellRange_[cell * 2 + 1] = (int16_t)(n - 1); continue; }
            int first = SHRT_MAX, last = SHRT_MIN;
            for (size_t e = 0; e < n; e++)
            {
                float sd[4];
                for (int c = 0; c < 4; c++) sd[c] = side(e, xs[c], ys[c]);
                // the reference compares against the double literal 0.0001 here (ConvexAcc.cpp:152-153)
                const bool allLeft = sd[0] > 0.0001 && sd[1] > 0.0001 && sd[2] > 0.0001 && sd[3] > 0.0001;
                const bool allRight = sd[0] < -0.0001 && sd[1] < -0.0001 && sd[2] < -0.0001 && sd[3] < -0.0001;
                if (allLeft || allRight) continue;
                const Point &a = crossSection[e], &b = crossSection[(e + 1) % n];
                if ((a.x > x1 && b.x > x1) || (a.x < x0 && b.x < x0) || (a.y > y1 && b.y > y1) || (a.y < y0 && b.y < y0)) continue;
                first = std::min(first, (int)e);
                last = std::max(last, (int)e);
            }
            cxCellRange_[cell * 2] = (int16_t)first;
            cxCellRange_[cell * 2 + 1] = (int16_t)last;
        }
    if (algorithm != Convex) return;
    // order of the edges by their angular distance from the 2D ray (0, height*(y+0.5)/100) + direction iAngle degrees;
    // the reference fills a multimap (equal keys keep insertion order) = a stable sort by that distance
    cxOrder_.resize((size_t)100 * 360 * 2 * n);
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < 100; y++)
    {
        std::vector<std::pair<float, int>> keyed(n);
        for (int iAngle = 0; iAngle < 360; iAngle++)
        {
            const float ty = height * (y + 0.5f) / 100.0f;
            const float target = iAngle / 360.0f * PI * 2;
            const float vx = std::cos(target), vy = std::sin(target);
            for (size_t e = 0; e < n; e++)
            {
                const Point &p1 = crossSection[e], &p2 = crossSection[(e + 1) % n];
                const float ax = p1.x - 0, ay = p1.y - ty, bx = p2.x - 0, by = p2.y - ty;
                float delta;
                if (ax * vy - ay * vx > 0 && vx * by - vy * bx > 0) delta = 0; // the ray passes between the edge's end points
                else
                {
                    const float mx = (p1.x + p2.x) / 2, my = (p1.y + p2.y) / 2;
                    const float angle = std::atan2(my - ty, mx - 0.0f);
                    delta = target - angle;
                    delta = (delta > PI) ? delta - 2 * PI : delta;
                    delta = std::fabs(delta);
                }
                keyed[e] = std::make_pair(delta, (int)e);
            }
            std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<float, int> &a, const std::pair<float, int> &b) { return a.first < b.first; });
            uint16_t *row = &cxOrder_[((size_t)y * 360 + iAngle) * 2 * n];
            for (size_t k = 0; k < n; k++)
            {
                row[2 * k] = (uint16_t)(keyed[k].second * 2);
                row[2 * k + 1] = (uint16_t)(keyed[k].second * 2 + 1);
            }
        }
    }
}

void Tunnel::initGrid(const std::vector<TunnelTriangle> &tris)
{ // reference Tunnel.cpp:346-465
    const Bounds b = sceneBounds(tris);
    const float width = b.mx[0] - b.mn[0], height = b.mx[1] - b.mn[1], depth = b.mx[2] - b.mn[2];
    const int R = gridResolution;
    if (algorithm == RegularGrid)
    { // the longest dimension is cut into R pieces; cubic cells
        const float maxLength = std::max(std::max(width, height), depth);
        const float size = maxLength / (R - 1);
        gridCell_[0] = gridCell_[1] = gridCell_[2] = size;
        for (int a = 0; a < 3; a++) gridOrigin_[a] = b.mn[a] - size / 2;
        gridDims_[0] = (int)(width / size + 1.5f);
        gridDims_[1] = (int)(height / size + 1.5f);
        gridDims_[2] = (int)(depth / size + 1.5f);
    }
    else
    { // every dimension is cut into R pieces; anisotropic cells
        gridCell_[0] = width / (R - 1); gridCell_[1] = height / (R - 1); gridCell_[2] = depth / (R - 1);
        for (int a = 0; a < 3; a++) { gridOrigin_[a] = b.mn[a] - gridCell_[a] / 2; gridDims_[a] = R; }
    }
    const int ny = gridDims_[1], nz = gridDims_[2];
    const int64_t cells = (int64_t)gridDims_[0] * ny * nz;

    // (cell, running entry number) keys in triangle order; sorting them groups by cell and keeps
    // the triangle order inside a cell (what push_back in triangle order gives the reference)
    std::vector<uint64_t> keys;
    std::vector<uint32_t> owner;
    for (size_t m = 0; m < tris.size(); m++)
    {
        float mn[3], mx[3];
        triangleBounds(tris[m], mn, mx);
        int lo[3], hi[3];
        for (int a = 0; a < 3; a++)
        {
            lo[a] = (int)((mn[a] - gridOrigin_[a]) / gridCell_[a]);
            hi[a] = (int)((mx[a] - gridOrigin_[a]) / gridCell_[a]);
        }
        for (int i = lo[0]; i <= hi[0]; i++)
            for (int j = lo[1]; j <= hi[1]; j++)
                for (int k = lo[2]; k <= hi[2]; k++)
                {
                    if (exactGridBinning)
                    { // the cell the reference's disabled branch constructs (Tunnel.cpp:437): origin + Vector(i * size, j * size, k * size)
                        const float pos[3] = {gridOrigin_[0] + i * gridCell_[0], gridOrigin_[1] + j * gridCell_[1], gridOrigin_[2] + k * gridCell_[2]};
                        const TunnelTriangle &T = tris[m];
                        const float rec[12] = {T.a.x, T.a.y, T.a.z, T.b.x, T.b.y, T.b.z, T.c.x, T.c.y, T.c.z, T.normal.x, T.normal.y, T.normal.z};
                        if (!rtb_sat::triangleOverlapsCell(rec, pos, gridCell_)) continue;
                    }
                    const uint64_t cell = (uint64_t)(((int64_t)i * ny + j) * nz + k);
                    keys.push_back((cell << 32) | (uint64_t)owner.size());
                    owner.push_back((uint32_t)m);
                }
    }
    std::sort(keys.begin(), keys.end());

    gridWords_.assign((size_t)((cells + 31) / 32), rtb_cellword{0, 0});
    gridCellStart_.clear();
    gridCellTris_.resize(keys.size());
    uint64_t prev = ~0ull;
    int64_t run = 0;
    for (size_t e = 0; e < keys.size(); e++)
    {
        const uint64_t cell = keys[e] >> 32;
        if (cell != prev)
        {
            gridCellStart_.push_back((uint32_t)e);
            gridWords_[cell >> 5].bits |= 1u << (cell & 31);
            prev = cell;
            run = 0;
        }
        run++;
        if (run > stats.cellMax) stats.cellMax = run;
        gridCellTris_[e] = owner[(uint32_t)keys[e]];
    }
    gridCellStart_.push_back((uint32_t)keys.size());
    uint32_t rank = 0;
    for (rtb_cellword &w : gridWords_) { w.rank = rank; rank += (uint32_t)__builtin_popcount(w.bits); }
    stats.gridX = gridDims_[0]; stats.gridY = gridDims_[1]; stats.gridZ = gridDims_[2];
    stats.cellsNonEmpty = (int64_t)gridCellStart_.size() - 1;
    stats.cellEntries = (int64_t)keys.size();
}

// ---- k-d tree ----------------------------------------------------------------------------------
namespace {

struct BuildNode
{
    int axis = 3;
    float split = 0;
    BuildNode *left = nullptr, *right = nullptr;
    std::vector<uint32_t> list; // leaf only
    ~BuildNode() { delete left; delete right; }
};

struct KdBuilder
{
    const std::vector<TunnelTriangle> &tris;
    bool sah, sweep;
    int leafSize, maxDepth, candidates;
    std::vector<float> lo[3], hi[3], centroid[3]; // per-triangle extent / centroid per axis

    KdBuilder(const std::vector<TunnelTriangle> &t, bool sah, bool sweep, int leafSize, int maxDepth, int candidates)
        : tris(t), sah(sah), sweep(sweep), leafSize(leafSize), maxDepth(maxDepth), candidates(candidates)
    {
        for (int a = 0; a < 3; a++)
        {
            lo[a].resize(t.size()); hi[a].resize(t.size()); centroid[a].resize(t.size());
            for (size_t i = 0; i < t.size(); i++)
            {
                lo[a][i] = std::min(std::min(t[i].a[a], t[i].b[a]), t[i].c[a]);
                hi[a][i] = std::max(std::max(t[i].a[a], t[i].b[a]), t[i].c[a]);
                centroid[a][i] = (t[i].a[a] + t[i].b[a] + t[i].c[a]) / 3; // Tunnel.cpp:523
            }
        }
    }

    // reference Tunnel.cpp:649-669: sort by centroid, split at the centroid of the middle element.
    // The sort reorders the node's own list, and the children inherit that order.
    float splitMedian(int axis, std::vector<uint32_t> &list) const
    {
        const std::vector<float> &c = centroid[axis];
        std::sort(list.begin(), list.end(), [&c](uint32_t p, uint32_t q) { return c[p] < c[q]; });
        return c[list[list.size() / 2]];
    }

    // reference Tunnel.cpp:671-784.  For candidate plane s: leftCount = #{triangles with a vertex < s}
    // = #{lo < s}; rightCount = #{triangles with a vertex >= s} = #{hi >= s}.  Candidates are
    // non-decreasing in i, so each triangle contributes to a suffix (left) / prefix (right) of them.
    float splitSAH(const float mn[3], const float mx[3], const std::vector<uint32_t> &list, int &bestAxis) const
    {
        float minSAH = FLT_MAX, minSplit = 0;
        const int N = candidates;
        std::vector<float> cand(N);
        std::vector<int> leftFrom(N + 1), rightTo(N + 1);
        for (int axis = 0; axis < 3; axis++)
        {
            for (int i = 1; i < N; i++) cand[i] = mn[axis] + (mx[axis] - mn[axis]) * i / N;
            std::fill(leftFrom.begin(), leftFrom.end(), 0);
            std::fill(rightTo.begin(), rightTo.end(), 0);
            const float *cb = cand.data() + 1, *ce = cand.data() + N;
            for (uint32_t t : list)
            {
                // first candidate index with cand > lo  -> counts as "left" for every candidate from there on
                leftFrom[(int)(std::upper_bound(cb, ce, lo[axis][t]) - cand.data())]++;
                // first candidate index with cand > hi  -> counts as "right" for every candidate before it
                rightTo[(int)(std::upper_bound(cb, ce, hi[axis][t]) - cand.data())]++;
            }
            const int nextAxis = (axis + 1) % 3, prevAxis = (axis + 2) % 3;
            const float height = mx[nextAxis] - mn[nextAxis], depth = mx[prevAxis] - mn[prevAxis];
            int leftCount = 0, rightCount = (int)list.size();
            for (int i = 1; i < N; i++)
            {
                leftCount += leftFrom[i];
                rightCount -= rightTo[i];
                const float leftWidth = cand[i] - mn[axis], rightWidth = mx[axis] - cand[i];
                const float cost = (leftWidth * height + leftWidth * depth + height * depth) * leftCount +
                                   (rightWidth * height + rightWidth * depth + height * depth) * rightCount;
                if (cost < minSAH) { minSAH = cost; minSplit = cand[i]; bestAxis = axis; } // first strict minimum wins
            }
        }
        return minSplit;
    }

    // PerformanceTest/KdTreeAcc.cpp:177-274: the exact SAH, cost = 1 + 1.5 * ((SAL/SA) * NL + (SAR/SA) * (NR + NP)),
    // evaluated at every bounding-box boundary of the node's triangles, first strict minimum wins (axis 0 -> 2,
    // positions ascending).  The reference sorts one list of {End, Planar, Start} events per axis; the counts it
    // derives at a position p are  PE = #{hi == p}, PS = #{lo == p} over the triangles with lo < hi, and
    // PP = #{lo == hi == p}, so three sorted position arrays merged in step give the same sweep.
    float splitSAHSweep(const float mn[3], const float mx[3], const std::vector<uint32_t> &list, int &bestAxis, float &minSAH) const
    {
        minSAH = FLT_MAX;
        float minPosition = 0;
        std::vector<float> starts, ends, planars;
        for (int axis = 0; axis < 3; axis++)
        {
            starts.clear(); ends.clear(); planars.clear();
            for (uint32_t t : list)
            {
                if (lo[axis][t] == hi[axis][t]) planars.push_back(lo[axis][t]);
                else { starts.push_back(lo[axis][t]); ends.push_back(hi[axis][t]); }
            }
            std::sort(starts.begin(), starts.end());
            std::sort(ends.begin(), ends.end());
            std::sort(planars.begin(), planars.end());
            const int nextAxis = (axis + 1) % 3, prevAxis = (axis + 2) % 3;
            const float width = mx[axis] - mn[axis], height = mx[nextAxis] - mn[nextAxis], depth = mx[prevAxis] - mn[prevAxis];
            const float SA = width * height + width * depth + height * depth;
            size_t is = 0, ie = 0, ip = 0;
            int NL = 0, NR = (int)list.size();
            while (is < starts.size() || ie < ends.size() || ip < planars.size())
            {
                float position = FLT_MAX;
                if (is < starts.size()) position = std::min(position, starts[is]);
                if (ie < ends.size()) position = std::min(position, ends[ie]);
                if (ip < planars.size()) position = std::min(position, planars[ip]);
                int PS = 0, PE = 0, PP = 0;
                while (ie < ends.size() && ends[ie] == position) { PE++; ie++; }
                while (ip < planars.size() && planars[ip] == position) { PP++; ip++; }
                while (is < starts.size() && starts[is] == position) { PS++; is++; }
                const int NP = PP;
                NR -= PP; NR -= PE;
                const float leftWidth = position - mn[axis], rightWidth = mx[axis] - position;
                const float SAL = leftWidth * height + leftWidth * depth + height * depth;
                const float SAR = rightWidth * height + rightWidth * depth + height * depth;
                const float cost = 1 + 1.5f * ((SAL / SA) * NL + SAR / SA * (NR + NP));
                if (cost < minSAH) { minSAH = cost; minPosition = position; bestAxis = axis; }
                NL += PS; NL += PP;
            }
        }
        return minPosition;
    }

    void build(BuildNode *node, std::vector<uint32_t> &list, const float mn[3], const float mx[3], int depth) const
    { // reference Tunnel.cpp:546-638; PerformanceTest/KdTreeAcc.cpp:38-144 when `sweep`
        if ((int)list.size() <= leafSize || depth > maxDepth)
        {
            node->axis = 3;
            node->list = list;
            return;
        }
        int axis = depth % 3;
        float split;
        if (sah && sweep)
        {
            float cost;
            split = splitSAHSweep(mn, mx, list, axis, cost);
            if (cost > 1.5f * list.size())
            { // automatic termination (KdTreeAcc.cpp:77-88): splitting would cost more than testing the whole list
                node->axis = 3;
                node->list = list;
                return;
            }
        }
        else if (sah) split = splitSAH(mn, mx, list, axis);
        else split = splitMedian(axis, list);
        node->axis = axis;
        node->split = split;
        node->left = new BuildNode();
        node->right = new BuildNode();
        std::vector<uint32_t> leftPart, rightPart;
        for (uint32_t t : list)
        { // straddlers go to both sides (619-634)
            if (lo[axis][t] < split) leftPart.push_back(t);
            if (hi[axis][t] >= split) rightPart.push_back(t);
        }
        std::vector<uint32_t>().swap(list);
        float lmx[3] = {mx[0], mx[1], mx[2]}, rmn[3] = {mn[0], mn[1], mn[2]};
        lmx[axis] = split;
        rmn[axis] = split;
        const bool spawn = leftPart.size() + rightPart.size() > 2048;
        float lmn[3] = {mn[0], mn[1], mn[2]}, rmx[3] = {mx[0], mx[1], mx[2]};
#pragma omp task default(shared) firstprivate(lmn, lmx, depth) if (spawn)
        build(node->left, leftPart, lmn, lmx, depth + 1);
#pragma omp task default(shared) firstprivate(rmn, rmx, depth) if (spawn)
        build(node->right, rightPart, rmn, rmx, depth + 1);
#pragma omp taskwait
    }
};

// pre-order emission into the 8-byte node layout of include/rtb.h
void emit(const BuildNode *n, int depth, std::vector<rtb_kdnode> &nodes, std::vector<uint32_t> &refs, Tunnel::BuildStats &st)
{
    const size_t self = nodes.size();
    nodes.push_back(rtb_kdnode{0, 0});
    st.kdNodes++;
    if (depth > st.kdMaxDepth) st.kdMaxDepth = depth;
    if (n->axis == 3)
    {
        nodes[self].a = (uint32_t)refs.size();
        nodes[self].b = ((uint32_t)n->list.size() << 2) | 3u;
        refs.insert(refs.end(), n->list.begin(), n->list.end());
        st.kdLeaves++;
        st.kdLeafRefs += (int64_t)n->list.size();
        return;
    }
    emit(n->left, depth + 1, nodes, refs, st);
    const uint32_t right = (uint32_t)nodes.size();
    emit(n->right, depth + 1, nodes, refs, st);
    uint32_t bits;
    memcpy(&bits, &n->split, 4);
    nodes[self].a = bits;
    nodes[self].b = (right << 2) | (uint32_t)n->axis;
}

} // namespace

void Tunnel::initKdTree(const std::vector<TunnelTriangle> &tris)
{ // reference Tunnel.cpp:467-517
    const Bounds b = sceneBounds(tris);
    for (int a = 0; a < 3; a++) { kdMin_[a] = b.mn[a]; kdMax_[a] = b.mx[a]; }
    KdBuilder builder(tris, algorithm == KdTreeSAH, performanceTestBuilders, kdLeafSize, kdMaxDepth, sahCandidates);
    std::vector<uint32_t> list(tris.size());
    for (size_t i = 0; i < list.size(); i++) list[i] = (uint32_t)i;
    BuildNode root;
#pragma omp parallel
#pragma omp single
    builder.build(&root, list, kdMin_, kdMax_, 0);
    emit(&root, 0, kdNodes_, kdLeafTris_, stats);
}

void Tunnel::flatten(FlatScene &out) const
{
    rtb_prim p;
    memset(&p, 0, sizeof(p));
    p.type = RTB_PRIM_TUNNEL;
    p.material = -1;
    p.base_id = out.nTop++;
    out.prims.push_back(p);
    const int wall = wallMaterial ? out.materialIndex(wallMaterial.get()) : 0;
    const int ground = groundMaterial ? out.materialIndex(groundMaterial.get()) : 0;
    for (const auto &seg : surface)
        for (const TunnelTriangle &t : seg)
        {
            const float v[12] = {t.a.x, t.a.y, t.a.z, t.b.x, t.b.y, t.b.z, t.c.x, t.c.y, t.c.z, t.normal.x, t.normal.y, t.normal.z};
            out.tri.insert(out.tri.end(), v, v + 12);
            out.triMaterial.push_back(t.material ? ground : wall);
        }
    rtb_flat_scene &f = out.view;
    f.accel = (int32_t)algorithm;
    memcpy(f.grid_origin, gridOrigin_, sizeof(gridOrigin_));
    memcpy(f.grid_cell, gridCell_, sizeof(gridCell_));
    memcpy(f.grid_dims, gridDims_, sizeof(gridDims_));
    memcpy(f.kd_min, kdMin_, sizeof(kdMin_));
    memcpy(f.kd_max, kdMax_, sizeof(kdMax_));
    out.gridWords = gridWords_;
    out.gridCellStart = gridCellStart_;
    out.gridCellTris = gridCellTris_;
    out.kdNodes = kdNodes_;
    out.kdLeafTris = kdLeafTris_;
    out.cxFrames = cxFrames_; out.cxEdges = cxEdges_; out.cxCellStatus = cxCellStatus_; out.cxCellRange = cxCellRange_; out.cxOrder = cxOrder_;
    f.cx_width = width; f.cx_height = height;
    f.cx_table_size = cxTable_; f.cx_round_bins = performanceTestBuilders ? 1 : 0;
    f.grid_build_resolution = ((algorithm == RegularGrid || algorithm == FlatGrid) && gridOnDevice) ? gridResolution : 0;
    f.grid_build_exact = exactGridBinning ? 1 : 0;
    const bool kdDevice = algorithm == KdTreeSAH && kdOnDevice && !performanceTestBuilders;
    f.kd_build_max_depth = kdDevice ? kdMaxDepth : 0;
    f.kd_build_leaf_size = kdDevice ? kdLeafSize : 0;
    f.kd_build_candidates = kdDevice ? sahCandidates : 0;
}

} // namespace rt
