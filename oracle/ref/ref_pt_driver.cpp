// Headless driver of the reference's PerformanceTest program (src/PerformanceTest), linked with its own
// sources by build_ref.sh into oracle/_ref/libref_pt.so.  TEST INFRASTRUCTURE ONLY (see oracle_abi.h).
//
// It replaces main.cpp: the scene of main.cpp:61-81 (tunnel + exit plane), Tunnel::init() with the
// accelerator classes of that program (GridAcc, KdTreeAcc with the event-sweep SAH builder, ConvexAcc) and
// the trace of main.cpp:29-59, unrolled into a loop, for the camera samples the caller passes (the
// reference draws them with rand()).  Also reports the size and a structure hash of the k-d tree that
// KdTreeAcc built, with the numbering / hashing of ref_driver.cpp.
#include "TunnelGenerator.h"
#include "Plane.h"
#include "Camera.h"
#include "Utils.h"
#include "GridAcc.h"
#include "KdTreeAcc.h"
#include "ConvexAcc.h"
#include "oracle_abi.h"

// ---- stand-in for the Win32-only Utils.cpp of PerformanceTest ------------------------------------------
int Utils::GetTickCount()
{
    using namespace std::chrono;
    static const steady_clock::time_point t0 = steady_clock::now();
    return (int)duration_cast<milliseconds>(steady_clock::now() - t0).count();
}
void Utils::DbgPrint(char *format, ...)
{
    if (!getenv("RTB_REF_VERBOSE")) return;
    va_list args;
    va_start(args, format);
    vfprintf(stderr, format, args);
    va_end(args);
}
void Utils::PrintTickCount(char *desc) { DbgPrint((char *)"%s: %.2lf\n", desc, GetTickCount() / 1000.0); }
int Utils::GetMemorySize() { return 0; }

static inline void hmix(uint64_t &h, uint32_t v) { h = (h ^ v) * 0x100000001b3ull; }
static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static void numberNodes(KdTreeAcc::KdNode *n, int depth, std::unordered_map<const Geometry *, int> &triId,
                        oracle_bounce_job *job, uint64_t &h)
{
    job->stats[ORACLE_STAT_KD_NODES]++;
    if (depth > job->stats[ORACLE_STAT_KD_MAX_DEPTH]) job->stats[ORACLE_STAT_KD_MAX_DEPTH] = depth;
    if (n->axis == KdTreeAcc::NoAxis)
    {
        job->stats[ORACLE_STAT_KD_LEAVES]++;
        job->stats[ORACLE_STAT_KD_LEAF_REFS] += (long long)n->list.size();
        hmix(h, 3);
        hmix(h, (uint32_t)n->list.size());
        for (size_t i = 0; i < n->list.size(); i++) hmix(h, (uint32_t)triId[n->list[i]]);
        return;
    }
    hmix(h, (uint32_t)n->axis);
    hmix(h, fbits(n->splitPlane));
    numberNodes(n->left, depth + 1, triId, job, h);
    numberNodes(n->right, depth + 1, triId, job, h);
}

extern "C" int ref_pt_bounce(oracle_bounce_job *job)
{
    static std::mutex once;
    std::lock_guard<std::mutex> guard(once);
    if (!job || job->n < 0 || job->algorithm < 0 || job->algorithm > 6 || !job->xy) return -1;
    GeometrySet scene;
    TunnelGenerator g;
    g.create(50, 25, 25, job->radius, job->angle, job->arch_seg, job->path_seg, scene, (Tunnel::Algorithm)job->algorithm); // main.cpp:63-69
    Tunnel *tunnel = (Tunnel *)scene.last();
    Vector normal(sin(job->angle), 0, -cos(job->angle)); // main.cpp:75
    float distance = job->radius * sin(job->angle);      // main.cpp:76
    Plane *plane = new Plane(normal, distance);
    scene.add(plane);
    const auto t0 = std::chrono::steady_clock::now();
    tunnel->init();
    job->prepare_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

    std::unordered_map<const Geometry *, int> geomId, triId;
    geomId[tunnel] = 0;
    geomId[plane] = 1;
    int k = 0;
    for (size_t s = 0; s < tunnel->surface.size(); s++)
        for (size_t j = 0; j < tunnel->surface[s].size(); j++, k++)
        {
            geomId[tunnel->surface[s][j]] = 2 + k;
            triId[tunnel->surface[s][j]] = k;
        }
    memset(job->stats, 0, sizeof(job->stats));
    job->stats[ORACLE_STAT_N_TRIS] = k;
    job->struct_hash = 0;
    if (tunnel->accGrid)
    { // same convention as ref_driver.cpp (grid of RayTracingOpt's Tunnel)
        GridAcc *G = tunnel->accGrid;
        uint64_t h = 0xcbf29ce484222325ull;
        hmix(h, 0x47524944u);
        hmix(h, (uint32_t)G->xLength); hmix(h, (uint32_t)G->yLength); hmix(h, (uint32_t)G->zLength);
        hmix(h, fbits(G->origin.x)); hmix(h, fbits(G->origin.y)); hmix(h, fbits(G->origin.z));
        hmix(h, fbits(G->cellSizeX)); hmix(h, fbits(G->cellSizeY)); hmix(h, fbits(G->cellSizeZ));
        job->stats[ORACLE_STAT_GRID_X] = G->xLength;
        job->stats[ORACLE_STAT_GRID_Y] = G->yLength;
        job->stats[ORACLE_STAT_GRID_Z] = G->zLength;
        for (size_t c = 0; c < G->data.size(); c++)
        {
            const std::vector<Triangle *> &l = G->data[c];
            if (l.empty()) continue;
            job->stats[ORACLE_STAT_CELLS_NONEMPTY]++;
            job->stats[ORACLE_STAT_CELL_ENTRIES] += (long long)l.size();
            if ((long long)l.size() > job->stats[ORACLE_STAT_CELL_MAX]) job->stats[ORACLE_STAT_CELL_MAX] = (long long)l.size();
            hmix(h, (uint32_t)c);
            hmix(h, (uint32_t)l.size());
            for (size_t i = 0; i < l.size(); i++) hmix(h, (uint32_t)triId[l[i]]);
        }
        job->struct_hash = h;
    }
    if (tunnel->accConvex)
    { // the convex accelerator's tables (ConvexAcc.h:24-27) + the per-polygon frames intersectWithPolygon derives
        ConvexAcc *X = tunnel->accConvex;
        uint64_t h = 0xcbf29ce484222325ull;
        hmix(h, 0x43565800u);
        hmix(h, (uint32_t)tunnel->path.size()); hmix(h, (uint32_t)tunnel->crossSection.vertices.size());
        hmix(h, fbits(tunnel->width)); hmix(h, fbits(tunnel->height));
        for (size_t i = 0; i < tunnel->path.size(); i++)
        {
            Point p = tunnel->path[i];
            Vector n = tunnel->nvs[i];
            float theta = PI - atan2(n.x, n.z); // ConvexAcc.cpp:14
            const float v[8] = {p.x, p.y, p.z, n.x, n.y, n.z, cos(theta), sin(theta)};
            for (int q = 0; q < 8; q++) hmix(h, fbits(v[q]));
        }
        for (size_t e = 0; e < X->edgeParams.size(); e++) { hmix(h, fbits(X->edgeParams[e].A)); hmix(h, fbits(X->edgeParams[e].B)); hmix(h, fbits(X->edgeParams[e].C)); }
        for (int i = 0; i < 100; i++)
            for (int j = 0; j < 100; j++)
                hmix(h, X->intersectionTable[i][j] == ConvexAcc::Hit ? 0u : (X->intersectionTable[i][j] == ConvexAcc::Partial ? 1u : 2u));
        for (int i = 0; i < 100; i++)
            for (int j = 0; j < 100; j++) { hmix(h, (uint32_t)(uint16_t)X->edgeRangeTable[i][j].start); hmix(h, (uint32_t)(uint16_t)X->edgeRangeTable[i][j].end); }
        if (tunnel->algorithm == Tunnel::Convex)
            for (int y = 0; y < 100; y++)
                for (int a = 0; a < 360; a++)
                    for (size_t k = 0; k < X->intersectionTableYAxis[y][a].size(); k++) hmix(h, (uint32_t)X->intersectionTableYAxis[y][a][k]);
        job->struct_hash = h;
    }
    if (tunnel->accKdTree && tunnel->accKdTree->root)
    {
        KdTreeAcc::KdNode *root = tunnel->accKdTree->root;
        uint64_t h = 0xcbf29ce484222325ull;
        hmix(h, 0x4b445452u);
        hmix(h, fbits(root->min.x)); hmix(h, fbits(root->min.y)); hmix(h, fbits(root->min.z));
        hmix(h, fbits(root->max.x)); hmix(h, fbits(root->max.y)); hmix(h, fbits(root->max.z));
        numberNodes(root, 0, triId, job, h);
        job->struct_hash = h;
    }

    Vector camFront(0, 0, -1);
    Camera camera(Point(0, 25, 5), camFront, Vector(0, 1, 0)); // main.cpp:143-147
    long long total = 0;
    omp_set_num_threads(job->threads > 0 ? job->threads : omp_get_num_procs());
    const auto t1 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
    for (int i = 0; i < job->n; i++)
    {
        Ray ray = camera.generateRay(job->xy[2 * i], job->xy[2 * i + 1]);
        int depth = 0, reached = 0, lastId = -1;
        Point lastPos(0, 0, 0);
        while (true)
        { // main.cpp:29-59, recursion unrolled
            IntersectResult result = scene.intersect(ray);
            total++;
            if (!result.hit) { lastId = -1; break; }
            lastId = geomId[result.geometry];
            lastPos = result.position;
            Vector &n = result.normal;
            Vector nl = (n.dot(ray.direction) < 0) ? n : n * -1;
            if (++depth > job->max_depth) break;
            if (result.geometry->type == GeometryType::PLANE) { reached = 1; break; }
            Vector v = ray.direction - nl * 2 * nl.dot(ray.direction);
            Ray newRay(result.position, v);
            newRay.context = ray.context;
            ray = newRay;
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
