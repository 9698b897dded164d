// ref_driver.cpp -- headless driver around the UNMODIFIED reference renderer.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_abi.h).  Compiled by oracle/build_ref.sh together
// with the reference's own translation units into oracle/_ref/libref.so.  It plays the role of
// the reference's Win32 MainWindow.cpp: it owns the file-static width/height/samples, hands a
// RenderProc to the reference's Script::Run (Scripts.h:11-12,38-39) and, inside that callback,
// runs the reference's own trace()/radiance()/Render() (lifted verbatim by the build script from
// MainWindow.cpp:69-303 into ref_lifted_render.inc) plus the observation passes the parity tests
// need (hit ids, traversal sequences, structure hashes).  No arithmetic lives here.
#include "GeometrySet.h"
#include "RenderSetting.h"
#include "Scripts.h"
#include "Tunnel.h"
#include "Triangle.h"
#include "Utils.h"
#include "Plane.h"
#include "TunnelGenerator.h"
#include "SolidColorMaterial.h"
#include "oracle_abi.h"

#define erand48 ref_erand48
#define lrand48 ref_lrand48
#define srand48 ref_srand48
#include "erand48.h"

double ref_last_tick_interval_ms();

// ---- the statics MainWindow.cpp:33-36 keeps ------------------------------------------------
static int width;
static int height;
static int samples;

// ---- observation state ----------------------------------------------------------------------
struct alignas(64) Counters { long long rays, tris, steps; };
static Counters g_cnt[512];
static bool g_counting = false;
static thread_local std::vector<int> *tls_rec = nullptr;
static std::unordered_map<const void *, int> g_node_id;
static const char *g_stl_path = "ball.stl";
static oracle_job *g_job = nullptr;
static float *g_rgb_out = nullptr;
static uint8_t *g_rgb8_out = nullptr;

// Win32 stand-ins for the output loop of Render() (MainWindow.cpp:305-312)
typedef int HDC;
typedef unsigned int COLORREF;
static HDC hdcBuffer = 0;
#define RGB(r, g, b) ((COLORREF)(((uint8_t)(r) | ((uint16_t)((uint8_t)(g)) << 8)) | (((uint32_t)(uint8_t)(b)) << 16)))
static void SetPixel(HDC, int x, int y, COLORREF c)
{
    if (!g_rgb8_out) return;
    uint8_t *p = g_rgb8_out + 3 * ((size_t)x * height + y);
    p[0] = (uint8_t)(c & 0xFF); p[1] = (uint8_t)((c >> 8) & 0xFF); p[2] = (uint8_t)((c >> 16) & 0xFF);
}

const char *ref_stl_path() { return g_stl_path; }

#if REF_HOOKS
void ref_hook_cell(int idx)
{
    if (tls_rec) tls_rec->push_back(idx);
    if (g_counting) g_cnt[omp_get_thread_num()].steps++;
}
void ref_hook_node(const void *n)
{
    if (tls_rec) tls_rec->push_back(g_node_id[n]);
    if (g_counting) g_cnt[omp_get_thread_num()].steps++;
}
void ref_hook_tri() { if (g_counting) g_cnt[omp_get_thread_num()].tris++; }
void ref_hook_ray() { if (g_counting) g_cnt[omp_get_thread_num()].rays++; }
#endif

static void ref_render_epilogue(Color *colors)
{
    if (g_rgb_out)
        for (int i = 0; i < width * height; i++)
        {
            g_rgb_out[3 * i + 0] = colors[i].r;
            g_rgb_out[3 * i + 1] = colors[i].g;
            g_rgb_out[3 * i + 2] = colors[i].b;
        }
}

#include "ref_lifted_render.inc" // trace(), radiance(), Render() of the reference

// ---- canonical hashing (same definition as rt_oracle.cpp and the product's python binding) ----
static inline void hmix(uint64_t &h, uint32_t v) { h = (h ^ v) * 0x100000001b3ull; }
static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static const uint64_t H0 = 0xcbf29ce484222325ull;

static void numberNodes(Tunnel::KdNode *n, int depth, std::unordered_map<const Triangle *, int> &triId,
                        oracle_job *job, int &next, uint64_t &h)
{
    g_node_id[n] = next++;
    job->stats[ORACLE_STAT_KD_NODES]++;
    if (depth > job->stats[ORACLE_STAT_KD_MAX_DEPTH]) job->stats[ORACLE_STAT_KD_MAX_DEPTH] = depth;
    if (n->axis == Tunnel::NoAxis)
    {
        job->stats[ORACLE_STAT_KD_LEAVES]++;
        job->stats[ORACLE_STAT_KD_LEAF_REFS] += (long long)n->list.size();
        hmix(h, 3);
        hmix(h, (uint32_t)n->list.size());
        for (size_t i = 0; i < n->list.size(); i++) hmix(h, (uint32_t)triId[(Triangle *)n->list[i]]);
        return;
    }
    hmix(h, (uint32_t)n->axis);
    hmix(h, fbits(n->splitPlane));
    numberNodes(n->left, depth + 1, triId, job, next, h);
    numberNodes(n->right, depth + 1, triId, job, next, h);
}

static void noProgress(int, int) {}
static void noLog(const char *) {}

static int jobRender(GeometrySet &scene, PerspectiveCamera &camera, RenderSetting &scriptSetting,
                     ProgressCallback progress)
{
    oracle_job *job = g_job;
    RenderSetting setting = scriptSetting;
    switch (job->setting)
    {
    case ORACLE_SETTING_SIMPLE: setting = RenderSetting::Simple(); break;
    case ORACLE_SETTING_DEFAULT: setting = RenderSetting::Default(); break;
    case ORACLE_SETTING_HIGHSPEED: setting = RenderSetting::HighSpeed(); break;
    case ORACLE_SETTING_HIGHQUALITY: setting = RenderSetting::HighQuality(); break;
    default: break;
    }

    // ---- number geometries, triangles, nodes; structure statistics -------------------------
    std::unordered_map<const Geometry *, int> geomId;
    std::unordered_map<const Triangle *, int> triId;
    const int nTop = (int)scene.geometries.size();
    Tunnel *tunnel = nullptr;
    for (int i = 0; i < nTop; i++)
    {
        geomId[scene.geometries[i]] = i;
        if (dynamic_cast<Tunnel *>(scene.geometries[i])) tunnel = (Tunnel *)scene.geometries[i];
    }
    memset(job->stats, 0, sizeof(job->stats));
    job->stats[ORACLE_STAT_N_TOP] = nTop;
    job->struct_hash = 0;
    job->tri_hash = 0;
    g_node_id.clear();
    if (tunnel)
    {
        int k = 0;
        uint64_t th = H0;
        const Material *ground = nullptr;
        if (!tunnel->surface.empty() && !tunnel->surface[0].empty())
            ground = &*tunnel->surface[0].back()->material; // last quad of a ring = ground (TunnelGenerator.cpp:331)
        for (size_t s = 0; s < tunnel->surface.size(); s++)
            for (size_t j = 0; j < tunnel->surface[s].size(); j++, k++)
            {
                Triangle *t = tunnel->surface[s][j];
                triId[t] = k;
                geomId[t] = nTop + k;
                const float v[12] = {t->a.x, t->a.y, t->a.z, t->b.x, t->b.y, t->b.z,
                                     t->c.x, t->c.y, t->c.z, t->normal.x, t->normal.y, t->normal.z};
                const int mat = (&*t->material == ground) ? 1 : 0;
                for (int q = 0; q < 12; q++) hmix(th, fbits(v[q]));
                hmix(th, (uint32_t)mat);
                if (job->tri_out && k < job->tri_cap) memcpy(job->tri_out + 12 * (size_t)k, v, sizeof(v));
                if (job->tri_mat && k < job->tri_cap) job->tri_mat[k] = mat;
            }
        job->stats[ORACLE_STAT_N_TRIS] = k;
        job->tri_hash = th;

        if (tunnel->algorithm == Tunnel::RegularGrid || tunnel->algorithm == Tunnel::FlatGrid)
        {
            uint64_t h = H0;
            hmix(h, 0x47524944u);
            hmix(h, (uint32_t)tunnel->grid.xLength);
            hmix(h, (uint32_t)tunnel->grid.yLength);
            hmix(h, (uint32_t)tunnel->grid.zLength);
            hmix(h, fbits(tunnel->grid.origin.x)); hmix(h, fbits(tunnel->grid.origin.y)); hmix(h, fbits(tunnel->grid.origin.z));
            hmix(h, fbits(tunnel->grid.cellSizeX)); hmix(h, fbits(tunnel->grid.cellSizeY)); hmix(h, fbits(tunnel->grid.cellSizeZ));
            job->stats[ORACLE_STAT_GRID_X] = tunnel->grid.xLength;
            job->stats[ORACLE_STAT_GRID_Y] = tunnel->grid.yLength;
            job->stats[ORACLE_STAT_GRID_Z] = tunnel->grid.zLength;
            for (size_t c = 0; c < tunnel->grid.data.size(); c++)
            {
                const std::vector<Triangle *> &l = tunnel->grid.data[c];
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
        else if (tunnel->algorithm == Tunnel::Convex || tunnel->algorithm == Tunnel::ConvexSimple)
        { // convex tables, convention of oracle/rt_oracle.cpp: hashConvex (a Partial cell tests every edge here)
            uint64_t h = H0;
            hmix(h, 0x43565800u);
            hmix(h, (uint32_t)tunnel->path.size()); hmix(h, (uint32_t)tunnel->crossSection.vertices.size());
            hmix(h, fbits(tunnel->width)); hmix(h, fbits(tunnel->height));
            for (size_t i = 0; i < tunnel->path.size(); i++)
            {
                Point p = tunnel->path[i];
                Vector n = tunnel->nvs[i];
                float theta = PI - atan2(n.x, n.z); // Tunnel.cpp:37
                const float v[8] = {p.x, p.y, p.z, n.x, n.y, n.z, cos(theta), sin(theta)};
                for (int q = 0; q < 8; q++) hmix(h, fbits(v[q]));
            }
            const int nEdges = (int)tunnel->edgeParams.size();
            for (int e = 0; e < nEdges; e++) { hmix(h, fbits(tunnel->edgeParams[e].A)); hmix(h, fbits(tunnel->edgeParams[e].B)); hmix(h, fbits(tunnel->edgeParams[e].C)); }
            for (int i = 0; i < 400; i++)
                for (int j = 0; j < 400; j++)
                    hmix(h, tunnel->intersectionTable[i][j] == Tunnel::Hit ? 0u : (tunnel->intersectionTable[i][j] == Tunnel::Partial ? 1u : 2u));
            for (int i = 0; i < 400; i++)
                for (int j = 0; j < 400; j++)
                {
                    const bool partial = tunnel->intersectionTable[i][j] == Tunnel::Partial;
                    hmix(h, (uint32_t)(uint16_t)(short)(partial ? 0 : -1));
                    hmix(h, (uint32_t)(uint16_t)(short)(partial ? nEdges - 1 : -1));
                }
            if (tunnel->algorithm == Tunnel::Convex)
                for (int y = 0; y < 100; y++)
                    for (int a = 0; a < 360; a++)
                        for (size_t k = 0; k < tunnel->intersectionTableYAxis[y][a].size(); k++) hmix(h, (uint32_t)tunnel->intersectionTableYAxis[y][a][k]);
            job->struct_hash = h;
        }
        else if (tunnel->root)
        {
            uint64_t h = H0;
            hmix(h, 0x4b445452u);
            hmix(h, fbits(tunnel->root->min.x)); hmix(h, fbits(tunnel->root->min.y)); hmix(h, fbits(tunnel->root->min.z));
            hmix(h, fbits(tunnel->root->max.x)); hmix(h, fbits(tunnel->root->max.y)); hmix(h, fbits(tunnel->root->max.z));
            int next = 0;
            numberNodes(tunnel->root, 0, triId, job, next, h);
            job->struct_hash = h;
        }
    }

    width = job->width;
    height = job->height;
    samples = job->samples;

    // ---- primary-ray pass: hit ids / distances / traversal sequences ------------------------
    if (job->hit_id || job->hit_t || job->seq_len || job->seq_hash || job->seq_buf)
    {
        const float dx = 1.0f / height, dy = 1.0f / height; // MainWindow.cpp:254-255
        std::vector<int> rec;
        for (int y = 0; y < height; y++)
            for (int x = 0; x < width; x++)
            {
                const float sx = (x + 0.5f) * dx;      // MainWindow.cpp:294-295
                const float sy = 1 - (y + 0.5f) * dy;
                Ray ray(camera.generateRay(sx, sy));
                rec.clear();
                tls_rec = &rec;
                IntersectResult res = scene.intersect(ray);
                tls_rec = nullptr;
                const size_t p = (size_t)y * width + x;
                if (job->hit_id) job->hit_id[p] = res.hit ? geomId[res.geometry] : -1;
                if (job->hit_t) job->hit_t[p] = res.hit ? res.distance : -1.0f;
                if (job->seq_len) job->seq_len[p] = (int)rec.size();
                if (job->seq_hash)
                {
                    uint64_t h = H0;
                    for (size_t i = 0; i < rec.size(); i++) hmix(h, (uint32_t)rec[i]);
                    job->seq_hash[p] = h;
                }
                if (job->seq_buf)
                    for (int i = 0; i < job->seq_cap; i++)
                        job->seq_buf[p * job->seq_cap + i] = i < (int)rec.size() ? rec[i] : -1;
            }
    }

    // ---- the render proper: the reference's own Render() -----------------------------------
    job->n_rays = job->n_tri_tests = job->n_steps = 0;
    job->render_ms = 0;
    if (job->rgb || job->repeat > 0)
    {
        if (job->threads > 0) omp_set_num_threads(job->threads);
        else omp_set_num_threads(omp_get_num_procs());
        const int reps = job->repeat > 0 ? job->repeat : 1;
        double best = 1e300;
        for (int r = 0; r < reps; r++)
        {
            memset(g_cnt, 0, sizeof(g_cnt));
            g_counting = true;
            g_rgb_out = (r == 0) ? job->rgb : nullptr;
            g_rgb8_out = (r == 0) ? job->rgb8 : nullptr;
            Render(scene, camera, setting, noProgress);
            g_counting = false;
            const double ms = ref_last_tick_interval_ms();
            if (job->render_ms_all) job->render_ms_all[r] = ms;
            if (ms < best) best = ms;
        }
        job->render_ms = best;
        for (int t = 0; t < 512; t++)
        {
            job->n_rays += g_cnt[t].rays;
            job->n_tri_tests += g_cnt[t].tris;
            job->n_steps += g_cnt[t].steps;
        }
    }
    return (int)job->render_ms;
}

extern "C" int ref_run(oracle_job *job)
{
    static std::mutex once;
    std::lock_guard<std::mutex> guard(once);
    if (!job || job->preset < 1 || job->preset > 5) return -1;
    if (job->algorithm < 0 || job->algorithm > 6) return -2; // 5, 6: Tunnel::fastIntersect (Convex / ConvexSimple)
    if (job->width <= 0 || job->height <= 0) return -3;
    g_job = job;
    g_stl_path = job->stl_path ? job->stl_path : "ball.stl";
    if (job->preset == 3)
    {
        FILE *fp = fopen(g_stl_path, "rb");
        if (!fp) return -4;
        fclose(fp);
    }
    Utils::RegisterOutputTarget(noLog);
    Script *script = scripts[job->preset - 1];
    script->tunnelSegments = job->segments;
    script->samples = job->samples;
    int prepare = 0, exec = 0;
    const auto t0 = std::chrono::steady_clock::now();
    script->Run(jobRender, job->algorithm, noLog, noProgress, prepare, exec);
    (void)t0;
    job->prepare_ms = prepare;
    g_job = nullptr;
    return 0;
}

// ---- PerformanceTest workload on the reference's own classes (see oracle_abi.h) -----------------
// The scene and the ray loop follow src/PerformanceTest/main.cpp:29-81,143-162; every intersection is
// the reference's GeometrySet / Tunnel / Triangle / Plane code.
extern "C" int ref_bounce(oracle_bounce_job *job)
{
    static std::mutex once;
    std::lock_guard<std::mutex> guard(once);
    if (!job || job->n < 0 || job->algorithm < 0 || job->algorithm > 4 || !job->xy) return -1;
    GeometrySet scene;
    TunnelGenerator g;
    Ptr<Material> dummy(new SolidColorMaterial(Color::Black(), Color::Black(), 1, 0, 0));
    g.create(50, 25, 25, job->radius, job->angle, job->arch_seg, job->path_seg, scene, dummy, dummy,
             (Tunnel::Algorithm)job->algorithm);
    Tunnel *tunnel = (Tunnel *)scene.last();
    Vector normal(sin(job->angle), 0, -cos(job->angle));   // main.cpp:75
    float distance = job->radius * sin(job->angle);        // main.cpp:76
    Plane *plane = new Plane(normal, distance);
    plane->material = dummy;
    scene.add(plane);
    const auto t0 = std::chrono::steady_clock::now();
    tunnel->init();
    job->prepare_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

    std::unordered_map<const Geometry *, int> geomId;
    geomId[tunnel] = 0;
    geomId[plane] = 1;
    int k = 0;
    for (size_t s = 0; s < tunnel->surface.size(); s++)
        for (size_t j = 0; j < tunnel->surface[s].size(); j++, k++) geomId[tunnel->surface[s][j]] = 2 + k;

    // PerformanceTest/Camera.cpp:5-21
    Vector front(0, 0, -1), upIn(0, 1, 0);
    front.norm();
    Vector right = front.cross(upIn).norm();
    Vector up = right.cross(front).norm();
    Point eye(0, 25, 5);
    long long total = 0;
    omp_set_num_threads(job->threads > 0 ? job->threads : omp_get_num_procs());
    const auto t1 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
    for (int i = 0; i < job->n; i++)
    {
        const float x = job->xy[2 * i], y = job->xy[2 * i + 1];
        Vector r = right * ((x - 0.5f) * 1.274f);
        Vector u = up * ((y - 0.5f) * 1.0f);
        Vector dir = (front + r + u).norm();
        Ray ray(eye, dir);
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
            if (result.geometry == plane) { reached = 1; break; }
            Vector v = ray.direction - nl * 2 * nl.dot(ray.direction);
            ray = Ray(result.position, v);
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
