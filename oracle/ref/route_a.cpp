// route_a.cpp -- INTEGRATION.md "Route A", compiled: the few files a maintainer of the REFERENCE adds so that its own
// Script::Run (Scripts.h:38-39) renders on the GPU.
//
// TEST INFRASTRUCTURE (built by oracle/build_ref.sh into oracle/_ref/libroute_a.so, used by tests/test_route_a.py).
// This translation unit is compiled against the reference's OWN headers and linked with the reference's OWN objects --
// its GeometrySet / Plane / Sphere / Triangle / Tunnel (with the grid and the k-d tree the reference's builders
// produce), its materials, its Scripts.cpp -- plus librtb200.so.  What is new is
//   * SceneFlattener: walks those objects and emits the rtb_flat_scene of include/rtb.h;
//   * CudaRender: a RenderProc (Scripts.h:11-12) -- the drop-in for MainWindow.cpp's Render (251-316);
//   * route_a_run: stands in for RenderThread (MainWindow.cpp:350-375), which calls scripts[i]->Run(Render, ...).
// Private members are reached the way the other headless translation units reach them (compat.h); in the reference
// tree proper that is one `friend class SceneFlattener;` per class.
#include "Camera.h"
#include "CheckerMaterial.h"
#include "GeometrySet.h"
#include "GlassMaterial.h"
#include "PhongMaterial.h"
#include "Plane.h"
#include "RadianceCheckerMaterial.h"
#include "RenderSetting.h"
#include "Scripts.h"
#include "SolidColorMaterial.h"
#include "Sphere.h"
#include "Triangle.h"
#include "Tunnel.h"

#include "../../include/rtb.h"

// ---- the statics MainWindow.cpp:33-36 keeps -----------------------------------------------------------------------
static int width = 400, height = 300, samples = 1;
static rtb_ctx *g_ctx = nullptr;
static rtb_multi *g_multi = nullptr;
static int g_devices = 1;
// Color colors[width * height], index = x * height + y (MainWindow.cpp:257, 276) -- in page-locked memory (rtb_host_alloc), so that
// the kernels store the frame straight into it while they render (no device frame, no copy behind the last kernel)
static float *g_colors = nullptr;
static size_t g_colors_floats = 0;
static bool colorsReserve(size_t floats)
{
    if (floats > g_colors_floats)
    {
        if (g_colors) rtb_host_free(g_colors);
        g_colors = nullptr; g_colors_floats = 0;
        void *p = nullptr;
        if (rtb_host_alloc(floats * sizeof(float), &p) != RTB_OK) return false;
        g_colors = (float *)p; g_colors_floats = floats;
    }
    memset(g_colors, 0, floats * sizeof(float));
    return true;
}
static rtb_stats g_stats;
static std::string g_log;
static int g_progress_calls = 0, g_progress_last = 0;

static void AddLog(const char *str) { g_log += str; }                                   // MainWindow.cpp:330-338
static void UpdateProgress(int cur, int total) { g_progress_calls++; g_progress_last = cur; (void)total; } // 340-347

class SceneFlattener
{
public:
    explicit SceneFlattener(GeometrySet &scene)
    {
        memset(&flat_, 0, sizeof(flat_));
        std::vector<Geometry *> &gs = scene.geometries;
        flat_.n_top = (int32_t)gs.size();
        for (size_t i = 0; i < gs.size(); i++)
        {
            Geometry *g = gs[i];
            rtb_prim p;
            memset(&p, 0, sizeof(p));
            p.base_id = (int32_t)i;
            if (Plane *pl = dynamic_cast<Plane *>(g))
            {
                p.type = RTB_PRIM_PLANE; p.material = material(pl->material);
                p.v[0] = pl->normal.x; p.v[1] = pl->normal.y; p.v[2] = pl->normal.z;
                p.v[3] = pl->position.x; p.v[4] = pl->position.y; p.v[5] = pl->position.z; p.v[6] = pl->dist;
                prims_.push_back(p);
            }
            else if (Sphere *sp = dynamic_cast<Sphere *>(g))
            {
                p.type = RTB_PRIM_SPHERE; p.material = material(sp->material);
                p.v[0] = sp->center.x; p.v[1] = sp->center.y; p.v[2] = sp->center.z; p.v[3] = sp->radius;
                prims_.push_back(p);
            }
            else if (Triangle *t = dynamic_cast<Triangle *>(g))
            { // loose triangles (addStlFile, GeometrySet.cpp:33-86): consecutive ones with one material form a run
                const int m = material(t->material);
                if (!prims_.empty() && prims_.back().type == RTB_PRIM_TRIANGLES && prims_.back().material == m &&
                    prims_.back().base_id + prims_.back().count == (int32_t)i)
                    prims_.back().count++;
                else
                {
                    p.type = RTB_PRIM_TRIANGLES; p.material = m; p.first = (int32_t)(loose_.size() / 12); p.count = 1;
                    prims_.push_back(p);
                }
                pushTriangle(loose_, *t);
            }
            else if (Tunnel *tn = dynamic_cast<Tunnel *>(g))
            {
                p.type = RTB_PRIM_TUNNEL;
                prims_.push_back(p);
                tunnel(*tn);
            }
            else ok_ = false; // a geometry class this binding does not know
        }
        flat_.n_prims = (int32_t)prims_.size(); flat_.prims = prims_.data();
        flat_.n_materials = (int32_t)mats_.size(); flat_.materials = mats_.data();
        flat_.n_loose = (int32_t)(loose_.size() / 12); flat_.loose_tri = loose_.empty() ? nullptr : loose_.data();
        flat_.n_tris = (int32_t)(tri_.size() / 12); flat_.tri = tri_.empty() ? nullptr : tri_.data();
        flat_.tri_material = triMat_.empty() ? nullptr : triMat_.data();
    }
    bool ok() const { return ok_; }
    const rtb_flat_scene *view() const { return &flat_; }

    static rtb_camera camera(PerspectiveCamera &c)
    { // Camera.h:10-18; the derived fields were computed by the reference's own constructor (Camera.cpp:4-18)
        rtb_camera r;
        r.eye[0] = c.eye.x; r.eye[1] = c.eye.y; r.eye[2] = c.eye.z;
        r.front[0] = c.front.x; r.front[1] = c.front.y; r.front[2] = c.front.z;
        r.up[0] = c.up.x; r.up[1] = c.up.y; r.up[2] = c.up.z;
        r.right[0] = c.right.x; r.right[1] = c.right.y; r.right[2] = c.right.z;
        r.xcenter = c.xcenter; r.fov_scale = c.fovScale; r.forward = c.forward;
        return r;
    }

private:
    static void pushTriangle(std::vector<float> &out, const Triangle &t)
    {
        const float v[12] = {t.a.x, t.a.y, t.a.z, t.b.x, t.b.y, t.b.z, t.c.x, t.c.y, t.c.z, t.normal.x, t.normal.y, t.normal.z};
        out.insert(out.end(), v, v + 12);
    }

    int material(Ptr<Material> &ptr)
    {
        Material *m = ptr.data;
        auto it = matIndex_.find(m);
        if (it != matIndex_.end()) return it->second;
        rtb_material r;
        memset(&r, 0, sizeof(r));
        r.diffusiveness = m->diffusiveness; r.reflectiveness = m->reflectiveness; r.refractiveness = m->refractiveness;
        r.refractive_index = m->refractiveness > 0 ? m->refractive_index : 0.0f; // uninitialised for the others (Material.cpp:3-8)
        if (PhongMaterial *ph = dynamic_cast<PhongMaterial *>(m))
        {
            r.kind = RTB_MAT_PHONG;
            r.a[0] = ph->diffuse.r; r.a[1] = ph->diffuse.g; r.a[2] = ph->diffuse.b;
            r.b[0] = ph->specular.r; r.b[1] = ph->specular.g; r.b[2] = ph->specular.b;
            r.p = ph->shininess;
        }
        else if (CheckerMaterial *ch = dynamic_cast<CheckerMaterial *>(m))
        {
            r.kind = RTB_MAT_CHECKER; r.scale = ch->scale;
            r.dir = ch->dir == CheckerMaterial::xoz ? RTB_DIR_XOZ : (ch->dir == CheckerMaterial::xoy ? RTB_DIR_XOY : RTB_DIR_YOZ);
        }
        else if (RadianceCheckerMaterial *rc = dynamic_cast<RadianceCheckerMaterial *>(m))
        {
            r.kind = RTB_MAT_RADIANCE_CHECKER; r.scale = rc->scale; r.p = rc->radiance;
            r.dir = rc->dir == RadianceCheckerMaterial::xoz ? RTB_DIR_XOZ : (rc->dir == RadianceCheckerMaterial::xoy ? RTB_DIR_XOY : RTB_DIR_YOZ);
        }
        else if (SolidColorMaterial *so = dynamic_cast<SolidColorMaterial *>(m)) // GlassMaterial included
        {
            r.kind = RTB_MAT_SOLID;
            r.a[0] = so->localColor.r; r.a[1] = so->localColor.g; r.a[2] = so->localColor.b;
            r.b[0] = so->emissionColor.r; r.b[1] = so->emissionColor.g; r.b[2] = so->emissionColor.b;
        }
        else ok_ = false;
        const int id = (int)mats_.size();
        mats_.push_back(r);
        matIndex_[m] = id;
        return id;
    }

    void tunnel(Tunnel &t)
    {
        std::unordered_map<const Triangle *, uint32_t> id;
        for (size_t s = 0; s < t.surface.size(); s++)
            for (size_t j = 0; j < t.surface[s].size(); j++)
            { // hit id of a tunnel triangle = n_top + its position in surface[seg][j] order
                Triangle *tr = t.surface[s][j];
                id[tr] = (uint32_t)(tri_.size() / 12);
                pushTriangle(tri_, *tr);
                triMat_.push_back(material(tr->material));
            }
        switch (t.algorithm)
        {
        case Tunnel::Linear: flat_.accel = RTB_ACCEL_LINEAR; break;
        case Tunnel::RegularGrid:
        case Tunnel::FlatGrid:
            flat_.accel = t.algorithm == Tunnel::RegularGrid ? RTB_ACCEL_REGULAR_GRID : RTB_ACCEL_FLAT_GRID;
            grid(t, id);
            break;
        case Tunnel::KdTreeStandard:
        case Tunnel::KdTreeSAH:
            flat_.accel = t.algorithm == Tunnel::KdTreeStandard ? RTB_ACCEL_KD_MEDIAN : RTB_ACCEL_KD_SAH;
            flat_.kd_min[0] = t.root->min.x; flat_.kd_min[1] = t.root->min.y; flat_.kd_min[2] = t.root->min.z;
            flat_.kd_max[0] = t.root->max.x; flat_.kd_max[1] = t.root->max.y; flat_.kd_max[2] = t.root->max.z;
            kd(t.root, id);
            flat_.n_kd_nodes = (int32_t)nodes_.size(); flat_.kd_nodes = nodes_.data();
            flat_.n_kd_refs = (int64_t)leafTris_.size(); flat_.kd_leaf_tris = leafTris_.data();
            break;
        default: ok_ = false; // the convex walk keeps its tables in fixed arrays of the class; not part of this binding
        }
    }

    // Tunnel::grid (Tunnel.h:51-67): vector<vector<Triangle *>> indexed (x * yLength + y) * zLength + z -> occupancy words,
    // list starts of the occupied cells, one array of triangle indices
    void grid(Tunnel &t, std::unordered_map<const Triangle *, uint32_t> &id)
    {
        flat_.grid_origin[0] = t.grid.origin.x; flat_.grid_origin[1] = t.grid.origin.y; flat_.grid_origin[2] = t.grid.origin.z;
        flat_.grid_cell[0] = t.grid.cellSizeX; flat_.grid_cell[1] = t.grid.cellSizeY; flat_.grid_cell[2] = t.grid.cellSizeZ;
        flat_.grid_dims[0] = t.grid.xLength; flat_.grid_dims[1] = t.grid.yLength; flat_.grid_dims[2] = t.grid.zLength;
        const size_t cells = t.grid.data.size();
        words_.assign((cells + 31) / 32, rtb_cellword{0, 0});
        cellStart_.clear(); cellTris_.clear();
        uint32_t used = 0;
        for (size_t c = 0; c < cells; c++)
        {
            if ((c & 31) == 0) words_[c >> 5].rank = used;
            const std::vector<Triangle *> &list = t.grid.data[c];
            if (list.empty()) continue;
            words_[c >> 5].bits |= 1u << (c & 31);
            cellStart_.push_back((uint32_t)cellTris_.size());
            for (Triangle *tr : list) cellTris_.push_back(id[tr]);
            used++;
        }
        cellStart_.push_back((uint32_t)cellTris_.size());
        flat_.n_cellwords = (int64_t)words_.size(); flat_.grid_words = words_.data();
        flat_.n_cells_used = used; flat_.grid_cell_start = cellStart_.data();
        flat_.n_cell_refs = (int64_t)cellTris_.size(); flat_.grid_cell_tris = cellTris_.data();
    }

    // Tunnel::root (Tunnel.h:75-93): pointer tree -> pre-order array of 8-byte nodes (left child = index + 1)
    void kd(Tunnel::KdNode *n, std::unordered_map<const Triangle *, uint32_t> &id)
    {
        const size_t me = nodes_.size();
        nodes_.push_back(rtb_kdnode{0, 0});
        if (n->axis == Tunnel::NoAxis)
        {
            nodes_[me].a = (uint32_t)leafTris_.size();
            nodes_[me].b = ((uint32_t)n->list.size() << 2) | 3u;
            for (Geometry *g : n->list) leafTris_.push_back(id[(Triangle *)g]);
            return;
        }
        float split = n->splitPlane;
        memcpy(&nodes_[me].a, &split, 4);
        kd(n->left, id);
        nodes_[me].b = ((uint32_t)nodes_.size() << 2) | (uint32_t)n->axis;
        kd(n->right, id);
    }

    rtb_flat_scene flat_;
    bool ok_ = true;
    std::vector<rtb_prim> prims_;
    std::vector<rtb_material> mats_;
    std::unordered_map<const Material *, int> matIndex_;
    std::vector<float> loose_, tri_;
    std::vector<int32_t> triMat_;
    std::vector<rtb_cellword> words_;
    std::vector<uint32_t> cellStart_, cellTris_, leafTris_;
    std::vector<rtb_kdnode> nodes_;
};

static ProgressCallback g_progress = nullptr;
static void progressTrampoline(int64_t done, int64_t total, void *)
{ // tiles -> the reference's unit, rows of `height` (MainWindow.cpp:271)
    if (g_progress) g_progress(total > 0 ? (int)(done * height / total) : height, height);
}

// The drop-in for `int Render(GeometrySet&, PerspectiveCamera&, RenderSetting&, ProgressCallback)` (MainWindow.cpp:251)
int CudaRender(GeometrySet &scene, PerspectiveCamera &camera, RenderSetting &setting, ProgressCallback progress)
{
    SceneFlattener flat(scene);
    if (!flat.ok()) { AddLog("CudaRender: the scene holds a geometry or material class the flattener does not know\r\n"); return -1; }
    const rtb_camera cam = SceneFlattener::camera(camera);
    const rtb_render_setting rs = {setting.enableMonteCarlo ? 1 : 0, setting.maxDepth, setting.terminationDepth, setting.singleTracingDepth};
    rtb_frame fr;
    memset(&fr, 0, sizeof(fr));
    fr.width = width; fr.height = height; fr.samples = samples;
    fr.world = 1; fr.row_block = 8; fr.layout = RTB_LAYOUT_REFERENCE; // Color colors[x * height + y]
    if (!colorsReserve((size_t)width * height * 3)) { AddLog("CudaRender: rtb_host_alloc failed\r\n"); return -1; }
    g_progress = progress;
    int rc;
    int t1 = Utils::GetTickCount();
    if (g_devices > 1)
    {
        rtb_multi_scene *dev = nullptr;
        if (rtb_multi_scene_upload(g_multi, flat.view(), &dev) != RTB_OK) { AddLog(rtb_multi_last_error(g_multi)); AddLog("\r\n"); return -1; }
        rtb_multi_set_progress(g_multi, progress ? progressTrampoline : nullptr, nullptr);
        rc = rtb_multi_render(g_multi, dev, &cam, &rs, &fr, g_colors, &g_stats);
        rtb_multi_scene_free(g_multi, dev);
        if (rc != RTB_OK) { AddLog(rtb_multi_last_error(g_multi)); AddLog("\r\n"); return -1; }
    }
    else
    {
        rtb_scene *dev = nullptr;
        if (rtb_scene_upload(g_ctx, flat.view(), &dev) != RTB_OK) { AddLog(rtb_last_error(g_ctx)); AddLog("\r\n"); return -1; }
        rtb_set_progress(g_ctx, progress ? progressTrampoline : nullptr, nullptr);
        rc = rtb_render(g_ctx, dev, &cam, &rs, &fr, g_colors, &g_stats);
        rtb_scene_free(g_ctx, dev);
        if (rc != RTB_OK) { AddLog(rtb_last_error(g_ctx)); AddLog("\r\n"); return -1; }
    }
    int t2 = Utils::GetTickCount();
    // MainWindow.cpp:305-312 (saturate, (int)(c * 255), SetPixel) would follow here, unchanged
    return t2 - t1 > 0 ? t2 - t1 : 1;
}

// Path of preset 3's STL file (build_ref.sh patch P6 routes Scripts.cpp:151 through this in the headless build)
static const char *g_stl_path = "ball.stl";
const char *ref_stl_path() { return g_stl_path; }

extern "C" {

// = RenderThread (MainWindow.cpp:350-375): scripts[index]->Run(Render, algorithm, AddLog, UpdateProgress, prepare, exec)
// with CudaRender in place of Render.  rgb_out: width * height * 3 floats in the reference's framebuffer order.
int route_a_run(int preset, int algorithm, int segments, int w, int h, int spp, int n_devices, const char *stl_path,
                float *rgb_out, int *prepare_ms, int *exec_ms, long long counts[3], int progress[2], char *log_out, int log_cap)
{
    if (preset < 1 || preset > 5) return -10;
    width = w; height = h; samples = spp;
    if (stl_path) g_stl_path = stl_path;
    g_log.clear(); g_progress_calls = 0; g_progress_last = 0;
    g_devices = n_devices > 1 ? n_devices : 1;
    if (g_devices > 1)
    {
        if (!g_multi && rtb_multi_init(g_devices, nullptr, &g_multi) != RTB_OK) return -11;
        if (rtb_multi_count(g_multi) != g_devices) { rtb_multi_shutdown(g_multi); g_multi = nullptr; if (rtb_multi_init(g_devices, nullptr, &g_multi) != RTB_OK) return -11; }
    }
    else if (!g_ctx && rtb_init(0, &g_ctx) != RTB_OK) return -11;
    Script *script = scripts[preset - 1];
    script->tunnelSegments = segments;
    script->samples = spp;
    int prepare = 0, exec = 0;
    script->Run(CudaRender, algorithm, AddLog, UpdateProgress, prepare, exec);
    if (prepare_ms) *prepare_ms = prepare;
    if (exec_ms) *exec_ms = exec;
    if (log_out && log_cap > 0) { strncpy(log_out, g_log.c_str(), (size_t)log_cap - 1); log_out[log_cap - 1] = 0; }
    if (progress) { progress[0] = g_progress_calls; progress[1] = g_progress_last; }
    if (exec < 0) return exec;
    if (counts) { counts[0] = g_stats.n_rays; counts[1] = g_stats.n_tri_tests; counts[2] = g_stats.n_steps; }
    if (rgb_out && g_colors) memcpy(rgb_out, g_colors, (size_t)w * h * 3 * sizeof(float));
    return 0;
}

void route_a_shutdown()
{
    if (g_ctx) rtb_shutdown(g_ctx);
    g_ctx = nullptr;
    if (g_multi) rtb_multi_shutdown(g_multi);
    g_multi = nullptr;
}

} // extern "C"
