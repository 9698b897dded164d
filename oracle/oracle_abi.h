/* oracle_abi.h -- job description shared by the two CPU checkers of this repo.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Two shared objects export the same entry point over this struct:
 *   oracle/librt_oracle.so : `int rt_oracle_run(oracle_job*)`  -- the CPU restatement (rt_oracle.cpp)
 *   oracle/_ref/libref.so  : `int ref_run(oracle_job*)`        -- the UNMODIFIED reference sources
 *                            compiled headless from /root/reference by oracle/build_ref.sh
 * so a test can run the same job through both and compare every output bit for bit.
 *
 * Conventions (all indices 0-based):
 *   hit id     : -1 miss; g for top-level geometry g of the GeometrySet (insertion order,
 *                reference GeometrySet.cpp:95-110); n_top + k for tunnel triangle k, where k runs
 *                over Tunnel::surface[seg][j] in (seg, j) order (reference Tunnel.h:13) and n_top
 *                is the number of top-level geometries (the tunnel itself included).
 *   sequences  : per primary ray, the accelerator steps in visiting order --
 *                grids: linear cell index (x*yLen + y)*zLen + z (reference Tunnel.h:63-66);
 *                k-d  : pre-order index of every inner node whose split is read and every leaf whose
 *                       list is scanned (reference Tunnel.cpp:1203-1280).
 *                seq_hash = fold over ids of h = (h ^ (uint32)id) * 0x100000001b3, h0 = 0xcbf29ce484222325.
 *   images     : float RGB, reference framebuffer order index = x*height + y, y = 0 top
 *                (reference MainWindow.cpp:276).
 *   per-ray arrays (hit_id, hit_t, seq_*) are row-major y*width + x.
 */
#ifndef RTB_ORACLE_ABI_H
#define RTB_ORACLE_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORACLE_STAT_N_TOP = 0,     /* top-level geometries                                   */
    ORACLE_STAT_N_TRIS = 1,    /* tunnel triangles                                       */
    ORACLE_STAT_GRID_X = 2,    /* grid dims (reference prints "Grid Size: x x y x z")   */
    ORACLE_STAT_GRID_Y = 3,
    ORACLE_STAT_GRID_Z = 4,
    ORACLE_STAT_CELLS_NONEMPTY = 5,
    ORACLE_STAT_CELL_ENTRIES = 6,
    ORACLE_STAT_CELL_MAX = 7,
    ORACLE_STAT_KD_NODES = 8,
    ORACLE_STAT_KD_LEAVES = 9, /* reference prints "Total leaves"                        */
    ORACLE_STAT_KD_LEAF_REFS = 10, /* "Average Leaf Size" = refs / leaves (int division) */
    ORACLE_STAT_KD_MAX_DEPTH = 11,
    ORACLE_STAT_COUNT = 16
};

enum { ORACLE_RNG_ERAND48 = 0, ORACLE_RNG_COUNTER = 1 };
enum { ORACLE_SETTING_PRESET = 0, ORACLE_SETTING_SIMPLE = 1, ORACLE_SETTING_DEFAULT = 2,
       ORACLE_SETTING_HIGHSPEED = 3, ORACLE_SETTING_HIGHQUALITY = 4 };

typedef struct oracle_job {
    /* ---- inputs ---- */
    int32_t preset;      /* 1..5 = reference Scripts.cpp Script1..Script5                 */
    int32_t algorithm;   /* reference Tunnel::Algorithm 0 linear, 1 regular grid, 2 flat grid,
                            3 k-d median, 4 k-d SAH (5/6 convex: out of scope)           */
    int32_t segments;    /* tunnel tessellation (arch = path segments)                    */
    int32_t width, height;
    int32_t samples;     /* samples per pixel when the setting enables Monte Carlo        */
    int32_t setting;     /* ORACLE_SETTING_*                                              */
    int32_t threads;     /* OpenMP threads for the render (<= 0: all)                     */
    int32_t rng;         /* ORACLE_RNG_* (libref.so only knows ERAND48)                   */
    int32_t seq_cap;     /* ids stored per primary ray in seq_buf                         */
    int32_t tri_cap;     /* capacity, in triangles, of tri_out / tri_mat                  */
    int32_t repeat;      /* render repeats; render_ms is the best                         */
    uint64_t seed;       /* counter-RNG frame seed                                        */
    const char *stl_path;/* binary STL for preset 3                                       */
    /* ---- optional output arrays (NULL = skip that product) ---- */
    float *rgb;          /* [w*h*3]                                                       */
    int32_t *hit_id;     /* [w*h]                                                         */
    float *hit_t;        /* [w*h]                                                         */
    int32_t *seq_len;    /* [w*h]                                                         */
    uint64_t *seq_hash;  /* [w*h]                                                         */
    int32_t *seq_buf;    /* [w*h*seq_cap]                                                 */
    float *tri_out;      /* [tri_cap*12] a, b, c, normal of tunnel triangles              */
    int32_t *tri_mat;    /* [tri_cap] 0 wall, 1 ground                                    */
    /* ---- scalar outputs ---- */
    int64_t n_rays;      /* scene intersections issued by the render                      */
    int64_t n_tri_tests; /* Triangle::intersect calls during the render                   */
    int64_t n_steps;     /* cells / k-d nodes visited during the render                   */
    double render_ms;
    double prepare_ms;   /* Tunnel::init()                                                */
    int64_t stats[ORACLE_STAT_COUNT];
    uint64_t struct_hash;/* canonical hash of the accelerator (see rt_oracle.cpp)         */
    uint64_t tri_hash;   /* hash of the tunnel triangle stream                            */
    double *render_ms_all; /* optional [repeat]: every render's interval, in order (input ptr)   */
    uint8_t *rgb8;       /* optional [w*h*3]: the reference's output stage -- saturate, (int)(c*255)
                            (MainWindow.cpp:305-311) -- R,G,B bytes in framebuffer order x*height + y */
    double *moments;     /* optional [w*h*6], Monte Carlo, librt_oracle.so only (libref.so ignores it: the reference keeps
                            no per-sample statistics): per pixel, framebuffer order, the sum over the samples of the
                            radiance (R, G, B) and of its square -- the per-pixel variance the statistical parity
                            criteria of SURVEY.md 8(d) need                                                        */
    int32_t grid_exact;  /* input, grids: 1 = bin triangles with the exact overlap test Triangle::intersectWithGrid
                            (Triangle.cpp:152-199), the branch the reference compiles out at Tunnel.cpp:435-445.
                            librt_oracle.so honours it; of the reference builds only libref_sat.so (that branch enabled
                            by build_ref.sh, patch P8) bins this way, whatever the field says                      */
    int32_t pad2_;
} oracle_job;

int rt_oracle_run(oracle_job *job);
int ref_run(oracle_job *job);

/* The "application independent" traversal benchmark of the reference's PerformanceTest console program
 * (src/PerformanceTest/main.cpp:29-59,127-162): a tunnel of radius `radius` / angle `angle` tessellated
 * arch_seg x path_seg, a plane closing its exit (added AFTER the tunnel), rays from the PT camera
 * (eye (0,25,5), PerformanceTest/Camera.cpp:15-21) at the given (x, y) in [0,1]^2 -- the reference draws
 * them with rand() -- each mirror-reflected until it hits the exit plane, misses, or exceeds max_depth.
 * PerformanceTest's Triangle / Plane / Grid sources equal RayTracingOpt's except for a `type` tag, so
 * libref.so runs this workload on the reference's own intersection code (grid, k-d median, SAH of
 * RayTracingOpt); libref_pt.so is built from PerformanceTest's own sources (its accelerator classes
 * GridAcc / KdTreeAcc with the event-sweep SAH builder) and runs main.cpp's trace on them.               */
typedef struct oracle_bounce_job {
    float radius, angle;
    int32_t arch_seg, path_seg;
    int32_t algorithm;     /* as oracle_job                                                  */
    int32_t n;             /* rays                                                           */
    int32_t max_depth;     /* reference: 200                                                 */
    int32_t threads;       /* <= 0: all; the reference loop is single-threaded (1)           */
    const float *xy;       /* [n][2] camera sample coordinates                               */
    int32_t *reached;      /* [n] 1 = stopped on the exit plane, 0 = miss / depth exceeded   */
    int32_t *depth;        /* [n] rays traced for this sample                                */
    int32_t *last_id;      /* [n] hit id of the last intersection (-1 miss)                  */
    float *last_pos;       /* [n][3] position of the last hit                                */
    int64_t total_rays;
    double trace_ms, prepare_ms;
    /* 0 = the tunnel generator and builders of RayTracingOpt (TunnelGenerator.cpp: rings turned by the segment
     * direction; Tunnel.cpp: 99-candidate SAH); 1 = those of PerformanceTest itself (TunnelGenerator.cpp:243-279:
     * rings turned by the averaged directions of the adjoining segments; KdTreeAcc.cpp: event-sweep SAH with
     * automatic termination, cost = 1 + 1.5 * (...), leaf when cost > 1.5 * n).  The two programs share
     * everything else on this path (cross section, median split, grids, traversal, intersection).          */
    int32_t pt_builders;
    int32_t pad_;
    int64_t stats[16];     /* ORACLE_STAT_*: k-d tree sizes of the accelerator that was built               */
    uint64_t struct_hash;  /* same hash as oracle_job.struct_hash (k-d trees)                                */
} oracle_bounce_job;

int rt_oracle_bounce(oracle_bounce_job *job);
int ref_bounce(oracle_bounce_job *job);    /* libref.so: RayTracingOpt classes (pt_builders must be 0)       */
int ref_pt_bounce(oracle_bounce_job *job); /* libref_pt.so: the PerformanceTest sources themselves           */

#ifdef __cplusplus
}
#endif
#endif
