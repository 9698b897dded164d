/* rtb.h -- C ABI of the B200 ray-tracing hot path (librtb200.so).
 *
 * This is the drop-in boundary for ONE path of windy32/win32-ray-tracing-demo: per-pixel camera
 * ray generation -> ray/scene intersection through the regular-grid / flat-grid / k-d (median,
 * SAH) accelerators and the plane / sphere / triangle primitives -> Whitted `trace` and
 * Monte-Carlo `radiance` shading.  Every entry point cites the reference interface it replaces
 * (paths relative to the reference's src/RayTracingOpt/).
 *
 * Plain pointers and sizes only; no C++ / torch types.  The caller owns every host buffer it
 * passes; the library owns device memory behind the opaque handles.  All functions return
 * RTB_OK (0) or a negative rtb_status and are re-entrant per context (one context = one CUDA
 * device + one stream).  There is NO CPU fallback: without a CUDA device rtb_init fails.
 *
 * Host-side builders (grid, k-d tree: reference Tunnel.cpp:346-784) stay on the host and hand
 * their output over as the flattened SoA buffers of rtb_flat_scene.
 */
#ifndef RTB_H
#define RTB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB_ABI_VERSION 5

typedef enum rtb_status {
    RTB_OK = 0,
    RTB_ERR_INVALID = -1,     /* bad argument / inconsistent flat scene          */
    RTB_ERR_NO_DEVICE = -2,   /* no usable CUDA device (there is no CPU path)    */
    RTB_ERR_CUDA = -3,        /* CUDA runtime error; see rtb_last_error          */
    RTB_ERR_UNSUPPORTED = -4, /* valid request outside the implemented path      */
    RTB_ERR_OOM = -5
} rtb_status;

/* ---- materials: reference Material.h:7-38 and subclasses, flattened to a tagged POD ------- */
enum { RTB_MAT_SOLID = 0,            /* SolidColorMaterial.cpp:11-19 (GlassMaterial.cpp:3-7 is SOLID white, (0,0,1), n=1.46) */
       RTB_MAT_CHECKER = 1,          /* CheckerMaterial.cpp:11-23                 */
       RTB_MAT_RADIANCE_CHECKER = 2, /* RadianceCheckerMaterial.cpp:12-29         */
       RTB_MAT_PHONG = 3 };          /* PhongMaterial.cpp:13-29                   */
enum { RTB_DIR_XOZ = 0, RTB_DIR_XOY = 1, RTB_DIR_YOZ = 2 }; /* CheckerMaterial.h:9 */

typedef struct rtb_material {
    int32_t kind;
    float diffusiveness, reflectiveness, refractiveness, refractive_index; /* Material.h:18-24 */
    float a[3];   /* SOLID: local colour;  PHONG: diffuse  */
    float b[3];   /* SOLID: emission;      PHONG: specular */
    float scale;  /* checker scale                         */
    float p;      /* RADIANCE_CHECKER: radiance; PHONG: shininess */
    int32_t dir;  /* RTB_DIR_*                             */
    int32_t pad_[2];
} rtb_material;   /* 64 bytes */

/* ---- top-level geometries: reference GeometrySet.cpp:95-110 iterates them in insertion order */
enum { RTB_PRIM_PLANE = 0,     /* Plane.cpp:3-34: v = normal xyz, position xyz, dist              */
       RTB_PRIM_SPHERE = 1,    /* Sphere.cpp:4-37: v = center xyz, radius                         */
       RTB_PRIM_TRIANGLES = 2, /* run of loose Triangle geometries (GeometrySet.cpp:33-86): first/count index loose_tri */
       RTB_PRIM_TUNNEL = 3 };  /* the Tunnel geometry (Tunnel.cpp:1299-1309), traversed with `accel` */

typedef struct rtb_prim {
    int32_t type;
    int32_t material; /* index into materials (ignored for TUNNEL: per-triangle) */
    int32_t base_id;  /* top-level geometry index of this record (of its first triangle for a run) */
    int32_t first, count; /* TRIANGLES: range in loose_tri */
    float v[7];
} rtb_prim;           /* 48 bytes */

/* ---- accelerators: reference Tunnel.h:16-22 ---------------------------------------------- */
enum { RTB_ACCEL_LINEAR = 0,       /* Tunnel.cpp:786-804  */
       RTB_ACCEL_REGULAR_GRID = 1, /* Tunnel.cpp:346-465 (uniform cell = maxExtent/399), 819-970 */
       RTB_ACCEL_FLAT_GRID = 2,    /* 400^3 anisotropic cells                                  */
       RTB_ACCEL_KD_MEDIAN = 3,    /* Tunnel.cpp:546-669, 1163-1297                            */
       RTB_ACCEL_KD_SAH = 4,       /* Tunnel.cpp:671-784                                       */
       RTB_ACCEL_CONVEX = 5,        /* PerformanceTest/ConvexAcc.cpp (Tunnel.cpp:972-1161): polygon-to-polygon walk,  */
       RTB_ACCEL_CONVEX_SIMPLE = 6 }; /* first accepted triangle of the wall segment, table-ordered / list-ordered     */

/* 8-byte k-d node, nodes stored in pre-order (left child = index + 1).
 *   inner: a = float bits of the split position, b = (right_child_index << 2) | axis (0,1,2)
 *   leaf : a = first index into kd_leaf_tris,    b = (count << 2) | 3                          */
typedef struct rtb_kdnode { uint32_t a, b; } rtb_kdnode;

/* Sparse grid cell directory: one word pair per 32 consecutive linear cell indices
 * ((x*ny + y)*nz + z, reference Tunnel.h:63-66): `bits` marks non-empty cells, `rank` is the
 * number of non-empty cells before this word.  Non-empty cell r owns
 * grid_cell_tris[grid_cell_start[r] .. grid_cell_start[r+1]).                                   */
typedef struct rtb_cellword { uint32_t bits, rank; } rtb_cellword;

typedef struct rtb_flat_scene {
    int32_t n_prims;      const rtb_prim *prims;
    int32_t n_materials;  const rtb_material *materials;
    int32_t n_top;        /* number of top-level geometries; tunnel triangle k has hit id n_top + k */
    int32_t n_loose;      const float *loose_tri;  /* [n_loose][12]: a, b, c, normal (Triangle.h:10-11) */
    int32_t n_tris;       const float *tri;        /* [n_tris][12] tunnel triangles in surface[seg][j] order */
                          const int32_t *tri_material; /* [n_tris] */
    int32_t accel;        /* RTB_ACCEL_* used by the TUNNEL prim */
    /* grid (RTB_ACCEL_REGULAR_GRID / FLAT_GRID): reference Tunnel.h:51-67 */
    float grid_origin[3], grid_cell[3];
    int32_t grid_dims[3];
    int32_t grid_build_resolution; /* > 1 with grid_words == NULL: the library builds the grid ON THE DEVICE from `tri`
                                      with the reference's rules (Tunnel.cpp:346-465; 400 in the reference): REGULAR
                                      cuts the longest extent into resolution - 1 cubic cells, FLAT every extent;
                                      grid_origin / grid_cell / grid_dims and the three arrays below are ignored        */
    int64_t n_cellwords;  const rtb_cellword *grid_words;      /* ceil(nx*ny*nz / 32)        */
    int64_t n_cells_used; const uint32_t *grid_cell_start;     /* [n_cells_used + 1]         */
    int64_t n_cell_refs;  const uint32_t *grid_cell_tris;      /* [n_cell_refs]              */
    /* k-d tree (RTB_ACCEL_KD_*): reference Tunnel.h:75-93 */
    float kd_min[3], kd_max[3];                                 /* root box                   */
    int32_t n_kd_nodes;   const rtb_kdnode *kd_nodes;
    int64_t n_kd_refs;    const uint32_t *kd_leaf_tris;
    /* convex accelerator (RTB_ACCEL_CONVEX*): reference PerformanceTest/ConvexAcc.h:8-28.  The tunnel is n_cx_path - 1
     * segments of 2 * n_cx_edges triangles each (tri order = surface[seg][j]).                                        */
    int32_t n_cx_path;    const float *cx_frames;  /* [n_cx_path][8]: path vertex xyz, ring normal xyz, cos / sin of the
                                                       rotation that maps the ring's plane to z = 0 (ConvexAcc.cpp:12-18;
                                                       host cosf / sinf / atan2f, as the reference evaluates them)        */
    int32_t n_cx_edges;   const float *cx_edges;   /* [n_cx_edges][3]: A, B, C of A*x + B*y + C > 0 (ConvexAcc.cpp:183-214) */
    float cx_width, cx_height;                      /* bounding rectangle of the cross section (Tunnel.h:19-20)              */
    int32_t cx_table_size;                          /* T: cells per side of the lookup over the cross section's bounding
                                                       rectangle -- 100 in PerformanceTest (ConvexAcc.h:12), 400 in
                                                       RayTracingOpt (Tunnel.cpp:255-284)                                     */
    int32_t cx_round_bins;                          /* direction bin of the order table: 1 = (int)(deg + 0.5) % 360
                                                       (ConvexAcc.cpp:372), 0 = (int)deg (Tunnel.cpp:1042)                    */
    const uint8_t *cx_cell_status;                  /* [T*T] 0 Hit, 1 Partial, 2 Miss (ConvexAcc.cpp:216-247)                */
    const int16_t *cx_cell_range;                   /* [T*T][2] first / last edge to test in a Partial cell (RayTracingOpt
                                                       tests them all: 0 .. n_cx_edges - 1)                                   */
    const uint16_t *cx_order;                       /* RTB_ACCEL_CONVEX: [100][360][2*n_cx_edges] triangle order per
                                                       (height, angle) bin (ConvexAcc.cpp:249-270)                           */
    int32_t grid_build_exact;                       /* device grid build (grid_build_resolution): 1 = a triangle enters a cell only
                                                       if it passes the reference's exact overlap test Triangle::intersectWithGrid
                                                       (Triangle.cpp:152-199), the alternative the reference compiles out at
                                                       Tunnel.cpp:435-445 ("8 times faster" to build without, "20% slower" to
                                                       traverse); 0 = bounding-box binning as the reference ships               */
    /* RTB_ACCEL_KD_SAH with kd_nodes == NULL and kd_build_max_depth > 0: the library builds the tree ON THE DEVICE from `tri`
     * with the reference's builder (Tunnel.cpp:546-638, 671-784): leaf when the list holds <= kd_build_leaf_size triangles or
     * depth > kd_build_max_depth (reference: 8 and 18), kd_build_candidates - 1 uniform planes per axis (reference: 100 - 1).
     * kd_min / kd_max / n_kd_* and the two arrays are ignored.                                                                */
    int32_t kd_build_leaf_size, kd_build_max_depth, kd_build_candidates;
    /* 1 = the arrays above are page-locked host memory (rtb_host_alloc / rtb_host_register) AND stay unchanged until a render
     * of the uploaded scene has returned or the scene has been freed: the upload then copies host -> device straight out of
     * them (truly asynchronous copies, no staging pass over the 4.5 MB of a 45,900-triangle scene, 0.6 -> 0.3 ms per upload).
     * 0 = the arrays may be pageable and may be reused as soon as the upload call returns (they are staged).  An array that
     * is not page-locked is staged whatever the flag says.                                                                  */
    int32_t arrays_page_locked;
} rtb_flat_scene;

/* ---- camera: reference Camera.h:7-23; the derived fields are computed on the host exactly as
 * Camera.cpp:4-18 does (fovScale needs the host tanf)                                        */
typedef struct rtb_camera {
    float eye[3], front[3], up[3], right[3];
    float xcenter, fov_scale, forward;
} rtb_camera;

/* ---- render policy: reference RenderSetting.h:6-79 ---------------------------------------- */
typedef struct rtb_render_setting {
    int32_t enable_monte_carlo;
    int32_t max_depth, termination_depth, single_tracing_depth;
} rtb_render_setting;

/* ---- frame description: the file-statics width/height/samples of reference
 * MainWindow.cpp:33-36 plus the tile-row shard this device renders.
 * Rows are dealt to ranks in blocks of `row_block` rows, block b going to rank b % world
 * (the reference's `omp parallel for schedule(dynamic,1)` over y, MainWindow.cpp:267-269, made
 * static across devices).  The output holds this rank's rows only, in increasing y, row-major
 * [local_row][x][3] float RGB -- unless RTB_LAYOUT_REFERENCE is set (world must be 1), which
 * gives the reference framebuffer order index = x*height + y (MainWindow.cpp:276).          */
enum { RTB_LAYOUT_ROWMAJOR = 0, RTB_LAYOUT_REFERENCE = 1,
       /* OR-able: the output buffer is the WHOLE frame, row-major [height][width][3], and this rank stores only the
        * pixels of its shard into it, at their place in the frame (slot = y * width + x).  Every rank of a frame is
        * given the same buffer -- rank 0's device frame opened on the peers through CUDA IPC (rtb_ipc_*: the stores
        * cross NVLink while the shard renders; no gather, no unshard pass), or one page-locked host frame that all
        * devices of a process store into (rtb_multi_render).  The reference has exactly one framebuffer
        * (MainWindow.cpp:257); this is that buffer, filled by several devices.                                     */
       RTB_LAYOUT_GLOBAL = 2,
       /* OR-able: the output buffer holds 3 BYTES per pixel (R, G, B) produced by the reference's output
        * stage -- saturate (upper clamp only) then (int)(c * 255), MainWindow.cpp:305-311 -- instead of 3
        * floats.  A quarter of the bytes to read back; the float framebuffer is what parity is judged on. */
       RTB_OUTPUT_RGB8 = 4,
       /* OR-able, Monte Carlo only: the output buffer holds SIX floats per pixel -- the sum over this call's samples of
        * the per-sample radiance (R, G, B; the reference accumulates radiance * (1 / samples), MainWindow.cpp:288: here
        * the plain sum) and the sum of its squares -- the accumulation buffer (sum, sum of squares, spp) of a progressive
        * or sample-sharded render: buffers of disjoint sample ranges add up, mean = sum / spp,
        * variance = (sumsq - sum^2 / spp) / (spp - 1).  Not with RTB_OUTPUT_RGB8 / RTB_LAYOUT_REFERENCE.            */
       RTB_OUTPUT_MOMENTS = 8 };
typedef struct rtb_frame {
    int32_t width, height;
    int32_t samples;      /* spp when enable_monte_carlo                                      */
    uint64_t seed;        /* counter-based RNG frame seed (key = pixel, sample)               */
    int32_t rank, world;  /* shard selector; 0,1 = whole frame                                */
    int32_t row_block;    /* rows per dealt block (multiple of 8; 0 = default 8)              */
    int32_t layout;       /* RTB_LAYOUT_*                                                     */
    int32_t counters;     /* 1: also count triangle tests / traversal steps (slower); 2: as 1 and write a
                             per-pixel cost map (thread cycles, rays, steps + tests) instead of the colour
                             (profiling aid, Whitted chain kernel only)                          */
    int32_t col_block;    /* 0: the shard is whole rows (blocks of row_block rows dealt round-robin).  > 0 (with world
                             > 1): COLUMN-BLOCK shard -- every rank renders every row and, of block row
                             by = y / row_block, the column blocks bx = x / col_block with (bx + by) % world == rank;
                             its output is a local image [height][width / world] in which local column block c holds
                             global block c * world + (rank - by) mod world.  Multiple of 8, width a multiple of
                             world * col_block, row-major layout.  The tunnel frames concentrate their cost around
                             the vanishing point: narrow column blocks give every rank the same mix of tiles.
                             (Occupies what was tail padding of this struct: zero-initialised callers get rows.)  */
    int32_t sample_first; /* Monte Carlo SAMPLE shard: this call renders samples [sample_first, sample_first +   */
    int32_t sample_count; /* sample_count) of the frame's `samples` (0 = all of them).  The RNG is keyed by (pixel,
                             sample index) and every sample is weighted 1 / samples, so the images of disjoint ranges
                             ADD UP to the full frame (float rounding of the per-pixel sum aside: the order of the
                             additions differs from the single pass): low-resolution / high-spp frames shard over
                             samples instead of rows, one NCCL reduce (sum) assembles them.              */
} rtb_frame;

typedef struct rtb_stats {
    int64_t n_rays;        /* scene intersections issued (primary + secondary)                */
    int64_t n_tri_tests;   /* frame.counters only                                             */
    int64_t n_steps;       /* cells / k-d nodes visited; frame.counters only                  */
    int64_t n_local_rows;  /* rows rendered by this rank                                      */
    float kernel_ms;       /* CUDA-event time of the render kernel(s) on the context stream   */
    float total_ms;        /* CUDA-event time of the whole call incl. copies                  */
    int32_t n_launches;    /* kernels launched by the call                                    */
    int32_t pad_;
} rtb_stats;

typedef struct rtb_ctx rtb_ctx;
typedef struct rtb_scene rtb_scene;

/* Context = one CUDA device.  `stream` may be 0 (library-owned stream) or a caller's
 * cudaStream_t cast to void* (e.g. torch's current stream) for rtb_render_device.            */
int rtb_init(int device, rtb_ctx **ctx);
int rtb_shutdown(rtb_ctx *ctx);
const char *rtb_last_error(const rtb_ctx *ctx); /* ctx may be NULL: last global error       */
int rtb_abi_version(void);

/* Page-locked host memory for the host-buffer calls (optional: any host pointer is accepted, but
 * device->host copies into pageable memory are staged by the driver and several times slower).   */
int rtb_host_alloc(size_t bytes, void **out);
int rtb_host_free(void *p);
/* Page-lock memory the caller already owns (the arrays of a flat scene, see rtb_flat_scene.arrays_page_locked; a frame
 * buffer): every device of the process may copy from / store into it.  Unregister before the memory is freed.          */
int rtb_host_register(void *p, size_t bytes);
int rtb_host_unregister(void *p);

/* Number of rows / first-row list of a shard (pure host arithmetic; no device needed).       */
int64_t rtb_shard_rows(const rtb_frame *frame);
/* width of the shard's local image: frame->width, or frame->width / world for a column-block shard */
int64_t rtb_shard_width(const rtb_frame *frame);

/* Host-only check (no device needed) that nodes[0 .. n) is a well-formed pre-order k-d array as rtb_scene_upload
 * requires it: every inner node's right child lies behind its left subtree and before the end of the enclosing
 * subtree, every node belongs to exactly one subtree, no node is deeper than 64.  RTB_OK and *max_depth (root = 0)
 * or RTB_ERR_INVALID.  rtb_scene_upload runs the same check (and refuses trees whose traversal stack,
 * 2 * max_depth + 2 entries, would exceed the reference's 50, Tunnel.cpp:1176).                                   */
int rtb_kd_validate(const rtb_kdnode *nodes, int32_t n, int32_t *max_depth);

/* Upload a flattened scene.  Replaces nothing in the reference (it has no device); it is the
 * device-side image of GeometrySet + Tunnel::grid / Tunnel::root.                           */
int rtb_scene_upload(rtb_ctx *ctx, const rtb_flat_scene *flat, rtb_scene **scene);
int rtb_scene_free(rtb_ctx *ctx, rtb_scene *scene);
int64_t rtb_scene_device_bytes(const rtb_scene *scene);
/* bytes rtb_scene_upload copied from host memory for this scene (raw triangle records, materials, accelerator arrays);
 * the packed triangle streams and the pair stream of the list scans are produced on the device                      */
int64_t rtb_scene_upload_bytes(const rtb_scene *scene);
/* Inspection: canonical structure hash of the grid resident on the device (the arrays are read back), the hash
 * tests/golden and the host API use for grids; stats = {dims x, y, z, occupied cells, triangle references, longest
 * cell list}.  RTB_ERR_UNSUPPORTED for other accelerators.                                                          */
int rtb_scene_grid_hash(rtb_ctx *ctx, const rtb_scene *scene, uint64_t *hash, int64_t stats[6]);
/* Inspection: the k-d tree resident on the device (uploaded, or built there) read back in the layout of rtb_flat_scene:
 * nodes[counts[0]] in pre-order, leaf_tris[counts[1]]; counts[2] = levels of a device build (0 otherwise); box = root box
 * (min xyz, size xyz).  Either array may be NULL to query the counts first.                                            */
int rtb_scene_kd_download(rtb_ctx *ctx, const rtb_scene *scene, rtb_kdnode *nodes, int64_t node_cap, uint32_t *leaf_tris,
                          int64_t ref_cap, int64_t counts[3], float box[6]);

/* Replaces `int Render(GeometrySet&, PerspectiveCamera&, RenderSetting&, ProgressCallback)`
 * (reference MainWindow.cpp:251-303, the RenderProc of Scripts.h:11-12): ray generation,
 * trace()/radiance() and the framebuffer store for this rank's rows.  `rgb_out` is a HOST
 * buffer of rtb_shard_rows(frame)*width*3 floats (bytes with RTB_OUTPUT_RGB8, which also runs the
 * output stage of MainWindow.cpp:305-311); the device->host copy is part of the call.         */
int rtb_render(rtb_ctx *ctx, const rtb_scene *scene, const rtb_camera *camera,
               const rtb_render_setting *setting, const rtb_frame *frame, void *rgb_out,
               rtb_stats *stats);

/* Same, writing to a DEVICE buffer on `stream` without synchronising (stats are filled only if
 * non-NULL, which forces a sync).  Used to keep the framebuffer resident for the NCCL gather.   */
int rtb_render_device(rtb_ctx *ctx, const rtb_scene *scene, const rtb_camera *camera,
                      const rtb_render_setting *setting, const rtb_frame *frame,
                      void *rgb_device, void *stream, rtb_stats *stats);

/* Multi-GPU assembly: `gathered` is the NCCL all-gather of every rank's rtb_render_device output,
 * [world][rows_per_rank][width][3] floats (each rank's rows padded to rows_per_rank); writes the
 * frame in image row order [height][width][3] to `image`.  Both are DEVICE buffers; asynchronous
 * on `stream`.  (The reference has a single framebuffer, MainWindow.cpp:257; this is its
 * reassembly after tile-row sharding.)                                                        */
int rtb_unshard_device(rtb_ctx *ctx, const void *gathered, void *image, int32_t width, int32_t height,
                       int32_t world, int32_t row_block, int64_t rows_per_rank, void *stream);

/* Parity hook: primary rays only (pixel centres, MainWindow.cpp:294-297) through
 * GeometrySet::intersect.  Arrays are row-major [y*width + x] HOST buffers, any may be NULL.
 * seq_* record the accelerator steps of the TUNNEL prim (cell indices / pre-order node ids):
 * seq_hash folds h = (h ^ id) * 0x100000001b3 from 0xcbf29ce484222325; seq_buf keeps the first
 * seq_cap ids per ray (-1 padded).                                                            */
int rtb_trace_primary(rtb_ctx *ctx, const rtb_scene *scene, const rtb_camera *camera,
                      int32_t width, int32_t height, int32_t *hit_id, float *hit_t,
                      int32_t *seq_len, uint64_t *seq_hash, int32_t *seq_buf, int32_t seq_cap);

/* Replaces `IntersectResult Geometry::intersect(Ray&)` (reference Geometry.h:17) and PT's
 * `Accelerator::intersect` (PerformanceTest/Accelerator.h:6-15) for a BATCH of rays:
 * rays = [n][6] origin, direction (HOST); outputs (HOST, any may be NULL): hit id (-1 miss),
 * distance, position xyz, normal xyz (the stored, unflipped normal).                         */
int rtb_intersect_rays(rtb_ctx *ctx, const rtb_scene *scene, int64_t n, const float *rays,
                       int32_t *hit_id, float *hit_t, float *position, float *normal);

/* Replaces the ray loop of the reference's PerformanceTest console benchmark
 * (src/PerformanceTest/main.cpp:29-59 `trace`, 151-162): every ray of the batch [n][6] (HOST) is
 * mirror-reflected off whatever it hits until it reaches a PLANE geometry (the plane closing the tunnel
 * exit), misses, or has been traced max_depth times (reference: 200).  Outputs (HOST, any may be NULL):
 * reached (1 = stopped on a plane), depth (hits counted), hit id and position of the last hit,
 * the total number of rays traced and the kernel's CUDA-event time.                                   */
int rtb_bounce_rays(rtb_ctx *ctx, const rtb_scene *scene, int64_t n, const float *rays, int32_t max_depth,
                    int32_t *reached, int32_t *depth, int32_t *last_id, float *last_pos, int64_t *total_rays,
                    float *kernel_ms);

/* As rtb_unshard_device for column-block shards (rtb_frame.col_block): `gathered` = [world][height][width / world][3]. */
int rtb_unshard_cols_device(rtb_ctx *ctx, const void *gathered, void *image, int32_t width, int32_t height,
                            int32_t world, int32_t row_block, int32_t col_block, void *stream);

/* Self-test of the list scans' rejection test (csrc/rtb_pretest.h): n random (ray, triangle) pairs -- triangles of
 * widely varying size, slivers, rays aimed at the triangle's boundary -- are decided by the scalar function (the one
 * tests/ checks against the oracle on the CPU) and by the packed-FP32 two-triangle variant the kernels run;
 * *mismatches = pairs on which the two disagree (must be 0), *rejected = pairs the test rejects (so a caller can see
 * that both outcomes occur).  No reference counterpart: diagnostic entry point.                                     */
int rtb_selftest_pretest(rtb_ctx *ctx, int64_t n, uint64_t seed, int64_t *mismatches, int64_t *rejected);

/* Progress reporting -- replaces `progress(y + 1, height)` of the reference's row loop (MainWindow.cpp:271; the
 * ProgressCallback of Scripts.h:9).  While rtb_render / rtb_multi_render wait for the frame they poll the number of
 * finished tiles (a device counter, read over a side stream; the kernels are not interrupted) and call `fn` from the
 * calling host thread, every ~0.25 ms and once more when the frame is complete (done == total).  NULL turns it off.  */
typedef void (*rtb_progress_fn)(int64_t tiles_done, int64_t tiles_total, void *user);
int rtb_set_progress(rtb_ctx *ctx, rtb_progress_fn fn, void *user);

/* Forget the tile schedule the context has learnt (heaviest-first order and latency tiers of the last view): the next
 * frame runs like the first frame of a new view -- raster order, one throughput kernel, every tile measured.  The
 * reference renderer keeps no state between frames (MainWindow.cpp:251-303); this makes that case measurable.       */
int rtb_forget_schedule(rtb_ctx *ctx);

/* ---- one frame buffer, several devices (RTB_LAYOUT_GLOBAL) -------------------------------------------------------
 * Replaces the single framebuffer `colors` every OpenMP thread of the reference's Render writes its rows into
 * (MainWindow.cpp:257, 267-276).
 *
 * A: one process per device (torch.distributed / MPI plumbing).  The owner rank allocates the frame with
 *    rtb_device_alloc (plain cudaMalloc memory: exportable), exports it with rtb_ipc_export, sends the 64 handle bytes
 *    to its peers by any means, each peer maps it with rtb_ipc_open (peer access over NVLink is enabled as needed) and
 *    passes the mapped pointer to rtb_render_device with RTB_LAYOUT_GLOBAL.  A barrier of the caller closes the step. */
#define RTB_IPC_HANDLE_BYTES 64
int rtb_device_alloc(rtb_ctx *ctx, size_t bytes, void **out);
int rtb_device_free(rtb_ctx *ctx, void *p);
int rtb_device_download(rtb_ctx *ctx, const void *device_ptr, void *host, size_t bytes); /* synchronous read-back of such a frame */
int rtb_ipc_export(rtb_ctx *ctx, void *device_ptr, unsigned char handle[RTB_IPC_HANDLE_BYTES]);
int rtb_ipc_open(rtb_ctx *ctx, const unsigned char handle[RTB_IPC_HANDLE_BYTES], void **mapped);
int rtb_ipc_close(rtb_ctx *ctx, void *mapped);
/* Alternative to rendering straight into the mapped frame: render the shard into a LOCAL device buffer (row-major layout of
 * rtb_render_device), then move it to its place in the whole frame with one streaming pass of 128-bit copies on `stream`
 * (the frame may be peer-mapped: one bulk NVLink transfer instead of tile-sized stores issued while the kernels render).
 * `frame` describes the shard (rank / world / row_block / col_block); float frames only.                              */
int rtb_scatter_shard_device(rtb_ctx *ctx, const void *local, void *frame_buffer, const rtb_frame *frame, void *stream);

/* B: ONE host thread drives n devices (SURVEY 8b threading row; the drop-in behind Scripts.h:11-12 RenderProc).
 *    rtb_multi_render shards the frame over the devices (tile rows up to 4 devices, column blocks above), every device
 *    stores its tiles straight into `rgb_out` -- ONE assembled host frame, row-major [height][width][3] floats (bytes
 *    with RTB_OUTPUT_RGB8), or the reference order with RTB_LAYOUT_REFERENCE -- and the call returns when all devices
 *    are done.  `rgb_out` from rtb_host_alloc is written directly over PCIe by all devices at once; any other host
 *    buffer goes through a page-locked frame owned by the library and one host copy.  frame->rank / world are ignored.
 *    stats: counts are summed over the devices, kernel_ms / total_ms are the maximum.                                */
typedef struct rtb_multi rtb_multi;
typedef struct rtb_multi_scene rtb_multi_scene;
int rtb_multi_init(int n_devices, const int *devices /* NULL: 0 .. n_devices - 1 */, rtb_multi **out);
int rtb_multi_shutdown(rtb_multi *m);
int rtb_multi_count(const rtb_multi *m);
rtb_ctx *rtb_multi_ctx(rtb_multi *m, int i);
const char *rtb_multi_last_error(const rtb_multi *m);
int rtb_multi_set_progress(rtb_multi *m, rtb_progress_fn fn, void *user);
int rtb_multi_scene_upload(rtb_multi *m, const rtb_flat_scene *flat, rtb_multi_scene **scene); /* replicated on every device */
int rtb_multi_scene_free(rtb_multi *m, rtb_multi_scene *scene);
int64_t rtb_multi_scene_upload_bytes(const rtb_multi_scene *scene); /* summed over the devices */
int rtb_multi_render(rtb_multi *m, const rtb_multi_scene *scene, const rtb_camera *camera,
                     const rtb_render_setting *setting, const rtb_frame *frame, void *rgb_out, rtb_stats *stats);

#ifdef __cplusplus
}
#endif
#endif /* RTB_H */
