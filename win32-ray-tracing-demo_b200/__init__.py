"""rtb200 -- Python harness over the C ABI (include/rtb.h) and the C++ host API (host/rt.h).

The product is the two in-tree shared libraries; this module only binds them with ctypes for the
tests, bench.py and torch.distributed plumbing.  The directory name is not a Python identifier, so
import it through `rtb200.py` at the repo root (`import rtb200`).

There is no CPU fallback anywhere below: every compute call goes to librtb200.so's CUDA kernels
and raises RtbError when the library or a CUDA device is missing.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CUDA_LIB = os.path.join(PKG, os.environ.get("RTB_CUDA_LIB_NAME", "librtb200.so"))  # RTB_CUDA_LIB_NAME: an experiment build (tools/)
HOST_LIB = os.path.join(PKG, "librtb200_host.so")

ALGORITHMS = {"linear": 0, "rgrid": 1, "fgrid": 2, "kd": 3, "sah": 4, "convex": 5, "convexsimple": 6}
STAT_NAMES = ["n_top", "n_tris", "grid_x", "grid_y", "grid_z", "cells_nonempty", "cell_entries",
              "cell_max", "kd_nodes", "kd_leaves", "kd_leaf_refs", "kd_max_depth"]
LAYOUT_ROWMAJOR, LAYOUT_REFERENCE, LAYOUT_GLOBAL, OUTPUT_RGB8, OUTPUT_MOMENTS = 0, 1, 2, 4, 8
IPC_HANDLE_BYTES = 64


class RtbError(RuntimeError):
    pass


# ---- ctypes mirrors of include/rtb.h ---------------------------------------------------------
class Material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("diffusiveness", C.c_float), ("reflectiveness", C.c_float),
                ("refractiveness", C.c_float), ("refractive_index", C.c_float), ("a", C.c_float * 3),
                ("b", C.c_float * 3), ("scale", C.c_float), ("p", C.c_float), ("dir", C.c_int32),
                ("pad_", C.c_int32 * 2)]


class Prim(C.Structure):
    _fields_ = [("type", C.c_int32), ("material", C.c_int32), ("base_id", C.c_int32),
                ("first", C.c_int32), ("count", C.c_int32), ("v", C.c_float * 7)]


class KdNode(C.Structure):
    _fields_ = [("a", C.c_uint32), ("b", C.c_uint32)]


class CellWord(C.Structure):
    _fields_ = [("bits", C.c_uint32), ("rank", C.c_uint32)]


class FlatScene(C.Structure):
    _fields_ = [
        ("n_prims", C.c_int32), ("prims", C.POINTER(Prim)),
        ("n_materials", C.c_int32), ("materials", C.POINTER(Material)),
        ("n_top", C.c_int32),
        ("n_loose", C.c_int32), ("loose_tri", C.POINTER(C.c_float)),
        ("n_tris", C.c_int32), ("tri", C.POINTER(C.c_float)), ("tri_material", C.POINTER(C.c_int32)),
        ("accel", C.c_int32),
        ("grid_origin", C.c_float * 3), ("grid_cell", C.c_float * 3), ("grid_dims", C.c_int32 * 3),
        ("grid_build_resolution", C.c_int32),
        ("n_cellwords", C.c_int64), ("grid_words", C.POINTER(CellWord)),
        ("n_cells_used", C.c_int64), ("grid_cell_start", C.POINTER(C.c_uint32)),
        ("n_cell_refs", C.c_int64), ("grid_cell_tris", C.POINTER(C.c_uint32)),
        ("kd_min", C.c_float * 3), ("kd_max", C.c_float * 3),
        ("n_kd_nodes", C.c_int32), ("kd_nodes", C.POINTER(KdNode)),
        ("n_kd_refs", C.c_int64), ("kd_leaf_tris", C.POINTER(C.c_uint32)),
        ("n_cx_path", C.c_int32), ("cx_frames", C.POINTER(C.c_float)),
        ("n_cx_edges", C.c_int32), ("cx_edges", C.POINTER(C.c_float)),
        ("cx_width", C.c_float), ("cx_height", C.c_float),
        ("cx_table_size", C.c_int32), ("cx_round_bins", C.c_int32),
        ("cx_cell_status", C.POINTER(C.c_uint8)), ("cx_cell_range", C.POINTER(C.c_int16)),
        ("cx_order", C.POINTER(C.c_uint16)),
        ("grid_build_exact", C.c_int32),
        ("kd_build_leaf_size", C.c_int32), ("kd_build_max_depth", C.c_int32), ("kd_build_candidates", C.c_int32),
        ("arrays_page_locked", C.c_int32),
    ]

    def arrays(self):
        """(address, bytes) of every array rtb_scene_upload copies for this scene (include/rtb.h: rtb_flat_scene)."""
        def a(ptr, n, elem):
            return (C.cast(ptr, C.c_void_p).value, int(n) * elem) if ptr and n > 0 else None
        out = [a(self.loose_tri, self.n_loose, 48), a(self.tri, self.n_tris, 48), a(self.tri_material, self.n_tris, 4)]
        if self.accel in (ALGORITHMS["rgrid"], ALGORITHMS["fgrid"]) and self.grid_words:
            out += [a(self.grid_words, self.n_cellwords, 8), a(self.grid_cell_start, self.n_cells_used + 1, 4),
                    a(self.grid_cell_tris, self.n_cell_refs, 4)]
        if self.accel in (ALGORITHMS["kd"], ALGORITHMS["sah"]) and self.kd_nodes:
            out += [a(self.kd_nodes, self.n_kd_nodes, 8), a(self.kd_leaf_tris, self.n_kd_refs, 4)]
        return [x for x in out if x]


class Camera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("front", C.c_float * 3), ("up", C.c_float * 3),
                ("right", C.c_float * 3), ("xcenter", C.c_float), ("fov_scale", C.c_float),
                ("forward", C.c_float)]


class RenderSetting(C.Structure):
    _fields_ = [("enable_monte_carlo", C.c_int32), ("max_depth", C.c_int32),
                ("termination_depth", C.c_int32), ("single_tracing_depth", C.c_int32)]


INT_MAX = 2 ** 31 - 1
SETTINGS = {  # reference RenderSetting.h:40-78
    "simple": (0, 20, INT_MAX, 0), "default": (1, INT_MAX, 5, 2),
    "highspeed": (1, 6, 2, 0), "highquality": (1, 8, INT_MAX, INT_MAX),
}


class Frame(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("samples", C.c_int32),
                ("seed", C.c_uint64), ("rank", C.c_int32), ("world", C.c_int32),
                ("row_block", C.c_int32), ("layout", C.c_int32), ("counters", C.c_int32),
                ("col_block", C.c_int32), ("sample_first", C.c_int32), ("sample_count", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n_rays", C.c_int64), ("n_tri_tests", C.c_int64), ("n_steps", C.c_int64),
                ("n_local_rows", C.c_int64), ("kernel_ms", C.c_float), ("total_ms", C.c_float),
                ("n_launches", C.c_int32), ("pad_", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "pad_"}


ABI_SYMBOLS = ["rtb_init", "rtb_shutdown", "rtb_last_error", "rtb_abi_version", "rtb_shard_rows", "rtb_shard_width",
               "rtb_unshard_cols_device",
               "rtb_host_alloc", "rtb_host_free", "rtb_host_register", "rtb_host_unregister",
               "rtb_scene_upload", "rtb_scene_free", "rtb_scene_device_bytes", "rtb_scene_upload_bytes", "rtb_scene_grid_hash", "rtb_scene_kd_download", "rtb_render",
               "rtb_render_device", "rtb_unshard_device", "rtb_trace_primary", "rtb_intersect_rays", "rtb_bounce_rays",
               "rtb_selftest_pretest", "rtb_kd_validate", "rtb_forget_schedule", "rtb_set_progress", "rtb_multi_set_progress",
               "rtb_device_alloc", "rtb_device_free", "rtb_device_download", "rtb_ipc_export", "rtb_ipc_open", "rtb_ipc_close", "rtb_scatter_shard_device",
               "rtb_multi_init", "rtb_multi_shutdown", "rtb_multi_count", "rtb_multi_ctx", "rtb_multi_last_error",
               "rtb_multi_scene_upload", "rtb_multi_scene_free", "rtb_multi_scene_upload_bytes", "rtb_multi_render"]

_cuda = None
_host = None


def cuda_lib():
    """librtb200.so; raises RtbError (never falls back) when it has not been built."""
    global _cuda
    if _cuda is None:
        if not os.path.exists(CUDA_LIB):
            raise RtbError(f"{CUDA_LIB} is missing: run `python __graft_entry__.py build` (there is no CPU fallback)")
        lib = C.CDLL(CUDA_LIB, mode=C.RTLD_GLOBAL)
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        lib.rtb_init.argtypes = [C.c_int, C.POINTER(vp)]
        lib.rtb_shutdown.argtypes = [vp]
        lib.rtb_last_error.argtypes = [vp]
        lib.rtb_last_error.restype = C.c_char_p
        lib.rtb_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
        lib.rtb_host_free.argtypes = [vp]
        lib.rtb_host_register.argtypes = [vp, C.c_size_t]
        lib.rtb_host_unregister.argtypes = [vp]
        lib.rtb_shard_rows.argtypes = [C.POINTER(Frame)]
        lib.rtb_shard_rows.restype = i64
        lib.rtb_shard_width.argtypes = [C.POINTER(Frame)]
        lib.rtb_shard_width.restype = i64
        lib.rtb_scene_upload.argtypes = [vp, C.POINTER(FlatScene), C.POINTER(vp)]
        lib.rtb_scene_free.argtypes = [vp, vp]
        lib.rtb_scene_device_bytes.argtypes = [vp]
        lib.rtb_scene_device_bytes.restype = i64
        lib.rtb_scene_upload_bytes.argtypes = [vp]
        lib.rtb_scene_upload_bytes.restype = i64
        lib.rtb_scene_grid_hash.argtypes = [vp, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
        lib.rtb_scene_grid_hash.restype = C.c_int
        lib.rtb_scene_kd_download.argtypes = [vp, vp, vp, i64, vp, i64, C.POINTER(C.c_int64), C.POINTER(C.c_float)]
        lib.rtb_render.argtypes = [vp, vp, C.POINTER(Camera), C.POINTER(RenderSetting), C.POINTER(Frame), vp,
                                   C.POINTER(Stats)]
        lib.rtb_render_device.argtypes = [vp, vp, C.POINTER(Camera), C.POINTER(RenderSetting), C.POINTER(Frame),
                                          vp, vp, C.POINTER(Stats)]
        lib.rtb_unshard_device.argtypes = [vp, vp, vp, i32, i32, i32, i32, i64, vp]
        lib.rtb_unshard_cols_device.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp]
        lib.rtb_trace_primary.argtypes = [vp, vp, C.POINTER(Camera), i32, i32, vp, vp, vp, vp, vp, i32]
        lib.rtb_intersect_rays.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
        lib.rtb_forget_schedule.argtypes = [vp]
        lib.rtb_device_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        lib.rtb_device_free.argtypes = [vp, vp]
        lib.rtb_device_download.argtypes = [vp, vp, vp, C.c_size_t]
        lib.rtb_ipc_export.argtypes = [vp, vp, C.c_char_p]
        lib.rtb_ipc_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
        lib.rtb_ipc_close.argtypes = [vp, vp]
        lib.rtb_scatter_shard_device.argtypes = [vp, vp, vp, C.POINTER(Frame), vp]
        lib.rtb_multi_init.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]
        lib.rtb_multi_shutdown.argtypes = [vp]
        lib.rtb_multi_count.argtypes = [vp]
        lib.rtb_multi_ctx.argtypes = [vp, C.c_int]
        lib.rtb_multi_ctx.restype = vp
        lib.rtb_multi_last_error.argtypes = [vp]
        lib.rtb_multi_last_error.restype = C.c_char_p
        lib.rtb_multi_scene_upload.argtypes = [vp, C.POINTER(FlatScene), C.POINTER(vp)]
        lib.rtb_multi_scene_free.argtypes = [vp, vp]
        lib.rtb_multi_scene_upload_bytes.argtypes = [vp]
        lib.rtb_multi_scene_upload_bytes.restype = i64
        lib.rtb_multi_render.argtypes = [vp, vp, C.POINTER(Camera), C.POINTER(RenderSetting), C.POINTER(Frame), vp, C.POINTER(Stats)]
        _cuda = lib
    return _cuda


def host_lib():
    global _host
    if _host is None:
        cuda_lib()
        if not os.path.exists(HOST_LIB):
            raise RtbError(f"{HOST_LIB} is missing: run `python __graft_entry__.py build`")
        lib = C.CDLL(HOST_LIB)
        vp = C.c_void_p
        lib.rtbh_preset_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p]
        lib.rtbh_preset_create.restype = vp
        lib.rtbh_perf_scene_create.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]
        lib.rtbh_perf_scene_create.restype = vp
        lib.rtbh_free.argtypes = [vp]
        lib.rtbh_flat.argtypes = [vp]
        lib.rtbh_flat.restype = C.POINTER(FlatScene)
        lib.rtbh_camera.argtypes = [vp]
        lib.rtbh_camera.restype = C.POINTER(Camera)
        lib.rtbh_setting.argtypes = [vp]
        lib.rtbh_setting.restype = C.POINTER(RenderSetting)
        for f in ("rtbh_prepare_ms", "rtbh_build_ms"):
            getattr(lib, f).argtypes = [vp]
            getattr(lib, f).restype = C.c_double
        lib.rtbh_host_bytes.argtypes = [vp]
        lib.rtbh_host_bytes.restype = C.c_int64
        lib.rtbh_stats.argtypes = [vp, C.POINTER(C.c_int64)]
        for f in ("rtbh_tri_hash", "rtbh_struct_hash"):
            getattr(lib, f).argtypes = [vp]
            getattr(lib, f).restype = C.c_uint64
        lib.rtbh_script_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int,
                                        C.c_char_p, vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(Stats)]
        lib.rtbh_script_run_ex.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                           C.c_char_p, vp, C.POINTER(C.c_int), C.POINTER(Stats), C.POINTER(C.c_int), C.c_char_p, C.c_int]
        lib.rtbh_script_run8.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int,
                                         C.c_char_p, vp, C.c_char_p, C.POINTER(C.c_int), C.POINTER(Stats)]
        lib.rtbh_perf_test.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp, vp,
                                       C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.rtbh_last_error.restype = C.c_char_p
        lib.rtbh_intersect_batch.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp]
        lib.rtbh_intersect_one.argtypes = [vp, vp, vp, vp, vp, vp]
        _host = lib
    return _host


def stl_fixture():
    return os.path.join(PKG, "assets", "ball.stl")  # the 528-triangle mesh of preset 3 (reference Scripts.cpp:113-168 reads ball.stl)


def _alg(a):
    return ALGORITHMS[a] if isinstance(a, str) else int(a)


def make_setting(name):
    return RenderSetting(*SETTINGS[name])


# ---- host side: preset scenes built by the C++ host API ----------------------------------------
class PresetScene:
    """A reference preset (Scripts.cpp Script1..5) built and flattened by the C++ host library."""

    def __init__(self, preset, algorithm="linear", segments=150, stl_path=None):
        lib = host_lib()
        self.preset, self.algorithm, self.segments = preset, algorithm, segments
        self._h = lib.rtbh_preset_create(preset, _alg(algorithm), segments, (stl_path or stl_fixture()).encode())
        if not self._h:
            raise RtbError("rtbh_preset_create failed (bad preset / algorithm / STL path)")
        self.flat = lib.rtbh_flat(self._h)
        self.camera = lib.rtbh_camera(self._h).contents
        self.setting = lib.rtbh_setting(self._h).contents

    def pin(self):
        """Page-lock the flat scene's arrays where they lie (rtb_host_register) and declare them page-locked and stable
        (rtb_flat_scene.arrays_page_locked): uploads then copy host -> device straight out of them.  An array that cannot be
        registered (it shares a page with one that is) stays pageable; the library stages such an array as before."""
        if getattr(self, "_pinned", None):
            return self
        lib = cuda_lib()
        self._pinned = [addr for addr, nbytes in self.flat.contents.arrays()
                        if nbytes >= (64 << 10) and lib.rtb_host_register(C.c_void_p(addr), C.c_size_t(nbytes)) == 0]
        self.flat.contents.arrays_page_locked = 1
        return self

    def unpin(self):
        for addr in getattr(self, "_pinned", None) or []:
            cuda_lib().rtb_host_unregister(C.c_void_p(addr))
        self._pinned = []
        if self._h:
            self.flat.contents.arrays_page_locked = 0

    def close(self):
        if self._h:
            self.unpin()
            host_lib().rtbh_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def prepare_ms(self):
        return host_lib().rtbh_prepare_ms(self._h)

    @property
    def build_ms(self):
        return host_lib().rtbh_build_ms(self._h)

    @property
    def host_bytes(self):
        return host_lib().rtbh_host_bytes(self._h)

    def stats(self):
        out = (C.c_int64 * 16)()
        host_lib().rtbh_stats(self._h, out)
        return {k: int(out[i]) for i, k in enumerate(STAT_NAMES)}

    def struct_hash(self):
        return int(host_lib().rtbh_struct_hash(self._h))

    def tri_hash(self):
        return int(host_lib().rtbh_tri_hash(self._h))

    def triangles(self):
        f = self.flat.contents
        n = f.n_tris
        if n == 0:
            return np.zeros((0, 12), np.float32), np.zeros(0, np.int32)
        tri = np.ctypeslib.as_array(f.tri, shape=(n, 12)).copy()
        mat = np.ctypeslib.as_array(f.tri_material, shape=(n,)).copy()
        return tri, (mat == mat[-1]).astype(np.int32)

    def intersect_batch(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        hid, ht = np.zeros(n, np.int32), np.zeros(n, np.float32)
        pos, nrm = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        rc = host_lib().rtbh_intersect_batch(self._h, n, rays.ctypes.data, hid.ctypes.data, ht.ctypes.data,
                                             pos.ctypes.data, nrm.ctypes.data)
        if rc != 0:
            raise RtbError("GeometrySet::intersectBatch failed: " + host_lib().rtbh_last_error().decode())
        return hid, ht, pos, nrm

    def intersect_one(self, ray):
        ray = np.ascontiguousarray(ray, np.float32).reshape(6)
        hid, ht = C.c_int32(-1), C.c_float(0)
        pos, nrm = np.zeros(3, np.float32), np.zeros(3, np.float32)
        host_lib().rtbh_intersect_one(self._h, ray.ctypes.data, C.byref(hid), C.byref(ht), pos.ctypes.data, nrm.ctypes.data)
        return hid.value, ht.value, pos, nrm


def set_exact_grid_binning(on):
    """Tunnels built from now on bin triangles into grid cells with the exact overlap test (reference
    Triangle.cpp:152-199, compiled out at Tunnel.cpp:435-445) instead of the bounding-box overlap."""
    host_lib().rtbh_set_exact_grid_binning(1 if on else 0)


def set_kd_on_device(on):
    """Tunnels built from now on leave a KdTreeSAH accelerator to the device builder of rtb_scene_upload."""
    host_lib().rtbh_set_kd_on_device(1 if on else 0)


def set_grid_on_device(on):
    """Tunnels built from now on leave RegularGrid / FlatGrid to the device builder of rtb_scene_upload."""
    host_lib().rtbh_set_grid_on_device(1 if on else 0)


class PerfScene(PresetScene):
    """The scene of the reference's PerformanceTest program (tunnel built by that program's own generator and
    k-d builder + the plane closing its exit), built and flattened by the C++ host library."""

    def __init__(self, radius=2000.0, angle=1.5708, arch_seg=150, path_seg=150, algorithm="sah"):
        lib = host_lib()
        self.preset, self.algorithm, self.segments = 0, algorithm, path_seg
        self._h = lib.rtbh_perf_scene_create(radius, angle, arch_seg, path_seg, _alg(algorithm))
        if not self._h:
            raise RtbError("rtbh_perf_scene_create failed (bad algorithm / tessellation)")
        self.flat = lib.rtbh_flat(self._h)
        self.camera = lib.rtbh_camera(self._h).contents
        self.setting = lib.rtbh_setting(self._h).contents


def script_run(preset, algorithm="linear", segments=150, width=400, height=300, samples=1, seed=0, device=0,
               stl_path=None):
    """The drop-in path: Script::Run(CudaRenderer::Render, ...) -- returns (image[y][x][3], info)."""
    lib = host_lib()
    rgb = np.zeros((width, height, 3), np.float32)
    prep, exe, st = C.c_int(0), C.c_int(0), Stats()
    rc = lib.rtbh_script_run(preset, _alg(algorithm), segments, width, height, samples, seed, device,
                             (stl_path or stl_fixture()).encode(), rgb.ctypes.data, C.byref(prep), C.byref(exe),
                             C.byref(st))
    if rc != 0:
        raise RtbError(f"Script::Run failed rc={rc}: {lib.rtbh_last_error().decode()}")
    info = st.as_dict()
    info.update(prepare_ms=prep.value, exec_ms=exe.value)
    return np.ascontiguousarray(rgb.transpose(1, 0, 2)), info


def script_run_ex(preset, algorithm="linear", segments=150, width=400, height=300, samples=1, seed=0, device=0, n_devices=1,
                  stl_path=None):
    """Script::Run(CudaRenderer::Render, ...) on n_devices GPUs with the reference's ProgressCallback and LogCallback attached;
    returns (image[y][x][3] or None, info) -- info carries rc, the progress calls seen and the log text."""
    lib = host_lib()
    rgb = np.zeros((width, height, 3), np.float32)
    exe, st, prog = C.c_int(0), Stats(), (C.c_int * 4)()
    log = C.create_string_buffer(4096)
    rc = lib.rtbh_script_run_ex(preset, _alg(algorithm), segments, width, height, samples, seed, device, n_devices,
                                (stl_path or stl_fixture()).encode(), rgb.ctypes.data, C.byref(exe), C.byref(st), prog, log, 4096)
    info = st.as_dict()
    info.update(rc=rc, exec_ms=exe.value, progress_calls=prog[0], progress_last=prog[1], progress_total=prog[2],
                progress_monotone=bool(prog[3]), log=log.value.decode(errors="replace"))
    return (np.ascontiguousarray(rgb.transpose(1, 0, 2)) if rc == 0 else None), info


def script_run8(preset, algorithm="linear", segments=150, width=400, height=300, samples=1, seed=0, device=0,
                stl_path=None, bmp_path=None):
    """Drop-in path with the reference's 8-bit output stage; returns (uint8 image[y][x][3], info)."""
    lib = host_lib()
    rgb = np.zeros((width, height, 3), np.uint8)
    exe, st = C.c_int(0), Stats()
    rc = lib.rtbh_script_run8(preset, _alg(algorithm), segments, width, height, samples, seed, device,
                              (stl_path or stl_fixture()).encode(), rgb.ctypes.data,
                              bmp_path.encode() if bmp_path else None, C.byref(exe), C.byref(st))
    if rc != 0:
        raise RtbError(f"Script::Run (8-bit) failed rc={rc}: {lib.rtbh_last_error().decode()}")
    info = st.as_dict()
    info.update(exec_ms=exe.value)
    return np.ascontiguousarray(rgb.transpose(1, 0, 2)), info


def perf_test(xy, radius=2000.0, angle=1.5708, arch_seg=150, path_seg=150, algorithm="sah", max_depth=200):
    """The reference's PerformanceTest console benchmark on the GPU (host C++: rt::PerformanceTest)."""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    n = xy.shape[0]
    out = {"reached": np.zeros(n, np.int32), "depth": np.zeros(n, np.int32), "last_id": np.zeros(n, np.int32),
           "last_pos": np.zeros((n, 3), np.float32)}
    total, b, pre, tr = C.c_int64(0), C.c_double(0), C.c_double(0), C.c_double(0)
    rc = host_lib().rtbh_perf_test(radius, angle, arch_seg, path_seg, _alg(algorithm), n, xy.ctypes.data, max_depth,
                                   out["reached"].ctypes.data, out["depth"].ctypes.data, out["last_id"].ctypes.data,
                                   out["last_pos"].ctypes.data, C.byref(total), C.byref(b), C.byref(pre), C.byref(tr))
    if rc != 0:
        raise RtbError(f"PerformanceTest failed rc={rc}: {host_lib().rtbh_last_error().decode()}")
    out.update(total_rays=total.value, build_ms=b.value, preprocess_ms=pre.value, trace_ms=tr.value)
    return out


# ---- device side: context + uploaded scene over the C ABI --------------------------------------
class Context:
    def __init__(self, device=0):
        self._lib = cuda_lib()
        self._h = C.c_void_p()
        rc = self._lib.rtb_init(device, C.byref(self._h))
        if rc != 0:
            raise RtbError(f"rtb_init({device}) failed rc={rc}: {self._lib.rtb_last_error(None).decode()}")
        self.device = device

    def close(self):
        if self._h:
            self._lib.rtb_shutdown(self._h)
            self._h = C.c_void_p()

    def _check(self, rc, what):
        if rc != 0:
            raise RtbError(f"{what} failed rc={rc}: {self._lib.rtb_last_error(self._h).decode()}")

    def upload(self, flat):
        return DeviceScene(self, flat)

    def forget_schedule(self):
        """The next frame runs like the first frame of a new view (raster order, no latency tiers)."""
        self._check(self._lib.rtb_forget_schedule(self._h), "rtb_forget_schedule")

    # ---- one frame buffer, several devices (RTB_LAYOUT_GLOBAL; include/rtb.h section A) ----
    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self._lib.rtb_device_alloc(self._h, nbytes, C.byref(p)), "rtb_device_alloc")
        return p.value

    def device_free(self, ptr):
        self._check(self._lib.rtb_device_free(self._h, C.c_void_p(ptr)), "rtb_device_free")

    def device_download(self, ptr, out):
        """synchronous copy of out.nbytes bytes of a device frame into the numpy array `out`"""
        self._check(self._lib.rtb_device_download(self._h, C.c_void_p(ptr), out.ctypes.data, out.nbytes), "rtb_device_download")
        return out

    def ipc_export(self, ptr):
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._check(self._lib.rtb_ipc_export(self._h, C.c_void_p(ptr), buf), "rtb_ipc_export")
        return buf.raw

    def ipc_open(self, handle):
        p = C.c_void_p()
        self._check(self._lib.rtb_ipc_open(self._h, C.create_string_buffer(bytes(handle), IPC_HANDLE_BYTES), C.byref(p)), "rtb_ipc_open")
        return p.value

    def scatter_shard_device(self, local_ptr, frame_ptr, frame, stream=0):
        """rtb_scatter_shard_device: a rank's local shard image -> its place in the whole (possibly peer-mapped) frame."""
        self._check(self._lib.rtb_scatter_shard_device(self._h, C.c_void_p(local_ptr), C.c_void_p(frame_ptr), C.byref(frame), C.c_void_p(stream)),
                    "rtb_scatter_shard_device")

    def ipc_close(self, ptr):
        self._check(self._lib.rtb_ipc_close(self._h, C.c_void_p(ptr)), "rtb_ipc_close")


class MultiContext:
    """rtb_multi_*: one host thread drives n devices; rtb_multi_render returns ONE assembled host frame."""

    def __init__(self, n_devices, devices=None):
        self._lib = cuda_lib()
        self._h = C.c_void_p()
        arr = (C.c_int * n_devices)(*devices) if devices is not None else None
        rc = self._lib.rtb_multi_init(n_devices, arr, C.byref(self._h))
        if rc != 0:
            raise RtbError(f"rtb_multi_init({n_devices}) failed rc={rc}: {self._lib.rtb_last_error(None).decode()}")
        self.n = n_devices

    def close(self):
        if self._h:
            self._lib.rtb_multi_shutdown(self._h)
            self._h = C.c_void_p()

    def _check(self, rc, what):
        if rc != 0:
            raise RtbError(f"{what} failed rc={rc}: {self._lib.rtb_multi_last_error(self._h).decode()}")

    def upload(self, flat):
        return MultiScene(self, flat)


class MultiScene:
    def __init__(self, multi, flat):
        self.multi = multi
        self._h = C.c_void_p()
        multi._check(multi._lib.rtb_multi_scene_upload(multi._h, flat, C.byref(self._h)), "rtb_multi_scene_upload")

    def close(self):
        if self._h:
            self.multi._lib.rtb_multi_scene_free(self.multi._h, self._h)
            self._h = C.c_void_p()

    @property
    def upload_bytes(self):
        return int(self.multi._lib.rtb_multi_scene_upload_bytes(self._h))

    def render(self, camera, setting, frame, out=None):
        """The whole frame into ONE host buffer: [height][width][3] (reference order [width][height][3] with LAYOUT_REFERENCE)."""
        if out is None:
            shape = (frame.width, frame.height, 3) if frame.layout & LAYOUT_REFERENCE else (frame.height, frame.width, 3)
            if frame.layout & OUTPUT_MOMENTS:
                shape = shape[:2] + (6,)
            out = np.zeros(shape, np.uint8 if frame.layout & OUTPUT_RGB8 else np.float32)
        st = Stats()
        self.multi._check(self.multi._lib.rtb_multi_render(self.multi._h, self._h, C.byref(camera), C.byref(setting), C.byref(frame),
                                                           out.ctypes.data, C.byref(st)), "rtb_multi_render")
        return out, st.as_dict()


def unshard_device(ctx, gathered_ptr, image_ptr, width, height, world, row_block, rows_per_rank, stream=0):
    ctx._check(ctx._lib.rtb_unshard_device(ctx._h, C.c_void_p(gathered_ptr), C.c_void_p(image_ptr), width, height, world,
                                           row_block, rows_per_rank, C.c_void_p(stream)), "rtb_unshard_device")


def unshard_cols_device(ctx, gathered_ptr, image_ptr, width, height, world, row_block, col_block, stream=0):
    ctx._check(ctx._lib.rtb_unshard_cols_device(ctx._h, C.c_void_p(gathered_ptr), C.c_void_p(image_ptr), width, height, world,
                                                row_block, col_block, C.c_void_p(stream)), "rtb_unshard_cols_device")


class PinnedArray:
    """float32 numpy view over page-locked host memory from rtb_host_alloc."""

    def __init__(self, shape):
        self.shape = tuple(int(x) for x in shape)
        n = int(np.prod(self.shape))
        self._p = C.c_void_p()
        rc = cuda_lib().rtb_host_alloc(max(n, 1) * 4, C.byref(self._p))
        if rc != 0:
            raise RtbError("rtb_host_alloc failed: " + cuda_lib().rtb_last_error(None).decode())
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_float)), shape=(max(n, 1),))[:n].reshape(self.shape)

    def close(self):
        if self._p:
            cuda_lib().rtb_host_free(self._p)
            self._p = C.c_void_p()


def shard_rows(frame):
    return int(cuda_lib().rtb_shard_rows(C.byref(frame)))


def shard_width(frame):
    return int(cuda_lib().rtb_shard_width(C.byref(frame)))


def make_frame(width, height, samples=1, seed=0, rank=0, world=1, row_block=8, layout=LAYOUT_ROWMAJOR, counters=0, col_block=0,
               sample_first=0, sample_count=0):
    """col_block > 0 (and world > 1): column-block shard, local image [height][width / world] (include/rtb.h).
    sample_count > 0: Monte-Carlo sample shard [sample_first, sample_first + sample_count) of `samples`."""
    return Frame(width, height, samples, seed, rank, world, row_block, layout, counters, col_block, sample_first, sample_count)


def shard_col_indices(width, y, rank, world, row_block, col_block):
    """Global x of each local column of row y of a column-block shard: block row by = y // row_block owns the
    column blocks bx with (bx + by) % world == rank (mirror of localToGlobal in csrc/rtb_kernels.cuh)."""
    by = y // row_block
    shift = (rank - by) % world
    xs = []
    for c in range(width // (world * col_block)):
        bx = c * world + shift
        xs.extend(range(bx * col_block, (bx + 1) * col_block))
    return np.asarray(xs, np.int64)


def sample_shard(samples, rank, world):
    """(sample_first, sample_count) of rank's share of a Monte-Carlo frame's samples: contiguous ranges, the first
    samples % world ranks take one more (rtb_frame.sample_first / sample_count)."""
    base, extra = divmod(samples, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def shard_pixel_mask(width, height, rank, world, row_block=8, col_block=0):
    """bool [height][width]: the pixels of the frame rank `rank` stores with RTB_LAYOUT_GLOBAL (mirror of localToGlobal
    in csrc/rtb_kernels.cuh: row blocks dealt round-robin, or column blocks rotated from block row to block row)."""
    m = np.zeros((height, width), bool)
    if col_block and world > 1:
        for y in range(height):
            m[y, shard_col_indices(width, y, rank, world, row_block, col_block)] = True
    else:
        m[shard_row_indices(height, rank, world, row_block)] = True
    return m


def shard_row_indices(height, rank, world, row_block=8):
    """Global y of each local row of a shard (block b of row_block rows goes to rank b % world)."""
    ys = []
    b = rank
    while b * row_block < height:
        ys.extend(range(b * row_block, min((b + 1) * row_block, height)))
        b += world
    return np.asarray(ys, np.int64)


class DeviceScene:
    def __init__(self, ctx, flat):
        self.ctx = ctx
        self._h = C.c_void_p()
        ctx._check(ctx._lib.rtb_scene_upload(ctx._h, flat, C.byref(self._h)), "rtb_scene_upload")

    def close(self):
        if self._h:
            self.ctx._lib.rtb_scene_free(self.ctx._h, self._h)
            self._h = C.c_void_p()

    @property
    def device_bytes(self):
        return int(self.ctx._lib.rtb_scene_device_bytes(self._h))

    @property
    def upload_bytes(self):
        """bytes the upload copied from host memory (the packed / pair streams are produced on the device)"""
        return int(self.ctx._lib.rtb_scene_upload_bytes(self._h))

    def grid_hash(self):
        """(canonical structure hash, {dims, occupied cells, references, longest list}) of the grid on the device."""
        h, st = C.c_uint64(0), (C.c_int64 * 6)()
        self.ctx._check(self.ctx._lib.rtb_scene_grid_hash(self.ctx._h, self._h, C.byref(h), st), "rtb_scene_grid_hash")
        return int(h.value), dict(grid_x=st[0], grid_y=st[1], grid_z=st[2], cells_nonempty=st[3], cell_entries=st[4], cell_max=st[5])

    def kd_download(self):
        """(nodes [n][2] uint32, leaf_tris [m] uint32, levels, box) of the k-d tree resident on the device."""
        counts, box = (C.c_int64 * 3)(), (C.c_float * 6)()
        self.ctx._check(self.ctx._lib.rtb_scene_kd_download(self.ctx._h, self._h, None, 0, None, 0, counts, box), "rtb_scene_kd_download")
        nodes, refs = np.zeros((counts[0], 2), np.uint32), np.zeros(counts[1], np.uint32)
        self.ctx._check(self.ctx._lib.rtb_scene_kd_download(self.ctx._h, self._h, nodes.ctypes.data, counts[0], refs.ctypes.data, counts[1],
                                                            counts, box), "rtb_scene_kd_download")
        return nodes, refs, int(counts[2]), np.array(list(box), np.float32)

    def render(self, camera, setting, frame, out=None):
        """rtb_render with HOST buffers; returns (rows x width x 3 float32 | reference-order array, stats)."""
        rows = shard_rows(frame)
        if rows < 0:
            raise RtbError("bad frame: size / rank / world / row_block (must be a multiple of 8)")
        if out is None:
            shape = (frame.width, frame.height, 3) if frame.layout & LAYOUT_REFERENCE else (rows, shard_width(frame), 3)
            if frame.layout & OUTPUT_MOMENTS:
                shape = shape[:2] + (6,)
            out = np.zeros(shape, np.uint8 if frame.layout & OUTPUT_RGB8 else np.float32)
        st = Stats()
        self.ctx._check(self.ctx._lib.rtb_render(self.ctx._h, self._h, C.byref(camera), C.byref(setting), C.byref(frame),
                                                 out.ctypes.data, C.byref(st)), "rtb_render")
        return out, st.as_dict()

    def render_device(self, camera, setting, frame, device_ptr, stream=0, want_stats=False):
        """rtb_render_device into DEVICE memory (e.g. a torch tensor's data_ptr()) on `stream`."""
        st = Stats()
        self.ctx._check(self.ctx._lib.rtb_render_device(self.ctx._h, self._h, C.byref(camera), C.byref(setting),
                                                        C.byref(frame), C.c_void_p(device_ptr), C.c_void_p(stream),
                                                        C.byref(st) if want_stats else None), "rtb_render_device")
        return st.as_dict() if want_stats else None

    def trace_primary(self, camera, width, height, seq=True, seq_cap=0):
        n = width * height
        hid, ht = np.zeros(n, np.int32), np.zeros(n, np.float32)
        slen = np.zeros(n, np.int32) if seq else None
        shash = np.zeros(n, np.uint64) if seq else None
        sbuf = np.zeros((n, seq_cap), np.int32) if seq_cap else None
        p = lambda a: a.ctypes.data if a is not None else None
        self.ctx._check(self.ctx._lib.rtb_trace_primary(self.ctx._h, self._h, C.byref(camera), width, height, p(hid), p(ht),
                                                        p(slen), p(shash), p(sbuf), seq_cap), "rtb_trace_primary")
        return {"hit_id": hid, "hit_t": ht, "seq_len": slen, "seq_hash": shash, "seq_buf": sbuf}

    def intersect_rays(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        hid, ht = np.zeros(n, np.int32), np.zeros(n, np.float32)
        pos, nrm = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        self.ctx._check(self.ctx._lib.rtb_intersect_rays(self.ctx._h, self._h, n, rays.ctypes.data, hid.ctypes.data,
                                                         ht.ctypes.data, pos.ctypes.data, nrm.ctypes.data),
                        "rtb_intersect_rays")
        return hid, ht, pos, nrm
