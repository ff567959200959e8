"""Host-side mirror of the reference's interface for the ray-casting-correspondence path, over the C ABI (ctypes).

Class / method names follow the reference (file:line cited per method); all arithmetic happens in librmcl_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .synth import CROSS_STATS_DTYPE, PARTICLE_ATTR_DTYPE, RANGE_MEAS_DTYPE, TRANSFORM_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("B2_LIB_PATH") or os.path.join(_HERE, "lib", "librmcl_b200.so")     # B2_LIB_PATH: experiment builds (scripts/)
_lib = None

B2_BUILD_HOST_SAH = 0
B2_BUILD_DEVICE_LBVH = 1


class B2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rmcl_b200 error {code}: {msg}")
        self.code = code


def lib_path():
    return _LIB_PATH


class _SphericalModel(C.Structure):
    _fields_ = [("phi_min", C.c_float), ("phi_inc", C.c_float), ("phi_size", C.c_uint32), ("theta_min", C.c_float), ("theta_inc", C.c_float),
                ("theta_size", C.c_uint32), ("range_min", C.c_float), ("range_max", C.c_float)]


class _PinholeModel(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("range_min", C.c_float), ("range_max", C.c_float)]


class PFParams(C.Structure):
    """PCDSensorUpdaterEmbree config (rmcl_ros/src/rmcl/PCDSensorUpdaterEmbree.cpp:122-134)."""
    _fields_ = [("dist_sigma", C.c_float), ("real_hit_sim_miss_error", C.c_float), ("real_miss_sim_hit_error", C.c_float),
                ("real_miss_sim_miss_error", C.c_float), ("range_min", C.c_float), ("range_max", C.c_float), ("ng_mode", C.c_int),
                ("correspondence_type", C.c_int)]       # 0: ray casting (evaluate_rcc), 1: closest point (evaluate_cpc, :88-95,219-222)

    @staticmethod
    def defaults(ng_mode=0, correspondence_type=0):
        return PFParams(2.0, 100.0, 100.0, 0.0, 0.05, 80.0, ng_mode, correspondence_type)


class GladiatorConfig(C.Structure):
    """GladiatorResamplerConfig (rmcl_ros/include/rmcl_ros/rmcl/GladiatorResamplerConfig.hpp:7-20)."""
    _fields_ = [("min_noise_tx", C.c_float), ("min_noise_ty", C.c_float), ("min_noise_tz", C.c_float), ("min_noise_roll", C.c_float),
                ("min_noise_pitch", C.c_float), ("min_noise_yaw", C.c_float), ("likelihood_forget_per_meter", C.c_float),
                ("likelihood_forget_per_radian", C.c_float)]

    @staticmethod
    def defaults():
        return GladiatorConfig(0.03, 0.03, 0.0, 0.0, 0.0, 0.01, 0.3, 0.2)


class _MeshInfo(C.Structure):
    _fields_ = [("n_faces", C.c_uint32), ("n_vertices", C.c_uint32), ("n_nodes", C.c_uint32), ("n_leaf_tris", C.c_uint32), ("max_depth", C.c_uint32),
                ("bvh_bytes", C.c_uint64), ("build_ms", C.c_float), ("device", C.c_int), ("build_mode", C.c_int), ("sah_cost", C.c_float)]


EXPORTS = [
    "b2_last_error", "b2_version", "b2_device_count", "b2_mesh_create", "b2_mesh_destroy", "b2_mesh_get_info", "b2_mesh_intersect",
    "b2_mesh_intersect_stats", "b2_rcc_create", "b2_rcc_destroy", "b2_rcc_set_stream", "b2_rcc_set_tsb", "b2_rcc_set_model_spherical",
    "b2_rcc_set_model_pinhole", "b2_rcc_set_model_o1dn", "b2_rcc_set_model_ondn", "b2_rcc_set_params", "b2_rcc_set_dataset", "b2_rcc_set_ranges",
    "b2_rcc_find", "b2_rcc_cross_statistics", "b2_rcc_model_view", "b2_rcc_dataset_view", "b2_rcc_download_model", "b2_rcc_download_dataset",
    "b2_rcc_correct_once", "b2_rcc_correct_once_ranges", "b2_rcc_correct_batch", "b2_umeyama_batch", "b2_pf_create", "b2_pf_destroy",
    "b2_pf_set_stream", "b2_pf_sensor_update", "b2_pf_sensor_update_host", "b2_kernel_launch_count", "b2_rcc_enable_timing", "b2_rcc_last_timing", "b2_pf_motion_update", "b2_pf_likelihood_stats",
    "b2_rcc_set_correspondence_type", "b2_pf_resample_gladiator", "b2_pf_gladiator_randoms", "b2_rcc_segment", "b2_mesh_create_from_file", "b2_mesh_file_load", "b2_mesh_file_free", "b2_peek_cuda_error", "b2_mesh_blob_size", "b2_mesh_export_blob", "b2_mesh_create_from_blob", "b2_mesh_refit",
    "b2_rcc_correct_once_async", "b2_rcc_correct_once_wait", "b2_micp_correct_once", "b2_rcc_set_exec_mode", "b2_debug_read_bandwidth", "b2_rcc_set_sim_options", "b2_rcc_bind_dataset", "b2_rcc_bind_model_buffers", "b2_rcc_benchmark_batch", "b2_pf_p2p_init", "b2_pf_p2p_connect", "b2_pf_p2p_publish", "b2_pf_resample_gladiator_p2p", "b2_pf_p2p_connect_local", "b2_rcc_set_cpc_options", "b2_pf_set_mapping", "b2_pf_get_mapping",
]


def load_library():
    """Load librmcl_b200.so.  Fails loudly when it is missing: there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise B2Error(-2, f"CUDA library not built: {_LIB_PATH} (run `python -c 'import __graft_entry__ as g; g.build()'`)")
    lib = C.CDLL(_LIB_PATH)
    for name in EXPORTS:
        getattr(lib, name)            # AttributeError if the ABI is incomplete
    lib.b2_last_error.restype = C.c_char_p
    lib.b2_kernel_launch_count.restype = C.c_uint64
    for name in EXPORTS:
        if name not in ("b2_last_error", "b2_kernel_launch_count", "b2_mesh_file_free", "b2_peek_cuda_error"):
            getattr(lib, name).restype = C.c_int
    lib.b2_mesh_file_free.restype = None
    lib.b2_peek_cuda_error.restype = C.c_char_p
    _lib = lib
    return lib


def _chk(rc):
    if rc != 0:
        raise B2Error(rc, load_library().b2_last_error().decode("utf-8", "replace"))


def read_bandwidth(nbytes, iters=20, device=0):
    """GB/s of the library's read micro-benchmark over a working set of `nbytes` (L2 below its capacity, HBM far above)."""
    out = C.c_double()
    _chk(load_library().b2_debug_read_bandwidth(C.c_int(device), C.c_uint64(int(nbytes)), C.c_int(int(iters)), C.byref(out)))
    return out.value


def kernel_launch_count():
    return int(load_library().b2_kernel_launch_count())


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _tf(a):
    a = np.ascontiguousarray(a)
    if a.dtype.itemsize != 32:
        raise TypeError("expected 32-byte Transform records (synth.TRANSFORM_DTYPE)")
    return a


def _devptr(x):
    """Device address of a torch CUDA tensor (or a raw int)."""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):
        if not x.is_cuda or not x.is_contiguous():
            raise TypeError("expected a contiguous CUDA tensor")
        return C.c_void_p(x.data_ptr())
    raise TypeError(type(x))


def read_mesh_file(path):
    """Import step of Map.from_file alone (host only): -> (vertices (nv,3) float32, faces (nf,3) uint32)."""
    lib = load_library()
    v, f, nv, nf = C.POINTER(C.c_float)(), C.POINTER(C.c_uint32)(), C.c_uint32(), C.c_uint32()
    _chk(lib.b2_mesh_file_load(os.fsencode(path), C.byref(v), C.byref(nv), C.byref(f), C.byref(nf)))
    try:
        V = np.ctypeslib.as_array(v, (nv.value, 3)).copy()
        F = np.ctypeslib.as_array(f, (nf.value, 3)).copy()
    finally:
        lib.b2_mesh_file_free(v, f)
    return V, F


class Map:
    """Triangle mesh + in-HBM BVH; stands in for rm::EmbreeMap / rm::OptixMap (rm::import_embree_map, micp_localization.cpp:188)."""

    def __init__(self, verts, faces, device=0, build_mode=None):
        lib = load_library()
        if build_mode is None:
            build_mode = int(os.environ.get("B2_BUILD_MODE", B2_BUILD_DEVICE_LBVH))     # device build: faster to build AND to trace
        verts = _f32(verts).reshape(-1, 3)
        faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
        h = C.c_void_p()
        _chk(lib.b2_mesh_create(_p(verts), C.c_uint32(len(verts)), _p(faces), C.c_uint32(len(faces)), C.c_int(device), C.c_int(build_mode), C.byref(h)))
        self._h = h
        self.device = device

    @classmethod
    def from_file(cls, path, device=0, build_mode=None):
        """rm::import_embree_map(file) twin: .ply (ascii / binary little endian) or .obj."""
        lib = load_library()
        if build_mode is None:
            build_mode = int(os.environ.get("B2_BUILD_MODE", B2_BUILD_DEVICE_LBVH))
        self = cls.__new__(cls)
        h = C.c_void_p()
        _chk(lib.b2_mesh_create_from_file(os.fsencode(path), C.c_int(device), C.c_int(build_mode), C.byref(h)))
        self._h, self.device = h, device
        return self

    def refit(self, verts):
        """Vertices moved, same faces: refit the resident BVH.  Device-built maps only."""
        verts = _f32(verts).reshape(-1, 3)
        _chk(load_library().b2_mesh_refit(self._h, _p(verts), C.c_uint32(len(verts)), C.c_int(0)))

    def export_blob(self):
        """The built BVH as bytes (numpy uint8): build once, broadcast / cache, `Map.from_blob` on the receiving side."""
        n = C.c_uint64()
        _chk(load_library().b2_mesh_blob_size(self._h, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        _chk(load_library().b2_mesh_export_blob(self._h, _p(buf), C.c_uint64(n.value)))
        return buf

    @classmethod
    def from_blob(cls, blob, device=0):
        blob = np.ascontiguousarray(blob, np.uint8)
        self = cls.__new__(cls)
        h = C.c_void_p()
        _chk(load_library().b2_mesh_create_from_blob(_p(blob), C.c_uint64(blob.size), C.c_int(device), C.byref(h)))
        self._h, self.device = h, device
        return self

    def close(self):
        if getattr(self, "_h", None):
            load_library().b2_mesh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        inf = _MeshInfo()
        _chk(load_library().b2_mesh_get_info(self._h, C.byref(inf)))
        return {k: getattr(inf, k) for k, _ in _MeshInfo._fields_}

    def intersect(self, origs, dirs, tfar=np.inf):
        """Closest hit for arbitrary rays (replaces rtcIntersect1, PCDSensorUpdaterEmbree.cpp:30-47)."""
        origs, dirs = _f32(origs).reshape(-1, 3), _f32(dirs).reshape(-1, 3)
        n = len(origs)
        t, f, ng, h = np.empty(n, np.float32), np.empty(n, np.uint32), np.empty((n, 3), np.float32), np.empty(n, np.uint8)
        _chk(load_library().b2_mesh_intersect(self._h, _p(origs), _p(dirs), C.c_uint32(n), C.c_float(tfar), _p(t), _p(f), _p(ng), _p(h)))
        return t, f, ng, h

    def traversal_stats(self, origs, dirs, tfar=np.inf):
        origs, dirs = _f32(origs).reshape(-1, 3), _f32(dirs).reshape(-1, 3)
        a, b = C.c_double(), C.c_double()
        _chk(load_library().b2_mesh_intersect_stats(self._h, _p(origs), _p(dirs), C.c_uint32(len(origs)), C.c_float(tfar), C.byref(a), C.byref(b)))
        return a.value, b.value


class RCCB200:
    """Ray-casting correspondences on the B200: the Correspondences_<VRAM_CUDA> interface
    (rmcl/include/rmcl/registration/Correspondences.hpp:16-88) + the fused drivers."""

    def __init__(self, map_: Map):
        self.map = map_
        h = C.c_void_p()
        _chk(load_library().b2_rcc_create(map_._h, C.byref(h)))
        self._h = h
        self.n = 0
        self.outdated = True                       # Correspondences.hpp:31
        self.max_dist = 1.0                        # params.max_dist, :22
        self.adaptive_max_dist_min = 0.15          # :23

    def close(self):
        if getattr(self, "_h", None):
            load_library().b2_rcc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setStream(self, cuda_stream):
        _chk(load_library().b2_rcc_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def enableTiming(self, on=True):
        _chk(load_library().b2_rcc_enable_timing(self._h, C.c_int(int(on))))

    def lastTiming(self):
        """(find_ms, reduce_ms) of the most recent correctOnce, from CUDA events recorded on the handle's stream."""
        a, b = C.c_float(), C.c_float()
        _chk(load_library().b2_rcc_last_timing(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def setTsb(self, Tsb):                         # Correspondences.hpp:33-36
        _chk(load_library().b2_rcc_set_tsb(self._h, _p(_tf(Tsb))))

    def setParams(self, max_dist, adaptive_max_dist_min):
        self.max_dist, self.adaptive_max_dist_min = float(max_dist), float(adaptive_max_dist_min)
        _chk(load_library().b2_rcc_set_params(self._h, C.c_float(max_dist), C.c_float(adaptive_max_dist_min)))

    def setModel(self, m):                         # ModelSetter<ModelT>::setModel (RCCEmbree.cpp:21-24 ...)
        lib = load_library()
        name = type(m).__name__
        if name == "SphericalModel":
            sm = _SphericalModel(m.phi_min, m.phi_inc, m.phi_size, m.theta_min, m.theta_inc, m.theta_size, m.range_min, m.range_max)
            _chk(lib.b2_rcc_set_model_spherical(self._h, C.byref(sm)))
        elif name == "PinholeModel":
            pm = _PinholeModel(m.width, m.height, m.fx, m.fy, m.cx, m.cy, m.range_min, m.range_max)
            _chk(lib.b2_rcc_set_model_pinhole(self._h, C.byref(pm)))
        elif name == "O1DnModel":
            o, d = _f32(m.orig).reshape(3), _f32(m.dirs).reshape(-1, 3)
            _chk(lib.b2_rcc_set_model_o1dn(self._h, C.c_uint32(m.width), C.c_uint32(m.height), _p(o), _p(d), C.c_float(m.range_min), C.c_float(m.range_max)))
        elif name == "OnDnModel":
            o, d = _f32(m.origs).reshape(-1, 3), _f32(m.dirs).reshape(-1, 3)
            _chk(lib.b2_rcc_set_model_ondn(self._h, C.c_uint32(m.width), C.c_uint32(m.height), _p(o), _p(d), C.c_float(m.range_min), C.c_float(m.range_max)))
        else:
            raise TypeError(name)
        self.model = m
        self.n = m.size

    def setDataset(self, points, mask=None):       # public field `dataset`, Correspondences.hpp:24
        if hasattr(points, "data_ptr"):
            n = points.numel() // 3
            _chk(load_library().b2_rcc_set_dataset(self._h, _devptr(points), _devptr(mask), C.c_uint32(n), C.c_int(1)))
        else:
            points = _f32(points).reshape(-1, 3)
            mask = None if mask is None else np.ascontiguousarray(mask, np.uint8)
            _chk(load_library().b2_rcc_set_dataset(self._h, _p(points), _p(mask), C.c_uint32(len(points)), C.c_int(0)))
        self.outdated = True

    def setRanges(self, ranges):                   # MICP..Sensor..::unpackMessage / v1 setInputData
        if hasattr(ranges, "data_ptr"):
            _chk(load_library().b2_rcc_set_ranges(self._h, _devptr(ranges), C.c_uint32(ranges.numel()), C.c_int(1)))
        else:
            ranges = _f32(ranges).reshape(-1)
            _chk(load_library().b2_rcc_set_ranges(self._h, _p(ranges), C.c_uint32(len(ranges)), C.c_int(0)))
        self.outdated = True

    setInputData = setRanges

    def find(self, Tbm_est):                       # RCCEmbree.cpp:26-36
        _chk(load_library().b2_rcc_find(self._h, _p(_tf(Tbm_est))))
        self.outdated = False

    def computeCrossStatistics(self, T_snew_sold, convergence_progress=0.0):   # CorrespondencesCPU.cpp:10-39
        out = np.zeros((), CROSS_STATS_DTYPE)
        _chk(load_library().b2_rcc_cross_statistics(self._h, _p(_tf(T_snew_sold)), C.c_double(convergence_progress), _p(out)))
        return out

    def modelView(self):                           # Correspondences.hpp:47-54 (host copies)
        nn = C.c_uint32()
        _chk(load_library().b2_rcc_model_view(self._h, None, None, None, None, None, C.byref(nn)))
        n = nn.value
        out = dict(points=np.empty((n, 3), np.float32), normals=np.empty((n, 3), np.float32), hits=np.empty(n, np.uint8),
                   face_ids=np.empty(n, np.uint32), ranges=np.empty(n, np.float32))
        _chk(load_library().b2_rcc_download_model(self._h, _p(out["points"]), _p(out["normals"]), _p(out["hits"]), _p(out["face_ids"]), _p(out["ranges"])))
        return out

    def datasetView(self):                         # Correspondences.hpp:56-62 (host copies)
        p, m, n = C.c_void_p(), C.c_void_p(), C.c_uint32()
        _chk(load_library().b2_rcc_dataset_view(self._h, C.byref(p), C.byref(m), C.byref(n)))
        out = dict(points=np.empty((n.value, 3), np.float32), mask=np.empty(n.value, np.uint8))
        _chk(load_library().b2_rcc_download_dataset(self._h, _p(out["points"]), _p(out["mask"])))
        return out

    def segment(self, min_dist_outlier_scan=0.15, min_dist_outlier_map=0.15):
        """ScanMapSegmentationEmbreeNode::scanCB classification (scan_map_segmentation_embree.cpp:110-187) after setRanges(real scan) and
        find(pose): -> (outlier_scan (k,3), outlier_map (m,3), labels (n,))."""
        nn = C.c_uint32()
        _chk(load_library().b2_rcc_model_view(self._h, None, None, None, None, None, C.byref(nn)))
        n = nn.value
        a, b, lab = np.zeros((max(n, 1), 3), np.float32), np.zeros((max(n, 1), 3), np.float32), np.zeros(max(n, 1), np.uint8)
        na, nb = C.c_uint32(), C.c_uint32()
        _chk(load_library().b2_rcc_segment(self._h, C.c_float(min_dist_outlier_scan), C.c_float(min_dist_outlier_map), _p(a), C.c_uint32(n), C.byref(na),
                                           _p(b), C.c_uint32(n), C.byref(nb), _p(lab)))
        return a[: na.value].copy(), b[: nb.value].copy(), lab[:n].copy()

    _CO_OUT = np.dtype([("Tn", TRANSFORM_DTYPE), ("Td", TRANSFORM_DTYPE), ("Cm", CROSS_STATS_DTYPE)])

    def _co_scratch(self):
        """Persistent argument block of correctOnce: at ~10 kHz call rates the per-call numpy/ctypes object churn of the generic path
        (~12 us) would be a tenth of the step; here the arguments go into one preallocated 192-byte block and the addresses are plain ints."""
        sc = getattr(self, "_co", None)
        if sc is None:
            buf = (C.c_char * 192)()                  # Tom 0:32 | Tbo 32:64 | Tom_new 64:96 | T_onew_oold 96:128 | Cmerged_o 128:192
            base = C.addressof(buf)
            lib = load_library()
            lib.b2_rcc_correct_once.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
            lib.b2_rcc_correct_once_ranges.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
            lib.b2_rcc_correct_once_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double]
            lib.b2_rcc_correct_once_wait.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            self._co_async, self._co_wait = lib.b2_rcc_correct_once_async, lib.b2_rcc_correct_once_wait
            sc = self._co = (buf, memoryview(buf).cast("B"), base, lib.b2_rcc_correct_once, lib.b2_rcc_correct_once_ranges)
        return sc

    def correctOnce(self, Tom, Tbo, iterations=5, convergence_progress=0.0, ranges=None):
        """MICPLocalizationNode::correctOnce for this sensor (micp_localization.cpp:899-984), fully on the device.
        With `ranges` (host array) the scan upload is part of the call (end-to-end entry point)."""
        buf, mv, base, f_once, f_ranges = self._co_scratch()
        a, b = np.asarray(Tom).tobytes(), np.asarray(Tbo).tobytes()
        if len(a) != 32 or len(b) != 32:
            raise TypeError("expected 32-byte Transform records (synth.TRANSFORM_DTYPE)")
        mv[0:32] = a
        mv[32:64] = b
        if ranges is None:
            rc = f_once(self._h, base, base + 32, iterations, convergence_progress, base + 64, base + 96, base + 128)
        else:
            if hasattr(ranges, "data_ptr"):       # pinned / pageable HOST torch tensor
                rp, rn = ranges.data_ptr(), ranges.numel()
            else:
                ranges = _f32(ranges).reshape(-1)
                rp, rn = ranges.ctypes.data, len(ranges)
            rc = f_ranges(self._h, rp, rn, base, base + 32, iterations, convergence_progress, base + 64, base + 96, base + 128)
        if rc != 0:
            _chk(rc)
        self.outdated = False
        out = np.frombuffer(bytearray(mv[64:192]), self._CO_OUT)
        return out["Tn"][0], out["Td"][0], out["Cm"][0]

    def bindDataset(self, points, mask):
        """use caller-owned CUDA tensors as the dataset (no copy): the reference's public `dataset` member in VRAM (Correspondences.hpp:24)"""
        n = points.numel() // 3
        _chk(load_library().b2_rcc_bind_dataset(self._h, _devptr(points), _devptr(mask), C.c_uint32(n)))
        self._bound = (points, mask)
        self.outdated = True

    def bindModelBuffers(self, points, normals, hits):
        """find() writes into these caller-owned CUDA tensors: the reference's protected `model_buffers_` (Correspondences.hpp:81-85); None unbinds"""
        if points is None:
            _chk(load_library().b2_rcc_bind_model_buffers(self._h, None, None, None, C.c_uint32(0)))
            self._bound_model = None
            return
        _chk(load_library().b2_rcc_bind_model_buffers(self._h, _devptr(points), _devptr(normals), _devptr(hits), C.c_uint32(hits.numel())))
        self._bound_model = (points, normals, hits)

    def setSimOptions(self, tfar_mode=0, min_mode=0, miss_fill=0):
        """the open rmagine simulate() semantics of SURVEY.md A.3: tfar = +inf, closest hit below range.min = miss, misses filled with zeros"""
        _chk(load_library().b2_rcc_set_sim_options(self._h, C.c_int(tfar_mode), C.c_int(min_mode), C.c_int(miss_fill)))

    def setExecMode(self, mode):
        """2 (default): one kernel, block sums exchanged by 64-bit atomics, programmatic launch; 1: cooperative launch, FP64 exchange behind a grid sync; 0: one reduction launch per inner iteration."""
        _chk(load_library().b2_rcc_set_exec_mode(self._h, C.c_int(int(mode))))

    def correctOnceAsync(self, Tom, Tbo, iterations=5, convergence_progress=0.0):
        """Enqueue one correctOnce on the handle's stream and return; collect with correctOnceWait()."""
        buf, mv, base, _, _ = self._co_scratch()
        mv[0:32] = np.asarray(Tom).tobytes()
        mv[32:64] = np.asarray(Tbo).tobytes()
        rc = self._co_async(self._h, base, base + 32, iterations, convergence_progress)
        if rc != 0:
            _chk(rc)

    def correctOnceWait(self):
        buf, mv, base, _, _ = self._co_scratch()
        rc = self._co_wait(self._h, base + 64, base + 96, base + 128)
        if rc != 0:
            _chk(rc)
        self.outdated = False
        out = np.frombuffer(bytearray(mv[64:192]), self._CO_OUT)
        return out["Tn"][0], out["Td"][0], out["Cm"][0]

    def benchmark(self, Tbm, n_runs=10):
        """v1 corrector.benchmark(T, Nruns) -> dict(sim, red, svd) seconds (lidar_corrector_optix_benchmark.cpp:143-155): stages run unfused"""
        Tbm = _tf(Tbm).reshape(-1)
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _chk(load_library().b2_rcc_benchmark_batch(self._h, _p(Tbm), C.c_uint32(len(Tbm)), C.c_uint32(n_runs), C.byref(a), C.byref(b), C.byref(c)))
        return dict(sim=a.value, red=b.value, svd=c.value)

    def correct(self, Tbm):
        """v1 {Sphere,Pinhole,O1Dn}Corrector::correct(Tbm[N]) -> (Tdelta[N], Ncorr[N], stats_b[N])
        (rmcl_ros/src/benchmarks/lidar_corrector_embree_benchmark.cpp:127-133)."""
        lib = load_library()
        if hasattr(Tbm, "data_ptr"):
            import torch
            n = Tbm.numel() * Tbm.element_size() // 32
            Td = torch.empty((n, 8), dtype=torch.float32, device=Tbm.device)
            nc = torch.empty((n,), dtype=torch.int32, device=Tbm.device)
            st = torch.empty((n, 16), dtype=torch.float32, device=Tbm.device)
            _chk(lib.b2_rcc_correct_batch(self._h, _devptr(Tbm), C.c_uint32(n), C.c_int(1), _devptr(Td), _devptr(nc), _devptr(st), C.c_int(1)))
            return Td, nc, st
        Tbm = _tf(Tbm).reshape(-1)
        n = len(Tbm)
        Td, nc, st = np.zeros(n, TRANSFORM_DTYPE), np.zeros(n, np.uint32), np.zeros(n, CROSS_STATS_DTYPE)
        _chk(lib.b2_rcc_correct_batch(self._h, _p(Tbm), C.c_uint32(n), C.c_int(0), _p(Td), _p(nc), _p(st), C.c_int(0)))
        return Td, nc, st


class RCCB200Spherical(RCCB200):
    """rmcl::RCCEmbreeSpherical / RCCOptixSpherical twin (rmcl/include/rmcl/registration/RCCEmbree.hpp:18-33)."""


class RCCB200Pinhole(RCCB200):
    """rmcl::RCCEmbreePinhole twin (RCCEmbree.hpp:35-49)."""


class RCCB200O1Dn(RCCB200):
    """rmcl::RCCEmbreeO1Dn twin (RCCEmbree.hpp:51-66)."""


class RCCB200OnDn(RCCB200):
    """rmcl::RCCEmbreeOnDn twin (RCCEmbree.hpp:68-83)."""


class CPCB200(RCCB200):
    """rmcl::CPCEmbree twin (rmcl/include/rmcl/registration/CPCEmbree.hpp:20-54): closest-point correspondences.  find() runs one
    closest-point query per dataset point on the same map BVH (CPCEmbree.cpp:17-43); no sensor model.  computeCrossStatistics /
    correctOnce are the shared Correspondences_ code."""

    def __init__(self, map_: Map):
        super().__init__(map_)
        _chk(load_library().b2_rcc_set_correspondence_type(self._h, C.c_int(1)))

    def setModel(self, m):
        raise B2Error(-1, "CPCB200 has no sensor model (closest-point correspondences use the dataset points)")

    def setOptions(self, skip_masked=False):
        """skip_masked: do not query dataset points whose mask is 0 (the reference queries them and never uses the result)"""
        _chk(load_library().b2_rcc_set_cpc_options(self._h, C.c_int(int(skip_masked))))


# v1 names used by the legacy benchmarks (lidar_corrector_{embree,optix}_benchmark.cpp:86)
SphereCorrectorB200 = RCCB200Spherical
PinholeCorrectorB200 = RCCB200Pinhole
O1DnCorrectorB200 = RCCB200O1Dn
OnDnCorrectorB200 = RCCB200OnDn


def micp_correct_once(sensors, Tbo, Tom, iterations=5, convergence_progress=0.0, merge_weights=None, ranges=None):
    """MICPLocalizationNode::correctOnce over all sensors (micp_localization.cpp:899-984): `sensors` RCCB200 / CPCB200 handles on one device,
    `Tbo` one Transform per sensor, `merge_weights` the sensors' merge_weight_multiplier (MICPSensor.hpp:103), `ranges` optional per-sensor
    host scans (None entries keep the resident dataset).  -> (Tom_new, T_onew_oold, Cmerged_o)."""
    n = len(sensors)
    hs = (C.c_void_p * n)(*[s._h for s in sensors])
    Tb = np.ascontiguousarray(np.asarray(Tbo, TRANSFORM_DTYPE).reshape(n))
    w = None if merge_weights is None else np.ascontiguousarray(merge_weights, np.float64).reshape(n)
    keep, rp = [], None
    if ranges is not None:
        rp = (C.c_void_p * n)()
        for k, r in enumerate(ranges):
            if r is None:
                rp[k] = None
            elif hasattr(r, "data_ptr"):
                rp[k] = r.data_ptr()
            else:
                r = _f32(r).reshape(-1)
                keep.append(r)
                rp[k] = r.ctypes.data
    Tn, Td, Cm = np.zeros((), TRANSFORM_DTYPE), np.zeros((), TRANSFORM_DTYPE), np.zeros((), CROSS_STATS_DTYPE)
    _chk(load_library().b2_micp_correct_once(hs, _p(Tb), _p(w), rp, C.c_uint32(n), _p(_tf(Tom)), C.c_uint32(iterations), C.c_double(convergence_progress),
                                             _p(Tn), _p(Td), _p(Cm)))
    for s in sensors:
        s.outdated = False
    return Tn, Td, Cm


def umeyama_transform(stats, device=0):
    """rm::umeyama_transform for an array of CrossStatistics (micp_localization.cpp:952-953)."""
    stats = np.ascontiguousarray(stats).reshape(-1)
    out = np.zeros(len(stats), TRANSFORM_DTYPE)
    _chk(load_library().b2_umeyama_batch(_p(stats), C.c_uint32(len(stats)), _p(out), C.c_int(0), C.c_int(device), None))
    return out


class PCDSensorUpdaterB200:
    """Particle-filter sensor update: ParticleUpdater<MemT>::update (rmcl_ros/include/rmcl_ros/rmcl/ParticleUpdater.hpp:39-43) with the
    hot loop of PCDSensorUpdaterEmbree::update (PCDSensorUpdaterEmbree.cpp:290-342).  Beams are an input (quirk D5)."""

    def __init__(self, map_: Map):
        self.map = map_
        h = C.c_void_p()
        _chk(load_library().b2_pf_create(map_._h, C.byref(h)))
        self._h = h
        self.config = PFParams.defaults()

    def close(self):
        if getattr(self, "_h", None):
            load_library().b2_pf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setStream(self, cuda_stream):
        _chk(load_library().b2_pf_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def setMapping(self, mode):
        """0 lanes = beams of one particle, 1 lanes = particles, 2 lanes = particles sorted by pose, 3 (default) auto by timing; results identical."""
        _chk(load_library().b2_pf_set_mapping(self._h, C.c_int(mode)))

    def mapping(self):
        mode, cur = C.c_int(0), C.c_int(0)
        _chk(load_library().b2_pf_get_mapping(self._h, C.byref(mode), C.byref(cur)))
        return mode.value, cur.value

    def update(self, particle_poses, particle_attrs, Tsb, beams, params: PFParams | None = None, inplace=False):
        """RAM variant: numpy arrays in, updated attrs array out (inplace=True: the caller's attrs array is updated itself, as the reference's
        update(MemoryView) does -- no staging copy; pin the arrays and the transfers overlap the kernel).  VRAM variant: torch CUDA tensors, attrs
        updated in place."""
        prm = params or self.config
        beams = np.ascontiguousarray(beams)
        assert beams.dtype.itemsize == 64
        Tsb = _tf(Tsb)
        lib = load_library()
        if hasattr(particle_poses, "data_ptr"):
            n = particle_poses.numel() * particle_poses.element_size() // 32
            _chk(lib.b2_pf_sensor_update(self._h, _devptr(particle_poses), _devptr(particle_attrs), C.c_uint32(n), _p(Tsb), _p(beams),
                                         C.c_uint32(len(beams)), C.byref(prm)))
            return particle_attrs
        poses = _tf(particle_poses).reshape(-1)
        if inplace:
            attrs = particle_attrs
            assert isinstance(attrs, np.ndarray) and attrs.flags.c_contiguous and attrs.flags.writeable
        else:
            attrs = np.ascontiguousarray(particle_attrs).copy()
        assert attrs.dtype.itemsize == 36
        _chk(lib.b2_pf_sensor_update_host(self._h, _p(poses), _p(attrs), C.c_uint32(len(poses)), _p(Tsb), _p(beams), C.c_uint32(len(beams)), C.byref(prm)))
        return attrs

    def motionUpdate(self, particle_poses, particle_attrs, T_bnew_bold, forget_rate, check_collision=False):
        """TFMotionUpdaterGPU (rmcl_ros/src/rmcl/particle_motion.cu:11-46): in place on torch CUDA tensors.  check_collision adds the CPU
        updater's wall check (TFMotionUpdaterCPU.cpp:17-50,205-216)."""
        n = particle_poses.numel() * particle_poses.element_size() // 32
        _chk(load_library().b2_pf_motion_update(self._h, _devptr(particle_poses), _devptr(particle_attrs), C.c_uint32(n), _p(_tf(T_bnew_bold)), C.c_double(forget_rate),
                                                C.c_int(int(check_collision))))

    def likelihoodStats(self, particle_attrs, dist=None):
        """compute_stats (rmcl_ros/src/rmcl/resampling.cu:41-92) of the local particles; with `dist` (torch.distributed, particles sharded
        across ranks) the 8 bytes are all-reduced (SUM / MAX) -- the only exchange step of the particle-filter cycle."""
        n = particle_attrs.numel() * particle_attrs.element_size() // 36
        s, m = C.c_float(), C.c_float()
        _chk(load_library().b2_pf_likelihood_stats(self._h, _devptr(particle_attrs), C.c_uint32(n), C.byref(s), C.byref(m)))
        s, m = s.value, m.value
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            import torch
            dev = particle_attrs.device if dist.get_backend() == "nccl" else "cpu"
            ts = torch.tensor([s], dtype=torch.float64, device=dev)
            tm = torch.tensor([m], dtype=torch.float64, device=dev)
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            s, m = float(ts[0]), float(tm[0])
        return s, m


    def gladiatorRandoms(self, seed, step, first, n, device=None):
        """The draws resample() uses for champions first..first+n-1 (raw u32 opponent word + 6 normals each), as torch CUDA tensors."""
        import torch
        dev = device if device is not None else torch.device("cuda", self.map.device)
        raw = torch.empty(n, dtype=torch.int32, device=dev)
        nrm = torch.empty((n, 6), dtype=torch.float32, device=dev)
        _chk(load_library().b2_pf_gladiator_randoms(self._h, C.c_uint64(seed), C.c_uint32(step), C.c_uint32(first), C.c_uint32(n), _devptr(raw), _devptr(nrm)))
        return raw, nrm

    def resample(self, particle_poses, particle_attrs, particle_poses_new, particle_attrs_new, config=None, seed=1234, step=0, first=0, raw=None, normals=None):
        """GladiatorResamplerGPU::resample (rmcl_ros/src/rmcl/resampling.cu:108-221) on torch CUDA tensors.  `particle_poses/attrs` hold all
        particles opponents are drawn from; the outputs receive champions first .. first+len(out)-1 (single GPU: first = 0, same length)."""
        cfg = config or GladiatorConfig.defaults()
        n_all = particle_poses.numel() * particle_poses.element_size() // 32
        n_local = particle_poses_new.numel() * particle_poses_new.element_size() // 32
        _chk(load_library().b2_pf_resample_gladiator(self._h, _devptr(particle_poses), _devptr(particle_attrs), C.c_uint32(n_all), C.c_uint32(first), C.c_uint32(n_local),
                                                     _devptr(particle_poses_new), _devptr(particle_attrs_new), C.byref(cfg), C.c_uint64(seed), C.c_uint32(step),
                                                     _devptr(raw) if raw is not None else None, _devptr(normals) if normals is not None else None))
        return particle_poses_new, particle_attrs_new

    def p2pConnect(self, dist, n_per_rank):
        """Map the other ranks' particle exchange buffers over NVLink (CUDA IPC): once per particle-set size.  Raises B2Error if the GPUs have no
        peer access or CUDA IPC is unavailable (the caller then stays with resampleSharded's all-gather)."""
        import torch
        ws, rank = dist.get_world_size(), dist.get_rank()
        mine = np.zeros(128, np.uint8)
        _chk(load_library().b2_pf_p2p_init(self._h, C.c_uint32(n_per_rank), _p(mine)))
        dev = torch.device("cuda", self.map.device) if dist.get_backend() == "nccl" else "cpu"
        allh = [torch.zeros(128, dtype=torch.uint8, device=dev) for _ in range(ws)]
        dist.all_gather(allh, torch.from_numpy(mine).to(dev))
        table = np.ascontiguousarray(torch.stack(allh).cpu().numpy())
        _chk(load_library().b2_pf_p2p_connect(self._h, _p(table), C.c_uint32(ws), C.c_uint32(rank), C.c_uint32(n_per_rank)))
        self._p2p = (ws, rank, n_per_rank)

    def resampleShardedP2P(self, poses_local, attrs_local, dist, config=None, seed=1234, step=0, want_traffic=False):
        """Gladiator resampling over peer memory (b2_pf_resample_gladiator_p2p): publish the shard, barrier, one kernel that reads opponents'
        likelihoods (4 B) and winners' records (68 B) from the owners' HBM, barrier.  -> (poses_new, attrs_new[, remote bytes read])"""
        import torch
        cfg = config or GladiatorConfig.defaults()
        n = poses_local.shape[0]
        if getattr(self, "_p2p", None) is None or self._p2p[2] != n:
            raise B2Error(-1, "resampleShardedP2P: call p2pConnect(dist, n_per_rank) first")
        _chk(load_library().b2_pf_p2p_publish(self._h, _devptr(poses_local), _devptr(attrs_local), C.c_uint32(n)))
        dist.barrier()
        P_new, A_new = torch.empty_like(poses_local), torch.empty_like(attrs_local)
        t = C.c_uint64()
        _chk(load_library().b2_pf_resample_gladiator_p2p(self._h, _devptr(P_new), _devptr(A_new), C.byref(cfg), C.c_uint64(seed), C.c_uint32(step),
                                                         C.byref(t) if want_traffic else None))
        torch.cuda.synchronize(poses_local.device)
        dist.barrier()                      # nobody overwrites a published shard while a peer may still read it
        return (P_new, A_new, int(t.value)) if want_traffic else (P_new, A_new)

    def resampleP2PLocalWorld(self, poses_all, attrs_all, world, rank, config=None, seed=1234, step=0):
        """test hook: `world` equal shards on ONE device; this handle plays `rank`.  -> (poses_new, attrs_new, bytes read from "remote" shards)"""
        import torch
        cfg = config or GladiatorConfig.defaults()
        n_all = poses_all.shape[0]
        n = n_all // world
        _chk(load_library().b2_pf_p2p_connect_local(self._h, _devptr(poses_all), _devptr(attrs_all), C.c_uint32(world), C.c_uint32(rank), C.c_uint32(n)))
        P_new = torch.empty((n, 8), dtype=torch.float32, device=poses_all.device)
        A_new = torch.empty((n, 9), dtype=torch.float32, device=poses_all.device)
        t = C.c_uint64()
        _chk(load_library().b2_pf_resample_gladiator_p2p(self._h, _devptr(P_new), _devptr(A_new), C.byref(cfg), C.c_uint64(seed), C.c_uint32(step), C.byref(t)))
        return P_new, A_new, int(t.value)

    def resampleSharded(self, poses_local, attrs_local, dist, config=None, seed=1234, step=0):
        """Particles sharded over ranks (equal contiguous slices in rank order): all-gather the particle set (the exchange step of this
        stage: 68 B per particle over NVLink), then every rank resamples its own champions against GLOBAL opponents.  Draws are keyed by the
        global particle index, so the result equals the single-GPU resample of the concatenated set.  (n, 8) / (n, 9) float32 CUDA tensors."""
        import torch
        from .shard import gladiator_resample_sharded

        def run(P_all, A_all, first, n_local):
            P_new, A_new = torch.empty_like(poses_local), torch.empty_like(attrs_local)
            return self.resample(P_all, A_all, P_new, A_new, config, seed, step, first=first)

        return gladiator_resample_sharded(run, poses_local, attrs_local, dist, sync=lambda: torch.cuda.synchronize(poses_local.device))
