"""ctypes binding of the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY -- see oracle/oracle.h.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

TRANSFORM = np.dtype([("R", np.float32, 4), ("t", np.float32, 3), ("stamp", np.uint32)])
CROSS_STATS = np.dtype([("dataset_mean", np.float32, 3), ("model_mean", np.float32, 3), ("covariance", np.float32, 9), ("n_meas", np.uint32)])
GAUSSIAN1D = np.dtype([("mean", np.float32), ("sigma", np.float32), ("n_meas", np.uint32)])
PARTICLE_ATTR = np.dtype([("likelihood", GAUSSIAN1D), ("state_sigma", np.float32, 6)])
RANGE_MEAS = np.dtype([("orig", np.float32, 3), ("dir", np.float32, 3), ("range", np.float32), ("cov", np.float32, 9)])


class PFParams(C.Structure):
    _fields_ = [("dist_sigma", C.c_float), ("real_hit_sim_miss_error", C.c_float), ("real_miss_sim_hit_error", C.c_float),
                ("real_miss_sim_miss_error", C.c_float), ("range_min", C.c_float), ("range_max", C.c_float), ("ng_mode", C.c_int),
                ("correspondence_type", C.c_int)]

    @staticmethod
    def defaults(ng_mode=0, correspondence_type=0):
        # rmcl_ros/src/rmcl/PCDSensorUpdaterEmbree.cpp:122-134
        return PFParams(2.0, 100.0, 100.0, 0.0, 0.05, 80.0, ng_mode, correspondence_type)


class _Tf(C.Structure):
    _fields_ = [("v", C.c_float * 7), ("stamp", C.c_uint32)]


class _MicpSensor(C.Structure):
    """orc_micp_sensor (oracle.h)"""
    _fields_ = [("n", C.c_uint32), ("n_origs", C.c_uint32), ("origs_s", C.c_void_p), ("dirs_s", C.c_void_p), ("dataset_pts", C.c_void_p), ("dataset_mask", C.c_void_p),
                ("Tbo", _Tf), ("Tsb", _Tf), ("range_max", C.c_float), ("max_dist", C.c_float), ("adaptive_max_dist_min", C.c_float), ("pad_", C.c_float),
                ("merge_weight", C.c_double)]


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(os.path.join(_HERE, "oracle.c")), os.path.getmtime(os.path.join(_HERE, "oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_scene_create.restype = C.c_void_p
        _lib.orc_scene_create.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        _lib.orc_scene_destroy.argtypes = [C.c_void_p]
        _lib.orc_pf_evaluate_rcc.restype = C.c_float
        _lib.orc_adaptive_max_dist.restype = C.c_float
        _lib.orc_adaptive_max_dist.argtypes = [C.c_float, C.c_float, C.c_double]
        _lib.orc_intersect.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _tf(a):
    a = np.ascontiguousarray(a)
    assert a.dtype.itemsize == 32, a.dtype
    return a


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


class Scene:
    """Stands in for rm::EmbreeMap (closest hit over one triangle mesh)."""

    def __init__(self, verts, faces):
        self.verts = _f32(verts).reshape(-1, 3)
        self.faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
        self._h = C.c_void_p(lib().orc_scene_create(_p(self.verts), self.verts.shape[0], _p(self.faces), self.faces.shape[0]))

    def __del__(self):
        try:
            if self._h:
                lib().orc_scene_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- rays -------------------------------------------------------------------------------
    def intersect(self, origs, dirs, tfar=np.inf, brute=False):
        origs = _f32(origs).reshape(-1, 3)
        dirs = _f32(dirs).reshape(-1, 3)
        n = origs.shape[0]
        t = np.empty(n, np.float32)
        face = np.empty(n, np.uint32)
        ng = np.empty((n, 3), np.float32)
        hit = np.empty(n, np.uint8)
        lib().orc_intersect_batch(self._h, C.c_uint32(n), _p(origs), _p(dirs), C.c_float(tfar), C.c_int(int(brute)), _p(t), _p(face), _p(ng), _p(hit))
        return t, face, ng, hit

    # ---- find -------------------------------------------------------------------------------
    def simulate(self, Tbm, Tsb, origs_s, dirs_s, range_max, range_min=0.0, tfar_mode=0, min_mode=0, miss_fill=0):
        """tfar_mode / min_mode / miss_fill: the open rmagine semantics of SURVEY.md A.3 (orc_sim_options); all 0 = the stated defaults"""
        origs_s = _f32(origs_s).reshape(-1, 3)
        dirs_s = _f32(dirs_s).reshape(-1, 3)
        n = dirs_s.shape[0]
        out = dict(points=np.empty((n, 3), np.float32), normals=np.empty((n, 3), np.float32), hits=np.empty(n, np.uint8),
                   face_ids=np.empty(n, np.uint32), ranges=np.empty(n, np.float32))
        Tbm, Tsb = _tf(Tbm), _tf(Tsb)
        opt = (C.c_int * 3)(int(tfar_mode), int(min_mode), int(miss_fill))
        lib().orc_simulate_opt(self._h, _p(Tbm), _p(Tsb), C.c_uint32(n), _p(origs_s), C.c_uint32(origs_s.shape[0]), _p(dirs_s), C.c_float(range_min), C.c_float(range_max),
                               opt, _p(out["points"]), _p(out["normals"]), _p(out["hits"]), _p(out["face_ids"]), _p(out["ranges"]))
        return out

    def micp_correct_once(self, origs_s, dirs_s, range_max, dataset_pts, dataset_mask, Tom, Tbo, Tsb, iterations=5, max_dist=1.0,
                          adaptive_max_dist_min=0.15, convergence_progress=0.0, f64_accum=False):
        dp = _f32(dataset_pts).reshape(-1, 3)
        dm = np.ascontiguousarray(dataset_mask, np.uint8)
        if dirs_s is None:                      # closest-point correspondences (CPCEmbree): one query per dataset point
            origs_s, n = np.zeros((1, 3), np.float32), dp.shape[0]
        else:
            origs_s = _f32(origs_s).reshape(-1, 3)
            dirs_s = _f32(dirs_s).reshape(-1, 3)
            n = dirs_s.shape[0]
        Tn = np.zeros((), TRANSFORM)
        Td = np.zeros((), TRANSFORM)
        Cm = np.zeros((), CROSS_STATS)
        Tom, Tbo, Tsb = _tf(Tom), _tf(Tbo), _tf(Tsb)
        lib().orc_micp_correct_once(self._h, C.c_uint32(n), _p(origs_s), C.c_uint32(origs_s.shape[0]), _p(dirs_s), C.c_float(range_max),
                                    _p(dp), _p(dm), _p(Tom), _p(Tbo), _p(Tsb), C.c_uint32(iterations), C.c_float(max_dist),
                                    C.c_float(adaptive_max_dist_min), C.c_double(convergence_progress), C.c_int(int(f64_accum)),   # 0 seq f32, 1 f64, 2 parallel f32
                                    _p(Tn), _p(Td), _p(Cm))
        return Tn, Td, Cm

    def micp_correct_once_multi(self, sensors, Tom, iterations=5, convergence_progress=0.0, f64_accum=True):
        """sensors: list of dicts(origs, dirs (None = closest point), range_max, dataset_points, dataset_mask, Tbo, Tsb, max_dist, adaptive_max_dist_min, weight)"""
        arr = (_MicpSensor * len(sensors))()
        keep = []
        for k, sd in enumerate(sensors):
            dp = _f32(sd["dataset_points"]).reshape(-1, 3)
            dm = np.ascontiguousarray(sd["dataset_mask"], np.uint8)
            if sd.get("dirs") is None:
                o, d, n = np.zeros((1, 3), np.float32), None, dp.shape[0]
            else:
                o, d = _f32(sd["origs"]).reshape(-1, 3), _f32(sd["dirs"]).reshape(-1, 3)
                n = d.shape[0]
            keep += [dp, dm, o, d]
            a = arr[k]
            a.n, a.n_origs = n, o.shape[0]
            a.origs_s, a.dirs_s = o.ctypes.data, (d.ctypes.data if d is not None else None)
            a.dataset_pts, a.dataset_mask = dp.ctypes.data, dm.ctypes.data
            C.memmove(C.byref(a, _MicpSensor.Tbo.offset), _tf(sd["Tbo"]).ctypes.data, 32)
            C.memmove(C.byref(a, _MicpSensor.Tsb.offset), _tf(sd["Tsb"]).ctypes.data, 32)
            a.range_max, a.max_dist, a.adaptive_max_dist_min = sd.get("range_max", 0.0), sd.get("max_dist", 1.0), sd.get("adaptive_max_dist_min", 0.15)
            a.merge_weight = sd.get("weight", 1.0)
        Tn, Td, Cm = np.zeros((), TRANSFORM), np.zeros((), TRANSFORM), np.zeros((), CROSS_STATS)
        lib().orc_micp_correct_once_multi(self._h, C.c_uint32(len(sensors)), arr, _p(_tf(Tom)), C.c_uint32(iterations), C.c_double(convergence_progress),
                                          C.c_int(int(f64_accum)), _p(Tn), _p(Td), _p(Cm))
        return Tn, Td, Cm

    def correct_batch(self, Tbm, Tsb, origs_s, dirs_s, range_min, range_max, ranges, max_dist, f64_accum=False):
        origs_s = _f32(origs_s).reshape(-1, 3)
        dirs_s = _f32(dirs_s).reshape(-1, 3)
        n = dirs_s.shape[0]
        Tbm = _tf(Tbm).reshape(-1)
        Tsb = _tf(Tsb)
        P = Tbm.shape[0]
        ranges = _f32(ranges)
        Td = np.zeros(P, TRANSFORM)
        nc = np.zeros(P, np.uint32)
        st = np.zeros(P, CROSS_STATS)
        lib().orc_correct_batch(self._h, C.c_uint32(P), _p(Tbm), _p(Tsb), C.c_uint32(n), _p(origs_s), C.c_uint32(origs_s.shape[0]), _p(dirs_s),
                                C.c_float(range_min), C.c_float(range_max), _p(ranges), C.c_float(max_dist), C.c_int(int(f64_accum)),
                                _p(Td), _p(nc), _p(st))
        return Td, nc, st

    # ---- closest-point correspondences (CPCEmbree::find) -------------------------------------
    def cpc_find(self, Tbm, Tsb, dataset_pts, max_dist, brute=False):
        dp = _f32(dataset_pts).reshape(-1, 3)
        n = len(dp)
        out = dict(points=np.empty((n, 3), np.float32), normals=np.empty((n, 3), np.float32), hits=np.empty(n, np.uint8),
                   face_ids=np.empty(n, np.uint32), dists=np.empty(n, np.float32))
        Tbm, Tsb = _tf(Tbm), _tf(Tsb)
        lib().orc_cpc_find(self._h, _p(Tbm), _p(Tsb), C.c_uint32(n), _p(dp), C.c_float(max_dist), C.c_int(int(brute)),
                           _p(out["points"]), _p(out["normals"]), _p(out["hits"]), _p(out["face_ids"]), _p(out["dists"]))
        return out

    # ---- particle filter --------------------------------------------------------------------
    def pf_update(self, poses, attrs, Tsb, beams, params: PFParams):
        poses = _tf(poses).reshape(-1)
        attrs = np.ascontiguousarray(attrs).copy()
        assert attrs.dtype.itemsize == 36
        beams = np.ascontiguousarray(beams)
        assert beams.dtype.itemsize == 64
        Tsb = _tf(Tsb)
        lib().orc_pf_update(self._h, C.c_uint32(poses.shape[0]), _p(poses), _p(attrs), _p(Tsb), C.c_uint32(beams.shape[0]), _p(beams), C.byref(params))
        return attrs


def segment(origs_s, dirs_s, range_min, range_max, ranges_real, ranges_sim, normals_sim, min_dist_outlier_scan, min_dist_outlier_map):
    origs_s, dirs_s = _f32(origs_s).reshape(-1, 3), _f32(dirs_s).reshape(-1, 3)
    n = len(dirs_s)
    rr, rs, ns_ = _f32(ranges_real).reshape(-1), _f32(ranges_sim).reshape(-1), _f32(normals_sim).reshape(-1, 3)
    a, b, lab = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.uint8)
    na, nb = C.c_uint32(), C.c_uint32()
    lib().orc_segment(C.c_uint32(n), _p(origs_s), C.c_uint32(len(origs_s)), _p(dirs_s), C.c_float(range_min), C.c_float(range_max), _p(rr), _p(rs), _p(ns_),
                      C.c_float(min_dist_outlier_scan), C.c_float(min_dist_outlier_map), _p(a), C.byref(na), _p(b), C.byref(nb), _p(lab))
    return a[: na.value].copy(), b[: nb.value].copy(), lab


class GladiatorConfig(C.Structure):
    """GladiatorResamplerConfig (rmcl_ros/include/rmcl_ros/rmcl/GladiatorResamplerConfig.hpp:7-20)."""
    _fields_ = [("min_noise_tx", C.c_float), ("min_noise_ty", C.c_float), ("min_noise_tz", C.c_float), ("min_noise_roll", C.c_float),
                ("min_noise_pitch", C.c_float), ("min_noise_yaw", C.c_float), ("likelihood_forget_per_meter", C.c_float),
                ("likelihood_forget_per_radian", C.c_float)]

    @staticmethod
    def defaults():
        return GladiatorConfig(0.03, 0.03, 0.0, 0.0, 0.0, 0.01, 0.3, 0.2)


def philox4x32_10(ctr, key):
    c, k, o = np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def pf_gladiator_randoms(seed, step, first, n):
    raw, nrm = np.empty(n, np.uint32), np.empty((n, 6), np.float32)
    lib().orc_pf_gladiator_randoms(C.c_uint64(seed), C.c_uint32(step), C.c_uint32(first), C.c_uint32(n), _p(raw), _p(nrm))
    return raw, nrm


def pf_gladiator_resample(poses, attrs, first, n_local, raw, normals, cfg):
    poses, attrs = _tf(poses).reshape(-1), np.ascontiguousarray(attrs)
    raw, normals = np.ascontiguousarray(raw, np.uint32), _f32(normals).reshape(-1, 6)
    assert len(raw) == n_local and len(normals) == n_local
    Pn, An = np.zeros(n_local, TRANSFORM), np.zeros(n_local, attrs.dtype)
    lib().orc_pf_gladiator_resample(C.c_uint32(len(poses)), _p(poses), _p(attrs), C.c_uint32(first), C.c_uint32(n_local), _p(raw), _p(normals),
                                    C.byref(cfg), _p(Pn), _p(An))
    return Pn, An


def spherical_dirs(m):
    d = np.empty((m.phi_size * m.theta_size, 3), np.float32)
    lib().orc_spherical_dirs(C.c_float(m.phi_min), C.c_float(m.phi_inc), C.c_uint32(m.phi_size), C.c_float(m.theta_min), C.c_float(m.theta_inc),
                             C.c_uint32(m.theta_size), _p(d))
    return d


def pinhole_dirs(m):
    d = np.empty((m.width * m.height, 3), np.float32)
    lib().orc_pinhole_dirs(C.c_uint32(m.width), C.c_uint32(m.height), C.c_float(m.fx), C.c_float(m.fy), C.c_float(m.cx), C.c_float(m.cy), _p(d))
    return d


def model_rays(m):
    """(origs_s (1|n,3), dirs_s (n,3)) for any synth sensor model."""
    name = type(m).__name__
    if name == "SphericalModel":
        return np.zeros((1, 3), np.float32), spherical_dirs(m)
    if name == "PinholeModel":
        return np.zeros((1, 3), np.float32), pinhole_dirs(m)
    if name == "O1DnModel":
        return _f32(m.orig).reshape(1, 3), _f32(m.dirs).reshape(-1, 3)
    if name == "OnDnModel":
        return _f32(m.origs).reshape(-1, 3), _f32(m.dirs).reshape(-1, 3)
    raise TypeError(name)


def dataset_from_ranges(origs_s, dirs_s, ranges, range_min, range_max):
    origs_s = _f32(origs_s).reshape(-1, 3)
    dirs_s = _f32(dirs_s).reshape(-1, 3)
    ranges = _f32(ranges)
    n = dirs_s.shape[0]
    pts = np.empty((n, 3), np.float32)
    mask = np.empty(n, np.uint8)
    nv = C.c_uint32(0)
    lib().orc_dataset_from_ranges(C.c_uint32(n), _p(origs_s), C.c_uint32(origs_s.shape[0]), _p(dirs_s), _p(ranges), C.c_float(range_min),
                                  C.c_float(range_max), _p(pts), _p(mask), C.byref(nv))
    return pts, mask, nv.value


def statistics_p2l(Tpre, dpts, dmask, mpts, mnrm, mmask, max_dist, f64=False):
    Tpre = _tf(Tpre)
    dpts, mpts, mnrm = _f32(dpts), _f32(mpts), _f32(mnrm)
    dmask = None if dmask is None else np.ascontiguousarray(dmask, np.uint8)
    mmask = None if mmask is None else np.ascontiguousarray(mmask, np.uint8)
    out = np.zeros((), CROSS_STATS)
    fn = lib().orc_statistics_p2l_f64 if f64 else lib().orc_statistics_p2l
    fn(_p(Tpre), C.c_uint32(dpts.reshape(-1, 3).shape[0]), _p(dpts), _p(dmask), _p(mpts), _p(mnrm), _p(mmask), C.c_float(max_dist), _p(out))
    return out


def adaptive_max_dist(max_dist, adaptive_min, cp):
    return float(lib().orc_adaptive_max_dist(C.c_float(max_dist), C.c_float(adaptive_min), C.c_double(cp)))


def umeyama(stats):
    stats = np.ascontiguousarray(stats)
    out = np.zeros((), TRANSFORM)
    lib().orc_umeyama(_p(stats), _p(out))
    return out


def cross_stats_merge(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    out = np.zeros((), CROSS_STATS)
    lib().orc_cross_stats_merge(_p(a), _p(b), _p(out))
    return out


def cross_stats_transform(T, s):
    T, s = _tf(T), np.ascontiguousarray(s)
    out = np.zeros((), CROSS_STATS)
    lib().orc_cross_stats_transform(_p(T), _p(s), _p(out))
    return out


def transform_mul(a, b):
    a, b = _tf(a), _tf(b)
    out = np.zeros((), TRANSFORM)
    lib().orc_transform_mul(_p(a), _p(b), _p(out))
    return out


def transform_inv(a):
    a = _tf(a)
    out = np.zeros((), TRANSFORM)
    lib().orc_transform_inv(_p(a), _p(out))
    return out


def pf_motion_update(poses, attrs, T_bnew_bold, forget_rate, scene=None):
    """scene: a Scene -> walls are checked (TFMotionUpdaterCPU with a map); None -> plain update (TFMotionUpdaterGPU)."""
    poses = _tf(poses).reshape(-1).copy()
    attrs = np.ascontiguousarray(attrs).copy()
    T = _tf(T_bnew_bold)
    lib().orc_pf_motion_update_collide(scene._h if scene is not None else None, C.c_uint32(len(poses)), _p(poses), _p(attrs), _p(T), C.c_double(forget_rate))
    return poses, attrs


def pf_likelihood_stats(attrs):
    attrs = np.ascontiguousarray(attrs)
    s, m = C.c_float(), C.c_float()
    lib().orc_pf_likelihood_stats(C.c_uint32(len(attrs)), _p(attrs), C.byref(s), C.byref(m))
    return s.value, m.value
