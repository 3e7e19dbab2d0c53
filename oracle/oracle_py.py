"""ctypes bindings of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (lio_mapping_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cc", ".h")) or f == "Makefile"]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(so):
        build()
    L = C.CDLL(so)
    L.orc_a_create.restype = C.c_void_p
    L.orc_a_create.argtypes = [C.c_float, C.c_float, C.c_int, C.c_double]
    L.orc_a_destroy.argtypes = [C.c_void_p]
    L.orc_a_run.argtypes = [C.c_void_p, f32p, C.c_int]
    L.orc_a_cloud_size.argtypes = [C.c_void_p, C.c_int]
    L.orc_a_cloud_copy.argtypes = [C.c_void_p, C.c_int, f32p]
    L.orc_a_idx_size.argtypes = [C.c_void_p, C.c_int]
    L.orc_a_idx_copy.argtypes = [C.c_void_p, C.c_int, i32p]
    L.orc_a_scan_ranges.argtypes = [C.c_void_p, i32p]
    L.orc_a_mask_labels.argtypes = [C.c_void_p, u8p, i8p]
    L.orc_a_start_ori.restype = C.c_float
    L.orc_a_start_ori.argtypes = [C.c_void_p]
    L.orc_normalize_rad.restype = C.c_double
    L.orc_normalize_rad.argtypes = [C.c_double]
    L.orc_normalize_deg.restype = C.c_double
    L.orc_normalize_deg.argtypes = [C.c_double]
    L.orc_voxel_grid.restype = C.c_int
    L.orc_voxel_grid.argtypes = [f32p, C.c_int, C.c_float, f32p]
    L.orc_transform_cloud.argtypes = [f32p, C.c_int, f32p, f32p, f32p]
    L.orc_knn.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, i32p, f32p]
    L.orc_calculate_features.restype = C.c_int
    L.orc_calculate_features.argtypes = [f32p, C.c_int, f32p, C.c_int, f32p, C.c_float, C.c_float, f32p, f32p, i32p]
    L.orc_laser_odom.restype = C.c_int
    L.orc_laser_odom.argtypes = [f32p, C.c_int, f32p, C.c_int, f32p, C.c_float, C.c_float, C.c_int, C.c_int,
                                 f32p, f32p, i32p, i32p]
    _LIB = L
    return L


CLOUDS = {"laser_scans": 0, "cloud_in_rings": 1, "sharp": 2, "less_sharp": 3, "flat": 4, "less_flat": 5}
IDX = {"sharp": 0, "less_sharp": 1, "flat": 2, "less_flat_prevoxel": 3, "orig_index": 4}


def stage_a(xyzi: np.ndarray, lower: float, upper: float, rings: int, scan_period: float = 0.1) -> dict:
    """PointProcessor::PointToRing + ExtractFeaturePoints on one sweep (oracle)."""
    L = lib()
    xyzi = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
    h = L.orc_a_create(lower, upper, rings, scan_period)
    try:
        L.orc_a_run(h, xyzi, xyzi.shape[0])
        out = {}
        for name, w in CLOUDS.items():
            n = L.orc_a_cloud_size(h, w)
            a = np.zeros((n, 4), np.float32)
            L.orc_a_cloud_copy(h, w, a)
            out[name] = a
        for name, w in IDX.items():
            n = L.orc_a_idx_size(h, w)
            a = np.zeros(n, np.int32)
            L.orc_a_idx_copy(h, w, a)
            out["idx_" + name] = a
        sr = np.zeros(2 * rings, np.int32)
        L.orc_a_scan_ranges(h, sr)
        out["scan_ranges"] = sr.reshape(rings, 2)
        n = out["laser_scans"].shape[0]
        m = np.zeros(max(n, 1), np.uint8)
        lab = np.zeros(max(n, 1), np.int8)
        L.orc_a_mask_labels(h, m, lab)
        out["mask"], out["labels"] = m[:n], lab[:n]
        out["start_ori"] = float(L.orc_a_start_ori(h))
        return out
    finally:
        L.orc_a_destroy(h)


def voxel_grid(cloud: np.ndarray, leaf: float) -> np.ndarray:
    L = lib()
    cloud = np.ascontiguousarray(cloud, np.float32).reshape(-1, 4)
    out = np.zeros_like(cloud)
    n = L.orc_voxel_grid(cloud, cloud.shape[0], leaf, out) if cloud.shape[0] else 0
    return out[:n].copy()


def transform_cloud(cloud, R, t):
    L = lib()
    cloud = np.ascontiguousarray(cloud, np.float32).reshape(-1, 4)
    out = np.zeros_like(cloud)
    L.orc_transform_cloud(cloud, cloud.shape[0], np.ascontiguousarray(R, np.float32).reshape(9),
                          np.ascontiguousarray(t, np.float32).reshape(3), out)
    return out


def knn(map_pts, queries, k=5):
    L = lib()
    m = np.ascontiguousarray(map_pts, np.float32).reshape(-1, 4)
    q = np.ascontiguousarray(queries, np.float32).reshape(-1, 4)
    idx = np.zeros((q.shape[0], k), np.int32)
    d2 = np.zeros((q.shape[0], k), np.float32)
    L.orc_knn(m, m.shape[0], q, q.shape[0], k, idx, d2)
    return idx, d2


def calculate_features(map_pts, surf, tf7, min_match_sq_dis=1.0, min_plane_dis=0.2):
    L = lib()
    m = np.ascontiguousarray(map_pts, np.float32).reshape(-1, 4)
    s = np.ascontiguousarray(surf, np.float32).reshape(-1, 4)
    pts = np.zeros((s.shape[0], 4), np.float32)
    coef = np.zeros((s.shape[0], 4), np.float32)
    src = np.zeros(s.shape[0], np.int32)
    n = L.orc_calculate_features(m, m.shape[0], s, s.shape[0], np.ascontiguousarray(tf7, np.float32), min_match_sq_dis,
                                 min_plane_dis, pts, coef, src)
    return pts[:n].copy(), coef[:n].copy(), src[:n].copy()


def laser_odom(map_pts, surf, tf7, min_match_sq_dis=1.0, min_plane_dis=0.2, keep_features=0, max_iter=10):
    L = lib()
    m = np.ascontiguousarray(map_pts, np.float32).reshape(-1, 4)
    s = np.ascontiguousarray(surf, np.float32).reshape(-1, 4)
    cap = s.shape[0] * (max_iter if keep_features else 1)
    pts = np.zeros((cap, 4), np.float32)
    coef = np.zeros((cap, 4), np.float32)
    src = np.zeros(cap, np.int32)
    tf = np.ascontiguousarray(tf7, np.float32).copy()
    it = np.zeros(1, np.int32)
    n = L.orc_laser_odom(m, m.shape[0], s, s.shape[0], tf, min_match_sq_dis, min_plane_dis, keep_features, max_iter,
                         pts, coef, src, it)
    return tf, pts[:n].copy(), coef[:n].copy(), src[:n].copy(), int(it[0])
