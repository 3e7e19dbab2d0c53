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


def stage_a(xyzi: np.ndarray, lower: float, upper: float, rings: int, scan_period: float = 0.1, ring_field=None) -> dict:
    """PointProcessor::PointToRing + ExtractFeaturePoints on one sweep (oracle).  ring_field: uint16 ring index per point
    (PointXYZIR input, PointProcessor.cc:428-536) instead of the elevation-derived ring."""
    L = lib()
    xyzi = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
    h = L.orc_a_create(lower, upper, rings, scan_period)
    try:
        if ring_field is None:
            L.orc_a_run(h, xyzi, xyzi.shape[0])
        else:
            rf = np.ascontiguousarray(ring_field, np.uint16)
            L.orc_a_run_ring.argtypes = [C.c_void_p, f32p, np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS"), C.c_int]
            L.orc_a_run_ring.restype = None
            L.orc_a_run_ring(h, xyzi, rf, xyzi.shape[0])
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


def calculate_line_features(corner_map, corner, tf7, min_match_sq_dis=1.0):
    L = lib()
    m = np.ascontiguousarray(corner_map, np.float32).reshape(-1, 4)
    s = np.ascontiguousarray(corner, np.float32).reshape(-1, 4)
    cap = max(2 * s.shape[0], 1)
    pts = np.zeros((cap, 4), np.float32)
    coef = np.zeros((cap, 4), np.float32)
    src = np.zeros(cap, np.int32)
    L.orc_calculate_line_features.argtypes = [f32p, C.c_int, f32p, C.c_int, f32p, C.c_float, f32p, f32p, i32p]
    L.orc_calculate_line_features.restype = C.c_int
    n = L.orc_calculate_line_features(m, m.shape[0], s, s.shape[0], np.ascontiguousarray(tf7, np.float32), min_match_sq_dis,
                                      pts, coef, src)
    return pts[:n].copy(), coef[:n].copy(), src[:n].copy()


def scan_to_map(corner_map, surf_map, corner, surf, tf7, min_match_sq_dis=1.0, min_plane_dis=0.2, max_iter=10,
                delta_r_abort=0.05, delta_t_abort=0.05, variant=0):
    """PointMapping::OptimizeTransformTobeMapped (PointMapping.cc:325-753; variant 1 = MapBuilder::OptimizeMap,
    MapBuilder.cc:624-1014): returns (tf7, pts, coef, src, iterations)."""
    L = lib()
    a = [np.ascontiguousarray(x, np.float32).reshape(-1, 4) for x in (corner_map, surf_map, corner, surf)]
    cap = max(a[2].shape[0] + a[3].shape[0], 1)
    pts = np.zeros((cap, 4), np.float32); coef = np.zeros((cap, 4), np.float32); src = np.zeros(cap, np.int32)
    tf = np.ascontiguousarray(tf7, np.float32).copy()
    it = np.zeros(1, np.int32)
    L.orc_scan_to_map.argtypes = [f32p, C.c_int, f32p, C.c_int, f32p, C.c_int, f32p, C.c_int, f32p, C.c_float, C.c_float, C.c_int,
                                  C.c_double, C.c_double, f32p, f32p, i32p, i32p, C.c_int]
    L.orc_scan_to_map.restype = C.c_int
    n = L.orc_scan_to_map(a[0], a[0].shape[0], a[1], a[1].shape[0], a[2], a[2].shape[0], a[3], a[3].shape[0], tf, min_match_sq_dis,
                          min_plane_dis, max_iter, delta_r_abort, delta_t_abort, pts, coef, src, it, variant)
    return tf, pts[:n].copy(), coef[:n].copy(), src[:n].copy(), int(it[0])


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


def compact_encode(tf7, corner, surf, full):
    """PointOdometry.cc:732-762: the compact cloud as (n, 4) float32."""
    L = lib()
    c, s, f = [np.ascontiguousarray(a, np.float32).reshape(-1, 4) for a in (corner, surf, full)]
    out = np.zeros((3 + c.shape[0] + s.shape[0] + f.shape[0], 4), np.float32)
    L.orc_compact_encode.argtypes = [f32p, f32p, C.c_int, f32p, C.c_int, f32p, C.c_int, f32p]
    L.orc_compact_encode.restype = C.c_int
    n = L.orc_compact_encode(np.ascontiguousarray(tf7, np.float32), c, c.shape[0], s, s.shape[0], f, f.shape[0], out)
    return out[:n]


def compact_decode(compact):
    """PointMapping::CompactDataHandler (PointMapping.cc:171-238); None on the reference's error paths."""
    L = lib()
    d = np.ascontiguousarray(compact, np.float32).reshape(-1, 4)
    n = d.shape[0]
    tf7 = np.zeros(7, np.float32)
    outs = [np.zeros((max(n, 1), 4), np.float32) for _ in range(3)]
    sz = np.zeros(3, np.int32)
    L.orc_compact_decode.argtypes = [f32p, C.c_int, f32p, f32p, f32p, f32p, i32p]
    L.orc_compact_decode.restype = C.c_int
    if not L.orc_compact_decode(d, n, tf7, outs[0], outs[1], outs[2], sz):
        return None
    return tf7, outs[0][:sz[0]], outs[1][:sz[1]], outs[2][:sz[2]]


def toy_solve(A, B, y, x0, amp=0.05, use_cauchy=False, max_iter=10):
    """Toy nonlinear least squares through the oracle's Problem / Solve (Ceres-1.14 restatement): returns (x, summary)."""
    L = lib()
    A = np.ascontiguousarray(A, np.float64); B = np.ascontiguousarray(B, np.float64)
    m, n = A.shape
    x = np.ascontiguousarray(x0, np.float64).copy()
    s = np.zeros(8)
    L.orc_toy_solve.argtypes = [C.c_int, C.c_int, f64p, f64p, f64p, C.c_double, C.c_int, f64p, C.c_int, f64p]
    L.orc_toy_solve(n, m, A, B, np.ascontiguousarray(y, np.float64), amp, int(use_cauchy), x, max_iter, s)
    return x, dict(iterations=int(s[0]), successful=int(s[1]), termination=int(s[2]), initial_cost=s[3], final_cost=s[4])


def transform_to_end(cloud, tf7, time_factor=10.0):
    """TransformToEnd (Estimator.cc:62-103) on a copy of `cloud`."""
    L = lib()
    c = np.ascontiguousarray(cloud, np.float32).reshape(-1, 4).copy()
    L.orc_transform_to_end.argtypes = [f32p, C.c_int, f32p, C.c_float]
    L.orc_transform_to_end.restype = None
    L.orc_transform_to_end(c, c.shape[0], np.ascontiguousarray(tf7, np.float32), time_factor)
    return c


# =================================================================================================
# fp64 factors / solver / estimator bindings
def _bind_factors(L):
    if getattr(L, "_factors_bound", False):
        return
    vp = C.c_void_p
    L.orc_ppp_evaluate.argtypes = [f64p, f64p, f64p, f64p, f64p, f64p, f64p, f64p, f64p]
    L.orc_prior_evaluate.argtypes = [f64p, f64p, f64p, f64p, f64p]
    L.orc_pose_plus.argtypes = [f64p, f64p, f64p]
    L.orc_pim_create.restype = vp
    L.orc_pim_create.argtypes = [f64p, f64p, f64p, f64p, f64p]
    L.orc_pim_destroy.argtypes = [vp]
    L.orc_pim_push_back.argtypes = [vp, C.c_double, f64p, f64p]
    L.orc_pim_get.argtypes = [vp, f64p, f64p, f64p]
    L.orc_imu_factor_evaluate.argtypes = [vp, f64p, f64p, f64p, f64p, f64p, f64p, f64p, f64p, f64p]
    L.orc_sym_eigen.argtypes = [f64p, C.c_int, f64p, f64p]
    L.orc_est_create.restype = vp
    L.orc_est_create.argtypes = [f64p]
    L.orc_est_destroy.argtypes = [vp]
    L.orc_est_set_extrinsic.argtypes = [vp, f32p]
    L.orc_est_init_frame.argtypes = [vp, C.c_int, f64p, f32p, C.c_int, vp]
    L.orc_est_finish_init.argtypes = [vp, f64p, f64p]
    L.orc_est_process_imu.argtypes = [vp, C.c_double, f64p, f64p, C.c_double]
    L.orc_est_process_scan.argtypes = [vp, f32p, C.c_int]
    L.orc_est_get_states.argtypes = [vp, f64p]
    L.orc_est_get_extrinsic.argtypes = [vp, f32p]
    L.orc_est_summary.argtypes = [vp, f64p]
    L.orc_est_feature_count.argtypes = [vp, C.c_int]
    L.orc_est_get_features.argtypes = [vp, C.c_int, f32p, f32p, i32p]
    L.orc_est_map_size.argtypes = [vp]
    L.orc_est_get_map.argtypes = [vp, f32p]
    L.orc_est_frame_size.argtypes = [vp, C.c_int]
    L.orc_est_get_frame.argtypes = [vp, C.c_int, f32p]
    L.orc_est_get_local_transform.argtypes = [vp, C.c_int, f32p]
    L.orc_est_prior_dim.argtypes = [vp]
    L.orc_est_prior_blocks.argtypes = [vp, i32p, i32p, i32p]
    L.orc_est_normal_dim.argtypes = [vp]
    L.orc_est_get_normal.argtypes = [vp, f64p, f64p]
    L.orc_est_get_prior.argtypes = [vp, f64p, f64p]
    L._factors_bound = True


def _d(a):
    return np.ascontiguousarray(a, np.float64)


def ppp_evaluate(point, coeff, pose_pivot, pose_i, pose_ex):
    L = lib(); _bind_factors(L)
    r = np.zeros(1); J = [np.zeros(7) for _ in range(3)]
    L.orc_ppp_evaluate(_d(point), _d(coeff), _d(pose_pivot), _d(pose_i), _d(pose_ex), r, J[0], J[1], J[2])
    return float(r[0]), J


def prior_evaluate(pos, quat_xyzw, pose):
    L = lib(); _bind_factors(L)
    r = np.zeros(6); J = np.zeros(42)
    L.orc_prior_evaluate(_d(pos), _d(quat_xyzw), _d(pose), r, J)
    return r, J.reshape(6, 7)


def pose_plus(x, delta):
    L = lib(); _bind_factors(L)
    out = np.zeros(7)
    L.orc_pose_plus(_d(x), _d(delta), out)
    return out


def sym_eigen(A):
    L = lib(); _bind_factors(L)
    A = _d(A); n = A.shape[0]
    ev = np.zeros(n); V = np.zeros((n, n))
    L.orc_sym_eigen(A, n, ev, V)
    return ev, V


class Pim:
    """IntegrationBase (include/imu_processor/IntegrationBase.h)."""

    def __init__(self, acc0, gyr0, ba, bg, acc_n=0.1, gyr_n=0.01, acc_w=2e-4, gyr_w=2e-5, g_norm=9.805):
        L = lib(); _bind_factors(L)
        self.L = L
        self.h = L.orc_pim_create(_d(acc0), _d(gyr0), _d(ba), _d(bg), _d([acc_n, gyr_n, acc_w, gyr_w, g_norm]))

    def push_back(self, dt, acc, gyr):
        self.L.orc_pim_push_back(self.h, float(dt), _d(acc), _d(gyr))

    def get(self):
        s = np.zeros(11); J = np.zeros(225); P = np.zeros(225)
        self.L.orc_pim_get(self.h, s, J, P)
        return dict(delta_p=s[0:3], delta_q=s[3:7], delta_v=s[7:10], sum_dt=s[10], jacobian=J.reshape(15, 15), covariance=P.reshape(15, 15))

    def imu_factor(self, pose_i, sb_i, pose_j, sb_j, jac=True):
        r = np.zeros(15)
        J = [np.zeros(15 * 7), np.zeros(15 * 9), np.zeros(15 * 7), np.zeros(15 * 9)]
        self.L.orc_imu_factor_evaluate(self.h, _d(pose_i), _d(sb_i), _d(pose_j), _d(sb_j), r, *J)
        return r, [J[0].reshape(15, 7), J[1].reshape(15, 9), J[2].reshape(15, 7), J[3].reshape(15, 9)]

    def __del__(self):
        try:
            self.L.orc_pim_destroy(self.h)
        except Exception:
            pass


EST_CFG_DEFAULT = dict(window_size=10, opt_window_size=10, min_match_sq_dis=1.0, min_plane_dis=0.2, surf_filter_size=0.4,
                       keep_features=0, estimate_extrinsic=1, opt_extrinsic=1, imu_factor=1, point_distance_factor=1,
                       prior_factor=0, marginalization_factor=1, enable_deskew=1, cutoff_deskew=1, acc_n=0.2, gyr_n=0.02,
                       acc_w=2e-4, gyr_w=2e-5, g_norm=9.805, max_num_iterations=10, odom_max_iterations=10)
EST_CFG_ORDER = ["window_size", "opt_window_size", "min_match_sq_dis", "min_plane_dis", "surf_filter_size", "keep_features",
                 "estimate_extrinsic", "opt_extrinsic", "imu_factor", "point_distance_factor", "prior_factor",
                 "marginalization_factor", "enable_deskew", "cutoff_deskew", "acc_n", "gyr_n", "acc_w", "gyr_w", "g_norm",
                 "max_num_iterations", "odom_max_iterations"]
SUMMARY_KEYS = ["iterations", "successful", "termination", "initial_cost", "final_cost", "cost_pim", "cost_ppp", "cost_marg",
                "turn_off", "convergence_flag", "map_size", "num_features", "odom_iters", "t_build_map", "t_features",
                "t_solve", "t_marg", "t_total", "has_prior", "linearizations", "cost_evals"]


class Estimator:
    """Steady-state lio::Estimator (oracle)."""

    def __init__(self, **cfg):
        L = lib(); _bind_factors(L)
        self.L = L
        c = dict(EST_CFG_DEFAULT); c.update(cfg)
        self.cfg = c
        self.W = int(c["window_size"])
        self.h = L.orc_est_create(_d([c[k] for k in EST_CFG_ORDER]))

    def set_extrinsic(self, tf7):
        self.L.orc_est_set_extrinsic(self.h, np.ascontiguousarray(tf7, np.float32))

    def init_frame(self, k, state16, surf_ds, pim: "Pim | None"):
        s = np.ascontiguousarray(surf_ds, np.float32).reshape(-1, 4)
        self.L.orc_est_init_frame(self.h, k, _d(state16), s, s.shape[0], pim.h if pim is not None else None)

    def finish_init(self, acc_last, gyr_last):
        self.L.orc_est_finish_init(self.h, _d(acc_last), _d(gyr_last))

    def process_imu(self, dt, acc, gyr, stamp):
        self.L.orc_est_process_imu(self.h, float(dt), _d(acc), _d(gyr), float(stamp))

    def process_scan(self, surf_last):
        s = np.ascontiguousarray(surf_last, np.float32).reshape(-1, 4)
        self.L.orc_est_process_scan(self.h, s, s.shape[0])

    def states(self):
        out = np.zeros((self.W + 1, 16))
        self.L.orc_est_get_states(self.h, out)
        return out

    def extrinsic(self):
        t = np.zeros(7, np.float32)
        self.L.orc_est_get_extrinsic(self.h, t)
        return t

    def summary(self):
        s = np.zeros(32)
        self.L.orc_est_summary(self.h, s)
        return dict(zip(SUMMARY_KEYS, s.tolist()))

    def features(self, frame):
        n = self.L.orc_est_feature_count(self.h, frame)
        p = np.zeros((max(n, 1), 4), np.float32); c = np.zeros((max(n, 1), 4), np.float32); s = np.zeros(max(n, 1), np.int32)
        self.L.orc_est_get_features(self.h, frame, p, c, s)
        return p[:n], c[:n], s[:n]

    def local_map(self):
        n = self.L.orc_est_map_size(self.h)
        m = np.zeros((max(n, 1), 4), np.float32)
        self.L.orc_est_get_map(self.h, m)
        return m[:n]

    def frame(self, k):
        n = self.L.orc_est_frame_size(self.h, k)
        m = np.zeros((max(n, 1), 4), np.float32)
        self.L.orc_est_get_frame(self.h, k, m)
        return m[:n]

    def local_transform(self, k):
        t = np.zeros(7, np.float32)
        self.L.orc_est_get_local_transform(self.h, k, t)
        return t

    def prior_canonical(self, O):
        """(Hp, bp) = (J^T J, J^T r0) permuted to the canonical order [pose_0,sb_0,...,pose_{O-1},sb_{O-1},ex]."""
        J, r = self.prior()
        kind = np.zeros(64, np.int32); idx = np.zeros(64, np.int32); off = np.zeros(64, np.int32)
        nb = self.L.orc_est_prior_blocks(self.h, kind, idx, off)
        n = 15 * O + 6
        H = np.zeros((n, n)); b = np.zeros(n)
        A = J.T @ J
        v = J.T @ r
        cols = []
        for k in range(nb):
            size = 9 if kind[k] == 1 else 6
            base = {0: 15 * idx[k], 1: 15 * idx[k] + 6, 2: 15 * O}[int(kind[k])]
            cols += [(off[k] + a, base + a) for a in range(size)]
        src = np.array([c[0] for c in cols]); dst = np.array([c[1] for c in cols])
        H[np.ix_(dst, dst)] = A[np.ix_(src, src)]
        b[dst] = v[src]
        return H, b

    def normal_equations(self):
        n = self.L.orc_est_normal_dim(self.h)
        H = np.zeros((max(n, 1), max(n, 1))); g = np.zeros(max(n, 1))
        if n:
            self.L.orc_est_get_normal(self.h, H, g)
        return H[:n, :n], g[:n]

    def prior(self):
        n = self.L.orc_est_prior_dim(self.h)
        J = np.zeros((max(n, 1), max(n, 1))); r = np.zeros(max(n, 1))
        if n:
            self.L.orc_est_get_prior(self.h, J, r)
        return J[:n, :n], r[:n]

    def __del__(self):
        try:
            self.L.orc_est_destroy(self.h)
        except Exception:
            pass


# =================================================================================================
# rolling cube map of PointMapping (oracle only: groundwork, see oracle/o_cubemap.cc)
class CubeMap:
    L_, W_, H_ = 21, 21, 11

    def __init__(self):
        self.L = lib()
        i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
        self.L.orc_cm_create.restype = C.c_void_p
        self.L.orc_cm_destroy.argtypes = [C.c_void_p]
        self.L.orc_cm_recentre.argtypes = [C.c_void_p, f32p, i32p]
        self.L.orc_cm_select.argtypes = [C.c_void_p, f32p, f32p, i32p, i64p, i64p, i32p]
        self.L.orc_cm_cube_size.argtypes = [C.c_void_p, C.c_longlong, C.c_int]
        self.L.orc_cm_cube_copy.argtypes = [C.c_void_p, C.c_longlong, C.c_int, f32p]
        self.L.orc_cm_update.argtypes = [C.c_void_p, f32p, C.c_int, f32p, C.c_int, i64p, C.c_int, f32p, i32p]
        self.h = self.L.orc_cm_create()

    def __del__(self):
        try:
            self.L.orc_cm_destroy(self.h)
        except Exception:
            pass

    @staticmethod
    def to_index(i, j, k):
        return i + 21 * j + 21 * 21 * k

    def recentre(self, pos):
        out = np.zeros(6, np.int32)
        self.L.orc_cm_recentre(self.h, np.ascontiguousarray(pos, np.float32), out)
        return tuple(out[:3].tolist()), tuple(out[3:].tolist())

    def select(self, pos, zaxis, centre):
        v = np.zeros(125, np.int64); s = np.zeros(125, np.int64); n = np.zeros(2, np.int32)
        self.L.orc_cm_select(self.h, np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(zaxis, np.float32),
                             np.ascontiguousarray(centre, np.int32), v, s, n)
        return v[:n[0]].copy(), s[:n[1]].copy()

    def cube(self, index, which):
        n = self.L.orc_cm_cube_size(self.h, int(index), 0 if which == "corner" else 1)
        out = np.zeros((max(n, 1), 4), np.float32)
        if n:
            self.L.orc_cm_cube_copy(self.h, int(index), 0 if which == "corner" else 1, out)
        return out[:n]

    def update(self, corner, surf, valid, tf7, margin_centre):
        c = np.ascontiguousarray(corner, np.float32).reshape(-1, 4); s = np.ascontiguousarray(surf, np.float32).reshape(-1, 4)
        cp = c if c.shape[0] else np.zeros((1, 4), np.float32)
        sp = s if s.shape[0] else np.zeros((1, 4), np.float32)
        v = np.ascontiguousarray(valid, np.int64)
        vp = v if v.shape[0] else np.zeros(1, np.int64)
        self.L.orc_cm_update(self.h, cp, c.shape[0], sp, s.shape[0], vp, v.shape[0], np.ascontiguousarray(tf7, np.float32),
                             np.ascontiguousarray(margin_centre, np.int32))


class PointMappingOracle:
    """PointMapping::Process (imu_inited_ == false path, PointMapping.cc:765-1052) around the cube map (oracle only)."""

    def __init__(self):
        self.L = lib()
        self.L.orc_pm_create.restype = C.c_void_p
        self.L.orc_pm_destroy.argtypes = [C.c_void_p]
        self.L.orc_pm_process.argtypes = [C.c_void_p, f32p, C.c_int, f32p, C.c_int, f32p, f32p, i32p]
        self.L.orc_pm_cube_size.argtypes = [C.c_void_p, C.c_longlong, C.c_int]
        self.L.orc_pm_cube_copy.argtypes = [C.c_void_p, C.c_longlong, C.c_int, f32p]
        self.L.orc_pm_centre.argtypes = [C.c_void_p, i32p]
        self.h = self.L.orc_pm_create()

    def __del__(self):
        try:
            self.L.orc_pm_destroy(self.h)
        except Exception:
            pass

    def process(self, corner_last, surf_last, transform_sum7):
        c = np.ascontiguousarray(corner_last, np.float32).reshape(-1, 4); s = np.ascontiguousarray(surf_last, np.float32).reshape(-1, 4)
        tobe = np.zeros(7, np.float32); info = np.zeros(3, np.int32)
        self.L.orc_pm_process(self.h, c if c.shape[0] else np.zeros((1, 4), np.float32), c.shape[0],
                              s if s.shape[0] else np.zeros((1, 4), np.float32), s.shape[0],
                              np.ascontiguousarray(transform_sum7, np.float32), tobe, info)
        return tobe, dict(iterations=int(info[0]), corner_from_map=int(info[1]), surf_from_map=int(info[2]))

    def centre(self):
        out = np.zeros(3, np.int32)
        self.L.orc_pm_centre(self.h, out)
        return tuple(out.tolist())

    def cube(self, index, which):
        w = 0 if which == "corner" else 1
        n = self.L.orc_pm_cube_size(self.h, int(index), w)
        out = np.zeros((max(n, 1), 4), np.float32)
        if n:
            self.L.orc_pm_cube_copy(self.h, int(index), w, out)
        return out[:n]

    def cube_sizes(self, which):
        w = 0 if which == "corner" else 1
        return np.array([self.L.orc_pm_cube_size(self.h, i, w) for i in range(21 * 21 * 11)], np.int64)


class PointOdometryOracle:
    """lio::PointOdometry (src/point_processor/PointOdometry.cc:294-766): scan-to-scan odometry + the /compact_data payload (oracle only)."""

    def __init__(self, scan_period=0.1, io_ratio=2, num_max_iterations=25):
        self.L = lib()
        self.L.orc_po_create.restype = C.c_void_p
        self.L.orc_po_create.argtypes = [C.c_float, C.c_int, C.c_int]
        self.L.orc_po_destroy.argtypes = [C.c_void_p]
        self.L.orc_po_set_enable_odom.argtypes = [C.c_void_p, C.c_int]
        self.L.orc_po_process.argtypes = [C.c_void_p] + [f32p, C.c_int] * 5 + [f32p, f32p, i32p]
        self.L.orc_po_cloud_size.argtypes = [C.c_void_p, C.c_int]
        self.L.orc_po_cloud_copy.argtypes = [C.c_void_p, C.c_int, f32p]
        self.L.orc_po_matches.argtypes = [C.c_void_p, C.c_int, i32p]
        self.h = self.L.orc_po_create(scan_period, io_ratio, num_max_iterations)

    def __del__(self):
        try:
            self.L.orc_po_destroy(self.h)
        except Exception:
            pass

    def set_enable_odom(self, enable):
        self.L.orc_po_set_enable_odom(self.h, int(bool(enable)))

    def process(self, sharp, less_sharp, flat, less_flat, full):
        args = []
        for c in (sharp, less_sharp, flat, less_flat, full):
            c = np.ascontiguousarray(c, np.float32).reshape(-1, 4)
            args += [c if c.shape[0] else np.zeros((1, 4), np.float32), c.shape[0]]
        ts = np.zeros(7, np.float32); te = np.zeros(7, np.float32); info = np.zeros(4, np.int32)
        self.L.orc_po_process(self.h, *args, ts, te, info)
        return ts, te, dict(iterations=int(info[0]), published=int(info[1]), frame_count=int(info[2]), matches=int(info[3]))

    def cloud(self, which):
        w = {"last_corner": 0, "last_surf": 1, "full": 2, "compact": 3}[which]
        n = self.L.orc_po_cloud_size(self.h, w)
        out = np.zeros((max(n, 1), 4), np.float32)
        if n:
            self.L.orc_po_cloud_copy(self.h, w, out)
        return out[:n]

    def matches(self, kind, n):
        k = 0 if kind == "corner" else 1
        out = np.zeros((max(n, 1), 2 + k), np.int32)
        m = self.L.orc_po_matches(self.h, k, out)
        return out[:m]


def toy_marginalize(nb, bs, bi, bj, W, y, x, use_cauchy=False):
    """Linear two-block factors over nb Euclidean blocks; block 0 is marginalised through the oracle's MarginalizationInfo.
    Returns dict(m, n, kept (block ids in prior order), J, r, A, b)."""
    L = lib()
    W = np.ascontiguousarray(W, np.float64); y = np.ascontiguousarray(y, np.float64); x = np.ascontiguousarray(x, np.float64).copy()
    bi = np.ascontiguousarray(bi, np.int32); bj = np.ascontiguousarray(bj, np.int32)
    nf, _, nr, _ = W.shape
    pos = nb * bs
    mn = np.zeros(2, np.int32); kept = np.full(nb, -1, np.int32)
    J = np.zeros(pos * pos); r = np.zeros(pos); A = np.zeros(pos * pos); b = np.zeros(pos)
    L.orc_toy_marginalize.argtypes = [C.c_int] * 4 + [i32p, i32p, f64p, f64p, f64p, C.c_int, i32p, i32p, f64p, f64p, f64p, f64p]
    nk = L.orc_toy_marginalize(nb, bs, nf, nr, bi, bj, W, y, x, int(use_cauchy), mn, kept, J, r, A, b)
    m, n = int(mn[0]), int(mn[1])
    return dict(m=m, n=n, kept=kept[:nk].copy(), J=J[:n * n].reshape(n, n).copy(), r=r[:n].copy(),
                A=A[:(m + n) ** 2].reshape(m + n, m + n).copy(), b=b[:m + n].copy())
