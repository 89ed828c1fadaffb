"""Test-side loaders: the CPU oracle (oracle/liboracle.so) and the synthetic LiDAR raycaster (synth/libsynth.so).

TEST INFRASTRUCTURE ONLY.  The product (lidar-slam-from-scratch_b200/) never imports this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lidar-slam-from-scratch_b200", "python"))

_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)
_LL = C.POINTER(C.c_longlong)


def _dp(a):
    return a.ctypes.data_as(_D)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 3)


def _build(dirname, so):
    path = os.path.join(ROOT, dirname, so)
    src_newer = False
    if os.path.exists(path):
        mt = os.path.getmtime(path)
        for f in os.listdir(os.path.join(ROOT, dirname)):
            if f.endswith((".cpp", ".h")) and os.path.getmtime(os.path.join(ROOT, dirname, f)) > mt:
                src_newer = True
    if not os.path.exists(path) or src_newer:
        subprocess.check_call(["sh", os.path.join(ROOT, dirname, "build.sh")])
    return path


class Oracle:
    def __init__(self):
        lib = C.CDLL(_build("oracle", "liboracle.so"))
        lib.orc_voxel_downsample.restype = C.c_longlong
        lib.orc_voxel_downsample.argtypes = [_D, C.c_longlong, C.c_double, _D, _LL]
        lib.orc_kdtree_build.restype = C.c_void_p
        lib.orc_kdtree_build.argtypes = [_D, C.c_int]
        lib.orc_kdtree_free.argtypes = [C.c_void_p]
        lib.orc_kdtree_nearest_batch.argtypes = [C.c_void_p, _D, C.c_int, _I, _D]
        lib.orc_kdtree_k_nearest_batch.argtypes = [C.c_void_p, _D, C.c_int, C.c_int, _I, _D]
        lib.orc_brute_knn.argtypes = [_D, C.c_int, _D, C.c_int, C.c_int, _I, _D]
        lib.orc_estimate_normals.argtypes = [C.c_void_p, _D, C.c_int, C.c_int, _D, _D]
        lib.orc_jacobi3.argtypes = [_D, _D, _D]
        lib.orc_ldlt6_solve.argtypes = [_D, _D, _D]
        lib.orc_solve_point_to_plane.argtypes = [_D, _D, _D, C.c_int, _D]
        lib.orc_icp_point_to_plane.restype = C.c_int
        lib.orc_icp_point_to_plane.argtypes = [_D, C.c_int, _D, C.c_int, C.c_int, C.c_double, C.c_double, _D, C.c_int,
                                               C.c_int, _D, _I, _I, _D, _D]
        lib.orc_transform_cloud.argtypes = [_D, C.c_longlong, _D, _D]
        lib.orc_occupancy_cells.restype = C.c_longlong
        lib.orc_occupancy_cells.argtypes = [_D, _LL, C.c_int, _D, C.c_double, C.c_double, C.c_double, C.c_double, _I,
                                            C.c_longlong]
        lib.orc_sc_compute.argtypes = [_D, C.c_longlong, _D]
        lib.orc_sc_distance.restype = C.c_double
        lib.orc_sc_distance.argtypes = [_D, _D]
        lib.orc_loop_create.restype = C.c_void_p
        lib.orc_loop_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int]
        lib.orc_loop_free.argtypes = [C.c_void_p]
        lib.orc_loop_add.argtypes = [C.c_void_p, _D, C.c_int, C.c_int]
        lib.orc_loop_size.argtypes = [C.c_void_p]
        lib.orc_loop_candidates.restype = C.c_int
        lib.orc_loop_candidates.argtypes = [C.c_void_p, C.c_int, _D, _I]
        lib.orc_loop_detect.restype = C.c_int
        lib.orc_loop_detect.argtypes = [C.c_void_p, C.c_int, _I, _D, _D, _D]
        self.lib = lib

    def voxel_downsample(self, pts, voxel):
        p = _f64(pts)
        n = p.shape[0]
        out = np.empty((max(n, 1), 3))
        keys = np.empty((max(n, 1), 3), dtype=np.int64)
        m = self.lib.orc_voxel_downsample(_dp(p), n, float(voxel), _dp(out), keys.ctypes.data_as(_LL))
        return out[:m].copy(), keys[:m].copy()

    class Tree:
        def __init__(self, lib, pts):
            self.lib = lib
            self.p = _f64(pts)
            self.h = lib.orc_kdtree_build(_dp(self.p), self.p.shape[0])

        def __del__(self):
            if getattr(self, "h", None):
                self.lib.orc_kdtree_free(self.h)
                self.h = None

        def nearest_batch(self, q):
            q = _f64(q)
            idx = np.empty(max(q.shape[0], 1), dtype=np.int32)
            d2 = np.empty(max(q.shape[0], 1))
            self.lib.orc_kdtree_nearest_batch(self.h, _dp(q), q.shape[0], idx.ctypes.data_as(_I), _dp(d2))
            return idx[:q.shape[0]], d2[:q.shape[0]]

        def k_nearest_batch(self, q, k):
            q = _f64(q)
            idx = np.empty((max(q.shape[0], 1), k), dtype=np.int32)
            d2 = np.empty((max(q.shape[0], 1), k))
            self.lib.orc_kdtree_k_nearest_batch(self.h, _dp(q), q.shape[0], k, idx.ctypes.data_as(_I), _dp(d2))
            return idx[:q.shape[0]], d2[:q.shape[0]]

        def estimate_normals(self, k):
            n = self.p.shape[0]
            nrm = np.empty((max(n, 1), 3))
            ev = np.empty((max(n, 1), 3))
            self.lib.orc_estimate_normals(self.h, _dp(self.p), n, k, _dp(nrm), _dp(ev))
            return nrm[:n], ev[:n]

    def tree(self, pts):
        return Oracle.Tree(self.lib, pts)

    def brute_knn(self, pts, q, k):
        p, q = _f64(pts), _f64(q)
        idx = np.empty((max(q.shape[0], 1), k), dtype=np.int32)
        d2 = np.empty((max(q.shape[0], 1), k))
        self.lib.orc_brute_knn(_dp(p), p.shape[0], _dp(q), q.shape[0], k, idx.ctypes.data_as(_I), _dp(d2))
        return idx[:q.shape[0]], d2[:q.shape[0]]

    def jacobi3(self, A):
        A = np.ascontiguousarray(A, dtype=np.float64).reshape(9)
        w, V = np.empty(3), np.empty(9)
        self.lib.orc_jacobi3(_dp(A), _dp(w), _dp(V))
        return w, V.reshape(3, 3)

    def ldlt6_solve(self, A, b):
        A = np.ascontiguousarray(A, dtype=np.float64).reshape(36)
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(6)
        x = np.empty(6)
        self.lib.orc_ldlt6_solve(_dp(A), _dp(b), _dp(x))
        return x

    def solve_point_to_plane(self, src, tgt, nrm):
        s, t, n = _f64(src), _f64(tgt), _f64(nrm)
        T = np.empty(16)
        self.lib.orc_solve_point_to_plane(_dp(s), _dp(t), _dp(n), s.shape[0], _dp(T))
        return T.reshape(4, 4)

    def icp_point_to_plane(self, src, tgt, max_iterations=50, tolerance=1e-6, min_error=1e-9, T0=None, normals_k=20,
                           faithful_cost=0):
        s, t = _f64(src), _f64(tgt)
        T = np.empty(16)
        conv, nit = C.c_int(0), C.c_int(0)
        fe = C.c_double(0)
        hist = np.empty(max_iterations + 2)
        t0 = None if T0 is None else np.ascontiguousarray(T0, dtype=np.float64).reshape(16)
        hl = self.lib.orc_icp_point_to_plane(_dp(s), s.shape[0], _dp(t), t.shape[0], max_iterations, tolerance,
                                             min_error, _dp(t0) if t0 is not None else None, normals_k, faithful_cost,
                                             _dp(T), C.byref(conv), C.byref(nit), C.byref(fe), _dp(hist))
        return dict(transformation=T.reshape(4, 4), converged=bool(conv.value), num_iterations=nit.value,
                    final_error=fe.value, error_history=hist[:hl].copy())

    def transform_cloud(self, pts, T):
        p, T = _f64(pts), np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        out = np.empty_like(p)
        self.lib.orc_transform_cloud(_dp(p), p.shape[0], _dp(T), _dp(out))
        return out

    def occupancy_cells(self, pts, offsets, poses, res=0.2, hmin=0.3, hmax=2.0, max_range=40.0):
        p = _f64(pts)
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        T = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 16)
        cells = np.empty((max(p.shape[0], 1), 2), dtype=np.int32)
        n = self.lib.orc_occupancy_cells(_dp(p), off.ctypes.data_as(_LL), off.shape[0] - 1, _dp(T), res, hmin, hmax,
                                         max_range, cells.ctypes.data_as(_I), cells.shape[0])
        return cells[:n].copy()

    def sc_compute(self, pts):
        p = _f64(pts)
        d = np.empty(1200)
        self.lib.orc_sc_compute(_dp(p), p.shape[0], _dp(d))
        return d

    def sc_distance(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
        return float(self.lib.orc_sc_distance(_dp(a), _dp(b)))

    class Loop:
        def __init__(self, lib, frame_gap, sc_thr, icp_thr, max_candidates):
            self.lib = lib
            self.h = lib.orc_loop_create(frame_gap, sc_thr, icp_thr, max_candidates)

        def __del__(self):
            if getattr(self, "h", None):
                self.lib.orc_loop_free(self.h)
                self.h = None

        def add(self, pts, frame_idx):
            p = _f64(pts)
            self.lib.orc_loop_add(self.h, _dp(p), p.shape[0], frame_idx)

        def candidates(self, cap=4096):
            dist = np.empty(cap)
            idx = np.empty(cap, dtype=np.int32)
            m = self.lib.orc_loop_candidates(self.h, cap, _dp(dist), idx.ctypes.data_as(_I))
            m = min(m, cap)
            return dist[:m].copy(), idx[:m].copy()

        def detect(self, cap=64):
            fr = np.empty(2 * cap, dtype=np.int32)
            T = np.empty(16 * cap)
            sd, fit = np.empty(cap), np.empty(cap)
            m = min(self.lib.orc_loop_detect(self.h, cap, fr.ctypes.data_as(_I), _dp(T), _dp(sd), _dp(fit)), cap)
            return [dict(query_frame=int(fr[2 * i]), match_frame=int(fr[2 * i + 1]),
                         transform=T[16 * i:16 * i + 16].reshape(4, 4).copy(), scan_context_distance=float(sd[i]),
                         icp_fitness=float(fit[i])) for i in range(m)]

    def loop(self, frame_gap=50, sc_thr=0.25, icp_thr=0.3, max_candidates=3):
        return Oracle.Loop(self.lib, frame_gap, sc_thr, icp_thr, max_candidates)


SENSOR64 = dict(beams=64, azimuth_steps=1875, elev_top_deg=2.0, elev_bot_deg=-24.8, max_range=120.0, noise_sigma=0.02,
                sensor_height=1.73)
SENSOR128 = dict(beams=128, azimuth_steps=2048, elev_top_deg=2.0, elev_bot_deg=-24.8, max_range=120.0,
                 noise_sigma=0.02, sensor_height=1.73)


class Synth:
    def __init__(self):
        lib = C.CDLL(_build("synth", "libsynth.so"))
        lib.syn_scene.restype = C.c_int
        lib.syn_scene.argtypes = [C.c_ulonglong, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float,
                                  C.POINTER(C.c_float)]
        lib.syn_pose.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _D]
        lib.syn_scan.restype = C.c_longlong
        lib.syn_scan.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.POINTER(C.c_float), C.c_int, C.c_double, C.c_double, C.c_double, C.c_ulonglong, _D,
                                 C.c_int]
        self.lib = lib

    def scene(self, seed, n_boxes=400, half_extent=90.0, path_kind=0, radius=0.0, corridor_half=6.0,
              sensor_height=1.73):
        b = np.zeros((n_boxes, 6), dtype=np.float32)
        m = self.lib.syn_scene(seed, n_boxes, half_extent, path_kind, radius, corridor_half, sensor_height,
                               b.ctypes.data_as(C.POINTER(C.c_float)))
        return b[:m].copy()

    def pose(self, path_kind, radius, arc_len, lateral=0.0, dyaw=0.0):
        p = np.zeros(3)
        self.lib.syn_pose(path_kind, radius, arc_len, lateral, dyaw, _dp(p))
        return p

    def scan(self, sensor, boxes, pose, noise_seed, threads=8):
        b = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 6)
        rays = sensor["beams"] * sensor["azimuth_steps"]
        out = np.empty((rays, 3))
        m = self.lib.syn_scan(sensor["beams"], sensor["azimuth_steps"], sensor["elev_top_deg"], sensor["elev_bot_deg"],
                              sensor["max_range"], sensor["noise_sigma"], sensor["sensor_height"],
                              b.ctypes.data_as(C.POINTER(C.c_float)), b.shape[0], float(pose[0]), float(pose[1]),
                              float(pose[2]), int(noise_seed), _dp(out), threads)
        return out[:m].copy()


def small_sensor(beams=32, az=450):
    """A reduced sensor for fast CPU-side parity cases (same FOV as SENSOR64)."""
    s = dict(SENSOR64)
    s["beams"] = beams
    s["azimuth_steps"] = az
    return s
