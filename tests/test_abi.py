"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/slam_b200.h declares,
the ctypes mirror has the C layout, and the library refuses to work without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

import oracle_lib  # noqa: F401  (sys.path)
import slam_b200

ROOT = oracle_lib.ROOT
HEADER = os.path.join(ROOT, "include", "slam_b200.h")
LIB = os.path.join(ROOT, "lidar-slam-from-scratch_b200", "libslam_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", src)))


def test_library_built():
    assert os.path.exists(LIB), "run `make -C lidar-slam-from-scratch_b200` (or __graft_entry__.build())"


def test_exports_every_declared_symbol():
    lib = C.CDLL(LIB)
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/slam_b200.h but not exported"
    assert set(names) == set(slam_b200.SYMBOLS), set(names) ^ set(slam_b200.SYMBOLS)


def test_ctypes_layout_matches_header():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "slam_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(sb_icp_config), sizeof(sb_icp_result), sizeof(sb_loop_config),
         sizeof(sb_loop_result), offsetof(sb_icp_result, error_history), offsetof(sb_loop_config, icp_tolerance),
         offsetof(sb_loop_result, transform));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        got = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [C.sizeof(slam_b200.ICPConfigC), C.sizeof(slam_b200.ICPResultC), C.sizeof(slam_b200.LoopConfigC),
            C.sizeof(slam_b200.LoopResultC), slam_b200.ICPResultC.error_history.offset,
            slam_b200.LoopConfigC.icp_tolerance.offset, slam_b200.LoopResultC.transform.offset]
    assert got == want


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = slam_b200.load_library()
    assert lib.sb_version().decode().startswith("slam_b200")
    with pytest.raises(slam_b200.SlamB200Error) as e:
        slam_b200.Engine(0)
    assert e.value.status == 4  # SB_ERR_NO_DEVICE


def test_defaults_match_reference():
    lib = slam_b200.load_library()
    cfg = slam_b200.ICPConfigC()
    lib.sb_default_icp_config(C.byref(cfg))
    assert (cfg.max_iterations, cfg.tolerance, cfg.min_error, cfg.normals_k) == (50, 1e-6, 1e-9, 20)  # types.hpp:143-148
    assert list(cfg.initial_transform) == [1.0 if i % 5 == 0 else 0.0 for i in range(16)]
    lc = slam_b200.LoopConfigC()
    lib.sb_default_loop_config(C.byref(lc))
    assert (lc.frame_gap, lc.sc_distance_threshold, lc.icp_fitness_threshold, lc.max_candidates) == (50, 0.25, 0.3, 3)
    assert (lc.icp_max_iterations, lc.icp_tolerance) == (30, 1e-6)  # loop_closure.hpp:106-107


def test_odometry_pose_chain_host_only():
    """sb_odometry_poses is host arithmetic (slam_node.cpp:139-145) and needs no device: identity for the frames whose
    registration did not converge or ended above the error limit, otherwise pose <- pose * delta."""
    import numpy as np
    lib = slam_b200.load_library()
    rng = np.random.default_rng(4)
    n = 9
    rec = np.zeros(n, dtype=slam_b200.ICP_DTYPE)
    Ts = []
    for i in range(n):
        a = rng.uniform(-0.05, 0.05)
        T = np.eye(4)
        T[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
        T[:3, 3] = rng.uniform(-1, 1, 3)
        Ts.append(T)
        rec[i]["transformation"] = T.reshape(-1)
        rec[i]["converged"] = 0 if i == 3 else 1
        rec[i]["final_error"] = 1.5 if i == 5 else (1.0 if i == 6 else 0.2)   # > 1.0 is dropped, == 1.0 is kept
        rec[i]["status"] = 0
    out = np.empty((n + 1, 16))
    D = C.POINTER(C.c_double)
    s = lib.sb_odometry_poses(None, rec.ctypes.data_as(C.POINTER(slam_b200.ICPResultC)), n, 1.0, None, out.ctypes.data_as(D))
    assert s == 0
    pose = np.eye(4)
    assert np.array_equal(out[0].reshape(4, 4), pose)
    for i in range(n):
        pose = pose @ (np.eye(4) if i in (3, 5) else Ts[i])
        assert np.allclose(out[i + 1].reshape(4, 4), pose, rtol=0, atol=1e-15)
    p0 = Ts[0].reshape(-1).copy()
    assert lib.sb_odometry_poses(None, None, 0, 1.0, p0.ctypes.data_as(D), out.ctypes.data_as(D)) == 0
    assert np.array_equal(out[0], p0)
    assert lib.sb_odometry_poses(None, None, 3, 1.0, None, out.ctypes.data_as(D)) != 0   # results missing


def test_pose_graph_factor_hand_off_host_only():
    """sb_odometry_factors / sb_loop_factors are host arithmetic: the arguments process_frame passes
    to PoseGraph::addOdometryFactor (slam_node.cpp:139-145; noise scale 1 + 10 * fitness, pose_graph.cpp:88) and to
    addLoopClosure(match, query, transform) (slam_node.cpp:163-167)."""
    import numpy as np
    lib = slam_b200.load_library()
    rng = np.random.default_rng(8)
    n = 6
    rec = np.zeros(n, dtype=slam_b200.ICP_DTYPE)
    for i in range(n):
        T = np.eye(4)
        T[:3, 3] = rng.uniform(-1, 1, 3)
        rec[i]["transformation"] = T.reshape(-1)
        rec[i]["converged"] = 0 if i == 2 else 1
        rec[i]["final_error"] = 1.25 if i == 4 else 0.1 * (i + 1)
    out = np.zeros(n, dtype=slam_b200.FACTOR_DTYPE)
    F = C.POINTER(slam_b200.PoseFactorC)
    assert lib.sb_odometry_factors(None, rec.ctypes.data_as(C.POINTER(slam_b200.ICPResultC)), n, 10, 1.0,
                                   out.ctypes.data_as(F)) == 0
    for i in range(n):
        assert (out[i]["kind"], out[i]["from"], out[i]["to"]) == (0, 10 + i, 11 + i)
        want = np.eye(4) if i in (2, 4) else rec[i]["transformation"].reshape(4, 4)
        assert np.array_equal(out[i]["relative"], want)
        assert out[i]["fitness"] == rec[i]["final_error"]          # passed on even when the delta is replaced
        assert out[i]["noise_scale"] == 1.0 + rec[i]["final_error"] * 10.0
    loops = (slam_b200.LoopResultC * 2)()
    for j, (q, m) in enumerate([(120, 7), (130, 12)]):
        loops[j].query_frame, loops[j].match_frame, loops[j].icp_fitness = q, m, 0.05 * (j + 1)
        loops[j].transform[:] = list(np.arange(16, dtype=float) + j)
    lout = np.zeros(2, dtype=slam_b200.FACTOR_DTYPE)
    assert lib.sb_loop_factors(None, loops, 2, lout.ctypes.data_as(F)) == 0
    assert [(int(r["kind"]), int(r["from"]), int(r["to"])) for r in lout] == [(1, 7, 120), (1, 12, 130)]
    assert np.array_equal(lout[1]["relative"].reshape(-1), np.arange(16, dtype=float) + 1)
    assert list(lout["noise_scale"]) == [1.0, 1.0]
    assert lib.sb_odometry_factors(None, None, 1, 0, 1.0, None) != 0   # invalid arguments are refused
