"""Generates tests/golden/reference_small.npz: outputs of the REFERENCE'S OWN SOURCES (oracle/_ref/libslam_ref.so =
/root/reference/slam_viz compiled unmodified against oracle/eigen_standin, oracle/build_ref.sh) on small seeded inputs.

The library can only be built where /root/reference exists, so its outputs are committed here and the tests that read
this file (tests/test_oracle.py::test_reference_golden_*, tests/test_gpu_parity.py::test_reference_golden_gpu) run
anywhere, including on the GPU box.  Run from the repo root:
    python tests/golden/make_reference_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # the synthetic raycaster only (input generation)
import ref_lib


def sort_rows(x):
    return x[np.lexsort((x[:, 2], x[:, 1], x[:, 0]))]


def main():
    ref = ref_lib.Reference()
    syn = oracle_lib.Synth()
    scene = syn.scene(1, n_boxes=400)
    s = oracle_lib.small_sensor(24, 480)
    raw_a = syn.scan(s, scene, (0.0, 0.0, 0.0), 21).astype(np.float32)      # float32 records, as on disk
    raw_b = syn.scan(s, scene, (0.9, 0.08, 0.008), 22).astype(np.float32)
    out = {"raw_a": raw_a, "raw_b": raw_b}
    a = sort_rows(ref.voxel_downsample(raw_a.astype(np.float64), 0.5))       # rows sorted: the reference's order is
    b = sort_rows(ref.voxel_downsample(raw_b.astype(np.float64), 0.5))       # unordered_map iteration order
    out["voxel_a"], out["voxel_b"] = a, b
    out["voxel_a_02"] = sort_rows(ref.voxel_downsample(raw_a.astype(np.float64), 0.2))
    ta = ref.tree(a)
    out["knn20_a"] = ta.k_nearest_batch(a, 20)
    out["knn10_a"] = ta.k_nearest_batch(a, 10)
    idx, d2 = ta.nearest_batch(b)
    out["nn_b_in_a"], out["nn_b_in_a_d2"] = idx, d2
    out["normals20_a"] = ta.estimate_normals(20)
    out["normals10_a"] = ta.estimate_normals(10)
    out["sc_a"], out["sc_b"] = ref.sc_compute(a), ref.sc_compute(b)
    out["sc_dist_ab"] = np.float64(ref.sc_distance_clouds(a, b))
    r, k = ref.sc_keys(a)
    out["sc_ring_key_a"], out["sc_sector_key_a"] = r, k
    nrm = out["normals20_a"]
    out["solve_T"] = ref.solve_point_to_plane(b, a[idx], nrm[idx])
    for name, cfg in (("icp50", dict()), ("icp3", dict(max_iterations=3)), ("icp30", dict(max_iterations=30))):
        res = ref.icp_point_to_plane(b, a, **cfg)
        out[name + "_T"] = res["transformation"]
        out[name + "_history"] = res["error_history"]
        out[name + "_meta"] = np.array([res["num_iterations"], int(res["converged"])], dtype=np.int32)
        out[name + "_final_error"] = np.float64(res["final_error"])
    # a short loop-closure sequence (loop_closure.hpp): out and back along x
    s2 = oracle_lib.small_sensor(16, 360)
    poses = [(float(i), 0.0, 0.0) for i in range(7)] + [(0.4, 0.05, 0.0), (1.1, -0.05, 0.01)]
    det = ref.loop(frame_gap=3, sc_thr=0.5, icp_thr=0.5, max_candidates=2)
    clouds, offs, found = [], [0], []
    for i, p in enumerate(poses):
        c = sort_rows(ref.voxel_downsample(syn.scan(s2, scene, p, 50 + i).astype(np.float32).astype(np.float64), 0.5))
        clouds.append(c)
        offs.append(offs[-1] + len(c))
        det.add(c, i)
        for x in det.detect():
            found.append(np.concatenate([[x["query_frame"], x["match_frame"], x["scan_context_distance"],
                                          x["icp_fitness"]], x["transform"].reshape(-1)]))
    out["loop_clouds"] = np.concatenate(clouds)
    out["loop_offsets"] = np.array(offs, dtype=np.int64)
    out["loop_results"] = np.array(found)  # rows: query, match, sc_distance, icp_fitness, T[16]
    # config C1 itself (BASELINE.json configs[0]): 64 beams x 1875 steps, poses (0,0,0) and (1.0, 0.1, 0.01), voxel 0.5,
    # ICPConfig defaults.  The raw scans (2 x 1.4 MB) are not stored: the raycaster is deterministic, the fixture keeps
    # their SHA-256 so that a test knows it regenerated the same input, and the reference's outputs.
    import hashlib
    c1a = syn.scan(oracle_lib.SENSOR64, scene, (0.0, 0.0, 0.0), 7)
    c1b = syn.scan(oracle_lib.SENSOR64, scene, (1.0, 0.1, 0.01), 8)
    out["c1_sha256"] = np.frombuffer(hashlib.sha256(c1a.tobytes() + c1b.tobytes()).digest(), dtype=np.uint8)
    va = sort_rows(ref.voxel_downsample(c1a, 0.5))
    vb = sort_rows(ref.voxel_downsample(c1b, 0.5))
    out["c1_voxel_counts"] = np.array([len(c1a), len(c1b), len(va), len(vb)], dtype=np.int64)
    out["c1_voxel_a_sum"], out["c1_voxel_b_sum"] = va.sum(axis=0), vb.sum(axis=0)
    c1 = ref.icp_point_to_plane(vb, va)
    out["c1_T"], out["c1_history"] = c1["transformation"], c1["error_history"]
    out["c1_meta"] = np.array([c1["num_iterations"], int(c1["converged"])], dtype=np.int32)
    out["c1_final_error"] = np.float64(c1["final_error"])
    path = os.path.join(HERE, "reference_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(a), "and", len(b), "voxel points;",
          len(found), "loop results; icp iterations", out["icp50_meta"][0])


if __name__ == "__main__":
    main()
