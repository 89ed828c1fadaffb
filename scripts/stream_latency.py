"""Streaming odometry latency: the reference's SlamNode::process_frame loop (slam_node.cpp:118-175), one frame at a
time through the single-call entry points (voxel_downsample, icp_point_to_plane, LoopClosureDetector::addFrame/detect,
world transform), host buffers in and out.  Prints one JSON line with ms/frame mean / p50 / p99 and the split per call.

  python scripts/stream_latency.py [--frames 120]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lidar-slam-from-scratch_b200", "python"))


def process_frames(eng, slam_b200, scans, voxel=0.5, skip=10):
    """scans: list of (n_i, 3) float64 host arrays.  Frames up to `skip` are warm-up (allocations, graph build)."""
    cfg = eng.icp_config()
    det = slam_b200.LoopClosureDetector(eng)
    split = {"voxel": [], "icp": [], "world": [], "add_frame": [], "detect": []}
    total, frames = [], []
    prev = eng.voxel_downsample(scans[0], voxel)          # slam_node.cpp:69-70 (first frame)
    det.addFrame(prev, 0)
    pose = np.eye(4)
    launches0 = eng.launch_count
    for i in range(1, len(scans)):
        t0 = time.perf_counter()
        cur = eng.voxel_downsample(scans[i], voxel)       # :122
        t1 = time.perf_counter()
        r = eng.icp_point_to_plane(cur, prev, cfg)        # :132-138
        t2 = time.perf_counter()
        delta = np.eye(4) if (not r.converged or r.final_error > 1.0) else r.transformation  # :139-140
        pose = pose @ delta
        eng.transform_clouds(cur, np.array([0, len(cur)], dtype=np.int64), pose[None])        # :147
        t3 = time.perf_counter()
        det.addFrame(cur, i)                              # :159
        t4 = time.perf_counter()
        hit = i % 10 == 0 and i > 50                      # :160
        if hit:
            det.detect()
        t5 = time.perf_counter()
        prev = cur
        if i > skip:
            split["voxel"].append(t1 - t0)
            split["icp"].append(t2 - t1)
            split["world"].append(t3 - t2)
            split["add_frame"].append(t4 - t3)
            if hit:
                split["detect"].append(t5 - t4)
            total.append(t5 - t0)
            frames.append((t5 - t0, i, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, r.num_iterations))
    launches = eng.launch_count - launches0
    det.close()
    t = np.array(total) * 1e3
    return {"workload": "process_frame loop (slam_node.cpp:118-175), one frame per call, host buffers in and out",
            "frames": len(t), "ms_per_frame_mean": float(t.mean()), "p50": float(np.percentile(t, 50)),
            "p99": float(np.percentile(t, 99)), "max": float(t.max()),
            "split_ms_mean": {k: float(np.mean(v) * 1e3) if v else None for k, v in split.items()},
            "slowest_frames": [{"frame": f[1], "ms": round(f[0] * 1e3, 2), "voxel": round(f[2] * 1e3, 2),
                                "icp": round(f[3] * 1e3, 2), "world": round(f[4] * 1e3, 2),
                                "add_frame": round(f[5] * 1e3, 2), "detect": round(f[6] * 1e3, 2), "iterations": f[7]}
                               for f in sorted(frames, reverse=True)[:3]],
            "launches_per_frame": launches / (len(scans) - 1), "voxel_points_last": int(len(prev)),
            "raw_points_last": int(len(scans[-1]))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--voxel", type=float, default=0.5)
    args = ap.parse_args()
    import torch
    import bench
    import slam_b200
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    syn = oracle_lib.Synth()  # pose / scene tables only (host side of the raycaster), not the oracle
    world = bench.make_world(syn)
    F = args.frames
    poses = bench.make_poses(syn, F + 1)
    eng = slam_b200.Engine(0)
    rays = bench.SENSOR["beams"] * bench.SENSOR["azimuth_steps"]
    d_raw = torch.empty((F + 1) * rays * 3, dtype=torch.float64, device="cuda")
    off = eng.synth_scans_dev(bench.SENSOR, world, poses, 1000, d_raw.data_ptr())
    h = d_raw[:int(off[-1]) * 3].cpu().numpy().reshape(-1, 3)
    scans = [np.ascontiguousarray(h[off[i]:off[i + 1]]) for i in range(F + 1)]
    print(json.dumps(process_frames(eng, slam_b200, scans, args.voxel)))


if __name__ == "__main__":
    main()
