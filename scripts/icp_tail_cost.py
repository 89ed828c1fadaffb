"""How much of the batch ICP time is the tail (few pairs iterating long after the rest converged)?"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lidar-slam-from-scratch_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import torch, oracle_lib, slam_b200, bench
eng = slam_b200.Engine(0)
syn = oracle_lib.Synth()
world = bench.make_world(syn)
F = 1000
poses = bench.make_poses(syn, F + 1)
d_raw = torch.empty((F + 1) * 64 * 1875 * 3, dtype=torch.float64, device="cuda")
off = eng.synth_scans_dev(bench.SENSOR, world, poses, 1000, d_raw.data_ptr())
src, tgt = np.arange(1, F + 1, dtype=np.int32), np.arange(0, F, dtype=np.int32)
for mi in (50, 50, 12, 9, 6, 3, 1, 0):
    cfg = eng.icp_config(max_iterations=mi)
    eng.set_profiling(True)
    for _ in range(2):
        res = eng.register_batch(None, off, src, tgt, voxel=0.5, cfg=cfg, device_ptr=d_raw.data_ptr())
    st = eng.stage_ms()
    print(f"max_iterations {mi:3d}: icp_loop {st['icp_loop']:7.3f} ms, mean iterations {res.num_iterations.mean():.2f}, converged {res.converged.mean():.3f}")
