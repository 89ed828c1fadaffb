"""Voxel-grid time vs number of clouds per call (device-resident scans): does the stage scale down to small chunks?"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lidar-slam-from-scratch_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import torch, oracle_lib, slam_b200, bench
eng = slam_b200.Engine(0)
syn = oracle_lib.Synth()
world = bench.make_world(syn)
F = 512
poses = bench.make_poses(syn, F)
rays = 64 * 1875
d_raw = torch.empty(F * rays * 3, dtype=torch.float64, device="cuda")
off = eng.synth_scans_dev(bench.SENSOR, world, poses, 1000, d_raw.data_ptr())
d_out = torch.empty(F * rays * 3, dtype=torch.float64, device="cuda")
I64 = C.POINTER(C.c_int64)
for n in (512, 512, 256, 128, 64, 32, 16):
    o = np.ascontiguousarray(off[:n + 1], dtype=np.int64)
    oo = np.zeros(n + 1, dtype=np.int64)
    ts = []
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s = eng.lib.sb_voxel_downsample_batch_dev(eng.h, C.c_void_p(d_raw.data_ptr()), o.ctypes.data_as(I64), n, 0.5,
                                                  C.c_void_p(d_out.data_ptr()), oo.ctypes.data_as(I64), None)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        assert s == 0
    print(f"clouds {n:4d}: {min(ts[1:]):7.3f} ms  ({min(ts[1:]) / n * 1e3:6.1f} us per cloud), voxels {oo[-1] / n:.0f} per cloud, path {eng.last_voxel_path}")
