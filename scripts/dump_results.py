"""Registers a small deterministic batch (64 consecutive frame pairs of the C2 sequence) and writes every result — poses,
error histories, iteration counts, and the normals of one cloud — to an .npz, so that two builds of the library
(SB_LIB_PATH) can be compared bit for bit:  python scripts/dump_results.py out.npz ;  python scripts/dump_results.py a.npz b.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lidar-slam-from-scratch_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

if len(sys.argv) == 3:
    a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
    bad = [k for k in a.files if not np.array_equal(a[k], b[k], equal_nan=True)]
    for k in bad:
        print(k, "differs: max abs diff", float(np.nanmax(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)))))
    print("identical" if not bad else f"{len(bad)} of {len(a.files)} arrays differ")
    sys.exit(1 if bad else 0)

import torch
import bench
import oracle_lib
import slam_b200

syn = oracle_lib.Synth()
world = bench.make_world(syn)
F = 64
poses = bench.make_poses(syn, F + 1)
eng = slam_b200.Engine(0)
rays = bench.SENSOR["beams"] * bench.SENSOR["azimuth_steps"]
d_raw = torch.empty((F + 1) * rays * 3, dtype=torch.float64, device="cuda")
off = eng.synth_scans_dev(bench.SENSOR, world, poses, 1000, d_raw.data_ptr())
src, tgt = np.arange(1, F + 1, dtype=np.int32), np.arange(0, F, dtype=np.int32)
res = eng.register_batch(None, off, src, tgt, voxel=0.5, device_ptr=d_raw.data_ptr())
h = d_raw[:int(off[1]) * 3].cpu().numpy().reshape(-1, 3)
ds = eng.voxel_downsample(h, 0.5)
ix = eng.index_build(ds) if hasattr(eng, "index_build") else None
out = {"T": np.stack([r.transformation for r in res]), "iters": np.array([r.num_iterations for r in res]),
       "conv": np.array([r.converged for r in res]), "err": np.array([r.final_error for r in res]),
       "hist": np.stack([np.pad(np.asarray(r.error_history, dtype=np.float64), (0, 64 - len(r.error_history))) for r in res]),
       "voxel": ds}
single = eng.icp_point_to_plane(eng.voxel_downsample(d_raw[int(off[1]) * 3:int(off[2]) * 3].cpu().numpy().reshape(-1, 3), 0.5), ds)
out["single_T"] = single.transformation
out["single_hist"] = np.asarray(single.error_history, dtype=np.float64)
np.savez(sys.argv[1], **out)
print("wrote", sys.argv[1], "mean iterations", out["iters"].mean())
