"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun).  Builds the C4 loop-closure database sharded over the
ranks (sb_loop_create(rank, world)), runs detect() through python/sharding.py (NCCL all-gathers) at several query
frames, and rank 0 writes what was found as JSON.  With world == 1 it also records the unsharded sb_loop_detect."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "lidar-slam-from-scratch_b200", "python"))


def main():
    import torch
    import torch.distributed as dist
    import bench
    import oracle_lib
    import sharding
    import slam_b200

    out_path, n_db = sys.argv[1], int(sys.argv[2])
    queries = [int(q) for q in sys.argv[3].split(",")]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = slam_b200.Engine(local)
    syn = oracle_lib.Synth()
    sensor = oracle_lib.small_sensor(32, 600)          # a reduced sensor keeps the 4000-frame build short
    scene = bench.make_world(syn)
    poses = bench.make_poses(syn, n_db)
    rays = sensor["beams"] * sensor["azimuth_steps"]
    chunk = 500
    d_raw = torch.empty(chunk * rays * 3, dtype=torch.float64, device="cuda")

    def detector(r, w, thr):
        return slam_b200.LoopClosureDetector(eng, frame_gap=50, sc_distance_threshold=thr, icp_fitness_threshold=0.3,
                                             max_candidates=10, rank=r, world=w)

    det = detector(rank, world, 1e300)                 # threshold "infinity": exactly the top 10 are verified
    plain = detector(0, 1, 1e300) if world == 1 else None
    found = {}
    for c0 in range(0, n_db, chunk):
        c1 = min(n_db, c0 + chunk)
        off = eng.synth_scans_dev(sensor, scene, poses[c0:c1], 7000 + c0, d_raw.data_ptr())
        h = d_raw[:int(off[-1]) * 3].cpu().numpy().reshape(-1, 3)
        ds, doff = eng.voxel_downsample_batch(h, off, 0.5)
        for f in range(c0, c1):
            cloud = ds[doff[f - c0]:doff[f - c0 + 1]]
            det.addFrame(cloud, f)
            if plain is not None:
                plain.addFrame(cloud, f)
            if f in queries:
                md, me, acc, rec = sharding.sharded_detect(det, rank, world, top_k=10)
                item = {"dist": [float(x) for x in md], "entries": [int(x) for x in me], "accepted": [int(x) for x in acc],
                        "fitness": [float(x) for x in rec[:len(me), 2]], "converged": [int(x > 0.5) for x in rec[:len(me), 1]],
                        "transforms": rec[:len(me), 4:].tolist()}
                if plain is not None:   # the unsharded entry point on the same database
                    item["plain"] = [{"match_frame": r["match_frame"], "fitness": r["icp_fitness"],
                                      "transform": np.asarray(r["transform"]).reshape(-1).tolist()} for r in plain.detect(64)]
                found[str(f)] = item
    if rank == 0:
        json.dump({"world": world, "n_db": n_db, "found": found}, open(out_path, "w"))
    det.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
