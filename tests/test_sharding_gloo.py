"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: round-robin sharding of independent scan pairs with one
all-gather of result records, and the sharded Scan Context candidate search + merge.  The per-rank compute is the CPU
oracle here (the checker standing in for a GPU rank); on the GPU box the same functions run over NCCL (bench.py)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

import oracle_lib  # noqa: F401

ROOT = oracle_lib.ROOT


class _Res:
    def __init__(self, d):
        self.transformation = d["transformation"]
        self.final_error = d["final_error"]
        self.num_iterations = d["num_iterations"]
        self.converged = d["converged"]
        self.status = 0


def _make_clouds(n):
    orc, syn = oracle_lib.Oracle(), oracle_lib.Synth()
    scene = syn.scene(1, n_boxes=400)
    s = oracle_lib.small_sensor(8, 180)
    return orc, [orc.voxel_downsample(syn.scan(s, scene, (0.7 * (i % 4), 0.0, 0.0), 80 + i), 0.5)[0] for i in range(n)]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "lidar-slam-from-scratch_b200", "python"))
    import sharding
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    orc, clouds = _make_clouds(7)
    # --- batched pair ICP: pair p = (p+1 -> p), sharded round-robin, one all-gather
    n_pairs = 6
    mine = sharding.shard_units(n_pairs, rank, world)
    local = sharding.pack_results([_Res(orc.icp_point_to_plane(clouds[p + 1], clouds[p], max_iterations=10)) for p in mine])
    full = sharding.all_gather_records(local, n_pairs)
    # --- sharded Scan Context search: entry i owned by rank i % world, query = last cloud
    descs = [orc.sc_compute(c) for c in clouds]
    own = [i for i in range(len(clouds) - 1) if sharding.owner(i, world) == rank]
    ld = np.array([orc.sc_distance(descs[-1], descs[i]) for i in own])
    keep = ld < 0.9
    order = np.lexsort((np.array(own)[keep], ld[keep]))
    md, me = sharding.all_gather_candidates(ld[keep][order], np.array(own)[keep][order], capacity=8)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), full=full, md=md, me=me)
    dist.destroy_process_group()


def test_two_rank_gather_and_merge(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    assert np.array_equal(r0["full"], r1["full"]) and np.array_equal(r0["me"], r1["me"])
    # single-process reference of the same work
    orc, clouds = _make_clouds(7)
    for p in range(6):
        ref = orc.icp_point_to_plane(clouds[p + 1], clouds[p], max_iterations=10)
        assert np.array_equal(r0["full"][p, :16], ref["transformation"].reshape(16))
        assert r0["full"][p, 17] == ref["num_iterations"]
    descs = [orc.sc_compute(c) for c in clouds]
    d = np.array([orc.sc_distance(descs[-1], descs[i]) for i in range(6)])
    keep = np.nonzero(d < 0.9)[0]
    order = np.lexsort((keep, d[keep]))
    assert np.array_equal(r0["me"], keep[order]) and np.array_equal(r0["md"], d[keep][order])


def test_accept_in_order_semantics():
    sys.path.insert(0, os.path.join(ROOT, "lidar-slam-from-scratch_b200", "python"))
    import sharding
    ent = [5, 2, 9, 4, 7]
    conv = [True, False, True, True, True]
    fit = [0.1, 0.1, 0.5, 0.2, 0.1]
    # failures (not converged / fitness too high) do not consume the budget (loop_closure.hpp:121)
    assert sharding.accept_in_order(ent, conv, fit, 0.3, 2) == [5, 4]
    assert sharding.accept_in_order(ent, conv, fit, 0.3, 10) == [5, 4, 7]
    assert list(sharding.shard_units(7, 1, 3)) == [1, 4]
