"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): voxel keys and kNN indices bit-exact; normals 1e-4 sign-canonicalised on
well-conditioned points (they are in fact expected to be bit-identical, the device Jacobi follows the oracle op for
op); final ICP poses 1e-4 m / 1e-5 rad; Scan Context distances 1e-5 with identical top-k IDs.
"""
import os

import numpy as np
import pytest

import oracle_lib
from conftest import rot_angle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def knn_cloud(rng, n, kind):
    if kind == "uniform":
        return rng.uniform(-10, 10, (n, 3))
    if kind == "lattice":
        return rng.integers(-4, 5, (n, 3)).astype(np.float64) * 0.5
    if kind == "dups":
        base = rng.uniform(-5, 5, (n // 4 + 1, 3))
        return base[rng.integers(0, len(base), n)]
    if kind == "collinear":
        t = rng.uniform(-10, 10, n)
        return np.stack([t, 2 * t, np.zeros(n)], axis=1)
    if kind == "clustered":  # wildly non-uniform density: dense blobs + far outliers (unbounded search radius)
        c = rng.normal(0, 0.05, (n - 20, 3)) + rng.integers(0, 3, (n - 20, 1)) * 40.0
        return np.vstack([c, rng.uniform(-500, 500, (20, 3))])
    raise ValueError(kind)


# ------------------------------------------------------------------ voxel grid
def test_voxel_full_scan_bit_exact(engine, oracle, synth, scene):
    raw = synth.scan(oracle_lib.SENSOR64, scene, (0.0, 0.0, 0.0), 7)
    assert raw.shape[0] > 100000
    for voxel in (0.5, 0.2):
        out, keys = engine.voxel_downsample(raw, voxel, return_keys=True)
        ref, rkeys = oracle.voxel_downsample(raw, voxel)
        assert np.array_equal(keys, rkeys)
        assert np.array_equal(out, ref)  # same members, same summation order -> bit-identical centroids


def test_voxel_hashed_and_sorted_paths_agree(engine, oracle, synth, scene, monkeypatch):
    """The one-pass hashed path (float32-born scans) and the sort-based path give the oracle's rows bit for bit;
    voxels whose integer sums cannot be proven exact are re-summed in input order; arbitrary fp64 input falls back."""
    import slam_b200
    raw = synth.scan(oracle_lib.SENSOR64, scene, (2.0, -1.0, 0.3), 21)
    ref, rkeys = oracle.voxel_downsample(raw, 0.5)
    out, keys = engine.voxel_downsample(raw, 0.5, return_keys=True)
    assert engine.last_voxel_path == 1
    assert np.array_equal(keys, rkeys) and np.array_equal(out, ref)
    monkeypatch.setenv("SB_VOXEL_SORT", "1")
    eng2 = slam_b200.Engine(0)
    monkeypatch.delenv("SB_VOXEL_SORT")
    out2, keys2 = eng2.voxel_downsample(raw, 0.5, return_keys=True)
    assert eng2.last_voxel_path == 2
    assert np.array_equal(keys2, rkeys) and np.array_equal(out2, ref)
    eng2.close()
    # a few members finer than 2^-44, members whose low bits make the fp64 loop round (1e-7 next to 40.x), and one
    # voxel whose sum of |v| is too large for the 64-bit fixed-point sums
    rng = np.random.default_rng(5)
    odd = raw.copy()
    pick = rng.choice(len(odd), 40, replace=False)
    odd[pick] += rng.uniform(-1e-9, 1e-9, (40, 3))
    pick2 = rng.choice(len(odd), 40, replace=False)
    odd[pick2, 2] = np.float32(1e-7) * rng.integers(1, 9, 40)
    odd = np.vstack([odd, np.tile([[400000.3, 0.1, 0.2]], (3, 1)) + rng.integers(0, 8, (3, 3)) / 64.0])
    ref3, rkeys3 = oracle.voxel_downsample(odd, 0.5)
    out3, keys3 = engine.voxel_downsample(odd, 0.5, return_keys=True)
    assert engine.last_voxel_path == 1
    assert np.array_equal(keys3, rkeys3) and np.array_equal(out3, ref3)
    # arbitrary doubles: nothing can be proven, the call falls back to the sort and is still exact
    arb = rng.uniform(-40, 40, (20000, 3))
    ref4, rkeys4 = oracle.voxel_downsample(arb, 0.5)
    out4, keys4 = engine.voxel_downsample(arb, 0.5, return_keys=True)
    assert engine.last_voxel_path == 2
    assert np.array_equal(keys4, rkeys4) and np.array_equal(out4, ref4)
    # one voxel per point (tables sized for 4 points per slot overflow), then the adapted sizing on the next call
    sparse = np.round(rng.uniform(-400, 400, (30000, 3)), 1).astype(np.float32).astype(np.float64)
    for _ in range(2):
        out5, keys5 = engine.voxel_downsample(sparse, 0.05, return_keys=True)
        ref5, rkeys5 = oracle.voxel_downsample(sparse, 0.05)
        assert np.array_equal(keys5, rkeys5) and np.array_equal(out5, ref5)
    assert engine.last_voxel_path == 1


def test_float32_ingest_matches_widened_input(engine, oracle, synth, scene):
    """sb_*_f32 (rows of 3 or 4 float32, widened on the device; file_utils.cpp:91-97) == the fp64 entry points on the
    host-widened rows, and the chunked host upload of sb_register_batch gives the device-resident result."""
    s = oracle_lib.small_sensor(32, 600)
    scans = [synth.scan(s, scene, (1.0 * i, 0.1 * i, 0.01 * i), 30 + i) for i in range(5)]
    off = np.r_[0, np.cumsum([len(c) for c in scans])]
    all64 = np.vstack(scans)
    all32 = all64.astype(np.float32)
    assert np.array_equal(all32.astype(np.float64), all64)  # the synthetic scans are float32-born
    xyzi = np.hstack([all32, np.ones((len(all32), 1), np.float32)])  # KITTI records: x, y, z, intensity
    ref, ref_off, ref_keys = engine.voxel_downsample_batch(all64, off, 0.5, return_keys=True)
    for pts in (all32, xyzi):
        out, out_off, keys = engine.voxel_downsample_batch_f32(pts, off, 0.5, return_keys=True)
        assert np.array_equal(out_off, ref_off) and np.array_equal(keys, ref_keys) and np.array_equal(out, ref)
    src, tgt = [1, 2, 3, 4], [0, 1, 2, 3]
    r64, sc64 = engine.register_batch(all64, off, src, tgt, voxel=0.5, want_sc=True)
    for pts in (all32, xyzi):
        r32, sc32 = engine.register_batch(pts, off, src, tgt, voxel=0.5, want_sc=True)
        assert np.array_equal(r32.transformations, r64.transformations)
        assert np.array_equal(r32.num_iterations, r64.num_iterations) and np.array_equal(sc32, sc64)
    o = oracle.icp_point_to_plane(oracle.voxel_downsample(scans[1], 0.5)[0], oracle.voxel_downsample(scans[0], 0.5)[0])
    check_icp(r64[0], o)
    # voxel <= 0 keeps the rows (file_utils.cpp:152)
    r0 = engine.register_batch(all32[:off[2]], off[:3], [1], [0], voxel=0.0)
    r1 = engine.register_batch(all64[:off[2]], off[:3], [1], [0], voxel=0.0)
    assert np.array_equal(r0.transformations, r1.transformations)


def test_voxel_edge_cases(engine, oracle):
    pts = np.array([[0.6, -0.6, 0.0], [0.6000000000000001, -0.2, 1e-300], [-1e-300, 0.2, -0.0]])
    out, keys = engine.voxel_downsample(pts, 0.2, return_keys=True)
    ref, rkeys = oracle.voxel_downsample(pts, 0.2)
    assert np.array_equal(keys, rkeys) and np.array_equal(out, ref)
    # voxel <= 0 returns the input (file_utils.cpp:152); empty and single-point clouds
    assert np.array_equal(engine.voxel_downsample(pts, 0.0), pts)
    assert engine.voxel_downsample(np.zeros((0, 3)), 0.5).shape == (0, 3)
    assert np.array_equal(engine.voxel_downsample(np.array([[1.0, 2.0, 3.0]]), 0.5), [[1.0, 2.0, 3.0]])
    # huge coordinates: keys need many bits
    big = np.array([[1e12, -1e12, 5.0], [1e12 + 0.3, -1e12, 5.0], [1e12 + 100.0, -1e12 + 3.0, 1e3], [1e12, -1e12, 5.1]])
    out, keys = engine.voxel_downsample(big, 0.25, return_keys=True)
    ref, rkeys = oracle.voxel_downsample(big, 0.25)
    assert np.array_equal(keys, rkeys) and np.array_equal(out, ref)
    # documented limits: non-finite input and key spans that do not pack into 64 bits are errors, not UB
    import slam_b200
    with pytest.raises(slam_b200.SlamB200Error):
        engine.voxel_downsample(np.array([[np.nan, 0, 0]]), 0.5)
    with pytest.raises(slam_b200.SlamB200Error) as e:
        engine.voxel_downsample(np.array([[1e12, -1e12, 5.0], [-7.0, 3.0, 1e9]]), 0.25)
    assert e.value.status == 5


def test_voxel_batch_ragged(engine, oracle):
    rng = np.random.default_rng(1)
    sizes = [0, 5000, 1, 0, 2049, 37]
    clouds = [np.round(rng.uniform(-30, 30, (n, 3)), 2) for n in sizes]
    off = np.r_[0, np.cumsum(sizes)]
    out, out_off, keys = engine.voxel_downsample_batch(np.vstack(clouds), off, 0.5, return_keys=True)
    for c, pts in enumerate(clouds):
        ref, rkeys = oracle.voxel_downsample(pts, 0.5) if len(pts) else (np.zeros((0, 3)), np.zeros((0, 3), np.int64))
        got = out[out_off[c]:out_off[c + 1]]
        assert np.array_equal(got, ref) and np.array_equal(keys[out_off[c]:out_off[c + 1]], rkeys)


def test_voxel_idempotent_full_size(engine, synth, scene):
    raw = synth.scan(oracle_lib.SENSOR128, scene, (3.0, 0.0, 0.2), 9)
    d1, k1 = engine.voxel_downsample(raw, 0.2, return_keys=True)
    assert np.all(np.lexsort((k1[:, 2], k1[:, 1], k1[:, 0])) == np.arange(len(k1)))  # sorted, unique keys
    assert len(np.unique(k1, axis=0)) == len(k1)
    d2, k2 = engine.voxel_downsample(d1, 0.2, return_keys=True)
    # a centroid stays in its voxel except for rounding at a face; every voxel then holds one point
    same = np.all(k1 == k2, axis=1) if len(k1) == len(k2) else np.zeros(1, bool)
    assert same.mean() > 0.999 and np.array_equal(d2[same], d1[same])


# ------------------------------------------------------------------ index: 1-NN and k-NN
@pytest.mark.parametrize("kind", ["uniform", "lattice", "dups", "collinear", "clustered"])
def test_knn_adversarial_bit_exact(engine, oracle, kind):
    import slam_b200
    rng = np.random.default_rng(21)
    pts = knn_cloud(rng, 3000, kind)
    q = np.vstack([pts[:300], knn_cloud(rng, 300, kind) + 0.125])
    tree = slam_b200.KDTree(engine, pts)
    for k in (1, 10, 20, 32):
        gi, gd = tree.k_nearest_batch(q, k)
        bi, bd = oracle.brute_knn(pts, q, k)
        assert np.array_equal(gi, bi), f"{kind} k={k}: {np.argwhere(gi != bi)[:5]}"
        assert np.array_equal(gd, bd)
    ni, nd = tree.nearest_batch(q)
    bi, bd = oracle.brute_knn(pts, q, 1)
    assert np.array_equal(ni, bi[:, 0]) and np.array_equal(nd, bd[:, 0])


def test_index_sort_paths_around_the_shared_memory_capacity(engine, oracle):
    """Morton sort of the index build: segments up to 12288 points are sorted by one CTA in shared memory, longer ones
    by the tiled device-wide sort; both sides of the boundary, odd sizes, and a batch that mixes them (then every
    segment takes the tiled path).  Also the voxel key sort of a cloud with > 12288 voxels."""
    import slam_b200
    for n in (12287, 12288, 12289, 4097):
        rng = np.random.default_rng(n)
        pts = rng.uniform(-40, 40, (n, 3))
        q = np.vstack([pts[:64], rng.uniform(-40, 40, (64, 3))])
        gi, gd = slam_b200.KDTree(engine, pts).k_nearest_batch(q, 7)
        bi, bd = oracle.brute_knn(pts, q, 7)
        assert np.array_equal(gi, bi) and np.array_equal(gd, bd), n
    rng = np.random.default_rng(5)
    sizes = [100, 12288, 12289, 1, 33, 5000]
    clouds = [np.round(rng.uniform(-60, 60, (m, 3)).astype(np.float32), 2).astype(np.float64) for m in sizes]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    res = engine.register_batch(np.vstack(clouds), off, [0, 2, 5], [1, 1, 2], voxel=0.0,
                                cfg=engine.icp_config(max_iterations=2))
    assert len(res) == 3
    for (s_, t_), r in zip(((0, 1), (2, 1), (5, 2)), res):
        o = oracle.icp_point_to_plane(clouds[s_], clouds[t_], max_iterations=2)
        assert r.num_iterations == o["num_iterations"] and np.allclose(r.error_history, o["error_history"], atol=1e-6)
    big = np.round(rng.uniform(-30, 30, (40000, 3)).astype(np.float32), 3).astype(np.float64)   # ~40 k voxels at 0.5 m
    out, keys = engine.voxel_downsample(big, 0.5, return_keys=True)
    oo, ok = oracle.voxel_downsample(big, 0.5)
    assert len(out) > 12288 and np.array_equal(keys, ok) and np.array_equal(out, oo)


def test_knn_tiny_and_empty(engine, oracle):
    import slam_b200
    for n in (1, 2, 3, 31, 32, 33, 1024, 1025):
        rng = np.random.default_rng(n)
        pts = rng.normal(size=(n, 3))
        q = rng.normal(size=(40, 3))
        tree = slam_b200.KDTree(engine, pts)
        gi, gd = tree.k_nearest_batch(q, 5)
        bi, bd = oracle.brute_knn(pts, q, 5)
        assert np.array_equal(gi, bi) and np.array_equal(gd, bd), n
    empty = slam_b200.KDTree(engine, np.zeros((0, 3)))
    i, d = empty.nearest_batch(np.zeros((3, 3)))
    assert np.all(i == -1) and np.all(d == np.finfo(np.float64).max)  # kdtree.hpp:33-36
    gi, _ = empty.k_nearest_batch(np.zeros((2, 3)), 4)
    assert np.all(gi == -1)
    with pytest.raises(slam_b200.SlamB200Error):
        tree.k_nearest_batch(q, 33)


def test_knn_scan_vs_oracle_tree(engine, oracle, synth, scene):
    """Config C3 shape: 128-beam scan, voxel 0.2, k = 10 — indices bit-exact vs the oracle KD-tree."""
    import slam_b200
    raw = synth.scan(oracle_lib.SENSOR128, scene, (0.0, 0.0, 0.0), 7)
    pts = engine.voxel_downsample(raw, 0.2)
    assert len(pts) > 20000
    tree = slam_b200.KDTree(engine, pts)
    gi, gd = tree.k_nearest_batch(pts, 10)
    oi, od = oracle.tree(pts).k_nearest_batch(pts, 10)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    assert np.array_equal(gi[:, 0], np.arange(len(pts)))  # a point is its own first neighbour (d2 = 0)
    assert np.all(np.diff(gd, axis=1) >= 0)


def test_find_correspondences(engine, oracle, small_pair):
    import slam_b200
    tree = slam_b200.KDTree(engine, small_pair["a"])
    m, d = tree.find_correspondences(small_pair["b"])
    oi, od = oracle.tree(small_pair["a"]).nearest_batch(small_pair["b"])
    assert np.array_equal(m, small_pair["a"][oi]) and np.array_equal(d, np.sqrt(od))


# ------------------------------------------------------------------ normals
def check_normals(gn, ge, on, oe, expect_excluded):
    """Normals within 1e-4 (north_star) on the points whose two smallest eigenvalues are separated — where they are not
    (neighbours almost on a line: one LiDAR ring) the direction is numerically arbitrary for ANY two eigen-solvers
    (SURVEY.md H3).  The excluded fraction is reported and must be the one SURVEY.md H3 measured for this
    configuration (about 3 % at 64 beams / 0.5 m / k = 20, about 13 % at 128 beams / 0.2 m / k = 10; this scene:
    4.5 % and 14.9 %), so that the filter cannot quietly swallow a broken kernel.  Eigenvalues, unit length and the
    z >= 0 orientation (icp.hpp:59-63) are checked on ALL points."""
    assert np.allclose(ge, oe, rtol=1e-9, atol=1e-12)
    ok = (oe[:, 1] - oe[:, 0]) / np.maximum(oe[:, 2], 1e-300) > 1e-2  # SURVEY.md H3 eigen-gap filter
    excluded = 1.0 - float(ok.mean())
    exact = float(np.mean(np.all(gn == on, axis=1)))
    print(f"normals: {len(gn)} points, excluded by the eigen-gap filter {excluded:.4f} (expected ~{expect_excluded}), "
          f"bit-identical to the oracle {exact:.4f}, max |diff| on the rest {np.max(np.abs(gn[ok] - on[ok])):.2e}")
    assert abs(excluded - expect_excluded) < 0.02, excluded
    assert np.max(np.abs(gn[ok] - on[ok])) < 1e-4
    assert np.allclose(np.linalg.norm(gn, axis=1), 1.0, atol=1e-12) and np.all(gn[:, 2] >= 0)
    return exact


def test_normals_k20_voxel05(engine, oracle, synth, scene):
    import slam_b200
    raw = synth.scan(oracle_lib.SENSOR64, scene, (0.0, 0.0, 0.0), 7)
    pts = engine.voxel_downsample(raw, 0.5)
    gn, ge = slam_b200.KDTree(engine, pts).estimate_normals(20, return_evals=True)
    on, oe = oracle.tree(pts).estimate_normals(20)
    exact = check_normals(gn, ge, on, oe, expect_excluded=0.045)
    assert exact > 0.99, f"only {exact:.4f} of the normals are bit-identical to the oracle"


def test_normals_k10_voxel02(engine, oracle, synth, scene):
    import slam_b200
    raw = synth.scan(oracle_lib.SENSOR128, scene, (0.0, 0.0, 0.0), 7)
    pts = engine.voxel_downsample(raw, 0.2)
    gn, ge = slam_b200.KDTree(engine, pts).estimate_normals(10, return_evals=True)
    on, oe = oracle.tree(pts).estimate_normals(10)
    exact = check_normals(gn, ge, on, oe, expect_excluded=0.149)
    assert exact > 0.99, f"only {exact:.4f} of the normals are bit-identical to the oracle"


def test_normals_packet_kernel_k10_in_a_subprocess():
    """k = 10 uses the warp-per-query kernel by default (faster on config C3); the packet kernel's k = 10 instantiation is
    selected with SB_KNN_PACKET_K10=1, read once per process — so it is checked in a process of its own: bit-identical
    normals and eigenvalues to the default kernel's on the same cloud (both equal the oracle's neighbour sets)."""
    import subprocess
    import sys
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle_lib, slam_b200
syn = oracle_lib.Synth(); orc = oracle_lib.Oracle()
raw = syn.scan(oracle_lib.small_sensor(64, 900), syn.scene(1, n_boxes=400), (0.0, 0.0, 0.0), 7)
eng = slam_b200.Engine(0)
pts = eng.voxel_downsample(raw, 0.2)
gn, ge = slam_b200.KDTree(eng, pts).estimate_normals(10, return_evals=True)
on, oe = orc.tree(pts).estimate_normals(10)
assert np.allclose(ge, oe, rtol=1e-9, atol=1e-12)
print("IDENTICAL", float(np.mean(np.all(gn == on, axis=1))), len(pts))
""" % (os.path.dirname(os.path.abspath(__file__)), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                              "lidar-slam-from-scratch_b200", "python"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, SB_KNN_PACKET_K10="1"))
    assert p.returncode == 0, p.stdout + p.stderr
    frac = float(p.stdout.split("IDENTICAL")[1].split()[0])
    assert frac > 0.99, p.stdout


def test_normals_degenerate(engine):
    import slam_b200
    n = slam_b200.KDTree(engine, np.array([[0, 0, 0], [1, 1, 1.0]])).estimate_normals(20)
    assert np.array_equal(n, [[0, 0, 1], [0, 0, 1.0]])  # icp.hpp:34-37


# ------------------------------------------------------------------ ICP
def test_solve_point_to_plane(engine, oracle):
    rng = np.random.default_rng(4)
    src = rng.normal(0, 3, (5000, 3))
    nrm = rng.normal(size=(5000, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    tgt = src + rng.normal(0, 0.01, (5000, 3)) + [0.05, -0.02, 0.01]
    assert np.allclose(engine.solve_point_to_plane(src, tgt, nrm), oracle.solve_point_to_plane(src, tgt, nrm), atol=1e-11)


def check_icp(g, o):
    assert g.status == 0
    assert g.converged == o["converged"]
    assert g.num_iterations == o["num_iterations"]
    assert len(g.error_history) == len(o["error_history"])
    assert np.max(np.abs(g.error_history - o["error_history"])) < 1e-6
    assert abs(g.final_error - o["final_error"]) < 1e-6
    dT = g.transformation @ np.linalg.inv(o["transformation"])
    assert np.linalg.norm(dT[:3, 3]) < 1e-4 and rot_angle(dT[:3, :3]) < 1e-5
    return np.linalg.norm(dT[:3, 3]), rot_angle(dT[:3, :3])


def test_icp_pair_c1(engine, oracle, synth, scene):
    """Config C1: one 64-beam pair, voxel 0.5, ICPConfig defaults."""
    a = synth.scan(oracle_lib.SENSOR64, scene, (0.0, 0.0, 0.0), 7)
    b = synth.scan(oracle_lib.SENSOR64, scene, (1.0, 0.1, 0.01), 8)
    da, db = engine.voxel_downsample(a, 0.5), engine.voxel_downsample(b, 0.5)
    g = engine.icp_point_to_plane(db, da)
    o = oracle.icp_point_to_plane(db, da)
    dt, dr = check_icp(g, o)
    assert dt < 1e-9 and dr < 1e-10  # in practice the two paths differ only by summation order


def test_icp_config_variants(engine, oracle, small_pair):
    a, b = small_pair["a"], small_pair["b"]
    for kw in (dict(max_iterations=2), dict(max_iterations=0), dict(tolerance=1e-12, max_iterations=50),
               dict(max_iterations=30, normals_k=10)):
        cfg = engine.icp_config(**kw)
        g = engine.icp_point_to_plane(b, a, cfg)
        o = oracle.icp_point_to_plane(b, a, max_iterations=cfg.max_iterations, tolerance=cfg.tolerance,
                                      normals_k=cfg.normals_k)
        check_icp(g, o)
    T0 = np.eye(4)
    T0[:3, 3] = [0.5, 0.0, 0.0]
    g = engine.icp_point_to_plane(b, a, engine.icp_config(initial_transform=T0))
    check_icp(g, oracle.icp_point_to_plane(b, a, T0=T0))


@pytest.mark.gpu
def test_icp_rank_deficient_cases(engine, oracle, small_pair):
    """J^T J exactly singular (icp.hpp:120: Eigen's pivoted ldlt() leaves the unconstrained components at zero): a planar
    target whose normals are all (0, 0, 1), and fewer than six source points.  The step must stay finite and equal
    the oracle's; nothing may be reported as converged with a NaN pose."""
    from test_reference_build import planar_pair
    src, tgt = planar_pair()
    g = engine.icp_point_to_plane(src, tgt)
    o = oracle.icp_point_to_plane(src, tgt)
    assert np.all(np.isfinite(g.transformation))
    check_icp(g, o)
    assert np.max(np.abs(g.transformation - o["transformation"])) < 1e-12
    # four source points: rank <= 4 with rounding-noise pivots, an ill-posed step (the oracle and the reference build
    # already take different paths there) — the result must be finite and flagged OK, no more can be asked
    few = small_pair["b"][[10, 500, 900, 1500]]
    g = engine.icp_point_to_plane(few, small_pair["a"])
    assert g.status == 0 and np.all(np.isfinite(g.transformation)) and np.all(np.isfinite(g.error_history))
    # the same two pairs inside a batch
    pts = np.vstack([src, tgt, few, small_pair["a"]])
    off = np.cumsum([0, len(src), len(tgt), len(few), len(small_pair["a"])]).astype(np.int64)
    res = engine.register_batch(pts, off, [0, 2], [1, 3])
    check_icp(res[0], o)
    assert res[1].status == 0 and np.all(np.isfinite(res[1].transformation))


@pytest.mark.gpu
def test_icp_without_any_correspondence_is_not_reported_converged(engine, small_pair):
    """Every source coordinate NaN: no query has a nearest neighbour (kdtree.hpp:125 never updates), the sums are
    empty and the RMS error is 0 — which must not pass for convergence (a NaN-free identity 'result' with
    final_error 0 would be accepted by the pose chain, slam_node.cpp:139-140, and by detect(), loop_closure.hpp:112)."""
    bad = np.full((50, 3), np.nan)
    pts = np.vstack([bad, small_pair["a"], small_pair["b"]])
    off = np.cumsum([0, len(bad), len(small_pair["a"]), len(small_pair["b"])]).astype(np.int64)
    res = engine.register_batch(pts, off, [0, 2], [1, 1])
    assert res[0].status != 0 and res[0].converged == 0
    assert res[1].status == 0 and res[1].converged == 1      # the healthy pair of the same batch is untouched
    poses = engine.odometry_poses(res)
    assert np.array_equal(poses[1], np.eye(4))               # the failed pair contributes the identity


def test_icp_empty_is_an_error(engine, small_pair):
    import slam_b200
    with pytest.raises(slam_b200.SlamB200Error) as e:
        engine.icp_point_to_plane(np.zeros((0, 3)), small_pair["a"])
    assert e.value.status == 2


def test_register_batch_equals_single_pairs(engine, oracle, synth, scene):
    """Batched pipeline (voxel + index + normals + ICP for several pairs in one call) vs the oracle pair by pair."""
    s = oracle_lib.small_sensor(32, 600)
    poses = [(0.0, 0.0, 0.0), (1.0, 0.1, 0.01), (2.1, 0.1, 0.02), (3.0, 0.3, 0.0), (3.9, 0.2, -0.02)]
    raws = [synth.scan(s, scene, p, 30 + i) for i, p in enumerate(poses)]
    off = np.r_[0, np.cumsum([len(r) for r in raws])]
    src = [1, 2, 3, 4, 0]
    tgt = [0, 1, 2, 3, 1]
    res, sc = engine.register_batch(np.vstack(raws), off, src, tgt, voxel=0.5, want_sc=True)
    ds = [oracle.voxel_downsample(r, 0.5)[0] for r in raws]
    for p in range(5):
        check_icp(res[p], oracle.icp_point_to_plane(ds[src[p]], ds[tgt[p]]))
    for c in range(5):
        assert np.array_equal(sc[c], oracle.sc_compute(ds[c]))


def test_chunked_upload_and_batch_composition_independence(engine, synth, scene, monkeypatch):
    """The host entry points upload, voxelise, index and compute normals chunk by chunk (SB_CHUNK_MB); the result does
    not depend on the chunking, nor on which other pairs share the batch (fixed summation order per pair)."""
    s = oracle_lib.small_sensor(32, 600)
    raws = [synth.scan(s, scene, (1.0 * i, 0.1 * (i % 3), 0.01 * i), 70 + i) for i in range(9)]
    off = np.r_[0, np.cumsum([len(r) for r in raws])]
    allpts = np.vstack(raws)
    src = [1, 2, 3, 4, 5, 6, 7, 8, 0, 4]
    tgt = [0, 1, 2, 3, 4, 5, 6, 7, 8, 2]
    ref, sc_ref = engine.register_batch(allpts, off, src, tgt, voxel=0.5, want_sc=True)
    monkeypatch.setenv("SB_CHUNK_MB", "1")  # ~2 clouds per chunk: 5 chunks, 5 batches of trees
    for pts in (allpts, allpts.astype(np.float32)):
        got, sc = engine.register_batch(pts, off, src, tgt, voxel=0.5, want_sc=True)
        assert np.array_equal(got.transformations, ref.transformations)
        assert np.array_equal(got.num_iterations, ref.num_iterations) and np.array_equal(sc, sc_ref)
        assert np.array_equal(got.final_errors, ref.final_errors)
    monkeypatch.delenv("SB_CHUNK_MB")
    for p in (0, 4, 9):  # each pair alone gives the bits it gave inside the batch
        a, b = src[p], tgt[p]
        pts = np.vstack([raws[a], raws[b]])
        one = engine.register_batch(pts, np.array([0, len(raws[a]), len(pts)]), [0], [1], voxel=0.5)
        assert np.array_equal(one.transformations[0], ref.transformations[p])
        assert one[0].num_iterations == ref[p].num_iterations


def test_register_batch_odd_pair_lists(engine, oracle, synth, scene):
    """Pair lists the odometry chain never produces: a cloud registered onto itself, a cloud that is only a source,
    one that is only a target, one that is in no pair, an empty cloud in the batch, and no pairs at all."""
    s = oracle_lib.small_sensor(32, 600)
    raws = [synth.scan(s, scene, (0.5 * i, 0.0, 0.01 * i), 90 + i) for i in range(4)]
    clouds = [raws[0], raws[1], np.zeros((0, 3)), raws[2], raws[3]]
    off = np.r_[0, np.cumsum([len(c) for c in clouds])]
    allpts = np.vstack(clouds)
    src, tgt = [0, 1, 3, 2], [0, 0, 1, 1]   # 0 onto itself; 1 is source and target; 3 only source; 4 unused; 2 is empty
    res, sc = engine.register_batch(allpts, off, src, tgt, voxel=0.5, want_sc=True)
    ds = [oracle.voxel_downsample(c, 0.5)[0] if len(c) else np.zeros((0, 3)) for c in clouds]
    assert np.allclose(res[0].transformation, np.eye(4), atol=1e-12) and res[0].final_error < 1e-12
    check_icp(res[1], oracle.icp_point_to_plane(ds[1], ds[0]))
    check_icp(res[2], oracle.icp_point_to_plane(ds[3], ds[1]))
    assert res[3].status == 2 and res[0].status == 0            # empty source: SB_ERR_EMPTY for that pair only
    assert np.array_equal(sc[2], np.zeros(1200)) and np.array_equal(sc[4], oracle.sc_compute(ds[4]))
    none, sc2 = engine.register_batch(allpts, off, [], [], voxel=0.5, want_sc=True)
    assert len(none) == 0 and np.array_equal(sc2, sc)


def test_icp_is_deterministic(engine, small_pair):
    r1 = engine.icp_point_to_plane(small_pair["b"], small_pair["a"])
    r2 = engine.icp_point_to_plane(small_pair["b"], small_pair["a"])
    assert np.array_equal(r1.transformation, r2.transformation) and np.array_equal(r1.error_history, r2.error_history)


# ------------------------------------------------------------------ Scan Context
def test_scan_context_descriptor_and_distance(engine, oracle, synth, scene):
    s = oracle_lib.small_sensor(32, 600)
    clouds = [oracle.voxel_downsample(synth.scan(s, scene, (2.0 * i, 0.1 * i, 0.05 * i), 40 + i), 0.5)[0] for i in range(12)]
    descs = np.stack([engine.sc_compute(c) for c in clouds])
    for c, d in zip(clouds, descs):
        assert np.array_equal(d, oracle.sc_compute(c))
    got = engine.sc_distance_batch(descs[0], descs)
    ref = np.array([oracle.sc_distance(descs[0], d) for d in descs])
    assert np.max(np.abs(got - ref)) < 1e-5
    assert np.array_equal(got, ref), "expected bit-identical distances (same order, no FMA)"
    assert np.array_equal(np.lexsort((np.arange(12), got))[:10], np.lexsort((np.arange(12), ref))[:10])
    assert engine.sc_distance(np.zeros(1200), descs[1]) == 1.0  # norm < 1e-10 rule, scan_context.hpp:137-138
    r, sk = engine.sc_keys(descs[0])
    D = descs[0].reshape(60, 20).T
    assert np.allclose(r, D.mean(axis=1), atol=1e-12) and np.allclose(sk, D.mean(axis=0), atol=1e-12)


def test_scan_context_edge_bins(engine, oracle):
    pts = np.array([[80.0, 0.0, 1.0], [0.1, 0.0, 2.0], [-5.0, 0.0, 3.0], [-5.0, -0.0, 4.0], [0.05, 0.0, 9.0],
                    [80.0000001, 0.0, 9.0], [3.0, 4.0, -2000.0], [0.0, 0.0, 5.0], [10.0, -10.0, np.nan]])
    assert np.array_equal(engine.sc_compute(pts), oracle.sc_compute(pts))
    assert np.array_equal(engine.sc_compute(np.zeros((0, 3))), np.zeros(1200))


def test_scan_context_random_db_topk(engine, oracle):
    rng = np.random.default_rng(8)
    db = np.where(rng.uniform(size=(300, 1200)) < 0.3, rng.uniform(-2, 8, (300, 1200)), 0.0)
    q = np.roll(db[17].reshape(60, 20), 5, axis=0).reshape(-1) + rng.normal(0, 0.01, 1200)
    got = engine.sc_distance_batch(q, db)
    ref = np.array([oracle.sc_distance(q, d) for d in db])
    assert np.max(np.abs(got - ref)) < 1e-5
    assert np.array_equal(np.lexsort((np.arange(300), got))[:10], np.lexsort((np.arange(300), ref))[:10])
    assert np.argmin(got) == 17


def test_scan_context_4k_database_top10(engine):
    """C4-sized search: one query against 4000 descriptors; distances vs a vectorised numpy restatement of
    scan_context.hpp:121-142 and identical top-10 ids (ties broken by entry id like loop_closure.hpp:92)."""
    rng = np.random.default_rng(11)
    n = 4000
    base = np.where(rng.uniform(size=(40, 60, 20)) < 0.35, rng.uniform(-1.7, 9.0, (40, 60, 20)), 0.0)
    db = base[rng.integers(0, 40, n)] + np.where(rng.uniform(size=(n, 60, 20)) < 0.05, rng.normal(0, 0.5, (n, 60, 20)), 0.0)
    db = np.stack([np.roll(d, int(sft), axis=0) for d, sft in zip(db, rng.integers(0, 60, n))])  # [sector][ring]
    q = db[1234] + rng.normal(0, 0.02, (60, 20))
    got = engine.sc_distance_batch(q.reshape(-1), db.reshape(n, -1))
    ref = np.full(n, np.inf)
    qa = np.sqrt((q * q).sum())
    nb = np.sqrt((db * db).sum(axis=(1, 2)))
    for sft in range(60):  # b(i, (j + shift) % 60): sector axis rolled by -shift
        dots = np.einsum("sr,nsr->n", q, np.roll(db, -sft, axis=1))
        norm = qa * nb
        ref = np.minimum(ref, np.where(norm < 1e-10, 1.0, 1.0 - dots / np.where(norm < 1e-10, 1.0, norm)))
    assert np.max(np.abs(got - ref)) < 1e-9
    assert np.array_equal(np.lexsort((np.arange(n), got))[:10], np.lexsort((np.arange(n), ref))[:10])
    assert np.argmin(got) == 1234


# ------------------------------------------------------------------ loop closure
def test_loop_detector_matches_oracle(engine, oracle, synth, scene):
    import slam_b200
    s = oracle_lib.small_sensor(16, 360)
    poses = [(float(i), 0.0, 0.0) for i in range(8)] + [(0.3, 0.05, 0.0), (1.2, -0.05, 0.01)]
    det = slam_b200.LoopClosureDetector(engine, frame_gap=3, sc_distance_threshold=0.5, icp_fitness_threshold=0.5,
                                        max_candidates=2)
    odet = oracle.loop(frame_gap=3, sc_thr=0.5, icp_thr=0.5, max_candidates=2)
    for i, p in enumerate(poses):
        c = oracle.voxel_downsample(synth.scan(s, scene, p, 50 + i), 0.5)[0]
        det.addFrame(c, i)
        odet.add(c, i)
        if i >= 4:
            gd, ge = det.candidates_local()
            od, oe = odet.candidates()
            assert np.array_equal(ge, oe) and np.array_equal(gd, od)
            g, o = det.detect(), odet.detect()
            assert [(r["query_frame"], r["match_frame"]) for r in g] == [(r["query_frame"], r["match_frame"]) for r in o]
            for rg, ro in zip(g, o):
                assert rg["scan_context_distance"] == ro["scan_context_distance"]
                assert abs(rg["icp_fitness"] - ro["icp_fitness"]) < 1e-6
                dT = rg["transform"] @ np.linalg.inv(ro["transform"])
                assert np.linalg.norm(dT[:3, 3]) < 1e-4 and rot_angle(dT[:3, :3]) < 1e-5
    assert det.size() == len(poses)
    det.clear()
    assert det.size() == 0 and det.detect() == []


def test_loop_sharded_candidates_merge(engine, oracle, synth, scene):
    """world = 2 emulated on one GPU: per-rank candidate lists merge to the single-rank list."""
    import slam_b200
    s = oracle_lib.small_sensor(16, 360)
    dets = [slam_b200.LoopClosureDetector(engine, frame_gap=2, sc_distance_threshold=0.6, rank=r, world=2) for r in (0, 1)]
    one = slam_b200.LoopClosureDetector(engine, frame_gap=2, sc_distance_threshold=0.6)
    for i in range(9):
        c = oracle.voxel_downsample(synth.scan(s, scene, (1.5 * (i % 5), 0.0, 0.0), 60 + i), 0.5)[0]
        for d in dets + [one]:
            d.addFrame(c, i)
    parts = [d.candidates_local() for d in dets]
    md = np.concatenate([p[0] for p in parts])
    me = np.concatenate([p[1] for p in parts])
    order = np.lexsort((me, md))
    d1, e1 = one.candidates_local()
    assert np.array_equal(me[order], e1) and np.array_equal(md[order], d1)
    # each rank verifies the entries it owns; together they reproduce the single-rank verification
    ref, rconv = one.verify_entries(e1[:4], d1[:4])
    for j, ent in enumerate(e1[:4]):
        res, conv = dets[ent % 2].verify_entries([ent], [d1[j]])
        assert conv[0] == rconv[j] and np.array_equal(res[0]["transform"], ref[j]["transform"])


def test_loop_pools_grow_without_changing_results(engine, oracle, synth, scene):
    """A detector whose pools were reserved far too small outgrows them several times (copy into the doubled pool,
    old pool retired); candidates and verified loops must equal those of a detector with the default pools."""
    import slam_b200
    s = oracle_lib.small_sensor(16, 360)
    small = slam_b200.LoopClosureDetector(engine, frame_gap=2, sc_distance_threshold=0.6, icp_fitness_threshold=0.5)
    small.reserve(1, 64)
    big = slam_b200.LoopClosureDetector(engine, frame_gap=2, sc_distance_threshold=0.6, icp_fitness_threshold=0.5)
    for i in range(12):
        c = oracle.voxel_downsample(synth.scan(s, scene, (1.5 * (i % 4), 0.0, 0.0), 80 + i), 0.5)[0]
        small.addFrame(c, i)
        big.addFrame(c, i)
        if i >= 3:
            ds, es = small.candidates_local()
            db, eb = big.candidates_local()
            assert np.array_equal(es, eb) and np.array_equal(ds, db)
    gs, gb = small.detect(), big.detect()
    assert len(gs) == len(gb) and len(gb) > 0
    for a, b in zip(gs, gb):
        assert a["match_frame"] == b["match_frame"] and np.array_equal(a["transform"], b["transform"])
    assert small.size() == big.size() == 12


# ------------------------------------------------------------------ committed golden fixtures (independent numpy)
def test_golden_fixtures_gpu(engine):
    import slam_b200
    g = np.load(os.path.join(GOLDEN, "golden_small.npz"))
    out, keys = engine.voxel_downsample(g["raw"], 0.5, return_keys=True)
    assert np.array_equal(keys, g["voxel_keys"]) and np.array_equal(out, g["voxel_xyz"])
    tree = slam_b200.KDTree(engine, g["voxel_xyz"])
    idx, d2 = tree.k_nearest_batch(g["voxel_xyz"], 10)
    assert np.array_equal(idx, g["knn_idx"]) and np.array_equal(d2, g["knn_d2"])
    nrm = tree.estimate_normals(10)
    ok = g["normal_ok"]
    assert np.max(np.abs(nrm[ok] - g["normals"][ok])) < 1e-4
    assert np.array_equal(engine.sc_compute(g["voxel_xyz"]), g["sc_desc"])
    assert abs(engine.sc_distance(g["sc_desc"], g["sc_desc_b"]) - float(g["sc_dist"])) < 1e-5
    r = engine.icp_point_to_plane(g["icp_src"], g["voxel_xyz"], engine.icp_config(max_iterations=int(g["icp_max_it"])))
    dT = r.transformation @ np.linalg.inv(g["icp_T"])
    assert np.linalg.norm(dT[:3, 3]) < 1e-4 and rot_angle(dT[:3, :3]) < 1e-5
    assert r.converged == bool(g["icp_converged"]) and np.allclose(r.error_history, g["icp_history"], atol=1e-6)


# ------------------------------------------------------------------ after the path: world frame, occupancy, map
def test_world_frame_occupancy_and_global_map(engine, oracle, synth, scene):
    """slam_node.cpp:147 (world clouds, bit-exact), :211-229 (occupancy cell set, identical) and
    :196-209 + :237 (global map = world clouds voxel-downsampled at 2 * voxel_size, bit-exact vs the oracle)."""
    s = oracle_lib.small_sensor(32, 600)
    poses = [(2.0 * i, 0.3 * i, 0.05 * i) for i in range(6)]
    clouds = [oracle.voxel_downsample(synth.scan(s, scene, p, 80 + i), 0.5)[0] for i, p in enumerate(poses)]
    off = np.r_[0, np.cumsum([len(c) for c in clouds])]
    Ts = []
    for x, y, yaw in poses:
        T = np.eye(4)
        T[:3, :3] = [[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]]
        T[:3, 3] = [x, y, 1.73]
        Ts.append(T)
    Ts = np.stack(Ts)
    allpts = np.vstack(clouds)
    world = engine.transform_clouds(allpts, off, Ts)
    ref_world = np.vstack([oracle.transform_cloud(c, T) for c, T in zip(clouds, Ts)])
    assert np.array_equal(world, ref_world)
    cells, count = engine.occupancy_cells(allpts, off, Ts)
    ref_cells = oracle.occupancy_cells(allpts, off, Ts)
    assert count == len(ref_cells) and np.array_equal(cells, ref_cells) and count > 100
    few, count2 = engine.occupancy_cells(allpts, off, Ts, capacity=10)
    assert count2 == count and np.array_equal(few, ref_cells[:10])
    gm = engine.global_map(allpts, off, Ts, 1.0)
    ref_gm, _ = oracle.voxel_downsample(ref_world, 1.0)
    assert np.array_equal(gm, ref_gm)
    assert np.array_equal(engine.global_map(allpts, off, Ts, 0.0), ref_world)
    assert engine.occupancy_cells(np.zeros((0, 3)), [0], np.zeros((0, 16)))[1] == 0
    # the PointCloud2 payload of what the node publishes (eigen_to_pointcloud2, slam_node.cpp:299-322): static_cast<float>
    w32 = engine.transform_clouds_f32(allpts, off, Ts)
    assert w32.dtype == np.float32 and np.array_equal(w32, ref_world.astype(np.float32))
    g32 = engine.global_map_f32(allpts, off, Ts, 1.0)
    assert g32.dtype == np.float32 and np.array_equal(g32, ref_gm.astype(np.float32))
    assert len(engine.global_map_f32(np.zeros((0, 3)), [0], np.zeros((0, 16)), 1.0)) == 0
    # the whole offline-mapping chain through the ABI: register the sequence as one batch, chain the poses the way
    # process_frame does (slam_node.cpp:139-145), build the map from them
    res = engine.register_batch(allpts, off, np.arange(1, 6), np.arange(0, 5), voxel=0.0)
    chain = engine.odometry_poses(res)
    pose = np.eye(4)
    for i in range(5):
        r = res[i]
        pose = pose @ (r.transformation if r.converged and not r.final_error > 1.0 else np.eye(4))
        assert np.allclose(chain[i + 1], pose, rtol=0, atol=1e-14)
    assert np.all(np.isfinite(chain)) and np.array_equal(chain[0], np.eye(4))
    assert len(engine.global_map(allpts, off, chain, 1.0)) > 1000


# ------------------------------------------------------------------ fixtures produced by the reference's own sources
def test_reference_golden_gpu(engine):
    """CUDA path against tests/golden/reference_small.npz = outputs of the reference's own sources (compiled against the
    Eigen stand-in, tests/golden/make_reference_golden.py).  Tolerances are north_star's: voxel keys / centroids and
    neighbour indices bit-exact, normals 1e-4 (sign-canonical), pose 1e-4 m / 1e-5 rad, Scan Context distance 1e-5."""
    import slam_b200
    g = np.load(os.path.join(GOLDEN, "reference_small.npz"))

    def sort_rows(x):
        return x[np.lexsort((x[:, 2], x[:, 1], x[:, 0]))]

    for raw, v, key in ((g["raw_a"], 0.5, "voxel_a"), (g["raw_b"], 0.5, "voxel_b"), (g["raw_a"], 0.2, "voxel_a_02")):
        assert np.array_equal(sort_rows(engine.voxel_downsample(raw.astype(np.float64), v)), g[key])
        out32, off32 = engine.voxel_downsample_batch_f32(raw, np.array([0, len(raw)], dtype=np.int64), v)[:2]
        assert np.array_equal(sort_rows(out32[off32[0]:off32[1]]), g[key])       # float32 records, widened on the device
    a, b = g["voxel_a"], g["voxel_b"]
    tree = slam_b200.KDTree(engine, a)
    assert np.array_equal(tree.k_nearest_batch(a, 20)[0], g["knn20_a"])
    assert np.array_equal(tree.k_nearest_batch(a, 10)[0], g["knn10_a"])
    idx, d2 = tree.nearest_batch(b)
    assert np.array_equal(idx, g["nn_b_in_a"]) and np.array_equal(d2, g["nn_b_in_a_d2"])
    for k, key in ((20, "normals20_a"), (10, "normals10_a")):
        n = tree.estimate_normals(k)
        dots = np.sum(n * g[key], axis=1)
        ok = dots > 1 - 1e-6                       # elsewhere the two smallest eigenvalues coincide: direction undefined
        assert ok.mean() > 0.999 and np.max(np.abs(n[ok] - g[key][ok])) < 1e-4
    assert np.array_equal(engine.sc_compute(a), g["sc_a"]) and np.array_equal(engine.sc_compute(b), g["sc_b"])
    assert abs(engine.sc_distance(g["sc_a"], g["sc_b"]) - float(g["sc_dist_ab"])) < 1e-5
    ring, sector = engine.sc_keys(g["sc_a"])
    assert np.allclose(ring, g["sc_ring_key_a"], atol=1e-12) and np.allclose(sector, g["sc_sector_key_a"], atol=1e-12)
    T = engine.solve_point_to_plane(b, a[g["nn_b_in_a"]], g["normals20_a"][g["nn_b_in_a"]])
    assert np.max(np.abs(T - g["solve_T"])) < 1e-9
    for name, it in (("icp50", 50), ("icp3", 3), ("icp30", 30)):
        r = engine.icp_point_to_plane(b, a, engine.icp_config(max_iterations=it))
        assert [r.num_iterations, int(r.converged)] == g[name + "_meta"].tolist()
        assert np.allclose(r.error_history, g[name + "_history"], rtol=0, atol=1e-6)
        dT = r.transformation @ np.linalg.inv(g[name + "_T"])
        assert np.linalg.norm(dT[:3, 3]) < 1e-4 and rot_angle(dT[:3, :3]) < 1e-5
    det = slam_b200.LoopClosureDetector(engine, frame_gap=3, sc_distance_threshold=0.5, icp_fitness_threshold=0.5,
                                        max_candidates=2)
    off, found = g["loop_offsets"], []
    for i in range(len(off) - 1):
        det.addFrame(g["loop_clouds"][off[i]:off[i + 1]], i)
        found += det.detect()
    want = g["loop_results"]
    assert len(found) == len(want)
    for x, w in zip(found, want):
        assert (x["query_frame"], x["match_frame"]) == (int(w[0]), int(w[1]))
        assert abs(x["scan_context_distance"] - w[2]) < 1e-5 and abs(x["icp_fitness"] - w[3]) < 1e-6
        dT = x["transform"] @ np.linalg.inv(w[4:].reshape(4, 4))
        assert np.linalg.norm(dT[:3, 3]) < 1e-4 and rot_angle(dT[:3, :3]) < 1e-5


def test_reference_golden_config_c1_gpu(engine, synth, scene):
    """BASELINE.json configs[0] through the CUDA path against the reference's own result for it (fixture made by
    tests/golden/make_reference_golden.py; the raw scans are regenerated and checked by SHA-256)."""
    import hashlib
    g = np.load(os.path.join(GOLDEN, "reference_small.npz"))
    a = synth.scan(oracle_lib.SENSOR64, scene, (0.0, 0.0, 0.0), 7)
    b = synth.scan(oracle_lib.SENSOR64, scene, (1.0, 0.1, 0.01), 8)
    sha = np.frombuffer(hashlib.sha256(a.tobytes() + b.tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, g["c1_sha256"])

    def sort_rows(x):
        return x[np.lexsort((x[:, 2], x[:, 1], x[:, 0]))]

    va, vb = sort_rows(engine.voxel_downsample(a, 0.5)), sort_rows(engine.voxel_downsample(b, 0.5))
    assert [len(a), len(b), len(va), len(vb)] == g["c1_voxel_counts"].tolist()
    assert np.array_equal(va.sum(axis=0), g["c1_voxel_a_sum"]) and np.array_equal(vb.sum(axis=0), g["c1_voxel_b_sum"])
    r = engine.icp_point_to_plane(vb, va)
    assert [r.num_iterations, int(r.converged)] == g["c1_meta"].tolist()
    assert np.allclose(r.error_history, g["c1_history"], rtol=0, atol=1e-6)
    dT = r.transformation @ np.linalg.inv(g["c1_T"])
    assert np.linalg.norm(dT[:3, 3]) < 1e-4 and rot_angle(dT[:3, :3]) < 1e-5
    # the same pair through the batch entry point (voxel grid + index + normals + ICP in one call, key-ordered rows)
    off = np.array([0, len(a), len(a) + len(b)], dtype=np.int64)
    rb = engine.register_batch(np.vstack([a, b]), off, [1], [0], voxel=0.5)[0]
    dT = rb.transformation @ np.linalg.inv(g["c1_T"])
    assert rb.num_iterations == int(g["c1_meta"][0]) and np.linalg.norm(dT[:3, 3]) < 1e-4 and rot_angle(dT[:3, :3]) < 1e-5
