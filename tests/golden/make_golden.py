"""Generates tests/golden/golden_small.npz with a PURE-NUMPY restatement of the reference's hot path
(slam_viz/src/core/file_utils.cpp:148-196, include/slam_viz/core/{kdtree,icp,scan_context}.hpp).

It shares no code with oracle/slam_oracle.cpp: brute-force neighbours, numpy.linalg.eigh normals, numpy.linalg.solve
Gauss-Newton steps.  The reference itself cannot run here (Eigen is absent), so these vectors pin the oracle and the
CUDA path against an independent implementation of the same source semantics.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # only for the synthetic raycaster (input generation)


def voxel(pts, v):
    keys = np.floor(pts / v).astype(np.int64)
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    ks = keys[order]
    starts = np.r_[0, np.nonzero(np.any(ks[1:] != ks[:-1], axis=1))[0] + 1, len(ks)]
    out = np.empty((len(starts) - 1, 3))
    for i in range(len(starts) - 1):
        acc = np.zeros(3)
        for j in order[starts[i]:starts[i + 1]]:
            acc = acc + pts[j]
        out[i] = acc / float(starts[i + 1] - starts[i])
    return out, ks[starts[:-1]]


def d2_matrix(P, Q):
    d = P[None, :, :] - Q[:, None, :]
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


def knn(P, Q, k):
    D = d2_matrix(P, Q)
    idx = np.empty((len(Q), k), dtype=np.int32)
    dd = np.empty((len(Q), k))
    ar = np.arange(len(P))
    for i in range(len(Q)):
        o = np.lexsort((ar, D[i]))[:k]
        idx[i], dd[i] = o, D[i][o]
    return idx, dd


def normals(P, k):
    idx, _ = knn(P, P, k)
    out = np.empty_like(P)
    ok = np.zeros(len(P), dtype=bool)
    for i in range(len(P)):
        nb = P[idx[i]]
        c = nb.sum(axis=0) / k
        C = (nb - c).T @ (nb - c) / k
        w, V = np.linalg.eigh(C)
        v = V[:, 0]
        if v[2] < 0:
            v = -v
        out[i] = v / np.linalg.norm(v)
        ok[i] = (w[1] - w[0]) / max(w[2], 1e-300) > 1e-2
    return out, ok


def sc(P):
    d = np.full((20, 60), -np.finfo(np.float64).max)
    for x, y, z in P:
        r = np.sqrt(x * x + y * y)
        a = np.arctan2(y, x) + np.pi
        if r > 80.0 or r < 0.1:
            continue
        i = min(max(int(r / 4.0), 0), 19)
        j = min(max(int(a / (2.0 * np.pi / 60)), 0), 59)
        d[i, j] = max(d[i, j], z)
    d[d < -1000] = 0.0
    return d.T.reshape(-1).copy()  # column-major 20x60


def sc_dist(a, b):
    A, B = a.reshape(60, 20).T, b.reshape(60, 20).T
    best = np.finfo(np.float64).max
    for s in range(60):
        Bs = np.roll(B, -s, axis=1)
        n = np.sqrt((A * A).sum()) * np.sqrt((Bs * Bs).sum())
        best = min(best, 1.0 if n < 1e-10 else 1.0 - (A * Bs).sum() / n)
    return best


def icp(src, tgt, k, max_it, tol=1e-6, min_err=1e-9):
    nrm, _ = normals(tgt, k)
    cur = src.copy()
    T = np.eye(4)
    prev = np.finfo(np.float64).max
    hist, conv = [], False
    for _ in range(max_it):
        j = knn(tgt, cur, 1)[0][:, 0]
        q, n = tgt[j], nrm[j]
        b = np.einsum("ij,ij->i", q - cur, n)
        e = np.sqrt(np.mean(b * b))
        hist.append(e)
        if e < min_err or abs(prev - e) < tol:
            conv = True
            break
        J = np.hstack([np.cross(cur, n), n])
        x = np.linalg.solve(J.T @ J, J.T @ b)
        th = np.linalg.norm(x[:3])
        R = np.eye(3)
        if th >= 1e-10:
            a = x[:3] / th
            K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
            R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
        D = np.eye(4)
        D[:3, :3], D[:3, 3] = R, x[3:]
        cur = cur @ R.T + x[3:]
        T = D @ T
        prev = e
    j = knn(tgt, cur, 1)[0][:, 0]
    b = np.einsum("ij,ij->i", tgt[j] - cur, nrm[j])
    hist.append(np.sqrt(np.mean(b * b)))
    return T, conv, np.array(hist)


def main():
    syn = oracle_lib.Synth()
    scene = syn.scene(1, n_boxes=400)
    s = oracle_lib.small_sensor(16, 360)
    raw = syn.scan(s, scene, (0.0, 0.0, 0.0), 7)
    raw_b = syn.scan(s, scene, (0.6, 0.05, 0.01), 8)
    vx, vk = voxel(raw, 0.5)
    vb, _ = voxel(raw_b, 0.5)
    ki, kd = knn(vx, vx, 10)
    nrm, ok = normals(vx, 10)
    da, db = sc(vx), sc(vb)
    T, conv, hist = icp(vb, vx, 20, 12)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), raw=raw, voxel_xyz=vx, voxel_keys=vk, knn_idx=ki,
                        knn_d2=kd, normals=nrm, normal_ok=ok, sc_desc=da, sc_desc_b=db, sc_dist=sc_dist(da, db),
                        icp_src=vb, icp_T=T, icp_converged=conv, icp_history=hist, icp_max_it=12)
    print("raw", raw.shape, "voxels", vx.shape, "icp", conv, len(hist), T[:3, 3])


if __name__ == "__main__":
    main()
