// slam_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A from-source-semantics restatement, in plain C++17 with no Eigen, of the
// front-end hot path of kaushik884/LiDAR-SLAM-from-scratch:
//   voxel_downsample      slam_viz/src/core/file_utils.cpp:148-196
//   KDTree                slam_viz/include/slam_viz/core/kdtree.hpp:18-186
//   NearestNeighborSearch slam_viz/include/slam_viz/core/kdtree.hpp:193-221
//   estimate_normals      slam_viz/include/slam_viz/core/icp.hpp:23-67
//   solve_point_to_plane  slam_viz/include/slam_viz/core/icp.hpp:89-144
//   icp_point_to_plane    slam_viz/include/slam_viz/core/icp.hpp:157-258
//   ScanContext           slam_viz/include/slam_viz/core/scan_context.hpp:24-145
//   LoopClosureDetector   slam_viz/include/slam_viz/core/loop_closure.hpp:41-149
//
// PARITY: the reference ships no tests, golden vectors or fixtures, and Eigen
// is absent here.  This restatement is checked (tests/test_reference_build.py,
// tests/golden/reference_small.npz) against the reference's OWN sources
// compiled unmodified over oracle/eigen_standin (oracle/build_ref.sh): control
// flow, operation order, indices, descriptors bit for bit; ICP history 1e-9.
// PARITY UNPINNED only for Eigen's internal kernels, which that stand-in and
// this file both replace (expected difference 1e-15 relative):
//   * 3x3 symmetric eigen: cyclic Jacobi (Eigen: tridiagonal QR), icp.hpp:55
//   * 6x6 solve: LDLT without pivoting (Eigen: pivoted LDLT),     icp.hpp:120
//   * dense sums/products: ascending-index scalar loops, no FMA.
// Also checked against independent implementations (brute-force kNN, scipy
// cKDTree, numpy eigh/solve, analytic ICP cases) in tests/test_oracle.py.
// Canonical tie rules (the reference is implementation-defined on exact ties,
// kdtree.hpp:125,160): neighbours are ranked by (d^2, index) lexicographically.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  Build: oracle/build.sh
// (g++ -O3 -DNDEBUG -ffp-contract=off; the reference's flags, CMakeLists.txt:13).

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <queue>
#include <unordered_map>
#include <utility>
#include <vector>

namespace orc {

using i64 = long long;

// ---------------------------------------------------------------------------
// d^2 with the association the oracle DEFINES: (dx*dx + dy*dy) + dz*dz, no FMA
// (kdtree.hpp:124 `(point - query).squaredNorm()`).
// ---------------------------------------------------------------------------
static inline double dist2(const double* a, const double* b) {
    double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return (dx * dx + dy * dy) + dz * dz;
}

// ---------------------------------------------------------------------------
// voxel_downsample — file_utils.cpp:148-196
// ---------------------------------------------------------------------------
struct VoxelKey {
    i64 x, y, z;
    bool operator==(const VoxelKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct VoxelHash {  // file_utils.cpp:162-170
    size_t operator()(const VoxelKey& v) const {
        size_t h = 0;
        h ^= std::hash<i64>{}(v.x) + 0x9e3779b9 + (h << 6) + (h >> 2);
        h ^= std::hash<i64>{}(v.y) + 0x9e3779b9 + (h << 6) + (h >> 2);
        h ^= std::hash<i64>{}(v.z) + 0x9e3779b9 + (h << 6) + (h >> 2);
        return h;
    }
};

// Returns m.  Output rows are emitted in CANONICAL order: ascending (kx,ky,kz)
// (the reference's order is unordered_map iteration order — unspecified).
static i64 voxel_downsample(const double* xyz, i64 n, double voxel, double* out_xyz, i64* out_keys) {
    if (voxel <= 0) {  // file_utils.cpp:152
        std::memcpy(out_xyz, xyz, sizeof(double) * 3 * n);
        if (out_keys) std::memset(out_keys, 0, sizeof(i64) * 3 * n);
        return n;
    }
    std::unordered_map<VoxelKey, std::vector<int>, VoxelHash> voxel_map;
    for (i64 i = 0; i < n; ++i) {  // file_utils.cpp:175-181
        VoxelKey key;
        key.x = static_cast<i64>(std::floor(xyz[3 * i + 0] / voxel));
        key.y = static_cast<i64>(std::floor(xyz[3 * i + 1] / voxel));
        key.z = static_cast<i64>(std::floor(xyz[3 * i + 2] / voxel));
        voxel_map[key].push_back(static_cast<int>(i));
    }
    struct Row { VoxelKey k; double c[3]; };
    std::vector<Row> rows;
    rows.reserve(voxel_map.size());
    for (const auto& kv : voxel_map) {  // file_utils.cpp:186-193
        double c[3] = {0, 0, 0};
        for (int i : kv.second) {
            c[0] += xyz[3 * i + 0];
            c[1] += xyz[3 * i + 1];
            c[2] += xyz[3 * i + 2];
        }
        double cnt = static_cast<double>(kv.second.size());
        Row r;
        r.k = kv.first;
        r.c[0] = c[0] / cnt; r.c[1] = c[1] / cnt; r.c[2] = c[2] / cnt;
        rows.push_back(r);
    }
    std::sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) {
        if (a.k.x != b.k.x) return a.k.x < b.k.x;
        if (a.k.y != b.k.y) return a.k.y < b.k.y;
        return a.k.z < b.k.z;
    });
    i64 m = static_cast<i64>(rows.size());
    for (i64 i = 0; i < m; ++i) {
        out_xyz[3 * i + 0] = rows[i].c[0];
        out_xyz[3 * i + 1] = rows[i].c[1];
        out_xyz[3 * i + 2] = rows[i].c[2];
        if (out_keys) {
            out_keys[3 * i + 0] = rows[i].k.x;
            out_keys[3 * i + 1] = rows[i].k.y;
            out_keys[3 * i + 2] = rows[i].k.z;
        }
    }
    return m;
}

// ---------------------------------------------------------------------------
// KDTree — kdtree.hpp:18-186.  Same build (nth_element median split, axis =
// depth%3, post-order node append) and the same near-first recursive search.
// The only change is the canonical tie rule: a candidate replaces the
// incumbent iff (d2, idx) is lexicographically smaller, and the far branch is
// pruned with `>`, not `>=`, so equal-distance points on the far side are seen.
// Without exact ties this is step-for-step the reference traversal.
// ---------------------------------------------------------------------------
class KDTree {
public:
    KDTree(const double* pts, int n) : pts_(pts, pts + 3 * (size_t)n), n_(n) {
        indices_.resize(n);
        for (int i = 0; i < n; ++i) indices_[i] = i;
        nodes_.reserve(n);
        root_ = build(0, n, 0);
    }
    int size() const { return n_; }
    const double* point(int i) const { return &pts_[3 * (size_t)i]; }

    void nearest(const double* q, int& best_idx, double& best_d2) const {
        best_idx = -1;
        best_d2 = std::numeric_limits<double>::max();
        search_nearest(root_, q, 0, best_idx, best_d2);
    }
    // kdtree.hpp:43-59
    void nearest_batch(const double* q, int nq, int* idx, double* d2) const {
        for (int i = 0; i < nq; ++i) {
            int bi; double bd;
            nearest(q + 3 * (size_t)i, bi, bd);
            idx[i] = bi;
            if (d2) d2[i] = bd;
        }
    }
    // kdtree.hpp:65-78; returns neighbours ascending by (d2, idx)
    int k_nearest(const double* q, int k, int* out, double* out_d2) const {
        std::priority_queue<std::pair<double, int>> heap;
        if (k > 0) search_k_nearest(root_, q, 0, k, heap);
        int m = static_cast<int>(heap.size());
        for (int i = m - 1; i >= 0; --i) {
            out[i] = heap.top().second;
            if (out_d2) out_d2[i] = heap.top().first;
            heap.pop();
        }
        return m;
    }

private:
    struct Node { int index; int left = -1; int right = -1; };

    int build(int start, int end, int depth) {  // kdtree.hpp:87-110
        if (start >= end) return -1;
        int axis = depth % 3;
        int mid = (start + end) / 2;
        std::nth_element(indices_.begin() + start, indices_.begin() + mid, indices_.begin() + end,
                         [this, axis](int a, int b) { return pts_[3 * (size_t)a + axis] < pts_[3 * (size_t)b + axis]; });
        Node node;
        node.index = indices_[mid];
        node.left = build(start, mid, depth + 1);
        node.right = build(mid + 1, end, depth + 1);
        nodes_.push_back(node);
        return static_cast<int>(nodes_.size()) - 1;
    }

    void search_nearest(int node_idx, const double* q, int depth, int& best_idx, double& best_d2) const {
        if (node_idx < 0) return;  // kdtree.hpp:112-142
        const Node& node = nodes_[node_idx];
        const double* p = &pts_[3 * (size_t)node.index];
        double d2 = dist2(p, q);
        if (d2 < best_d2 || (d2 == best_d2 && node.index < best_idx)) {
            best_d2 = d2;
            best_idx = node.index;
        }
        int axis = depth % 3;
        double diff = q[axis] - p[axis];
        int first = diff < 0 ? node.left : node.right;
        int second = diff < 0 ? node.right : node.left;
        search_nearest(first, q, depth + 1, best_idx, best_d2);
        if (!(diff * diff > best_d2)) search_nearest(second, q, depth + 1, best_idx, best_d2);
    }

    void search_k_nearest(int node_idx, const double* q, int depth, int k,
                          std::priority_queue<std::pair<double, int>>& heap) const {
        if (node_idx < 0) return;  // kdtree.hpp:144-180
        const Node& node = nodes_[node_idx];
        const double* p = &pts_[3 * (size_t)node.index];
        double d2 = dist2(p, q);
        std::pair<double, int> cand{d2, node.index};
        if (static_cast<int>(heap.size()) < k) {
            heap.push(cand);
        } else if (cand < heap.top()) {
            heap.pop();
            heap.push(cand);
        }
        int axis = depth % 3;
        double diff = q[axis] - p[axis];
        int first = diff < 0 ? node.left : node.right;
        int second = diff < 0 ? node.right : node.left;
        search_k_nearest(first, q, depth + 1, k, heap);
        double threshold = static_cast<int>(heap.size()) < k ? std::numeric_limits<double>::max() : heap.top().first;
        if (!(diff * diff > threshold)) search_k_nearest(second, q, depth + 1, k, heap);
    }

    std::vector<double> pts_;
    int n_;
    std::vector<int> indices_;
    std::vector<Node> nodes_;
    int root_ = -1;
};

// Brute-force canonical (d2, idx) top-k: the independent definition the tree is
// checked against.
static void brute_knn(const double* pts, int n, const double* q, int nq, int k, int* idx, double* d2out) {
    std::vector<std::pair<double, int>> all(n);
    for (int qi = 0; qi < nq; ++qi) {
        for (int i = 0; i < n; ++i) all[i] = {dist2(pts + 3 * (size_t)i, q + 3 * (size_t)qi), i};
        int m = std::min(k, n);
        std::partial_sort(all.begin(), all.begin() + m, all.end());
        for (int j = 0; j < k; ++j) {
            idx[(size_t)qi * k + j] = j < m ? all[j].second : -1;
            if (d2out) d2out[(size_t)qi * k + j] = j < m ? all[j].first : std::numeric_limits<double>::max();
        }
    }
}

// ---------------------------------------------------------------------------
// 3x3 symmetric eigen-decomposition: cyclic Jacobi, fixed operation order so a
// device restatement can follow it op for op.  Returns eigenvalues w[3]
// (unsorted) and eigenvectors as the COLUMNS of V (V[r][c]).
// ---------------------------------------------------------------------------
static void jacobi3(const double Ain[3][3], double w[3], double V[3][3]) {
    double A[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { A[i][j] = Ain[i][j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
    const int P[3] = {0, 0, 1}, Q[3] = {1, 2, 2};
    for (int sweep = 0; sweep < 12; ++sweep) {
        // Converged once the off-diagonal mass is below 1e-25 of the diagonal mass: every further rotation has
        // c == 1.0 and s * v below half an ulp of what it is added to, i.e. it is the identity in fp64 (checked
        // against iterating until the off-diagonals are exactly zero: same bits on 2e5 covariance matrices, 3.9
        // sweeps instead of 5.7).  The CUDA path (forest.cu jacobi3) applies the same test.
        double off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
        if (off <= 1.0e-25 * ((std::fabs(A[0][0]) + std::fabs(A[1][1])) + std::fabs(A[2][2]))) break;
        for (int r = 0; r < 3; ++r) {
            int p = P[r], q = Q[r];
            double apq = A[p][q];
            if (apq == 0.0) continue;
            double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
            double t = 1.0 / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            if (theta < 0.0) t = -t;
            double c = 1.0 / std::sqrt(t * t + 1.0);
            double s = t * c;
            // A <- J^T A J  for the (p,q) plane
            double app = A[p][p], aqq = A[q][q];
            A[p][p] = app - t * apq;
            A[q][q] = aqq + t * apq;
            A[p][q] = 0.0; A[q][p] = 0.0;
            int k = 3 - p - q;  // the remaining index
            double akp = A[k][p], akq = A[k][q];
            A[k][p] = c * akp - s * akq; A[p][k] = A[k][p];
            A[k][q] = s * akp + c * akq; A[q][k] = A[k][q];
            for (int i = 0; i < 3; ++i) {
                double vip = V[i][p], viq = V[i][q];
                V[i][p] = c * vip - s * viq;
                V[i][q] = s * vip + c * viq;
            }
        }
    }
    w[0] = A[0][0]; w[1] = A[1][1]; w[2] = A[2][2];
}

// ---------------------------------------------------------------------------
// estimate_normals — icp.hpp:23-67.  evals_out (optional, 3 per point,
// ascending) lets tests apply the eigen-gap filter of SURVEY.md H3.
// ---------------------------------------------------------------------------
static void estimate_normals(const KDTree& tree, const double* pts, int n, int k, double* normals, double* evals_out) {
    std::vector<int> nb(std::max(k, 1));
    for (int i = 0; i < n; ++i) {
        const double* q = pts + 3 * (size_t)i;
        int m = tree.k_nearest(q, k, nb.data(), nullptr);
        double* nrm = normals + 3 * (size_t)i;
        if (m < 3) {  // icp.hpp:34-37
            nrm[0] = 0; nrm[1] = 0; nrm[2] = 1;
            if (evals_out) { evals_out[3 * (size_t)i] = evals_out[3 * (size_t)i + 1] = evals_out[3 * (size_t)i + 2] = 0; }
            continue;
        }
        double c[3] = {0, 0, 0};
        for (int j = 0; j < m; ++j) {  // icp.hpp:40-44
            const double* p = pts + 3 * (size_t)nb[j];
            c[0] += p[0]; c[1] += p[1]; c[2] += p[2];
        }
        double md = static_cast<double>(m);
        c[0] /= md; c[1] /= md; c[2] /= md;
        double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int j = 0; j < m; ++j) {  // icp.hpp:47-52
            const double* p = pts + 3 * (size_t)nb[j];
            double d[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) C[a][b] += d[a] * d[b];
        }
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) C[a][b] /= md;
        double w[3], V[3][3];
        jacobi3(C, w, V);
        int s = 0;  // smallest eigenvalue, first index on ties (icp.hpp:56 col(0))
        if (w[1] < w[s]) s = 1;
        if (w[2] < w[s]) s = 2;
        double v[3] = {V[0][s], V[1][s], V[2][s]};
        if (v[2] < 0) { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; }  // icp.hpp:59-61
        double nn = std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);  // icp.hpp:63
        nrm[0] = v[0] / nn; nrm[1] = v[1] / nn; nrm[2] = v[2] / nn;
        if (evals_out) {
            double e[3] = {w[0], w[1], w[2]};
            std::sort(e, e + 3);
            evals_out[3 * (size_t)i] = e[0]; evals_out[3 * (size_t)i + 1] = e[1]; evals_out[3 * (size_t)i + 2] = e[2];
        }
    }
}

// ---------------------------------------------------------------------------
// Rigid transforms — types.hpp:74-136, row-major 4x4.
// ---------------------------------------------------------------------------
static void mat4_identity(double T[16]) {
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
}
static void mat4_mul(const double A[16], const double B[16], double C[16]) {  // types.hpp:118-120
    double R[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
            R[4 * i + j] = s;
        }
    std::memcpy(C, R, sizeof(R));
}
// P <- P * R^T + t (row-wise) — types.hpp:110-115, icp.hpp:174-176,225-226
static void apply_rt(const double T[16], double* pts, int n) {
    for (int i = 0; i < n; ++i) {
        double* p = pts + 3 * (size_t)i;
        double x = p[0], y = p[1], z = p[2];
        p[0] = ((x * T[0] + y * T[1]) + z * T[2]) + T[3];
        p[1] = ((x * T[4] + y * T[5]) + z * T[6]) + T[7];
        p[2] = ((x * T[8] + y * T[9]) + z * T[10]) + T[11];
    }
}

// 6x6 symmetric solve by LDL^T with diagonal pivoting, the decomposition behind (J^T J).ldlt().solve(J^T b)
// (icp.hpp:120).  At every step the largest remaining diagonal entry becomes the pivot (first one on ties), rows and
// columns are swapped symmetrically, and a pivot that is exactly zero leaves its component of the solution at
// zero — the behaviour of Eigen's LDLT on a rank-deficient J^T J (a planar target: every normal (0, 0, 1); fewer
// than six points), where an unpivoted factorisation divides 0 by 0.  Same operation order as csrc/icp.cu.
static void ldlt6_solve(const double Ain[6][6], const double bin[6], double x[6]) {
    double a[6][6];
    int perm[6];
    for (int i = 0; i < 6; ++i) {
        perm[i] = i;
        for (int j = 0; j < 6; ++j) a[i][j] = Ain[i][j];
    }
    for (int k = 0; k < 6; ++k) {
        int p = k;
        double best = std::fabs(a[k][k]);
        for (int i = k + 1; i < 6; ++i)
            if (std::fabs(a[i][i]) > best) { best = std::fabs(a[i][i]); p = i; }
        if (p != k) {
            for (int j = 0; j < 6; ++j) { double t = a[k][j]; a[k][j] = a[p][j]; a[p][j] = t; }
            for (int i = 0; i < 6; ++i) { double t = a[i][k]; a[i][k] = a[i][p]; a[i][p] = t; }
            int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        }
        const double d = a[k][k];
        if (d == 0.0) continue;
        for (int i = k + 1; i < 6; ++i) a[i][k] /= d;  // column k of L
        for (int i = k + 1; i < 6; ++i)
            for (int j = k + 1; j <= i; ++j) {
                a[i][j] -= a[i][k] * d * a[j][k];
                a[j][i] = a[i][j];
            }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) y[i] = bin[perm[i]];
    for (int i = 0; i < 6; ++i)  // L y = P b
        for (int j = 0; j < i; ++j) y[i] -= a[i][j] * y[j];
    for (int i = 0; i < 6; ++i) y[i] = a[i][i] == 0.0 ? 0.0 : y[i] / a[i][i];  // D z = y
    for (int i = 5; i >= 0; --i)  // L^T w = z
        for (int j = i + 1; j < 6; ++j) y[i] -= a[j][i] * y[j];
    for (int i = 0; i < 6; ++i) x[perm[i]] = y[i];
}

// solve_point_to_plane — icp.hpp:89-144
static void solve_point_to_plane(const double* src, const double* tgt, const double* nrm, int n, double T[16]) {
    double A[6][6] = {{0}}, g[6] = {0};
    for (int i = 0; i < n; ++i) {
        const double* p = src + 3 * (size_t)i;
        const double* q = tgt + 3 * (size_t)i;
        const double* nn = nrm + 3 * (size_t)i;
        double J[6];
        J[0] = p[1] * nn[2] - p[2] * nn[1];  // p x n (icp.hpp:105)
        J[1] = p[2] * nn[0] - p[0] * nn[2];
        J[2] = p[0] * nn[1] - p[1] * nn[0];
        J[3] = nn[0]; J[4] = nn[1]; J[5] = nn[2];
        double b = ((q[0] - p[0]) * nn[0] + (q[1] - p[1]) * nn[1]) + (q[2] - p[2]) * nn[2];  // icp.hpp:116
        for (int a = 0; a < 6; ++a) {
            for (int c = 0; c < 6; ++c) A[a][c] += J[a] * J[c];
            g[a] += J[a] * b;
        }
    }
    double x[6];
    ldlt6_solve(A, g, x);
    double angle = std::sqrt((x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]);  // icp.hpp:127
    double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    if (!(angle < 1e-10)) {  // icp.hpp:130-142
        double ax = x[0] / angle, ay = x[1] / angle, az = x[2] / angle;
        double K[3][3] = {{0, -az, ay}, {az, 0, -ax}, {-ay, ax, 0}};
        double K2[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += K[i][k] * K[k][j];
                K2[i][j] = s;
            }
        double sn = std::sin(angle), cs = 1 - std::cos(angle);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[i][j] = (R[i][j] + sn * K[i][j]) + cs * K2[i][j];
    }
    mat4_identity(T);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[4 * i + j] = R[i][j];
        T[4 * i + 3] = x[3 + i];
    }
}

// ---------------------------------------------------------------------------
// icp_point_to_plane — icp.hpp:157-258
// faithful_cost != 0 repeats the nearest-neighbour search exactly as the
// reference does (twice per iteration, icp.hpp:185,190) so CPU timing matches
// the reference's cost; the numerical result is identical either way.
// ---------------------------------------------------------------------------
struct ICPOut {
    double T[16];
    int converged;
    int num_iterations;
    double final_error;
    std::vector<double> history;
};

static double icp_rms(const double* cur, int ns, const KDTree& tree, const double* normals,
                      const std::vector<int>& idx) {
    double e = 0;
    for (int i = 0; i < ns; ++i) {  // icp.hpp:198-206
        const double* p = cur + 3 * (size_t)i;
        const double* q = tree.point(idx[i]);
        const double* nn = normals + 3 * (size_t)idx[i];
        double pd = ((q[0] - p[0]) * nn[0] + (q[1] - p[1]) * nn[1]) + (q[2] - p[2]) * nn[2];
        e += pd * pd;
    }
    return std::sqrt(e / ns);
}

static void icp_point_to_plane(const double* src, int ns, const double* tgt, int nt, int max_iterations,
                               double tolerance, double min_error, const double T0[16], int normals_k,
                               int faithful_cost, ICPOut& out) {
    KDTree tree(tgt, nt);  // icp.hpp:166
    std::vector<double> normals(3 * (size_t)nt);
    estimate_normals(tree, tgt, nt, normals_k, normals.data(), nullptr);  // icp.hpp:169-171
    std::vector<double> cur(src, src + 3 * (size_t)ns);
    apply_rt(T0, cur.data(), ns);  // icp.hpp:174-176
    double total[16];
    std::memcpy(total, T0, sizeof(total));
    double prev_error = std::numeric_limits<double>::max();
    out.converged = 0;
    out.history.clear();
    std::vector<int> idx(ns), idx2(ns);
    std::vector<double> d2(ns);
    std::vector<double> matched(3 * (size_t)ns), mnorm(3 * (size_t)ns);
    for (int iter = 0; iter < max_iterations; ++iter) {
        tree.nearest_batch(cur.data(), ns, idx.data(), d2.data());               // icp.hpp:185
        if (faithful_cost) tree.nearest_batch(cur.data(), ns, idx2.data(), d2.data());  // icp.hpp:190
        for (int i = 0; i < ns; ++i) {
            std::memcpy(&matched[3 * (size_t)i], tree.point(idx[i]), 24);
            std::memcpy(&mnorm[3 * (size_t)i], &normals[3 * (size_t)idx[i]], 24);
        }
        double error = icp_rms(cur.data(), ns, tree, normals.data(), idx);
        out.history.push_back(error);
        if (error < min_error) { out.converged = 1; break; }                     // icp.hpp:210-213
        if (std::fabs(prev_error - error) < tolerance) { out.converged = 1; break; }  // icp.hpp:214-217
        double delta[16];
        solve_point_to_plane(cur.data(), matched.data(), mnorm.data(), ns, delta);  // icp.hpp:220
        apply_rt(delta, cur.data(), ns);                                         // icp.hpp:225-226
        mat4_mul(delta, total, total);                                           // icp.hpp:229
        prev_error = error;
    }
    tree.nearest_batch(cur.data(), ns, idx.data(), d2.data());                   // icp.hpp:238
    if (faithful_cost) tree.nearest_batch(cur.data(), ns, idx2.data(), d2.data());      // icp.hpp:242
    out.final_error = icp_rms(cur.data(), ns, tree, normals.data(), idx);
    out.history.push_back(out.final_error);
    std::memcpy(out.T, total, sizeof(total));
    out.num_iterations = static_cast<int>(out.history.size()) - 1;               // icp.hpp:255
}

// ---------------------------------------------------------------------------
// ScanContext — scan_context.hpp:24-145.  Descriptor stored COLUMN-MAJOR
// (Eigen::MatrixXd default): element (ring i, sector j) at desc[j*20 + i].
// ---------------------------------------------------------------------------
constexpr int SC_RINGS = 20, SC_SECTORS = 60;
constexpr double SC_MAX_RANGE = 80.0;

static void sc_compute(const double* xyz, i64 n, double* desc) {
    for (int t = 0; t < SC_RINGS * SC_SECTORS; ++t) desc[t] = -std::numeric_limits<double>::max();
    double ring_size = SC_MAX_RANGE / SC_RINGS;
    double sector_size = 2.0 * M_PI / SC_SECTORS;
    for (i64 i = 0; i < n; ++i) {  // scan_context.hpp:50-74
        double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        double range = std::sqrt(x * x + y * y);
        double angle = std::atan2(y, x) + M_PI;
        if (range > SC_MAX_RANGE || range < 0.1) continue;
        int ring_idx = static_cast<int>(range / ring_size);
        int sector_idx = static_cast<int>(angle / sector_size);
        ring_idx = std::clamp(ring_idx, 0, SC_RINGS - 1);
        sector_idx = std::clamp(sector_idx, 0, SC_SECTORS - 1);
        double& cell = desc[sector_idx * SC_RINGS + ring_idx];
        if (z > cell) cell = z;
    }
    for (int t = 0; t < SC_RINGS * SC_SECTORS; ++t)
        if (desc[t] < -1000) desc[t] = 0;  // scan_context.hpp:77-83
}

static double sc_shifted(const double* A, const double* B, int shift) {  // scan_context.hpp:121-142
    double sum_ab = 0, sum_aa = 0, sum_bb = 0;
    for (int i = 0; i < SC_RINGS; ++i)
        for (int j = 0; j < SC_SECTORS; ++j) {
            double a = A[j * SC_RINGS + i];
            double b = B[((j + shift) % SC_SECTORS) * SC_RINGS + i];
            sum_ab += a * b;
            sum_aa += a * a;
            sum_bb += b * b;
        }
    double norm = std::sqrt(sum_aa) * std::sqrt(sum_bb);
    if (norm < 1e-10) return 1.0;
    return 1.0 - sum_ab / norm;
}

static double sc_distance(const double* A, const double* B) {  // scan_context.hpp:90-102
    double best = std::numeric_limits<double>::max();
    for (int s = 0; s < SC_SECTORS; ++s) {
        double d = sc_shifted(A, B, s);
        if (d < best) best = d;
    }
    return best;
}

// ---------------------------------------------------------------------------
// LoopClosureDetector — loop_closure.hpp:41-149
// ---------------------------------------------------------------------------
struct LoopResult {
    int query_frame, match_frame;
    double T[16];
    double sc_distance, icp_fitness;
};

struct LoopDetector {
    int frame_gap = 50;
    double sc_thr = 0.25, icp_thr = 0.3;
    int max_candidates = 3;
    int icp_max_iterations = 30;   // loop_closure.hpp:106
    double icp_tolerance = 1e-6;   // loop_closure.hpp:107
    int normals_k = 20;            // icp.hpp:170
    std::vector<std::vector<double>> desc;
    std::vector<std::vector<double>> clouds;
    std::vector<int> frames;

    void add(const double* xyz, int n, int frame_idx) {  // loop_closure.hpp:53-59
        std::vector<double> d(SC_RINGS * SC_SECTORS);
        sc_compute(xyz, n, d.data());
        desc.push_back(std::move(d));
        clouds.emplace_back(xyz, xyz + 3 * (size_t)n);
        frames.push_back(frame_idx);
    }
    // candidate list (dist, db index) ascending — loop_closure.hpp:75-92
    void candidates(std::vector<std::pair<double, int>>& c) const {
        c.clear();
        if (desc.size() < 2) return;
        size_t q = desc.size() - 1;
        for (size_t i = 0; i + 1 < desc.size(); ++i) {
            if (frames[q] - frames[i] < frame_gap) continue;
            double d = sc_distance(desc[q].data(), desc[i].data());
            if (d < sc_thr) c.push_back({d, static_cast<int>(i)});
        }
        std::sort(c.begin(), c.end());
    }
    void detect(std::vector<LoopResult>& res) const {  // loop_closure.hpp:66-126
        res.clear();
        std::vector<std::pair<double, int>> c;
        candidates(c);
        if (c.empty()) return;
        size_t q = desc.size() - 1;
        double I[16];
        mat4_identity(I);
        int verified = 0;
        for (const auto& pr : c) {
            if (verified >= max_candidates) break;
            const auto& qc = clouds[q];
            const auto& cc = clouds[pr.second];
            ICPOut o;
            icp_point_to_plane(qc.data(), (int)(qc.size() / 3), cc.data(), (int)(cc.size() / 3), icp_max_iterations,
                               icp_tolerance, 1e-9, I, normals_k, 0, o);
            if (o.converged && o.final_error < icp_thr) {
                LoopResult r;
                r.query_frame = frames[q];
                r.match_frame = frames[pr.second];
                std::memcpy(r.T, o.T, sizeof(r.T));
                r.sc_distance = pr.first;
                r.icp_fitness = o.final_error;
                res.push_back(r);
                ++verified;
            }
        }
    }
};

// ===========================================================================
// The steps right after the path in slam_node.cpp: world-frame clouds, occupancy cells, global map
// ===========================================================================
// world = cloud * R^T + t, row by row (slam_viz/src/ros/slam_node.cpp:147, 189, 201-203)
static void transform_cloud(const double* xyz, i64 n, const double T[16], double* out) {
    for (i64 i = 0; i < n; ++i) {
        const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        for (int a = 0; a < 3; ++a) out[3 * i + a] = ((x * T[4 * a] + y * T[4 * a + 1]) + z * T[4 * a + 2]) + T[4 * a + 3];
    }
}

// update_occupancy_grid (slam_node.cpp:211-221) into an ordered set of (x, y) cells
static void occupancy_insert(const double* world, i64 n, const double sensor[3], double res, double hmin, double hmax,
                             double max_range, std::vector<std::pair<int, int>>& cells) {
    for (i64 i = 0; i < n; ++i) {
        const double x = world[3 * i], y = world[3 * i + 1], z = world[3 * i + 2];
        if (z < hmin || z > hmax) continue;
        const double dx = x - sensor[0], dy = y - sensor[1];
        const double r = std::sqrt(dx * dx + dy * dy);
        if (r > max_range || r < 0.5) continue;
        cells.emplace_back(static_cast<int>(std::floor(x / res)), static_cast<int>(std::floor(y / res)));
    }
}

}  // namespace orc

// ===========================================================================
// C ABI for ctypes (tests / smoke / bench cpu_baseline only)
// ===========================================================================
extern "C" {

long long orc_voxel_downsample(const double* xyz, long long n, double voxel, double* out_xyz, long long* out_keys) {
    return orc::voxel_downsample(xyz, n, voxel, out_xyz, out_keys);
}

void* orc_kdtree_build(const double* xyz, int n) { return new orc::KDTree(xyz, n); }
void orc_kdtree_free(void* t) { delete static_cast<orc::KDTree*>(t); }
void orc_kdtree_nearest_batch(void* t, const double* q, int nq, int* idx, double* d2) {
    static_cast<orc::KDTree*>(t)->nearest_batch(q, nq, idx, d2);
}
// out: nq*k ints, padded with -1; d2 optional
void orc_kdtree_k_nearest_batch(void* t, const double* q, int nq, int k, int* out, double* d2) {
    auto* tree = static_cast<orc::KDTree*>(t);
    std::vector<int> tmp(std::max(k, 1));
    std::vector<double> td(std::max(k, 1));
    for (int i = 0; i < nq; ++i) {
        int m = tree->k_nearest(q + 3 * (size_t)i, k, tmp.data(), td.data());
        for (int j = 0; j < k; ++j) {
            out[(size_t)i * k + j] = j < m ? tmp[j] : -1;
            if (d2) d2[(size_t)i * k + j] = j < m ? td[j] : std::numeric_limits<double>::max();
        }
    }
}
void orc_brute_knn(const double* pts, int n, const double* q, int nq, int k, int* idx, double* d2) {
    orc::brute_knn(pts, n, q, nq, k, idx, d2);
}
void orc_estimate_normals(void* t, const double* pts, int n, int k, double* normals, double* evals) {
    orc::estimate_normals(*static_cast<orc::KDTree*>(t), pts, n, k, normals, evals);
}
void orc_jacobi3(const double* A9, double* w3, double* V9) {
    double A[3][3], V[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = A9[3 * i + j];
    orc::jacobi3(A, w3, V);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V9[3 * i + j] = V[i][j];
}
void orc_ldlt6_solve(const double* A36, const double* b6, double* x6) {
    double A[6][6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) A[i][j] = A36[6 * i + j];
    orc::ldlt6_solve(A, b6, x6);
}
void orc_solve_point_to_plane(const double* src, const double* tgt, const double* nrm, int n, double* T16) {
    orc::solve_point_to_plane(src, tgt, nrm, n, T16);
}
// history must hold max_iterations+1 doubles; returns history length
int orc_icp_point_to_plane(const double* src, int ns, const double* tgt, int nt, int max_iterations, double tolerance,
                           double min_error, const double* T0, int normals_k, int faithful_cost, double* T16,
                           int* converged, int* num_iterations, double* final_error, double* history) {
    orc::ICPOut o;
    double I[16];
    orc::mat4_identity(I);
    orc::icp_point_to_plane(src, ns, tgt, nt, max_iterations, tolerance, min_error, T0 ? T0 : I, normals_k,
                            faithful_cost, o);
    std::memcpy(T16, o.T, sizeof(o.T));
    *converged = o.converged;
    *num_iterations = o.num_iterations;
    *final_error = o.final_error;
    if (history) std::memcpy(history, o.history.data(), sizeof(double) * o.history.size());
    return static_cast<int>(o.history.size());
}
void orc_transform_cloud(const double* xyz, long long n, const double* T16, double* out) {
    orc::transform_cloud(xyz, n, T16, out);
}
// rebuild_occupancy_grid (slam_node.cpp:223-229): all clouds (sensor frame, CSR) with their poses -> unique cells in
// ascending (x, y) order (the reference's unordered_set has no order).  Returns the number of cells; out_cells may
// be null to count only.
long long orc_occupancy_cells(const double* xyz, const long long* offsets, int n_clouds, const double* poses16,
                              double res, double hmin, double hmax, double max_range, int* out_cells,
                              long long capacity) {
    std::vector<std::pair<int, int>> cells;
    std::vector<double> w;
    for (int c = 0; c < n_clouds; ++c) {
        const long long n = offsets[c + 1] - offsets[c];
        w.resize((size_t)(3 * n));
        const double* T = poses16 + 16 * c;
        orc::transform_cloud(xyz + 3 * offsets[c], n, T, w.data());
        const double sensor[3] = {T[3], T[7], T[11]};
        orc::occupancy_insert(w.data(), n, sensor, res, hmin, hmax, max_range, cells);
    }
    std::sort(cells.begin(), cells.end());
    cells.erase(std::unique(cells.begin(), cells.end()), cells.end());
    if (out_cells)
        for (size_t i = 0; i < cells.size() && (long long)i < capacity; ++i) {
            out_cells[2 * i] = cells[i].first;
            out_cells[2 * i + 1] = cells[i].second;
        }
    return (long long)cells.size();
}
void orc_sc_compute(const double* xyz, long long n, double* desc1200) { orc::sc_compute(xyz, n, desc1200); }
double orc_sc_distance(const double* a, const double* b) { return orc::sc_distance(a, b); }

void* orc_loop_create(int frame_gap, double sc_thr, double icp_thr, int max_candidates) {
    auto* d = new orc::LoopDetector();
    d->frame_gap = frame_gap; d->sc_thr = sc_thr; d->icp_thr = icp_thr; d->max_candidates = max_candidates;
    return d;
}
void orc_loop_free(void* d) { delete static_cast<orc::LoopDetector*>(d); }
void orc_loop_add(void* d, const double* xyz, int n, int frame_idx) { static_cast<orc::LoopDetector*>(d)->add(xyz, n, frame_idx); }
int orc_loop_size(void* d) { return (int)static_cast<orc::LoopDetector*>(d)->desc.size(); }
// candidate list: returns count (<= cap)
int orc_loop_candidates(void* d, int cap, double* dist, int* idx) {
    std::vector<std::pair<double, int>> c;
    static_cast<orc::LoopDetector*>(d)->candidates(c);
    int m = std::min<int>(cap, (int)c.size());
    for (int i = 0; i < m; ++i) { dist[i] = c[i].first; idx[i] = c[i].second; }
    return (int)c.size();
}
// results: per result {query_frame, match_frame} ints, T[16], sc_distance, icp_fitness
int orc_loop_detect(void* d, int cap, int* frames2, double* T16s, double* sc_dist, double* fitness) {
    std::vector<orc::LoopResult> r;
    static_cast<orc::LoopDetector*>(d)->detect(r);
    int m = std::min<int>(cap, (int)r.size());
    for (int i = 0; i < m; ++i) {
        frames2[2 * i] = r[i].query_frame; frames2[2 * i + 1] = r[i].match_frame;
        std::memcpy(T16s + 16 * i, r[i].T, sizeof(r[i].T));
        sc_dist[i] = r[i].sc_distance; fitness[i] = r[i].icp_fitness;
    }
    return (int)r.size();
}

}  // extern "C"
