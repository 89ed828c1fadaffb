// ref_harness.cpp -- test infrastructure.  A C interface over the REFERENCE'S OWN sources, compiled unmodified from
// /root/reference/slam_viz (never copied into this repository) against oracle/eigen_standin (Eigen is absent from
// this image; see that header for what the stand-in does and does not pin).  Built by oracle/build_ref.sh into
// oracle/_ref/libslam_ref.so; used by tests/test_reference_build.py to check the CPU oracle against the reference's
// real control flow and by tests/golden/make_reference_golden.py to produce the committed reference fixtures.
//
// Every function below is a thin call into the reference API named in its comment; no algorithm lives here.
#include "slam_viz/core/file_utils.hpp"
#include "slam_viz/core/icp.hpp"
#include "slam_viz/core/kdtree.hpp"
#include "slam_viz/core/loop_closure.hpp"
#include "slam_viz/core/scan_context.hpp"
#include "slam_viz/core/types.hpp"

#include <cstring>
#include <limits>
#include <string>

namespace {

using Mat = slam::PointCloud::Matrix;

Mat to_mat(const double* xyz, long long n) {
    Mat m(n, 3);
    if (n > 0) std::memcpy(m.data(), xyz, sizeof(double) * 3 * (size_t)n);  // row-major n x 3
    return m;
}

void put_T(const slam::Transformation& T, double* T16) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) T16[4 * i + j] = T.matrix()(i, j);
}

struct Tree {
    Mat pts;
    slam::KDTree tree;
    explicit Tree(const Mat& p) : pts(p), tree(p) {}
};

}  // namespace

extern "C" {

// slam::voxel_downsample (file_utils.cpp:148-196); rows in the unordered_map's iteration order
long long ref_voxel_downsample(const double* xyz, long long n, double voxel, double* out_xyz) {
    Mat out = slam::voxel_downsample(to_mat(xyz, n), voxel);
    if (out_xyz && out.rows() > 0) std::memcpy(out_xyz, out.data(), sizeof(double) * 3 * (size_t)out.rows());
    return (long long)out.rows();
}

// slam::load_bin / slam::load_ply (file_utils.cpp:20-141): float32 records widened to double
long long ref_load_points(const char* path, int is_bin, double* out_xyz, long long capacity) {
    Mat m = is_bin ? slam::load_bin(path) : slam::load_ply(path);
    if (out_xyz && m.rows() <= capacity && m.rows() > 0)
        std::memcpy(out_xyz, m.data(), sizeof(double) * 3 * (size_t)m.rows());
    return (long long)m.rows();
}

// slam::KDTree (kdtree.hpp:18-186)
void* ref_kdtree_build(const double* xyz, int n) { return new Tree(to_mat(xyz, n)); }
void ref_kdtree_free(void* t) { delete static_cast<Tree*>(t); }
void ref_kdtree_nearest_batch(void* t, const double* q, int nq, int* idx, double* d2) {
    std::vector<int> i;
    std::vector<double> d;
    static_cast<Tree*>(t)->tree.nearest_batch(to_mat(q, nq), i, d);
    for (int k = 0; k < nq; ++k) {
        idx[k] = i[(size_t)k];
        d2[k] = d[(size_t)k];
    }
}
// KDTree::nearest (kdtree.hpp:32-37)
int ref_kdtree_nearest(void* t, const double* q3) {
    return static_cast<Tree*>(t)->tree.nearest(Eigen::Vector3d(q3[0], q3[1], q3[2]));
}
// KDTree::k_nearest (kdtree.hpp:65-78) per query; out padded with -1
void ref_kdtree_k_nearest_batch(void* t, const double* q, int nq, int k, int* out) {
    for (int i = 0; i < nq; ++i) {
        std::vector<int> r = static_cast<Tree*>(t)->tree.k_nearest(Eigen::Vector3d(q[3 * i], q[3 * i + 1], q[3 * i + 2]), k);
        for (int j = 0; j < k; ++j) out[(size_t)i * k + j] = j < (int)r.size() ? r[(size_t)j] : -1;
    }
}
// NearestNeighborSearch::find_correspondences (kdtree.hpp:198-214)
void ref_find_correspondences(const double* tgt, int nt, const double* src, int ns, double* matched, double* dist) {
    slam::NearestNeighborSearch nn{slam::PointCloud(to_mat(tgt, nt))};
    Mat m;
    Eigen::VectorXd d;
    nn.find_correspondences(to_mat(src, ns), m, d);
    std::memcpy(matched, m.data(), sizeof(double) * 3 * (size_t)ns);
    for (int i = 0; i < ns; ++i) dist[i] = d(i);
}
// slam::estimate_normals (icp.hpp:23-67)
void ref_estimate_normals(void* t, int k, double* normals) {
    Tree* T = static_cast<Tree*>(t);
    Mat nrm = slam::estimate_normals(T->pts, T->tree, k);
    std::memcpy(normals, nrm.data(), sizeof(double) * 3 * (size_t)nrm.rows());
}
// slam::solve_point_to_plane (icp.hpp:89-144)
void ref_solve_point_to_plane(const double* src, const double* tgt, const double* nrm, int n, double* T16) {
    put_T(slam::solve_point_to_plane(to_mat(src, n), to_mat(tgt, n), to_mat(nrm, n)), T16);
}
// slam::icp_point_to_plane (icp.hpp:157-258); history must hold max_iterations + 1 doubles; returns its length
int ref_icp_point_to_plane(const double* src, int ns, const double* tgt, int nt, int max_iterations, double tolerance,
                           double min_error, const double* T0, double* T16, int* converged, int* num_iterations,
                           double* final_error, double* history) {
    slam::ICPConfig cfg;
    cfg.max_iterations = max_iterations;
    cfg.tolerance = tolerance;
    cfg.min_error = min_error;
    if (T0) {
        Eigen::Matrix4d M;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) M(i, j) = T0[4 * i + j];
        cfg.initial_transform = slam::Transformation(M);
    }
    slam::ICPResult r = slam::icp_point_to_plane(slam::PointCloud(to_mat(src, ns)), slam::PointCloud(to_mat(tgt, nt)), cfg);
    put_T(r.transformation, T16);
    *converged = r.converged ? 1 : 0;
    *num_iterations = r.num_iterations;
    *final_error = r.final_error;
    for (size_t i = 0; i < r.error_history.size(); ++i) history[i] = r.error_history[i];
    return (int)r.error_history.size();
}
// Transformation::apply / compose / inverse (types.hpp:105-133)
void ref_transform_apply(const double* T16, const double* xyz, long long n, double* out) {
    Eigen::Matrix4d M;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) M(i, j) = T16[4 * i + j];
    slam::PointCloud c = slam::Transformation(M).apply(slam::PointCloud(to_mat(xyz, n)));
    if (n > 0) std::memcpy(out, c.points().data(), sizeof(double) * 3 * (size_t)n);
}
void ref_transform_compose_inverse(const double* A16, const double* B16, double* AB16, double* Ainv16) {
    Eigen::Matrix4d A, B;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            A(i, j) = A16[4 * i + j];
            B(i, j) = B16[4 * i + j];
        }
    put_T(slam::Transformation(A) * slam::Transformation(B), AB16);
    put_T(slam::Transformation(A).inverse(), Ainv16);
}
// ScanContext::compute / descriptor (scan_context.hpp:44-82, 118); desc1200 in the storage order of the
// reference's Eigen::MatrixXd (column-major: element (ring, sector) at sector * 20 + ring), as the C ABI has it
void ref_sc_compute(const double* xyz, long long n, double* desc1200) {
    slam::ScanContext sc(to_mat(xyz, n));
    for (int i = 0; i < slam::ScanContext::NUM_RINGS; ++i)
        for (int j = 0; j < slam::ScanContext::NUM_SECTORS; ++j)
            desc1200[j * slam::ScanContext::NUM_RINGS + i] = sc.descriptor()(i, j);
}
// ScanContext::distance (scan_context.hpp:90-102) between the descriptors of two clouds
double ref_sc_distance_clouds(const double* a, long long na, const double* b, long long nb) {
    return slam::ScanContext(to_mat(a, na)).distance(slam::ScanContext(to_mat(b, nb)));
}
// ScanContext::ring_key / sector_key (scan_context.hpp:107-116)
void ref_sc_keys(const double* xyz, long long n, double* ring20, double* sector60) {
    slam::ScanContext sc(to_mat(xyz, n));
    Eigen::VectorXd r = sc.ring_key(), s = sc.sector_key();
    for (int i = 0; i < 20; ++i) ring20[i] = r(i);
    for (int j = 0; j < 60; ++j) sector60[j] = s(j);
}
// LoopClosureDetector (loop_closure.hpp:41-149)
void* ref_loop_create(int frame_gap, double sc_thr, double icp_thr, int max_candidates) {
    slam::LoopClosureConfig c;
    c.frame_gap = frame_gap;
    c.sc_distance_threshold = sc_thr;
    c.icp_fitness_threshold = icp_thr;
    c.max_candidates = max_candidates;
    return new slam::LoopClosureDetector(c);
}
void ref_loop_free(void* d) { delete static_cast<slam::LoopClosureDetector*>(d); }
void ref_loop_add(void* d, const double* xyz, int n, int frame_idx) {
    static_cast<slam::LoopClosureDetector*>(d)->addFrame(to_mat(xyz, n), frame_idx);
}
int ref_loop_size(void* d) { return (int)static_cast<slam::LoopClosureDetector*>(d)->size(); }
void ref_loop_clear(void* d) { static_cast<slam::LoopClosureDetector*>(d)->clear(); }
// results: per result {query_frame, match_frame} ints, T[16], sc_distance, icp_fitness; returns the count
int ref_loop_detect(void* d, int cap, int* frames2, double* T16s, double* sc_dist, double* fitness) {
    std::vector<slam::LoopClosureResult> r = static_cast<slam::LoopClosureDetector*>(d)->detect();
    for (int i = 0; i < (int)r.size() && i < cap; ++i) {
        frames2[2 * i] = r[(size_t)i].query_frame;
        frames2[2 * i + 1] = r[(size_t)i].match_frame;
        put_T(r[(size_t)i].transform, T16s + 16 * i);
        sc_dist[i] = r[(size_t)i].scan_context_distance;
        fitness[i] = r[(size_t)i].icp_fitness;
    }
    return (int)r.size();
}

}  // extern "C"
