// host_api_test.cpp — the reference-shaped C++17 API (lidar-slam-from-scratch_b200/host/slam_viz/core/*.hpp) driven
// the way slam_node.cpp drives the reference (slam_viz/src/ros/slam_node.cpp:118-175), checked against the CPU
// oracle (oracle/liboracle.so).  TEST CODE: links the oracle as the checker only.  Needs a B200; run by
// tests/test_cpp_host.py (-m gpu).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "slam_viz/core/file_utils.hpp"
#include "slam_viz/core/icp.hpp"
#include "slam_viz/core/loop_closure.hpp"
#include "slam_viz/core/scan_context.hpp"

extern "C" {
// synth/libsynth.so
int syn_scene(unsigned long long seed, int n_boxes, float half_extent, int path_kind, float radius, float corridor_half,
              float sensor_height, float* boxes6);
long long syn_scan(int beams, int azimuth_steps, float elev_top_deg, float elev_bot_deg, float max_range,
                   float noise_sigma, float sensor_height, const float* boxes6, int n_boxes, double x, double y,
                   double yaw, unsigned long long noise_seed, double* out_xyz, int n_threads);
// oracle/liboracle.so
long long orc_voxel_downsample(const double* xyz, long long n, double voxel, double* out_xyz, long long* out_keys);
void* orc_kdtree_build(const double* xyz, int n);
void orc_kdtree_free(void* t);
void orc_kdtree_nearest_batch(void* t, const double* q, int nq, int* idx, double* d2);
void orc_kdtree_k_nearest_batch(void* t, const double* q, int nq, int k, int* out, double* d2);
int orc_icp_point_to_plane(const double* src, int ns, const double* tgt, int nt, int max_iterations, double tolerance,
                           double min_error, const double* T0, int normals_k, int faithful_cost, double* T16,
                           int* converged, int* num_iterations, double* final_error, double* history);
void orc_sc_compute(const double* xyz, long long n, double* desc1200);
double orc_sc_distance(const double* a, const double* b);
}

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);    \
            ++failures;                                                    \
        }                                                                  \
    } while (0)

using Matrix = slam::PointCloud::Matrix;

static Matrix scan(const std::vector<float>& boxes, double x, double y, double yaw, unsigned long long seed) {
    const int beams = 32, az = 600;
    std::vector<double> buf((size_t)3 * beams * az);
    long long n = syn_scan(beams, az, 2.0f, -24.8f, 120.0f, 0.02f, 1.73f, boxes.data(), (int)boxes.size() / 6, x, y, yaw,
                           seed, buf.data(), 4);
    Matrix m((long)n, 3);
    for (long long i = 0; i < 3 * n; ++i) m.data()[i] = buf[(size_t)i];
    return m;
}

int main() {
    std::vector<float> boxes((size_t)6 * 400);
    int nb = syn_scene(1, 400, 90.0f, 0, 0.0f, 6.0f, 1.73f, boxes.data());
    boxes.resize((size_t)6 * nb);
    Matrix raw_prev = scan(boxes, 0.0, 0.0, 0.0, 7), raw_curr = scan(boxes, 1.0, 0.1, 0.01, 8);

    // ---- slam_node.cpp:122 voxel_downsample
    Matrix prev = slam::voxel_downsample(raw_prev, 0.5), curr = slam::voxel_downsample(raw_curr, 0.5);
    {
        std::vector<double> ref((size_t)3 * raw_prev.rows());
        std::vector<long long> keys((size_t)3 * raw_prev.rows());
        long long m = orc_voxel_downsample(raw_prev.data(), raw_prev.rows(), 0.5, ref.data(), keys.data());
        CHECK(m == prev.rows());
        bool same = m == prev.rows();
        for (long long i = 0; same && i < 3 * m; ++i) same = ref[(size_t)i] == prev.data()[i];
        CHECK(same);  // bit-exact centroids in the canonical (kx, ky, kz) order
    }

    // ---- KDTree (kdtree.hpp): nearest, nearest_batch, k_nearest
    slam::KDTree tree(prev);
    {
        void* ot = orc_kdtree_build(prev.data(), (int)prev.rows());
        std::vector<int> gi, oi((size_t)curr.rows());
        std::vector<double> gd, od((size_t)curr.rows());
        tree.nearest_batch(curr, gi, gd);
        orc_kdtree_nearest_batch(ot, curr.data(), (int)curr.rows(), oi.data(), od.data());
        CHECK(gi == oi);
        CHECK(gd == od);
        Eigen::Vector3d q(curr(5, 0), curr(5, 1), curr(5, 2));
        CHECK(tree.nearest(q) == oi[5]);
        std::vector<int> kn = tree.k_nearest(q, 20), okn(20);
        std::vector<double> okd(20);
        orc_kdtree_k_nearest_batch(ot, curr.data() + 15, 1, 20, okn.data(), okd.data());
        CHECK(kn == okn);
        orc_kdtree_free(ot);
        slam::NearestNeighborSearch nn{slam::PointCloud(prev)};
        Matrix matched;
        Eigen::VectorXd dist;
        nn.find_correspondences(curr, matched, dist);
        bool ok = matched.rows() == curr.rows();
        for (long i = 0; ok && i < curr.rows(); ++i)
            ok = matched(i, 0) == prev(oi[(size_t)i], 0) && matched(i, 2) == prev(oi[(size_t)i], 2) &&
                 dist(i) == std::sqrt(od[(size_t)i]);
        CHECK(ok);
        Matrix normals = slam::estimate_normals(prev, tree, 20);
        CHECK(normals.rows() == prev.rows());
        double worst = 0.0;
        for (long i = 0; i < normals.rows(); ++i)
            worst = std::fmax(worst, std::fabs(std::sqrt(normals(i, 0) * normals(i, 0) + normals(i, 1) * normals(i, 1) +
                                                        normals(i, 2) * normals(i, 2)) - 1.0));
        CHECK(worst < 1e-12);
    }

    // ---- slam_node.cpp:132-142 icp_point_to_plane(source = current, target = previous) and pose composition
    slam::ICPConfig cfg;
    cfg.max_iterations = 50;
    cfg.tolerance = 1e-6;
    slam::ICPResult res = slam::icp_point_to_plane(slam::PointCloud(curr), slam::PointCloud(prev), cfg);
    {
        double T0[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, T[16], hist[256], fe = 0;
        int conv = 0, iters = 0;
        int hl = orc_icp_point_to_plane(curr.data(), (int)curr.rows(), prev.data(), (int)prev.rows(), 50, 1e-6, 1e-9, T0,
                                        20, 0, T, &conv, &iters, &fe, hist);
        CHECK(res.converged == (conv != 0));
        CHECK((int)res.error_history.size() == hl);
        CHECK(res.num_iterations == iters && iters == hl - 1);
        CHECK(std::fabs(res.final_error - fe) < 1e-6);
        double dt = 0.0;
        for (int i = 0; i < 3; ++i) dt = std::fmax(dt, std::fabs(res.transformation.matrix()(i, 3) - T[4 * i + 3]));
        CHECK(dt < 1e-4);  // north_star: poses within 1e-4 m
        // the sensor moved by (1.0, 0.1, yaw 0.01): T maps the current frame into the previous one
        CHECK(std::fabs(res.transformation.t()(0) - 1.0) < 0.2 && std::fabs(res.transformation.t()(1) - 0.1) < 0.1);
        slam::Transformation pose = slam::Transformation::identity() * res.transformation;  // slam_node.cpp:142
        slam::Transformation back = pose * pose.inverse();
        CHECK(std::fabs(back.matrix()(0, 3)) < 1e-12 && std::fabs(back.matrix()(1, 1) - 1.0) < 1e-12);
        std::printf("icp: %d iterations, final_error %.6f, t = (%.4f, %.4f, %.4f)\n", res.num_iterations, res.final_error,
                    res.transformation.t()(0), res.transformation.t()(1), res.transformation.t()(2));
    }

    // ---- ScanContext (scan_context.hpp) and LoopClosureDetector (loop_closure.hpp, slam_node.cpp:77-81,159-167)
    slam::ScanContext sa(prev), sb(curr);
    {
        std::vector<double> da(1200), db(1200);
        orc_sc_compute(prev.data(), prev.rows(), da.data());
        orc_sc_compute(curr.data(), curr.rows(), db.data());
        bool same = true;
        for (int i = 0; i < 1200; ++i) same = same && sa.descriptor().data()[i] == da[(size_t)i];
        CHECK(same);
        CHECK(std::fabs(sa.distance(sb) - orc_sc_distance(da.data(), db.data())) < 1e-5);
        CHECK(sa.ring_key().size() == 20 && sa.sector_key().size() == 60);
    }
    slam::LoopClosureConfig lc;
    lc.frame_gap = 2;
    lc.sc_distance_threshold = 0.5;
    slam::LoopClosureDetector det(lc);
    det.reserve(8, 100000);  // extension: size the device pools up front (must not change any result)
    det.addFrame(prev, 1);
    det.addFrame(curr, 2);
    det.addFrame(slam::voxel_downsample(scan(boxes, 30.0, 0.0, 0.0, 9), 0.5), 3);
    det.addFrame(slam::voxel_downsample(scan(boxes, 0.3, 0.0, 0.0, 10), 0.5), 4);  // revisits frame 1
    std::vector<slam::LoopClosureResult> loops = det.detect();
    CHECK(det.size() == 4);
    CHECK(!loops.empty() && loops[0].query_frame == 4 && (loops[0].match_frame == 1 || loops[0].match_frame == 2));
    det.clear();
    CHECK(det.size() == 0);

    // ---- errors come back as exceptions, not UB (kdtree.hpp:25,119 are undefined on an empty cloud)
    bool threw = false;
    try {
        slam::icp_point_to_plane(slam::PointCloud(Matrix(0, 3)), slam::PointCloud(prev));
    } catch (const std::runtime_error&) {
        threw = true;
    }
    CHECK(threw);

    std::printf(failures ? "host_api_test: %d FAILURES\n" : "host_api_test: all checks passed\n", failures);
    return failures ? 1 : 0;
}
