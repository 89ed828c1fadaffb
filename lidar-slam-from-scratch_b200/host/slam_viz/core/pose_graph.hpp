// pose_graph.hpp — the back-end interface, kept so that the whole slam_viz/core/ directory can be shadowed.
//
// The pose graph is NOT part of this engine: it stays on the host with GTSAM (BASELINE.json north_star; SURVEY.md 8f
// N4).  This header exists because the reference's own pose_graph.hpp:3 says #include "types.hpp", which a compiler
// resolves next to that file — i.e. to the reference's types.hpp — so including it after the mirror's types.hpp
// defines slam::PointCloud twice.  With this file in the mirror directory nothing under
// slam_viz/include/slam_viz/core/ is read any more, and the reference's unmodified src/core/pose_graph.cpp compiles
// against it: the class below declares exactly the members that file defines (pose_graph.hpp:22-147) — the
// implementation and every GTSAM type stay in the reference.
//
// Below the class: the hand-off from the engine's batch results to this interface (sb_pose_factor,
// include/slam_b200.h), i.e. what slam_node.cpp:145 and :163-167 do one frame at a time.
#pragma once
#include <cstddef>
#include <memory>
#include <vector>

#include "backend.hpp"
#include "types.hpp"

namespace gtsam {  // pose_graph.hpp:8-12: only named, never defined here
class NonlinearFactorGraph;
class Values;
class Pose3;
}  // namespace gtsam

namespace slam {

struct PoseGraphConfig {  // pose_graph.hpp:22-40; sigmas in the order roll, pitch, yaw, x, y, z
    double odom_rotation_sigma = 0.01;
    double odom_translation_sigma = 0.05;
    double prior_rotation_sigma = 0.001;
    double prior_translation_sigma = 0.001;
    double loop_rotation_sigma = 0.005;
    double loop_translation_sigma = 0.025;
    int max_iterations = 100;
    double relative_error_tol = 1e-5;
    double absolute_error_tol = 1e-5;
};

class PoseGraph {  // pose_graph.hpp:49-147, defined by the reference's pose_graph.cpp
public:
    explicit PoseGraph(const PoseGraphConfig& config = PoseGraphConfig());
    ~PoseGraph();
    PoseGraph(const PoseGraph&) = delete;
    PoseGraph& operator=(const PoseGraph&) = delete;
    PoseGraph(PoseGraph&&) noexcept;
    PoseGraph& operator=(PoseGraph&&) noexcept;

    void addPrior(size_t index, const Transformation& pose);
    void addOdometryFactor(size_t from_idx, size_t to_idx, const Transformation& relative_transform,
                           double fitness_score = 0.0);
    void addLoopClosure(size_t from_idx, size_t to_idx, const Transformation& relative_transform);
    bool optimize();
    Transformation getPose(size_t index) const;
    std::vector<Transformation> getAllPoses() const;
    size_t size() const { return num_poses_; }
    size_t loopClosureCount() const { return num_loop_closures_; }
    double getFinalError() const { return final_error_; }
    int getIterations() const { return iterations_; }

private:
    static gtsam::Pose3 toGtsamPose(const Transformation& t);
    static Transformation fromGtsamPose(const gtsam::Pose3& p);

    PoseGraphConfig config_;
    std::unique_ptr<gtsam::NonlinearFactorGraph> graph_;
    std::unique_ptr<gtsam::Values> initial_estimates_;
    std::unique_ptr<gtsam::Values> optimized_estimates_;
    size_t num_poses_ = 0;
    size_t num_loop_closures_ = 0;
    bool optimized_ = false;
    double final_error_ = 0.0;
    int iterations_ = 0;
};

namespace b200 {

// Feeds a batch of factors (sb_odometry_factors / sb_loop_factors) to the back end in order: the calls
// slam_node.cpp:145 (addOdometryFactor(i-1, i, delta, final_error)) and :165 (addLoopClosure(match, query, T)) make.
// Any class with those two members works (the reference's PoseGraph, a test double).
template <class Graph>
inline void hand_off(Graph& graph, const sb_pose_factor* factors, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        const sb_pose_factor& f = factors[i];
        const Transformation rel = Transformation::from_row_major(f.relative);
        if (f.kind == SB_FACTOR_LOOP) graph.addLoopClosure((size_t)f.from, (size_t)f.to, rel);
        else graph.addOdometryFactor((size_t)f.from, (size_t)f.to, rel, f.fitness);
    }
}

}  // namespace b200
}  // namespace slam
