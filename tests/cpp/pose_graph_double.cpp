// pose_graph_double.cpp — TEST DOUBLE of slam::PoseGraph (slam_viz/src/core/pose_graph.cpp), without GTSAM.
//
// Implements the members the mirror's pose_graph.hpp declares, recording every factor with the noise sigmas the
// reference computes (pose_graph.cpp:58-137: prior sigmas; odometry sigmas scaled by 1 + 10 * fitness, :88; loop
// sigmas) and propagating initial estimates like pose_graph.cpp:107-113.  optimize() "succeeds" without moving
// anything (the solver is the part that stays with GTSAM).  On destruction the record is written as JSON to
// $SLAM_DOUBLE_OUT so that a test can compare what the node handed to its back end.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>

#include "slam_viz/core/pose_graph.hpp"

namespace gtsam {
struct RecordedFactor {
    int kind;  // 0 prior, 1 odometry, 2 loop
    size_t from, to;
    double rel[16];
    double sigmas[6];
};
class NonlinearFactorGraph {
public:
    std::vector<RecordedFactor> factors;
};
class Values {
public:
    std::map<size_t, slam::Transformation> poses;
};
class Pose3 {};
}  // namespace gtsam

namespace slam {

static void fill(gtsam::RecordedFactor& f, int kind, size_t from, size_t to, const Transformation& rel, double rs, double ts) {
    f.kind = kind; f.from = from; f.to = to;
    rel.to_row_major(f.rel);
    for (int i = 0; i < 3; ++i) { f.sigmas[i] = rs; f.sigmas[3 + i] = ts; }
}

PoseGraph::PoseGraph(const PoseGraphConfig& config)
    : config_(config), graph_(new gtsam::NonlinearFactorGraph()), initial_estimates_(new gtsam::Values()),
      optimized_estimates_(new gtsam::Values()) {}

PoseGraph::~PoseGraph() {
    const char* path = std::getenv("SLAM_DOUBLE_OUT");
    if (!path || !graph_) return;
    FILE* f = std::fopen(path, "w");
    if (!f) return;
    std::fprintf(f, "{\"num_poses\": %zu, \"num_loop_closures\": %zu, \"factors\": [", num_poses_, num_loop_closures_);
    for (size_t i = 0; i < graph_->factors.size(); ++i) {
        const gtsam::RecordedFactor& r = graph_->factors[i];
        std::fprintf(f, "%s\n{\"kind\": %d, \"from\": %zu, \"to\": %zu, \"relative\": [", i ? "," : "", r.kind, r.from, r.to);
        for (int k = 0; k < 16; ++k) std::fprintf(f, "%s%.17g", k ? ", " : "", r.rel[k]);
        std::fprintf(f, "], \"sigmas\": [");
        for (int k = 0; k < 6; ++k) std::fprintf(f, "%s%.17g", k ? ", " : "", r.sigmas[k]);
        std::fprintf(f, "]}");
    }
    std::fprintf(f, "],\n\"poses\": [");
    bool first = true;
    for (const auto& kv : initial_estimates_->poses) {
        double m[16];
        kv.second.to_row_major(m);
        std::fprintf(f, "%s\n[", first ? "" : ",");
        for (int k = 0; k < 16; ++k) std::fprintf(f, "%s%.17g", k ? ", " : "", m[k]);
        std::fprintf(f, "]");
        first = false;
    }
    std::fprintf(f, "]}\n");
    std::fclose(f);
}

PoseGraph::PoseGraph(PoseGraph&&) noexcept = default;
PoseGraph& PoseGraph::operator=(PoseGraph&&) noexcept = default;

void PoseGraph::addPrior(size_t index, const Transformation& pose) {  // pose_graph.cpp:58-79
    gtsam::RecordedFactor f;
    fill(f, 0, index, index, pose, config_.prior_rotation_sigma, config_.prior_translation_sigma);
    graph_->factors.push_back(f);
    if (!initial_estimates_->poses.count(index)) {
        initial_estimates_->poses[index] = pose;
        num_poses_ = num_poses_ > index + 1 ? num_poses_ : index + 1;
    }
}

void PoseGraph::addOdometryFactor(size_t from_idx, size_t to_idx, const Transformation& rel, double fitness_score) {
    const double scale = 1.0 + fitness_score * 10.0;  // pose_graph.cpp:88
    gtsam::RecordedFactor f;
    fill(f, 1, from_idx, to_idx, rel, config_.odom_rotation_sigma * scale, config_.odom_translation_sigma * scale);
    graph_->factors.push_back(f);
    if (!initial_estimates_->poses.count(to_idx)) {  // pose_graph.cpp:107-113 (.at() throws on a missing key)
        const Transformation from_pose = initial_estimates_->poses.at(from_idx);
        initial_estimates_->poses[to_idx] = from_pose * rel;
        num_poses_ = num_poses_ > to_idx + 1 ? num_poses_ : to_idx + 1;
    }
    optimized_ = false;
}

void PoseGraph::addLoopClosure(size_t from_idx, size_t to_idx, const Transformation& rel) {  // pose_graph.cpp:118-137
    gtsam::RecordedFactor f;
    fill(f, 2, from_idx, to_idx, rel, config_.loop_rotation_sigma, config_.loop_translation_sigma);
    graph_->factors.push_back(f);
    num_loop_closures_++;
    optimized_ = false;
}

bool PoseGraph::optimize() {  // pose_graph.cpp:147-171 minus the solver
    if (num_poses_ == 0) return false;
    *optimized_estimates_ = *initial_estimates_;
    final_error_ = 0.0;
    iterations_ = 0;
    optimized_ = true;
    return true;
}

Transformation PoseGraph::getPose(size_t index) const {  // pose_graph.cpp:177-185
    const gtsam::Values& v = optimized_ ? *optimized_estimates_ : *initial_estimates_;
    auto it = v.poses.find(index);
    if (it == v.poses.end()) throw std::out_of_range("Pose index " + std::to_string(index) + " not found");
    return it->second;
}

std::vector<Transformation> PoseGraph::getAllPoses() const {  // pose_graph.cpp:187-200
    std::vector<Transformation> out;
    const gtsam::Values& v = optimized_ ? *optimized_estimates_ : *initial_estimates_;
    for (size_t i = 0; i < num_poses_; ++i) {
        auto it = v.poses.find(i);
        if (it != v.poses.end()) out.push_back(it->second);
    }
    return out;
}

}  // namespace slam
