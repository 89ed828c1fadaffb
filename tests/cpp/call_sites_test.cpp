// call_sites_test.cpp — the include order and the call sites of the reference's ROS node, restated without ROS and
// without Eigen expressions, so that it also compiles in the mirror's NO-Eigen branch (dense.hpp stand-in).  The
// Eigen branch is covered by compiling the reference's own slam_node.cpp (tests/cpp/build.sh, slam_node_dropin).
//   includes    slam_viz/include/slam_viz/ros/slam_node.hpp:16-18, slam_viz/src/ros/slam_node.cpp:2-3
//   call sites  slam_node.cpp:64-66 (prior), :69-70 (first frame), :77-81 (detector), :92 (discover_frames),
//               :121-167 (process_frame), :177-185 (optimise), :237 (global map)
// Compile-only (-c): the loaders and PoseGraph are declared by the mirror and defined by the reference.
#include "slam_viz/core/types.hpp"
#include "slam_viz/core/pose_graph.hpp"
#include "slam_viz/core/loop_closure.hpp"
#include "slam_viz/core/file_utils.hpp"
#include "slam_viz/core/icp.hpp"

#include <string>
#include <utility>
#include <vector>

struct NodeState {  // slam_node.hpp:149-161
    std::vector<std::pair<long long, std::string>> frames_;
    std::vector<slam::Transformation> poses_;
    std::vector<slam::PointCloud::Matrix> downsampled_clouds_;
    slam::PointCloud::Matrix prev_points_;
    slam::PoseGraph pose_graph_;
    slam::LoopClosureDetector loop_detector_;
    int loop_closures_found_ = 0;
    bool has_loop_closure_pending_ = false;
};

void start(NodeState& n, const std::string& data_dir, double voxel_size) {
    n.frames_ = slam::discover_frames(data_dir);                                     // :92
    n.poses_.push_back(slam::Transformation::identity());                            // :64
    n.pose_graph_.addPrior(0, slam::Transformation::identity());                     // :66
    auto first_frame_raw = slam::load_ply(n.frames_[0].second);                      // :69
    n.prev_points_ = slam::voxel_downsample(first_frame_raw, voxel_size);            // :70
    n.downsampled_clouds_.push_back(n.prev_points_);
    slam::LoopClosureConfig lc_config;                                               // :77-81
    lc_config.frame_gap = 50;
    lc_config.sc_distance_threshold = 0.2;
    lc_config.icp_fitness_threshold = 0.3;
    n.loop_detector_ = slam::LoopClosureDetector(lc_config);
    (void)slam::extract_timestamp("000001.ply");
    (void)slam::load_bin;
}

void process_frame(NodeState& n, int frame_idx, double voxel_size, int max_iterations, double tolerance) {
    auto raw = slam::load_ply(n.frames_[frame_idx].second);                          // :121
    auto curr = slam::voxel_downsample(raw, voxel_size);                             // :122
    n.downsampled_clouds_.push_back(curr);
    if (curr.rows() < 1000) {                                                        // :125-130
        n.poses_.push_back(n.poses_.back());
        n.prev_points_ = curr;
        return;
    }
    slam::PointCloud source(curr);                                                   // :132-138
    slam::PointCloud target(n.prev_points_);
    slam::ICPConfig icp_cfg;
    icp_cfg.max_iterations = max_iterations;
    icp_cfg.tolerance = tolerance;
    auto result = slam::icp_point_to_plane(source, target, icp_cfg);
    auto delta = (!result.converged || result.final_error > 1.0) ? slam::Transformation::identity()
                                                                 : result.transformation;  // :139-140
    auto new_pose = n.poses_.back() * delta;                                         // :142
    n.poses_.push_back(new_pose);
    (void)new_pose.t();
    (void)new_pose.R();
    n.pose_graph_.addOdometryFactor(n.poses_.size() - 2, n.poses_.size() - 1, delta, result.final_error);  // :145
    (void)source.row(0);                                                             // types.hpp:36-37
    n.prev_points_ = curr;
    n.loop_detector_.addFrame(curr, frame_idx);                                      // :159
    if (frame_idx % 10 == 0 && frame_idx > 50) {                                     // :160-167
        for (const auto& lc : n.loop_detector_.detect()) {
            n.pose_graph_.addLoopClosure(lc.match_frame, lc.query_frame, lc.transform);
            n.loop_closures_found_++;
            n.has_loop_closure_pending_ = true;
        }
    }
}

void finish(NodeState& n, double voxel_size) {
    if (n.pose_graph_.optimize()) n.poses_ = n.pose_graph_.getAllPoses();            // :177-185
    slam::PointCloud::Matrix all = n.downsampled_clouds_.front();
    auto ds = slam::voxel_downsample(all, voxel_size * 2);                           // :237
    (void)ds;
}
