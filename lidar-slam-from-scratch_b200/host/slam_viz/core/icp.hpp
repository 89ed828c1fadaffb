// icp.hpp — mirror of slam_viz/include/slam_viz/core/icp.hpp:23-258.
#pragma once
#include <cstring>

#include "kdtree.hpp"

namespace slam {

// icp.hpp:23-67.  `points` must be the cloud the tree was built from (it is at the reference's only call site,
// icp.hpp:169-171); the device index already holds them.
inline PointCloud::Matrix estimate_normals(const PointCloud::Matrix& points, const KDTree& tree, int k = 20) {
    if ((int64_t)points.rows() != tree.size())
        throw std::runtime_error("estimate_normals: points must be the cloud the tree was built from");
    PointCloud::Matrix normals(points.rows(), 3);
    b200::check(sb_estimate_normals(tree.handle(), k, normals.data(), nullptr), "estimate_normals");
    return normals;
}

// icp.hpp:89-144: one Gauss-Newton step from given correspondences
inline Transformation solve_point_to_plane(const PointCloud::Matrix& source, const PointCloud::Matrix& target,
                                           const PointCloud::Matrix& normals) {
    double T[16];
    b200::check(sb_solve_point_to_plane(b200::context(), source.data(), target.data(), normals.data(),
                                        (int64_t)source.rows(), T),
                "solve_point_to_plane");
    return Transformation::from_row_major(T);
}

// icp.hpp:157-258
inline ICPResult icp_point_to_plane(const PointCloud& source, const PointCloud& target,
                                    const ICPConfig& config = ICPConfig()) {
    sb_icp_config cfg;
    sb_default_icp_config(&cfg);
    cfg.max_iterations = config.max_iterations;
    cfg.tolerance = config.tolerance;
    cfg.min_error = config.min_error;
    config.initial_transform.to_row_major(cfg.initial_transform);
    sb_icp_result r;
    b200::check(sb_icp_point_to_plane(b200::context(), source.points().data(), (int64_t)source.size(),
                                      target.points().data(), (int64_t)target.size(), &cfg, &r),
                "icp_point_to_plane");
    ICPResult out;
    out.transformation = Transformation::from_row_major(r.transformation);
    out.converged = r.converged != 0;
    out.num_iterations = r.num_iterations;
    out.error_history.assign(r.error_history, r.error_history + r.history_len);
    out.final_error = r.final_error;
    return out;
}

}  // namespace slam
