// file_utils.hpp — mirror of the hot-path part of slam_viz/include/slam_viz/core/file_utils.hpp:41-44.
// (load_ply / load_bin / discover_frames are file I/O and stay with the reference)
#pragma once
#include "backend.hpp"
#include "types.hpp"

namespace slam {

// file_utils.cpp:148-196 on the GPU (csrc/voxel.cu).  Same voxels, same centroids bit for bit; rows come out in
// ascending (kx, ky, kz) order where the reference's order is the unspecified unordered_map iteration order.
inline PointCloud::Matrix voxel_downsample(const PointCloud::Matrix& points, double voxel_size) {
    PointCloud::Matrix out(points.rows(), 3);
    int64_t m = 0;
    b200::check(sb_voxel_downsample(b200::context(), points.data(), (int64_t)points.rows(), voxel_size, out.data(), &m,
                                    nullptr),
                "voxel_downsample");
    PointCloud::Matrix res((long)m, 3);
    for (int64_t i = 0; i < 3 * m; ++i) res.data()[i] = out.data()[i];
    return res;
}

}  // namespace slam
