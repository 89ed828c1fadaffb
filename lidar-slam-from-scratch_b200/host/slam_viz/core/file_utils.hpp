// file_utils.hpp — mirror of slam_viz/include/slam_viz/core/file_utils.hpp.
//
// voxel_downsample (file_utils.hpp:41-44, file_utils.cpp:148-196) runs on the GPU (csrc/voxel.cu).  The loaders
// (file_utils.hpp:18, 28, 54, 63) are file I/O and stay with the reference: they are DECLARED here with the
// reference's signatures and resolved at link time by the reference's own, unmodified src/core/file_utils.cpp
// (compiled against these mirror headers).
//
// That file also defines slam::voxel_downsample.  The GPU version therefore lives in the inline namespace
// slam::b200_impl: callers that write slam::voxel_downsample(points, voxel) (slam_node.cpp:70, 122, 237) get it —
// it is the only declaration they see — while the reference's definition becomes a separate, unused symbol instead
// of a redefinition or a silent link-time override.
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "backend.hpp"
#include "types.hpp"

namespace slam {

PointCloud::Matrix load_ply(const std::string& filepath);                                   // file_utils.hpp:18
PointCloud::Matrix load_bin(const std::string& filepath);                                   // file_utils.hpp:28
long long extract_timestamp(const std::string& filename);                                   // file_utils.hpp:54
std::vector<std::pair<long long, std::string>> discover_frames(const std::string& data_dir);  // file_utils.hpp:63

inline namespace b200_impl {

// file_utils.cpp:148-196 on the GPU.  Same voxels, same centroids bit for bit; rows come out in ascending
// (kx, ky, kz) order where the reference's order is the unspecified unordered_map iteration order.
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

}  // namespace b200_impl
}  // namespace slam
