// kdtree.hpp — mirror of slam_viz/include/slam_viz/core/kdtree.hpp:18-221.
// KDTree owns a device-resident spatial index (csrc/forest.cu) instead of a host KD-tree; the queries return the
// same indices as the reference's exact searches (ties: smallest (d2, index)).
#pragma once
#include <limits>
#include <memory>
#include <vector>

#include "backend.hpp"
#include "types.hpp"

namespace slam {

class KDTree {
public:
    explicit KDTree(const PointCloud::Matrix& points) {  // kdtree.hpp:20-26
        sb_index* ix = nullptr;
        b200::check(sb_index_build(b200::context(), points.data(), (int64_t)points.rows(), &ix), "KDTree build");
        index_ = std::shared_ptr<sb_index>(ix, [](sb_index* p) { sb_index_free(p); });
    }

    int nearest(const Eigen::Vector3d& query) const {  // kdtree.hpp:32-37
        double q[3] = {query(0), query(1), query(2)};
        int32_t idx = -1;
        b200::check(sb_index_nearest_batch(index_.get(), q, 1, &idx, nullptr), "KDTree::nearest");
        return idx;
    }

    void nearest_batch(const PointCloud::Matrix& queries, std::vector<int>& indices,
                       std::vector<double>& distances_sq) const {  // kdtree.hpp:43-59
        indices.resize((size_t)queries.rows());
        distances_sq.resize((size_t)queries.rows());
        static_assert(sizeof(int) == sizeof(int32_t), "int is 32 bits");
        b200::check(sb_index_nearest_batch(index_.get(), queries.data(), (int64_t)queries.rows(),
                                           reinterpret_cast<int32_t*>(indices.data()), distances_sq.data()),
                    "KDTree::nearest_batch");
    }

    std::vector<int> k_nearest(const Eigen::Vector3d& query, int k) const {  // kdtree.hpp:65-78
        double q[3] = {query(0), query(1), query(2)};
        std::vector<int32_t> idx((size_t)(k > 0 ? k : 0));
        if (k <= 0) return {};
        b200::check(sb_index_knn(index_.get(), q, 1, k, idx.data(), nullptr), "KDTree::k_nearest");
        std::vector<int> out;
        for (int32_t v : idx)
            if (v >= 0) out.push_back(v);  // fewer than k points: the reference returns them all
        return out;
    }

    // batch form (not in the reference): row q holds min(k, size) indices, padded with -1
    void k_nearest_batch(const PointCloud::Matrix& queries, int k, std::vector<int>& indices) const {
        indices.assign((size_t)queries.rows() * (size_t)k, -1);
        b200::check(sb_index_knn(index_.get(), queries.data(), (int64_t)queries.rows(), k,
                                 reinterpret_cast<int32_t*>(indices.data()), nullptr),
                    "KDTree::k_nearest_batch");
    }

    int64_t size() const { return sb_index_size(index_.get()); }
    sb_index* handle() const { return index_.get(); }

private:
    std::shared_ptr<sb_index> index_;
};

// kdtree.hpp:193-221
class NearestNeighborSearch {
public:
    explicit NearestNeighborSearch(const PointCloud& target) : tree_(target.points()) {}

    void find_correspondences(const PointCloud::Matrix& source, PointCloud::Matrix& matched_target,
                              Eigen::VectorXd& distances) const {  // kdtree.hpp:198-214
        matched_target.resize(source.rows(), 3);
        distances.resize(source.rows());
        b200::check(sb_index_find_correspondences(tree_.handle(), source.data(), (int64_t)source.rows(),
                                                  matched_target.data(), distances.data()),
                    "find_correspondences");
    }

    const KDTree& tree() const { return tree_; }

private:
    KDTree tree_;
};

}  // namespace slam
