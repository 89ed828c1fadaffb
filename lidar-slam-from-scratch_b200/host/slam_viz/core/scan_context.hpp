// scan_context.hpp — mirror of slam_viz/include/slam_viz/core/scan_context.hpp:24-145.
#pragma once
#include "backend.hpp"
#include "types.hpp"

namespace slam {

class ScanContext {
public:
    static constexpr int NUM_RINGS = 20;       // scan_context.hpp:27
    static constexpr int NUM_SECTORS = 60;     // scan_context.hpp:28
    static constexpr double MAX_RANGE = 80.0;  // scan_context.hpp:29

    ScanContext() : descriptor_(NUM_RINGS, NUM_SECTORS) { descriptor_.setZero(); }
    explicit ScanContext(const PointCloud::Matrix& cloud) : descriptor_(NUM_RINGS, NUM_SECTORS) { compute(cloud); }

    void compute(const PointCloud::Matrix& cloud) {  // scan_context.hpp:44-82
        // Eigen::MatrixXd is column-major, exactly the layout sb_sc_compute writes
        b200::check(sb_sc_compute(b200::context(), cloud.data(), (int64_t)cloud.rows(), descriptor_.data()),
                    "ScanContext::compute");
    }

    double distance(const ScanContext& other) const {  // scan_context.hpp:90-102
        double d = 0.0;
        b200::check(sb_sc_distance(b200::context(), descriptor_.data(), other.descriptor_.data(), &d),
                    "ScanContext::distance");
        return d;
    }

    Eigen::VectorXd ring_key() const {  // scan_context.hpp:107-109
        Eigen::VectorXd r(NUM_RINGS), s(NUM_SECTORS);
        b200::check(sb_sc_keys(b200::context(), descriptor_.data(), r.data(), s.data()), "ScanContext::ring_key");
        return r;
    }
    Eigen::VectorXd sector_key() const {  // scan_context.hpp:114-116
        Eigen::VectorXd r(NUM_RINGS), s(NUM_SECTORS);
        b200::check(sb_sc_keys(b200::context(), descriptor_.data(), r.data(), s.data()), "ScanContext::sector_key");
        return s;
    }

    const Eigen::MatrixXd& descriptor() const { return descriptor_; }

private:
    Eigen::MatrixXd descriptor_;
};

}  // namespace slam
