// loop_closure.hpp — mirror of slam_viz/include/slam_viz/core/loop_closure.hpp:14-149.
// The descriptor database and the cloud copies live in device memory (csrc/loop.cu).
#pragma once
#include <memory>
#include <vector>

#include "icp.hpp"
#include "scan_context.hpp"

namespace slam {

struct LoopClosureConfig {  // loop_closure.hpp:14-19
    int frame_gap = 50;
    double sc_distance_threshold = 0.25;
    double icp_fitness_threshold = 0.3;
    int max_candidates = 3;
};

struct LoopClosureResult {  // loop_closure.hpp:25-31
    int query_frame;
    int match_frame;
    Transformation transform;
    double scan_context_distance;
    double icp_fitness;
};

class LoopClosureDetector {
public:
    explicit LoopClosureDetector(const LoopClosureConfig& config = LoopClosureConfig()) : config_(config) {
        sb_loop_config c;
        sb_default_loop_config(&c);  // loop ICP: 30 iterations, 1e-6 (loop_closure.hpp:105-107), normals k = 20
        c.frame_gap = config.frame_gap;
        c.sc_distance_threshold = config.sc_distance_threshold;
        c.icp_fitness_threshold = config.icp_fitness_threshold;
        c.max_candidates = config.max_candidates;
        sb_loop* L = nullptr;
        b200::check(sb_loop_create(b200::context(), &c, 0, 1, &L), "LoopClosureDetector");
        loop_ = std::shared_ptr<sb_loop>(L, [](sb_loop* p) { sb_loop_free(p); });
    }

    void addFrame(const PointCloud::Matrix& cloud, int frame_idx) {  // loop_closure.hpp:53-59
        b200::check(sb_loop_add_frame(loop_.get(), cloud.data(), (int64_t)cloud.rows(), frame_idx), "addFrame");
    }

    std::vector<LoopClosureResult> detect() {  // loop_closure.hpp:66-126
        int cap = config_.max_candidates > 0 ? config_.max_candidates : 0;
        std::vector<sb_loop_result> raw((size_t)(cap > 0 ? cap : 1));
        int32_t count = 0;
        b200::check(sb_loop_detect(loop_.get(), raw.data(), cap, &count), "detect");
        std::vector<LoopClosureResult> out;
        for (int i = 0; i < count && i < cap; ++i) {
            LoopClosureResult r;
            r.query_frame = raw[i].query_frame;
            r.match_frame = raw[i].match_frame;
            r.transform = Transformation::from_row_major(raw[i].transform);
            r.scan_context_distance = raw[i].scan_context_distance;
            r.icp_fitness = raw[i].icp_fitness;
            out.push_back(r);
        }
        return out;
    }

    size_t size() const { return (size_t)sb_loop_size(loop_.get()); }                // loop_closure.hpp:131
    void clear() { b200::check(sb_loop_clear(loop_.get()), "clear"); }               // loop_closure.hpp:136-141
    // extension: size the device pools for a known sequence (frames, total downsampled rows) before streaming
    void reserve(size_t frames, size_t total_rows) {
        b200::check(sb_loop_reserve(loop_.get(), (int64_t)frames, (int64_t)total_rows), "reserve");
    }

private:
    LoopClosureConfig config_;
    std::shared_ptr<sb_loop> loop_;
};

}  // namespace slam
