// backend.hpp — glue between the slam:: mirror headers and the C ABI (include/slam_b200.h).
//
// One sb_ctx per host thread (the reference is called from the single ROS executor thread, slam_node.cpp:344).
// Every ABI failure becomes a std::runtime_error carrying sb_last_error(): the reference reports errors through
// exceptions only (file_utils.cpp:23,118) and the ABI itself never throws.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>

#include "slam_b200.h"

namespace slam {
namespace b200 {

inline int& device_ordinal() {
    static int d = 0;  // set before the first call on a thread to use another GPU
    return d;
}

struct ContextHolder {
    sb_ctx* ctx = nullptr;
    ContextHolder() {
        int s = sb_ctx_create(device_ordinal(), nullptr, &ctx);
        if (s != SB_OK)
            throw std::runtime_error("slam_b200: no usable sm_100a device (status " + std::to_string(s) +
                                     "); this build has no CPU implementation");
    }
    ~ContextHolder() { sb_ctx_destroy(ctx); }
    ContextHolder(const ContextHolder&) = delete;
    ContextHolder& operator=(const ContextHolder&) = delete;
};

inline sb_ctx* context() {
    thread_local ContextHolder holder;
    return holder.ctx;
}

inline void check(int status, const char* what) {
    if (status != SB_OK)
        throw std::runtime_error(std::string("slam_b200: ") + what + " failed (status " + std::to_string(status) +
                                 "): " + sb_last_error(context()));
}

}  // namespace b200
}  // namespace slam
