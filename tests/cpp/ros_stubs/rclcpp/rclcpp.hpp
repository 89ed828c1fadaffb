// rclcpp.hpp — TEST STUB of the sliver of ROS 2 (rclcpp) that slam_viz/src/ros/slam_node.cpp uses.  Not ROS, not part
// of the product: it lets tests/cpp/build.sh compile the reference's UNMODIFIED slam_node.cpp against the mirror
// headers, and run its process_frame loop without a ROS installation.
//   * parameters come from the environment: SLAM_PARAM_<name> overrides the declared default;
//   * publishers count messages; the timer callback is driven by rclcpp::spin for SLAM_STUB_TICKS ticks.
#pragma once
#include <chrono>
#include <climits>   // the real rclcpp headers bring it in; slam_node.cpp:284 uses INT_MAX
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace rclcpp {

struct Time {
    long long ns = 0;
};
struct Logger {
    std::string name;
};
inline Logger get_logger(const std::string& name) { return Logger{name}; }

class Parameter {
public:
    explicit Parameter(std::string v = "") : v_(std::move(v)) {}
    std::string as_string() const { return v_; }
    double as_double() const { return std::atof(v_.c_str()); }
    long as_int() const { return std::atol(v_.c_str()); }

private:
    std::string v_;
};

class TimerBase {
public:
    using SharedPtr = std::shared_ptr<TimerBase>;
    std::function<void()> callback;
};

template <class Msg>
class Publisher {
public:
    using SharedPtr = std::shared_ptr<Publisher<Msg>>;
    explicit Publisher(std::string t) : topic(std::move(t)) {}
    void publish(const Msg&) { ++count; }
    std::string topic;
    size_t count = 0;
};

class Node {
public:
    explicit Node(const std::string& name) : name_(name) {}
    virtual ~Node() = default;
    template <class T>
    void declare_parameter(const std::string& name, const T& def) {
        const char* env = std::getenv(("SLAM_PARAM_" + name).c_str());
        if (env) params_[name] = env;
        else params_[name] = to_text(def);
    }
    Parameter get_parameter(const std::string& name) const {
        auto it = params_.find(name);
        return Parameter(it == params_.end() ? std::string() : it->second);
    }
    Logger get_logger() const { return Logger{name_}; }
    Time now() const { return Time{++clock_}; }
    template <class Msg>
    typename Publisher<Msg>::SharedPtr create_publisher(const std::string& topic, int /*qos*/) {
        return std::make_shared<Publisher<Msg>>(topic);
    }
    template <class Duration, class Callback>
    TimerBase::SharedPtr create_wall_timer(Duration, Callback cb) {
        auto t = std::make_shared<TimerBase>();
        t->callback = cb;
        timers_.push_back(t);
        return t;
    }
    const std::vector<TimerBase::SharedPtr>& timers() const { return timers_; }

private:
    static std::string to_text(const std::string& s) { return s; }
    static std::string to_text(const char* s) { return s; }
    static std::string to_text(double d) {
        char b[64];
        std::snprintf(b, sizeof(b), "%.17g", d);
        return b;
    }
    static std::string to_text(int i) { return std::to_string(i); }
    std::string name_;
    std::map<std::string, std::string> params_;
    std::vector<TimerBase::SharedPtr> timers_;
    mutable long long clock_ = 0;
};

inline void init(int, char**) {}
inline void shutdown() {}
// drives every timer of the node SLAM_STUB_TICKS times (default 3), in order
inline void spin(const std::shared_ptr<Node>& node) {
    const char* env = std::getenv("SLAM_STUB_TICKS");
    const long ticks = env ? std::atol(env) : 3;
    for (long i = 0; i < ticks; ++i)
        for (const auto& t : node->timers()) t->callback();
}

}  // namespace rclcpp

#define RCLCPP_STUB_LOG(level, logger, ...)                     \
    do {                                                        \
        if (std::getenv("SLAM_STUB_VERBOSE")) {                 \
            std::fprintf(stderr, "[%s] [%s] ", level, (logger).name.c_str()); \
            std::fprintf(stderr, __VA_ARGS__);                  \
            std::fprintf(stderr, "\n");                         \
        }                                                       \
    } while (0)
#define RCLCPP_INFO(logger, ...) RCLCPP_STUB_LOG("INFO", logger, __VA_ARGS__)
#define RCLCPP_ERROR(logger, ...) RCLCPP_STUB_LOG("ERROR", logger, __VA_ARGS__)
