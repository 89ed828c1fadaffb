// synth_host.cpp — host build of the synthetic LiDAR raycaster (TEST/BENCH INPUT GENERATOR, see lidar_synth.h).
// Built by synth/build.sh into synth/libsynth.so; loaded with ctypes by tests/ and bench.py.
#include "lidar_synth.h"

#include <thread>
#include <vector>

extern "C" {

int syn_scene(unsigned long long seed, int n_boxes, float half_extent, int path_kind, float radius,
              float corridor_half, float sensor_height, float* boxes6) {
    return syn_make_scene(seed, n_boxes, half_extent, path_kind, radius, corridor_half, sensor_height,
                          reinterpret_cast<SynBox*>(boxes6));
}

void syn_pose(int path_kind, double radius, double arc_len, double lateral, double dyaw, double* xyyaw) {
    SynPose p = syn_pose_on_path(path_kind, radius, arc_len, lateral, dyaw);
    xyyaw[0] = p.x; xyyaw[1] = p.y; xyyaw[2] = p.yaw;
}

// One scan; returns the number of points written (<= beams*azimuth_steps), ray order preserved.
long long syn_scan(int beams, int azimuth_steps, float elev_top_deg, float elev_bot_deg, float max_range,
                   float noise_sigma, float sensor_height, const float* boxes6, int n_boxes, double x, double y,
                   double yaw, unsigned long long noise_seed, double* out_xyz, int n_threads) {
    SynSensor s{beams, azimuth_steps, elev_top_deg, elev_bot_deg, max_range, noise_sigma, sensor_height};
    const SynBox* boxes = reinterpret_cast<const SynBox*>(boxes6);
    SynPose pose{x, y, yaw};
    // cull boxes that cannot be reached
    std::vector<SynBox> near;
    for (int b = 0; b < n_boxes; ++b) {
        double ox = boxes[b].cx - x, oy = boxes[b].cy - y;
        double reach = sqrt((double)boxes[b].hx * boxes[b].hx + (double)boxes[b].hy * boxes[b].hy);
        if (sqrt(ox * ox + oy * oy) - reach < max_range) near.push_back(boxes[b]);
    }
    int rays = beams * azimuth_steps;
    std::vector<double> tmp(3 * (size_t)rays);
    std::vector<unsigned char> hit(rays);
    if (n_threads < 1) n_threads = 1;
    auto work = [&](int t) {
        for (int r = t; r < rays; r += n_threads)
            hit[r] = (unsigned char)syn_cast_ray(s, near.data(), (int)near.size(), pose, noise_seed, r, &tmp[3 * (size_t)r]);
    };
    if (n_threads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    long long m = 0;
    for (int r = 0; r < rays; ++r)
        if (hit[r]) { out_xyz[3 * m] = tmp[3 * (size_t)r]; out_xyz[3 * m + 1] = tmp[3 * (size_t)r + 1]; out_xyz[3 * m + 2] = tmp[3 * (size_t)r + 2]; ++m; }
    return m;
}

}  // extern "C"
