// lidar_synth.h — deterministic synthetic spinning-LiDAR raycaster (TEST/BENCH INPUT GENERATOR).
//
// Datasets (KITTI) are not available offline, so every test and bench input is a seeded raycast of a
// "box city": a ground plane at z = -sensor_height plus axis-aligned boxes standing on it, with a free
// corridor along the ego path.  The same code is compiled for the host
// (g++, synth/synth_host.cpp -> libsynth.so, used by the CPU tests) and for the device (nvcc, inside
// libslam_b200.so, used by bench.py to create inputs directly in HBM).  Points are emitted in the SENSOR
// frame, rounded to float32 and widened to fp64 exactly as the reference's load_ply does
// (slam_viz/src/core/file_utils.cpp:91-97).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SYN_HD __host__ __device__ inline
#else
#define SYN_HD inline
#endif

struct SynBox {      // axis-aligned box standing on the ground
    float cx, cy;    // footprint centre (world)
    float hx, hy;    // half extents
    float z0, z1;    // bottom / top (world z, sensor at z = 0)
};

struct SynSensor {
    int beams;           // 64 (C1/C2/C4/C5) or 128 (C3)
    int azimuth_steps;   // 1875 or 2048
    float elev_top_deg;  // +2
    float elev_bot_deg;  // -24.8
    float max_range;     // 120 m
    float noise_sigma;   // 0.02 m
    float sensor_height; // 1.73 m (ground plane z = -1.73 in the sensor frame)
};

struct SynPose { double x, y, yaw; };

SYN_HD uint64_t syn_mix(uint64_t z) {  // splitmix64 finaliser
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
SYN_HD double syn_u01(uint64_t h) { return ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

// Casts ray `ray` (= beam * azimuth_steps + az) of one scan.  Returns 1 and writes the float32-rounded
// sensor-frame point if something within max_range is hit, else 0.
SYN_HD int syn_cast_ray(const SynSensor& s, const SynBox* boxes, int n_boxes, SynPose pose, uint64_t noise_seed,
                        int ray, double out[3]) {
    int beam = ray / s.azimuth_steps, az = ray % s.azimuth_steps;
    double el_deg = s.beams > 1 ? s.elev_top_deg + (s.elev_bot_deg - s.elev_top_deg) * beam / (double)(s.beams - 1)
                                : s.elev_top_deg;
    double el = el_deg * (3.14159265358979323846 / 180.0);
    double a_s = 2.0 * 3.14159265358979323846 * az / (double)s.azimuth_steps;  // sensor-frame azimuth
    double a_w = a_s + pose.yaw;
    double ce = cos(el), se = sin(el);
    double dx = ce * cos(a_w), dy = ce * sin(a_w), dz = se;
    double best = (double)s.max_range;
    int hit = 0;
    if (dz < 0.0) {
        double t = -(double)s.sensor_height / dz;
        if (t < best) { best = t; hit = 1; }
    }
    double idx = 1.0 / dx, idy = 1.0 / dy, idz = 1.0 / dz;
    for (int b = 0; b < n_boxes; ++b) {
        SynBox B = boxes[b];
        double ox = (double)B.cx - pose.x, oy = (double)B.cy - pose.y;
        double tx0 = (ox - B.hx) * idx, tx1 = (ox + B.hx) * idx;
        double ty0 = (oy - B.hy) * idy, ty1 = (oy + B.hy) * idy;
        double tz0 = ((double)B.z0) * idz, tz1 = ((double)B.z1) * idz;
        double tmin = fmax(fmax(fmin(tx0, tx1), fmin(ty0, ty1)), fmin(tz0, tz1));
        double tmax = fmin(fmin(fmax(tx0, tx1), fmax(ty0, ty1)), fmax(tz0, tz1));
        if (tmax >= tmin && tmin > 0.5 && tmin < best) { best = tmin; hit = 1; }
    }
    if (!hit) return 0;
    uint64_t h1 = syn_mix(noise_seed * 0x100000001b3ULL + (uint64_t)ray);
    uint64_t h2 = syn_mix(h1);
    double g = sqrt(-2.0 * log(syn_u01(h1))) * cos(2.0 * 3.14159265358979323846 * syn_u01(h2));
    double r = best + (double)s.noise_sigma * g;
    double csx = ce * cos(a_s), csy = ce * sin(a_s);
    out[0] = (double)(float)(r * csx);
    out[1] = (double)(float)(r * csy);
    out[2] = (double)(float)(r * se);
    return 1;
}

// ---- scene / trajectory generation (host only, tiny) -------------------------------------------------
// path_kind 0: straight along +x through the origin, corridor |y| < corridor_half.
// path_kind 1: circle of radius `radius` centred at the origin, driven counter-clockwise starting at
//              (radius, 0); corridor | |p| - radius | < corridor_half.
#ifndef __CUDA_ARCH__
inline int syn_make_scene(uint64_t seed, int n_boxes, float half_extent, int path_kind, float radius,
                          float corridor_half, float sensor_height, SynBox* out) {
    int made = 0;
    uint64_t ctr = syn_mix(seed ^ 0x5ce7e5eedULL);
    int guard = 0;
    while (made < n_boxes && guard < 100 * n_boxes + 1000) {
        ++guard;
        double u[5];
        for (int i = 0; i < 5; ++i) { ctr = syn_mix(ctr); u[i] = syn_u01(ctr); }
        SynBox b;
        b.cx = (float)((2.0 * u[0] - 1.0) * half_extent);
        b.cy = (float)((2.0 * u[1] - 1.0) * half_extent);
        b.hx = (float)(0.5 * (1.5 + 10.5 * u[2]));
        b.hy = (float)(0.5 * (1.5 + 10.5 * u[3]));
        b.z0 = -sensor_height;
        b.z1 = -sensor_height + (float)(1.0 + 9.0 * u[4]);
        double reach = sqrt((double)b.hx * b.hx + (double)b.hy * b.hy);
        double d;
        if (path_kind == 0) d = fabs((double)b.cy);
        else d = fabs(sqrt((double)b.cx * b.cx + (double)b.cy * b.cy) - radius);
        if (d < corridor_half + reach) continue;
        out[made++] = b;
    }
    return made;
}

inline SynPose syn_pose_on_path(int path_kind, double radius, double arc_len, double lateral, double dyaw) {
    SynPose p;
    if (path_kind == 0) { p.x = arc_len; p.y = lateral; p.yaw = dyaw; }
    else {
        double th = arc_len / radius;
        p.x = (radius + lateral) * cos(th);
        p.y = (radius + lateral) * sin(th);
        p.yaw = th + 1.57079632679489661923 + dyaw;
    }
    return p;
}
#endif
