// synth.cu — device build of the synthetic LiDAR raycaster (bench/test INPUT GENERATOR, synth/lidar_synth.h).
// Not part of the reference's API: datasets are unavailable offline, so bench.py creates its scans directly in HBM.
#include "../../synth/lidar_synth.h"
#include "common.cuh"

#include <cmath>

namespace sb {

__global__ void __launch_bounds__(256) k_synth_cast(SynSensor s, const SynBox* __restrict__ boxes,
                                                    const int* __restrict__ box_off, const double* __restrict__ poses,
                                                    uint64_t noise_seed, int rays, double* __restrict__ tmp,
                                                    uint32_t* __restrict__ hit) {
    int scan = blockIdx.y;
    int ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= rays) return;
    SynPose pose;
    pose.x = poses[3 * scan]; pose.y = poses[3 * scan + 1]; pose.yaw = poses[3 * scan + 2];
    double out[3] = {0, 0, 0};
    int b0 = box_off[scan], b1 = box_off[scan + 1];
    int h = syn_cast_ray(s, boxes + b0, b1 - b0, pose, noise_seed + (uint64_t)scan, ray, out);
    i64 o = (i64)scan * rays + ray;
    hit[o] = (uint32_t)h;
    tmp[3 * o] = out[0]; tmp[3 * o + 1] = out[1]; tmp[3 * o + 2] = out[2];
}

__global__ void __launch_bounds__(256) k_synth_compact(const double* __restrict__ tmp, const uint32_t* __restrict__ hit,
                                                       const uint32_t* __restrict__ pos, i64 n,
                                                       double* __restrict__ out) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !hit[i]) return;
    i64 p = pos[i];
    out[3 * p] = tmp[3 * i]; out[3 * p + 1] = tmp[3 * i + 1]; out[3 * p + 2] = tmp[3 * i + 2];
}

__global__ void k_synth_offsets(const uint32_t* __restrict__ pos, const uint32_t* __restrict__ total, int rays,
                                int n_scans, i64* __restrict__ off) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_scans) return;
    off[s] = s < n_scans ? (i64)pos[(i64)s * rays] : (i64)*total;
}

int synth_scans_dev(Ctx* ctx, int beams, int azimuth_steps, float elev_top_deg, float elev_bot_deg, float max_range,
                    float noise_sigma, float sensor_height, const float* boxes6, int n_boxes, const double* poses,
                    int n_scans, uint64_t noise_seed, double* d_xyz, i64* out_offsets) {
    SynSensor s{beams, azimuth_steps, elev_top_deg, elev_bot_deg, max_range, noise_sigma, sensor_height};
    const SynBox* boxes = reinterpret_cast<const SynBox*>(boxes6);
    int rays = beams * azimuth_steps;
    i64 n = (i64)n_scans * rays;
    if (n_scans <= 0 || rays <= 0) {
        if (n_scans >= 0) for (int i = 0; i <= n_scans; ++i) out_offsets[i] = 0;
        return SB_OK;
    }
    if (n >= (i64)0xffffffffLL) return fail(ctx, SB_ERR_RANGE, "synth: too many rays in one call");
    // per-scan culling of unreachable boxes (same rule as synth/synth_host.cpp)
    std::vector<SynBox> near;
    std::vector<int> box_off((size_t)n_scans + 1, 0);
    for (int sc = 0; sc < n_scans; ++sc) {
        double x = poses[3 * sc], y = poses[3 * sc + 1];
        for (int b = 0; b < n_boxes; ++b) {
            double ox = boxes[b].cx - x, oy = boxes[b].cy - y;
            double reach = sqrt((double)boxes[b].hx * boxes[b].hx + (double)boxes[b].hy * boxes[b].hy);
            if (sqrt(ox * ox + oy * oy) - reach < max_range) near.push_back(boxes[b]);
        }
        box_off[sc + 1] = (int)near.size();
    }
    SynBox* d_boxes;
    int* d_box_off;
    double *d_poses, *d_tmp;
    uint32_t *d_hit, *d_pos, *d_total;
    i64* d_off;
    SB_TRY(arena_get(ctx, near.size() ? near.size() : 1, &d_boxes));
    SB_TRY(arena_get(ctx, box_off.size(), &d_box_off));
    SB_TRY(arena_get(ctx, (size_t)3 * n_scans, &d_poses));
    SB_TRY(arena_get(ctx, (size_t)3 * n, &d_tmp));
    SB_TRY(arena_get(ctx, (size_t)n, &d_hit));
    SB_TRY(arena_get(ctx, (size_t)n, &d_pos));
    SB_TRY(arena_get(ctx, 1, &d_total));
    SB_TRY(arena_get(ctx, (size_t)n_scans + 1, &d_off));
    if (!near.empty())
        SB_CUDA(ctx, cudaMemcpyAsync(d_boxes, near.data(), sizeof(SynBox) * near.size(), cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(d_box_off, box_off.data(), sizeof(int) * box_off.size(), cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(d_poses, poses, sizeof(double) * 3 * n_scans, cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((unsigned)ceil_div(rays, 256), (unsigned)n_scans, 1);
    SB_LAUNCH(ctx, k_synth_cast, grid, 256, 0, s, d_boxes, d_box_off, d_poses, noise_seed, rays, d_tmp, d_hit);
    SB_TRY(exclusive_scan_u32(ctx, d_hit, d_pos, n, d_total));
    SB_LAUNCH(ctx, k_synth_compact, (unsigned)ceil_div(n, 256), 256, 0, d_tmp, d_hit, d_pos, n, d_xyz);
    SB_LAUNCH(ctx, k_synth_offsets, (unsigned)ceil_div(n_scans + 1, 256), 256, 0, d_pos, d_total, rays, n_scans, d_off);
    SB_CUDA(ctx, cudaMemcpyAsync(out_offsets, d_off, sizeof(i64) * (n_scans + 1), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

}  // namespace sb
