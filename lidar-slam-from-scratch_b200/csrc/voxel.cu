// voxel.cu — batched voxel-grid downsampling; replaces slam::voxel_downsample
// (slam_viz/src/core/file_utils.cpp:148-196; declared slam_viz/include/slam_viz/core/file_utils.hpp:41-44).
//
// Reference: unordered_map<VoxelKey, vector<int>> keyed on (long long)floor(coord / voxel), centroid = sum of the
// member points in ascending input index / count.  Here: per-point int64 keys with true IEEE division
// (__ddiv_rn; x * (1/voxel) gives different keys), packed into one 64-bit word relative to the
// batch-wide key minimum, stable segmented radix sort (scan_sort.cu), run-start flags + exclusive scan for the
// compaction, and one thread per voxel that adds its members in sorted (= ascending input) order with __dadd_rn.
// Output rows are ascending in (kx, ky, kz) inside each cloud.
#include "common.cuh"

namespace sb {

struct VoxelPack {
    i64 minx, miny, minz;
    int sx, sy;  // shifts of the x and y fields; z field at bit 0
    u64 mask_y, mask_z;
};

__device__ __forceinline__ bool voxel_key(double c, double voxel, i64* k) {
    double q = floor(__ddiv_rn(c, voxel));
    if (!(fabs(q) < 4.0e18)) {  // also catches NaN / inf
        *k = 0;
        return false;
    }
    *k = (i64)q;
    return true;
}

__global__ void __launch_bounds__(256) k_voxel_minmax(const double* __restrict__ xyz, i64 n, double voxel,
                                                      i64* __restrict__ mm, int* __restrict__ flags) {
    i64 lo[3] = {INT64_MAX, INT64_MAX, INT64_MAX}, hi[3] = {INT64_MIN, INT64_MIN, INT64_MIN};
    bool bad = false;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            i64 k;
            if (!voxel_key(xyz[3 * i + a], voxel, &k)) bad = true;
            lo[a] = min(lo[a], k);
            hi[a] = max(hi[a], k);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    bad = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (lo[a] <= hi[a]) {
                atomicMin(&mm[a], lo[a]);
                atomicMax(&mm[3 + a], hi[a]);
            }
        }
        if (bad) atomicOr(flags, FLAG_NONFINITE);
    }
}

__global__ void __launch_bounds__(256) k_voxel_pack(const double* __restrict__ xyz, i64 n, double voxel, VoxelPack P,
                                                    u64* __restrict__ keys, uint32_t* __restrict__ vals) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    i64 kx, ky, kz;
    voxel_key(xyz[3 * i + 0], voxel, &kx);
    voxel_key(xyz[3 * i + 1], voxel, &ky);
    voxel_key(xyz[3 * i + 2], voxel, &kz);
    u64 ux = (u64)(kx - P.minx), uy = (u64)(ky - P.miny), uz = (u64)(kz - P.minz);
    keys[i] = (ux << P.sx) | (uy << P.sy) | uz;
    vals[i] = (uint32_t)i;
}

__global__ void k_mark_starts(const i64* __restrict__ off, int n_seg, uint32_t* __restrict__ flags) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_seg && off[s] < off[s + 1]) flags[off[s]] = 1u;
}

__global__ void __launch_bounds__(256) k_voxel_flag(const u64* __restrict__ keys, i64 n, uint32_t* __restrict__ flags) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i == 0) return;
    if (keys[i] != keys[i - 1]) flags[i] = 1u;
}

__global__ void __launch_bounds__(256) k_voxel_centroid(const double* __restrict__ xyz, const u64* __restrict__ keys,
                                                        const uint32_t* __restrict__ vals,
                                                        const uint32_t* __restrict__ flags,
                                                        const uint32_t* __restrict__ vpos, i64 n, VoxelPack P,
                                                        double* __restrict__ out_xyz, i64* __restrict__ out_keys) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || flags[i] == 0u) return;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    i64 j = i;
    do {  // members in ascending input index (stable sort), file_utils.cpp:186-190
        const double* p = xyz + 3 * (i64)vals[j];
        sx = __dadd_rn(sx, p[0]);
        sy = __dadd_rn(sy, p[1]);
        sz = __dadd_rn(sz, p[2]);
        ++j;
    } while (j < n && flags[j] == 0u);
    double cnt = (double)(j - i);
    i64 v = vpos[i];
    out_xyz[3 * v + 0] = __ddiv_rn(sx, cnt);  // file_utils.cpp:191
    out_xyz[3 * v + 1] = __ddiv_rn(sy, cnt);
    out_xyz[3 * v + 2] = __ddiv_rn(sz, cnt);
    if (out_keys) {
        u64 k = keys[i];
        out_keys[3 * v + 0] = (i64)(P.sx >= 64 ? 0ull : (k >> P.sx)) + P.minx;
        out_keys[3 * v + 1] = (i64)((k >> P.sy) & P.mask_y) + P.miny;
        out_keys[3 * v + 2] = (i64)(k & P.mask_z) + P.minz;
    }
}

__global__ void k_voxel_offsets(const i64* __restrict__ off, int n_seg, const uint32_t* __restrict__ vpos,
                                const uint32_t* __restrict__ total, i64 n, i64* __restrict__ out_off) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_seg) return;
    i64 o = off[s];
    out_off[s] = o < n ? (i64)vpos[o] : (i64)*total;
}

static int bits_for(u64 range) {
    int b = 0;
    while (range) { ++b; range >>= 1; }
    return b;
}

int voxel_downsample_dev(Ctx* ctx, const double* d_xyz, const i64* h_off, int n_clouds, double voxel,
                         double* d_out_xyz, i64* h_out_off, i64* d_out_keys) {
    i64 n = h_off[n_clouds];
    if (h_off[0] != 0) return fail(ctx, SB_ERR_INVALID_ARG, "voxel: offsets must start at 0");
    if (n >= (i64)0xffffffffLL) return fail(ctx, SB_ERR_RANGE, "voxel: more than 2^32-1 rows in one call");
    if (!(voxel > 0)) {  // file_utils.cpp:152: voxel_size <= 0 returns the input
        if (n > 0)
            SB_CUDA(ctx, cudaMemcpyAsync(d_out_xyz, d_xyz, sizeof(double) * 3 * n, cudaMemcpyDeviceToDevice, ctx->stream));
        if (d_out_keys && n > 0) SB_CUDA(ctx, cudaMemsetAsync(d_out_keys, 0, sizeof(i64) * 3 * n, ctx->stream));
        for (int s = 0; s <= n_clouds; ++s) h_out_off[s] = h_off[s];
        SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return SB_OK;
    }
    if (n == 0) {
        for (int s = 0; s <= n_clouds; ++s) h_out_off[s] = 0;
        return SB_OK;
    }
    // ---- pass 1: key range over the whole batch
    i64* d_mm;
    SB_TRY(arena_get(ctx, 6, &d_mm));
    i64 init[6] = {INT64_MAX, INT64_MAX, INT64_MAX, INT64_MIN, INT64_MIN, INT64_MIN};
    SB_CUDA(ctx, cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    int grid = (int)((n + 255) / 256);
    int cap = ctx->sm_count * 8;
    SB_LAUNCH(ctx, k_voxel_minmax, grid < cap ? grid : cap, 256, 0, d_xyz, n, voxel, d_mm, ctx->d_flags);
    i64 mm[6];
    int flags = 0;
    SB_CUDA(ctx, cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(&flags, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (flags & FLAG_NONFINITE) return fail(ctx, SB_ERR_RANGE, "voxel: non-finite coordinate or |coord/voxel| >= 4e18");
    int bx = bits_for((u64)(mm[3] - mm[0])), by = bits_for((u64)(mm[4] - mm[1])), bz = bits_for((u64)(mm[5] - mm[2]));
    if (bx + by + bz > 64) return fail(ctx, SB_ERR_RANGE, "voxel: key range needs %d bits (> 64)", bx + by + bz);
    VoxelPack P;
    P.minx = mm[0]; P.miny = mm[1]; P.minz = mm[2];
    P.sy = bz;
    P.sx = by + bz;
    P.mask_z = bz >= 64 ? ~0ull : ((1ull << bz) - 1ull);
    P.mask_y = by >= 64 ? ~0ull : ((1ull << by) - 1ull);
    VoxelPack Pk = P;  // shifts clamped for the pack kernel (a field with 0 bits always holds 0)
    if (Pk.sx > 63) Pk.sx = 63;
    if (Pk.sy > 63) Pk.sy = 63;
    // ---- pass 2: packed keys, sort inside each cloud
    u64 *ka, *kb, *ks;
    uint32_t *va, *vb, *vs;
    SB_TRY(arena_get(ctx, (size_t)n, &ka));
    SB_TRY(arena_get(ctx, (size_t)n, &kb));
    SB_TRY(arena_get(ctx, (size_t)n, &va));
    SB_TRY(arena_get(ctx, (size_t)n, &vb));
    SB_LAUNCH(ctx, k_voxel_pack, grid, 256, 0, d_xyz, n, voxel, Pk, ka, va);
    SB_TRY(segmented_sort_pairs(ctx, ka, kb, va, vb, h_off, n_clouds, bx + by + bz, &ks, &vs));
    // ---- run starts, compaction positions
    uint32_t *d_flag, *d_vpos, *d_total;
    i64 *d_off, *d_out_off;
    SB_TRY(arena_get(ctx, (size_t)n, &d_flag));
    SB_TRY(arena_get(ctx, (size_t)n, &d_vpos));
    SB_TRY(arena_get(ctx, 1, &d_total));
    SB_TRY(arena_get(ctx, (size_t)n_clouds + 1, &d_off));
    SB_TRY(arena_get(ctx, (size_t)n_clouds + 1, &d_out_off));
    SB_CUDA(ctx, cudaMemcpyAsync(d_off, h_off, sizeof(i64) * (n_clouds + 1), cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(uint32_t) * n, ctx->stream));
    SB_LAUNCH(ctx, k_mark_starts, ceil_div(n_clouds, 256), 256, 0, d_off, n_clouds, d_flag);
    SB_LAUNCH(ctx, k_voxel_flag, grid, 256, 0, ks, n, d_flag);
    SB_TRY(exclusive_scan_u32(ctx, d_flag, d_vpos, n, d_total));
    SB_LAUNCH(ctx, k_voxel_centroid, grid, 256, 0, d_xyz, ks, vs, d_flag, d_vpos, n, P, d_out_xyz, d_out_keys);
    SB_LAUNCH(ctx, k_voxel_offsets, ceil_div(n_clouds + 1, 256), 256, 0, d_off, n_clouds, d_vpos, d_total, n, d_out_off);
    SB_CUDA(ctx, cudaMemcpyAsync(h_out_off, d_out_off, sizeof(i64) * (n_clouds + 1), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

}  // namespace sb
