// voxel.cu — batched voxel-grid downsampling; replaces slam::voxel_downsample
// (slam_viz/src/core/file_utils.cpp:148-196; declared slam_viz/include/slam_viz/core/file_utils.hpp:41-44).
//
// Reference: unordered_map<VoxelKey, vector<int>> keyed on (long long)floor(coord / voxel), centroid = sum of the
// member points in ascending input index / count.  Here: per-point int64 keys with true IEEE division
// (__ddiv_rn; x * (1/voxel) gives different keys), packed into one 64-bit word relative to the
// batch-wide key minimum, stable segmented radix sort (scan_sort.cu), run-start flags + exclusive scan for the
// compaction, and one thread per voxel that adds its members in sorted (= ascending input) order with __dadd_rn.
// Output rows are ascending in (kx, ky, kz) inside each cloud.
//
// That sort-based pipeline is the GENERAL path.  Scans that come from a float32 file (file_utils.cpp:91-97 widens
// float32 to double) take a one-pass FAST path first: every point is hashed into a per-cloud table of voxels and its
// coordinates are added to the voxel's sums as 20.44 fixed-point integers with 64-bit atomics.  Integer addition
// is exact and order-free.  When every member coordinate of a voxel is a multiple of 2^g and sum|v| < 2^(53+g),
// every partial sum of the reference's left-to-right fp64 loop is a multiple of 2^g below 2^(53+g), i.e. exactly
// representable — so that loop never rounds and its result IS the integer sum.  The kernel proves this per voxel
// and axis with g = -33 (every float32 of magnitude >= 2^-10) or, flagged, g = -44.  Voxels that cannot be proven
// exact (a member finer than 2^-44, or too large a sum) are re-summed in input order from a list of their members;
// if there are too many of them, or keys do not fit 21 bits per axis, the call falls back to the sort.
#include "common.cuh"

#include <climits>
#include <cmath>
#include <cstdlib>

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

// floor(c / voxel) without the division where that is provably the same integer.  rinv = fl(1 / voxel).  If voxel is
// a power of two (pow2) c * rinv IS c / voxel.  Otherwise q~ = fl(c * rinv) is within 2^-51 |q~| of the exact quotient
// and fl(c / voxel) within 2^-53: unless q~ lies within 2^-49 |q~| of an integer all three have the same floor; the
// rare rest (and huge quotients) takes the IEEE division the reference performs (file_utils.cpp:177-179).
// the rare exact path, out of line: inlined, its twelve copies (three axes x four rows) made k_vox_insert 64 KB of code
__device__ __noinline__ bool voxel_key_slow(double c, double voxel, i64* k) { return voxel_key(c, voxel, k); }

// -> 0 and the key as a 32-bit integer inside the 21-bit window of the packed table keys, or the flag that makes the
// call fail (non-finite) / fall back to the sort (key outside the window)
__device__ __forceinline__ int voxel_key_fast(double c, double voxel, double rinv, bool pow2, int* k) {
    const double q = c * rinv;
    const double f = floor(q);
    bool exact = !(fabs(q) < 1048575.0);   // outside the window, NaN, inf: the exact path classifies it
    if (!pow2 && !exact) {
        const double r = q - f, tol = fabs(q) * 1.7763568394002505e-15 + 1.0e-300;   // 2^-49
        exact = !(r > tol && (1.0 - r) > tol);
    }
    if (exact) {
        i64 kk;
        if (!voxel_key_slow(c, voxel, &kk)) return FLAG_NONFINITE;
        if (kk < -(i64)(1 << 20) || kk >= (i64)(1 << 20)) return FLAG_KEY_RANGE;
        *k = (int)kk;
        return 0;
    }
    *k = (int)f;   // |f| < 2^20
    return 0;
}

// slot of `key` in a cloud's table (it is there: k_vox_insert put it)
__device__ __forceinline__ unsigned vox_hash(unsigned long long key, unsigned size) {
    return __umulhi((unsigned)((key * 0x9E3779B97F4A7C15ull) >> 32), size);
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

// General path: any key range that packs into 64 bits, any fp64 input.
static int voxel_sorted_dev(Ctx* ctx, const double* d_xyz, const i64* h_off, int n_clouds, double voxel,
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

// =============================================================================================================
// fast path
// =============================================================================================================
static constexpr double VQ_SCALE = 17592186044416.0;   // 2^44: the sums carry 44 fixed-point fraction bits
static constexpr int VQ_COARSE_TZ = 11;                // a coordinate is "coarse" if it is a multiple of 2^(11-44)
static constexpr double VQ_LIMIT_COARSE = 524288.0;    // 2^19: sum|v| below this -> no wrap, fp64 loop exact (coarse)
static constexpr double VQ_LIMIT_FINE = 512.0;         // 2^(53-44): sum|v| bound when some member is finer
static constexpr int VK_BIAS = 1 << 20;                // keys are packed as (k + 2^20), 21 bits per axis
#ifndef SB_VGROUP
#define SB_VGROUP 8
#endif
static constexpr int VGROUP = SB_VGROUP;                       // lanes of one run whose sums are combined before the atomics
static constexpr int VROWS = 4;                        // rows per thread and block iteration
static constexpr int VTILE = 256 * VROWS;              // rows per block iteration
static constexpr int PATCH_CAP = 4096;                 // members of unproven voxels that the patch-up pass can take
#define VOX_EMPTY 0xffffffffffffffffull
enum { FLAG_TABLE_FULL = 4, FLAG_PATCH_OVERFLOW = 8 };
// VoxAcc::flags while inserting: bit a (0..2): a member's coordinate a is finer than 2^-33; bit 3: finer than
// 2^-44 (not representable in the sums).  After k_vox_finalize: 0, or 1 + output row of a voxel that needs the
// ordered sum.
enum { VF_FINE_X = 1, VF_FINE_Y = 2, VF_FINE_Z = 4, VF_UNREPRESENTABLE = 8 };

struct VoxCloud {   // per cloud, device-resident
    i64 pt_off;     // first input row
    i64 tab_off;    // first slot of the cloud's table
    i64 tile_off;   // first tile (VTILE rows) of the cloud
    int n;
    unsigned size;  // slots of the cloud's table (a multiple of 256, not a power of two: no rounding up to 2x)
};

// One voxel of a cloud's hash table = one 8-byte key (packed (kx, ky, kz) or VOX_EMPTY) in the key array + one
// 32-byte accumulator — exactly one sector — in the accumulator array.  Probing walks the keys only: four to a sector,
// 8 bytes per slot instead of 48, so a cloud's keys (~0.2 MB) stay in L1/L2 while the scan streams through.
struct alignas(32) VoxAcc {
    long long sx, sy, sz;     // fixed-point sums of the member coordinates
    unsigned cnt;
    unsigned flags;
};

__device__ __forceinline__ void load_xyz(const PointSrc& S, i64 row, double& x, double& y, double& z) {
    if (S.f32) {  // float32 rows widened exactly like load_ply does (file_utils.cpp:91-97)
        const float* p = static_cast<const float*>(S.base) + row * S.stride;
        x = (double)p[0]; y = (double)p[1]; z = (double)p[2];
    } else {
        const double* p = static_cast<const double*>(S.base) + row * S.stride;
        x = p[0]; y = p[1]; z = p[2];
    }
}

// fixed-point image of a coordinate; bit: this axis' "fine" flag
__device__ __forceinline__ long long vox_fixed(double v, unsigned bit, unsigned* fl) {
    double t = v * VQ_SCALE;  // exact: a power of two
    long long f = (long long)t;
    if (!(t == rint(t) && fabs(t) < 4.0e18)) *fl |= VF_UNREPRESENTABLE;
    else if (f != 0 && (f & ((1ll << VQ_COARSE_TZ) - 1)) != 0) *fl |= bit;
    return f;
}

// hash + accumulate.  mm: [0..2] min keys, [3..5] max keys; n_vox[c]: distinct voxels of cloud c.
__global__ void __launch_bounds__(256) k_vox_insert(const PointSrc src, const VoxCloud* __restrict__ clouds,
                                                    const int* __restrict__ tile_cloud, i64 n_tiles, double voxel,
                                                    double rinv, int pow2, unsigned long long* __restrict__ tkeys,
                                                    VoxAcc* __restrict__ tacc, int* __restrict__ n_vox, i64* __restrict__ mm,
                                                    int* __restrict__ flags, i64 win, i64 perm, i64 perm_last) {
    __shared__ i64 s_red[8][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    int bad = 0;
    for (i64 tile_seq = blockIdx.x; tile_seq < n_tiles; tile_seq += gridDim.x) {
        // scattered inside a window of `win` consecutive tiles (see the host side)
        const i64 w0 = (tile_seq / win) * win;
        const i64 wn = n_tiles - w0 < win ? n_tiles - w0 : win;
        const i64 pw = wn == win ? perm : perm_last;
        const i64 tile = pw ? w0 + (i64)(((unsigned long long)(tile_seq - w0) * (unsigned long long)pw) % (unsigned long long)wn) : tile_seq;
        const int c = tile_cloud[tile];
        const VoxCloud C = clouds[c];
        const i64 t0 = (tile - C.tile_off) * VTILE;
        unsigned long long* tab = tkeys + C.tab_off;
        VoxAcc* acc = tacc + C.tab_off;
        int claimed = 0;  // voxels this thread created in this tile: one counter update per warp and tile
#pragma unroll
        for (int r = 0; r < VROWS; ++r) {
            const i64 i = t0 + r * 256 + threadIdx.x;  // row inside the cloud
            unsigned slot = 0xffffffffu - (unsigned)lane;  // lanes without a voxel never share a run
            long long fx = 0, fy = 0, fz = 0;
            unsigned fl = 0u;
            if (i < C.n) {
                double x, y, z;
                load_xyz(src, C.pt_off + i, x, y, z);
                int kx = 0, ky = 0, kz = 0;
                const int st = voxel_key_fast(x, voxel, rinv, pow2, &kx) | voxel_key_fast(y, voxel, rinv, pow2, &ky) |
                               voxel_key_fast(z, voxel, rinv, pow2, &kz);
                if (st) {
                    bad |= st;
                } else {
                    lo[0] = min(lo[0], kx); hi[0] = max(hi[0], kx);
                    lo[1] = min(lo[1], ky); hi[1] = max(hi[1], ky);
                    lo[2] = min(lo[2], kz); hi[2] = max(hi[2], kz);
                    const unsigned long long key = ((unsigned long long)(kx + VK_BIAS) << 42) |
                                                   ((unsigned long long)(ky + VK_BIAS) << 21) |
                                                   (unsigned long long)(kz + VK_BIAS);
                    unsigned s = vox_hash(key, C.size);
                    unsigned probes = 0;
                    while (true) {
                        unsigned long long cur = __ldcg(&tab[s]);
                        if (cur == VOX_EMPTY) {
#if !defined(SB_CAS_DEDUPE) || SB_CAS_DEDUPE
                            // The rows of a run (neighbouring rays, neighbouring lanes) find their voxel's slot empty
                            // together, and their compare-and-swaps on the one address queue up in L2 (18 % of the
                            // kernel's stall samples for 1 % of its instructions): one lane per key claims, the
                            // others take its answer.  Same key = same probe sequence = same slot.
                            const unsigned peers = __match_any_sync(__activemask(), key);
                            const int leader = __ffs(peers) - 1;
                            if (lane == leader) cur = atomicCAS(&tab[s], VOX_EMPTY, key);
                            cur = __shfl_sync(peers, cur, leader);
                            if (cur == VOX_EMPTY) {
                                if (lane == leader) ++claimed;
                                cur = key;
                            }
#else
                            cur = atomicCAS(&tab[s], VOX_EMPTY, key);
                            if (cur == VOX_EMPTY) {
                                ++claimed;
                                cur = key;
                            }
#endif
                        }
                        if (cur == key) { slot = s; break; }
                        s = s + 1u == C.size ? 0u : s + 1u;
                        if (++probes >= C.size) { bad |= FLAG_TABLE_FULL; break; }
                    }
                    fx = vox_fixed(x, VF_FINE_X, &fl);
                    fy = vox_fixed(y, VF_FINE_Y, &fl);
                    fz = vox_fixed(z, VF_FINE_Z, &fl);
                }
            }
            // runs of consecutive lanes in the same voxel (neighbouring rays) add up inside the warp first: a segmented
            // inclusive scan, integers, exact in any order.  Runs are cut into aligned groups of VGROUP lanes — two
            // shuffle steps instead of five (the full scan was 36 % of this kernel's instructions; most runs are
            // short) — and the last lane of a group sends the group's sums to the table.
            const unsigned prev_slot = __shfl_up_sync(0xffffffffu, slot, 1);
            const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || slot != prev_slot);
            const int run_start = 31 - __clz(heads & (lanemask_lt() | (1u << lane)));
            const int group_start = run_start + ((lane - run_start) & ~(VGROUP - 1));
#pragma unroll
            for (int d = 1; d < VGROUP; d <<= 1) {
                const long long ox = __shfl_up_sync(0xffffffffu, fx, d), oy = __shfl_up_sync(0xffffffffu, fy, d),
                                oz = __shfl_up_sync(0xffffffffu, fz, d);
                const unsigned of = __shfl_up_sync(0xffffffffu, fl, d);
                if (lane - d >= group_start) {
                    fx += ox; fy += oy; fz += oz;
                    fl |= of;
                }
            }
            const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u) || ((lane - run_start) & (VGROUP - 1)) == VGROUP - 1;
            if (tail && slot < 0xffffffe0u) {
                VoxAcc* S = acc + slot;
                atomicAdd(reinterpret_cast<unsigned long long*>(&S->sx), (unsigned long long)fx);
                atomicAdd(reinterpret_cast<unsigned long long*>(&S->sy), (unsigned long long)fy);
                atomicAdd(reinterpret_cast<unsigned long long*>(&S->sz), (unsigned long long)fz);
                atomicAdd(&S->cnt, (unsigned)(lane - group_start + 1));
                if (fl) atomicOr(&S->flags, fl);
            }
        }
        claimed = __reduce_add_sync(0xffffffffu, claimed);
        if (lane == 0 && claimed) atomicAdd(&n_vox[c], claimed);
    }
    // key range + flags of this block
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
        hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
    }
    bad = (int)__reduce_or_sync(0xffffffffu, (unsigned)bad);
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            s_red[warp][a] = lo[a] == INT_MAX ? INT64_MAX : (i64)lo[a];
            s_red[warp][3 + a] = hi[a] == INT_MIN ? INT64_MIN : (i64)hi[a];
        }
        if (bad) atomicOr(flags, bad);
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int a = threadIdx.x;
        i64 v = s_red[0][a];
        for (int w = 1; w < 8; ++w) v = a < 3 ? min(v, s_red[w][a]) : max(v, s_red[w][a]);
        // thousands of blocks end here and all of them would hit the same six words: look first, only a block that
        // improves the bound pays for the atomic
        const i64 cur = *reinterpret_cast<volatile i64*>(&mm[a]);
        if (a < 3) { if (v != INT64_MAX && v < cur) atomicMin(&mm[a], v); }
        else if (v != INT64_MIN && v > cur) atomicMax(&mm[a], v);
    }
}

// empty table: key = VOX_EMPTY, accumulators 0
__global__ void __launch_bounds__(256) k_vox_clear(unsigned long long* __restrict__ tkeys, VoxAcc* __restrict__ tacc,
                                                   i64 n_slots) {
    const i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_slots) return;
    tkeys[g] = VOX_EMPTY;
    uint4* w = reinterpret_cast<uint4*>(tacc + g);
    w[0] = make_uint4(0u, 0u, 0u, 0u);
    w[1] = make_uint4(0u, 0u, 0u, 0u);
}

// occupied slots -> per-cloud compact (relative key, slot) lists.  Tables are multiples of 256 slots, so the 256
// slots of a block belong to one cloud: one cursor atomic per block.
__global__ void __launch_bounds__(256) k_vox_list(const VoxCloud* __restrict__ clouds, int n_clouds, i64 n_slots,
                                                  const unsigned long long* __restrict__ tkeys,
                                                  const i64* __restrict__ out_off,
                                                  int* __restrict__ cursor, VoxelPack P, u64* __restrict__ lst_key,
                                                  uint32_t* __restrict__ lst_slot) {
    __shared__ int s_wcnt[8];
    __shared__ i64 s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const i64 g = (i64)blockIdx.x * 256 + threadIdx.x;
    const unsigned long long key = g < n_slots ? tkeys[g] : VOX_EMPTY;
    const bool occ = key != VOX_EMPTY;
    const unsigned bal = __ballot_sync(0xffffffffu, occ);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
        for (int w = 0; w < 8; ++w) { int c = s_wcnt[w]; s_wcnt[w] = total; total += c; }
        i64 base = 0;
        if (total > 0) {
            const i64 g0 = (i64)blockIdx.x * 256;
            int lo = 0, hi = n_clouds;  // cloud of the block's slots
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (clouds[mid].tab_off <= g0) lo = mid; else hi = mid;
            }
            base = out_off[lo] + atomicAdd(&cursor[lo], total);
        }
        s_base = base;
    }
    __syncthreads();
    if (!occ) return;
    const i64 kx = (i64)(key >> 42) - VK_BIAS, ky = (i64)((key >> 21) & 0x1fffffu) - VK_BIAS,
              kz = (i64)(key & 0x1fffffu) - VK_BIAS;
    const i64 dst = s_base + s_wcnt[warp] + __popc(bal & lanemask_lt());
    lst_key[dst] = ((u64)(kx - P.minx) << P.sx) | ((u64)(ky - P.miny) << P.sy) | (u64)(kz - P.minz);
    lst_slot[dst] = (uint32_t)g;
}

// one thread per output voxel (sorted by key): centroid from the integer sums, or a request for the ordered sum
__global__ void __launch_bounds__(256) k_vox_finalize(const u64* __restrict__ keys, const uint32_t* __restrict__ slots,
                                                      i64 m, double voxel, VoxelPack P, VoxAcc* __restrict__ table,
                                                      double* __restrict__ out_xyz, i64* __restrict__ out_keys,
                                                      int* __restrict__ n_unproven) {
    const i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= m) return;
    const u64 k = keys[v];
    VoxAcc* S = table + slots[v];
    const i64 kk[3] = {(i64)(P.sx >= 64 ? 0ull : (k >> P.sx)) + P.minx, (i64)((k >> P.sy) & P.mask_y) + P.miny,
                       (i64)(k & P.mask_z) + P.minz};
    const unsigned cnt = S->cnt, fl = S->flags;
    const double dc = (double)cnt;
    // |member| <= (|key| + 1) * voxel bounds B = the sum of the absolute values on each axis.  The 64-bit sums did
    // not wrap and — all members being multiples of 2^-33 — the fp64 loop was exact if B < 2^19; if some member is
    // finer (but a multiple of 2^-44) the loop was exact if B < 2^9.
    bool proven = !(fl & VF_UNREPRESENTABLE);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double B = dc * ((double)(kk[a] < 0 ? -kk[a] : kk[a]) + 1.0) * voxel * 1.0001;
        proven = proven && B < ((fl >> a) & 1u ? VQ_LIMIT_FINE : VQ_LIMIT_COARSE);
    }
    const unsigned want = proven ? 0u : (unsigned)(v + 1);
    if (fl != want) S->flags = want;   // (almost never: an unconditional store dirtied every accumulator sector)
    if (!proven) atomicAdd(n_unproven, 1);  // rare; lets the member collection skip its pass over all rows
    out_xyz[3 * v + 0] = __ddiv_rn((double)S->sx / VQ_SCALE, dc);  // file_utils.cpp:191
    out_xyz[3 * v + 1] = __ddiv_rn((double)S->sy / VQ_SCALE, dc);
    out_xyz[3 * v + 2] = __ddiv_rn((double)S->sz / VQ_SCALE, dc);
    if (out_keys) { out_keys[3 * v + 0] = kk[0]; out_keys[3 * v + 1] = kk[1]; out_keys[3 * v + 2] = kk[2]; }
}

// members of the voxels that asked for the ordered sum: (output row << 32 | input row)
__global__ void __launch_bounds__(256) k_vox_collect(const VoxCloud* __restrict__ clouds,
                                                     const int* __restrict__ tile_cloud, i64 n_tiles, const PointSrc src,
                                                     double voxel, const unsigned long long* __restrict__ tkeys,
                                                     const VoxAcc* __restrict__ tacc, u64* __restrict__ list,
                                                     int* __restrict__ list_n, int* __restrict__ flags,
                                                     const int* __restrict__ n_unproven) {
    if (*n_unproven == 0) return;  // every voxel was proven exact (always, for scans born as float32)
    for (i64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const VoxCloud C = clouds[tile_cloud[tile]];
        const i64 t0 = (tile - C.tile_off) * VTILE;
#pragma unroll
        for (int r = 0; r < VROWS; ++r) {
            const i64 i = t0 + r * 256 + threadIdx.x;
            if (i >= C.n) continue;
            // the row's voxel again (this pass only runs when some voxel asked for it: never for float32-born scans)
            double x, y, z;
            load_xyz(src, C.pt_off + i, x, y, z);
            i64 kx, ky, kz;
            if (!(voxel_key(x, voxel, &kx) & voxel_key(y, voxel, &ky) & voxel_key(z, voxel, &kz))) continue;
            const unsigned long long key = ((unsigned long long)(kx + VK_BIAS) << 42) |
                                           ((unsigned long long)(ky + VK_BIAS) << 21) | (unsigned long long)(kz + VK_BIAS);
            const unsigned long long* tab = tkeys + C.tab_off;
            unsigned s = vox_hash(key, C.size);
            for (unsigned probes = 0; probes < C.size && __ldg(&tab[s]) != key; ++probes) s = s + 1u == C.size ? 0u : s + 1u;
            const unsigned row1 = __ldg(&tacc[C.tab_off + s].flags);
            if (row1 == 0u) continue;
            const int at = atomicAdd(list_n, 1);
            if (at < PATCH_CAP) list[at] = ((u64)(row1 - 1u) << 32) | (u64)(C.pt_off + i);
            else atomicOr(flags, FLAG_PATCH_OVERFLOW);
        }
    }
}

// one block: bitonic sort of the member list, then one thread per voxel run adds its members in input order
__global__ void __launch_bounds__(1024) k_vox_patch(const PointSrc src, u64* __restrict__ list,
                                                    const int* __restrict__ list_n, double* __restrict__ out_xyz) {
    __shared__ u64 s[PATCH_CAP];
    int n = *list_n;
    if (n <= 0) return;
    if (n > PATCH_CAP) n = PATCH_CAP;  // overflow: the host discards this result
    int cap = 2;
    while (cap < n) cap <<= 1;
    for (int i = threadIdx.x; i < cap; i += 1024) s[i] = i < n ? list[i] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= cap; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < cap; i += 1024) {
                int l = i ^ j;
                if (l > i) {
                    u64 a = s[i], b = s[l];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { s[i] = b; s[l] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += 1024) {
        const u64 e = s[i];
        if (i > 0 && (s[i - 1] >> 32) == (e >> 32)) continue;  // not the first member of its voxel
        double sx = 0.0, sy = 0.0, sz = 0.0;
        int j = i;
        do {  // members in ascending input row, file_utils.cpp:186-190
            double x, y, z;
            load_xyz(src, (i64)(s[j] & 0xffffffffull), x, y, z);
            sx = __dadd_rn(sx, x); sy = __dadd_rn(sy, y); sz = __dadd_rn(sz, z);
            ++j;
        } while (j < n && (s[j] >> 32) == (e >> 32));
        const double cnt = (double)(j - i);
        const i64 v = (i64)(e >> 32);
        out_xyz[3 * v + 0] = __ddiv_rn(sx, cnt);
        out_xyz[3 * v + 1] = __ddiv_rn(sy, cnt);
        out_xyz[3 * v + 2] = __ddiv_rn(sz, cnt);
    }
}

// returns SB_OK with *done = 1 if the fast path produced the result, *done = 0 if the caller must use the sort
static int voxel_hashed_dev(Ctx* ctx, const PointSrc src, const i64* h_off, int n_clouds, double voxel,
                            double* d_out_xyz, i64* h_out_off, i64* d_out_keys, int* done) {
    *done = 0;
    const i64 n = h_off[n_clouds];
    if (n >= (i64)0xffffffffLL) return SB_OK;
    // ---- per-cloud tables: a power of two of slots, vox_slots_per_point * rows (the ratio adapts, see below)
    std::vector<VoxCloud> hc((size_t)n_clouds);
    i64 n_slots = 0, n_tiles = 0;
    for (int c = 0; c < n_clouds; ++c) {
        VoxCloud& C = hc[c];
        i64 nc = h_off[c + 1] - h_off[c];
        if (nc > 0x7fffffffLL) return SB_OK;
        i64 want = (i64)(ctx->vox_slots_per_point * (double)nc);
        i64 sl = ((want > 256 ? want : 256) + 255) / 256 * 256;
        if (sl > 0x40000000LL) return SB_OK;
        C.pt_off = h_off[c];
        C.tab_off = n_slots;
        C.tile_off = n_tiles;
        C.n = (int)nc;
        C.size = (unsigned)sl;
        n_slots += sl;
        n_tiles += (nc + VTILE - 1) / VTILE;
    }
    if (n_slots >= (i64)0xffffffffLL) return SB_OK;
    std::vector<int> h_tile_cloud((size_t)n_tiles);
    for (int c = 0; c < n_clouds; ++c) {
        i64 t1 = c + 1 < n_clouds ? hc[c + 1].tile_off : n_tiles;
        for (i64 t = hc[c].tile_off; t < t1; ++t) h_tile_cloud[(size_t)t] = c;
    }
    VoxCloud* d_clouds;
    int* d_tile_cloud;
    unsigned long long* d_tkeys;
    VoxAcc* d_table;
    int *d_nvox, *d_cursor, *d_list_n;
    i64* d_mm;
    u64* d_list;
    SB_TRY(arena_get(ctx, (size_t)n_clouds, &d_clouds));
    SB_TRY(arena_get(ctx, (size_t)(n_tiles > 0 ? n_tiles : 1), &d_tile_cloud));
    SB_TRY(arena_get(ctx, (size_t)n_slots, &d_tkeys));
    SB_TRY(arena_get(ctx, (size_t)n_slots, &d_table));
    SB_TRY(arena_get(ctx, (size_t)2 * n_clouds + 2, &d_nvox));
    SB_TRY(arena_get(ctx, 6, &d_mm));
    SB_TRY(arena_get(ctx, (size_t)PATCH_CAP, &d_list));
    d_cursor = d_nvox + n_clouds;
    d_list_n = d_nvox + 2 * n_clouds;
    int* d_unproven = d_nvox + 2 * n_clouds + 1;
    SB_TRY(table_upload(ctx, d_clouds, hc.data(), sizeof(VoxCloud) * n_clouds));
    SB_TRY(table_upload(ctx, d_tile_cloud, h_tile_cloud.data(), sizeof(int) * (size_t)n_tiles));
    trace_mark(ctx, "vox:begin");
    SB_LAUNCH(ctx, k_vox_clear, ceil_div(n_slots, 256), 256, 0, d_tkeys, d_table, n_slots);
    trace_mark(ctx, "vox:clear");
    SB_CUDA(ctx, cudaMemsetAsync(d_nvox, 0, sizeof(int) * (2 * (size_t)n_clouds + 2), ctx->stream));
    SB_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    i64 init[6] = {INT64_MAX, INT64_MAX, INT64_MAX, INT64_MIN, INT64_MIN, INT64_MIN};
    SB_CUDA(ctx, cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    static const int per_sm = getenv("SB_VOX_GRID") ? atoi(getenv("SB_VOX_GRID")) : 64;
    // at most sm_count * 64 CTAs, and at least ~12 tiles for each of them: a CTA that takes one or two tiles spends
    // as long starting up, reducing its key range and waiting for its slowest warp as it spends on rows (a 225-scan
    // chunk of the pipelined upload ran at half the rate of the 1000-scan batch with one CTA per tile)
    static const int tiles_per_cta = getenv("SB_VOX_TPC") ? atoi(getenv("SB_VOX_TPC")) : 4;
    i64 want_grid = (n_tiles + tiles_per_cta - 1) / (tiles_per_cta > 0 ? tiles_per_cta : 1);
    if (want_grid < (i64)ctx->sm_count * 4) want_grid = (i64)ctx->sm_count * 4;
    if (want_grid > (i64)ctx->sm_count * per_sm) want_grid = (i64)ctx->sm_count * per_sm;
    const int pgrid = (int)(n_tiles < want_grid ? n_tiles : want_grid);
    // Tiles are visited in a scattered order: in sequence order the resident CTAs all work on the same five or six
    // scans, whose neighbouring beams hit the same voxels, and their atomics queue up on the same slots (measured:
    // 18-20 % faster scattered).  But scattered over the WHOLE batch every cloud's table is live at once — 1.9 GB for
    // 2048 scans against 126 MB of L2, so every accumulator update went to DRAM (ncu: 12.3 GB moved for 5.8 GB of
    // input).  So the scatter stays inside a window of consecutive tiles worth ~48 clouds (tile = w0 + (seq - w0) *
    // stride mod window, stride coprime with the window): their tables (~45 MB) stay in L2.
    auto gcd = [](i64 a, i64 b) { while (b) { i64 t = a % b; a = b; b = t; } return a; };
    auto coprime_stride = [&](i64 n) -> i64 {
        static const i64 stride = getenv("SB_VOX_PERM") ? atoll(getenv("SB_VOX_PERM")) : 7919;
        if (n <= 2) return 0;
        i64 p = stride % n;
        while (p > 1 && gcd(p, n) != 1) ++p;
        return (p <= 1 || p >= n) ? 0 : p;
    };
    static const int win_clouds = getenv("SB_VOX_WINDOW") ? atoi(getenv("SB_VOX_WINDOW")) : 48;
    i64 win = n_clouds > 0 ? (n_tiles * (i64)(win_clouds > 0 ? win_clouds : 1) + n_clouds - 1) / n_clouds : n_tiles;
    if (win < 1) win = 1;
    if (win > n_tiles) win = n_tiles > 0 ? n_tiles : 1;
    const i64 perm = coprime_stride(win), perm_last = coprime_stride(n_tiles % win);
    int expo = 0;
    const int pow2 = frexp(voxel, &expo) == 0.5 ? 1 : 0;   // voxel = 2^e: c * (1 / voxel) is the exact quotient
    SB_LAUNCH(ctx, k_vox_insert, pgrid, 256, 0, src, d_clouds, d_tile_cloud, n_tiles, voxel, 1.0 / voxel, pow2, d_tkeys,
              d_table, d_nvox, d_mm, ctx->d_flags, win, perm, perm_last);
    trace_mark(ctx, "vox:insert");
    // ---- the only host round trip: flags, key range, voxels per cloud (into pinned memory: a pageable target would
    // make the driver stage the copies)
    std::vector<int> nvox((size_t)n_clouds);
    i64 mm[6];
    int flags = 0;
    {
        const size_t nv_bytes = sizeof(int) * (size_t)n_clouds;
        SB_TRY(pinned_reserve(ctx, nv_bytes + sizeof(mm) + 64));
        char* hp = ctx->pinned;
        const size_t o_mm = (nv_bytes + 15) & ~(size_t)15, o_fl = o_mm + sizeof(mm);
        SB_CUDA(ctx, cudaMemcpyAsync(hp, d_nvox, nv_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(ctx, cudaMemcpyAsync(hp + o_mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(ctx, cudaMemcpyAsync(hp + o_fl, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(nvox.data(), hp, nv_bytes);
        memcpy(mm, hp + o_mm, sizeof(mm));
        memcpy(&flags, hp + o_fl, sizeof(int));
    }
    if (flags & FLAG_NONFINITE) return fail(ctx, SB_ERR_RANGE, "voxel: non-finite coordinate or |coord/voxel| >= 4e18");
    if (flags & FLAG_TABLE_FULL) {  // more voxels per point than the tables were sized for: this call takes the
        ctx->vox_slots_per_point = 2.0;  // sort, the next one gets worst-case tables (one voxel per point)
        return SB_OK;
    }
    if (flags & FLAG_KEY_RANGE) return SB_OK;
    h_out_off[0] = 0;
    for (int c = 0; c < n_clouds; ++c) h_out_off[c + 1] = h_out_off[c] + nvox[c];
    const i64 m = h_out_off[n_clouds];
    {  // Size the next call's tables for 3 slots per voxel at this batch's AVERAGE density.  The tables are sized per
        // cloud from its row count, so a cloud that is richer than the average runs at a higher load: with 2 slots
        // per voxel the 64 different scenes of config C5 put some clouds at 80-100 % load and the linear probe chains
        // made the whole insert 25 % slower (9.5 vs 7.6 ms per 2048 scans; 1.5: 13.6 ms; 4: 7.7 ms).
        static const double per_voxel = getenv("SB_VOX_SLOTS") ? atof(getenv("SB_VOX_SLOTS")) : 3.0;
        double r = per_voxel * (double)m / (double)n;
        ctx->vox_slots_per_point = r < 0.0625 ? 0.0625 : (r > 2.0 ? 2.0 : r);
    }
    int bx = bits_for((u64)(mm[3] - mm[0])), by = bits_for((u64)(mm[4] - mm[1])), bz = bits_for((u64)(mm[5] - mm[2]));
    VoxelPack P;
    P.minx = mm[0]; P.miny = mm[1]; P.minz = mm[2];
    P.sy = bz;
    P.sx = by + bz;
    P.mask_z = (1ull << bz) - 1ull;
    P.mask_y = (1ull << by) - 1ull;
    // ---- voxels of every cloud in key order
    u64 *ka, *kb, *ks;
    uint32_t *va, *vb, *vs;
    i64* d_out_off;
    SB_TRY(arena_get(ctx, (size_t)m, &ka));
    SB_TRY(arena_get(ctx, (size_t)m, &kb));
    SB_TRY(arena_get(ctx, (size_t)m, &va));
    SB_TRY(arena_get(ctx, (size_t)m, &vb));
    SB_TRY(arena_get(ctx, (size_t)n_clouds + 1, &d_out_off));
    SB_TRY(table_upload(ctx, d_out_off, h_out_off, sizeof(i64) * (n_clouds + 1)));
    trace_mark(ctx, "vox:after-sync1");
    SB_LAUNCH(ctx, k_vox_list, ceil_div(n_slots, 256), 256, 0, d_clouds, n_clouds, n_slots, d_tkeys, d_out_off, d_cursor, P,
              ka, va);
    trace_mark(ctx, "vox:list");
    SB_TRY(segmented_sort_pairs(ctx, ka, kb, va, vb, h_out_off, n_clouds, bx + by + bz, &ks, &vs));
    trace_mark(ctx, "vox:sort");
    SB_LAUNCH(ctx, k_vox_finalize, ceil_div(m, 256), 256, 0, ks, vs, m, voxel, P, d_table, d_out_xyz, d_out_keys,
              d_unproven);
    // ---- ordered re-summation of the voxels that could not be proven exact
    SB_LAUNCH(ctx, k_vox_collect, pgrid, 256, 0, d_clouds, d_tile_cloud, n_tiles, src, voxel, d_tkeys, d_table, d_list, d_list_n,
              ctx->d_flags, d_unproven);
    SB_LAUNCH(ctx, k_vox_patch, 1, 1024, 0, src, d_list, d_list_n, d_out_xyz);
    trace_mark(ctx, "vox:finalize+patch");
    SB_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(&flags, ctx->pinned, sizeof(int));
    if (flags & FLAG_PATCH_OVERFLOW) return SB_OK;  // e.g. arbitrary fp64 input: the sort handles it
    if (getenv("SB_VOX_TRACE")) trace_dump(ctx);
    *done = 1;
    return SB_OK;
}

__global__ void __launch_bounds__(256) k_widen(const PointSrc src, i64 n, double* __restrict__ out) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y, z;
    load_xyz(src, i, x, y, z);
    out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = z;
}

int voxel_downsample_src(Ctx* ctx, const PointSrc src, const i64* h_off, int n_clouds, double voxel, double* d_out_xyz,
                         i64* h_out_off, i64* d_out_keys) {
    if (h_off[0] != 0) return fail(ctx, SB_ERR_INVALID_ARG, "voxel: offsets must start at 0");
    const i64 n = h_off[n_clouds];
    if (voxel > 0 && n > 0 && !ctx->vox_force_sort) {
        int done = 0;
        SB_TRY(voxel_hashed_dev(ctx, src, h_off, n_clouds, voxel, d_out_xyz, h_out_off, d_out_keys, &done));
        ctx->vox_last_path = done ? 1 : 2;
        if (done) return SB_OK;
    } else {
        ctx->vox_last_path = 2;
    }
    const double* d_xyz = static_cast<const double*>(src.base);
    if (src.f32 || src.stride != 3) {  // the sort path works on packed fp64 rows
        double* tmp;
        SB_TRY(arena_get(ctx, (size_t)3 * (n > 0 ? n : 1), &tmp));
        if (n > 0) SB_LAUNCH(ctx, k_widen, ceil_div(n, 256), 256, 0, src, n, tmp);
        d_xyz = tmp;
    }
    return voxel_sorted_dev(ctx, d_xyz, h_off, n_clouds, voxel, d_out_xyz, h_out_off, d_out_keys);
}

int voxel_downsample_dev(Ctx* ctx, const double* d_xyz, const i64* h_off, int n_clouds, double voxel,
                         double* d_out_xyz, i64* h_out_off, i64* d_out_keys) {
    PointSrc src;
    src.base = d_xyz; src.f32 = 0; src.stride = 3;
    return voxel_downsample_src(ctx, src, h_off, n_clouds, voxel, d_out_xyz, h_out_off, d_out_keys);
}

}  // namespace sb
