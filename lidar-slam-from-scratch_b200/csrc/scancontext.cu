// scancontext.cu — Scan Context descriptor build and column-shift distance search; replaces
// slam::ScanContext::compute / distance / column_shifted_distance / ring_key / sector_key
// (slam_viz/include/slam_viz/core/scan_context.hpp:44-82, 90-102, 121-142, 107-118).
//
// Build: one CTA per cloud, 1200 shared-memory bins, atomicMax on the order-preserving integer image of z, so the
// result is exact and independent of the order in which points arrive.  Range/sector use the reference's fp64
// expressions (sqrt, atan2 + pi, division by the bin size, truncation, clamp).
// Search: one (database entry, shift) pair per thread.  Each thread walks the 1200 cells in the reference's order
// (ring outer, sector inner) with separately rounded multiply and add (this file is compiled with -fmad=false), so
// sum_ab, sum_aa and sum_bb — and therefore the distance — are bit-identical to the scalar reference; the minimum
// over the 60 shifts is an exact reduction.  The query and the entry are staged ring-major in shared memory, so a
// warp's 32 shifts read 32 consecutive doubles (no bank conflicts) and the query cell is a broadcast.
#include "common.cuh"

namespace sb {

static constexpr int RINGS = SB_SC_RINGS, SECTORS = SB_SC_SECTORS, CELLS = SB_SC_SIZE;

__global__ void __launch_bounds__(256) k_sc_compute(const double* __restrict__ xyz, const i64* __restrict__ off,
                                                    double* __restrict__ desc) {
    __shared__ long long bins[CELLS];
    const int c = blockIdx.x;
    const long long empty = ordered_from_double(-1.7976931348623157e308);  // scan_context.hpp:46
    for (int t = threadIdx.x; t < CELLS; t += blockDim.x) bins[t] = empty;
    __syncthreads();
    const double ring_size = 80.0 / RINGS;                            // scan_context.hpp:47
    const double sector_size = 2.0 * 3.14159265358979323846 / SECTORS;  // scan_context.hpp:48
    const i64 b = off[c], e = off[c + 1];
    for (i64 i = b + threadIdx.x; i < e; i += blockDim.x) {
        double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        double range = sqrt(x * x + y * y);                       // scan_context.hpp:56
        double angle = atan2(y, x) + 3.14159265358979323846;      // scan_context.hpp:57
        if (range > 80.0 || range < 0.1) continue;                // scan_context.hpp:59
        int ring = (int)(range / ring_size);                      // scan_context.hpp:62
        int sector = (int)(angle / sector_size);                  // scan_context.hpp:63
        ring = min(max(ring, 0), RINGS - 1);                      // scan_context.hpp:65-66
        sector = min(max(sector, 0), SECTORS - 1);
        if (z == z) atomicMax(&bins[sector * RINGS + ring], ordered_from_double(z));  // scan_context.hpp:69-71
    }
    __syncthreads();
    for (int t = threadIdx.x; t < CELLS; t += blockDim.x) {
        double v = double_from_ordered(bins[t]);
        if (v < -1000.0) v = 0.0;  // scan_context.hpp:77-81
        desc[(i64)c * CELLS + t] = v;  // column-major 20x60: (ring i, sector j) at j*20 + i
    }
}

int sc_compute_dev(Ctx* ctx, const double* d_xyz, const i64* d_off, int n_clouds, double* d_desc) {
    if (n_clouds <= 0) return SB_OK;
    SB_LAUNCH(ctx, k_sc_compute, (unsigned)n_clouds, 256, 0, d_xyz, d_off, d_desc);
    return SB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
static constexpr int ENTRIES_PER_CTA = 2;

__global__ void __launch_bounds__(64 * ENTRIES_PER_CTA) k_sc_distance(const double* __restrict__ query,
                                                                      const double* __restrict__ db, int n_db,
                                                                      double* __restrict__ out) {
    __shared__ double sa[CELLS];                      // query, ring-major: (i, j) at i*60 + j
    __shared__ double sb_[ENTRIES_PER_CTA][CELLS];    // entries, ring-major
    __shared__ double smin[ENTRIES_PER_CTA][2];
    const int sub = threadIdx.x >> 6;       // entry within the CTA
    const int s = threadIdx.x & 63;         // shift handled by this thread (60..63 idle)
    const int entry = blockIdx.x * ENTRIES_PER_CTA + sub;
    for (int t = threadIdx.x; t < CELLS; t += blockDim.x) {
        int j = t / RINGS, i = t - j * RINGS;
        sa[i * SECTORS + j] = query[t];
    }
    if (entry < n_db) {
        const double* g = db + (i64)entry * CELLS;
        for (int t = s; t < CELLS; t += 64) {
            int j = t / RINGS, i = t - j * RINGS;
            sb_[sub][i * SECTORS + j] = g[t];
        }
    }
    __syncthreads();
    double d = 1.7976931348623157e308;
    if (entry < n_db && s < SECTORS) {
        double sum_ab = 0.0, sum_aa = 0.0, sum_bb = 0.0;
        const double* B = sb_[sub];
        for (int i = 0; i < RINGS; ++i) {           // scan_context.hpp:126-135, same order
            const double* ar = sa + i * SECTORS;
            const double* br = B + i * SECTORS;
            int jj = s;
#pragma unroll 4
            for (int j = 0; j < SECTORS; ++j) {
                double a = ar[j];
                double b = br[jj];
                sum_ab += a * b;
                sum_aa += a * a;
                sum_bb += b * b;
                jj = jj + 1 == SECTORS ? 0 : jj + 1;
            }
        }
        double norm = sqrt(sum_aa) * sqrt(sum_bb);                 // scan_context.hpp:137
        d = norm < 1e-10 ? 1.0 : 1.0 - sum_ab / norm;             // scan_context.hpp:138-141
        if (!(d == d)) d = 1.7976931348623157e308;                 // NaN never wins `dist < min_dist` (:96)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = fmin(d, shfl_d_xor(d, o));
    if ((threadIdx.x & 31) == 0) smin[sub][(threadIdx.x >> 5) & 1] = d;
    __syncthreads();
    if (entry < n_db && s == 0) out[entry] = fmin(smin[sub][0], smin[sub][1]);  // scan_context.hpp:93-99
}

int sc_distance_dev(Ctx* ctx, const double* d_query_desc, const double* d_db, int n_db, double* d_out) {
    if (n_db <= 0) return SB_OK;
    SB_LAUNCH(ctx, k_sc_distance, (unsigned)ceil_div(n_db, ENTRIES_PER_CTA), 64 * ENTRIES_PER_CTA, 0, d_query_desc,
              d_db, n_db, d_out);
    return SB_OK;
}

// ring_key: row means over the 60 sectors; sector_key: column means over the 20 rings (scan_context.hpp:107-118)
__global__ void k_sc_keys(const double* __restrict__ desc, double* __restrict__ ring_key, double* __restrict__ sector_key) {
    int t = threadIdx.x;
    if (t < RINGS) {
        double s = 0.0;
        for (int j = 0; j < SECTORS; ++j) s += desc[j * RINGS + t];
        ring_key[t] = s / (double)SECTORS;
    } else if (t < RINGS + SECTORS) {
        int j = t - RINGS;
        double s = 0.0;
        for (int i = 0; i < RINGS; ++i) s += desc[j * RINGS + i];
        sector_key[j] = s / (double)RINGS;
    }
}

int sc_keys_dev(Ctx* ctx, const double* d_desc, double* d_ring, double* d_sector) {
    SB_LAUNCH(ctx, k_sc_keys, 1, 96, 0, d_desc, d_ring, d_sector);
    return SB_OK;
}

}  // namespace sb
