// scan_sort.cu — device-wide exclusive scan and a stable, SEGMENTED least-significant-digit radix sort.
//
// The sort is the grouping primitive of the engine:
//   * voxel grid   (file_utils.cpp:148-196 in the reference uses an unordered_map): points are sorted by packed
//     voxel key inside each cloud; stability keeps the members of a voxel in ascending input order so that the
//     centroid sum has the reference's summation order;
//   * spatial index (kdtree.hpp:87-110 in the reference is a recursive nth_element build): points are sorted by
//     space-filling-curve index (Hilbert) inside each cloud.
// Segments (clouds) never mix: tiles do not cross segment boundaries and the digit histogram is laid out
// [segment][digit][tile-in-segment], so one flat exclusive scan yields segment-local stable destinations.
#include "common.cuh"

namespace sb {

// =============================================================================================================
// exclusive scan (uint32), three phases; 4096 elements per block
// =============================================================================================================
static constexpr int SCAN_THREADS = 1024;
static constexpr int SCAN_PER_THREAD = 4;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t* smem33, uint32_t* total) {
    // warp inclusive scan
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem33[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = smem33[lane];
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem33[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) smem33[32] = winc;
    }
    __syncthreads();
    uint32_t res = smem33[warp] + inc - v;
    if (total) *total = smem33[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t* __restrict__ in, i64 n,
                                                              uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t sm[33];
    i64 base = (i64)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; ++j) {
        i64 i = base + (i64)j * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t tot;
    block_exclusive_scan_1024(s, sm, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_blocks(uint32_t* __restrict__ block_sums, int nb,
                                                              uint32_t* __restrict__ total_out) {
    __shared__ uint32_t sm[33];
    uint32_t carry = 0;
    for (int base = 0; base < nb; base += SCAN_THREADS) {
        int i = base + threadIdx.x;
        uint32_t v = i < nb ? block_sums[i] : 0u;
        uint32_t tot;
        uint32_t ex = block_exclusive_scan_1024(v, sm, &tot);
        if (i < nb) block_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_down(const uint32_t* __restrict__ in,
                                                            uint32_t* __restrict__ out, i64 n,
                                                            const uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t sm[33];
    // blocked arrangement: thread t owns elements [4t, 4t+4) of the tile
    i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)threadIdx.x * SCAN_PER_THREAD;
    uint32_t v[SCAN_PER_THREAD];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; ++j) {
        v[j] = (base + j < n) ? in[base + j] : 0u;
        s += v[j];
    }
    uint32_t ex = block_exclusive_scan_1024(s, sm, nullptr) + block_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; ++j) {
        if (base + j < n) out[base + j] = ex;
        ex += v[j];
    }
}

int exclusive_scan_u32(Ctx* ctx, const uint32_t* d_in, uint32_t* d_out, i64 n, uint32_t* d_total) {
    if (n <= 0) {
        if (d_total) SB_CUDA(ctx, cudaMemsetAsync(d_total, 0, sizeof(uint32_t), ctx->stream));
        return SB_OK;
    }
    int nb = ceil_div(n, SCAN_TILE);
    uint32_t* bs;
    SB_TRY(arena_get(ctx, (size_t)nb, &bs));
    SB_LAUNCH(ctx, k_scan_reduce, nb, SCAN_THREADS, 0, d_in, n, bs);
    SB_LAUNCH(ctx, k_scan_blocks, 1, SCAN_THREADS, 0, bs, nb, d_total);
    SB_LAUNCH(ctx, k_scan_down, nb, SCAN_THREADS, 0, d_in, d_out, n, bs);
    return SB_OK;
}

// =============================================================================================================
// segmented stable radix sort, 8-bit digits, 2048-element tiles
// =============================================================================================================
static constexpr int SORT_THREADS = 256;
static constexpr int SORT_ITEMS = 8;
static constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 2048
static constexpr int RADIX = 256;

struct SortTile {
    i64 start;       // first element (global)
    i64 hist_base;   // 256 * (tiles in earlier segments)
    int count;       // 1..2048
    int tiles_seg;   // tiles in this tile's segment
    int t_local;     // index of this tile inside its segment
    int pad;
};

__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const u64* __restrict__ keys,
                                                            const SortTile* __restrict__ tiles, int shift,
                                                            uint32_t* __restrict__ H) {
    __shared__ uint32_t hist[RADIX];
    SortTile t = tiles[blockIdx.x];
    hist[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        int i = j * SORT_THREADS + threadIdx.x;
        if (i < t.count) {
            unsigned d = (unsigned)(keys[t.start + i] >> shift) & 0xffu;
            atomicAdd(&hist[d], 1u);
        }
    }
    __syncthreads();
    H[t.hist_base + (i64)threadIdx.x * t.tiles_seg + t.t_local] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const u64* __restrict__ keys_in,
                                                               const uint32_t* __restrict__ vals_in,
                                                               u64* __restrict__ keys_out,
                                                               uint32_t* __restrict__ vals_out,
                                                               const SortTile* __restrict__ tiles, int shift,
                                                               const uint32_t* __restrict__ Hs) {
    __shared__ uint32_t cnt[SORT_THREADS / 32][RADIX];
    SortTile t = tiles[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (SORT_THREADS / 32) * RADIX; i += SORT_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();

    u64 key[SORT_ITEMS];
    uint32_t val[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
    // element order inside the tile: warp-major, then row, then lane  (index = warp*256 + row*32 + lane)
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        int i = warp * (32 * SORT_ITEMS) + r * 32 + lane;
        bool valid = i < t.count;
        key[r] = valid ? keys_in[t.start + i] : 0ull;
        val[r] = valid ? vals_in[t.start + i] : 0u;
        unsigned d = valid ? ((unsigned)(key[r] >> shift) & 0xffu) : 256u;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (valid && lane == leader) {
            prev = cnt[warp][d];
            cnt[warp][d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[r] = prev + __popc(peers & lanemask_lt());
        __syncwarp();
    }
    __syncthreads();
    {
        // thread d turns the per-warp counts of digit d into exclusive offsets + the tile's global base
        unsigned d = threadIdx.x;
        uint32_t running = Hs[t.hist_base + (i64)d * t.tiles_seg + t.t_local];
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; ++w) {
            uint32_t c = cnt[w][d];
            cnt[w][d] = running;
            running += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        int i = warp * (32 * SORT_ITEMS) + r * 32 + lane;
        if (i < t.count) {
            unsigned d = (unsigned)(key[r] >> shift) & 0xffu;
            uint32_t dst = cnt[warp][d] + rank[r];
            keys_out[dst] = key[r];
            vals_out[dst] = val[r];
        }
    }
}


// =============================================================================================================
// one-CTA-per-segment sort in shared memory (segments of up to SEG_CAP rows, keys of up to 32 bits)
// =============================================================================================================
// A downsampled scan is 7-10 k rows: the tiled sort above spends five launches per digit on it (20 for a 30-bit
// code) and all of them are latency.  Here a CTA keeps the segment's keys and a 16-bit permutation in shared memory
// and does every digit pass there: rank by __match_any inside each warp's contiguous block of rows (stable), one
// block scan over the [digit][warp] counters, scatter of the permutation; keys and values move once, at the end.
static constexpr int SEG_CAP = 12288;
static constexpr int SEG_THREADS = 1024;
static constexpr int SEG_EPT = SEG_CAP / SEG_THREADS;   // batches of 32 rows per warp, at most
// counters: [warp][digit] with a row stride of 258 shorts = 129 words — odd, so that neither the ranking (one warp,
// 32 different digits) nor the scan (eight warps' counters of one digit per thread) piles onto a few banks; as
// [digit][warp] the 32 lanes of a ranking step hit two banks.
#ifndef SB_SORT_CSTR
#define SB_SORT_CSTR 258
#endif
static constexpr int SEG_CSTR = SB_SORT_CSTR;
static constexpr size_t SEG_SMEM = (size_t)SEG_CAP * 4 + (size_t)SEG_CAP * 2 * 2 + (size_t)SEG_CSTR * 32 * 2;

__global__ void __launch_bounds__(SEG_THREADS) k_seg_sort_smem(const u64* __restrict__ kin, const uint32_t* __restrict__ vin,
                                                               u64* __restrict__ kout, uint32_t* __restrict__ vout,
                                                               const i64* __restrict__ seg_off, int key_bits) {
    extern __shared__ __align__(16) unsigned char seg_sm[];
    uint32_t* keys = reinterpret_cast<uint32_t*>(seg_sm);
    unsigned short* pin = reinterpret_cast<unsigned short*>(keys + SEG_CAP);
    unsigned short* pout = pin + SEG_CAP;
    unsigned short* cnt = pout + SEG_CAP;  // [warp][digit]: rows of this digit in this warp's block, then their base
    __shared__ uint32_t s_scan[33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const i64 base = seg_off[blockIdx.x];
    const int n = (int)(seg_off[blockIdx.x + 1] - base);
    if (n <= 0) return;
    for (int i = tid; i < n; i += SEG_THREADS) {
        keys[i] = (uint32_t)kin[base + i];
        pin[i] = (unsigned short)i;
    }
    const int per = (((n + 31) / 32) + 31) & ~31;  // rows per warp: a multiple of 32, 32 * per >= n
    const int w0 = warp * per, w1 = min(n, w0 + per);
    for (int shift = 0; shift < key_bits; shift += 8) {
        for (int i = tid; i < SEG_CSTR * 32; i += SEG_THREADS) cnt[i] = 0;
        __syncthreads();
        unsigned short rank[SEG_EPT];
#pragma unroll
        for (int j = 0; j < SEG_EPT; ++j) {
            rank[j] = 0;
            if (j * 32 < per) {  // warp-uniform
                const int i = w0 + j * 32 + lane;
                const bool act = i < w1;
                const unsigned d = act ? ((keys[pin[i]] >> shift) & (RADIX - 1)) : (unsigned)RADIX + lane;
                const unsigned m = __match_any_sync(0xffffffffu, d);
                unsigned short old = 0;
                if (act) old = cnt[warp * SEG_CSTR + d];
                __syncwarp();
                if (act) {
                    rank[j] = (unsigned short)(old + __popc(m & lanemask_lt()));
                    if ((int)(__ffs(m) - 1) == lane) cnt[warp * SEG_CSTR + d] = (unsigned short)(old + __popc(m));
                }
                __syncwarp();
            }
        }
        __syncthreads();
        {   // exclusive scan over the 8192 counters in (digit, warp) order: thread t takes digit t / 4, warps 8 (t % 4) ..
            unsigned short* c8 = cnt + (tid & 3) * 8 * SEG_CSTR + (tid >> 2);
            uint32_t loc[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { loc[q] = sum; sum += c8[q * SEG_CSTR]; }
            uint32_t total;
            const uint32_t off = block_exclusive_scan_1024(sum, s_scan, &total);
#pragma unroll
            for (int q = 0; q < 8; ++q) c8[q * SEG_CSTR] = (unsigned short)(off + loc[q]);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SEG_EPT; ++j) {
            if (j * 32 < per) {
                const int i = w0 + j * 32 + lane;
                if (i < w1) {
                    const unsigned short idx = pin[i];
                    const unsigned d = (keys[idx] >> shift) & (RADIX - 1);
                    pout[cnt[warp * SEG_CSTR + d] + rank[j]] = idx;
                }
            }
        }
        __syncthreads();
        unsigned short* t = pin; pin = pout; pout = t;
    }
    for (int i = tid; i < n; i += SEG_THREADS) {
        const unsigned short idx = pin[i];
        kout[base + i] = (u64)keys[idx];
        vout[base + i] = vin[base + idx];
    }
}

int segmented_sort_pairs(Ctx* ctx, u64* keys_a, u64* keys_b, uint32_t* vals_a, uint32_t* vals_b,
                         const i64* h_seg_off, int n_seg, int key_bits, u64** out_keys, uint32_t** out_vals) {
    *out_keys = keys_a;
    *out_vals = vals_a;
    i64 n_total = h_seg_off[n_seg] - h_seg_off[0];
    if (h_seg_off[0] != 0) return fail(ctx, SB_ERR_INVALID_ARG, "sort: offsets must start at 0");
    if (n_total <= 0 || key_bits <= 0) return SB_OK;
    if (h_seg_off[n_seg] >= (i64)0xffffffffLL) return fail(ctx, SB_ERR_RANGE, "sort: more than 2^32-1 rows in one call");
    // short segments with short keys: one CTA per segment, everything in shared memory
    {
        i64 longest = 0;
        for (int s = 0; s < n_seg; ++s) longest = std::max(longest, h_seg_off[s + 1] - h_seg_off[s]);
        static const bool no_smem_sort = getenv("SB_SORT_TILED") != nullptr;
        if (!no_smem_sort && key_bits <= 32 && longest <= SEG_CAP) {
            // per device, and cheap: set on every call rather than tracked per context
            SB_CUDA(ctx, cudaFuncSetAttribute(k_seg_sort_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEG_SMEM));
            i64* d_seg;
            SB_TRY(arena_get(ctx, (size_t)n_seg + 1, &d_seg));
            SB_TRY(table_upload(ctx, d_seg, h_seg_off, sizeof(i64) * ((size_t)n_seg + 1)));
            SB_LAUNCH(ctx, k_seg_sort_smem, (unsigned)n_seg, SEG_THREADS, SEG_SMEM, keys_a, vals_a, keys_b, vals_b, d_seg,
                      key_bits);
            *out_keys = keys_b;
            *out_vals = vals_b;
            return SB_OK;
        }
    }
    // tile table (host-built, staged through pinned memory)
    i64 n_tiles = 0;
    for (int s = 0; s < n_seg; ++s) n_tiles += (h_seg_off[s + 1] - h_seg_off[s] + SORT_TILE - 1) / SORT_TILE;
    std::vector<SortTile> ht_vec((size_t)n_tiles);
    SortTile* ht = ht_vec.data();
    i64 ti = 0;
    for (int s = 0; s < n_seg; ++s) {
        i64 n = h_seg_off[s + 1] - h_seg_off[s];
        int nt = (int)((n + SORT_TILE - 1) / SORT_TILE);
        i64 hb = (i64)RADIX * ti;
        for (int t = 0; t < nt; ++t) {
            SortTile& T = ht[ti + t];
            T.start = h_seg_off[s] + (i64)t * SORT_TILE;
            T.hist_base = hb;
            i64 rem = n - (i64)t * SORT_TILE;
            T.count = (int)(rem < SORT_TILE ? rem : SORT_TILE);
            T.tiles_seg = nt;
            T.t_local = t;
            T.pad = 0;
        }
        ti += nt;
    }
    SortTile* d_tiles;
    uint32_t* H;
    SB_TRY(arena_get(ctx, (size_t)n_tiles, &d_tiles));
    SB_TRY(arena_get(ctx, (size_t)n_tiles * RADIX, &H));
    SB_TRY(table_upload(ctx, d_tiles, ht, (size_t)n_tiles * sizeof(SortTile)));
    u64* kin = keys_a; u64* kout = keys_b;
    uint32_t* vin = vals_a; uint32_t* vout = vals_b;
    for (int shift = 0; shift < key_bits; shift += 8) {
        SB_LAUNCH(ctx, k_sort_hist, (unsigned)n_tiles, SORT_THREADS, 0, kin, d_tiles, shift, H);
        SB_TRY(exclusive_scan_u32(ctx, H, H, n_tiles * RADIX, nullptr));
        SB_LAUNCH(ctx, k_sort_scatter, (unsigned)n_tiles, SORT_THREADS, 0, kin, vin, kout, vout, d_tiles, shift, H);
        u64* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    *out_keys = kin;
    *out_vals = vin;
    return SB_OK;
}

}  // namespace sb
