// loop.cu — device-resident loop-closure database; replaces slam::LoopClosureDetector
// (slam_viz/include/slam_viz/core/loop_closure.hpp:41-149).
//
// The reference keeps a vector of Eigen descriptors plus a full copy of every cloud on the host and scans it
// linearly (loop_closure.hpp:78-89), then verifies candidates one by one with ICP (:95-122).  Here descriptors
// and clouds live in two growing device pools; the Scan Context search is one kernel over the whole database and
// the ICP verifications of a chunk of candidates run as ONE batch (they are independent); acceptance then walks the
// candidates in the reference's (distance, entry) order so the accepted set is the same as the sequential loop's.
// With world > 1, entry i is owned by rank i % world; every rank also keeps the newest entry (the query).
#include "common.cuh"

#include <algorithm>

namespace sb {

// Pools only ever grow, by doubling, and the outgrown buffer is retired (freed with the detector) rather than
// released: cudaMalloc / cudaFree stall the calling thread for up to hundreds of milliseconds (measured on the B200
// pool: 100 ms / 580 ms worst case) while a frame takes 1.7 ms, so the streaming path must not reach them in steady
// state.  The first allocation holds ~1300 voxel-downsampled scans / 4096 descriptors; sb_loop_reserve sizes the
// pools for a known sequence length up front.
static const size_t FIRST_CLOUD_ROWS = (size_t)11 << 20;   // 264 MB of fp64 xyz rows
static const size_t FIRST_DESC_SLOTS = 4096;               // 39 MB

static int grow(sb_loop* L, double** buf, size_t* cap, size_t need, size_t used, size_t unit, size_t first) {
    if (need <= *cap) return SB_OK;
    Ctx* ctx = L->ctx;
    size_t ncap = *cap ? *cap : first;
    while (ncap < need) ncap *= 2;
    double* nb;
    SB_CUDA(ctx, cudaMalloc(&nb, ncap * unit * sizeof(double)));
    if (used) SB_CUDA(ctx, cudaMemcpyAsync(nb, *buf, used * unit * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    if (*buf) L->retired.push_back(*buf);
    *buf = nb;
    *cap = ncap;
    return SB_OK;
}

int loop_reserve(sb_loop* L, i64 n_entries, i64 total_rows) {
    size_t slots = L->entry_id.size();
    i64 rows = L->cloud_off.empty() ? 0 : L->cloud_off.back();
    if (n_entries > 0) SB_TRY(grow(L, &L->d_desc, &L->desc_cap, (size_t)n_entries, slots, SB_SC_SIZE, (size_t)n_entries));
    if (n_entries > 0) SB_TRY(grow(L, &L->d_meta, &L->meta_cap, (size_t)n_entries, slots, 1, (size_t)n_entries));
    if (total_rows > 0) SB_TRY(grow(L, &L->d_clouds, &L->cloud_cap, (size_t)total_rows, (size_t)rows, 3, (size_t)total_rows));
    SB_CUDA(L->ctx, cudaStreamSynchronize(L->ctx->stream));
    return SB_OK;
}

int loop_add(sb_loop* L, const double* xyz, i64 n, int frame_idx, const double* desc) {
    Ctx* ctx = L->ctx;
    // drop a guest (non-owned former query) before appending
    if (L->last_is_guest) {
        L->entry_id.pop_back();
        L->frame_idx.pop_back();
        L->cloud_off.pop_back();
        L->last_is_guest = false;
    }
    int id = L->n_global++;
    size_t slots = L->entry_id.size();
    i64 rows = L->cloud_off.empty() ? 0 : L->cloud_off.back();
    if (L->cloud_off.empty()) L->cloud_off.push_back(0);
    SB_TRY(grow(L, &L->d_desc, &L->desc_cap, slots + 1, slots, SB_SC_SIZE, FIRST_DESC_SLOTS));
    SB_TRY(grow(L, &L->d_meta, &L->meta_cap, slots + 1, slots, 1, FIRST_DESC_SLOTS));
    {
        const int meta[2] = {frame_idx, id};
        SB_CUDA(ctx, cudaMemcpyAsync(L->d_meta + slots, meta, sizeof(meta), cudaMemcpyHostToDevice, ctx->stream));
    }
    SB_TRY(grow(L, &L->d_clouds, &L->cloud_cap, (size_t)(rows + n), (size_t)rows, 3, FIRST_CLOUD_ROWS));
    if (n > 0)
        SB_CUDA(ctx, cudaMemcpyAsync(L->d_clouds + 3 * rows, xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    double* d_slot = L->d_desc + slots * SB_SC_SIZE;
    if (desc) {
        SB_CUDA(ctx, cudaMemcpyAsync(d_slot, desc, sizeof(double) * SB_SC_SIZE, cudaMemcpyHostToDevice, ctx->stream));
    } else {  // ScanContext sc(points), loop_closure.hpp:55
        i64* d_off;
        SB_TRY(arena_get(ctx, 2, &d_off));
        i64 h[2] = {rows, rows + n};
        SB_CUDA(ctx, cudaMemcpyAsync(d_off, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
        SB_TRY(sc_compute_dev(ctx, L->d_clouds, d_off, 1, d_slot));
    }
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    L->entry_id.push_back(id);
    L->frame_idx.push_back(frame_idx);
    L->cloud_off.push_back(rows + n);
    L->last_frame = frame_idx;
    L->last_is_guest = (id % L->world) != L->rank;
    return SB_OK;
}

// -------------------------------------------------------------------------------------------------------------
// Candidate selection on the device (loop_closure.hpp:78-92: frame-gap filter, threshold, sort by (distance, entry)):
// one block.  Every warp keeps the KSEL best (distance, entry) pairs of its share of the database in registers, entry
// j in lane j, ascending; warp 0 then merges the 32 lists.  The host receives KSEL x 12 bytes and the number of
// entries under the threshold instead of one distance per database entry.
// -------------------------------------------------------------------------------------------------------------
static constexpr int KSEL = SB_LOOP_SELECT_MAX;

struct SelList {   // lane j of a warp holds the j-th best pair
    unsigned long long d;   // order-preserving image of the distance; ~0: empty
    int e;
    __device__ __forceinline__ bool less_than(unsigned long long od, int oe) const { return d < od || (d == od && e < oe); }
    // inserts the candidates of the lanes in `m` (warp-uniform mask), each lane's own (cd, ce)
    __device__ __forceinline__ void insert(unsigned m, unsigned long long cd, int ce, int lane) {
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1u;
            const unsigned long long bd = __shfl_sync(0xffffffffu, cd, b);
            const int be = __shfl_sync(0xffffffffu, ce, b);
            const int pos = __popc(__ballot_sync(0xffffffffu, less_than(bd, be)));   // sorted: a prefix
            if (pos >= 32) continue;
            const unsigned long long ud = __shfl_up_sync(0xffffffffu, d, 1);
            const int ue = __shfl_up_sync(0xffffffffu, e, 1);
            if (lane > pos) { d = ud; e = ue; }
            else if (lane == pos) { d = bd; e = be; }
        }
    }
};

__global__ void __launch_bounds__(1024) k_loop_select(const double* __restrict__ dist, const int2* __restrict__ meta,
                                                      int n_db, int last_frame, int frame_gap, double threshold,
                                                      unsigned long long* __restrict__ out_d, int* __restrict__ out_e,
                                                      int* __restrict__ out_total) {
    __shared__ unsigned long long s_d[32][KSEL];
    __shared__ int s_e[32][KSEL];
    __shared__ int s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    SelList L;
    L.d = ~0ull; L.e = 0x7fffffff;
    int mine = 0;
    for (int base = warp * 32; base < n_db; base += 1024) {
        const int i = base + lane;
        unsigned long long cd = ~0ull;
        int ce = 0x7fffffff;
        bool ok = false;
        if (i < n_db) {
            const int2 mt = meta[i];
            const double dv = dist[i];
            ok = (last_frame - mt.x >= frame_gap) && (dv < threshold);   // loop_closure.hpp:79-81, 86-88 (NaN: false)
            if (ok) {   // order-preserving image of the double (a cosine distance can round to -1e-16)
                const unsigned long long b = (unsigned long long)__double_as_longlong(dv);
                cd = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
                if (cd == ~0ull) cd = ~0ull - 1ull;
                ce = mt.y;
                ++mine;
            }
        }
        // only candidates better than the list's last entry can enter it
        const unsigned long long ld = __shfl_sync(0xffffffffu, L.d, 31);
        const int le = __shfl_sync(0xffffffffu, L.e, 31);
        const unsigned m = __ballot_sync(0xffffffffu, ok && (cd < ld || (cd == ld && ce < le)));
        L.insert(m, cd, ce, lane);
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if (lane == 0 && mine) atomicAdd(&s_total, mine);
    s_d[warp][lane] = L.d;
    s_e[warp][lane] = L.e;
    __syncthreads();
    if (warp != 0) return;
    SelList M;
    M.d = ~0ull; M.e = 0x7fffffff;
    for (int w = 0; w < 32; ++w) {
        const unsigned long long cd = s_d[w][lane];
        const int ce = s_e[w][lane];
        const unsigned long long ld = __shfl_sync(0xffffffffu, M.d, 31);
        const int le = __shfl_sync(0xffffffffu, M.e, 31);
        const unsigned m = __ballot_sync(0xffffffffu, cd != ~0ull && (cd < ld || (cd == ld && ce < le)));
        M.insert(m, cd, ce, lane);
    }
    out_d[lane] = M.d;
    out_e[lane] = M.e;
    if (lane == 0) *out_total = s_total;
}

// loop_closure.hpp:75-92 restricted to the entries this rank owns
int loop_candidates(sb_loop* L, std::vector<std::pair<double, int>>& cand, int limit, int* total) {
    Ctx* ctx = L->ctx;
    cand.clear();
    if (total) *total = 0;
    if (L->n_global < 2) return SB_OK;  // loop_closure.hpp:69
    int slots = (int)L->entry_id.size();
    int n_db = slots - 1;  // the newest entry (the query) is always the last slot
    if (n_db <= 0) return SB_OK;
    double* d_out;
    SB_TRY(arena_get(ctx, (size_t)n_db, &d_out));
    SB_TRY(sc_distance_dev(ctx, L->d_desc + (size_t)(slots - 1) * SB_SC_SIZE, L->d_desc, n_db, d_out));
    if (limit > 0 && limit <= KSEL) {
        unsigned long long* d_sel_d;
        int* d_sel_e;
        SB_TRY(arena_get(ctx, (size_t)KSEL, &d_sel_d));
        SB_TRY(arena_get(ctx, (size_t)KSEL + 1, &d_sel_e));
        SB_LAUNCH(ctx, k_loop_select, 1, 1024, 0, d_out, reinterpret_cast<const int2*>(L->d_meta), n_db, L->last_frame,
                  L->cfg.frame_gap, L->cfg.sc_distance_threshold, d_sel_d, d_sel_e, d_sel_e + KSEL);
        SB_TRY(pinned_reserve(ctx, KSEL * 12 + 16));
        unsigned long long* h_d = reinterpret_cast<unsigned long long*>(ctx->pinned);
        int* h_e = reinterpret_cast<int*>(ctx->pinned + KSEL * 8);
        SB_CUDA(ctx, cudaMemcpyAsync(h_d, d_sel_d, KSEL * 8, cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(ctx, cudaMemcpyAsync(h_e, d_sel_e, (KSEL + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (total) *total = h_e[KSEL];
        for (int i = 0; i < KSEL && i < limit && h_d[i] != ~0ull; ++i) {
            const unsigned long long b = (h_d[i] >> 63) ? (h_d[i] & 0x7fffffffffffffffull) : ~h_d[i];
            double d;
            memcpy(&d, &b, sizeof(d));
            cand.push_back({d, h_e[i]});
        }
        return SB_OK;
    }
    std::vector<double> dist((size_t)n_db);
    SB_CUDA(ctx, cudaMemcpyAsync(dist.data(), d_out, sizeof(double) * n_db, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n_db; ++i) {
        if (L->last_frame - L->frame_idx[i] < L->cfg.frame_gap) continue;            // loop_closure.hpp:79-81
        if (dist[i] < L->cfg.sc_distance_threshold) cand.push_back({dist[i], L->entry_id[i]});  // :86-88
    }
    std::sort(cand.begin(), cand.end());  // loop_closure.hpp:92: (distance, entry) ascending
    if (total) *total = (int)cand.size();
    return SB_OK;
}

static int slot_of(const sb_loop* L, int entry) {
    // owned entries are stored in ascending id order: entry = rank + slot * world
    int slot = (entry - L->rank) / L->world;
    if (entry % L->world != L->rank || slot < 0 || slot >= (int)L->entry_id.size() || L->entry_id[slot] != entry) return -1;
    return slot;
}

// loop_closure.hpp:99-109 for a set of entries at once
int loop_verify(sb_loop* L, const int* entries, const double* dist, int n, sb_loop_result* results, int* converged) {
    Ctx* ctx = L->ctx;
    if (n <= 0) return SB_OK;
    int qslot = (int)L->entry_id.size() - 1;
    if (qslot < 0) return fail(ctx, SB_ERR_EMPTY, "loop: empty database");
    std::vector<int> slots((size_t)n);
    for (int i = 0; i < n; ++i) {
        slots[i] = slot_of(L, entries[i]);
        if (slots[i] < 0 || slots[i] == qslot)
            return fail(ctx, SB_ERR_INVALID_ARG, "loop: entry %d is not owned by rank %d", entries[i], L->rank);
    }
    Forest F;
    F.in_arena = true;
    int s = forest_reserve(ctx, &F, n + 1);
    if (s == SB_OK) s = forest_append(ctx, &F, L->d_clouds, L->cloud_off.data(), slots.data(), n);   // candidates
    if (s == SB_OK) s = forest_normals(ctx, &F, L->cfg.normals_k, nullptr, nullptr);
    if (s == SB_OK) s = forest_append(ctx, &F, L->d_clouds, L->cloud_off.data(), &qslot, 1);         // the query
    std::vector<sb_icp_result> res((size_t)n);
    if (s == SB_OK) {
        sb_icp_config cfg;
        sb_default_icp_config(&cfg);
        cfg.max_iterations = L->cfg.icp_max_iterations;  // loop_closure.hpp:106
        cfg.tolerance = L->cfg.icp_tolerance;            // loop_closure.hpp:107
        cfg.normals_k = L->cfg.normals_k;
        std::vector<PairDesc> pairs((size_t)n);
        for (int i = 0; i < n; ++i) {
            memset(&pairs[i], 0, sizeof(PairDesc));
            pairs[i].src_tree = n;                   // source = query cloud (loop_closure.hpp:102)
            pairs[i].n_src = (int)(L->cloud_off[qslot + 1] - L->cloud_off[qslot]);
            pairs[i].tree = i;                       // target = candidate cloud (loop_closure.hpp:103)
        }
        s = icp_batch(ctx, &F, pairs, &cfg, res.data());
    }
    const cudaError_t sync_err = cudaStreamSynchronize(ctx->stream);
    forest_free(&F);
    if (s != SB_OK) return s;
    if (sync_err != cudaSuccess) return fail(ctx, SB_ERR_CUDA, "loop: verification failed on the device: %s", cudaGetErrorString(sync_err));
    for (int i = 0; i < n; ++i) {
        sb_loop_result& R = results[i];
        R.query_frame = L->last_frame;
        R.match_frame = L->frame_idx[slots[i]];
        memcpy(R.transform, res[i].transformation, sizeof(R.transform));
        R.scan_context_distance = dist ? dist[i] : 0.0;
        R.icp_fitness = res[i].final_error;
        converged[i] = res[i].converged && res[i].status == SB_OK;
    }
    return SB_OK;
}

}  // namespace sb
