// selfknn.cuh — k_self_knn: k-NN of the trees' own points + surface normals, one LANE per query.
// (included by forest.cu inside namespace sb, after normal_of_point / find_segment / load_tree)
//
// Replaces KDTree::k_nearest called once per point by estimate_normals (slam_viz/include/slam_viz/core/icp.hpp:23-67,
// kdtree.hpp:65-78, 144-180) for the k the pipelines use.  k_knn<1> answers one query per warp: every step of the
// descent, every list insertion and the seed ordering are warp-uniform bookkeeping paid once per QUERY (1 742
// warp-instructions per query, profiles/r01_ncu_knn_sass_segments_final.txt).  Here the 32 queries of a leaf (a
// "packet") share ONE nearest-first traversal — the query is the leaf's box — and everything that is per query
// runs in that query's lane on registers with compile-time indices, so the 32 lanes execute the same instructions
// whatever their data:
//
//   scan    each visited leaf's 32 points are read as LEAF-LOCAL float32 offsets (TreeDesc::pts32: one broadcast
//           16-byte load per candidate).  A lane computes the float32 squared distance d~ and, from the error
//           analysis below, an interval [lo, hi] that contains the exact d2.  It keeps the candidate iff lo <= U,
//           U = the lane's current bound on its k-th nearest distance, by appending (hi, position) to its column of
//           a shared-memory buffer.
//   flush   the buffered candidates are merged, NB at a time, into the lane's sorted list of k+1 entries (key = hi)
//           with a sorting network + a bitonic merge on registers.  U = key of entry k-1: k points are no farther.
//   verify  at the end the lane evaluates the EXACT fp64 distances D_0 .. D_{k-1} of its list (the oracle's
//           (dx*dx + dy*dy) + dz*dz without FMA) and accepts the list iff
//             (i)  D_0 < D_1 < ... < D_{k-1}            (the float32 order is the exact order, no exact ties), and
//             (ii) D_{k-1} < V (1 - 2 rho) - 2 beta_max  (V = key of entry k, the best candidate left out):
//           every candidate that was ever kept but is not in the list has hi >= V, hence exact d2 >= lo > D_{k-1};
//           every candidate dropped by the scan had lo > U(then) >= U(final) >= D_{k-1}; every leaf the traversal
//           skipped has a box farther than max over lanes of U.  So the list IS the exact k nearest in the oracle's
//           (d2, index) order.  A query that fails (i) or (ii) — exact ties, or a gap below the float32 resolution,
//           ~1e-2 of the queries — is appended to a redo list and answered by the warp-per-query search of
//           traverse.cuh in a second kernel (k_knn_redo).
// The normals are a third kernel (k_normals_from_graph) over the neighbour lists this one writes: the three pieces
// of code are each far smaller than their sum, and instruction fetch — not issue slots — was what bound the fused
// version (ncu: 25 % issue-active, 7.9 cycles of `no_instruction` stall per issue with 94 KB of code).
//
// Error analysis ([lo, hi] = [d~ (1 - rho) - beta, d~ (1 + rho) + beta]): traverse.cuh, "Leaf-local float32 arithmetic".
#pragma once

static constexpr int PWARPS = 4;   // warps per CTA

// SB_KNN_STATS=1 (debugging aid): per-launch totals printed by forest_normals
enum { PS_PACKETS, PS_LEAVES, PS_SCANNED, PS_CAND, PS_APPENDED, PS_FLUSHES, PS_ROUNDS, PS_MERGES, PS_REDO, PS_N };

// Shared memory through 32-bit shared-window addresses: with a generic pointer kept in a struct the compiler
// re-derives the window base (S2R SR_CgaCtaId ...) at every access — 9 of the 22 instructions per scanned candidate.
__device__ __forceinline__ void sts64(unsigned addr, int a, int b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b));
}
__device__ __forceinline__ void sts32(unsigned addr, int a) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a)); }
__device__ __forceinline__ int2 lds64(unsigned addr) {
    int2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds32(unsigned addr) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ void ce_asc(float& a, int& pa, float& b, int& pb) {  // afterwards a <= b (no NaNs)
    const bool sw = b < a;
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    const int qa = sw ? pb : pa, qb = sw ? pa : pb;
    a = lo; b = hi; pa = qa; pb = qb;
}

// Bitonic sorting network for ANY length (H. W. Lang's construction), written as template recursion so that every
// index is a compile-time constant and the arrays live in registers.  ASC: ascending result.
template <bool ASC>
__device__ __forceinline__ void ce_dir(float& a, int& pa, float& b, int& pb) {
    if (ASC) ce_asc(a, pa, b, pb); else ce_asc(b, pb, a, pa);
}
__host__ __device__ constexpr int pow2_below(int n) { int k = 1; while (k * 2 < n) k *= 2; return k; }   // largest power of two < n
// x[LO .. LO+N) is bitonic -> sorted
template <int LO, int N, bool ASC, int TOT>
__device__ __forceinline__ void bitonic_merge(float (&xk)[TOT], int (&xp)[TOT]) {
    if constexpr (N > 1) {
        constexpr int M = pow2_below(N);
#pragma unroll
        for (int i = LO; i < LO + N - M; ++i) ce_dir<ASC>(xk[i], xp[i], xk[i + M], xp[i + M]);
        bitonic_merge<LO, M, ASC, TOT>(xk, xp);
        bitonic_merge<LO + M, N - M, ASC, TOT>(xk, xp);
    }
}
template <int LO, int N, bool ASC, int TOT>
__device__ __forceinline__ void sort_network(float (&xk)[TOT], int (&xp)[TOT]) {
    if constexpr (N > 1) {
        constexpr int M = N / 2;
        sort_network<LO, M, !ASC, TOT>(xk, xp);
        sort_network<LO + M, N - M, ASC, TOT>(xk, xp);
        bitonic_merge<LO, N, ASC, TOT>(xk, xp);
    }
}

template <int K, int TOT, int PCAP, bool STATS>
struct PacketVisitor {
    static constexpr int KL = K + 1, NB = TOT - KL;
    static_assert(NB >= 1 && PCAP >= 40, "a leaf adds up to 32 candidates per lane between two flushes");
    const TreeDesc& T;
    const int lane;
    const int own_p0;
    const unsigned buf;        // shared-window address of this lane's column: entry `slot` at buf + slot * 256 = (key bits, position)
    const unsigned stage;      // shared-window address of the warp's 32 x float4 staging row (SB_LEAF_STAGE)
    double qx, qy, qz;         // this lane's query (NaN for lanes without one: nothing is ever kept)
    float xk[TOT];             // [0, KL): the list's keys, ascending; [KL, TOT): the batch being merged
    int xp[TOT];               // positions (cloud-local, sorted order)
    float U;                   // key of entry K-1: an upper bound of the K-th nearest d2 (+inf until K points are known)
    float U_warp;              // max over lanes
    float beta_max;            // largest beta of any leaf this lane KEPT a candidate of
    int cnt;                   // buffered candidates of this lane
    unsigned st[STATS ? PS_N : 1];
    __device__ __forceinline__ PacketVisitor(const TreeDesc& t, int l, int p0, unsigned b, unsigned st_)
        : T(t), lane(l), own_p0(p0), buf(b), stage(st_) {}
    __device__ __forceinline__ double tau() const { return (double)U_warp; }
    __device__ __forceinline__ void refresh_bounds() {
        U = xk[K - 1];
        U_warp = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(U)));  // non-negative floats
    }
    // leaf-local query offsets and the error term beta of this (lane, leaf): traverse.cuh
    __device__ __forceinline__ void leaf_frame(float2 b0, float2 b1, float2 b2, float& ox, float& oy, float& oz,
                                               float& beta) {
        leaf_local_frame(qx, qy, qz, b0, b1, b2, ox, oy, oz, beta);
    }
    __device__ __forceinline__ void init(bool valid, double x, double y, double z) {
        const double nan_ = __longlong_as_double(0x7ff8000000000000LL);
        const float inf = __int_as_float(0x7f800000);
        qx = valid ? x : nan_; qy = valid ? y : nan_; qz = valid ? z : nan_;
        beta_max = 0.f;
        cnt = 0;
        if (STATS) for (int i = 0; i < PS_N; ++i) st[i] = 0;
#pragma unroll
        for (int i = 0; i < TOT; ++i) { xk[i] = inf; xp[i] = -1; }
        U = inf;
        U_warp = inf;
    }
    // candidates [p0, p0 + n) of one leaf
    __device__ __forceinline__ void scan_leaf(int p0, int n, float ox, float oy, float oz, float beta) {
        const float4* __restrict__ c = T.pts32 + T.pt_off + p0;
        unsigned w = buf + (unsigned)cnt * 256u;
        const unsigned w0 = w;
        const float thr = U;
#if SB_LEAF_STAGE
        // the leaf's 32 candidates: ONE coalesced 512-byte load, lane i fetching candidate i, handed round through
        // shared memory — as 32 broadcast loads the four lines of a leaf missed one after the other (14 % of the
        // kernel's stall samples sat on the first use of a candidate)
        __syncwarp();
        if (lane < n) sts128(stage + 16u * lane, __ldg(c + lane));
        __syncwarp();
#endif
#pragma unroll 8
        for (int i = 0; i < n; ++i) {
#if SB_LEAF_STAGE
            const float4 P = lds128(stage + 16u * i);
#else
            const float4 P = __ldg(c + i);  // same address in every lane: one broadcast load
#endif
            const float dx = P.x - ox, dy = P.y - oy, dz = P.z - oz;
            const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
            const float lo = __fmaf_rd(d, SB_RHO_DN, -beta);
            if (lo <= thr) {   // false for NaN
                sts64(w, __float_as_int(__fmaf_ru(d, SB_RHO_UP, beta)), p0 + i);
                w += 256u;
            }
        }
        const int added = (int)((w - w0) >> 8);
        cnt += added;
        if (added) beta_max = fmaxf(beta_max, beta);   // only candidates that were kept enter condition (ii)
        if (STATS) { st[PS_SCANNED]++; st[PS_CAND] += n; st[PS_APPENDED] += added; }
    }
    // merges the buffered candidates into the sorted list, NB at a time
    __device__ __forceinline__ void flush() {
        const float inf = __int_as_float(0x7f800000);
        if (STATS) st[PS_FLUSHES]++;
        while (__any_sync(0xffffffffu, cnt > 0)) {
            bool fresh = false;
            if (STATS) st[PS_ROUNDS]++;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int slot = cnt - 1 - b;
                float key = inf;
                int p = -1;
                if (slot >= 0) {
                    const int2 e = lds64(buf + (unsigned)slot * 256u);
                    // a candidate can only enter the K+1 list if its key is below the list's last key
                    if (__int_as_float(e.x) < xk[KL - 1]) { key = __int_as_float(e.x); p = e.y; fresh = true; }
                }
                xk[KL + b] = key;
                xp[KL + b] = p;
            }
            cnt = cnt > NB ? cnt - NB : 0;
            if (!__any_sync(0xffffffffu, fresh)) continue;
            if (STATS) st[PS_MERGES]++;
            sort_network<KL, NB, false, TOT>(xk, xp);      // the batch, descending: list + batch is bitonic
            bitonic_merge<0, TOT, true, TOT>(xk, xp);
        }
        refresh_bounds();
    }
    // the packet's own leaf: every point is a candidate of every lane (U = +inf: everything but NaN is kept)
    __device__ __forceinline__ void scan_own(int n) {
        const float2* b = reinterpret_cast<const float2*>(T.boxes + 6 * (T.box_off[0] + (own_p0 >> 5)));
        const float2 b0 = __ldg(b), b1 = __ldg(b + 1), b2 = __ldg(b + 2);
        float ox, oy, oz, beta;
        leaf_frame(b0, b1, b2, ox, oy, oz, beta);
        scan_leaf(own_p0, n, ox, oy, oz, beta);
    }
    // returns true when the buffer may not take another leaf: flush before going on
    __device__ __forceinline__ bool leaf(int p0, int n) {
        if (p0 == own_p0) return false;  // the packet's own leaf was taken first
        if (STATS) st[PS_LEAVES]++;
        // does ANY lane need this leaf?  (the traversal only knows that the packet's box might)
        const float2* b = reinterpret_cast<const float2*>(T.boxes + 6 * (T.box_off[0] + (p0 >> 5)));
        const float2 b0 = __ldg(b), b1 = __ldg(b + 1), b2 = __ldg(b + 2);
        PointQuery Q;
        Q.x = qx; Q.y = qy; Q.z = qz;
        const float lb = Q.lb(b0, b1, b2);
        if (!__any_sync(0xffffffffu, lb <= U)) return false;
        float ox, oy, oz, beta;
        leaf_frame(b0, b1, b2, ox, oy, oz, beta);
        scan_leaf(p0, n, ox, oy, oz, beta);
        return __any_sync(0xffffffffu, cnt > PCAP - 32);
    }
};

// traverse() of traverse.cuh made resumable: runs until the visitor asks for a flush (returns true; call again
// afterwards with the same `level`) or the tree is exhausted (returns false).
template <class Query, class Visitor>
__device__ __forceinline__ bool traverse_until_full(const ForestView& F, const TreeDesc& T, const Query& Q, WarpStack& S,
                                                    Visitor& V, int lane, int& level) {
    while (true) {
        unsigned m = S.mask[level];
        if (m == 0u) {
            if (level == T.top) return false;
            ++level;
            continue;
        }
        bool mine = (m >> lane) & 1u;
        unsigned key = mine ? __float_as_uint(S.dm[level][lane]) : 0xffffffffu;
        unsigned kmin = __reduce_min_sync(0xffffffffu, key);
        if ((double)__uint_as_float(kmin) > V.tau()) {
            S.mask[level] = 0u;
            continue;
        }
        int c = __ffs(__ballot_sync(0xffffffffu, mine && key == kmin)) - 1;
        S.mask[level] = m & ~(1u << c);
        int node = S.base[level] + c;
        if (level == 0) {
            int p0 = node * 32;
            int cnt = T.n - p0;
            if (V.leaf(p0, cnt > 32 ? 32 : cnt)) return true;
        } else {
            --level;
            test_children(F, T, level, node * 32, Q, V.tau(), S, lane);
        }
    }
}

// One entry of the redo list: a query the packet search could not settle.
struct RedoEntry {
    int tree;   // tree id relative to the ForestView
    int pos;    // cloud-local sorted position of the query point
};

template <int K, int TOT, int PCAP, bool STATS>
__global__ void __launch_bounds__(PWARPS * 32, 4) k_self_knn(ForestView F, const i64* __restrict__ tio, i64 n_items,
                                                          int n_trees, NbrEntry* __restrict__ nbr_sorted,
                                                          RedoEntry* __restrict__ redo_list, int* __restrict__ redo_count,
                                                          unsigned long long* __restrict__ stats, int* __restrict__ work) {
    constexpr int ROW = K | 1;   // int2 entries per row of the output staging: odd, so that rows start on different banks
    constexpr int ENTRIES = (32 * ROW > PCAP * 32) ? 32 * ROW : PCAP * 32;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ WarpStack stacks[PWARPS];
    __shared__ TreeDesc s_tree[PWARPS];
#if SB_LEAF_STAGE
    __shared__ __align__(16) float4 s_stage[PWARPS][32];
#endif
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // candidate buffer during the search (entry `slot` of lane l at + (slot * 32 + l) * 8), then the packet's result
    // rows (entry j of query l at + (l * ROW + j) * 8); always addressed through the shared window (sts64 / lds64)
    const unsigned rows = (unsigned)__cvta_generic_to_shared(s_dyn) + (unsigned)warp * (unsigned)(ENTRIES * 8);
    WarpStack& S = stacks[warp];
    // Packets are handed out by a device-wide counter (work != nullptr) to a grid of resident CTAs: a packet costs
    // anything between a few and a few dozen leaves, and with a fixed share per warp the CTA slots of the early
    // finishers idled (ncu: 13 of 16 warps per SM active on average).
    auto next_item = [&](i64 prev) -> i64 {
        if (!work) return prev < 0 ? (i64)blockIdx.x * PWARPS + warp : prev + (i64)gridDim.x * PWARPS;
        int v = 0;
        if (lane == 0) v = atomicAdd(work, 1);
        return (i64)__shfl_sync(0xffffffffu, v, 0);
    };
    for (i64 it = next_item(-1); it < n_items; it = next_item(it)) {
        const int t = find_segment(tio, n_trees, it);
        const int q_off = (int)(it - tio[t]) * 32;
        __syncwarp();
        load_tree(&s_tree[warp], &F.trees[t], lane);
        const TreeDesc& T = s_tree[warp];
        const int count = T.n - q_off < 32 ? T.n - q_off : 32;
        double mx = 0, my = 0, mz = 0;
        if (lane < count) {
            TreePoint P = load_point(T.pts + T.pt_off + q_off + lane);
            mx = P.x; my = P.y; mz = P.z;
        }
        bool ok = false;
        if (T.gext < 1.0e15) {   // finite, moderate extent (false for NaN too): the float32 error bounds cannot overflow
#if SB_LEAF_STAGE
            const unsigned stage_addr = (unsigned)__cvta_generic_to_shared(&s_stage[warp][0]);
#else
            const unsigned stage_addr = 0u;
#endif
            PacketVisitor<K, TOT, PCAP, STATS> V(T, lane, q_off, rows + (unsigned)lane * 8u, stage_addr);
            V.init(lane < count, mx, my, mz);
            V.scan_own(count);
            const float2* b = reinterpret_cast<const float2*>(T.boxes + 6 * (T.box_off[0] + (q_off >> 5)));
            const float2 b0 = __ldg(b), b1 = __ldg(b + 1), b2 = __ldg(b + 2);
            BoxQuery Q;
            Q.lox = b0.x; Q.loy = b0.y; Q.loz = b1.x; Q.hix = b1.y; Q.hiy = b2.x; Q.hiz = b2.y;
            int level = T.top;
            bool more = true, started = false;
            while (true) {   // one flush site: the merge networks are several hundred instructions
                V.flush();
                if (!more) break;
                if (!started) { test_children(F, T, level, 0, Q, V.tau(), S, lane); started = true; }
                more = traverse_until_full(F, T, Q, S, V, lane, level);
            }
            __syncwarp();   // the candidate buffer is dead from here on: its storage takes the result rows
            // verify (see the header comment): exact distances of the list in list order
            const TreePoint* __restrict__ TP = T.pts + T.pt_off;
            ok = true;
            double prev = -1.0;
            int m = 0;
#pragma unroll
            const unsigned my_row = rows + (unsigned)(lane * ROW) * 8u;
#pragma unroll
            for (int j = 0; j < K; ++j) sts64(my_row + 8u * j, V.xp[j], 0x7f800000);
#ifndef SB_VERIFY_PIPE
#define SB_VERIFY_PIPE 1
#endif
#if SB_VERIFY_PIPE
            // rolled (code size), software-pipelined by one: the gather of entry j + 1 is in flight while entry j is checked
            int p_nx = lds32(my_row);
            TreePoint P_nx = load_point(TP + (p_nx >= 0 ? p_nx : 0));
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
                const int p = p_nx;
                const TreePoint P = P_nx;
                if (j + 1 < K) {
                    p_nx = lds32(my_row + 8u * (j + 1));
                    P_nx = load_point(TP + (p_nx >= 0 ? p_nx : 0));
                }
                if (p >= 0) {
                    const double D = dist2_rn(P.x, P.y, P.z, V.qx, V.qy, V.qz);
                    ok = ok && (D > prev);
                    prev = D;
                    sts32(my_row + 8u * j + 4u, __float_as_int(sqrt_lower(D)));
                    ++m;
                }
            }
#else
#pragma unroll 1
            for (int j = 0; j < K; ++j) {   // rolled on purpose (code size): the positions come back from shared memory
                const int p = lds32(my_row + 8u * j);
                if (p >= 0) {
                    const TreePoint P = load_point(TP + p);
                    const double D = dist2_rn(P.x, P.y, P.z, V.qx, V.qy, V.qz);
                    ok = ok && (D > prev);
                    prev = D;
                    sts32(my_row + 8u * j + 4u, __float_as_int(sqrt_lower(D)));
                    ++m;
                }
            }
#endif
            // (ii): the best candidate left out is provably farther than the list's last point
            const float v_lo = __fmaf_rd(V.xk[K], SB_RHO2_DN, -2.0f * V.beta_max);
            ok = ok && (m < K || __double2float_ru(prev) < v_lo * 0.99999988f);
            if (STATS) {
                V.st[PS_PACKETS] = 1; V.st[PS_REDO] = __popc(__ballot_sync(0xffffffffu, !ok && lane < count));
                for (int i = 0; i < PS_N; ++i) {
                    unsigned v = V.st[i];
                    if (i == PS_APPENDED) v = __reduce_add_sync(0xffffffffu, v);
                    if (lane == 0 && stats) atomicAdd(&stats[i], (unsigned long long)v);
                }
            }
        }
        // queries with exactly tied distances, an unresolved float32 gap, or a tree of extreme extent: k_knn_redo
        const unsigned redo = __ballot_sync(0xffffffffu, !ok && lane < count);
        if (redo) {
            int base = 0;
            if (lane == 0) base = atomicAdd(redo_count, __popc(redo));
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((redo >> lane) & 1u) {
                RedoEntry e;
                e.tree = t; e.pos = q_off + lane;
                redo_list[base + __popc(redo & lanemask_lt())] = e;
            }
        }
        __syncwarp();
        // the neighbour graph (NbrEntry rows of the packet are contiguous: coalesced copy); rows of redo queries are
        // rewritten by k_knn_redo
        if (T.gext < 1.0e15) {
            int2* out = reinterpret_cast<int2*>(nbr_sorted + (T.pt_off + q_off) * (i64)K);
            for (int e = lane; e < count * K; e += 32) out[e] = lds64(rows + (unsigned)((e / K) * ROW + (e % K)) * 8u);
        }
        __syncwarp();
    }
}

// The queries k_self_knn left open: one warp per query, the exact search of traverse.cuh in the oracle's (d2, index)
// order; rewrites the query's row of the neighbour graph.
__global__ void __launch_bounds__(QWARPS * 32) k_knn_redo(ForestView F, const RedoEntry* __restrict__ redo_list,
                                                          const int* __restrict__ redo_count, int k,
                                                          NbrEntry* __restrict__ nbr_sorted, int* __restrict__ work) {
    __shared__ WarpStack stacks[QWARPS];
    __shared__ TreeDesc s_tree[QWARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpStack& S = stacks[warp];
    const int n = *redo_count;
    for (i64 e = next_work(work, -1, lane, warp, QWARPS); e < n; e = next_work(work, e, lane, warp, QWARPS)) {
        const RedoEntry E = redo_list[e];
        __syncwarp();
        load_tree(&s_tree[warp], &F.trees[E.tree], lane);
        const TreeDesc& T = s_tree[warp];
        const TreePoint P = load_point(T.pts + T.pt_off + E.pos);
        KnnVisitor V(F, T, P.x, P.y, P.z, lane, k);
        traverse(F, T, P.x, P.y, P.z, S, V, lane);
        if (lane < k) {
            const bool have = V.lidx != 0x7fffffff;
            NbrEntry o;
            o.pos = have ? V.lpos : -1;
            o.r = have ? sqrt_lower(V.ld) : __int_as_float(0x7f800000);
            nbr_sorted[(T.pt_off + E.pos) * (i64)k + lane] = o;
        }
    }
}

// Normals from the neighbour graph (icp.hpp:34-63): one thread per sorted point, one warp per leaf (implicit items
// like k_self_knn).  Also accumulates the mean distance to the nearest other point (sizes the seed grid).
__global__ void __launch_bounds__(256) k_normals_from_graph(ForestView F, const i64* __restrict__ tio, i64 n_items,
                                                            int n_trees, int k, const NbrEntry* __restrict__ nbr_sorted,
                                                            TreePoint* __restrict__ pts, TreeNormal* __restrict__ nrm_sorted,
                                                            double* __restrict__ nrm_orig, double* __restrict__ evals_orig,
                                                            unsigned long long* __restrict__ spacing_acc) {
    const int lane = threadIdx.x & 31;
    const i64 it = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (it >= n_items) return;
    const int t = find_segment(tio, n_trees, it);
    const TreeDesc& T = F.trees[t];
    const int pos = (int)(it - tio[t]) * 32 + lane;
    unsigned long long spacing = 0ull;
    if (pos < T.n) {
        const int* row = reinterpret_cast<const int*>(nbr_sorted + (T.pt_off + pos) * (i64)k);
        int m = k;   // valid entries come first: all k of them unless the cloud has fewer points (one look at the last)
        if (row[2 * (k - 1)] < 0) {
            m = 0;
            while (m < k && row[2 * m] >= 0) ++m;
        }
        if (k > 1 && m > 1) spacing = (unsigned long long)(fminf(__int_as_float(row[3]), 1.0e6f) * 65536.0f);
        // the nearest-other-point bound into the point's own sector (TreePoint::pad); with k == 1 or a single point
        // nothing is known: 0
        pts[T.pt_off + pos].pad = (k > 1 && m > 1) ? row[3] : 0;
        normal_of_point(T, row, 2, m, T.pt_off + pos, nrm_sorted, nrm_orig, evals_orig);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) spacing += __shfl_xor_sync(0xffffffffu, spacing, o);
    if (lane == 0 && spacing) atomicAdd(&spacing_acc[t], spacing);  // integer: order-independent
}
