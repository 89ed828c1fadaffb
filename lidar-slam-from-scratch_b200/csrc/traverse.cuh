// traverse.cuh — warp-cooperative exact nearest-neighbour traversal of the forest (forest.cu, icp.cu).
//
// Replaces the recursive KD-tree searches of the reference (slam_viz/include/slam_viz/core/kdtree.hpp:112-142
// search_nearest and :144-180 search_k_nearest).  The index of one cloud is an implicit 32-ary tree of axis-aligned
// boxes over the curve-sorted (Hilbert) points: level-0 box b bounds sorted points [32b, 32b+32), level-l box b bounds the
// level-(l-1) boxes [32b, 32b+32).  One warp answers one query: the 32 lanes test the 32 children of a node (one
// coalesced 768-byte load), descend nearest-box-first, and evaluate the 32 points of a leaf with one coalesced
// 32-byte load per lane (TreePoint).  A subtree is skipped only if a conservative lower bound of its squared distance is
// strictly greater than the current k-th best, so the result is exact for any density and any radius.
//
// Distances are the oracle's (dx*dx + dy*dy) + dz*dz in fp64 without FMA; candidates are ranked by
// (d2, original index) lexicographically.
#pragma once
#include "common.cuh"

namespace sb {

struct WarpStack {                       // shared memory, one per warp
    float dm[SB_MAX_LEVELS][32];         // lower-bound d^2 of "my" child at each level
    unsigned mask[SB_MAX_LEVELS];        // children still to visit at each level (warp-uniform)
    int base[SB_MAX_LEVELS];             // first child index at each level (warp-uniform)
};

struct ForestView {
    const TreeDesc* __restrict__ trees;
};

// ---- seed grid (GridSlot, common.cuh): cell coordinates are offset by +1 so that the cells just outside the
// bounding box are representable; 21 bits per axis.
#define SB_GRID_EMPTY 0xffffffffffffffffull
__device__ __forceinline__ bool grid_cell(const TreeDesc& T, double x, double y, double z, int& ix, int& iy, int& iz) {
    double fx = floor((x - T.glo[0]) * T.ginv), fy = floor((y - T.glo[1]) * T.ginv), fz = floor((z - T.glo[2]) * T.ginv);
    if (!(fx >= -1.0 && fx < 2097150.0 && fy >= -1.0 && fy < 2097150.0 && fz >= -1.0 && fz < 2097150.0)) return false;
    ix = (int)fx + 1; iy = (int)fy + 1; iz = (int)fz + 1;
    return true;
}
__device__ __forceinline__ unsigned long long grid_key(int ix, int iy, int iz) {
    return (unsigned long long)ix | ((unsigned long long)iy << 21) | ((unsigned long long)iz << 42);
}
__device__ __forceinline__ unsigned grid_hash(unsigned long long key, int shift) {
    return (unsigned)((key * 0x9E3779B97F4A7C15ull) >> shift);
}

// one 32-byte sector as two 16-byte read-only loads
__device__ __forceinline__ TreePoint load_point(const TreePoint* p) {
    const int4* q = reinterpret_cast<const int4*>(p);
    int4 a = __ldg(q), b = __ldg(q + 1);
    TreePoint r;
    r.x = __hiloint2double(a.y, a.x);
    r.y = __hiloint2double(a.w, a.z);
    r.z = __hiloint2double(b.y, b.x);
    r.idx = b.z;
    r.pad = b.w;
    return r;
}

__device__ __forceinline__ bool lex_less(double da, int ia, double db, int ib) {
    return da < db || (da == db && ia < ib);
}

// What is searched for: a point, or (self k-NN of a whole leaf at once) the bounding box of a packet of queries.
// lb() is a conservative lower bound of the squared distance from the query to a box given as 3 float2
// (lo.x lo.y | lo.z hi.x | hi.y hi.z; lo rounded down, hi rounded up): a few ulps shaved off, then rounded towards
// -inf into a float.
struct PointQuery {
    double x, y, z;
    __device__ __forceinline__ float lb(float2 b0, float2 b1, float2 b2) const {
        double ex = fmax(fmax((double)b0.x - x, x - (double)b1.y), 0.0);
        double ey = fmax(fmax((double)b0.y - y, y - (double)b2.x), 0.0);
        double ez = fmax(fmax((double)b1.x - z, z - (double)b2.y), 0.0);
        double d = (ex * ex + ey * ey) + ez * ez;
        return __double2float_rd(d * (1.0 - 1.0e-14));
    }
};
struct BoxQuery {
    double lox, loy, loz, hix, hiy, hiz;
    __device__ __forceinline__ float lb(float2 b0, float2 b1, float2 b2) const {
        double ex = fmax(fmax((double)b0.x - hix, lox - (double)b1.y), 0.0);
        double ey = fmax(fmax((double)b0.y - hiy, loy - (double)b2.x), 0.0);
        double ez = fmax(fmax((double)b1.x - hiz, loz - (double)b2.y), 0.0);
        double d = (ex * ex + ey * ey) + ez * ez;
        return __double2float_rd(d * (1.0 - 1.0e-14));
    }
};

// Tests the (up to 32) boxes [first, first+32) of `level` against the query; records the survivors.
template <class Query>
__device__ __forceinline__ void test_children(const ForestView& F, const TreeDesc& T, int level, int first,
                                              const Query& Q, double tau, WarpStack& S, int lane) {
    int ci = first + lane;
    bool valid = ci < T.box_cnt[level];
    float dmf = __int_as_float(0x7f800000);
    if (valid) {
        const float2* b = reinterpret_cast<const float2*>(T.boxes + 6 * (T.box_off[level] + ci));
        dmf = Q.lb(b[0], b[1], b[2]);
    }
    bool pass = valid && !((double)dmf > tau);
    unsigned m = __ballot_sync(0xffffffffu, pass);
    S.dm[level][lane] = dmf;
    S.mask[level] = m;   // every lane writes the same value; each lane only ever reads back what the warp wrote
    S.base[level] = first;
}

// Generic nearest-first traversal.  V must provide:
//   double tau() const            — current pruning bound (k-th best d2; +inf/DBL_MAX while the list is not full)
//   void leaf(int p0, int cnt)    — visit sorted points [p0, p0+cnt) (cloud-local positions)
template <class Query, class Visitor>
__device__ __forceinline__ void traverse(const ForestView& F, const TreeDesc& T, const Query& Q, WarpStack& S,
                                         Visitor& V, int lane) {
    if (T.n <= 0) return;
    int level = T.top;
    test_children(F, T, level, 0, Q, V.tau(), S, lane);
    while (true) {
        unsigned m = S.mask[level];
        if (m == 0u) {
            if (level == T.top) break;
            ++level;
            continue;
        }
        bool mine = (m >> lane) & 1u;
        unsigned key = mine ? __float_as_uint(S.dm[level][lane]) : 0xffffffffu;  // non-negative floats order as uints
        unsigned kmin = __reduce_min_sync(0xffffffffu, key);
        if ((double)__uint_as_float(kmin) > V.tau()) {  // the nearest remaining child is too far: so are the rest
            S.mask[level] = 0u;
            continue;
        }
        int c = __ffs(__ballot_sync(0xffffffffu, mine && key == kmin)) - 1;
        S.mask[level] = m & ~(1u << c);
        int node = S.base[level] + c;
        if (level == 0) {
            int p0 = node * 32;
            int cnt = T.n - p0;
            V.leaf(p0, cnt > 32 ? 32 : cnt);
        } else {
            --level;
            test_children(F, T, level, node * 32, Q, V.tau(), S, lane);
        }
    }
}

template <class Visitor>
__device__ __forceinline__ void traverse(const ForestView& F, const TreeDesc& T, double qx, double qy, double qz,
                                         WarpStack& S, Visitor& V, int lane) {
    PointQuery Q;
    Q.x = qx; Q.y = qy; Q.z = qz;
    traverse(F, T, Q, S, V, lane);
}

// -------------------------------------------------------------------------------------------------------------
// 1-NN visitor
// -------------------------------------------------------------------------------------------------------------
struct NearestVisitor {
    const ForestView& F;
    const TreeDesc& T;
    double qx, qy, qz;
    int lane;
    double best_d;   // warp-uniform
    int best_idx;    // original (cloud-local) row, -1 if none
    int best_pos;    // cloud-local sorted position
    __device__ NearestVisitor(const ForestView& f, const TreeDesc& t, double x, double y, double z, int l)
        : F(f), T(t), qx(x), qy(y), qz(z), lane(l), best_d(1.7976931348623157e308), best_idx(-1), best_pos(-1) {}
    __device__ __forceinline__ double tau() const { return best_d; }
    // Starts the search from a known tree point (cloud-local sorted position): any real point bounds the nearest
    // distance from above, so the traversal only has to visit boxes inside that ball.  Exactness is unaffected.
    __device__ __forceinline__ void seed(int pos) {
        if (pos < 0 || pos >= T.n) return;
        TreePoint P = load_point(T.pts + T.pt_off + pos);
        double d = dist2_rn(P.x, P.y, P.z, qx, qy, qz);
        int idx = P.idx;
        if (d < best_d || (d == best_d && idx < best_idx)) {
            best_d = d;
            best_idx = idx;
            best_pos = pos;
        }
    }
    __device__ __forceinline__ void leaf(int p0, int cnt) {
        bool valid = lane < cnt;
        unsigned hi = 0xffffffffu, lo = 0xffffffffu;
        int idx = 0x7fffffff;
        if (valid) {
            TreePoint P = load_point(T.pts + T.pt_off + p0 + lane);
            double d = dist2_rn(P.x, P.y, P.z, qx, qy, qz);
            idx = P.idx;
            if (d == d) {  // NaN never wins (kdtree.hpp:125 strict <)
                long long b = __double_as_longlong(d);
                hi = (unsigned)((unsigned long long)b >> 32);
                lo = (unsigned)b;
            }
        }
        unsigned mh = __reduce_min_sync(0xffffffffu, hi);
        if (mh == 0xffffffffu) return;
        unsigned lo2 = hi == mh ? lo : 0xffffffffu;
        unsigned ml = __reduce_min_sync(0xffffffffu, lo2);
        bool tie = hi == mh && lo == ml;
        unsigned mi = __reduce_min_sync(0xffffffffu, tie ? (unsigned)idx : 0xffffffffu);
        int src = __ffs(__ballot_sync(0xffffffffu, tie && (unsigned)idx == mi)) - 1;
        double d = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
        if (d < best_d || (d == best_d && (int)mi < best_idx)) {  // kdtree.hpp:125 + canonical tie rule
            best_d = d;
            best_idx = (int)mi;
            best_pos = p0 + src;
        }
    }
};

// -------------------------------------------------------------------------------------------------------------
// k-NN visitor (k <= 32): lane j holds the j-th best (d2, idx, pos), ascending
// -------------------------------------------------------------------------------------------------------------
struct KnnVisitor {
    const ForestView& F;
    const TreeDesc& T;
    double qx, qy, qz;
    int lane, k;
    double ld;       // this lane's entry
    int lidx, lpos;
    double tau_d;    // entry k-1 (warp-uniform)
    int tau_idx;
    __device__ KnnVisitor(const ForestView& f, const TreeDesc& t, double x, double y, double z, int l, int kk)
        : F(f), T(t), qx(x), qy(y), qz(z), lane(l), k(kk), ld((double)INFINITY),
          lidx(0x7fffffff), lpos(-1), tau_d((double)INFINITY), tau_idx(0x7fffffff) {}
    __device__ __forceinline__ double tau() const { return tau_d; }
    // Seeds the list with up to 32 DISTINCT tree points (this lane's `pos`, cloud-local sorted position, or -1):
    // distances to the new query are evaluated and the 32 entries are put in (d2, idx) order across the warp.  The
    // seeds arrive in the order of the previous query's list, which is almost the order for this query (consecutive
    // queries are neighbours on the space-filling curve), so an odd-even transposition sort that stops as soon as a whole
    // pass swaps nothing takes a few passes instead of the 15 stages of a full sorting network.  Any k distinct real
    // points bound the k-th nearest distance from above, so the result stays exact.
    __device__ __forceinline__ void seed(int pos) {
        ld = (double)INFINITY; lidx = 0x7fffffff; lpos = -1;
        if (pos >= 0 && pos < T.n) {
            TreePoint P = load_point(T.pts + T.pt_off + pos);
            double d = dist2_rn(P.x, P.y, P.z, qx, qy, qz);
            if (d == d) { ld = d; lidx = P.idx; lpos = pos; }
        }
        // comb passes at shrinking strides move far-displaced entries cheaply before the stride-1 passes
#pragma unroll
        for (int stride = 16; stride >= 2; stride >>= 1) {
            const int partner = lane ^ stride;
            const double od = shfl_d(ld, partner);
            const int oi = __shfl_sync(0xffffffffu, lidx, partner);
            const int op = __shfl_sync(0xffffffffu, lpos, partner);
            const bool other_less = lex_less(od, oi, ld, lidx);
            const bool take = lane < partner ? other_less : (!other_less && !(od == ld && oi == lidx));
            if (take) { ld = od; lidx = oi; lpos = op; }
        }
        bool swapped;
        do {
            swapped = false;
#pragma unroll
            for (int phase = 0; phase < 2; ++phase) {
                const int partner = phase == 0 ? (lane ^ 1) : ((lane & 1) ? lane + 1 : lane - 1);
                const bool valid = partner >= 0 && partner < 32;
                const int src = valid ? partner : lane;
                const double od = shfl_d(ld, src);
                const int oi = __shfl_sync(0xffffffffu, lidx, src);
                const int op = __shfl_sync(0xffffffffu, lpos, src);
                const bool other_less = lex_less(od, oi, ld, lidx);
                const bool take = valid && (lane < partner ? other_less : (!other_less && !(od == ld && oi == lidx)));
                if (take) { ld = od; lidx = oi; lpos = op; swapped = true; }
            }
        } while (__any_sync(0xffffffffu, swapped));
        tau_d = shfl_d(ld, k - 1);
        tau_idx = __shfl_sync(0xffffffffu, lidx, k - 1);
    }
    __device__ __forceinline__ void leaf(int p0, int cnt) {
        bool valid = lane < cnt;
        double cd = 0.0;
        int cidx = 0x7fffffff;
        if (valid) {
            TreePoint P = load_point(T.pts + T.pt_off + p0 + lane);
            cd = dist2_rn(P.x, P.y, P.z, qx, qy, qz);
            cidx = P.idx;
        }
        bool want = valid && lex_less(cd, cidx, tau_d, tau_idx);  // NaN compares false: never inserted
        // points of this leaf that are in the list already (seeds): each leaf is visited once per query, so these
        // are the only possible duplicates
        const unsigned off = (unsigned)(lpos - p0);
        const unsigned listed = __reduce_or_sync(0xffffffffu, (lpos >= 0 && off < 32u) ? (1u << off) : 0u);
        unsigned cm = __ballot_sync(0xffffffffu, want) & ~listed;
        while (cm) {
            int b = __ffs(cm) - 1;
            cm &= cm - 1u;
            double bd = shfl_d(cd, b);
            int bi = __shfl_sync(0xffffffffu, cidx, b);
            if (!lex_less(bd, bi, tau_d, tau_idx)) continue;  // warp-uniform: the bound moved past it
            int bp = p0 + b;
            int pos = __popc(__ballot_sync(0xffffffffu, lex_less(ld, lidx, bd, bi)));  // sorted: a prefix
            double ud = __shfl_up_sync(0xffffffffu, ld, 1);
            int ui = __shfl_up_sync(0xffffffffu, lidx, 1);
            int up = __shfl_up_sync(0xffffffffu, lpos, 1);
            if (lane > pos) { ld = ud; lidx = ui; lpos = up; }
            else if (lane == pos) { ld = bd; lidx = bi; lpos = bp; }
            tau_d = shfl_d(ld, k - 1);
            tau_idx = __shfl_sync(0xffffffffu, lidx, k - 1);
        }
    }
};

// -------------------------------------------------------------------------------------------------------------
// Leaf-local float32 arithmetic (TreeDesc::pts32): bounds of the exact squared distance from a float32 evaluation.
//
// pts32 holds every tree point as float32 offsets from its leaf box's lower corner (as stored: rounded down to
// float32).  Let S bound every leaf-local coordinate involved (the leaf's extent and the query's offsets from that
// corner).  Stored offsets and the query's offsets are within 2^-24 S of the exact ones, their float32 difference
// within 2^-22 S (1 + 2^-10) =: a of the exact coordinate difference, so | |d~vec| - d | <= sqrt(3) a.  With
// 2xy <= r x^2 + y^2 / r, r = 2^-19, and 3 * 2^-24 for the float32 evaluation of the sum of squares:
//     d2 <= d~ (1 + rho) + beta,   d2 >= d~ (1 - rho) - beta,   rho = 2^-18,  beta = a^2 (3 * 2^19 + 4),
// evaluated with directed rounding.  Valid while the tree's extent is moderate (callers check TreeDesc::gext < 1e15).
// -------------------------------------------------------------------------------------------------------------
#define SB_RHO_UP 1.000003814697265625f    // 1 + 2^-18
#define SB_RHO_DN 0.999996185302734375f    // 1 - 2^-18
#define SB_RHO2_DN 0.99999237060546875f    // 1 - 2^-17

// query offsets from the corner of the leaf box (b0, b1, b2) and the error term beta of this (query, leaf)
__device__ __forceinline__ void leaf_local_frame(double qx, double qy, double qz, float2 b0, float2 b1, float2 b2,
                                                 float& ox, float& oy, float& oz, float& beta) {
    ox = (float)(qx - (double)b0.x);
    oy = (float)(qy - (double)b0.y);
    oz = (float)(qz - (double)b1.x);
    float S = fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz));
    S = fmaxf(S, fmaxf(fmaxf(__fsub_ru(b1.y, b0.x), __fsub_ru(b2.x, b0.y)), __fsub_ru(b2.y, b1.x)));
    const float a = __fmul_ru(S, 2.386520565e-07f);          // 2^-22 (1 + 2^-10), rounded up
    beta = __fmul_ru(__fmul_ru(a, a), 1572868.0f);           // 3 * 2^19 + 4
}

// SB_LEAF_STAGE (default 1): the packet visitors fetch a leaf's 32 float32 candidates with ONE coalesced 512-byte load,
// lane i taking candidate i, and hand them round through a 512-byte row of shared memory per warp; as 32 broadcast
// loads the four lines of a leaf missed one after the other (14 % of k_self_knn's and 31 % of the first fallback pass's
// stall samples sat on the first use of a candidate).
#ifndef SB_LEAF_STAGE
#define SB_LEAF_STAGE 1
#endif
__device__ __forceinline__ void sts128(unsigned addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// -------------------------------------------------------------------------------------------------------------
// 1-NN of a PACKET of up to 32 nearby queries, one per lane (icp.cu): the packet shares one traversal (the query is
// the packet's bounding box); a visited leaf's points are read as leaf-local float32 (one broadcast load per
// candidate) and a lane evaluates a candidate in the oracle's fp64 arithmetic only if the float32 LOWER bound of its
// distance does not exceed the lane's best distance so far — so the result is the exact (d2, index) minimum, like
// NearestVisitor's.  Lanes with need == false take part in the warp-wide instructions but never search.
// -------------------------------------------------------------------------------------------------------------
struct NearestPacketVisitor {
    const TreeDesc& T;
    const int lane;
    bool need;
    double qx, qy, qz;
    double bd;       // best exact d2 so far (+inf: none)
    int bidx, bpos;  // its original row and sorted position
    float U;         // >= bd (0 for lanes that do not search)
    float U_warp;    // max over lanes
    const unsigned stage;   // shared-window address of the warp's 32 x float4 staging row (SB_LEAF_STAGE)
    __device__ __forceinline__ NearestPacketVisitor(const TreeDesc& t, int l, unsigned st_) : T(t), lane(l), stage(st_) {}
    __device__ __forceinline__ double tau() const { return (double)U_warp; }
    __device__ __forceinline__ void refresh() {
        U = need ? __double2float_ru(bd) : 0.f;
        U_warp = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(U)));  // non-negative floats
    }
    // (x, y, z): the lane's query; seed_pos: a tree point to start from (-1: none)
    __device__ __forceinline__ void init(bool want, double x, double y, double z, int seed_pos) {
        qx = x; qy = y; qz = z;
        bd = (double)INFINITY; bidx = 0x7fffffff; bpos = -1;
        need = want && x == x && y == y && z == z;   // a NaN query never matches (kdtree.hpp:125 strict <)
        if (need && seed_pos >= 0 && seed_pos < T.n) {
            const TreePoint P = load_point(T.pts + T.pt_off + seed_pos);
            const double d = dist2_rn(P.x, P.y, P.z, qx, qy, qz);
            if (d == d) { bd = d; bidx = P.idx; bpos = seed_pos; }
        }
        refresh();
    }
    __device__ __forceinline__ void leaf(int p0, int n) {
        const float2* b = reinterpret_cast<const float2*>(T.boxes + 6 * (T.box_off[0] + (p0 >> 5)));
        const float2 b0 = __ldg(b), b1 = __ldg(b + 1), b2 = __ldg(b + 2);
        PointQuery Q;
        Q.x = qx; Q.y = qy; Q.z = qz;
        const bool want = need && Q.lb(b0, b1, b2) <= U;
        if (!__any_sync(0xffffffffu, want)) return;
        float ox, oy, oz, beta;
        leaf_local_frame(qx, qy, qz, b0, b1, b2, ox, oy, oz, beta);
        const float4* __restrict__ c = T.pts32 + T.pt_off + p0;
        const TreePoint* __restrict__ TP = T.pts + T.pt_off + p0;
        float thr = want ? U : -1.f;   // lanes that do not want this leaf pass nothing
#if SB_LEAF_STAGE
        __syncwarp();
        if (lane < n) sts128(stage + 16u * lane, __ldg(c + lane));
        __syncwarp();
#endif
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
#if SB_LEAF_STAGE
            const float4 P32 = lds128(stage + 16u * i);
#else
            const float4 P32 = __ldg(c + i);  // same address in every lane: one broadcast load
#endif
            const float dx = P32.x - ox, dy = P32.y - oy, dz = P32.z - oz;
            const float dd = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
            const float lo = __fmaf_rd(dd, SB_RHO_DN, -beta);
            if (lo <= thr) {   // false for NaN
                const TreePoint P = load_point(TP + i);   // also one address for the whole warp
                const double d = dist2_rn(P.x, P.y, P.z, qx, qy, qz);
                if (d < bd || (d == bd && P.idx < bidx)) {   // kdtree.hpp:125 + the canonical tie rule
                    bd = d; bidx = P.idx; bpos = p0 + i;
                    thr = __double2float_ru(bd);
                }
            }
        }
        refresh();
    }
};

}  // namespace sb
