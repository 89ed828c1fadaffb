"""Host-side models of device-side bookkeeping whose invariants the kernels rely on (CPU, no GPU needed).

1. The two warp reductions icp.cu has used for a work item's 29 sums.

k_icp_match / k_icp_accum leave sum t of the item's 32 source points in lane t.  Until round 2 every sum was its own
xor-butterfly (145 shuffle steps for 29 sums); `item_sum` in icp.cu is a transposing butterfly: at the step with
partner lane ^ m a lane keeps the sums whose index has its own bit m and sends the others (31 shuffle steps).
DESIGN.md claims that both perform the same additions on the same operands, i.e. give the same bits — which is what
lets the kernel change without changing any ICP result.  This test checks the claim on a 32-lane model in fp64."""
import numpy as np


def butterfly(v):
    """v[lane][t]: 32 x 32 doubles.  Every sum t by its own xor-butterfly; the value lane t ends up with."""
    out = np.empty(32)
    for t in range(32):
        x = v[:, t].copy()
        for m in (16, 8, 4, 2, 1):
            x = x + x[np.arange(32) ^ m]          # v += shfl_xor(v, m), every lane
        out[t] = x[t]
    return out


def transposing(v):
    """item_sum<1, 0>: slot T after the step with mask M combines the slots T and T + M of the step before."""
    w = [list(v[lane]) for lane in range(32)]     # w[lane][slot]
    for m in (16, 8, 4, 2, 1):
        nxt = []
        for lane in range(32):
            up = (lane & m) != 0
            row = []
            for t in range(m):
                lo, hi = w[lane][t], w[lane][t + m]
                plo, phi = w[lane ^ m][t], w[lane ^ m][t + m]
                keep = hi if up else lo
                partner_up = ((lane ^ m) & m) != 0            # (= not up)
                recv = plo if partner_up else phi             # the partner sends the slot IT does not keep
                row.append(keep + recv)
            nxt.append(row)
        w = nxt
    return np.array([w[lane][0] for lane in range(32)])


def test_transposing_butterfly_is_bit_identical_to_per_sum_butterflies():
    rng = np.random.default_rng(7)
    for trial in range(50):
        scale = 10.0 ** rng.integers(-8, 8, size=(32, 32))
        v = rng.standard_normal((32, 32)) * scale          # cancellation-prone: any re-association would show
        a, b = butterfly(v), transposing(v)
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), trial


def test_transposing_butterfly_sums_the_right_terms():
    v = np.zeros((32, 32))
    for lane in range(32):
        for t in range(32):
            v[lane, t] = (lane + 1) * 1000 + t           # exact in fp64
    got = transposing(v)
    want = np.array([sum((lane + 1) * 1000 + t for lane in range(32)) for t in range(32)], dtype=np.float64)
    assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------------------------
# Two more host-side models of device-side bookkeeping (icp.cu / forest.cu), for their edge cases.
# ---------------------------------------------------------------------------------------------------------------
def warp_search(off, x):
    """find_segment / find_active (n <= 256): largest s with off[s] <= x, 32 evenly spaced probes per step."""
    n = len(off)
    lo, length, steps = 0, n, 0
    while length > 1:
        step = (length + 31) >> 5
        le = [(lo + lane * step < lo + length) and off[lo + lane * step] <= x for lane in range(32)]
        ballot = sum(1 << lane for lane in range(32) if le[lane]) | 1
        j = ballot.bit_length() - 1               # 31 - clz
        end = lo + length
        lo += j * step
        length = min(step, end - lo)
        steps += 1
    return lo, steps


def test_warp_wide_segment_search_matches_bisect_including_empty_segments():
    import bisect
    rng = np.random.default_rng(3)
    for n in (1, 2, 31, 32, 33, 255, 256, 257, 1024, 4096, 5000):
        sizes = rng.integers(0, 4, size=n)         # empty segments give repeated offsets
        sizes[0] = max(sizes[0], 1)
        off = np.concatenate([[0], np.cumsum(sizes)])[:n].tolist()
        total = int(np.sum(sizes))
        for x in list(range(0, min(total, 200))) + rng.integers(0, max(total, 1), size=200).tolist():
            want = bisect.bisect_right(off, x) - 1
            got, steps = warp_search(off, x)
            assert got == want, (n, x)
            assert steps <= 3 or n > 32768


def test_work_iterator_hands_every_item_out_exactly_once():
    """WorkIter (icp.cu): warps take `chunk` consecutive items per fetch from a shared counter; with no more items than
    warps, warp w takes item w without touching the counter."""
    rng = np.random.default_rng(5)
    for n, warps, chunk in ((0, 64, 4), (1, 64, 1), (63, 64, 4), (64, 64, 4), (65, 64, 4), (1000, 64, 4), (1003, 7, 16),
                            (300000, 9472, 4)):
        counter = [0]
        seen = np.zeros(n, dtype=np.int32)
        state = []
        for w in range(warps):
            if n <= warps:
                state.append({"chunk": 0, "cur": w, "end": 0})
            else:
                state.append({"chunk": chunk, "cur": -1, "end": 0})
        live = list(range(warps))
        while live:
            w = live[int(rng.integers(len(live)))]      # warps advance in any interleaving
            s = state[w]
            if s["chunk"] == 0:
                it, have = s["cur"], s["cur"] < n
                s["cur"] = n
            else:
                if s["cur"] + 1 >= s["end"]:
                    s["cur"] = counter[0]
                    counter[0] += s["chunk"]
                    s["end"] = s["cur"] + s["chunk"]
                else:
                    s["cur"] += 1
                it, have = s["cur"], s["cur"] < n
            if have:
                seen[it] += 1
            else:
                live.remove(w)
        assert np.all(seen == 1), (n, warps, chunk)
