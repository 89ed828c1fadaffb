"""Host-side model of the two warp reductions icp.cu has used for a work item's 29 sums (CPU, no GPU needed).

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
