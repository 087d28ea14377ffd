"""Size-independent properties of the sensing / routing math (hypothesis), checked on the oracle; the GPU parity
tests compare the kernels with this oracle, so these are the invariants the kernels inherit."""
import numpy as np
from hypothesis import given, settings, strategies as st
from hypothesis.extra import numpy as hnp

from oracle import forde_oracle as O

finite = st.floats(min_value=-1e3, max_value=1e3, allow_nan=False, allow_infinity=False, width=32)


@settings(max_examples=60, deadline=None)
@given(hnp.arrays(np.float32, st.integers(2, 64), elements=finite))
def test_hoyer_is_bounded_and_scale_invariant(x):
    h = float(O.hoyer_sparsity(x))
    assert -1e-4 <= h <= 1.0 + 1e-4
    if np.abs(x).max() > 1e-3:
        assert abs(float(O.hoyer_sparsity(x * np.float32(3.5))) - h) < 1e-4
    assert abs(float(O.hoyer_sparsity(-x)) - h) < 1e-6


@settings(max_examples=40, deadline=None)
@given(hnp.arrays(np.float32, st.tuples(st.integers(2, 40), st.integers(1, 6)), elements=finite), st.integers(1, 38))
def test_moments_pool_over_any_split(a, cut):
    cut = min(cut, a.shape[0] - 1)
    g = a[::-1].copy()
    m, n = O.raw_moments(a, g)
    m1, n1 = O.raw_moments(a[:cut], g[:cut]); m2, n2 = O.raw_moments(a[cut:], g[cut:])
    assert n1 + n2 == n
    np.testing.assert_allclose(m1 + m2, m, rtol=1e-9, atol=1e-9)
    s = O.stats_from_moments(m, n)
    assert np.all(s[:, 4] >= 0) and np.all(s[:, [1, 3]] >= 0)      # variance and mean |.| are non-negative


@settings(max_examples=60, deadline=None)
@given(hnp.arrays(np.float32, st.tuples(st.integers(1, 8), st.integers(2, 16)), elements=st.sampled_from([0.0, 1.0, -1.0, 2.5])),
       st.integers(1, 4))
def test_top_k_order_with_heavy_ties(x, k):
    k = min(k, x.shape[1])
    vals, idx = O.top_k(x, k)
    for r in range(x.shape[0]):
        assert np.all(np.diff(vals[r]) <= 0)                                  # descending
        for j in range(k - 1):
            if vals[r, j] == vals[r, j + 1]:
                assert idx[r, j] < idx[r, j + 1]                              # ties: lower index first
        rest = np.delete(x[r], idx[r])
        assert rest.size == 0 or rest.max() <= vals[r, -1]                    # nothing larger was left out


@settings(max_examples=40, deadline=None)
@given(hnp.arrays(np.float32, st.tuples(st.integers(1, 6), st.integers(3, 9)), elements=finite),
       hnp.arrays(np.int64, st.integers(3, 9), elements=st.integers(0, 3)))
def test_specialists_pass_gradients_straight_through(z, a):
    a = a[: z.shape[1]]
    z = z[:, : a.shape[0]]
    g = O.multiplex_grad(z, a) + O.gate_of(a)
    spec = ~np.isin(a, (0, 1))
    assert np.allclose(g[:, spec], 1.1)                                       # STE slope 1 + gate 0.1, whatever z is
    relu = a == 0
    assert np.all(np.isin(g[:, relu], (1.0, 2.0)))                            # gate 1 + relu' in {0, 1}
