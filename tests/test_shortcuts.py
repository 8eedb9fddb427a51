"""The two arithmetic shortcuts the CUDA kernels take, restated in numpy and checked against the
literal formula on random and adversarial values (the kernels themselves are checked bit for bit
against the oracle on the GPU; this pins the ARGUMENT on the CPU, where it can be hammered).

1. encode (csrc/quadtree.cu, cell_index): floor(a / scale) == trunc(a * (1 / scale)) whenever the
   product's fractional part is farther than `guard` from 0 and 1 and the product is below 65536.
2. nearest linestring (csrc/nearest_linestring.cu): with m = RN(d3*d3),
   m > RN(RN(d2*d2) * 1.00001)  implies  RN(m / d2) >= d2.
"""
import numpy as np
import pytest


def _adversarial_quotients(rng, dt, scale, n):
    """a such that a / scale sits within a few ulps of an integer, plus random ones."""
    k = rng.integers(0, 65536, n).astype(np.float64)
    near = (k * np.float64(scale)).astype(dt)
    steps = rng.integers(-6, 7, n)
    for _ in range(6):                       # walk up to 6 ulps either side of k * scale
        near = np.where(steps > 0, np.nextafter(near, dt(np.inf)), near)
        near = np.where(steps < 0, np.nextafter(near, dt(-np.inf)), near)
        steps = steps - np.sign(steps)
    rand = (rng.random(n) * 65536.0 * np.float64(scale)).astype(dt)
    return np.abs(np.concatenate([near, rand])).astype(dt)


@pytest.mark.parametrize("scale", [1.0, 1.0 / 3.0, 0.00030517578125, 7.345e-5, 12.5, 1e-9 + 1e-3])
def test_fp64_cell_index_guard_band_never_disagrees_with_the_division(scale):
    rng = np.random.default_rng(int(scale * 1e9) % 2**31)
    dt = np.float64
    s = dt(scale)
    inv = dt(1) / s
    guard = dt(2.0 ** -30)
    a = _adversarial_quotients(rng, dt, s, 2_000_000)
    q = a * inv
    t = np.trunc(q)
    f = q - t
    fast = (q < 65536) & (f > guard) & (f < 1 - guard)
    exact = np.trunc(a / s)                       # IEEE division, then cvt.rzi
    assert fast.mean() > 0.4                      # the shortcut is actually taken
    bad = fast & (t != exact)
    assert not bad.any(), (a[bad][:3], q[bad][:3], exact[bad][:3])
    # and the guard is needed: without it the product and the quotient do disagree somewhere
    assert (np.trunc(q) != exact)[q < 65536].any() or scale in (1.0, 12.5, 0.00030517578125)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_division_skip_rule_of_the_nearest_linestring_kernel(dt):
    rng = np.random.default_rng(5)
    n = 3_000_000
    d2 = np.exp(rng.uniform(-30, 30, n)).astype(dt)
    # d3 around d2 (the decision boundary r == d2 is d3 == d2), random elsewhere
    rel = np.concatenate([1 + rng.uniform(-1e-5, 1e-5, n // 2) * (4 if dt == np.float32 else 1e-8),
                          np.exp(rng.uniform(-3, 3, n - n // 2))])
    d3 = (d2.astype(np.float64) * rel).astype(dt)
    with np.errstate(over="ignore", under="ignore", invalid="ignore", divide="ignore"):
        m = d3 * d3
        certain = m > (d2 * d2) * dt(1.00001)
        r = m / d2
        literal = r >= d2
    assert certain.sum() > n // 10
    assert literal[certain].all()                 # the skipped division could not have said "<"
    # tiny / huge magnitudes: overflow and underflow must fall through to the literal path or
    # still be right
    for scale in (dt(1e-30 if dt == np.float32 else 1e-200), dt(1e18 if dt == np.float32 else 1e150)):
        with np.errstate(over="ignore", under="ignore", invalid="ignore", divide="ignore"):
            e2, e3 = d2 * scale, d3 * scale
            mm = e3 * e3
            cc = mm > (e2 * e2) * dt(1.00001)
            ok = (mm / e2 >= e2) | ~cc | (e2 == 0)
        assert ok.all()
