"""What the reference's predicate MEANS: "point strictly within polygon (even-odd over rings,
boundary excluded)", i.e. GeoPandas/shapely `sjoin(predicate="within")` -- BASELINE.json names
that as the semantic oracle.  shapely/geopandas are not installable offline, so the check is an
exact one: rational arithmetic on the doubles themselves (fractions.Fraction).  The reference
(hence the oracle, hence the CUDA path, which is bit-identical to it) may differ from the exact
answer only for points within its 4-ULP collinearity band of an edge
(detail/utility/floating_point.cuh:118-130, python/.../tests/binpreds/test_contains_properly.py:84-109).
"""
from fractions import Fraction as F

import numpy as np
import pytest

from cuspatial_b200 import datagen as D


def exact_within(px, py, rings):
    """Even-odd rule with exact arithmetic; a point on the boundary is NOT within."""
    x, y = F(float(px)), F(float(py))
    inside = False
    for ring in rings:
        n = len(ring)
        for i in range(n):
            ax, ay = F(float(ring[i][0])), F(float(ring[i][1]))
            bx, by = F(float(ring[(i + 1) % n][0])), F(float(ring[(i + 1) % n][1]))
            if (ax, ay) == (bx, by):
                continue
            cross = (bx - ax) * (y - ay) - (x - ax) * (by - ay)
            if cross == 0 and min(ax, bx) <= x <= max(ax, bx) and min(ay, by) <= y <= max(ay, by):
                return False, 0.0                                    # on the boundary
            if (ay > y) != (by > y):
                # x coordinate of the edge at height y, compared exactly with x
                if (x - ax) * (by - ay) * (1 if by > ay else -1) < (bx - ax) * (y - ay) * (1 if by > ay else -1):
                    inside = not inside
    return inside, None


def rel_distance_to_boundary(px, py, rings):
    """Distance to the nearest edge relative to that edge's length (float is enough here)."""
    best = np.inf
    for ring in rings:
        a = np.asarray(ring, dtype=np.float64)
        b = np.roll(a, -1, axis=0)
        e = b - a
        l2 = (e * e).sum(1)
        ok = l2 > 0
        t = np.clip(((px - a[:, 0]) * e[:, 0] + (py - a[:, 1]) * e[:, 1])[ok] / l2[ok], 0, 1)
        dx = a[ok, 0] + t * e[ok, 0] - px
        dy = a[ok, 1] + t * e[ok, 1] - py
        best = min(best, float(np.min(np.sqrt(dx * dx + dy * dy) / np.sqrt(l2[ok]))))
    return best


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_reference_predicate_is_strict_within_outside_the_ulp_band(oracle_lib, dtype):
    rng = np.random.default_rng(17)
    po, ro, vx, vy = D.taxi_zone_like_polygons(31, seed=5, dtype=dtype, median_vertices=20)
    ext = D.polygon_extent(vx, vy)
    n = 1500
    px, py = D.uniform_points(n, ext, seed=6, dtype=dtype)
    # a third of the points on / next to edges and vertices
    for i in range(0, n, 3):
        j = int(rng.integers(0, len(vx) - 1))
        t = dtype(rng.uniform()) if i % 2 else dtype(0)
        px[i] = vx[j] + t * (vx[j + 1] - vx[j])
        py[i] = vy[j] + t * (vy[j + 1] - vy[j])
        if i % 4 == 0:
            px[i] = np.nextafter(px[i], dtype(np.inf))
    mask = oracle_lib.point_in_polygon(px, py, po.astype(np.int32), ro.astype(np.int32), vx, vy)
    band = 1e-4 if dtype == np.float32 else 1e-12   # >> 4 ulp of the products, << any real gap
    checked = disagreements_in_band = 0
    for p in range(31):
        rings = [list(zip(vx[ro[r]:ro[r + 1] - 1], vy[ro[r]:ro[r + 1] - 1]))   # drop closing vertex
                 for r in range(po[p], po[p + 1])]
        for i in range(0, n, 2):
            want, _ = exact_within(px[i], py[i], rings)
            got = bool((mask[i] >> p) & 1)
            if got != want:
                assert rel_distance_to_boundary(float(px[i]), float(py[i]), rings) < band, \
                    (p, i, got, want)
                disagreements_in_band += 1
            checked += 1
    assert checked > 20000
    # exact boundary points are "not within" for both
    assert disagreements_in_band < checked * 0.01
