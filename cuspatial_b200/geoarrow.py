"""GeoArrow buffers in, the hot path's columns out.

The reference's spatial-join functions take `cuspatial.GeoSeries` objects and read four things
from them: `points.points.x / .y`, `polygons.polygons.part_offset / .ring_offset / .x / .y`
(+ `.geometry_offset` for the "no multipolygon" check) and `linestrings.lines.part_offset /
.geometry_offset / .x / .y` (core/spatial/join.py:89-95,230-242,311-330, bounding.py:62-65,
123-128).  A GeoSeries of that kind is built from raw GeoArrow buffers by

    GeoSeries.from_points_xy(points_xy)                               core/geoseries.py:670
    GeoSeries.from_linestrings_xy(xy, part_offset, geometry_offset)    core/geoseries.py:729
    GeoSeries.from_polygons_xy(xy, ring_offset, part_offset, geometry_offset)   :776

This module provides the same three constructors over device tensors, with no cuDF: the result
objects expose exactly the accessors above (`.points`, `.polygons`, `.lines`), and every function
of `cuspatial_b200.api` accepts them wherever it accepts `(x, y)` / `(part_offset, ring_offset,
x, y)` / `(part_offset, x, y)` tuples.  The interleaved xy buffer is kept as is (zero-copy); the
x and y columns the C ABI wants are strided copies made once, on first use, like the reference's
`GeoColumnAccessor.x / .y` (core/geoseries.py:238-243).
"""
import torch


def _as_cuda_1d(a, name, dtype=None):
    # tensors stay where they are (the API functions insist on device memory when they are
    # called); anything else (numpy, lists) is uploaded when a GPU is present
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.as_tensor(a, device="cuda" if torch.cuda.is_available() else "cpu")
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.dim() == 2 and t.shape[1] == 2 and name.endswith("xy"):
        t = t.reshape(-1)
    if t.dim() != 1:
        raise ValueError("%s must be one-dimensional" % name)
    return t.contiguous()


def _check_coords(xy, name):
    if xy.dtype not in (torch.float32, torch.float64):
        raise TypeError("%s must be float32 or float64" % name)  # geoseries.py _check_coords_dtype
    if xy.numel() % 2:
        raise ValueError("%s must hold an even number of interleaved x, y values" % name)


class _Coords:
    """Interleaved xy buffer + lazily materialised x / y columns."""

    def __init__(self, xy):
        self.xy = xy
        self._x = self._y = None

    @property
    def x(self):
        if self._x is None:
            self._x = self.xy[0::2].contiguous()
        return self._x

    @property
    def y(self):
        if self._y is None:
            self._y = self.xy[1::2].contiguous()
        return self._y

    def __len__(self):
        return self.xy.numel() // 2


class PointAccessor(_Coords):
    """`GeoSeries.points` (x, y, xy)."""


class LineStringAccessor(_Coords):
    """`GeoSeries.lines` (geometry_offset, part_offset, x, y, xy)."""

    def __init__(self, xy, part_offset, geometry_offset):
        super().__init__(xy)
        self.part_offset = part_offset
        self.geometry_offset = geometry_offset


class PolygonAccessor(_Coords):
    """`GeoSeries.polygons` (geometry_offset, part_offset, ring_offset, x, y, xy)."""

    def __init__(self, xy, ring_offset, part_offset, geometry_offset):
        super().__init__(xy)
        self.ring_offset = ring_offset
        self.part_offset = part_offset
        self.geometry_offset = geometry_offset


class GeoArrowSeries:
    """A single-geometry-type column of GeoArrow buffers on the device: the part of
    `cuspatial.GeoSeries` the spatial-join path reads."""

    def __init__(self, kind, accessor, length):
        self.kind = kind
        self._acc = accessor
        self._len = int(length)

    def __len__(self):
        return self._len

    def _get(self, kind):
        if self.kind != kind:
            raise ValueError("this series holds %s geometries, not %s" % (self.kind, kind))
        return self._acc

    @property
    def points(self):
        return self._get("points")

    @property
    def lines(self):
        return self._get("linestrings")

    @property
    def polygons(self):
        return self._get("polygons")

    # what cuspatial_b200.api reads directly (a series passed where a tuple is accepted)
    @property
    def part_offset(self):
        return self._acc.part_offset

    @property
    def ring_offset(self):
        return self._acc.ring_offset

    @property
    def x(self):
        return self._acc.x

    @property
    def y(self):
        return self._acc.y


def from_points_xy(points_xy):
    """GeoSeries.from_points_xy: POINTs from interleaved xy coordinates."""
    xy = _as_cuda_1d(points_xy, "points_xy")
    _check_coords(xy, "points_xy")
    return GeoArrowSeries("points", PointAccessor(xy), xy.numel() // 2)


def from_linestrings_xy(linestrings_xy, part_offset, geometry_offset):
    """GeoSeries.from_linestrings_xy: (MULTI)LINESTRINGs from interleaved xy coordinates,
    offsets of each part into the coordinates and of each geometry into the parts."""
    xy = _as_cuda_1d(linestrings_xy, "linestrings_xy")
    _check_coords(xy, "linestrings_xy")
    po = _as_cuda_1d(part_offset, "part_offset", torch.int32)
    go = _as_cuda_1d(geometry_offset, "geometry_offset", torch.int32)
    return GeoArrowSeries("linestrings", LineStringAccessor(xy, po, go), max(go.numel() - 1, 0))


def from_polygons_xy(polygons_xy, ring_offset, part_offset, geometry_offset):
    """GeoSeries.from_polygons_xy: (MULTI)POLYGONs from interleaved xy coordinates, offsets of
    each ring into the coordinates, of each part (polygon) into the rings and of each geometry
    into the parts."""
    xy = _as_cuda_1d(polygons_xy, "polygons_xy")
    _check_coords(xy, "polygons_xy")
    ro = _as_cuda_1d(ring_offset, "ring_offset", torch.int32)
    po = _as_cuda_1d(part_offset, "part_offset", torch.int32)
    go = _as_cuda_1d(geometry_offset, "geometry_offset", torch.int32)
    return GeoArrowSeries("polygons", PolygonAccessor(xy, ro, po, go), max(go.numel() - 1, 0))


def is_multi(series):
    """True if some geometry has more than one part (the reference's
    `len(part_offset) != len(geometry_offset)` test, join.py:75-78,323-326)."""
    acc = series._acc
    return hasattr(acc, "geometry_offset") and acc.part_offset.numel() != acc.geometry_offset.numel()
