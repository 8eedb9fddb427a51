"""ctypes bindings of the CPU checkers -- TEST INFRASTRUCTURE ONLY.

Two libraries with the same C ABI shape (oracle/oracle.cpp, oracle/ref_driver.cpp):
  oracle()    -> oracle/liboracle.so                  (prefix orc_) the repo's CPU restatement
  reference() -> oracle/_ref/libcuspatial_ref_host.so (prefix ref_) the reference's own
                 header-only implementation compiled in place for the host (Thrust OpenMP)
Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing in the product package (cuspatial_b200/) imports this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_PATH = os.path.join(_HERE, "liboracle.so")
REF_HOST_PATH = os.path.join(_HERE, "_ref", "libcuspatial_ref_host.so")


def _dt(a):
    if a.dtype == np.float32:
        return 0
    if a.dtype == np.float64:
        return 1
    raise TypeError("float32/float64 only")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class HostLib:
    """Host-pointer implementation of the four hot-path entry points (+ polygon bboxes)."""

    def __init__(self, path, prefix, kind):
        self.path, self.prefix, self.kind = path, prefix, kind
        self._lib = C.CDLL(path)
        self._f("last_error").restype = C.c_char_p
        self._f("free").argtypes = [C.c_void_p]

    def _f(self, name):
        return getattr(self._lib, self.prefix + name)

    def _take(self, ptr, n, dtype):
        if n == 0 or not ptr:
            return np.empty(0, dtype=dtype)
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        out = np.frombuffer(buf, dtype=dtype, count=n).copy()
        self._f("free")(C.c_void_p(ptr))
        return out

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self._f("last_error")().decode())

    def quadtree_on_points(self, x, y, x_min, x_max, y_min, y_max, scale, max_depth, max_size):
        x = np.ascontiguousarray(x)
        y = np.ascontiguousarray(y, dtype=x.dtype)
        out = (C.c_void_p * 6)()
        out_n = (C.c_uint64 * 2)()
        self._check(self._f("quadtree_on_points")(
            _p(x), _p(y), _dt(x), C.c_uint64(len(x)), C.c_double(x_min), C.c_double(x_max),
            C.c_double(y_min), C.c_double(y_max), C.c_double(scale), int(max_depth),
            int(max_size), out, out_n))
        n, q = out_n[0], out_n[1]
        return {
            "point_indices": self._take(out[0], n, np.uint32),
            "key": self._take(out[1], q, np.uint32),
            "level": self._take(out[2], q, np.uint8),
            "is_internal_node": self._take(out[3], q, np.uint8),
            "length": self._take(out[4], q, np.uint32),
            "offset": self._take(out[5], q, np.uint32),
        }

    @staticmethod
    def _tree_args(tree):
        key = np.ascontiguousarray(tree["key"], dtype=np.uint32)
        level = np.ascontiguousarray(tree["level"], dtype=np.uint8)
        internal = np.ascontiguousarray(tree["is_internal_node"], dtype=np.uint8)
        length = np.ascontiguousarray(tree["length"], dtype=np.uint32)
        offset = np.ascontiguousarray(tree["offset"], dtype=np.uint32)
        keep = (key, level, internal, length, offset)
        return keep, [_p(a) for a in keep] + [C.c_uint64(len(key))]

    def join_quadtree_and_bounding_boxes(self, tree, bx0, by0, bx1, by1, x_min, y_min, scale,
                                         max_depth):
        bx0 = np.ascontiguousarray(bx0)
        by0, bx1, by1 = (np.ascontiguousarray(a, dtype=bx0.dtype) for a in (by0, bx1, by1))
        keep, targs = self._tree_args(tree)
        out = (C.c_void_p * 2)()
        out_n = (C.c_uint64 * 1)()
        self._check(self._f("join_quadtree_and_bounding_boxes")(
            *targs, _p(bx0), _p(by0), _p(bx1), _p(by1), _dt(bx0), C.c_uint64(len(bx0)),
            C.c_double(x_min), C.c_double(y_min), C.c_double(scale), int(max_depth), out, out_n))
        p = out_n[0]
        return self._take(out[0], p, np.uint32), self._take(out[1], p, np.uint32)

    def quadtree_point_in_polygon(self, pair_poly, pair_quad, tree, point_indices, px, py,
                                  poly_offsets, ring_offsets, vx, vy):
        px = np.ascontiguousarray(px)
        py, vx, vy = (np.ascontiguousarray(a, dtype=px.dtype) for a in (py, vx, vy))
        pair_poly = np.ascontiguousarray(pair_poly, dtype=np.uint32)
        pair_quad = np.ascontiguousarray(pair_quad, dtype=np.uint32)
        point_indices = np.ascontiguousarray(point_indices, dtype=np.uint32)
        poly_offsets = np.ascontiguousarray(poly_offsets, dtype=np.uint32)
        ring_offsets = np.ascontiguousarray(ring_offsets, dtype=np.uint32)
        keep, targs = self._tree_args(tree)
        out = (C.c_void_p * 2)()
        out_n = (C.c_uint64 * 1)()
        self._check(self._f("quadtree_point_in_polygon")(
            _p(pair_poly), _p(pair_quad), C.c_uint64(len(pair_poly)), *targs, _p(point_indices),
            _p(px), _p(py), _dt(px), C.c_uint64(len(px)), _p(poly_offsets),
            C.c_uint64(len(poly_offsets)), _p(ring_offsets), C.c_uint64(len(ring_offsets)),
            _p(vx), _p(vy), C.c_uint64(len(vx)), out, out_n))
        h = out_n[0]
        return self._take(out[0], h, np.uint32), self._take(out[1], h, np.uint32)

    def point_in_polygon(self, px, py, poly_offsets, ring_offsets, vx, vy):
        px = np.ascontiguousarray(px)
        py, vx, vy = (np.ascontiguousarray(a, dtype=px.dtype) for a in (py, vx, vy))
        poly_offsets = np.ascontiguousarray(poly_offsets, dtype=np.int32)
        ring_offsets = np.ascontiguousarray(ring_offsets, dtype=np.int32)
        out = np.zeros(len(px), dtype=np.int32)
        self._check(self._f("point_in_polygon")(
            _p(px), _p(py), _dt(px), C.c_uint64(len(px)), _p(poly_offsets),
            C.c_uint64(len(poly_offsets)), _p(ring_offsets), C.c_uint64(len(ring_offsets)),
            _p(vx), _p(vy), C.c_uint64(len(vx)), _p(out)))
        return out

    def pairwise_point_in_polygon(self, px, py, poly_offsets, ring_offsets, vx, vy):
        px = np.ascontiguousarray(px)
        py, vx, vy = (np.ascontiguousarray(a, dtype=px.dtype) for a in (py, vx, vy))
        poly_offsets = np.ascontiguousarray(poly_offsets, dtype=np.int32)
        ring_offsets = np.ascontiguousarray(ring_offsets, dtype=np.int32)
        out = np.zeros(len(px), dtype=np.uint8)
        self._check(self._f("pairwise_point_in_polygon")(
            _p(px), _p(py), _dt(px), C.c_uint64(len(px)), _p(poly_offsets),
            C.c_uint64(len(poly_offsets)), _p(ring_offsets), C.c_uint64(len(ring_offsets)),
            _p(vx), _p(vy), C.c_uint64(len(vx)), _p(out)))
        return out

    def quadtree_point_to_nearest_linestring(self, pair_line, pair_quad, tree, point_indices, px, py,
                                             line_offsets, lx, ly):
        px = np.ascontiguousarray(px)
        py, lx, ly = (np.ascontiguousarray(a, dtype=px.dtype) for a in (py, lx, ly))
        pair_line = np.ascontiguousarray(pair_line, dtype=np.uint32)
        pair_quad = np.ascontiguousarray(pair_quad, dtype=np.uint32)
        point_indices = np.ascontiguousarray(point_indices, dtype=np.uint32)
        line_offsets = np.ascontiguousarray(line_offsets, dtype=np.uint32)
        keep, targs = self._tree_args(tree)
        out = (C.c_void_p * 3)()
        out_n = (C.c_uint64 * 1)()
        self._check(self._f("quadtree_point_to_nearest_linestring")(
            _p(pair_line), _p(pair_quad), C.c_uint64(len(pair_line)), *targs, _p(point_indices),
            _p(px), _p(py), _dt(px), C.c_uint64(len(px)), _p(line_offsets),
            C.c_uint64(len(line_offsets)), _p(lx), _p(ly), C.c_uint64(len(lx)), out, out_n))
        n = out_n[0]
        return (self._take(out[0], n, np.uint32), self._take(out[1], n, np.uint32),
                self._take(out[2], n, px.dtype))

    def linestring_bounding_boxes(self, line_offsets, lx, ly, expansion=0.0):
        lx = np.ascontiguousarray(lx)
        ly = np.ascontiguousarray(ly, dtype=lx.dtype)
        line_offsets = np.ascontiguousarray(line_offsets, dtype=np.uint32)
        n = len(line_offsets) - 1
        outs = [np.zeros(n, dtype=lx.dtype) for _ in range(4)]
        self._check(self._f("linestring_bounding_boxes")(
            _p(line_offsets), C.c_uint64(len(line_offsets)), _p(lx), _p(ly), _dt(lx),
            C.c_uint64(len(lx)), C.c_double(expansion), *[_p(o) for o in outs]))
        return tuple(outs)

    def polygon_bounding_boxes(self, poly_offsets, ring_offsets, vx, vy, expansion=0.0):
        vx = np.ascontiguousarray(vx)
        vy = np.ascontiguousarray(vy, dtype=vx.dtype)
        poly_offsets = np.ascontiguousarray(poly_offsets, dtype=np.uint32)
        ring_offsets = np.ascontiguousarray(ring_offsets, dtype=np.uint32)
        n = len(poly_offsets) - 1
        outs = [np.zeros(n, dtype=vx.dtype) for _ in range(4)]
        self._check(self._f("polygon_bounding_boxes")(
            _p(poly_offsets), C.c_uint64(len(poly_offsets)), _p(ring_offsets),
            C.c_uint64(len(ring_offsets)), _p(vx), _p(vy), _dt(vx), C.c_uint64(len(vx)),
            C.c_double(expansion), *[_p(o) for o in outs]))
        return tuple(outs)


_cache = {}


def oracle():
    """The repo's CPU restatement (always available once oracle/Makefile has run)."""
    if "orc" not in _cache:
        _cache["orc"] = HostLib(ORACLE_PATH, "orc_", "port")
    return _cache["orc"]


def reference_available():
    return os.path.exists(REF_HOST_PATH)


def reference():
    """The reference's own implementation on the host (present if it was compiled here)."""
    if "ref" not in _cache:
        _cache["ref"] = HostLib(REF_HOST_PATH, "ref_", "reference")
    return _cache["ref"]
