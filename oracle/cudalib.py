"""ctypes bindings of the reference's CUDA flavour -- TEST / BASELINE INFRASTRUCTURE ONLY.

oracle/_ref/libcuspatial_ref_cuda.so is the reference's own header-only implementation
(Thrust/CUB kernels, /root/reference/cpp/include compiled in place by `make -C oracle ref_cuda`)
behind the same C ABI shape as the host checkers, but every data pointer is a DEVICE pointer.
It is used by tests/ (GPU-vs-GPU parity at sizes the CPU checkers cannot reach) and by bench.py's
`gpu_reference` leg (the reference's CUDA build timed on the same B200, BASELINE.json north_star).
Nothing in the product package (cuspatial_b200/) imports this module.
"""
import ctypes as C
import os
import time

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_CUDA_PATH = os.path.join(_HERE, "_ref", "libcuspatial_ref_cuda.so")

_NP_TO_TORCH = {"u32": torch.uint32, "u8": torch.uint8}


def available():
    return os.path.exists(REF_CUDA_PATH) and torch.cuda.is_available()


def _dt(t):
    if t.dtype == torch.float32:
        return 0
    if t.dtype == torch.float64:
        return 1
    raise TypeError("float32/float64 only")


def _p(t):
    return C.c_void_p(t.data_ptr() if t.numel() else 0)


class CudaRefLib:
    """Device-pointer driver of the reference's four hot-path entry points.

    Every method returns (result, seconds): the reference call is synchronous (it ends with a
    stream synchronize, like libcuspatial's column API seen from Python), so host wall-clock
    around the call is its device time plus its own launch/sync overhead -- exactly what a
    cuspatial user pays.  Copying the result into torch tensors happens outside that interval.
    """

    def __init__(self, path=REF_CUDA_PATH):
        self._lib = C.CDLL(path)
        self._lib.ref_last_error.restype = C.c_char_p
        self._lib.ref_free.argtypes = [C.c_void_p]
        assert self._lib.ref_is_cuda() == 1
        self._rt = C.CDLL("libcudart.so.12")
        self._rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self._lib.ref_last_error().decode())

    def _take(self, ptr, n, dtype, dev):
        out = torch.empty(int(n), dtype=dtype, device=dev)
        if n and ptr:
            rc = self._rt.cudaMemcpy(C.c_void_p(out.data_ptr()), C.c_void_p(ptr),
                                     out.numel() * out.element_size(), 3)
            assert rc == 0, "cudaMemcpy failed: %d" % rc
            self._lib.ref_free(C.c_void_p(ptr))
        return out

    def quadtree_on_points(self, x, y, x_min, x_max, y_min, y_max, scale, max_depth, max_size):
        out = (C.c_void_p * 6)()
        out_n = (C.c_uint64 * 2)()
        torch.cuda.synchronize(x.device)
        t0 = time.perf_counter()
        rc = self._lib.ref_quadtree_on_points(
            _p(x), _p(y), _dt(x), C.c_uint64(x.numel()), C.c_double(x_min), C.c_double(x_max),
            C.c_double(y_min), C.c_double(y_max), C.c_double(scale), int(max_depth),
            int(max_size), out, out_n)
        dt = time.perf_counter() - t0
        self._check(rc)
        n, q = out_n[0], out_n[1]
        d = x.device
        return {
            "point_indices": self._take(out[0], n, torch.uint32, d),
            "key": self._take(out[1], q, torch.uint32, d),
            "level": self._take(out[2], q, torch.uint8, d),
            "is_internal_node": self._take(out[3], q, torch.uint8, d),
            "length": self._take(out[4], q, torch.uint32, d),
            "offset": self._take(out[5], q, torch.uint32, d),
        }, dt

    @staticmethod
    def _tree_args(tree):
        keep = tuple(tree[k].contiguous() for k in
                     ("key", "level", "is_internal_node", "length", "offset"))
        return keep, [_p(a) for a in keep] + [C.c_uint64(keep[0].numel())]

    def join_quadtree_and_bounding_boxes(self, tree, bx0, by0, bx1, by1, x_min, y_min, scale,
                                         max_depth):
        keep, targs = self._tree_args(tree)
        out = (C.c_void_p * 2)()
        out_n = (C.c_uint64 * 1)()
        torch.cuda.synchronize(bx0.device)
        t0 = time.perf_counter()
        rc = self._lib.ref_join_quadtree_and_bounding_boxes(
            *targs, _p(bx0), _p(by0), _p(bx1), _p(by1), _dt(bx0), C.c_uint64(bx0.numel()),
            C.c_double(x_min), C.c_double(y_min), C.c_double(scale), int(max_depth), out, out_n)
        dt = time.perf_counter() - t0
        self._check(rc)
        p = out_n[0]
        d = bx0.device
        return (self._take(out[0], p, torch.uint32, d), self._take(out[1], p, torch.uint32, d)), dt

    def quadtree_point_in_polygon(self, pair_poly, pair_quad, tree, point_indices, px, py,
                                  poly_offsets, ring_offsets, vx, vy):
        keep, targs = self._tree_args(tree)
        po = poly_offsets.to(torch.uint32).contiguous()
        ro = ring_offsets.to(torch.uint32).contiguous()
        out = (C.c_void_p * 2)()
        out_n = (C.c_uint64 * 1)()
        torch.cuda.synchronize(px.device)
        t0 = time.perf_counter()
        rc = self._lib.ref_quadtree_point_in_polygon(
            _p(pair_poly), _p(pair_quad), C.c_uint64(pair_poly.numel()), *targs,
            _p(point_indices), _p(px), _p(py), _dt(px), C.c_uint64(px.numel()), _p(po),
            C.c_uint64(po.numel()), _p(ro), C.c_uint64(ro.numel()), _p(vx), _p(vy),
            C.c_uint64(vx.numel()), out, out_n)
        dt = time.perf_counter() - t0
        self._check(rc)
        h = out_n[0]
        d = px.device
        return (self._take(out[0], h, torch.uint32, d), self._take(out[1], h, torch.uint32, d)), dt

    def point_in_polygon(self, px, py, poly_offsets, ring_offsets, vx, vy):
        po = poly_offsets.to(torch.int32).contiguous()
        ro = ring_offsets.to(torch.int32).contiguous()
        out = torch.zeros(px.numel(), dtype=torch.int32, device=px.device)
        torch.cuda.synchronize(px.device)
        t0 = time.perf_counter()
        rc = self._lib.ref_point_in_polygon(
            _p(px), _p(py), _dt(px), C.c_uint64(px.numel()), _p(po), C.c_uint64(po.numel()),
            _p(ro), C.c_uint64(ro.numel()), _p(vx), _p(vy), C.c_uint64(vx.numel()), _p(out))
        dt = time.perf_counter() - t0
        self._check(rc)
        return out, dt


    def quadtree_point_to_nearest_linestring(self, pair_line, pair_quad, tree, point_indices, px, py,
                                             line_offsets, lx, ly):
        keep, targs = self._tree_args(tree)
        lo = line_offsets.to(torch.uint32).contiguous()
        out = (C.c_void_p * 3)()
        out_n = (C.c_uint64 * 1)()
        torch.cuda.synchronize(px.device)
        t0 = time.perf_counter()
        rc = self._lib.ref_quadtree_point_to_nearest_linestring(
            _p(pair_line), _p(pair_quad), C.c_uint64(pair_line.numel()), *targs,
            _p(point_indices), _p(px), _p(py), _dt(px), C.c_uint64(px.numel()), _p(lo),
            C.c_uint64(lo.numel()), _p(lx), _p(ly), C.c_uint64(lx.numel()), out, out_n)
        dt = time.perf_counter() - t0
        self._check(rc)
        n = out_n[0]
        d = px.device
        return (self._take(out[0], n, torch.uint32, d), self._take(out[1], n, torch.uint32, d),
                self._take(out[2], n, px.dtype, d)), dt

    def linestring_bounding_boxes(self, line_offsets, lx, ly, expansion=0.0):
        import numpy as np

        lo = line_offsets.to(torch.uint32).contiguous()
        n = lo.numel() - 1
        npdt = np.float32 if lx.dtype == torch.float32 else np.float64
        outs = [np.zeros(n, dtype=npdt) for _ in range(4)]  # host outputs in both flavours
        rc = self._lib.ref_linestring_bounding_boxes(
            _p(lo), C.c_uint64(lo.numel()), _p(lx), _p(ly), _dt(lx), C.c_uint64(lx.numel()),
            C.c_double(expansion), *[o.ctypes.data_as(C.c_void_p) for o in outs])
        self._check(rc)
        return tuple(torch.as_tensor(o, device=lx.device) for o in outs)

    def polygon_bounding_boxes(self, poly_offsets, ring_offsets, vx, vy, expansion=0.0):
        import numpy as np

        po = poly_offsets.to(torch.uint32).contiguous()
        ro = ring_offsets.to(torch.uint32).contiguous()
        n = po.numel() - 1
        npdt = np.float32 if vx.dtype == torch.float32 else np.float64
        outs = [np.zeros(n, dtype=npdt) for _ in range(4)]  # host outputs in both flavours
        rc = self._lib.ref_polygon_bounding_boxes(
            _p(po), C.c_uint64(po.numel()), _p(ro), C.c_uint64(ro.numel()), _p(vx), _p(vy),
            _dt(vx), C.c_uint64(vx.numel()), C.c_double(expansion),
            *[o.ctypes.data_as(C.c_void_p) for o in outs])
        self._check(rc)
        return tuple(torch.as_tensor(o, device=vx.device) for o in outs)


_cache = {}


def reference_cuda():
    if "lib" not in _cache:
        _cache["lib"] = CudaRefLib()
    return _cache["lib"]
