"""ctypes loader of libcuspatial_b200.so (the C ABI declared in include/cuspatial_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or a call fails, the error is
raised.  Nothing here (or anywhere in this package) imports oracle/.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BSJ_LIBRARY_PATH") or os.path.join(_HERE, "libcuspatial_b200.so")

BSJ_SUCCESS, BSJ_INVALID_ARGUMENT, BSJ_CUDA_ERROR, BSJ_OUT_OF_MEMORY = 0, 1, 2, 3

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p)
FREE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p)


class bsj_allocator(C.Structure):
    _fields_ = [("allocate", ALLOC_FN), ("deallocate", FREE_FN), ("ctx", C.c_void_p)]


class bsj_grid(C.Structure):
    _fields_ = [
        ("valid", C.c_int32), ("max_depth", C.c_int32),
        ("min_x", C.c_double), ("min_y", C.c_double), ("max_x", C.c_double), ("max_y", C.c_double),
        ("scale", C.c_double), ("has_nan", C.c_int32), ("has_out_of_bbox", C.c_int32),
        ("sorted_keys", C.c_void_p), ("n_sorted_keys", C.c_uint64),
    ]


class bsj_quadtree(C.Structure):
    _fields_ = [
        ("point_indices", C.c_void_p), ("num_points", C.c_uint64),
        ("key", C.c_void_p), ("level", C.c_void_p), ("is_internal_node", C.c_void_p),
        ("length", C.c_void_p), ("offset", C.c_void_p), ("num_nodes", C.c_uint64),
        ("grid", bsj_grid), ("sorted_keys", C.c_void_p),
    ]


class bsj_pip_compact(C.Structure):
    _fields_ = [
        ("pair_offset", C.c_void_p), ("pair_length", C.c_void_p), ("pair_hits", C.c_void_p),
        ("pair_class", C.c_void_p), ("pair_word_base", C.c_void_p), ("pair_row_base", C.c_void_p),
        ("mask_words", C.c_void_p), ("n_pairs", C.c_uint64), ("n_words", C.c_uint64),
        ("n_hits", C.c_uint64),
    ]


BSJ_MAX_RANKS = 32


class bsj_shard_plan(C.Structure):
    """Mirror of the device-resident sharding plan (include/cuspatial_b200.h)."""
    _fields_ = [
        ("n_ranks", C.c_uint32), ("rank", C.c_uint32), ("hist_shift", C.c_uint32),
        ("sub_shift", C.c_uint32), ("n_sub", C.c_uint32), ("n_targets", C.c_uint32),
        ("status", C.c_uint32), ("reserved", C.c_uint32),
        ("gid_base", C.c_uint32 * (BSJ_MAX_RANKS + 1)),
        ("bound_bin", C.c_uint32 * BSJ_MAX_RANKS), ("bound_missing", C.c_uint32 * BSJ_MAX_RANKS),
        ("target_bin", C.c_uint32 * BSJ_MAX_RANKS), ("splitter", C.c_uint32 * BSJ_MAX_RANKS),
        ("send_count", C.c_uint32 * BSJ_MAX_RANKS), ("send_offset", C.c_uint32 * BSJ_MAX_RANKS),
        ("recv_total", C.c_uint32 * BSJ_MAX_RANKS),
    ]


class bsj_coord_segments(C.Structure):
    _fields_ = [
        ("n_segments", C.c_int32), ("first_id", C.c_uint32 * (BSJ_MAX_RANKS + 1)),
        ("x", C.c_void_p * BSJ_MAX_RANKS), ("y", C.c_void_p * BSJ_MAX_RANKS),
    ]


class bsj_pairs(C.Structure):
    _fields_ = [("first", C.c_void_p), ("second", C.c_void_p), ("size", C.c_uint64)]


# every symbol include/cuspatial_b200.h declares
EXPORTED_SYMBOLS = [
    "bsj_quadtree_on_points", "bsj_join_quadtree_and_bounding_boxes",
    "bsj_quadtree_point_in_polygon", "bsj_quadtree_point_in_polygon_ex", "bsj_quadtree_point_in_polygon_compact",
    "bsj_expand_pip_compact", "bsj_point_in_polygon", "bsj_pairwise_point_in_polygon",
    "bsj_polygon_bounding_boxes", "bsj_quadtree_point_to_nearest_linestring",
    "bsj_linestring_bounding_boxes",
    "bsj_point_keys_histogram", "bsj_shard_plan_level1", "bsj_shard_subhistogram",
    "bsj_shard_plan_level2", "bsj_shard_plan_finalize", "bsj_partition_keys",
    "bsj_quadtree_on_keys", "bsj_quadtree_point_in_polygon_compact_seg", "bsj_free", "bsj_free_quadtree", "bsj_free_pairs", "bsj_last_error", "bsj_version",
    "bsj_kernel_launch_count", "bsj_set_profiling", "bsj_get_profile",
]

_lib = None


def lib():
    """Load the CUDA library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "cuspatial_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C cuspatial_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64, dbl, i32 = C.c_void_p, C.c_uint64, C.c_double, C.c_int32
    L.bsj_quadtree_on_points.argtypes = [vp, vp, C.c_int, u64, dbl, dbl, dbl, dbl, dbl, C.c_int8,
                                         i32, C.POINTER(bsj_allocator), vp,
                                         C.POINTER(bsj_quadtree)]
    L.bsj_join_quadtree_and_bounding_boxes.argtypes = [vp, vp, vp, vp, vp, u64, vp, vp, vp, vp,
                                                       C.c_int, u64, dbl, dbl, dbl, dbl, dbl,
                                                       C.c_int8, C.POINTER(bsj_allocator), vp,
                                                       C.POINTER(bsj_pairs)]
    L.bsj_quadtree_point_in_polygon.argtypes = [vp, vp, u64, vp, vp, vp, vp, vp, u64, vp, vp, vp,
                                                C.c_int, u64, vp, u64, vp, u64, vp, vp, u64,
                                                C.POINTER(bsj_allocator), vp, C.POINTER(bsj_pairs)]
    L.bsj_quadtree_point_in_polygon_ex.argtypes = [vp, vp, u64, vp, vp, vp, vp, vp, u64, vp, vp,
                                                   vp, C.c_int, u64, vp, u64, vp, u64, vp, vp,
                                                   u64, C.POINTER(bsj_grid),
                                                   C.POINTER(bsj_allocator), vp,
                                                   C.POINTER(bsj_pairs)]
    L.bsj_quadtree_point_in_polygon_compact.argtypes = [
        vp, vp, u64, vp, vp, vp, vp, vp, u64, vp, vp, vp, C.c_int, u64, vp, u64, vp, u64, vp, vp,
        u64, C.POINTER(bsj_grid), C.POINTER(bsj_allocator), vp, C.POINTER(bsj_pip_compact)]
    L.bsj_expand_pip_compact.argtypes = [vp, C.POINTER(bsj_pip_compact), C.c_uint32, vp, vp, vp]
    L.bsj_point_in_polygon.argtypes = [vp, vp, C.c_int, u64, vp, u64, vp, u64, vp, vp, u64, vp, vp]
    L.bsj_pairwise_point_in_polygon.argtypes = [vp, vp, C.c_int, u64, vp, u64, vp, u64, vp, vp, u64,
                                                vp, vp]
    L.bsj_quadtree_point_to_nearest_linestring.argtypes = [
        vp, vp, u64, vp, vp, vp, vp, vp, u64, vp, vp, vp, C.c_int, u64, vp, u64, vp, vp, u64, vp,
        vp, vp, vp, C.POINTER(u64)]
    L.bsj_linestring_bounding_boxes.argtypes = [vp, u64, vp, vp, C.c_int, u64, dbl, vp, vp, vp, vp,
                                                vp]
    L.bsj_polygon_bounding_boxes.argtypes = [vp, u64, vp, u64, vp, vp, C.c_int, u64, dbl, vp,
                                             vp, vp, vp, vp]
    L.bsj_point_keys_histogram.argtypes = [vp, vp, C.c_int, u64, dbl, dbl, dbl, dbl, dbl, C.c_int8,
                                           C.c_int, vp, vp, u64, vp, vp]
    L.bsj_shard_plan_level1.argtypes = [vp, u64, vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_uint32, vp, vp]
    L.bsj_shard_subhistogram.argtypes = [vp, u64, vp, C.c_int, C.c_uint32, vp, vp]
    L.bsj_shard_plan_level2.argtypes = [vp, u64, vp, vp, vp, vp]
    L.bsj_shard_plan_finalize.argtypes = [vp, u64, vp, vp]
    L.bsj_partition_keys.argtypes = [vp, u64, vp, C.c_int, vp, vp, C.c_int, vp]
    L.bsj_quadtree_on_keys.argtypes = [vp, vp, u64, C.POINTER(bsj_grid), i32,
                                       C.POINTER(bsj_allocator), vp, C.POINTER(bsj_quadtree)]
    L.bsj_quadtree_point_in_polygon_compact_seg.argtypes = [
        vp, vp, u64, vp, vp, vp, vp, vp, u64, vp, C.POINTER(bsj_coord_segments), C.c_int, u64, vp,
        u64, vp, u64, vp, vp, u64, C.POINTER(bsj_grid), C.POINTER(bsj_allocator), vp,
        C.POINTER(bsj_pip_compact)]
    L.bsj_free.argtypes = [vp, vp]
    L.bsj_last_error.restype = C.c_char_p
    L.bsj_version.restype = C.c_char_p
    L.bsj_kernel_launch_count.restype = u64
    L.bsj_set_profiling.argtypes = [C.c_int]
    L.bsj_get_profile.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    L.bsj_get_profile.restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc == BSJ_SUCCESS:
        return
    msg = lib().bsj_last_error().decode(errors="replace")
    if rc == BSJ_INVALID_ARGUMENT:
        # the reference's cuspatial::logic_error surfaces in Python as RuntimeError (Cython `except +`)
        raise RuntimeError(msg)
    if rc == BSJ_OUT_OF_MEMORY:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def kernel_launch_count():
    return int(lib().bsj_kernel_launch_count())


def set_profiling(on):
    lib().bsj_set_profiling(1 if on else 0)


def get_profile():
    cap = 8192
    names = (C.c_char_p * cap)()
    ms = (C.c_float * cap)()
    n = lib().bsj_get_profile(names, ms, cap)
    return [(names[i].decode(), float(ms[i])) for i in range(n)]
