"""CPU tests (no GPU): the C-ABI library builds, loads and exports exactly what the header declares;
host-side argument validation mirrors the reference's error conditions."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cuspatial_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bsj_[a-z_0-9]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    from cuspatial_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    L = _lib.lib()
    declared = _declared_symbols()
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s
    assert b"sm_100a" in L.bsj_version()


def test_python_api_mirrors_reference_signatures():
    import inspect

    import cuspatial_b200 as cs

    assert list(inspect.signature(cs.quadtree_on_points).parameters) == [
        "points", "x_min", "x_max", "y_min", "y_max", "scale", "max_depth", "max_size"]
    assert list(inspect.signature(cs.join_quadtree_and_bounding_boxes).parameters) == [
        "quadtree", "bounding_boxes", "x_min", "x_max", "y_min", "y_max", "scale", "max_depth"]
    assert list(inspect.signature(cs.quadtree_point_in_polygon).parameters) == [
        "poly_quad_pairs", "quadtree", "point_indices", "points", "polygons"]
    assert list(inspect.signature(cs.point_in_polygon).parameters) == ["points", "polygons"]


def test_no_cpu_fallback_inputs_must_be_on_device():
    import torch

    import cuspatial_b200 as cs

    x = torch.zeros(4, dtype=torch.float64)
    with pytest.raises(ValueError):
        cs.quadtree_on_points((x, x), 0, 1, 0, 1, 1, 3, 2)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cuspatial_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "liboracle" not in txt and "libcuspatial_ref" not in txt, f
