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
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's
    baseline legs may touch it -- not the package, not include/, not scripts/."""
    for top in ("cuspatial_b200", "include", "scripts"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                    assert "liboracle" not in txt and "libcuspatial_ref" not in txt, f


def test_cpp_header_compiles_against_the_c_abi(tmp_path):
    """include/cuspatial_b200.hpp: the reference's C++ signatures over raw device spans."""
    import shutil
    import subprocess

    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        pytest.skip("no host C++ compiler")
    src = tmp_path / "use_header.cpp"
    src.write_text(
        '#include "cuspatial_b200.hpp"\n'
        "using namespace cuspatial_b200;\n"
        "int main() {\n"
        "  column_view<double> d; column_view<uint32_t> u; column_view<int32_t> i;\n"
        "  pair_table p; quadtree_table q;\n"
        "  if (d.size) {\n"
        "    auto t = quadtree_on_points<double>(d, d, 0, 1, 0, 1, 1, 15, 512);\n"
        "    auto j = join_quadtree_and_bounding_boxes<double>(t.second, d, d, d, d, 0, 1, 0, 1, 1,"
        " 15);\n"
        "    auto h = quadtree_point_in_polygon<double>(j, t.second, u, d, d, u, u, d, d);\n"
        "    point_in_polygon<double>(d, d, i, i, d, d, nullptr);\n"
        "    pairwise_point_in_polygon<double>(d, d, i, i, d, d, nullptr);\n"
        "    quadtree_point_to_nearest_linestring<double>(p, q, u, d, d, u, d, d, nullptr, nullptr,"
        " nullptr);\n"
        "    linestring_bounding_boxes<double>(u, d, d, 0.0, nullptr, nullptr, nullptr, nullptr);\n"
        "  }\n"
        "  return 0;\n"
        "}\n")
    inc = os.path.join(ROOT, "include")
    r = subprocess.run([cxx, "-std=c++17", "-fsyntax-only", "-I", inc, "-I",
                        "/usr/local/cuda/include", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU reference on the host cores) needs no GPU."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "points/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True
