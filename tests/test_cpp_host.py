"""The C++ host layer (include/cuspatial_b200.hpp over the C ABI): a C++ program written like the
reference's own gtest cases is built here (CPU) and run on the GPU box against the golden vectors."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "golden_join.cpp")


def _vec(name, ctype, values, fmt):
    return "static const std::vector<%s> %s = {%s};\n" % (ctype, name, ", ".join(fmt(v) for v in values))


def _write_golden_inc(path, golden):
    sj, nl = golden["small_join"], golden["nearest_linestring"]
    tree = [c for c in golden["quadtree_cases"] if len(c.get("points", [])) == 71]
    d17 = lambda v: repr(float(v))          # noqa: E731  shortest round-trip decimal of the double
    f9 = lambda v: "%sf" % repr(float(np.float32(v)))   # noqa: E731
    u = lambda v: "%uu" % int(v)            # noqa: E731
    with open(path, "w") as f:
        f.write("// generated from tests/golden/cuspatial_golden.json by tests/test_cpp_host.py\n")
        f.write("#include <cstdint>\n#include <vector>\n")
        f.write(_vec("kPoints", "double", np.array(sj["points"]).reshape(-1), d17))
        f.write(_vec("kVertices", "double", np.array(sj["vertices"]).reshape(-1), d17))
        f.write(_vec("kRingOffsets", "uint32_t", sj["ring_offsets"], u))
        f.write(_vec("kPartOffsets", "uint32_t", sj["part_offsets"], u))
        f.write(_vec("kPairPoly", "uint32_t", sj["pair_poly"], u))
        f.write(_vec("kPairQuad", "uint32_t", sj["pair_quad"], u))
        f.write(_vec("kPipPoly", "uint32_t", sj["pip_poly"], u))
        f.write(_vec("kPipPoint", "uint32_t", sj["pip_point"], u))
        f.write(_vec("kLineVertices", "double", np.array(nl["vertices"]).reshape(-1), d17))
        f.write(_vec("kLineOffsets", "uint32_t", nl["line_offsets"], u))
        f.write(_vec("kLinePairLine", "uint32_t", nl["pair_line"], u))
        f.write(_vec("kLinePairQuad", "uint32_t", nl["pair_quad"], u))
        f.write(_vec("kNearestPoint", "uint32_t", nl["point_index"], u))
        f.write(_vec("kNearestLine", "uint32_t", nl["linestring_index"], u))
        f.write(_vec("kNearestDistanceF32", "float", nl["distance_f32"], f9))
        f.write(_vec("kNearestDistanceF64", "double", nl["distance_f64"], d17))
        assert tree, "the 71-point quadtree case is missing from the golden file"
        t = tree[0]
        assert (t["scale"], t["max_depth"], t["max_size"]) == (1.0, 3, 12) and \
            t["points"] == sj["points"]
        f.write(_vec("kTreeKey", "uint32_t", t["key"], u))
        f.write(_vec("kTreeLength", "uint32_t", t["length"], u))
        f.write(_vec("kTreeOffset", "uint32_t", t["offset"], u))


def _build(tmp_path, golden):
    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        pytest.skip("no host C++ compiler")
    _write_golden_inc(str(tmp_path / "golden_data.inc"), golden)
    exe = str(tmp_path / "golden_join")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = [cxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", str(tmp_path),
           "-I", os.path.join(cuda, "include"), SRC, "-L", os.path.join(ROOT, "cuspatial_b200"),
           "-lcuspatial_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
           "-Wl,-rpath," + os.path.join(ROOT, "cuspatial_b200"),
           "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_cpp_host_program_builds_and_links_against_the_library(tmp_path, golden):
    _build(tmp_path, golden)


@pytest.mark.gpu
def test_cpp_host_program_reproduces_the_golden_vectors(tmp_path, golden):
    exe = _build(tmp_path, golden)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "CPP_GOLDEN OK" in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]
