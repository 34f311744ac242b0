"""The C-ABI shared library loads without a GPU and exports every symbol include/tfusion_b200.h declares;
argument validation works without touching the device.  No compute calls here (CPU box)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tfusion_b200.h")
LIB = os.path.join(ROOT, "topfusion_b200", "libtfusion_b200.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return C.CDLL(LIB)


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"TFB_API\s+[\w\s\*]+?\b(tfb_\w+)\s*\(", src)))


def test_header_declares_the_path():
    names = declared_symbols()
    assert len(names) >= 45
    for must in ("tfb_create", "tfb_process_frame", "tfb_allocate_scene_from_depth", "tfb_integrate_into_scene",
                 "tfb_create_expected_depths", "tfb_create_icp_maps", "tfb_icp_estimate", "tfb_bilateral_filter",
                 "tfb_compute_point_normals", "tfb_resize_points_normals", "tfb_depth_pyr", "tfb_compute_dists"):
        assert must in names


def test_every_declared_symbol_is_exported(lib):
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_every_entry_point_cites_the_reference():
    """each block of the header names the reference file it replaces"""
    src = open(HEADER).read()
    for ref in ("src/topfu.cpp", "src/cuda/imgproc.cu", "src/cuda/proj_icp.cu", "src/projective_icp.cpp",
                "SceneReconstructionEngine_host.cu", "VisualisationEngine_CUDA.cu", "src/internal.hpp"):
        assert ref in src, ref


def test_no_torch_or_cxx_types_in_the_abi():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)   # declarations only, comments stripped
    for bad in ("torch", "at::", "std::", "cv::", "template", "class "):
        assert bad not in src, bad


def test_default_params_match_reference_defaults(lib):
    from topfusion_b200.capi import Params
    p = Params()
    assert lib.tfb_default_params(C.byref(p)) == 0
    assert (p.cols, p.rows) == (640, 480)
    assert abs(p.fx - 504.261) < 1e-3 and abs(p.cy - 272.202) < 1e-3          # topfu.cpp:24
    assert list(p.icp_iters) == [10, 5, 4, 0]                                   # topfu.cpp:14
    assert abs(p.mu - 0.02) < 1e-7 and p.max_w == 100 and abs(p.voxel_size - 0.005) < 1e-7   # topfu.cpp:50
    assert (p.num_blocks, p.num_buckets, p.excess_size) == (0x10000, 0x100000, 0x20000)      # VoxelBlockHash.hpp:14-18
    assert p.depth_cutoff_mm == 2047                                            # imgproc.cu:277


def test_argument_errors_do_not_need_a_device(lib):
    assert lib.tfb_default_params(None) == -1
    assert lib.tfb_create(None, None, None) == -1
    assert lib.tfb_reset(None) == -1
    assert lib.tfb_process_frame(None, None, C.c_size_t(0), None) == -1
    lib.tfb_version.restype = C.c_char_p
    assert b"sm_100a" in lib.tfb_version()


def test_create_fails_loudly_without_gpu(lib):
    """no CPU fallback: on a box without a CUDA device tfb_create must return an error, never a working context"""
    from conftest import has_cuda
    if has_cuda():
        pytest.skip("a GPU is present")
    from topfusion_b200.capi import Params
    p = Params()
    lib.tfb_default_params(C.byref(p))
    h = C.c_void_p()
    lib.tfb_create.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    rc = lib.tfb_create(C.byref(p), None, C.byref(h))
    assert rc != 0 and not h.value


def test_product_does_not_touch_the_oracle():
    """the oracle is test infrastructure: nothing under topfusion_b200/, src/, include/ or apps/ may reference it"""
    bad = []
    for sub in ("topfusion_b200", "src", "include", "apps"):
        for dp, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    # code references only: includes, imports, load/link names, calls (prose may mention the oracle)
                    if re.search(r"#\s*include[^\n]*(oracle|tfo_)|\bimport\s+oracle|\bfrom\s+oracle|libtfo|\btfo_\w+\s*\(|-ltfo", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
