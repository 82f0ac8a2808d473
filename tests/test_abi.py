"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/b2s.h declares, and fails loudly
(no fallback) when no CUDA device is visible.  No compute calls are made here."""
import ctypes as C
import os
import re

import pytest

from open3d_slam_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "b2s.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    declared = header_functions()
    assert len(declared) >= 50
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, f"declared in include/b2s.h but not exported by libb2s.so: {missing}"
    assert set(L.SYMBOLS) == set(declared)     # the ctypes mirror lists exactly the header's entry points


def test_struct_layouts_match_the_header():
    # sizes the C side computes implicitly: b2s_result is 16 + 2 doubles + 2 int32 = 152 bytes, cropper 64 bytes
    assert C.sizeof(L.Result) == 152
    assert C.sizeof(L.Cropper) == 64
    assert C.sizeof(L.IcpParams) == 48
    assert C.sizeof(L.Config) == C.sizeof(L.IcpParams) + C.sizeof(L.ScanParams) + 32   # 3 doubles + icp_cluster_ctas + reserved_


def test_default_config_is_the_lua_default():
    cfg = L.Config()
    L.lib().b2s_default_config(C.byref(cfg))
    assert cfg.icp.reg_type == L.REG_POINT_TO_PLANE and cfg.icp.max_iter == 50 and cfg.icp.knn == 20
    assert cfg.icp.max_corr_dist == 1.0 and cfg.icp.knn_radius == 3.0 and cfg.icp.rel_fitness == 1e-6
    assert cfg.scan.voxel_size == 0.1 and cfg.scan.downsampling_ratio == 0.3 and cfg.map_voxel_size == 0.1
    for c in (cfg.scan.map_builder_cropper, cfg.scan.scan_matcher_cropper):
        assert c.kind == L.CROP_MINMAX_RADIUS and c.rmin == 2.0 and c.rmax == 30.0


def test_no_gpu_means_loud_failure_not_fallback():
    lib = L.lib()
    if lib.b2s_device_count() > 0:
        pytest.skip("a GPU is visible: covered by the -m gpu tests")
    h = C.c_void_p()
    rc = lib.b2s_create(None, C.c_int32(0), None, C.byref(h))
    assert rc == L.E_CUDA and not h.value
    assert b"no CPU fallback" in lib.b2s_last_error() or b"CUDA" in lib.b2s_last_error()
    from open3d_slam_b200 import engine as E
    with pytest.raises(L.B2SError):
        E.Engine()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under open3d_slam_b200/ may import, load or link it."""
    pkg = os.path.join(ROOT, "open3d_slam_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "o3d_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(dirpath, f)
