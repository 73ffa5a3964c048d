"""CPU tests of the C-ABI boundary: the library loads and exports every symbol that include/pfx_b200.h
declares; the ctypes signature table mirrors the header; the product path has no CPU fallback."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "pfx_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pfx_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    names = header_functions()
    for need in ["pfx_set_surface", "pfx_set_queries", "pfx_set_surface_normals", "pfx_normals", "pfx_knn",
                 "pfx_radius_search", "pfx_cloud_resolution", "pfx_iss", "pfx_harris3d", "pfx_fpfh", "pfx_shot352",
                 "pfx_match", "pfx_voxel_grid"]:
        assert need in names


def test_library_exports_every_declared_symbol():
    import pcl_feature_extraction_b200 as pfx
    lib = pfx.capi.load()
    for name in header_functions():
        assert hasattr(lib, name), f"{name} is declared in pfx_b200.h but not exported"
    assert lib.pfx_version() >= 100


def test_ctypes_table_matches_header():
    import pcl_feature_extraction_b200 as pfx
    assert sorted(pfx.capi.SIGNATURES) == header_functions()


def test_no_cpu_fallback_and_no_oracle_in_product():
    import pcl_feature_extraction_b200 as pfx
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(pfx.PfxError):
            pfx.Context(0)
    # the product package must never import / link the oracle
    pkg = os.path.join(ROOT, "pcl_feature_extraction_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                for pat in (r"import\s+oracle", r"from\s+oracle", r"liboracle", r"oracle/", r"pcl_oracle\.h", r"\borc_[a-z]"):
                    assert not re.search(pat, txt), (dirpath, f, pat)


def test_pod_layouts_match_pcl():
    import pcl_feature_extraction_b200 as pfx
    assert ctypes.sizeof(pfx.capi.Correspondence) == 12
    assert pfx.capi.CORR_DTYPE.itemsize == 12
