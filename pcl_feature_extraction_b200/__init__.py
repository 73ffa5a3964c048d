"""pcl_feature_extraction_b200 — B200-native (sm_100a) feature-extraction hot path.

Product = the CUDA library behind the C ABI in include/pfx_b200.h (lib/libpfx_b200.so, built from
csrc/).  This package is only the Python binding used by tests and bench.py, the PCD reader and the
synthetic cloud generator.  There is no CPU fallback: `Context()` raises without an sm_100 GPU and
`capi.load()` raises when the library has not been built.
"""
from . import capi, pcd, synth  # noqa: F401
from .capi import Context, PfxError  # noqa: F401
