"""Dense SHOT352 from the resident k-search rows (shot_fused.cu, ROWS variant) against the stencil walk on a radius
grid: the same neighbourhoods, the same arithmetic, hence the same bits; rows the k-search cannot close take the
stencil pass on the k-search grid, and a radius beyond the grid's cells the generic kernels with m-ring stencils
(reference call site: SHOTEstimationOMP through features.h:181-195, evaluation.cpp:770-775)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIDE = 320  # 102 400 points: above the size from which the dense path takes the rows


@pytest.fixture(scope="module")
def sheet():
    from pcl_feature_extraction_b200.synth import sheet_cloud
    return sheet_cloud(side=SIDE, pitch=0.004)


def _context(rows):
    import pcl_feature_extraction_b200 as pfx
    old = os.environ.get("PFX_SHOT_ROWS")
    os.environ["PFX_SHOT_ROWS"] = "1" if rows else "0"  # read once, by pfx_create
    try:
        c = pfx.Context(0)
    finally:
        if old is None:
            del os.environ["PFX_SHOT_ROWS"]
        else:
            os.environ["PFX_SHOT_ROWS"] = old
    c.set_viewpoint(0, 0, 0)
    return c


def _same_bits(a, b):
    return ((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all(axis=1)


@pytest.mark.parametrize("k,radius", [(32, 0.0128), (32, 0.009), (16, 0.008)])
def test_rows_variant_gives_the_bits_of_the_stencil_walk(sheet, orc, k, radius):
    out = {}
    launches = {}
    for rows in (False, True):
        c = _context(rows)
        c.set_surface(sheet)
        c.set_queries(None)
        nr = c.normals(k=k)
        l0 = c.launches
        s, rf = c.shot352(radius)
        launches[rows] = c.launches - l0
        out[rows] = (np.concatenate([s, rf], axis=1), nr)
        c.close()
    assert launches[True] <= 5  # rows pass, open-rows pass, three work-list launches: no index build
    same = _same_bits(out[False][0], out[True][0])
    assert same.all(), (same.mean(), np.nanmax(np.abs(out[False][0] - out[True][0])))
    # and the rows are PCL's: a sample against the oracle, frames given (the frame solve is compared elsewhere)
    s, nr = out[True]
    sel = np.arange(0, len(sheet), 97)
    ref, _ = orc.shot352(sheet, nr, np.ascontiguousarray(sheet[sel]), radius, lrf_in=np.ascontiguousarray(s[sel, 352:361]))
    ok = ~np.isnan(ref[:, 0])
    assert np.array_equal(np.isnan(s[sel, 0]), ~ok)
    assert ok.mean() > 0.9
    assert np.abs(s[sel][ok, :352] - ref[ok]).max() <= 1e-6


def test_radius_beyond_the_k_search_cells(sheet):
    """r = 20 mm holds ~60 points: every k = 32 row stays open and the k-search grid's cells are finer than the
    radius, so the generic kernels walk two rings of cells.  The next call knows (asynchronous read-back of the
    open share) and goes back to a radius grid: then the bits are those of a context that never used the rows."""
    radius = 0.02
    c0 = _context(False)
    c0.set_surface(sheet)
    c0.set_queries(None)
    c0.normals(k=32)
    s0, rf0 = c0.shot352(radius)
    c0.close()
    c1 = _context(True)
    c1.set_surface(sheet)
    c1.set_queries(None)
    c1.normals(k=32)
    s1, rf1 = c1.shot352(radius)   # rows -> all open -> generic kernels on the k-search grid
    s2, rf2 = c1.shot352(radius)   # open share known: radius grid
    c1.close()
    assert np.array_equal(np.isnan(s0[:, 0]), np.isnan(s1[:, 0]))
    ok = ~np.isnan(s0[:, 0])
    assert ok.mean() > 0.99
    # generic (double) against fused (float) accumulation of the same neighbourhoods
    assert np.abs(s1[ok] - s0[ok]).max() <= 2e-6
    assert np.abs(rf1[ok] - rf0[ok]).max() <= 1e-6
    assert _same_bits(np.concatenate([s2, rf2], 1), np.concatenate([s0, rf0], 1)).all()
