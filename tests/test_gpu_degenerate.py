"""Degenerate clouds through the whole dense path (set_surface -> normals -> FPFH -> SHOT -> matching): no crash, no
hang, PCL's NaN conventions, and agreement with the oracle where the answer is well defined."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_dense(ctx, pts, k=8, r=0.05):
    ctx.set_viewpoint(0, 0, 0)
    ctx.set_surface(pts)
    ctx.set_queries(None)
    nr = ctx.normals(k=k)
    f = ctx.fpfh(k=k)
    s, rf = ctx.shot352(r)
    return nr, f, s, rf


def test_tiny_clouds(ctx, orc):
    for n in (0, 1, 2, 3, 7):
        pts = np.random.default_rng(n).uniform(0, 0.05, (n, 3)).astype(np.float32)
        nr, f, s, rf = run_dense(ctx, pts)
        assert nr.shape == (n, 4) and f.shape == (n, 33) and s.shape == (n, 352)
        if n:
            onr, _, _ = orc.normals(pts, k=8)
            assert np.array_equal(np.isnan(nr[:, 0]), np.isnan(onr[:, 0]))
            of = orc.fpfh(pts, nr, k=8)
            assert np.array_equal(np.isnan(f), np.isnan(of))
            os_, orf = orc.shot352(pts, nr, None, 0.05)
            assert np.array_equal(np.isnan(s[:, 0]), np.isnan(os_[:, 0]))
            idx, d2 = ctx.knn(8)
            oi, od = orc.knn(pts, pts, 8)
            assert np.array_equal(idx, oi) and np.array_equal(d2, od)


def test_all_points_identical(ctx, orc):
    pts = np.full((3000, 3), 0.25, np.float32)
    nr, f, s, rf = run_dense(ctx, pts, k=8, r=0.01)
    idx, d2 = ctx.knn(8)
    oi, od = orc.knn(pts, pts, 8)
    assert np.array_equal(idx, oi) and np.all(d2 == 0)          # all ties: ascending index decides
    assert np.isnan(s).all()                                    # SHOT: no neighbour differs from the query -> NaN frame
    onr, _, _ = orc.normals(pts, k=8)
    assert np.array_equal(np.isnan(nr[:, 0]), np.isnan(onr[:, 0]))


def test_collinear_and_exact_grid(ctx, orc):
    line = np.c_[np.arange(2000) * 0.001, np.zeros(2000), np.zeros(2000)].astype(np.float32)
    nr, f, s, rf = run_dense(ctx, line, k=8, r=0.01)
    idx, d2 = ctx.knn(8)
    oi, od = orc.knn(line, line, 8)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)   # symmetric ties on a line
    assert not np.isnan(f).any() or np.isnan(nr[:, 0]).any()
    # an exact lattice: every k-th distance is tied many times over
    g = np.stack(np.meshgrid(np.arange(40), np.arange(40), np.arange(3)), -1).reshape(-1, 3).astype(np.float32) * 0.01
    ctx.set_surface(g)
    ctx.set_queries(None)
    for k in (1, 7, 27, 32):
        idx, d2 = ctx.knn(k)
        oi, od = orc.knn(g, g, k)
        assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    off, ridx, rd2 = ctx.radius_search(0.0101)
    ooff, oidx, od2 = orc.radius_search(g, g, 0.0101)
    assert np.array_equal(off, ooff) and np.array_equal(ridx, oidx) and np.array_equal(rd2, od2)
    nr, f, s, rf = run_dense(ctx, g, k=16, r=0.025)
    onr, _, gap = orc.normals(g, k=16)
    assert np.array_equal(np.isnan(nr[:, 0]), np.isnan(onr[:, 0]))
    ok = gap > 1e-3
    assert np.abs(np.abs(np.sum(nr[ok, :3] * onr[ok, :3], 1)) - 1).max() < 1e-5


def test_nan_and_far_outliers(ctx, orc):
    rng = np.random.default_rng(4)
    pts = rng.uniform(0, 0.3, (6000, 3)).astype(np.float32)
    pts[::17] = np.nan
    pts[5] = [1e6, -1e6, 1e6]          # outliers stretch the bounding box by seven orders of magnitude
    pts[6] = [-1e6, 1e6, -1e6]
    nr, f, s, rf = run_dense(ctx, pts, k=12, r=0.04)
    bad = np.isnan(pts[:, 0])
    assert np.isnan(nr[bad]).all() and np.isnan(f[bad]).all() and np.isnan(s[bad]).all()
    idx, d2 = ctx.knn(12)
    oi, od = orc.knn(pts, pts, 12)
    assert np.array_equal(idx, oi) and np.array_equal(d2[~bad], od[~bad])
    # matching with NaN rows and duplicates on both sides
    a = f[:500].copy()
    b = f[500:1500].copy()
    b[10] = a[20]
    corr = ctx.match(a, b, reciprocal=False)
    oi2, od2 = orc.match_nn(a, b)
    keep = oi2 >= 0
    assert np.array_equal(corr["index_match"], oi2[keep]) and np.array_equal(corr["distance"], od2[keep])


def test_two_contexts_side_by_side(ctx):
    """contexts own all of their state (grids, caches, shared-memory attributes, streams): interleaved use of two
    contexts on one device gives the results of using either alone"""
    import pcl_feature_extraction_b200 as pfx
    rng = np.random.default_rng(8)
    a = rng.uniform(0, 0.4, (9000, 3)).astype(np.float32)
    b = rng.uniform(0, 0.3, (7000, 3)).astype(np.float32)
    ref_a = run_dense(ctx, a, k=10, r=0.03)
    ref_b = run_dense(ctx, b, k=10, r=0.03)
    other = pfx.Context(0)
    try:
        other.set_viewpoint(0, 0, 0)
        ctx.set_surface(a)
        other.set_surface(b)
        ctx.prepare_radius(0.03)
        other.prepare_radius(0.03)
        na, nb_ = ctx.normals(k=10), other.normals(k=10)
        fa, fb = ctx.fpfh(k=10), other.fpfh(k=10)
        sa, sb = ctx.shot352(0.03), other.shot352(0.03)
        for got, ref in ((na, ref_a[0]), (fa, ref_a[1]), (sa[0], ref_a[2]), (sa[1], ref_a[3]),
                         (nb_, ref_b[0]), (fb, ref_b[1]), (sb[0], ref_b[2]), (sb[1], ref_b[3])):
            assert np.array_equal(got, ref, equal_nan=True)
        m1 = ctx.match(fa[:300], fa[300:900])
        m2 = other.match(fa[:300], fa[300:900])
        assert np.array_equal(m1, m2)
    finally:
        other.close()


@pytest.mark.gpu
def test_host_input_reuse_keeps_grids_and_normals(ctx, orc, clouds):
    """the reference announces the same cloud (and hands back the same normals) for every descriptor type
    (features.h:186-193): an unchanged PFX_HOST cloud is not uploaded again, its normals are not recomputed, and
    the results are the bits of a fresh run; an edit in place is noticed"""
    import pcl_feature_extraction_b200 as pfx
    pts = np.ascontiguousarray(clouds["underwater_source"][:30000]).copy()
    q = np.ascontiguousarray(pts[::50])
    ctx.set_reuse(True)
    ctx.set_viewpoint(0, 0, 0)
    ctx.set_surface(pts)
    i0 = ctx.reuse_info()
    nbuf = np.zeros((len(pts), 4), np.float32)

    def normals_into(buf):
        ctx._chk(ctx.lib.pfx_normals(ctx.h, 0.03, 0, pfx.capi._ptr(buf), 16, 3, pfx.capi.HOST))

    normals_into(nbuf)
    ctx.set_surface_normals(nbuf)          # the buffer pfx_normals just filled: no upload
    ctx.set_queries(q)
    f1 = ctx.fpfh(radius=0.05)
    l1 = ctx.launches
    # the reference's next descriptor type: same cloud, normals recomputed, same keypoints
    ctx.set_surface(pts)
    with pytest.raises(pfx.PfxError):       # a fresh surface has no input normals yet (FeatureFromNormals::initCompute)
        ctx.set_queries(q)
        ctx.fpfh(radius=0.05)
    ctx.set_queries(None)
    nbuf2 = np.zeros_like(nbuf)
    normals_into(nbuf2)                     # answered from the resident normals
    assert np.array_equal(nbuf2.view(np.uint32), nbuf.view(np.uint32))
    ctx.set_surface_normals(nbuf2)
    ctx.set_queries(q)
    l2 = ctx.launches
    f2 = ctx.fpfh(radius=0.05)
    assert np.array_equal(f1.view(np.uint32), f2.view(np.uint32))
    i1 = ctx.reuse_info()
    assert i1["surface_uploads"] - i0["surface_uploads"] == 0 and i1["surface_reused"] - i0["surface_reused"] == 1
    assert i1["normals_passes"] - i0["normals_passes"] == 1 and i1["normals_reused"] - i0["normals_reused"] == 1
    assert i1["normals_upload_skipped"] - i0["normals_upload_skipped"] == 2
    # the second FPFH found its radius grid: it launched far fewer kernels than the first descriptor pass did
    assert ctx.launches - l2 < 12
    # an edit in place is a different cloud
    pts[123] += np.float32(0.5)
    ctx.set_surface(pts)
    i2 = ctx.reuse_info()
    assert i2["surface_uploads"] - i1["surface_uploads"] == 1
    normals_into(nbuf2)
    ref, _, gap = orc.normals(pts, radius=0.03)
    ok = gap > 1e-3
    assert np.abs(nbuf2[ok, :3] - ref[ok, :3]).max() < 1e-4        # recomputed for the edited cloud
    assert ctx.reuse_info()["normals_passes"] - i2["normals_passes"] == 1
    # switched off: every announcement uploads
    ctx.set_reuse(False)
    ctx.set_surface(pts)
    ctx.set_surface(pts)
    assert ctx.reuse_info()["surface_uploads"] - i2["surface_uploads"] == 2
    ctx.set_reuse(True)
    ctx.set_queries(None)


def test_long_strip_with_nan_points_beyond_256_cells_per_axis(ctx, orc):
    """A 16 m long strip: the k-search grid has ~800 cells along x, so the Morton keys use more than 24 bits and the top
    pass of the radix sort has real work to do (on compact clouds it only copies); non-finite points carry the invalid
    key and must end up behind every finite one.  Sampled k-search rows and the points' normals against the oracle."""
    from pcl_feature_extraction_b200.synth import sheet_cloud
    full = sheet_cloud(side=1024, pitch=0.016, seed=7)           # 16.4 m x 16.4 m, generation order shuffled
    pts = np.ascontiguousarray(full[full[:, 1] < 0.016 * 64])    # a strip 64 rows wide
    assert 60000 < len(pts) < 70000
    rng = np.random.default_rng(1)
    bad = rng.choice(len(pts), 50, replace=False)
    pts[bad[:25], 0] = np.nan
    pts[bad[25:], 2] = np.inf
    ctx.set_viewpoint(0, 0, 0)
    ctx.set_surface(pts)
    ctx.set_queries(None)
    nr = ctx.normals(k=16)
    info = ctx.grid_info()
    assert max(info["dims"]) > 256, info
    idx, d2 = ctx.knn(16)
    rows = rng.choice(len(pts), 3000, replace=False)
    rows = rows[np.isfinite(pts[rows]).all(1)]
    finite = np.isfinite(pts).all(1)
    oi, od = orc.knn(pts[finite], pts[rows], 16)
    back = np.flatnonzero(finite)
    assert np.array_equal(back[oi], idx[rows]) and np.array_equal(od, d2[rows])
    assert np.isnan(nr[bad, 0]).all() and not np.isnan(nr[finite, 0]).any()
    f = ctx.fpfh(k=16)
    blocks = f[finite].reshape(-1, 3, 11).sum(2)
    assert np.abs(blocks - 100).max() < 1e-2
