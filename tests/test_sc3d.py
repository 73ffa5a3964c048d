"""3D Shape Context (reference evaluation.cpp:319-345, the first entry of its descriptor list):
the CPU oracle against known answers, and the CUDA path (through the C ABI) against the oracle.

Contract: the frame of a query is the normal of its nearest surface point plus a random tangent
direction; upstream's wall-clock-seeded mt19937 cannot be pinned, so the three uniform draws of query i are the top
24 bits of SplitMix64(seed + golden (3 i + t + 1)) on both sides.  Frames are then BIT-identical (same float
operations, no FMA); the bins follow USC's tolerance (test_usc.py): 1e-5 relative given identical frames, at most
2 % of the rows carrying one weight moved across a bin boundary by an ulp of atan2f / acosf."""
import numpy as np
import pytest

from test_usc import bumpy


def test_oracle_sc3d_known_answers(orc):
    pts = bumpy(20000, 1)
    nr, _, _ = orc.normals(pts, k=16)
    q = pts[:80]
    r = 0.08
    out, fr = orc.sc3d1980(pts, nr, q, r, seed=7)
    assert not np.isnan(out).any() and (out >= 0).all()
    # orthonormal right-handed frames whose z axis is the normal of the nearest surface point (the query itself here)
    x, y, z = fr[:, :3], fr[:, 3:6], fr[:, 6:]
    assert np.array_equal(z, nr[:80, :3])
    assert np.abs((x * z).sum(1)).max() < 1e-5 and np.abs((x * x).sum(1) - 1).max() < 1e-6
    assert np.abs(np.cross(z, x) - y).max() < 1e-6
    # the same seed gives the same rows, another seed another tangent direction but the same elevation / radius
    # marginals: summing over the azimuth sectors removes the random direction
    out_b, fr_b = orc.sc3d1980(pts, nr, q, r, seed=7)
    assert np.array_equal(out, out_b)
    out_c, fr_c = orc.sc3d1980(pts, nr, q, r, seed=8)
    assert not np.array_equal(fr_c[:, :3], fr[:, :3])
    m1 = out.reshape(-1, 12, 11 * 15).sum(1)
    m2 = out_c.reshape(-1, 12, 11 * 15).sum(1)
    assert np.abs(m1 - m2).max() <= 1e-4 * m1.max()
    # given the same frames USC computes the same bins
    usc, _, _ = orc.usc1980(pts, q, r, lrf_in=fr)
    assert np.array_equal(usc, out)
    # a query without neighbours: NaN row
    far = np.array([[9, 9, 9]], np.float32)
    o, f = orc.sc3d1980(pts, nr, far, r)
    assert np.isnan(o).all() and np.isnan(f).all()


@pytest.mark.gpu
@pytest.mark.parametrize("n,r,dense,seed", [(20000, 0.08, False, 2), (4000, 0.1, True, 3), (40000, 0.05, False, 4)])
def test_gpu_sc3d_equals_oracle(ctx, orc, n, r, dense, seed):
    pts = bumpy(n, seed)
    nr, _, _ = orc.normals(pts, k=16)
    sel = np.random.default_rng(seed).choice(n, 200, replace=False)
    q = pts if dense else np.concatenate([pts[sel], pts[sel[:20]] + np.float32(0.002)])  # on and off the surface
    ref, fr_ref = orc.sc3d1980(pts, nr, q, r, seed=1000 + seed)
    ctx.set_surface(pts)
    ctx.set_surface_normals(nr)
    ctx.set_queries(None if dense else q)
    g, fr, rf_out = ctx.sc3d1980(r, seed=1000 + seed)
    ctx.set_queries(None)
    assert np.array_equal(fr.view(np.uint32), fr_ref.view(np.uint32)), "3DSC frames must be bit-identical"
    assert np.all(rf_out == 0)                                      # upstream zeroes rf: no repeatable frame
    assert np.array_equal(np.isnan(g[:, 0]), np.isnan(ref[:, 0]))
    ok = ~np.isnan(ref[:, 0])
    assert ok.mean() > 0.9
    tol = 1e-5 * np.abs(ref[ok]) + 1e-6 * ref[ok].max(1, keepdims=True)
    row_ok = np.all(np.abs(g[ok] - ref[ok]) <= tol, axis=1)
    assert row_ok.mean() >= 0.98, row_ok.mean()
    assert np.abs(g[ok].sum(1) - ref[ok].sum(1)).max() <= 1e-4 * ref[ok].sum(1).max()


@pytest.mark.gpu
def test_gpu_sc3d_preconditions_and_nan_rows(ctx, orc):
    import pcl_feature_extraction_b200 as pfx
    pts = bumpy(3000, 5)
    nr, _, _ = orc.normals(pts, k=16)
    ctx.set_surface(pts)
    with pytest.raises(pfx.PfxError) as e:      # FeatureFromNormals::initCompute: normals required
        ctx.sc3d1980(0.08)
    assert e.value.code == pfx.capi.E_STATE
    ctx.set_surface_normals(nr)
    ctx.set_queries(np.array([[9, 9, 9], pts[0]], np.float32))
    g, fr, rf = ctx.sc3d1980(0.08)
    assert np.isnan(g[0]).all() and np.isnan(fr[0]).all() and not np.isnan(g[1]).any() and np.all(rf == 0)
    with pytest.raises(pfx.PfxError) as e:      # search radius below the minimal radius
        ctx.sc3d1980(0.01, min_radius=0.02)
    assert e.value.code == pfx.capi.E_PRECOND
    ctx.set_queries(None)
